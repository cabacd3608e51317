set -x
timeout 1800 python tools/ab.py --tag r2g_hdl64 --repeats 2 p2d6: p3d4:p3d4 p4d3:p4d3 p6d2:p6d2 p5d2:p5d2 p3d4s25:p3d4s25 p3d4s23:p3d4s23 2>&1 | tee gpurun_out/r2g_ab_hdl64.txt
timeout 1200 python tools/ab.py --tag r2g_hdl32 --repeats 2 --args "--shape hdl32 --scans 4096" p2d6: p3d4:p3d4 p4d3:p4d3 p6d2:p6d2 p5d2:p5d2 p3d4s25:p3d4s25 p3d4s23:p3d4s23 2>&1 | tee gpurun_out/r2g_ab_hdl32.txt
timeout 600 python tools/ab.py --tag r2g_b128 --repeats 1 --args "--shape beam128 --scans 2048" p2d6: p3d4:p3d4 p4d3:p4d3 2>&1 | tee gpurun_out/r2g_ab_b128.txt

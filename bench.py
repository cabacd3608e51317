#!/usr/bin/env python
"""Benchmark of the spectral encoding front end (BASELINE.json metric: scans/s encoded to 800-D).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the fused encode kernel over this rank's batch of synthetic scans
(BASELINE.json configs[1]: 4541 HDL-64-shaped scans of ~120 k xyzi points, 8.7 GB, resident in
HBM before timing), followed for N > 1 by the gather of the 800-D descriptors into the database
replicated on every GPU. Weak scaling: every rank holds its own 4541 scans.

Prints ONE JSON line (rank 0). ``value`` is device-timed whole-job throughput; ``e2e`` is the
same metric through ``SpectralEncoder.encode_scans`` with pinned HOST buffers (H2D + kernel +
D2H inside the timed region); ``roofline`` is the fused kernel against the measured HBM peak;
``cpu_baseline`` is the CPU oracle (a port of the reference encoder) on the host cores.

``--impl reference`` times the reference's CPU algorithm (the oracle port: the reference is
Python and /root/reference does not exist on the GPU box) with all host cores on the same
workload, a bounded sample per step.
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "scans_per_sec_encoded_to_800d"
UNIT = "scans/s"
N_SCANS = 4541            # KITTI sequence 00 length (BASELINE.json configs[1])
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only without MEASURED_PEAKS.json


SHAPE_DESC = {"hdl64": "HDL-64-shaped scans x ~120k xyzi points", "hdl32": "NCLT HDL-32-shaped scans x ~70k xyzi points",
              "beam128": "128-beam dense scans x ~260k xyzi points"}


def workload_name(n, shape="hdl64", shuffle=False):
    head = "synthetic KITTI seq-00-length batch" if shape == "hdl64" else "synthetic batch"
    return f"{head}: {n} {SHAPE_DESC[shape]} per GPU" + (" (shuffled point order)" if shuffle else "")


# ----------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """Encode ``count`` scans starting at ``first`` with the oracle; returns (count, seconds)."""
    first, count = args
    import torch
    torch.set_num_threads(1)
    from neural_spectral_codec_b200 import synth
    from oracle import nsc_oracle as orc
    cfg = orc.OracleConfig()
    scans = [synth.make_scan(synth.HDL64, first + i).numpy() for i in range(count)]
    t0 = time.perf_counter()
    for s in scans:
        orc.encode_points(s, cfg)
    return count, time.perf_counter() - t0


def cpu_oracle_throughput(scans_per_core: int, cores: int, pool=None):
    """All-cores throughput of the oracle: each worker generates its own scans from seeds (not
    timed) and encodes them single-threaded; throughput = sum over workers of count / time."""
    jobs = [(100000 + w * scans_per_core, scans_per_core) for w in range(cores)]
    if pool is not None:
        res = pool.map(_cpu_worker, jobs, chunksize=1)
    else:
        with mp.get_context("spawn").Pool(cores) as p:
            res = p.map(_cpu_worker, jobs, chunksize=1)
    return sum(c / t for c, t in res), sum(c for c, _ in res)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = 16
    steps = max(1, min(args.steps, 5))
    warmup = min(args.warmup, 1)
    vals, ms = [], []
    with mp.get_context("spawn").Pool(cores) as pool:
        for _ in range(warmup):
            cpu_oracle_throughput(1, cores, pool)
        for _ in range(steps):
            t0 = time.perf_counter()
            v, n = cpu_oracle_throughput(per_core, cores, pool)
            ms.append(1e3 * (time.perf_counter() - t0))
            vals.append(v)
    value = statistics.median(vals)
    sample = (f"{per_core} scans per core x {cores} cores per step, {steps} steps (of {args.steps} asked); "
              "oracle/nsc_oracle.py (port of the reference's Python encoder), one single-threaded process per core; "
              "scan generation is outside the timed part")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": statistics.median(ms), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(N_SCANS), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index: int, period: float = 0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from neural_spectral_codec_b200 import SpectralEncoder, synth
    from neural_spectral_codec_b200.distributed import ShardedEncoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep each rank (and the pinned host buffers it allocates) on the CPUs next to its GPU
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[local_rank]) if visible else local_rank
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
        except Exception:
            pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n_scans = args.scans
    enc = SpectralEncoder(n_elevation=16, n_azimuth=360, n_bins=50, alpha=2.0, learnable_alpha=True,
                          target_elevation_bins=16).to(dev)

    # synthetic scans of this rank, generated on the device from per-scan seeds
    first = rank * n_scans
    scans = [synth.make_scan(synth.SHAPES[args.shape], first + i, device=dev, shuffle=args.shuffle)
             for i in range(n_scans)]
    counts = torch.tensor([0] + [s.shape[0] for s in scans], dtype=torch.int64)
    offsets = torch.cumsum(counts, 0).to(dev)
    points = torch.cat(scans, 0)
    del scans
    total_points = int(points.shape[0])
    alg_bytes = 16 * total_points + 3200 * n_scans

    sharded = None
    if world > 1:
        try:
            sharded = ShardedEncoder(enc, world * n_scans, mode=args.gather)
        except Exception as exc:   # symmetric memory unavailable on this box: use the NCCL gather
            if args.gather != "fused":
                raise
            print(f"[bench] fused gather unavailable ({type(exc).__name__}: {exc}); using nccl", file=sys.stderr)
            ok = torch.tensor([0], device=dev)
        else:
            ok = torch.tensor([1], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)      # every rank must take the same path
        if int(ok.item()) == 0 and args.gather == "fused":
            args.gather = "nccl"
            sharded = ShardedEncoder(enc, world * n_scans, mode="nccl")
    out = torch.empty((n_scans, enc.output_dim), dtype=torch.float32, device=dev)

    def step():
        if sharded is None:
            enc.encode_points_batch(points, offsets, out=out)
        else:
            sharded.encode(points, offsets)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()

    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ
                           else int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    sampler.start()
    stream = torch.cuda.current_stream(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps)]
    barrier()
    for k in range(args.steps):
        ev[2 * k].record(stream)
        step()
        ev[2 * k + 1].record(stream)
    barrier()
    clocks = sampler.stop()
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(args.steps)]
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * n_scans * args.steps / (total_ms * 1e-3)

    # the fused kernel alone (same launches, N=1 path) for the roofline
    kern_ms = step_ms
    if world > 1:
        kev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps)]
        for k in range(args.steps):
            kev[2 * k].record(stream)
            enc.encode_points_batch(points, offsets, out=out)
            kev[2 * k + 1].record(stream)
        torch.cuda.synchronize(dev)
        kern_ms = [kev[2 * k].elapsed_time(kev[2 * k + 1]) for k in range(args.steps)]
    kern_avg_ms = sum(kern_ms) / len(kern_ms)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (kern_avg_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("scans") == n_scans and args.shape == "hdl64" and not args.shuffle:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "kernel": "encode_points_kernel<4,0>",
                "kernel_ms": kern_avg_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "peak_source": peak_src}

    # end to end: pinned host buffers -> descriptors on the host, through the public API
    n_e2e = min(n_scans, args.e2e_scans)
    e_off = offsets[: n_e2e + 1].cpu()
    h_points = torch.empty((int(e_off[-1]), 4), dtype=torch.float32).pin_memory()
    h_points.copy_(points[: int(e_off[-1])])
    h_out = torch.empty((n_e2e, enc.output_dim), dtype=torch.float32).pin_memory()
    hp, ho, hoff = h_points.numpy(), h_out.numpy(), e_off.numpy()
    enc.encode_scans((hp, hoff), out=ho)
    barrier()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        enc.encode_scans((hp, hoff), out=ho)     # synchronous: returns when ho is complete
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": world * n_e2e * e2e_steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": int(hp.nbytes + hoff.nbytes), "d2h_bytes_per_step": int(ho.nbytes),
           "scans_per_step_per_gpu": n_e2e, "steps": e2e_steps,
           "api": "SpectralEncoder.encode_scans -> nsc_pipeline_encode (pinned host buffers)"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        per_core = 64
        v, n = cpu_oracle_throughput(per_core, cores)
        # the reference's native call pattern: one process looping encode_points (pipeline.py:336-354)
        from oracle import nsc_oracle as orc
        cfg = orc.OracleConfig()
        host_scans = [points[int(offsets[i]):int(offsets[i + 1])].cpu().numpy() for i in range(12)]
        orc.encode_points(host_scans[0], cfg)
        t0 = time.perf_counter()
        for s_ in host_scans:
            orc.encode_points(s_, cfg)
        single = len(host_scans) / (time.perf_counter() - t0)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{n} HDL-64 scans ({per_core} per core), oracle/nsc_oracle.py, "
                                  "one single-threaded process per core",
                        "single_process_value": single,
                        "single_process_sample": f"{len(host_scans)} scans in one process, torch threads = "
                                                 f"{torch.get_num_threads()} (the reference's per-scan loop)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(n_scans, args.shape, args.shuffle), "scans_per_gpu": n_scans,
                       "points_per_gpu": total_points, "input_bytes_per_gpu": 16 * total_points,
                       "l2": f"inputs ({16 * total_points / 1e9:.1f} GB) larger than L2, no flush needed",
                       "gather": ("none (single GPU)" if world == 1 else args.gather),
                       "encoder": "n_elevation=16 n_azimuth=360 n_bins=50 alpha=2.0 target_rows=16"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": args.steps, "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- retrieval (secondary)
def run_retrieval(args):
    """Secondary workload: Wasserstein top-K over a 100 k x 800 database.
    The reference's only stated target for this stage is 27 ms per query at 100 k descriptors
    (configs/training.yaml:99). Not the BASELINE.json metric; printed in the same JSON shape."""
    import torch

    from neural_spectral_codec_b200.retrieval import WassersteinRetriever
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(7)
    n_db, nq, k = args.db, args.queries, args.topk
    db = torch.rand((n_db, 800), generator=g, device=dev) ** 4
    db /= db.sum(1, keepdim=True)
    q = db[torch.randint(0, n_db, (nq,), device=dev)] * (1 + 0.05 * torch.rand((nq, 800), device=dev))
    r = WassersteinRetriever(device=dev)
    r.add_to_database(db)
    for _ in range(max(args.warmup, 3)):
        r.query_batch(q, top_k=k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        r.query_batch(q, top_k=k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    passes = -(-nq // 8)                                   # one pass over the CDF rows serves 8 queries
    alg_bytes = passes * n_db * 800 * 4 + nq * n_db * 4 * 3
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else FALLBACK_HBM_GBS
    line = {"metric": "retrieval_queries_per_sec_at_100k_db", "value": nq / (ms * 1e-3), "unit": "queries/s",
            "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"Wasserstein top-{k} of {nq} queries over a {n_db} x 800 descriptor database",
                       "ms_per_query": ms / nq, "reference_target_ms_per_query": 27.0},
            "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg_bytes / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                         "note": "whole call (distance pass + top-K kernels + host launch gaps)"},
            "gpu_launches": args.steps * (passes + 1)}
    if not args.no_cpu:
        from oracle import retrieval_oracle as ro
        dbc, qc = db.cpu(), q.cpu()
        n = min(nq, 4)
        t0 = time.perf_counter()
        for i in range(n):
            ro.query_topk(qc[i], dbc, k)
        cpu_ms = 1e3 * (time.perf_counter() - t0) / n
        line["cpu_baseline"] = {"value": 1e3 / cpu_ms, "unit": "queries/s", "cores": torch.get_num_threads(),
                                "kind": "port", "sample": f"{n} queries, oracle/retrieval_oracle.py (torch CPU)"}
    emit(line)


def emit(line: dict) -> None:
    """The one JSON line goes to the REAL stdout; everything else a library prints to fd 1
    (e.g. NCCL's version banner) was redirected to stderr in main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scans", type=int, default=N_SCANS, help="scans per GPU")
    ap.add_argument("--e2e-scans", type=int, default=1024, help="scans per end-to-end step")
    ap.add_argument("--gather", default="fused", choices=["nccl", "fused"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--shape", default="hdl64", choices=sorted(SHAPE_DESC),
                    help="sensor shape of the synthetic scans (BASELINE.json configs 2-4; default = the metric's config)")
    ap.add_argument("--shuffle", action="store_true", help="random point order inside each scan")
    ap.add_argument("--workload", default="encode", choices=["encode", "retrieval"],
                    help="encode = the BASELINE.json metric (default); retrieval = secondary stage-1 retrieval line")
    ap.add_argument("--db", type=int, default=100000, help="retrieval: database rows")
    ap.add_argument("--queries", type=int, default=8, help="retrieval: queries per call")
    ap.add_argument("--topk", type=int, default=10, help="retrieval: K")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "retrieval":
        run_retrieval(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()

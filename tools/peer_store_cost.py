#!/usr/bin/env python
"""What do the extra descriptor stores of the fused gather cost by themselves? Single GPU: the
"peers" are 1, 2, 4 or 8 LOCAL buffers, so there is no NVLink in the picture -- only the store
instructions of the tail warps and the write traffic they add to the read stream.

    python tools/peer_store_cost.py [--scans 4541] [--steps 20]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_spectral_codec_b200 import SpectralEncoder, _lib, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=4541)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda")
    lib = _lib.load()
    enc = SpectralEncoder(n_elevation=16, target_elevation_bins=16).to(dev)
    points, offsets = synth.make_batch_resident(synth.HDL64, 0, a.scans, dev)
    out = torch.empty((a.scans, 800), device=dev)
    ws = torch.empty(64, dtype=torch.int32, device=dev)
    p, lut = enc._params(), enc.freq_to_bin()
    stream = torch.cuda.current_stream(dev).cuda_stream
    dbs = [torch.zeros((8 * a.scans, 800), device=dev) for _ in range(8)]

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps

    res = {}
    for rnd in range(2):
        res.setdefault("plain", []).append(timed(lambda: enc.encode_points_batch(points, offsets, out=out)))
        for n in (1, 2, 4, 8):
            ptrs = (C.c_void_p * n)(*[d.data_ptr() for d in dbs[:n]])

            def fn():
                st = lib.nsc_encode_batch_peers(points.data_ptr(), 4, offsets.data_ptr(), 0, a.scans, C.byref(p),
                                                lut.ctypes.data, ptrs, n, 3 * a.scans, ws.data_ptr(), 256, stream)
                assert st == 0
            res.setdefault(f"local_x{n}", []).append(timed(fn))
    print(json.dumps({k: min(v) for k, v in res.items()}))
    assert torch.equal(dbs[7][3 * a.scans:4 * a.scans], out)


if __name__ == "__main__":
    main()

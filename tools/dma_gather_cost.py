#!/usr/bin/env python
"""Single-GPU emulation of a copy-engine gather overlapped with the encode kernel: while the
kernel streams its 8.7 GB, a side stream copies n blocks of 14.5 MB (the descriptors of the
previous step) with cudaMemcpyAsync, as the DMA engines of 7 peers would write them into this
GPU's database. Compare with tools/peer_store_cost.py (the same bytes as 3.2 KB stores from the
kernel's tail warps).

    python tools/dma_gather_cost.py [--scans 4541] [--steps 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_spectral_codec_b200 import SpectralEncoder, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=4541)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda")
    enc = SpectralEncoder(n_elevation=16, target_elevation_bins=16).to(dev)
    points, offsets = synth.make_batch_resident(synth.HDL64, 0, a.scans, dev)
    out = torch.empty((a.scans, 800), device=dev)
    src = torch.zeros((a.scans, 800), device=dev)
    db = torch.zeros((8, a.scans, 800), device=dev)
    side = torch.cuda.Stream()
    main_s = torch.cuda.current_stream()

    def run(n_copies, chunk_rows=None):
        def step():
            ev = torch.cuda.Event()
            ev.record(main_s)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                for c in range(n_copies):
                    if chunk_rows is None:
                        db[c].copy_(src, non_blocking=True)
                    else:
                        for r0 in range(0, a.scans, chunk_rows):
                            db[c, r0:r0 + chunk_rows].copy_(src[r0:r0 + chunk_rows], non_blocking=True)
            enc.encode_points_batch(points, offsets, out=out)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step()
        main_s.wait_stream(side)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps

    res = {}
    for rnd in range(2):
        for n in (0, 1, 3, 7):
            res.setdefault(f"dma_x{n}", []).append(run(n))
        res.setdefault("dma_x7_chunks_of_512_rows", []).append(run(7, 512))
    print(json.dumps({k: min(v) for k, v in res.items()}))


if __name__ == "__main__":
    main()

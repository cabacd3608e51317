#!/usr/bin/env python
"""BASELINE.json configs[4]: a 100 000-scan encode sharded over the GPUs of one box, descriptors
gathered into the database replicated on every GPU (320 MB), with the checks of SURVEY.md 8(d) C5:
every rank holds the same database bytes, and every 391st scan matches the CPU oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29513 tests/multi_gpu_c5.py [--scans 100000] [--gather fused|nccl]

Lives under tests/ because it checks against the oracle (test infrastructure); it is a script for
a multi-GPU box, not collected by pytest.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_spectral_codec_b200 import SpectralEncoder, synth  # noqa: E402
from neural_spectral_codec_b200.distributed import ShardedEncoder  # noqa: E402
from oracle import nsc_oracle as orc  # noqa: E402  (checker only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=100000)
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"])
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    enc = SpectralEncoder(n_elevation=16, target_elevation_bins=16).to(dev)
    lo, hi = synth.shard_range(a.scans, world, rank)
    scans = [synth.make_scan(synth.HDL64, i, device=dev) for i in range(lo, hi)]
    offs = torch.cumsum(torch.tensor([0] + [s.shape[0] for s in scans], dtype=torch.int64), 0).to(dev)
    pts = torch.cat(scans, 0)
    del scans
    se = ShardedEncoder(enc, a.scans, mode=a.gather)
    for _ in range(2):
        se.encode(pts, offs)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        db = se.encode(pts, offs)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)

    # 1. identical bytes on every rank: compare a 64-bit checksum of the raw descriptor bits
    bits = db.contiguous().view(torch.int32).to(torch.int64)
    w = torch.arange(1, bits.numel() + 1, device=dev, dtype=torch.int64).view_as(bits)
    checksum = ((bits * (w % 1000003)).sum() % (1 << 61)).view(1)
    sums = [torch.zeros_like(checksum) for _ in range(world)]
    dist.all_gather(sums, checksum)
    same = all(int(s.item()) == int(sums[0].item()) for s in sums)

    # 2. every 391st scan against the CPU oracle (each rank checks the ones it owns)
    cfg = orc.OracleConfig()
    o = offs.cpu().numpy()
    worst, checked = 0.0, 0
    for g in range(0, a.scans, 391):
        if lo <= g < hi:
            s = pts[o[g - lo]:o[g - lo + 1]].cpu().numpy()
            ref = orc.encode_points(s, cfg).numpy()
            worst = max(worst, float(np.abs(db[g].cpu().numpy() - ref).max()))
            checked += 1
    stats = torch.tensor([worst, float(checked)], dtype=torch.float64, device=dev)
    mx = stats.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"config": "100k-scan encode sharded over GPUs with descriptors gathered into the replicated DB",
                          "n_gpus": world, "scans": a.scans, "gather": a.gather, "ms_per_pass": float(ms.item()),
                          "scans_per_s": a.scans / (float(ms.item()) * 1e-3),
                          "db_bytes": int(db.numel() * 4), "db_identical_on_all_ranks": same,
                          "oracle_spot_checks": int(stats[1].item()), "max_abs_diff_vs_oracle": float(mx[0].item()),
                          "points_per_gpu": int(pts.shape[0])}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not same or float(mx[0].item()) > 1e-4:
        raise SystemExit("C5 check failed")


if __name__ == "__main__":
    main()

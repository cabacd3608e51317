#!/usr/bin/env python
"""Multi-GPU invariant: the replicated descriptor database is bit-identical
for every rank and for both gather modes, and equals a single-GPU encode of the same scans.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_sharded.py [--scans 64]
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_spectral_codec_b200 import SpectralEncoder, synth  # noqa: E402
from neural_spectral_codec_b200.distributed import ShardedEncoder  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=61)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    shape = synth.SensorShape("s", 64, -24.8, 2.0, 800)
    enc = SpectralEncoder(n_elevation=16, target_elevation_bins=16).to(dev)
    n = args.scans
    lo, hi = synth.shard_range(n, world, rank)
    pts, offs = synth.make_batch(shape, lo, hi - lo, device=dev)
    full_pts, full_offs = synth.make_batch(shape, 0, n, device=dev)
    want = enc.encode_points_batch(full_pts, full_offs)
    ok = True
    for mode, lag in (("nccl", 0), ("fused", 0), ("fused", 1)):
        se = ShardedEncoder(enc, n, mode=mode, lag=lag)
        for rep in range(5):
            db = se.encode(pts, offs)
            if lag:
                assert (db is None) == (rep == 0)
                db = se.flush() if rep % 2 else db
            if db is None:
                continue
            torch.cuda.synchronize()
            same = bool(torch.equal(db, want))
            ok &= same
            if rank == 0 or not same:
                print(f"rank {rank} mode {mode} lag {lag} rep {rep}: database == single-GPU encode: {same}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        raise SystemExit("sharded encode mismatch")
    if rank == 0:
        print("SHARDED OK", flush=True)


if __name__ == "__main__":
    main()

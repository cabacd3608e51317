"""World-size-2 (and 3) CPU tests of the sharding and gather logic over the gloo backend."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neural_spectral_codec_b200 import synth
from neural_spectral_codec_b200.distributed import gather_descriptors, padded_rows


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def fake_descriptor(scan_index: int, dim: int = 800) -> torch.Tensor:
    g = torch.Generator().manual_seed(1234 + scan_index)
    d = torch.rand(dim, generator=g)
    return d / d.sum()


def _worker(rank, world, port, n_scans, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = synth.shard_range(n_scans, world, rank)
        per = padded_rows(n_scans, world)
        local = torch.zeros(per, 800)
        for i in range(lo, hi):
            local[i - lo] = fake_descriptor(i)      # stands in for the CUDA encode of scan i
        db = gather_descriptors(local, n_scans)
        np.save(os.path.join(result_dir, f"db_{rank}.npy"), db.numpy())
        with pytest.raises(ValueError):
            gather_descriptors(local[:-1], n_scans)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_scans", [(2, 10), (2, 7), (3, 8), (2, 1)])
def test_gathered_database_is_identical_on_every_rank_and_to_one_rank(tmp_path, world, n_scans):
    mp.spawn(_worker, args=(world, free_port(), n_scans, str(tmp_path)), nprocs=world, join=True)
    want = torch.stack([fake_descriptor(i) for i in range(n_scans)]).numpy()
    for r in range(world):
        got = np.load(tmp_path / f"db_{r}.npy")
        np.testing.assert_array_equal(got, want)    # bitwise: the gather moves bytes only


def test_shard_range_partitions_every_scan_once():
    for n in (0, 1, 7, 4541, 100000):
        for g in (1, 2, 4, 8):
            covered = []
            for r in range(g):
                lo, hi = synth.shard_range(n, g, r)
                assert 0 <= lo <= hi <= n and hi - lo <= padded_rows(n, g)
                assert lo == min(n, r * padded_rows(n, g))
                covered += list(range(lo, hi))
            assert covered == list(range(n))


def test_scan_content_is_independent_of_sharding():
    a = synth.make_scan(synth.SensorShape("s", 16, -24.8, 2.0, 100), 5)
    pts, offs = synth.make_batch(synth.SensorShape("s", 16, -24.8, 2.0, 100), 3, 4)
    o = offs.numpy()
    np.testing.assert_array_equal(pts[o[2]:o[3]].numpy(), a.numpy())   # NaN rows compare equal

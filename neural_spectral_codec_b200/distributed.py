"""Multi-GPU encode: scans are independent, so ranks encode contiguous blocks of scan indices
and the descriptors are gathered into a database replicated on every GPU -- what
``WassersteinRetriever.add_to_database`` (reference ``src/retrieval/wasserstein.py:300-326``)
would hold after the same scans were added one at a time.

One process per GPU (``torch.distributed``). Two ways to fill the database:
  * ``mode="nccl"``  -- fused encode kernel into a local block, then ONE
    ``all_gather_into_tensor`` over NVLink;
  * ``mode="fused"`` -- the database lives in symmetric memory and the encode kernel's epilogue
    stores every descriptor straight into all peers' copies (``nsc_encode_batch_peers``). The
    database is double-buffered (step s fills buffer s & 1), so nothing has to be synchronised
    before the encode; the step ends with one small signal-and-wait kernel
    (``nsc_peer_signal_wait``: a release store into every peer's flag array, acquire-polls on
    this rank's own) instead of two symmetric-memory barriers.
On the CPU test backend (gloo) the gather runs on host tensors produced elsewhere; the encode
itself always needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import torch
import torch.distributed as dist

from . import _lib
from .synth import shard_range


def padded_rows(n_scans: int, world_size: int) -> int:
    """Rows per rank of the gathered database: ceil(B / G) (tail rank padded)."""
    return -(-n_scans // world_size)


def gather_descriptors(local: torch.Tensor, n_scans: int, group=None) -> torch.Tensor:
    """All-gather ``(rows_per_rank, D)`` blocks into the replicated ``(n_scans, D)`` database.

    ``local`` holds this rank's descriptors in its first ``hi - lo`` rows (``shard_range``);
    every rank must pass the same padded shape. Works on any backend (NCCL on GPUs, gloo on
    CPU tensors in the tests).
    """
    world = dist.get_world_size(group)
    per = padded_rows(n_scans, world)
    if local.shape[0] != per:
        raise ValueError(f"local block must have {per} rows (ceil(B/G)), got {local.shape[0]}")
    db = torch.empty((world * per, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(db, local.contiguous(), group=group)
    return db[:n_scans]


class ShardedEncoder:
    """Encodes this rank's block of a global batch and keeps the replicated database."""

    def __init__(self, encoder, n_scans: int, mode: str = "nccl", group=None):
        if mode not in ("nccl", "fused"):
            raise ValueError("mode must be 'nccl' or 'fused'")
        self.encoder = encoder
        self.n_scans = int(n_scans)
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.lo, self.hi = shard_range(self.n_scans, self.world, self.rank)
        self.per = padded_rows(self.n_scans, self.world)
        self.mode = mode
        self.D = encoder.output_dim
        dev = encoder.alpha.device
        if dev.type != "cuda":
            raise RuntimeError("ShardedEncoder needs the encoder on a CUDA device")
        self.device = dev
        if mode == "nccl":
            self.local = torch.zeros((self.per, self.D), dtype=torch.float32, device=dev)
            self.db = torch.empty((self.world * self.per, self.D), dtype=torch.float32, device=dev)
            self._peer_ptrs = None
        else:
            import torch.distributed._symmetric_memory as symm
            gname = (group or dist.group.WORLD).group_name
            rows = self.world * self.per
            self._dbs = symm.empty((2, rows, self.D), dtype=torch.float32, device=dev)   # ping-pong
            self._dbs.zero_()
            self._hdl = symm.rendezvous(self._dbs, gname)
            half = rows * self.D * 4
            self._peer_ptrs = [(C.c_void_p * self.world)(*[int(self._hdl.buffer_ptrs[r]) + b * half
                                                           for r in range(self.world)]) for b in (0, 1)]
            self._flags = symm.empty((64,), dtype=torch.int32, device=dev)
            self._flags.zero_()
            self._flag_hdl = symm.rendezvous(self._flags, gname)
            self._flag_ptrs = (C.c_void_p * self.world)(*[int(self._flag_hdl.buffer_ptrs[r])
                                                          for r in range(self.world)])
            self._flag_hdl.barrier()          # every rank's flags are zero before anyone signals
            torch.cuda.synchronize(dev)
            self._step = 0
            self.db = self._dbs[0]
            self._ws = torch.empty(64, dtype=torch.int32, device=dev)
            self.local = None

    def encode(self, points: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
        """``points`` / ``offsets`` describe THIS rank's scans ``[lo, hi)``. Returns the
        replicated ``(n_scans, D)`` database (valid on return for "nccl"; for "fused" after
        the barrier this method issues)."""
        n_local = self.hi - self.lo
        if offsets.numel() - 1 != n_local:
            raise ValueError(f"rank {self.rank} owns {n_local} scans, got {offsets.numel() - 1}")
        if self.mode == "nccl":
            if n_local:
                self.encoder.encode_points_batch(points, offsets, out=self.local[:n_local])
            dist.all_gather_into_tensor(self.db, self.local, group=self.group)
        else:
            lib = _lib.load()
            from .encoder import _check_batch
            points, offsets, n, stride = _check_batch(points, offsets)
            # Peers store into this rank's database from THEIR streams. Step s fills buffer s & 1:
            # a rank that has left the wait of step s-1 knows every peer is past (in stream order)
            # whatever read that buffer after step s-2, so the stores below need no barrier first.
            b = self._step & 1
            p = self.encoder._params()
            lut = self.encoder.freq_to_bin()
            stream = torch.cuda.current_stream(self.device).cuda_stream
            with torch.cuda.device(self.device):
                st = lib.nsc_encode_batch_peers(
                    points.data_ptr(), stride, offsets.data_ptr(), 0, n, C.byref(p),
                    lut.ctypes.data, self._peer_ptrs[b], self.world, self.rank * self.per,
                    self._ws.data_ptr(), self._ws.numel() * 4, stream)
                _lib.check(st, "nsc_encode_batch_peers")
                self._step += 1
                st = lib.nsc_peer_signal_wait(self._flag_ptrs, self.world, self.rank, self._step, stream)
            _lib.check(st, "nsc_peer_signal_wait")
            self.db = self._dbs[b]
        # rank r owns global rows [r*per, r*per + n_r): the database is contiguous in scan index
        return self.db[:self.n_scans]

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
    """Encodes this rank's block of a global batch and keeps the replicated database.

    ``mode="fused"``, ``lag=0``: ``encode`` returns the database of THIS step (complete in stream
    order). ``lag=1`` pipelines the gather: ``encode`` enqueues step s and waits only for step
    s-1, so a rank is never held up by the slowest peer of the current step; it returns the
    database of step s-1 (``None`` on the first call) and ``flush()`` completes the last one.
    Four database buffers instead of two (see ``nsc_peer_signal_wait`` in the header)."""

    def __init__(self, encoder, n_scans: int, mode: str = "nccl", group=None, lag: int = 0):
        if mode not in ("nccl", "fused"):
            raise ValueError("mode must be 'nccl' or 'fused'")
        if lag not in (0, 1) or (lag and mode != "fused"):
            raise ValueError("lag must be 0, or 1 with mode='fused'")
        self.encoder = encoder
        self.n_scans = int(n_scans)
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.lo, self.hi = shard_range(self.n_scans, self.world, self.rank)
        self.per = padded_rows(self.n_scans, self.world)
        self.mode = mode
        self.lag = lag
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
            self._nbuf = 4 if lag else 2
            self._dbs = symm.empty((self._nbuf, rows, self.D), dtype=torch.float32, device=dev)
            self._dbs.zero_()
            self._hdl = symm.rendezvous(self._dbs, gname)
            one = rows * self.D * 4
            self._peer_ptrs = [(C.c_void_p * self.world)(*[int(self._hdl.buffer_ptrs[r]) + b * one
                                                           for r in range(self.world)]) for b in range(self._nbuf)]
            self._flags = symm.empty((64,), dtype=torch.int32, device=dev)
            self._flags.zero_()
            self._flag_hdl = symm.rendezvous(self._flags, gname)
            self._flag_ptrs = (C.c_void_p * self.world)(*[int(self._flag_hdl.buffer_ptrs[r])
                                                          for r in range(self.world)])
            self._flag_hdl.barrier()          # every rank's flags are zero before anyone signals
            torch.cuda.synchronize(dev)
            self._step = 0                    # steps enqueued
            self._complete = 0                # steps whose database is complete in stream order
            self.db = self._dbs[0]
            need = int(_lib.load().nsc_workspace_bytes(self.per, C.byref(encoder._params())))
            self._ws = torch.empty((need + 3) // 4, dtype=torch.int32, device=dev)
            self.local = None

    def _signal_wait(self, wait_value: int) -> None:
        with torch.cuda.device(self.device):
            st = _lib.load().nsc_peer_signal_wait(self._flag_ptrs, self.world, self.rank, self._step, wait_value,
                                                  torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(st, "nsc_peer_signal_wait")
        self._complete = max(self._complete, wait_value)
        if self._complete:
            self.db = self._dbs[(self._complete - 1) % self._nbuf]

    def flush(self) -> torch.Tensor:
        """Completes every enqueued step (a no-op unless ``lag=1``) and returns the database."""
        if self.mode == "fused" and self._complete < self._step:
            self._signal_wait(self._step)
        return self.db[:self.n_scans]

    def encode(self, points: torch.Tensor, offsets: torch.Tensor, wait: bool = True):
        """``points`` / ``offsets`` describe THIS rank's scans ``[lo, hi)``. Returns the latest
        complete replicated ``(n_scans, D)`` database in stream order: this step's (``lag=0``) or
        the previous step's (``lag=1``; ``None`` before there is one). ``wait=False`` (fused mode,
        measurements only) skips the synchronisation kernel: the caller must synchronise all
        ranks before the next call."""
        n_local = self.hi - self.lo
        if offsets.numel() - 1 != n_local:
            raise ValueError(f"rank {self.rank} owns {n_local} scans, got {offsets.numel() - 1}")
        if self.mode == "nccl":
            if n_local:
                self.encoder.encode_points_batch(points, offsets, out=self.local[:n_local])
            dist.all_gather_into_tensor(self.db, self.local, group=self.group)
            return self.db[:self.n_scans]
        lib = _lib.load()
        from .encoder import _check_batch
        points, offsets, n, stride = _check_batch(points, offsets)
        # Peers store into this rank's database from THEIR streams. Step s fills buffer s mod 2
        # (mod 4 with lag): a rank that has left the wait for step s-1 (s-2) knows every peer is
        # past, in stream order, whatever read that buffer last, so no barrier precedes the stores.
        b = self._step % self._nbuf
        p = self.encoder._params()
        lut = self.encoder.freq_to_bin()
        with torch.cuda.device(self.device):
            st = lib.nsc_encode_batch_peers(
                points.data_ptr(), stride, offsets.data_ptr(), 0, n, C.byref(p),
                lut.ctypes.data, self._peer_ptrs[b], self.world, self.rank * self.per,
                self._ws.data_ptr(), self._ws.numel() * 4, torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(st, "nsc_encode_batch_peers")
        if wait:
            self._step += 1
            self._signal_wait(self._step - self.lag)
        # rank r owns global rows [r*per, r*per + n_r): the database is contiguous in scan index
        return self.db[:self.n_scans] if self._complete else None

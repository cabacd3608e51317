"""uint16 wire format of the descriptors on the GPU.

``HistogramQuantizer`` mirrors the reference class (reference
``src/encoding/quantization.py:112-192``) with the row length as a parameter (the reference
instantiates it for 50 bins; the encoder's descriptor has ``target_elevation_bins * n_bins`` =
800) and accepts batches. ``CompressedDescriptor`` is the reference's record
(:22-109) with the histogram field sized to the descriptor: ``2 * n_bins`` bytes of uint16
followed by the same 120 bytes of metadata (pose 7 x float32, timestamp float64, keyframe id
uint32, SHA-1 of the cloud, 60 reserved bytes) -- 1720 bytes for 800 bins, 220 for 50.
The quantisation arithmetic runs in ``libnsc_b200.so``; record packing is host-side ``struct``.
"""
from __future__ import annotations

import hashlib
import struct
from dataclasses import dataclass
from typing import Union

import numpy as np
import torch

from . import _lib

METADATA_BYTES = 120


class HistogramQuantizer:
    def __init__(self, n_bins: int = 800, epsilon: float = 1e-8, device: str = "cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("HistogramQuantizer: no CPU implementation; pass a CUDA device")
        _lib.load()
        self.n_bins = n_bins
        self.epsilon = epsilon
        self.max_value = 65535
        self.device = dev

    def _rows(self, x, dtype):
        is_np = isinstance(x, np.ndarray)
        t = torch.from_numpy(np.ascontiguousarray(x)) if is_np else x
        single = t.dim() == 1
        if single:
            t = t.unsqueeze(0)
        if t.dim() != 2 or t.shape[1] != self.n_bins:
            raise AssertionError(f"Expected {self.n_bins} bins, got {t.shape[-1]}")
        return t.detach().to(self.device, dtype).contiguous(), is_np, single

    def quantize(self, histogram: Union[np.ndarray, torch.Tensor]):
        """``(n_bins,)`` or ``(B, n_bins)`` float histograms -> uint16 of the same shape
        (numpy in -> numpy out; CUDA tensor in -> CUDA int32-free uint16 tensor out)."""
        lib = _lib.load()
        h, is_np, single = self._rows(histogram, torch.float32)
        q = torch.empty(h.shape, dtype=torch.uint16, device=self.device)
        with torch.cuda.device(self.device):
            st = lib.nsc_quantize_histograms(h.data_ptr(), h.shape[0], self.n_bins, self.epsilon,
                                             q.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(st, "nsc_quantize_histograms")
        if single:
            q = q[0]
        return q.cpu().numpy() if is_np else q

    def dequantize(self, quantized: Union[np.ndarray, torch.Tensor]):
        """uint16 rows -> normalised float32 rows (reference :169-192)."""
        lib = _lib.load()
        q, is_np, single = self._rows(quantized, torch.uint16)
        h = torch.empty(q.shape, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = lib.nsc_dequantize_histograms(q.data_ptr(), q.shape[0], self.n_bins, self.epsilon,
                                               h.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(st, "nsc_dequantize_histograms")
        if single:
            h = h[0]
        return h.cpu().numpy() if is_np else h


def compute_point_cloud_hash(points: np.ndarray) -> bytes:
    """SHA-1 of the float32 xyz bytes (reference quantization.py:195-211)."""
    return hashlib.sha1(np.ascontiguousarray(points[:, :3], dtype=np.float32).tobytes()).digest()


@dataclass
class CompressedDescriptor:
    """One keyframe record: ``2 * len(histogram) + 120`` bytes (reference :22-109)."""
    histogram: np.ndarray          # (n_bins,) uint16
    pose: np.ndarray               # (7,) [x, y, z, qw, qx, qy, qz]
    timestamp: float
    keyframe_id: int
    point_cloud_hash: bytes        # 20 bytes

    def to_bytes(self) -> bytes:
        if len(self.point_cloud_hash) != 20:
            raise ValueError("point_cloud_hash must be 20 bytes (SHA-1)")
        total = (np.asarray(self.histogram).astype(np.uint16).tobytes()
                 + np.asarray(self.pose).astype(np.float32).tobytes()
                 + struct.pack("d", self.timestamp) + struct.pack("I", self.keyframe_id)
                 + self.point_cloud_hash + bytes(60))
        assert len(total) == 2 * len(self.histogram) + METADATA_BYTES
        return total

    @staticmethod
    def from_bytes(data: bytes) -> "CompressedDescriptor":
        n = len(data) - METADATA_BYTES
        if n <= 0 or n % 2:
            raise AssertionError(f"record of {len(data)} bytes is not 2 * n_bins + {METADATA_BYTES}")
        return CompressedDescriptor(
            histogram=np.frombuffer(data[:n], dtype=np.uint16),
            pose=np.frombuffer(data[n:n + 28], dtype=np.float32),
            timestamp=struct.unpack("d", data[n + 28:n + 36])[0],
            keyframe_id=struct.unpack("I", data[n + 36:n + 40])[0],
            point_cloud_hash=data[n + 40:n + 60])

"""CPU oracle for the uint16 descriptor quantiser -- TEST INFRASTRUCTURE ONLY.

Restates ``HistogramQuantizer`` of the reference (``src/encoding/quantization.py:131-192``) for
any row length; pinned by ``tests/golden/quantization.npz`` (outputs of the unmodified reference,
``tests/golden/make_golden_quantization.py``). Only ``tests/`` imports it.
"""
import numpy as np

MAX_VALUE = 65535


def quantize(histogram: np.ndarray, epsilon: float = 1e-8) -> np.ndarray:
    """quantization.py:131-167."""
    h = np.asarray(histogram, np.float32)
    s = h.sum()                                                 # :144 (NumPy pairwise float32 sum)
    if s > epsilon:                                             # :145
        h = h / (s + epsilon)                                   # :146
    q = np.round(h * MAX_VALUE).astype(np.uint16)               # :150
    total = int(q.astype(np.int64).sum())                       # :154
    if total > 0:                                               # :155
        error = MAX_VALUE - total                               # :157
        if error != 0:                                          # :159
            i = int(q.argmax())                                 # :161 first largest bin
            q[i] = np.uint16(min(max(int(q[i]) + error, 0), MAX_VALUE))   # :162-166
    return q


def dequantize(quantized: np.ndarray, epsilon: float = 1e-8) -> np.ndarray:
    """quantization.py:169-192."""
    h = np.asarray(quantized).astype(np.float32)                # :182
    s = h.sum()                                                 # :184
    if s > epsilon:                                             # :185
        return h / (s + epsilon)                                # :186
    return np.ones(len(h), dtype=np.float32) / len(h)           # :189

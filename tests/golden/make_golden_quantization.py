#!/usr/bin/env python
"""Golden vectors for the uint16 quantiser from the UNMODIFIED reference
(/root/reference/src/encoding/quantization.py, loaded by file path). Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_quantization.py
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
spec = importlib.util.spec_from_file_location("ref_quant", "/root/reference/src/encoding/quantization.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def main():
    rng = np.random.default_rng(99)
    out = {}
    for n_bins in (800, 50, 181, 7, 129, 2896):
        q = ref.HistogramQuantizer(n_bins=n_bins)
        rows = []
        if n_bins == 800:
            for name in ("hdl64_full", "hdl32_small", "beam128_small", "sparse_rows", "single_point", "empty"):
                rows.append(np.load(os.path.join(HERE, name + ".npz"))["descriptor"])
        g = rng.gamma(0.4, 1.0, (120, n_bins)).astype(np.float32)
        g /= g.sum(1, keepdims=True)
        rows += list(g)
        rows += list((rng.random((10, n_bins)) * 5).astype(np.float32))        # not normalised
        rows.append(np.zeros(n_bins, np.float32))
        rows.append(np.full(n_bins, 1e-12, np.float32))
        one = np.zeros(n_bins, np.float32); one[n_bins // 3] = 1.0
        rows.append(one)
        h = np.stack(rows).astype(np.float32)
        quant = np.stack([q.quantize(r) for r in h])
        deq = np.stack([q.dequantize(r) for r in quant])
        assert quant.dtype == np.uint16 and deq.dtype == np.float32
        out[f"hist{n_bins}"], out[f"quant{n_bins}"], out[f"deq{n_bins}"] = h, quant, deq
        print(n_bins, h.shape, "row sums of quantised:", np.unique(quant.astype(np.int64).sum(1))[:6])
    rec = ref.CompressedDescriptor(histogram=out["quant50"][0], pose=np.arange(7, dtype=np.float32) / 7,
                                   timestamp=1234.5678, keyframe_id=4242, point_cloud_hash=bytes(range(20)))
    out["record50"] = np.frombuffer(rec.to_bytes(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "quantization.npz"), **out)


if __name__ == "__main__":
    main()

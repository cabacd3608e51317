"""uint16 descriptor quantiser: oracle pinned to the reference's vectors (CPU), CUDA kernels
bit-exact against them (GPU), record packing compatible with the reference's 220-byte layout."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import quantization_oracle as qo

SIZES = (800, 50, 181, 7, 129, 2896)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "quantization.npz"))


@pytest.mark.parametrize("n_bins", SIZES)
def test_oracle_matches_reference_vectors(golden, n_bins):
    h, q, d = golden[f"hist{n_bins}"], golden[f"quant{n_bins}"], golden[f"deq{n_bins}"]
    for i in range(len(h)):
        np.testing.assert_array_equal(qo.quantize(h[i]), q[i])
        np.testing.assert_array_equal(qo.dequantize(q[i]), d[i])


def test_record_layout_matches_reference(golden):
    from neural_spectral_codec_b200.quantization import CompressedDescriptor, compute_point_cloud_hash
    rec = CompressedDescriptor(histogram=golden["quant50"][0], pose=np.arange(7, dtype=np.float32) / 7,
                               timestamp=1234.5678, keyframe_id=4242, point_cloud_hash=bytes(range(20)))
    raw = rec.to_bytes()
    assert len(raw) == 220
    np.testing.assert_array_equal(np.frombuffer(raw, np.uint8), golden["record50"])
    back = CompressedDescriptor.from_bytes(raw)
    np.testing.assert_array_equal(back.histogram, golden["quant50"][0])
    assert back.keyframe_id == 4242 and back.timestamp == 1234.5678 and back.point_cloud_hash == bytes(range(20))
    big = CompressedDescriptor(histogram=golden["quant800"][0], pose=np.zeros(7, np.float32), timestamp=0.5,
                               keyframe_id=7, point_cloud_hash=compute_point_cloud_hash(np.ones((5, 4), np.float32)))
    assert len(big.to_bytes()) == 1720
    np.testing.assert_array_equal(CompressedDescriptor.from_bytes(big.to_bytes()).histogram, golden["quant800"][0])


@pytest.mark.gpu
@pytest.mark.parametrize("n_bins", SIZES)
def test_cuda_quantiser_is_bit_exact(golden, n_bins):
    from neural_spectral_codec_b200.quantization import HistogramQuantizer
    h, q, d = golden[f"hist{n_bins}"], golden[f"quant{n_bins}"], golden[f"deq{n_bins}"]
    qz = HistogramQuantizer(n_bins=n_bins)
    got_q = qz.quantize(h)
    assert got_q.dtype == np.uint16
    np.testing.assert_array_equal(got_q, q)
    np.testing.assert_array_equal(qz.dequantize(q), d)
    np.testing.assert_array_equal(qz.quantize(h[3]), q[3])              # single-row signature
    dq = qz.dequantize(torch.from_numpy(q.astype(np.int32)).to(torch.uint16).cuda())
    assert dq.is_cuda and torch.equal(dq.cpu(), torch.from_numpy(d))
    with pytest.raises(AssertionError):
        qz.quantize(np.zeros(n_bins + 1, np.float32))


@pytest.mark.gpu
def test_cuda_quantiser_unaligned_rows_take_the_scalar_path(golden):
    """Rows that do not start on 16-byte boundaries (a view shifted by one element) cannot use the
    16-byte row moves; the result is the same bits."""
    from neural_spectral_codec_b200.quantization import HistogramQuantizer
    h, q, d = golden["hist800"], golden["quant800"], golden["deq800"]
    qz = HistogramQuantizer(n_bins=800)
    buf = torch.zeros(h.size + 1, dtype=torch.float32, device="cuda")
    buf[1:] = torch.from_numpy(h).cuda().flatten()
    shifted = buf[1:].view(h.shape)
    assert shifted.data_ptr() % 16 != 0
    got = qz.quantize(shifted)
    assert torch.equal(got.cpu().to(torch.int32), torch.from_numpy(q.astype(np.int32)))
    qbuf = torch.zeros(q.size + 1, dtype=torch.uint16, device="cuda")
    qbuf[1:] = got.flatten()
    qshift = qbuf[1:].view(q.shape)
    assert qshift.data_ptr() % 16 != 0
    assert torch.equal(qz.dequantize(qshift).cpu(), torch.from_numpy(d))


@pytest.mark.gpu
@pytest.mark.parametrize("n_bins", (800, 2896, 50))
def test_cuda_quantiser_unusual_rows(n_bins):
    """Rows outside the descriptor regime (the kernel divides those the long way): a negative
    element, huge and tiny sums, an all-zero row, one spike, zeros and denormals between ordinary
    values, a row below epsilon. Same integers / floats as the oracle."""
    from neural_spectral_codec_b200.quantization import HistogramQuantizer
    rng = np.random.default_rng(n_bins)
    base = (rng.random((10, n_bins)) ** 4).astype(np.float32)
    base /= base.sum(1, keepdims=True)
    rows = base.copy()
    rows[0, 5] = -1e-9                                   # sign bit set, rounds to 0 on both sides
    rows[1] *= np.float32(1e15)                          # sum beyond the moderate range
    rows[2] *= np.float32(1e-6)                          # small but moderate sum
    rows[3] = 0.0                                        # nothing to normalise, nothing to fix up
    rows[4] = 0.0
    rows[4, n_bins // 3] = 0.75                          # one spike takes all 65535
    rows[5, ::3] = 0.0
    rows[5, 1::7] = np.float32(1e-42)                    # denormals
    rows[6] *= np.float32(1e-9)                          # sum <= epsilon: not normalised
    rows[7] *= np.float32(3e-13)                         # far below epsilon
    rows[8, :] = np.float32(1.0 / n_bins)                # the uniform fallback descriptor
    rows[9] *= np.float32(2.0 ** 30)                     # large, still moderate
    qz = HistogramQuantizer(n_bins=n_bins)
    got = qz.quantize(rows)
    want = np.stack([qo.quantize(r) for r in rows])
    np.testing.assert_array_equal(got, want)
    back = qz.dequantize(got)
    np.testing.assert_array_equal(back, np.stack([qo.dequantize(r) for r in got]))
    # dequantise rows the quantiser never emits: all zero (uniform fallback), a single count, all 65535
    odd = np.zeros((3, n_bins), np.uint16)
    odd[1, n_bins - 1] = 1
    odd[2] = 65535
    np.testing.assert_array_equal(qz.dequantize(odd), np.stack([qo.dequantize(r) for r in odd]))


@pytest.mark.gpu
def test_encoder_descriptors_round_trip():
    """encode -> quantise -> dequantise keeps the descriptor within one quantisation step and the
    retrieval ranking of the exact descriptors."""
    from neural_spectral_codec_b200 import SpectralEncoder, synth
    from neural_spectral_codec_b200.quantization import HistogramQuantizer
    small = synth.SensorShape("s", 64, -24.8, 2.0, 600)
    pts, offs = synth.make_batch(small, 0, 32, device="cuda")
    desc = SpectralEncoder(n_elevation=16, target_elevation_bins=16).to("cuda").encode_points_batch(pts, offs)
    qz = HistogramQuantizer(n_bins=800)
    q = qz.quantize(desc)
    assert q.dtype == torch.uint16 and (q.to(torch.int64).sum(1) == 65535).all()
    back = qz.dequantize(q)
    # every bin is within half a quantisation step, except the largest one per row, which
    # absorbs the summed rounding error of the other 799 (quantization.py:154-167)
    err = (back - desc).abs()
    big = desc.argmax(1, keepdim=True)
    assert err.gather(1, big).max().item() <= 40.0 / 65535
    assert err.scatter(1, big, 0.0).max().item() <= 0.51 / 65535
    for i in (0, 31):
        np.testing.assert_array_equal(q[i].cpu().numpy(), qo.quantize(desc[i].cpu().numpy()))

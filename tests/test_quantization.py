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


def _kernel_sum_plan(n_bins):
    """The plan the kernels follow (host-side hook of the C ABI; no CUDA call)."""
    import ctypes as C
    from neural_spectral_codec_b200 import _lib
    lib = _lib.load()
    n_leaves, n_adds, result = C.c_int32(), C.c_int32(), C.c_int32()
    start, length = np.zeros(64, np.uint16), np.zeros(64, np.uint16)
    add_a, add_b = np.zeros(64, np.uint8), np.zeros(64, np.uint8)
    st = lib.nsc_test_pairwise_sum_plan(n_bins, C.byref(n_leaves), C.byref(n_adds), C.byref(result),
                                        start.ctypes.data, length.ctypes.data, add_a.ctypes.data, add_b.ctypes.data)
    return st, n_leaves.value, n_adds.value, result.value, start, length, add_a, add_b


def _sum_like_the_kernel(row):
    """csrc/nsc_quantize.cu numpy_sum, step for step in float32: per leaf 8 strided accumulators,
    ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), the leaf's tail, then the plan's additions."""
    f = np.float32
    st, n_leaves, n_adds, result, start, length, add_a, add_b = _kernel_sum_plan(len(row))
    assert st == 0
    slots = np.zeros(128, np.float32)
    for l in range(n_leaves):
        a = row[start[l]:start[l] + length[l]]
        if len(a) < 8:
            res = f(-0.0)
            for x in a:
                res = f(res + x)
        else:
            body = len(a) - len(a) % 8
            r = a[:8].copy()
            for i in range(8, body, 8):
                r = r + a[i:i + 8]
            res = f(f(f(r[0] + r[1]) + f(r[2] + r[3])) + f(f(r[4] + r[5]) + f(r[6] + r[7])))
            for x in a[body:]:
                res = f(res + x)
        slots[l] = res
    for t in range(n_adds):
        slots[n_leaves + t] = f(slots[add_a[t]] + slots[add_b[t]])
    return slots[result]


def test_kernel_sum_plan_is_numpys_pairwise_sum_for_every_row_length():
    """The quantised integers are bit-exact only if the float32 row sum is NumPy's; the GPU tests pin
    six row lengths against the reference, this one walks the kernels' plan for all 4096."""
    rng = np.random.default_rng(0)
    for n_bins in range(1, 4097):
        row = (rng.random(n_bins) ** 6 * 10.0 ** rng.integers(-3, 4, n_bins)).astype(np.float32)
        assert _sum_like_the_kernel(row).tobytes() == row.sum().tobytes(), n_bins
    # an all-zero row sums to zero (the sign of that zero is not NumPy's for rows of -0.0; the kernels
    # only compare the sum with epsilon and add epsilon to it, where the sign cannot matter)
    for row in (np.array([-0.0], np.float32), np.array([0.0, -0.0], np.float32), np.zeros(800, np.float32)):
        assert _sum_like_the_kernel(row) == row.sum() == 0.0
    assert _kernel_sum_plan(0)[0] != 0 and _kernel_sum_plan(4097)[0] != 0
    st, n_leaves, n_adds, result, start, length, _, _ = _kernel_sum_plan(800)
    assert (st, n_leaves, n_adds, result) == (0, 8, 7, 14)
    assert list(length[:8]) == [96, 104] * 4 and list(start[:3]) == [0, 96, 200]


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

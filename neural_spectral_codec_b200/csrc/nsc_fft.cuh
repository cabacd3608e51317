// 360-point complex FFT as three Stockham passes of radix 8, 9 and 5 with the butterflies held
// in registers (torch.fft.rfft at reference spectral_encoder.py:180 is a 360-point real
// transform; two rows ride in one complex transform, see spectrum_and_bins()).
//
// Pass (R, NS): butterfly j of N/R reads x[j + r*N/R], r < R, multiplies by the twiddle
// W^(k*r*N/(NS*R)) with k = j mod NS, takes an R-point DFT and writes y[(j/NS)*NS*R + k + q*NS].
// The butterfly functions are __host__ __device__ so tools/fft_selftest.cu can check them on
// the CPU against a float64 DFT.
#pragma once
#include <cuda_runtime.h>

namespace nsc {

#define NSC_FFT_HD __host__ __device__ __forceinline__

NSC_FFT_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
NSC_FFT_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
NSC_FFT_HD float2 cmul(float2 a, float2 w) {
    return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}
NSC_FFT_HD float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// forward DFTs: X[q] = sum_r a[r] exp(-2 pi i q r / R), in place.
NSC_FFT_HD void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_neg_i(csub(a1, a3));
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
}

NSC_FFT_HD void dft8(float2* a) {
    const float h = 0.70710678118654752440f;
    float2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];
    float2 o0 = a[1], o1 = a[3], o2 = a[5], o3 = a[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    o1 = make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));      // * (1 - i)/sqrt2
    o2 = mul_neg_i(o2);                                           // * -i
    o3 = make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y));     // * (-1 - i)/sqrt2
    a[0] = cadd(e0, o0);
    a[4] = csub(e0, o0);
    a[1] = cadd(e1, o1);
    a[5] = csub(e1, o1);
    a[2] = cadd(e2, o2);
    a[6] = csub(e2, o2);
    a[3] = cadd(e3, o3);
    a[7] = csub(e3, o3);
}

NSC_FFT_HD void dft3(float2& a0, float2& a1, float2& a2) {
    const float s = 0.86602540378443864676f;
    const float2 sum = cadd(a1, a2), d = csub(a1, a2);
    const float2 m = make_float2(fmaf(-0.5f, sum.x, a0.x), fmaf(-0.5f, sum.y, a0.y));
    const float2 r = make_float2(s * d.y, -s * d.x);              // -i * s * d
    a0 = cadd(a0, sum);
    a1 = cadd(m, r);
    a2 = csub(m, r);
}

NSC_FFT_HD void dft9(float2* a) {
    // n = 3 n1 + n2, k = k1 + 3 k2
    float2 b[3][3];   // b[n2][k1]
#pragma unroll
    for (int n2 = 0; n2 < 3; ++n2) {
        b[n2][0] = a[n2];
        b[n2][1] = a[n2 + 3];
        b[n2][2] = a[n2 + 6];
        dft3(b[n2][0], b[n2][1], b[n2][2]);
    }
    // twiddles w9^(n2*k1): w9 = exp(-2 pi i / 9)
    const float2 w1 = make_float2(0.76604444311897803520f, -0.64278760968653932632f);
    const float2 w2 = make_float2(0.17364817766693034885f, -0.98480775301220805937f);
    const float2 w4 = make_float2(-0.93969262078590838405f, -0.34202014332566873304f);
    b[1][1] = cmul(b[1][1], w1);
    b[1][2] = cmul(b[1][2], w2);
    b[2][1] = cmul(b[2][1], w2);
    b[2][2] = cmul(b[2][2], w4);
#pragma unroll
    for (int k1 = 0; k1 < 3; ++k1) {
        float2 c0 = b[0][k1], c1 = b[1][k1], c2 = b[2][k1];
        dft3(c0, c1, c2);
        a[k1] = c0;
        a[k1 + 3] = c1;
        a[k1 + 6] = c2;
    }
}

NSC_FFT_HD void dft5(float2* a) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const float2 t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]);
    const float2 t3 = csub(a[1], a[4]), t4 = csub(a[2], a[3]);
    const float2 m1 = make_float2(fmaf(c1, t1.x, fmaf(c2, t2.x, a[0].x)), fmaf(c1, t1.y, fmaf(c2, t2.y, a[0].y)));
    const float2 m2 = make_float2(fmaf(c2, t1.x, fmaf(c1, t2.x, a[0].x)), fmaf(c2, t1.y, fmaf(c1, t2.y, a[0].y)));
    const float2 n1 = make_float2(fmaf(s1, t3.x, s2 * t4.x), fmaf(s1, t3.y, s2 * t4.y));
    const float2 n2 = make_float2(fmaf(s2, t3.x, -s1 * t4.x), fmaf(s2, t3.y, -s1 * t4.y));
    const float2 r1 = mul_neg_i(n1), r2 = mul_neg_i(n2);
    a[0] = make_float2(a[0].x + t1.x + t2.x, a[0].y + t1.y + t2.y);
    a[1] = cadd(m1, r1);
    a[4] = csub(m1, r1);
    a[2] = cadd(m2, r2);
    a[3] = csub(m2, r2);
}

template <int R>
NSC_FFT_HD void dft(float2* a) {
    if (R == 8) dft8(a);
    else if (R == 9) dft9(a);
    else dft5(a);
}

// One butterfly of a Stockham pass over a 360-point signal. tw[m] = exp(-2 pi i m / 360).
template <int R, int NS>
NSC_FFT_HD void stockham_butterfly(const float2* x, float2* y, const float2* tw, int j) {
    constexpr int N = 360;
    constexpr int kS1 = N / (NS * R);
    const int blk = j / NS, k = j - blk * NS;
    float2 a[R];
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] = x[j + r * (N / R)];
    if (NS > 1) {
#pragma unroll
        for (int r = 1; r < R; ++r) a[r] = cmul(a[r], tw[k * r * kS1]);   // k*r*kS1 < 360
    }
    dft<R>(a);
    float2* out = y + blk * (NS * R) + k;
#pragma unroll
    for (int q = 0; q < R; ++q) out[q * NS] = a[q];
}

}  // namespace nsc

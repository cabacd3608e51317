// Per-point arithmetic of the projection (reference src/encoding/range_image.py:146-198),
// written once for the device kernels and for the host-side test hook
// (nsc_test_host_classify) that lets the CPU-only test suite compare the pixel assignment
// with the oracle. No product path evaluates it on the host.
//
// What the reference computes per point, and what is computed here instead:
//   range   = sqrt((x*x + y*y) + z*z) in float32, kept iff 1 <= range <= 80  (:159-162,:174)
//             -> s = (x*x + y*y) + z*z formed with the same three roundings; kept iff
//                s_lo <= s <= s_hi where the two float32 thresholds are the exact preimages of
//                the range test under the correctly rounded sqrt (host_tables.cu). NaN / Inf
//                coordinates make s NaN / Inf and fail the test, which is the finite filter
//                (:151-155). The min over s is the min over range (sqrt is monotone); the
//                sqrt is taken once per pixel after the scatter.
//   column  = floor(((atan2(y,x) + pi) mod 2pi) / 2pi * 360)                       (:166-167,:194)
//             -> octant-reduced odd polynomial of atan in column units, error 3e-7 rad; the
//                north star excuses points within 1e-5 rad of a column edge.
//   row     = floor((atan2(z, sqrt(x*x+y*y)) - el_min) / (el_max - el_min) * E), clipped (:170,:186)
//             -> u = z / rho clamped to the field of view, odd polynomial of atan scaled to
//                row units (kRowPoly), or a binary search of z|z| against tan^2 thresholds
//                times rho^2 (kRowSearch, any field of view). Same 1e-5 rad excuse.
#pragma once
#include "nsc_internal.h"

#include <cmath>
#include <cstring>

#if defined(__CUDACC__)
#define NSC_HD __host__ __device__ __forceinline__
#else
#define NSC_HD inline
#endif

namespace nsc {

// float32 multiply / add with exactly one rounding each (no FMA contraction).
NSC_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}
NSC_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
// MUFU.RCP / MUFU.RSQ on the device: one special-function op each, relative error <= 2^-22.
NSC_HD float rcp_fast(float a) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
#else
    return 1.0f / a;
#endif
}
NSC_HD float rsqrt_fast(float a) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
#else
    return 1.0f / sqrtf(a);
#endif
}
NSC_HD uint32_t f2u(float a) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(a);
#else
    uint32_t u;
    memcpy(&u, &a, 4);
    return u;
#endif
}


// floor of a float in [0, 2^22) as an integer, without the conversion pipe: adding 2^23 with
// round-down leaves the integer part in the low mantissa bits.
NSC_HD uint32_t floor_bits(float v) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(__fadd_rd(v, 8388608.0f)) - 0x4B000000u;
#else
    return (uint32_t)floorf(v);
#endif
}

// Column in [0, 360]; 360 only for azimuth == 2*pi exactly, which the reference wraps to 0.
template <typename P>
NSC_HD uint32_t column_of(float x, float y, const P& dp) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(fmaxf(ax, ay), 1e-30f);
    const float mn = fminf(ax, ay);
    const float t = mn * rcp_fast(mx);
    const float t2 = t * t;
    float p = dp.col_c[kColTerms - 1];
#pragma unroll
    for (int i = kColTerms - 2; i >= 0; --i) p = fmaf(p, t2, dp.col_c[i]);
    float r = p * t;                                  // [0, 45] column units
    if (ay > ax) r = 90.0f - r;                       // |y| > |x|: reflect about the bisector
    if ((int32_t)f2u(x) < 0) r = 180.0f - r;      // sign BIT of x: atan2(+-0, -0) = +-pi
    r = copysignf(r, y);                              // sign bit of y
    return floor_bits(r + 180.0f);
}

template <typename P>
NSC_HD uint32_t row_of(float z, float rho2, const P& dp, int row_mode) {
    if (row_mode == kRowPoly) {
        float u = z * rsqrt_fast(rho2);                // rho2 == 0 -> +-inf, clamped below
        u = fminf(fmaxf(u, dp.u_lo), dp.u_hi);
        const float u2 = u * u;
        float p = dp.row_p[kRowTerms - 1];
#pragma unroll
        for (int i = kRowTerms - 2; i >= 0; --i) p = fmaf(p, u2, dp.row_p[i]);
        return floor_bits(fmaf(p, u, dp.row_off));    // in (0, E) by construction of u_lo/u_hi
    }
    const float q = z * fabsf(z);
    int row = 0;
#pragma unroll
    for (int step = NSC_MAX_ELEVATION / 2; step > 0; step >>= 1) {
        const int k = row + step;
        if (k < dp.E && q >= dp.row_c[k] * rho2) row = k;
    }
    return (uint32_t)row;
}

// One point -> (keep, pixel index into the kPitch-wide min image, bits of s).
template <typename P>
NSC_HD bool classify(float x, float y, float z, const P& dp, int row_mode, uint32_t& pix,
                     uint32_t& sbits) {
    const float xx = mul_rn(x, x), yy = mul_rn(y, y), zz = mul_rn(z, z);
    const float rho2 = add_rn(xx, yy);
    const float s = add_rn(rho2, zz);
    const bool keep = (s >= dp.s_lo) && (s <= dp.s_hi);   // false for NaN / Inf
    const uint32_t col = column_of(x, y, dp);
    const uint32_t row = row_of(z, rho2, dp, row_mode);
    pix = row * (uint32_t)kPitch + col;
    sbits = f2u(s);
    return keep;
}

}  // namespace nsc

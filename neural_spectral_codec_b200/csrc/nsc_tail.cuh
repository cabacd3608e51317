// Shared-memory tail of the encoder: everything after the per-pixel min scatter.
//   validity masks -> circular linear hole interpolation + empty-row fill
//   (reference src/encoding/range_image.py:15-89) -> optional row pooling
//   (spectral_encoder.py:171-176) -> 360-point FFT magnitude per row (:180-186) ->
//   contiguous-frequency bin sums (:118-158) -> L1 normalisation (:197-202).
// All functions are CTA-collective: every thread of the block must call them.
#pragma once
#include "nsc_fft.cuh"
#include "nsc_internal.h"

namespace nsc {

#ifndef NSC_THREADS
#define NSC_THREADS 512
#endif
#ifndef NSC_MIN_BLOCKS
#define NSC_MIN_BLOCKS 2
#endif
constexpr int kThreads = NSC_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kMinBlocks = NSC_MIN_BLOCKS;   // resident CTAs per SM the kernels are compiled for

// Per-thread cp.async ring (LDGSTS feed): kCpDepth stages of kCpPts 16-byte points per thread.
#ifndef NSC_CP_PTS
#define NSC_CP_PTS 2
#endif
#ifndef NSC_CP_DEPTH
#define NSC_CP_DEPTH 4
#endif
constexpr int kCpPts = NSC_CP_PTS;
constexpr int kCpDepth = NSC_CP_DEPTH;
constexpr int kCpRingBytes = kCpDepth * kCpPts * kThreads * 16;   // 65 536 B per CTA

// Shared-memory carve-up, identical on host (size) and device (pointers). With a ring the two
// FFT buffers alias it: the ring is idle (fully consumed) while the tail runs.
struct SmemLayout {
    int img_off, tw_off, fa_off, fb_off, hist_off, mask_off, nvalid_off, src_off, red_off;
    int ring_off, total;
    int n_sig;  // complex FFTs per batch
    __host__ __device__ SmemLayout(int rows, int T, int n_bins, int ring_bytes = 0) {
        const bool ring = ring_bytes > 0;
        int o = 0;
        auto take = [&o](int bytes) { int r = o; o += (bytes + 127) & ~127; return r; };
        n_sig = (T + 1) / 2 < kMaxSignals ? (T + 1) / 2 : kMaxSignals;
        img_off = take(rows * kPitch * 4);
        tw_off = take(kAz * 8);
        // The second FFT buffer may reuse the image: once the (single) batch of signals has been
        // loaded from the image into fa, the image is dead.
        const int sig_bytes = n_sig * kAz * 8;
        const bool fb_on_img = (T + 1) / 2 <= kMaxSignals && rows * kPitch * 4 >= sig_bytes;
        const int scratch = fb_on_img ? sig_bytes : 2 * sig_bytes;
        ring_off = take(ring && ring_bytes > scratch ? ring_bytes : scratch);
        fa_off = ring_off;
        fb_off = fb_on_img ? img_off : ring_off + sig_bytes;
        hist_off = take(T * n_bins * 4);
        mask_off = take(rows * kMaskWords * 4);
        nvalid_off = take(rows * 4);
        src_off = take(rows * 4);
        red_off = take(kWarps * 8 + 16);
        total = o;
    }
};

struct TailSmem {
    float* img;        // rows x kPitch
    float2* tw;        // exp(-2 pi i m / 360)
    float2* fa;
    float2* fb;
    float* hist;
    uint32_t* mask;    // rows x kMaskWords validity bits (value > 0)
    int* nvalid;
    int* src;          // row r of the filled image is stored row src[r]
    double* red;
    __device__ TailSmem(unsigned char* base, const SmemLayout& L)
        : img((float*)(base + L.img_off)), tw((float2*)(base + L.tw_off)),
          fa((float2*)(base + L.fa_off)), fb((float2*)(base + L.fb_off)),
          hist((float*)(base + L.hist_off)), mask((uint32_t*)(base + L.mask_off)),
          nvalid((int*)(base + L.nvalid_off)), src((int*)(base + L.src_off)),
          red((double*)(base + L.red_off)) {}
};

__device__ __forceinline__ void init_twiddles(float2* tw) {
    for (int m = threadIdx.x; m < kAz; m += kThreads) {
        float s, c;
        sincospif((float)m * (1.0f / 180.0f), &s, &c);   // angle = 2 pi m / 360
        tw[m] = make_float2(c, -s);
    }
}

// mask / nvalid from the float image (valid <=> value > 0, range_image.py:35).
__device__ __forceinline__ void build_masks(const TailSmem& S, int rows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < rows; r += kWarps) {
        int cnt = 0;
        for (int w = 0; w < kMaskWords; ++w) {
            const int c = w * 32 + lane;
            const bool valid = (c < kAz) && (S.img[r * kPitch + c] > 0.0f);
            const uint32_t m = __ballot_sync(0xffffffffu, valid);
            cnt += __popc(m);
            if (lane == 0) S.mask[r * kMaskWords + w] = m;
        }
        if (lane == 0) S.nvalid[r] = cnt;
    }
}

// Nearest valid column strictly left / right of x on the circular row; the returned position
// is unwrapped (left in (x-360, x), right in (x, x+360)), as np.interp sees it on the tiled
// abscissa (range_image.py:55-64). Requires at least one valid bit in the row.
__device__ __forceinline__ int prev_valid(const uint32_t* m, int x) {
    int w = x >> 5, off = 0;
    uint32_t bits = m[w] & ((1u << (x & 31)) - 1u);
    while (bits == 0) {
        if (--w < 0) { w = kMaskWords - 1; off -= kAz; }
        bits = m[w];
    }
    return off + w * 32 + 31 - __clz(bits);
}
__device__ __forceinline__ int next_valid(const uint32_t* m, int x) {
    int w = x >> 5, off = 0;
    uint32_t bits = m[w] & ~((2u << (x & 31)) - 1u);
    while (bits == 0) {
        if (++w >= kMaskWords) { w = 0; off += kAz; }
        bits = m[w];
    }
    return off + w * 32 + __ffs(bits) - 1;
}

// interpolate_range_image(img, 'linear') in place (range_image.py:33-64 and :77-87).
// Pass 1 writes only hole pixels and reads only valid ones, so it is race-free in place.
// Pass 2 is expressed as a row indirection src[] instead of copying rows.
__device__ __forceinline__ void interpolate_and_fill(const TailSmem& S, int rows, bool enabled) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (enabled) {
        for (int r = warp; r < rows; r += kWarps) {
            const int nv = S.nvalid[r];
            if (nv == 0 || nv == kAz) continue;
            const uint32_t* m = S.mask + r * kMaskWords;
            float* row = S.img + r * kPitch;
            for (int x = lane; x < kAz; x += 32) {
                if ((m[x >> 5] >> (x & 31)) & 1u) continue;
                const int xl = prev_valid(m, x), xr = next_valid(m, x);
                const double fl = (double)row[xl < 0 ? xl + kAz : xl];
                const double fr = (double)row[xr >= kAz ? xr - kAz : xr];
                // np.interp: slope = (fp[j+1]-fp[j])/(xp[j+1]-xp[j]); slope*(x-xp[j]) + fp[j], float64
                const double slope = __ddiv_rn(__dsub_rn(fr, fl), (double)(xr - xl));
                row[x] = __double2float_rn(__dadd_rn(__dmul_rn(slope, (double)(x - xl)), fl));
            }
        }
    }
    __syncthreads();
    // Rows with no pixel > 0 copy the nearest filled row below, leading ones the first non-empty
    // row above; an all-empty image stays zero (sequential in-place semantics of :77-87).
    if (threadIdx.x < rows) {
        const int r = threadIdx.x;
        int s = r;
        if (enabled && S.nvalid[r] == 0) {
            int k = r - 1;
            while (k >= 0 && S.nvalid[k] == 0) --k;
            if (k < 0) {
                k = r + 1;
                while (k < rows && S.nvalid[k] == 0) ++k;
            }
            if (k >= 0 && k < rows) s = k;
        }
        S.src[r] = s;
    }
    __syncthreads();
}

// Value of target row i at column n: the stored row, or the mean of its source rows when the
// image height differs from T (adaptive_avg_pool2d on (1,1,rows,360) -> (T,360)).
__device__ __forceinline__ float pooled_value(const TailSmem& S, int rows, int T, int i, int n) {
    if (rows == T) return S.img[S.src[i] * kPitch + n];
    const int r0 = (i * rows) / T, r1 = ((i + 1) * rows + T - 1) / T;
    float sum = 0.0f;
    for (int r = r0; r < r1; ++r) sum += S.img[S.src[r] * kPitch + n];
    return sum / (float)(r1 - r0);
}

// One Stockham pass of radix R over n_sig complex signals of length 360: one thread per
// butterfly (nsc_fft.cuh), R inputs and outputs in registers.
template <int R, int NS>
__device__ __forceinline__ void fft_pass(const float2* __restrict__ x, float2* __restrict__ y,
                                         const float2* __restrict__ tw, int n_sig) {
    constexpr int kBfly = kAz / R;
    for (int t = threadIdx.x; t < n_sig * kBfly; t += kThreads) {
        const int g = t / kBfly, j = t - g * kBfly;
        stockham_butterfly<R, NS>(x + g * kAz, y + g * kAz, tw, j);
    }
}

// Rows [2*g0, 2*(g0+n)) of the (pooled) image -> complex signals -> spectra -> bin sums.
template <typename P>
__device__ __forceinline__ void spectrum_and_bins(const TailSmem& S, const P& dp, int rows) {
    const int T = dp.T, nb = dp.n_bins;
    const int n_sig_total = (T + 1) / 2;
    const int cap = n_sig_total < kMaxSignals ? n_sig_total : kMaxSignals;
    for (int g0 = 0; g0 < n_sig_total; g0 += cap) {
        const int n_sig = min(cap, n_sig_total - g0);
        for (int t = threadIdx.x; t < n_sig * kAz; t += kThreads) {
            const int g = t / kAz, n = t - g * kAz;
            const int ra = 2 * (g0 + g), rb = ra + 1;
            const float a = pooled_value(S, rows, T, ra, n);
            const float b = rb < T ? pooled_value(S, rows, T, rb, n) : 0.0f;
            S.fa[t] = make_float2(a, b);
        }
        __syncthreads();
        fft_pass<8, 1>(S.fa, S.fb, S.tw, n_sig);
        __syncthreads();
        fft_pass<9, 8>(S.fb, S.fa, S.tw, n_sig);
        __syncthreads();
        fft_pass<5, 72>(S.fa, S.fb, S.tw, n_sig);
        __syncthreads();
        // Z = FFT(a + i b): A[k] = (Z[k] + conj Z[-k]) / 2, B[k] = (Z[k] - conj Z[-k]) / 2i.
        for (int t = threadIdx.x; t < n_sig * nb; t += kThreads) {
            const int g = t / nb, b = t - g * nb;
            const float2* z = S.fb + g * kAz;
            float ha = 0.0f, hb = 0.0f;
            const int k1 = dp.bin_start[b + 1];
            for (int k = dp.bin_start[b]; k < k1; ++k) {   // ascending k: scatter_add_ order
                const float2 p = z[k], m = z[k == 0 ? 0 : kAz - k];
                const float ar = p.x + m.x, ai = p.y - m.y;
                const float br = p.y + m.y, bi = p.x - m.x;
                ha += 0.5f * __fsqrt_rn(fmaf(ar, ar, ai * ai));
                hb += 0.5f * __fsqrt_rn(fmaf(br, br, bi * bi));
            }
            const int ra = 2 * (g0 + g);
            S.hist[ra * nb + b] = ha;
            if (ra + 1 < T) S.hist[(ra + 1) * nb + b] = hb;
        }
        __syncthreads();
    }
}

// h / (sum h + eps), or the uniform descriptor when sum h <= eps (spectral_encoder.py:197-202).
struct PeerOut {
    float* ptr[NSC_MAX_PEERS];
    int n;
    long long row0;
};

template <typename P>
__device__ __forceinline__ void normalise_and_store(const TailSmem& S, const P& dp, float* out,
                                                    const PeerOut& peers, long long peer_row) {
    const int D = dp.T * dp.n_bins;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int i = threadIdx.x; i < D; i += kThreads) acc += (double)S.hist[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) S.red[warp] = acc;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) tot += S.red[w];
    const float total = (float)tot;
    const bool ok = total > dp.eps;
    const float denom = __fadd_rn(total, dp.eps);
    for (int i = threadIdx.x; i < D; i += kThreads) {
        const float v = ok ? __fdiv_rn(S.hist[i], denom) : dp.uniform;
        if (out) out[i] = v;
        for (int p = 0; p < peers.n; ++p) peers.ptr[p][peer_row * D + i] = v;
    }
    __syncthreads();   // hist / red are reused by the next scan
}

}  // namespace nsc

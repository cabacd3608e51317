// Fused encode kernel: points -> min-range image (shared memory) -> interpolation -> spectrum
// -> descriptor. One persistent CTA per resident slot; each CTA pulls whole scans from a work
// counter, so the range image, its interpolation and the spectrum never leave shared memory:
// HBM traffic is the 16 B per point and the descriptor.
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "nsc_point.h"
#include "nsc_tail.cuh"

namespace nsc {

namespace {

constexpr int kUnroll = 4;   // independent 16-byte loads in flight per thread (LDG feed)
constexpr int kDefaultFeed = 0;   // kFeedLdg until the TMA ring is measured faster

struct EncodeArgs {
    const float* points;
    const long long* offsets;
    long long origin;
    int n_scans;
    float* out;          // n_scans x D, may be null (projection only)
    float* img_out;      // n_scans x E x 360, may be null
    int stage;
    unsigned* counter;
    PeerOut peers;
};

__device__ __forceinline__ void scatter_min(uint32_t* img, bool keep, uint32_t pix, uint32_t sbits) {
    // A plain read first: after the first few hits most points are not a new minimum, and a
    // stale read can only cause a redundant atomic, never a missed one.
    if (keep && sbits < img[pix]) atomicMin(img + pix, sbits);
}

enum Feed { kFeedLdg = 0, kFeedTma = 1 };

// ---- mbarrier / bulk-copy (TMA) primitives ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
// global -> shared bulk copy (SASS UBLKCP), completion counted in bytes on an mbarrier.
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                              uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

template <int STRIDE, int ROWMODE, int FEED>
__global__ void __launch_bounds__(kThreads, 2)
encode_points_kernel(const __grid_constant__ EncodeArgs a, const __grid_constant__ DeviceParams dp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_scan;
    const SmemLayout L(dp.E, dp.T, dp.n_bins, FEED == kFeedTma);
    const TailSmem S(smem_raw, L);
    uint32_t* img = reinterpret_cast<uint32_t*>(S.img);
    const int tid = threadIdx.x;
    const int n_pix = dp.E * kPitch;
    const int D = dp.T * dp.n_bins;

    init_twiddles(S.tw);

    // TMA feed: every warp owns kStages x kStageBytes of the ring and one mbarrier per stage.
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t ring0 = smem_u32(smem_raw + L.ring_off) + warp * (kStages * kStageBytes);
    const uint32_t bar0 = smem_u32(smem_raw + L.bar_off) + warp * (kStages * 8);
    const float4* ring_w = reinterpret_cast<const float4*>(smem_raw + L.ring_off) +
                           warp * (kStages * kStagePts);
    uint32_t g_issue = 0, g_cons = 0;   // chunks issued / consumed by this warp since launch
    if (FEED == kFeedTma) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < kStages; ++s) mbar_init(bar0 + 8 * s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }

    for (;;) {
        __syncthreads();
        if (tid == 0) s_scan = (int)atomicAdd(a.counter, 1u);
        for (int i = tid; i < n_pix; i += kThreads) img[i] = kInfBits;
        __syncthreads();
        const int scan = s_scan;
        if (scan >= a.n_scans) break;

        const long long beg = a.offsets[scan] - a.origin;
        const int n = (int)(a.offsets[scan + 1] - a.offsets[scan]);

        if (FEED == kFeedTma) {
            // Chunk c of the scan (kStagePts points) belongs to warp c % kWarps. Each warp keeps
            // kStages bulk copies in flight into its own ring; lane 0 arms the stage's mbarrier
            // with the byte count and issues the copy, all lanes wait on the phase parity.
            const float4* p4 = reinterpret_cast<const float4*>(a.points) + beg;
            const int n_chunks = (n + kStagePts - 1) / kStagePts;
            const int mine = n_chunks > warp ? (n_chunks - warp + kWarps - 1) / kWarps : 0;
            // The ring doubles as FFT scratch in the tail: order those generic-proxy writes
            // before the async-proxy writes of the copies issued below.
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            auto issue = [&](int j) {
                const int c = warp + j * kWarps;
                const int pts = min(kStagePts, n - c * kStagePts);
                const uint32_t slot = g_issue % kStages;
                if (lane == 0) {
                    mbar_expect_tx(bar0 + 8 * slot, (uint32_t)pts * 16u);
                    bulk_copy_g2s(ring0 + slot * kStageBytes, p4 + (long long)c * kStagePts,
                                  (uint32_t)pts * 16u, bar0 + 8 * slot);
                }
                ++g_issue;
            };
            int issued = 0;
            for (; issued < mine && issued < kStages; ++issued) issue(issued);
            for (int j = 0; j < mine; ++j) {
                const uint32_t slot = g_cons % kStages, parity = (g_cons / kStages) & 1u;
                mbar_wait(bar0 + 8 * slot, parity);
                ++g_cons;
                const int c = warp + j * kWarps;
                const int valid = min(kStagePts, n - c * kStagePts);
                const float4* st = ring_w + slot * kStagePts + 3 * lane;
                const float4 v0 = st[0], v1 = st[1], v2 = st[2];
                uint32_t p0, p1, p2, s0, s1, s2;
                bool k0 = classify(v0.x, v0.y, v0.z, dp, ROWMODE, p0, s0) && (3 * lane + 0 < valid);
                bool k1 = classify(v1.x, v1.y, v1.z, dp, ROWMODE, p1, s1) && (3 * lane + 1 < valid);
                bool k2 = classify(v2.x, v2.y, v2.z, dp, ROWMODE, p2, s2) && (3 * lane + 2 < valid);
                // consecutive points of a spinning sensor mostly share a pixel: fold them first
                if (k0 && k1 && p0 == p1) { s1 = min(s0, s1); k0 = false; }
                if (k1 && k2 && p1 == p2) { s2 = min(s1, s2); k1 = false; }
                scatter_min(img, k0, p0, s0);
                scatter_min(img, k1, p1, s1);
                scatter_min(img, k2, p2, s2);
                __syncwarp();   // every lane has read the stage before it is refilled
                if (issued < mine) issue(issued++);
            }
        } else if (STRIDE == 4) {
            const float4* p4 = reinterpret_cast<const float4*>(a.points) + beg;
            for (int base = 0; base < n; base += kThreads * kUnroll) {
                float4 v[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int i = base + u * kThreads + tid;
                    v[u] = i < n ? __ldcs(p4 + i) : make_float4(NAN, NAN, NAN, 0.0f);
                }
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    uint32_t pix, sb;
                    const bool keep = classify(v[u].x, v[u].y, v[u].z, dp, ROWMODE, pix, sb);
                    scatter_min(img, keep, pix, sb);
                }
            }
        } else {
            const float* p = a.points + beg * 3;
            for (int base = 0; base < n; base += kThreads * kUnroll) {
                float x[kUnroll], y[kUnroll], z[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int i = base + u * kThreads + tid;
                    const bool in = i < n;
                    x[u] = in ? __ldcs(p + 3 * (long long)i) : NAN;
                    y[u] = in ? __ldcs(p + 3 * (long long)i + 1) : NAN;
                    z[u] = in ? __ldcs(p + 3 * (long long)i + 2) : NAN;
                }
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    uint32_t pix, sb;
                    const bool keep = classify(x[u], y[u], z[u], dp, ROWMODE, pix, sb);
                    scatter_min(img, keep, pix, sb);
                }
            }
        }
        __syncthreads();

        // bits of min s -> range = sqrt_rn(s); empty -> 0 (range_image.py:162,:214). Column 360
        // (azimuth exactly 2 pi) belongs to column 0.
        for (int i = tid; i < dp.E * kAz; i += kThreads) {
            const int r = i / kAz, c = i - r * kAz;
            uint32_t b = img[r * kPitch + c];
            if (c == 0) b = min(b, img[r * kPitch + kAz]);
            const float v = (b == kInfBits) ? 0.0f : __fsqrt_rn(__uint_as_float(b));
            S.img[r * kPitch + c] = v;
            if (a.img_out && a.stage == NSC_STAGE_PROJECTED)
                a.img_out[(long long)scan * dp.E * kAz + i] = v;
        }
        __syncthreads();
        build_masks(S, dp.E);
        __syncthreads();
        interpolate_and_fill(S, dp.E, dp.interpolate != 0);
        if (a.img_out && a.stage == NSC_STAGE_INTERPOLATED) {
            for (int i = tid; i < dp.E * kAz; i += kThreads) {
                const int r = i / kAz, c = i - r * kAz;
                a.img_out[(long long)scan * dp.E * kAz + i] = S.img[S.src[r] * kPitch + c];
            }
        }
        if (a.out || a.peers.n > 0) {
            spectrum_and_bins(S, dp, dp.E);
            normalise_and_store(S, dp, a.out ? a.out + (long long)scan * D : nullptr, a.peers,
                                a.peers.row0 + scan);
        }
    }
}

// Range images in, descriptors out: SpectralEncoder.forward / encode_range_image
// (spectral_encoder.py:160-204, :231-261). No projection, no interpolation.
__global__ void __launch_bounds__(kThreads, 2)
encode_images_kernel(const float* __restrict__ images, int n_images, int rows,
                     const __grid_constant__ DeviceParams dp, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SmemLayout L(rows, dp.T, dp.n_bins);
    const TailSmem S(smem_raw, L);
    const int D = dp.T * dp.n_bins;
    init_twiddles(S.tw);
    if (threadIdx.x < rows) S.src[threadIdx.x] = threadIdx.x;
    PeerOut none;
    none.n = 0;
    none.row0 = 0;
    for (int im = blockIdx.x; im < n_images; im += gridDim.x) {
        __syncthreads();
        const float* src = images + (long long)im * rows * kAz;
        for (int i = threadIdx.x; i < rows * kAz; i += kThreads) {
            const int r = i / kAz, c = i - r * kAz;
            S.img[r * kPitch + c] = src[i];
        }
        __syncthreads();
        spectrum_and_bins(S, dp, rows);
        normalise_and_store(S, dp, out + (long long)im * D, none, 0);
    }
}

// interpolate_range_image on device images (range_image.py:15-89).
__global__ void __launch_bounds__(kThreads, 2)
interpolate_kernel(const float* __restrict__ in, int n_images, int rows, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SmemLayout L(rows, 1, 1);
    const TailSmem S(smem_raw, L);
    for (int im = blockIdx.x; im < n_images; im += gridDim.x) {
        __syncthreads();
        const float* src = in + (long long)im * rows * kAz;
        for (int i = threadIdx.x; i < rows * kAz; i += kThreads) {
            const int r = i / kAz, c = i - r * kAz;
            S.img[r * kPitch + c] = src[i];
        }
        __syncthreads();
        build_masks(S, rows);
        __syncthreads();
        interpolate_and_fill(S, rows, true);
        float* dst = out + (long long)im * rows * kAz;
        for (int i = threadIdx.x; i < rows * kAz; i += kThreads) {
            const int r = i / kAz, c = i - r * kAz;
            dst[i] = S.img[S.src[r] * kPitch + c];
        }
    }
}

struct DeviceInfo {
    int sms = 0;
    int max_smem_optin = 0;
};

int device_info(DeviceInfo* out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return record_cuda(e);
    static std::mutex mu;
    static DeviceInfo cache[64];
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && cache[dev].sms > 0) { *out = cache[dev]; return NSC_OK; }
    DeviceInfo di;
    e = cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return record_cuda(e);
    e = cudaDeviceGetAttribute(&di.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return record_cuda(e);
    if (dev >= 0 && dev < 64) cache[dev] = di;
    *out = di;
    return NSC_OK;
}

template <typename K>
int configure(K kernel, int smem, const DeviceInfo& di, int* blocks_per_sm) {
    if (smem > di.max_smem_optin) return NSC_ERR_BAD_PARAMS;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return record_cuda(e);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kernel, kThreads, smem);
    if (e != cudaSuccess) return record_cuda(e);
    if (*blocks_per_sm < 1) return NSC_ERR_BAD_PARAMS;
    return NSC_OK;
}

}  // namespace

size_t workspace_bytes_for(int n_scans, int E) {
    (void)n_scans;
    (void)E;
    return 256;   // work counter (+ padding)
}

int launch_encode(const float* d_points, int stride, const long long* d_offsets, long long origin,
                  int n_scans, const DeviceParams& dp, float* d_out, float* d_img_out, int stage,
                  float* const* d_peer_out, int n_peers, long long peer_row0,
                  unsigned* d_workspace, cudaStream_t stream) {
    if (n_scans == 0) return NSC_OK;
    DeviceInfo di;
    int st = device_info(&di);
    if (st != NSC_OK) return st;
    EncodeArgs a;
    a.points = d_points;
    a.offsets = d_offsets;
    a.origin = origin;
    a.n_scans = n_scans;
    a.out = d_out;
    a.img_out = d_img_out;
    a.stage = stage;
    a.counter = d_workspace;
    a.peers.n = n_peers;
    a.peers.row0 = peer_row0;
    for (int i = 0; i < NSC_MAX_PEERS; ++i) a.peers.ptr[i] = i < n_peers ? d_peer_out[i] : nullptr;

    // Feed of the point pass: per-warp TMA bulk-copy ring (4-float points) or plain vector
    // loads (3-float points are not 16-byte granular). NSC_FEED=ldg|tma overrides for A/B runs.
    static const int feed_override = [] {
        const char* e = getenv("NSC_FEED");
        if (!e) return -1;
        return strcmp(e, "tma") == 0 ? (int)kFeedTma : strcmp(e, "ldg") == 0 ? (int)kFeedLdg : -1;
    }();
    int feed = kDefaultFeed;
    if (feed_override >= 0) feed = feed_override;
    if (stride != 4) feed = kFeedLdg;
    const SmemLayout L(dp.E, dp.T, dp.n_bins, feed == kFeedTma);
    void (*kernel)(const EncodeArgs, const DeviceParams) = nullptr;
    if (feed == kFeedTma) {
        kernel = dp.row_mode == kRowPoly ? encode_points_kernel<4, kRowPoly, kFeedTma>
                                         : encode_points_kernel<4, kRowSearch, kFeedTma>;
    } else if (stride == 4) {
        kernel = dp.row_mode == kRowPoly ? encode_points_kernel<4, kRowPoly, kFeedLdg>
                                         : encode_points_kernel<4, kRowSearch, kFeedLdg>;
    } else {
        kernel = dp.row_mode == kRowPoly ? encode_points_kernel<3, kRowPoly, kFeedLdg>
                                         : encode_points_kernel<3, kRowSearch, kFeedLdg>;
    }
    int per_sm = 0;
    st = configure(kernel, L.total, di, &per_sm);
    if (st != NSC_OK) return st;
    cudaError_t e = cudaMemsetAsync(d_workspace, 0, 2 * sizeof(unsigned), stream);
    if (e != cudaSuccess) return record_cuda(e);
    int grid = di.sms * per_sm;
    if (grid > n_scans) grid = n_scans;
    kernel<<<grid, kThreads, L.total, stream>>>(a, dp);
    return record_cuda(cudaGetLastError());
}

int launch_encode_images(const float* d_images, int n_images, int rows, const DeviceParams& dp,
                         float* d_out, cudaStream_t stream) {
    if (n_images == 0) return NSC_OK;
    DeviceInfo di;
    int st = device_info(&di);
    if (st != NSC_OK) return st;
    const SmemLayout L(rows, dp.T, dp.n_bins);
    int per_sm = 0;
    st = configure(encode_images_kernel, L.total, di, &per_sm);
    if (st != NSC_OK) return st;
    int grid = di.sms * per_sm;
    if (grid > n_images) grid = n_images;
    encode_images_kernel<<<grid, kThreads, L.total, stream>>>(d_images, n_images, rows, dp, d_out);
    return record_cuda(cudaGetLastError());
}

int launch_interpolate(const float* d_in, int n_images, int rows, float* d_out, cudaStream_t stream) {
    if (n_images == 0) return NSC_OK;
    DeviceInfo di;
    int st = device_info(&di);
    if (st != NSC_OK) return st;
    const SmemLayout L(rows, 1, 1);
    int per_sm = 0;
    st = configure(interpolate_kernel, L.total, di, &per_sm);
    if (st != NSC_OK) return st;
    int grid = di.sms * per_sm;
    if (grid > n_images) grid = n_images;
    interpolate_kernel<<<grid, kThreads, L.total, stream>>>(d_in, n_images, rows, d_out);
    return record_cuda(cudaGetLastError());
}

}  // namespace nsc

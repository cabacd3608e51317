// Feed ceiling of one SM-resident stream per CTA (tuning aid, not product code): how fast can
// 148 CTAs each pull their own contiguous ~1.9 MB "scan" of 16-byte points through
//   (a) plain LDG.128 into registers,
//   (b) a per-thread cp.async (LDGSTS) ring in shared memory, consumed with LDS.128,
//   (c) whole-stage cp.async.bulk (TMA, UBLKCP) copies by one producer thread + mbarriers,
// with no arithmetic at all. The encode kernel cannot beat the best of these.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a feedbw.cu -o _build/feedbw && _build/feedbw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                 ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// (a) LDG: chunk c of `chunk` points per CTA round-robin, U loads in flight per thread
template <int NT, int U>
__global__ void __launch_bounds__(NT) ldg_chunks(const float4* __restrict__ p, size_t n, size_t chunk, float* out) {
    float acc = 0.f;
    for (size_t c = blockIdx.x; c * chunk < n; c += gridDim.x) {
        const float4* q = p + c * chunk;
        const size_t m = (c + 1) * chunk <= n ? chunk : n - c * chunk;
        for (size_t base = 0; base < m; base += (size_t)NT * U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t i = base + u * NT + threadIdx.x;
                v[u] = i < m ? __ldcs(q + i) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

// (b) per-thread cp.async ring: NS streaming threads of an NT-thread CTA (the others idle), PTS
// points per stage per thread, D stages; the ring restarts at every chunk (like the round-1 kernel)
// or flows across chunks (CONT, like the warp-specialised kernel).
template <int NT, int NS, int PTS, int D, bool CONT>
__global__ void __launch_bounds__(NT) ldgsts_ring(const float4* __restrict__ p, size_t n, size_t chunk, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    if (threadIdx.x >= NS) return;
    const int tid = threadIdx.x;
    constexpr int SP = PTS * NS, SLOT = SP * 16;
    const uint32_t ring_t = smem_u32(smem) + tid * 16;
    float acc = 0.f;
    const size_t n_chunks = (n + chunk - 1) / chunk;
    if (!CONT) {
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            const float4* q = p + c * chunk + tid;
            const int m = (int)((c + 1) * chunk <= n ? chunk : n - c * chunk);
            const int n_iter = (m + SP - 1) / SP;
            auto issue = [&](int it) {
                if (it < n_iter)
#pragma unroll
                    for (int u = 0; u < PTS; ++u) {
                        const int i = it * SP + u * NS;
                        if (i + tid < m) cp_async16(ring_t + (it % D) * SLOT + u * (NS * 16), q + i);
                    }
                cp_commit();
            };
            for (int d = 0; d < D - 1; ++d) issue(d);
            for (int it = 0; it < n_iter; ++it) {
                issue(it + D - 1);
                cp_wait<D - 1>();
#pragma unroll
                for (int u = 0; u < PTS; ++u) {
                    const float4 v = lds128(ring_t + (it % D) * SLOT + u * (NS * 16));
                    if (it * SP + u * NS + tid < m) acc += v.x + v.y + v.z + v.w;
                }
            }
            cp_wait<0>();
        }
    } else {
        // one long stream: the CTA's chunks back to back, stage index running across them
        size_t c = blockIdx.x;
        const float4* q = p + c * chunk + tid;
        int m = c < n_chunks ? (int)((c + 1) * chunk <= n ? chunk : n - c * chunk) : 0;
        int n_iter = (m + SP - 1) / SP;
        int it_issue = 0;                 // next stage to issue within the issuing chunk
        const float4* qi = q; int mi = m, ni = n_iter; size_t ci = c;
        unsigned g_issue = 0, g_use = 0;
        auto issue = [&]() {
            while (it_issue >= ni && ci < n_chunks) {      // move the issue side to the next chunk
                ci += gridDim.x;
                qi = p + ci * chunk + tid;
                mi = ci < n_chunks ? (int)((ci + 1) * chunk <= n ? chunk : n - ci * chunk) : 0;
                ni = (mi + SP - 1) / SP;
                it_issue = 0;
            }
            if (ci < n_chunks) {
#pragma unroll
                for (int u = 0; u < PTS; ++u) {
                    const int i = it_issue * SP + u * NS;
                    if (i + tid < mi) cp_async16(ring_t + (g_issue % D) * SLOT + u * (NS * 16), qi + i);
                }
                ++it_issue;
            }
            ++g_issue;
            cp_commit();
        };
        for (int d = 0; d < D - 1; ++d) issue();
        while (c < n_chunks) {
            for (int it = 0; it < n_iter; ++it) {
                issue();
                cp_wait<D - 1>();
#pragma unroll
                for (int u = 0; u < PTS; ++u) {
                    const float4 v = lds128(ring_t + (g_use % D) * SLOT + u * (NS * 16));
                    if (it * SP + u * NS + tid < m) acc += v.x + v.y + v.z + v.w;
                }
                ++g_use;
            }
            c += gridDim.x;
            m = c < n_chunks ? (int)((c + 1) * chunk <= n ? chunk : n - c * chunk) : 0;
            n_iter = (m + SP - 1) / SP;
        }
        cp_wait<0>();
    }
    if (acc == 123.456f) out[0] = acc;
}

// (c) TMA bulk copies: thread NT-32 is the producer, the first NC threads consume; stage = SB bytes.
template <int NT, int NC, int SB, int D>
__global__ void __launch_bounds__(NT) bulk_ring(const float4* __restrict__ p, size_t n, size_t chunk, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const uint32_t ring = smem_u32(smem), bars = ring + D * SB;     // full[D], empty[D]
    constexpr int SPTS = SB / 16;
    if (tid == 0) {
        for (int s = 0; s < D; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (D + s), NC / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t n_chunks = (n + chunk - 1) / chunk;
    float acc = 0.f;
    unsigned g = 0;
    if (tid == NT - 32) {
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            const float4* q = p + c * chunk;
            const int m = (int)((c + 1) * chunk <= n ? chunk : n - c * chunk);
            for (int base = 0; base < m; base += SPTS, ++g) {
                const uint32_t slot = g % D;
                if (g >= D) mbar_wait(bars + 8 * (D + slot), ((g / D) + 1) & 1);
                const uint32_t bytes = (uint32_t)(m - base < SPTS ? m - base : SPTS) * 16u;
                mbar_expect_tx(bars + 8 * slot, bytes);
                bulk_g2s(ring + slot * SB, q + base, bytes, bars + 8 * slot);
            }
        }
    } else if (tid < NC) {
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            const int m = (int)((c + 1) * chunk <= n ? chunk : n - c * chunk);
            for (int base = 0; base < m; base += SPTS, ++g) {
                const uint32_t slot = g % D;
                mbar_wait(bars + 8 * slot, (g / D) & 1);
                for (int i = tid; i < SPTS; i += NC) {
                    const float4 v = lds128(ring + slot * SB + i * 16);
                    if (base + i < m) acc += v.x + v.y + v.z + v.w;
                }
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(bars + 8 * (D + slot));
            }
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

template <typename F>
static void run(const char* name, size_t bytes, F f, int iters = 12) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f(); f();
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-58s ERROR %s\n", name, cudaGetErrorString(e)); return; }
    double best = 1e30, sum = 0;
    for (int i = 0; i < iters; ++i) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
        sum += ms;
    }
    printf("%-58s best %7.1f GB/s   mean %7.1f GB/s\n", name, bytes / best / 1e6, bytes / (sum / iters) / 1e6);
}

template <typename K>
static void optin(K k, int smem) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); }

int main() {
    const size_t chunk = 120704;                      // points per "scan"
    const size_t n = chunk * 4541, bytes = n * 16;    // the benchmark batch: 8.77 GB
    float4* src; float* out;
    if (cudaMalloc(&src, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&out, 4);
    cudaMemset(src, 1, bytes);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("%d SMs, %zu chunks of %zu points (%.2f GB)\n", sms, n / chunk, chunk, bytes / 1e9);

    run("LDG 512 thr x2 CTA/SM, U=4 (64 B/thr)", bytes, [&] { ldg_chunks<512, 4><<<sms * 2, 512>>>(src, n, chunk, out); });
    run("LDG 512 thr x2 CTA/SM, U=8 (128 B/thr)", bytes, [&] { ldg_chunks<512, 8><<<sms * 2, 512>>>(src, n, chunk, out); });
    run("LDG 1024 thr x1 CTA/SM, U=4", bytes, [&] { ldg_chunks<1024, 4><<<sms, 1024>>>(src, n, chunk, out); });
    run("LDG 1024 thr x1 CTA/SM, U=8", bytes, [&] { ldg_chunks<1024, 8><<<sms, 1024>>>(src, n, chunk, out); });
    run("LDG 768 thr x1 CTA/SM, U=8", bytes, [&] { ldg_chunks<768, 8><<<sms, 768>>>(src, n, chunk, out); });

#define RING(NT, NS, PTS, D, CONT, CTAS)                                                              \
    {                                                                                                 \
        auto k = ldgsts_ring<NT, NS, PTS, D, CONT>;                                                   \
        const int sm = PTS * NS * 16 * D;                                                             \
        optin(k, sm);                                                                                 \
        run("LDGSTS ring NT=" #NT " NS=" #NS " PTS=" #PTS " D=" #D " cont=" #CONT " x" #CTAS, bytes,      \
            [&] { k<<<sms * CTAS, NT, sm>>>(src, n, chunk, out); });                                  \
    }
    RING(512, 512, 2, 4, false, 2)      // the round-1 kernel's feed
    RING(512, 512, 2, 4, true, 2)
    RING(1024, 768, 2, 5, false, 1)
    RING(1024, 768, 2, 5, true, 1)      // the warp-specialised kernel's feed
    RING(1024, 768, 2, 4, true, 1)
    RING(1024, 768, 2, 6, true, 1)
    RING(1024, 768, 4, 3, true, 1)
    RING(1024, 768, 1, 8, true, 1)
    RING(1024, 1024, 2, 5, true, 1)
    RING(1024, 1024, 2, 6, true, 1)
    RING(1024, 896, 2, 5, true, 1)
    RING(512, 384, 2, 5, true, 2)
    RING(256, 256, 2, 5, true, 4)
    RING(768, 768, 2, 5, true, 1)

#define BULK(NT, NC, SB, D, CTAS)                                                                     \
    {                                                                                                 \
        auto k = bulk_ring<NT, NC, SB, D>;                                                            \
        const int sm = SB * D + 16 * D + 64;                                                          \
        optin(k, sm);                                                                                 \
        run("TMA bulk NT=" #NT " NC=" #NC " stage=" #SB " D=" #D " x" #CTAS, bytes,                      \
            [&] { k<<<sms * CTAS, NT, sm>>>(src, n, chunk, out); });                                  \
    }
    BULK(1024, 768, 16384, 6, 1)
    BULK(1024, 768, 24576, 5, 1)
    BULK(1024, 768, 32768, 4, 1)
    BULK(1024, 768, 32768, 6, 1)
    BULK(1024, 768, 8192, 12, 1)
    BULK(512, 480, 16384, 4, 2)
    BULK(512, 480, 8192, 8, 2)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

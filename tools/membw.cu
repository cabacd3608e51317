// Read-only streaming bandwidth ceiling on this GPU (tuning aid, not product code):
// what a kernel that only LOADS 16-byte elements can reach, next to a device copy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a membw.cu -o _build/membw && _build/membw
#include <cstdio>
#include <cuda_runtime.h>

template <int U>
__global__ void __launch_bounds__(512, 2) read_kernel(const float4* __restrict__ p, size_t n, float* out) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldcs(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; i < n; i += stride) { float4 v = __ldcs(p + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) out[0] = acc;
}

// contiguous chunk per CTA (like one scan per CTA), U loads in flight per thread
template <int U>
__global__ void __launch_bounds__(512, 2) read_chunked(const float4* __restrict__ p, size_t n, size_t chunk, float* out) {
    float acc = 0.f;
    for (size_t c = blockIdx.x; c * chunk < n; c += gridDim.x) {
        const float4* q = p + c * chunk;
        size_t m = (c + 1) * chunk <= n ? chunk : n - c * chunk;
        for (size_t base = 0; base < m; base += 512 * U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { size_t i = base + u * 512 + threadIdx.x; v[u] = i < m ? __ldcs(q + i) : make_float4(0, 0, 0, 0); }
#pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

template <typename F>
static double time_ms(F f, int iters = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f(); cudaDeviceSynchronize();
    double best = 1e30;
    for (int i = 0; i < iters; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main() {
    const size_t bytes = (size_t)8 << 30, n = bytes / 16;
    float4 *src, *dst; float* out;
    cudaMalloc(&src, bytes); cudaMalloc(&dst, bytes); cudaMalloc(&out, 4);
    cudaMemset(src, 1, bytes);
    double ms = time_ms([&] { cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice); });
    printf("memcpy D2D            %8.1f GB/s (read+write bytes)\n", 2.0 * bytes / ms / 1e6);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int mult : {2, 4, 8}) {
        ms = time_ms([&] { read_kernel<4><<<sms * mult, 512>>>(src, n, out); });
        printf("read grid-stride U=4  x%d %8.1f GB/s\n", mult, bytes / ms / 1e6);
        ms = time_ms([&] { read_kernel<8><<<sms * mult, 512>>>(src, n, out); });
        printf("read grid-stride U=8  x%d %8.1f GB/s\n", mult, bytes / ms / 1e6);
    }
    ms = time_ms([&] { read_chunked<4><<<sms * 2, 512>>>(src, n, 120000, out); });
    printf("read 120k-pt chunks U=4    %8.1f GB/s\n", bytes / ms / 1e6);
    ms = time_ms([&] { read_chunked<8><<<sms * 2, 512>>>(src, n, 120000, out); });
    printf("read 120k-pt chunks U=8    %8.1f GB/s\n", bytes / ms / 1e6);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

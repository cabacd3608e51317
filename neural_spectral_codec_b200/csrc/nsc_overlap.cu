// Keyframe-gate geometry: voxel-set IoU of cloud pairs, the expensive
// criterion of the reference's keyframe gate (src/data/pose_utils.py:323-389 compute_overlap,
// called from src/keyframe/criteria.py:96-131). One CTA per pair (cloud A, cloud B, T_AB):
//   1. every point -> an int32 voxel coordinate triple, with the reference's arithmetic:
//      cloud A is transformed in float64 (transform_points stacks a float64 ones column, :121-129),
//      clipped to +-1e6 and divided by voxel_size in float64; cloud B keeps its type, so float32
//      points are clipped and divided in float32 (NumPy weak-scalar promotion); floor, int32 cast.
//      Points with a non-finite column (intensity included) are dropped (:358-359).
//   2. np.unique + Python set algebra (:371-387) become one open-addressing hash set in shared
//      memory whose slots hold POINT INDICES: a slot is claimed with atomicCAS(slot, EMPTY, index)
//      and a collision is resolved by comparing the full 96-bit keys of the two points, so there is
//      no key packing and no hash ambiguity. A's points are inserted first; B's points then either
//      claim a fresh slot (a voxel only B has), hit one of B's own slots (duplicate), or hit one of
//      A's (counted once per A voxel through a bit set).
// Counts are exact integers and independent of thread order; IoU = |A & B| / |A | B| in float64.
#include <math.h>

#include "nsc_internal.h"

namespace nsc {

namespace {

constexpr int kOvThreads = 512;
constexpr int kEmptySlot = -1;
constexpr int kInvalidX = INT32_MIN;          // no clipped coordinate reaches it (|v| <= 1e6 / voxel)

struct OverlapArgs {
    const void* points;          // float32 or float64, stride 3 or 4
    const long long* offsets;    // 2 * n_pairs + 1: A_0, B_0, A_1, B_1, ...
    const double* T;             // n_pairs x 16, row-major, maps A into B's frame
    int stride, is_f64, n_pairs;
    double voxel;
    float voxel_f;
    int* keys;                   // workspace: 3 ints per point of the batch
    int* table_g;                // workspace hash tables for pairs too large for shared memory
    long long table_g_stride;
    int smem_slots;              // capacity of the shared-memory table (power of two)
    int* counts;                 // n_pairs x 3: |A|, |B|, |A & B|
    double* iou;                 // n_pairs
};

__device__ __forceinline__ unsigned hash3(int x, int y, int z) {
    unsigned h = (unsigned)x * 73856093u ^ (unsigned)y * 19349663u ^ (unsigned)z * 83492791u;
    h ^= h >> 15;
    h *= 0x2c1b3c6du;
    h ^= h >> 12;
    return h;
}

template <typename T>
__device__ __forceinline__ void voxel_of_a(const T* p, int stride, const double* M, double voxel, int* key) {
    // transform_points: float64 row . [x y z 1] (pose_utils.py:121-129), then voxelize_fast (:356-369)
    const double x = (double)p[0], y = (double)p[1], z = (double)p[2];
    bool ok = stride == 3 || isfinite((double)p[3]);
    double c[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double acc = __dmul_rn(M[4 * r], x);
        acc = fma(M[4 * r + 1], y, acc);
        acc = fma(M[4 * r + 2], z, acc);
        c[r] = __dadd_rn(acc, M[4 * r + 3]);
        ok = ok && isfinite(c[r]);
    }
    if (!ok) { key[0] = kInvalidX; key[1] = 0; key[2] = 0; return; }
#pragma unroll
    for (int r = 0; r < 3; ++r)
        key[r] = (int)floor(__ddiv_rn(fmin(fmax(c[r], -1e6), 1e6), voxel));
}

__device__ __forceinline__ void voxel_of_b(const float* p, int stride, float voxel_f, int* key) {
    bool ok = isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]) && (stride == 3 || isfinite(p[3]));
    if (!ok) { key[0] = kInvalidX; key[1] = 0; key[2] = 0; return; }
#pragma unroll
    for (int r = 0; r < 3; ++r)       // float32 clip and float32 division (weak Python-float promotion)
        key[r] = (int)floorf(__fdiv_rn(fminf(fmaxf(p[r], -1e6f), 1e6f), voxel_f));
}

__device__ __forceinline__ void voxel_of_b(const double* p, int stride, double voxel, int* key) {
    bool ok = isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]) && (stride == 3 || isfinite(p[3]));
    if (!ok) { key[0] = kInvalidX; key[1] = 0; key[2] = 0; return; }
#pragma unroll
    for (int r = 0; r < 3; ++r)
        key[r] = (int)floor(__ddiv_rn(fmin(fmax(p[r], -1e6), 1e6), voxel));
}

__global__ void __launch_bounds__(kOvThreads)
voxel_overlap_kernel(const __grid_constant__ OverlapArgs a) {
    extern __shared__ int smem[];
    __shared__ int s_cnt[3];     // voxels of A, voxels only B has, A voxels B hit
    __shared__ double s_T[16];
    const int tid = threadIdx.x;
    for (int pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
        const long long a0 = a.offsets[2 * pair], b0 = a.offsets[2 * pair + 1], b1 = a.offsets[2 * pair + 2];
        const int na = (int)(b0 - a0), nb = (int)(b1 - b0), n = na + nb;
        int cap = 64;
        while (cap < 2 * n) cap <<= 1;
        int* table = cap <= a.smem_slots ? smem : a.table_g + (long long)blockIdx.x * a.table_g_stride;
        unsigned* hit = reinterpret_cast<unsigned*>(table + cap);        // one bit per slot
        int* keys = a.keys + 3 * a0;
        __syncthreads();
        if (tid < 16) s_T[tid] = a.T[16 * (long long)pair + tid];
        if (tid < 3) s_cnt[tid] = 0;
        for (int i = tid; i < cap; i += kOvThreads) table[i] = kEmptySlot;
        for (int i = tid; i < cap / 32; i += kOvThreads) hit[i] = 0u;
        __syncthreads();
        for (int i = tid; i < n; i += kOvThreads) {
            int k[3];
            if (a.is_f64) {
                const double* p = reinterpret_cast<const double*>(a.points) + (a0 + i) * a.stride;
                if (i < na) voxel_of_a(p, a.stride, s_T, a.voxel, k);
                else voxel_of_b(p, a.stride, a.voxel, k);
            } else {
                const float* p = reinterpret_cast<const float*>(a.points) + (a0 + i) * a.stride;
                if (i < na) voxel_of_a(p, a.stride, s_T, a.voxel, k);
                else voxel_of_b(p, a.stride, a.voxel_f, k);
            }
            keys[3 * i] = k[0];
            keys[3 * i + 1] = k[1];
            keys[3 * i + 2] = k[2];
        }
        __syncthreads();
        const unsigned mask = (unsigned)cap - 1u;
        int mine[3] = {0, 0, 0};
        for (int phase = 0; phase < 2; ++phase) {          // A's points, then (after a barrier) B's
            const int lo = phase == 0 ? 0 : na, hi = phase == 0 ? na : n;
            for (int i = lo + tid; i < hi; i += kOvThreads) {
                const int x = keys[3 * i], y = keys[3 * i + 1], z = keys[3 * i + 2];
                if (x == kInvalidX) continue;
                unsigned h = hash3(x, y, z) & mask;
                for (;;) {
                    const int prev = atomicCAS(&table[h], kEmptySlot, i);
                    if (prev == kEmptySlot) { ++mine[phase]; break; }          // a voxel nobody had
                    if (keys[3 * prev] == x && keys[3 * prev + 1] == y && keys[3 * prev + 2] == z) {
                        if (phase == 1 && prev < na) {                          // B meets a voxel of A
                            const unsigned bit = 1u << (h & 31);
                            if (!(atomicOr(&hit[h >> 5], bit) & bit)) ++mine[2];
                        }
                        break;
                    }
                    h = (h + 1) & mask;
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int v = __reduce_add_sync(0xffffffffu, mine[c]);
            if ((tid & 31) == 0 && v) atomicAdd(&s_cnt[c], v);
        }
        __syncthreads();
        if (tid == 0) {
            const int va = s_cnt[0], inter = s_cnt[2], vb = s_cnt[1] + inter, uni = va + s_cnt[1];
            a.counts[3 * pair] = va;
            a.counts[3 * pair + 1] = vb;
            a.counts[3 * pair + 2] = inter;
            a.iou[pair] = uni > 0 ? (double)inter / (double)uni : 0.0;       // pose_utils.py:385-389
        }
    }
}

constexpr int kSmemSlots = 32768;                       // 128 KB of slots + 4 KB of bits

size_t table_words(long long max_pair_points) {
    long long cap = 64;
    while (cap < 2 * max_pair_points) cap <<= 1;
    return (size_t)(cap + cap / 32);
}

}  // namespace

}  // namespace nsc

using namespace nsc;

extern "C" {

size_t nsc_voxel_overlap_workspace_bytes(int64_t total_points, int64_t max_pair_points, int n_pairs) {
    if (total_points < 0 || max_pair_points < 0 || n_pairs < 0) return 0;
    size_t bytes = (size_t)total_points * 12 + 256;
    if (2 * max_pair_points > kSmemSlots) {               // pairs that need a table in global memory
        const int ctas = n_pairs < 1024 ? n_pairs : 1024;
        bytes += (size_t)ctas * table_words(max_pair_points) * 4;
    }
    return bytes;
}

int nsc_voxel_overlap_batch(const void* d_points, int point_stride, int points_are_f64,
                            const int64_t* d_offsets, int64_t total_points, int64_t max_pair_points,
                            const double* d_T, int n_pairs, double voxel_size, int32_t* d_counts,
                            double* d_iou, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (n_pairs < 0 || total_points < 0 || max_pair_points < 0) return NSC_ERR_BAD_COUNT;
    if (point_stride != 3 && point_stride != 4) return NSC_ERR_BAD_STRIDE;
    if (!(voxel_size > 0.0) || !isfinite(voxel_size)) return NSC_ERR_BAD_PARAMS;
    if (max_pair_points > (1ll << 28)) return NSC_ERR_BAD_COUNT;
    if (n_pairs == 0) return NSC_OK;
    if (!d_offsets || !d_T || !d_counts || !d_iou) return NSC_ERR_NULL_POINTER;
    if (total_points > 0 && !d_points) return NSC_ERR_NULL_POINTER;
    if (!d_workspace || workspace_bytes < nsc_voxel_overlap_workspace_bytes(total_points, max_pair_points, n_pairs))
        return NSC_ERR_WORKSPACE;
    OverlapArgs a;
    a.points = d_points;
    a.offsets = (const long long*)d_offsets;
    a.T = d_T;
    a.stride = point_stride;
    a.is_f64 = points_are_f64 ? 1 : 0;
    a.n_pairs = n_pairs;
    a.voxel = voxel_size;
    a.voxel_f = (float)voxel_size;
    a.keys = (int*)d_workspace;
    const size_t keys_bytes = ((size_t)total_points * 12 + 255) & ~(size_t)255;
    a.table_g = (int*)((char*)d_workspace + keys_bytes);
    a.table_g_stride = (long long)table_words(max_pair_points);
    const bool in_smem = 2 * max_pair_points <= kSmemSlots;
    long long cap = 64;
    while (cap < 2 * max_pair_points) cap <<= 1;
    a.smem_slots = in_smem ? (int)cap : 0;
    a.counts = d_counts;
    a.iou = d_iou;
    const size_t smem = in_smem ? (size_t)(cap + cap / 32) * 4 : 0;
    cudaError_t e = cudaFuncSetAttribute(voxel_overlap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(kSmemSlots + kSmemSlots / 32) * 4);
    if (e != cudaSuccess) return record_cuda(e);
    const int grid = n_pairs < 1024 ? n_pairs : 1024;
    voxel_overlap_kernel<<<grid, kOvThreads, smem, (cudaStream_t)stream>>>(a);
    return record_cuda(cudaGetLastError());
}

}  // extern "C"

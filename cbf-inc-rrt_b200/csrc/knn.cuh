// knn.cuh -- planner front-end: brute-force nearest-neighbour / k-nearest / radius queries in configuration space.
//
// Replaces the numpy distance matrices of the planners that call the hot path:
//   batch_tsa.global_explore  (motion_planning/baseline/batch_tsa.py:135-145): argmin / argsort[:batch] of
//       ||sample - tree_node|| over the non-terminal states;
//   tsa.global_explore        (motion_planning/baseline/tsa.py:151-153): the same for one sample;
//   BITStarPlanner.expand_vertex (motion_planning/baseline/bit_star.py:196-216): nodes within radius r.
//
// Contract (shared with the oracle): d_k = q_k - t_k in float32, dsq = fma(d_{n-1}, d_{n-1}, ... fma(d_1, d_1, d_0*d_0));
// candidates are ordered by the 64-bit key (bits(dsq) << 32 | row) -- dsq >= 0, so its bit pattern is monotone -- i.e.
// by squared distance, ties to the lowest row (np.argmin / stable argsort order); dist = sqrt_rn(dsq).
// HBM/L2-bound integer-key reduction.  k = 1: the tree is streamed once per 128 queries, one atomicMin per (query,
// chunk).  1 < k <= 1024: two launches -- per-chunk shared-memory sort keeping k candidates, then a per-query merge.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cbfrrt {
namespace knn {

constexpr int QT = 128;        // queries per CTA (one per thread)
constexpr int CHUNK = 512;     // tree rows staged per CTA
constexpr unsigned long long NONE = ~0ull;

template <int N>
__device__ __forceinline__ float dsq_of(const float (&q)[N], const float* __restrict__ t) {
    float d = __fsub_rn(q[0], t[0]);
    float acc = __fmul_rn(d, d);
#pragma unroll
    for (int k = 1; k < N; ++k) { d = __fsub_rn(q[k], t[k]); acc = fmaf(d, d, acc); }
    return acc;
}

// best[m] = min over rows of this chunk of keys > lower[m] (lower == nullptr: no bound); best preset to NONE
template <int N>
__global__ void __launch_bounds__(QT) nn_kernel(const float* __restrict__ q, const float* __restrict__ tree, long long M,
                                                long long Nt, const unsigned long long* __restrict__ lower,
                                                unsigned long long* __restrict__ best) {
    __shared__ float rows[CHUNK * N];
    const long long r0 = (long long)blockIdx.x * CHUNK;
    const int nr = (int)((Nt - r0) < CHUNK ? (Nt - r0) : CHUNK);
    for (int i = threadIdx.x; i < nr * N; i += QT) rows[i] = __ldg(tree + r0 * N + i);
    __syncthreads();
    const long long m = (long long)blockIdx.y * QT + threadIdx.x;
    if (m >= M) return;
    float qq[N];
#pragma unroll
    for (int k = 0; k < N; ++k) qq[k] = __ldg(q + m * N + k);
    const unsigned long long lo = lower ? lower[m] : 0ull;
    const bool bounded = lower != nullptr;
    unsigned long long b = NONE;
    for (int i = 0; i < nr; ++i) {
        const float dsq = dsq_of<N>(qq, rows + i * N);                       // broadcast LDS
        const unsigned long long key = ((unsigned long long)__float_as_uint(dsq) << 32) | (unsigned)(r0 + i);
        if (key < b && (!bounded || key > lo)) b = key;
    }
    if (b != NONE) atomicMin(best + m, b);
}

// idx[m*stride] / dist[m*stride] from best[m]; also copies best -> lower for the next round of a k-NN query
static __global__ void nn_finish(const unsigned long long* __restrict__ best, unsigned long long* __restrict__ lower,
                                 long long M, int* __restrict__ idx, float* __restrict__ dist, int stride) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const unsigned long long b = best[m];
    idx[m * stride] = b == NONE ? -1 : (int)(unsigned)(b & 0xffffffffull);
    dist[m * stride] = b == NONE ? __int_as_float(0x7f800000) : __fsqrt_rn(__uint_as_float((unsigned)(b >> 32)));
    if (lower) lower[m] = b;
}

// ---- k > 1 in two launches (instead of k rounds): per (chunk of TOPK_CHUNK rows, query) sort the chunk's keys in shared
// memory and keep its k smallest; then one CTA per query merges the per-chunk candidate lists, 2048 keys at a time.
constexpr int TOPK_CHUNK = 1024;   // rows per CTA of the candidate kernel (256 threads, 4 rows each)
constexpr int TOPK_MAXK = 1024;    // larger k falls back to the rounds
constexpr int MERGE_TILE = 2048;

// ascending bitonic sort of a power-of-two array of unique-or-NONE keys in shared memory (all threads of the CTA)
template <int LEN>
__device__ __forceinline__ void bitonic_sort(unsigned long long* s) {
    for (int size = 2; size <= LEN; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < LEN / 2; t += blockDim.x) {
                const int i = 2 * t - (t & (stride - 1)), j = i + stride;
                const unsigned long long x = s[i], y = s[j];
                if ((x > y) == ((i & size) == 0)) { s[i] = y; s[j] = x; }
            }
        }
    }
    __syncthreads();
}

// cand[(m * n_chunks + chunk) * k + j] = j-th smallest key of the chunk (NONE where the chunk has fewer rows)
template <int N>
__global__ void __launch_bounds__(256) topk_chunk_kernel(const float* __restrict__ q, const float* __restrict__ tree,
                                                         long long Nt, int k, unsigned long long* __restrict__ cand) {
    __shared__ unsigned long long keys[TOPK_CHUNK];
    const long long m = blockIdx.y, r0 = (long long)blockIdx.x * TOPK_CHUNK;
    float qq[N];
#pragma unroll
    for (int j = 0; j < N; ++j) qq[j] = __ldg(q + m * N + j);
    for (int i = threadIdx.x; i < TOPK_CHUNK; i += 256) {
        const long long row = r0 + i;
        unsigned long long key = NONE;
        if (row < Nt) {
            float t[N];
#pragma unroll
            for (int j = 0; j < N; ++j) t[j] = __ldg(tree + row * N + j);
            key = ((unsigned long long)__float_as_uint(dsq_of<N>(qq, t)) << 32) | (unsigned)row;
        }
        keys[i] = key;
    }
    bitonic_sort<TOPK_CHUNK>(keys);
    unsigned long long* out = cand + (m * gridDim.x + blockIdx.x) * (long long)k;
    for (int i = threadIdx.x; i < k; i += 256) out[i] = keys[i];
}

// one CTA (1024 threads) per query: the k smallest of its n_cand candidate keys -> idx / dist
static __global__ void __launch_bounds__(1024) topk_merge_kernel(const unsigned long long* __restrict__ cand, long long n_cand,
                                                                 int k, int* __restrict__ idx, float* __restrict__ dist) {
    __shared__ unsigned long long s[MERGE_TILE];
    const long long m = blockIdx.x;
    const unsigned long long* c = cand + m * n_cand;
    int kept = 0;                                  // s[0, kept) holds the best keys so far (sorted)
    for (long long pos = 0; pos < n_cand || pos == 0;) {
        const long long take = (n_cand - pos) < (MERGE_TILE - kept) ? (n_cand - pos) : (MERGE_TILE - kept);
        for (int i = threadIdx.x; i < MERGE_TILE - kept; i += 1024) s[kept + i] = i < take ? c[pos + i] : NONE;
        bitonic_sort<MERGE_TILE>(s);
        pos += take;
        kept = k;
        if (take == 0) break;
    }
    for (int i = threadIdx.x; i < k; i += 1024) {
        const unsigned long long b = s[i];
        idx[m * k + i] = b == NONE ? -1 : (int)(unsigned)(b & 0xffffffffull);
        dist[m * k + i] = b == NONE ? __int_as_float(0x7f800000) : __fsqrt_rn(__uint_as_float((unsigned)(b >> 32)));
    }
}

// within[m, j] = sqrt_rn(dsq(m, j)) <= r
template <int N>
__global__ void __launch_bounds__(256) radius_kernel(const float* __restrict__ q, const float* __restrict__ set, long long M,
                                                     long long Nt, float r, uint8_t* __restrict__ within) {
    const long long total = M * Nt;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long m = i / Nt, j = i - m * Nt;
        float qq[N];
#pragma unroll
        for (int k = 0; k < N; ++k) qq[k] = __ldg(q + m * N + k);
        within[i] = __fsqrt_rn(dsq_of<N>(qq, set + j * N)) <= r;
    }
}

}  // namespace knn
}  // namespace cbfrrt

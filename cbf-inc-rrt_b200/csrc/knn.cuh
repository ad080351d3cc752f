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
// HBM/L2-bound integer-key reduction: the tree is streamed once per 128 queries, one atomicMin per (query, chunk).
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

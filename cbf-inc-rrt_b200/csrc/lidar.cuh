// lidar.cuh -- the LiDAR / point-cloud barrier of the reference on sm_100a:
//     datax_to_x            neural_cbf/systems/arm_lidar.py:219-248   sampled cloud points -> sensor frames, R_s^T (x - p_s)
//     PointNetfeat          neural_cbf/controllers/utils/pointnet.py:13-70   [point | normal | one-hot sensor] -> 64 -> 128 -> 512
//                           (LeakyReLU 0.1), max over the rays of a sensor, 512 -> 256 -> per_feature_dim
//     PointNetVanillaEncoder pointnet.py:73-97                        S * per_feature_dim -> 512 -> 256 -> feature_dim
//     h                     neural_cbf/controllers/neural_lidar_cbf_controller.py:115-139   h_nn([q | encoded])
//     batch_lookahead       arm_lidar.py:250-288                      q + dq, frames advanced to first order by (J_p, J_R)
//     h_with_jacobian       neural_lidar_cbf_controller.py:141-219    one-sided differences, (h(q + e_k d) - h(q)) / (2 d)
//
// This is the one GEMM-sized workload of the reference: 149 kflop per point, S x R points per evaluation, n + 1
// evaluations per state.  backbone_kernel runs the three per-point layers on the tcgen05 tensor cores with the same
// building blocks as mlp_tc.cuh (A_hi in tensor memory, A_lo and the weight tiles in shared memory, 3-term TF32 split --
// the 1e-5 bar rules out plain TF32: 6e-4, and a 3-term bf16 split: 3e-5; tools/experiments/pointnet_tf32x3_sim.py):
//   * one CTA = 128 threads = one tile of 128 points = the M of every MMA; thread t <-> point t <-> TMEM lane t;
//   * the tile's A operand of layer 1 is PRODUCED in registers (gather of the sampled point, frame transform with the
//     look-ahead increment, one-hot): datax_to_x and batch_lookahead never touch memory;
//   * TMEM: D in columns 0..127, A_hi in columns 128..255 (K <= 128); layer 3 (128 -> 512) runs as 8 chunks of 64 output
//     columns, its weights streamed through one shared-memory chunk buffer;
//   * layer-3 epilogue: bias, LeakyReLU, max over the warp's 32 rows per column by REDUX on the order-preserving integer
//     image of the float, one atomicMax per (warp, column) into the [evaluation, sensor, 512] key buffer.
// head_kernel does everything after the max-pool in FP32 (< 1 % of the flops): one CTA per evaluation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cbfrrt.h"
#include "mlp_tc.cuh"

namespace cbfrrt {
namespace lidar {

using namespace tc;

constexpr int PT = 128;                                   // points per tile
constexpr int K1 = 16, N1 = 64, N2 = 128, N3 = 512, NCH = 64, CHUNKS = N3 / NCH;
constexpr int COLD = 0, COLA = 128;                       // TMEM columns of D and A_hi
constexpr int MAX_S = 7, MAX_N = 8;
// shared memory (floats)
constexpr int S_W1 = 0;                                   // [hi | lo] 16 x 64
constexpr int S_W2 = S_W1 + 2 * K1 * N1;                  // [hi | lo] 64 x 128
constexpr int S_W3 = S_W2 + 2 * N1 * N2;                  // [hi | lo] 128 x 64 (one chunk)
constexpr int S_ALO = S_W3 + 2 * N2 * NCH;                // [128/4][128] float4
constexpr int S_B = S_ALO + N2 * PT;                      // b1[64] b2[128] b3[512]
constexpr int S_MBAR = S_B + N1 + N2 + N3;                // one 64-bit mbarrier (+ pad)
constexpr int S_TMEM = S_MBAR + 4;
constexpr int S_TOTAL = S_TMEM + 4;

struct Tile {
    float* sm;
    uint32_t tmem, lane, phase;
    int t;
};

__device__ __forceinline__ float leaky(float s) { return s > 0.f ? s : 0.1f * s; }
__device__ __forceinline__ int ordered(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float unordered(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

template <int K>
__device__ __forceinline__ void store_a(const Tile& c, const float* v) {
    const uint32_t row = c.tmem + c.lane + COLA;
    float4* alo = reinterpret_cast<float4*>(c.sm + S_ALO) + c.t;
#pragma unroll
    for (int kc = 0; kc < K / 4; ++kc) {
        float4 lo;
        lo.x = v[4 * kc] - __uint_as_float(__float_as_uint(v[4 * kc]) & 0xffffe000u);
        lo.y = v[4 * kc + 1] - __uint_as_float(__float_as_uint(v[4 * kc + 1]) & 0xffffe000u);
        lo.z = v[4 * kc + 2] - __uint_as_float(__float_as_uint(v[4 * kc + 2]) & 0xffffe000u);
        lo.w = v[4 * kc + 3] - __uint_as_float(__float_as_uint(v[4 * kc + 3]) & 0xffffe000u);
        alo[kc * PT] = lo;
    }
#pragma unroll
    for (int j = 0; j < K / 16; ++j) tmem_st16(row + 16 * j, v + 16 * j);
    fence_proxy_async();
    tmem_st_wait();
}

// D[128 x N] = A[128 x K] . B^T, 3-term split; all 128 threads call it
template <int K, int N>
__device__ __forceinline__ void gemm(Tile& c, const float* w_hi, const float* w_lo) {
    fence_before_sync();
    __syncthreads();
    if (c.t < 32) {
        fence_after_sync();
        constexpr uint32_t idesc = instr_desc(N);
        constexpr uint32_t B_LBO = N * 16, B_SBO = 128, B_STEP = 2 * N * 16;
        constexpr uint32_t A_LBO = PT * 16, A_SBO = 128, A_STEP = 2 * PT * 16;
        const uint32_t b_hi = smem_addr(w_hi), b_lo = smem_addr(w_lo);
        const uint32_t a_hi = c.tmem + COLA, a_lo = smem_addr(c.sm + S_ALO);
        uint64_t* bar = reinterpret_cast<uint64_t*>(c.sm + S_MBAR);
#pragma unroll
        for (int ks = 0; ks < K / 8; ++ks)
            mma_tf32_ts(c.tmem + COLD, a_hi + 8 * ks, smem_desc(b_hi + ks * B_STEP, B_LBO, B_SBO), idesc, ks ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < K / 8; ++ks)
            mma_tf32_ts(c.tmem + COLD, a_hi + 8 * ks, smem_desc(b_lo + ks * B_STEP, B_LBO, B_SBO), idesc, 1u);
#pragma unroll
        for (int ks = 0; ks < K / 8; ++ks)
            mma_tf32_ss(c.tmem + COLD, smem_desc(a_lo + ks * A_STEP, A_LBO, A_SBO), smem_desc(b_hi + ks * B_STEP, B_LBO, B_SBO), idesc, 1u);
        mma_commit(bar);
        mbar_wait_parity(bar, c.phase);
        fence_before_sync();
    }
    c.phase ^= 1u;
    __syncthreads();
    fence_after_sync();
}

template <int NCOL>
__device__ __forceinline__ void load_d(const Tile& c, float* v) {
    const uint32_t row = c.tmem + c.lane + COLD;
#pragma unroll
    for (int j = 0; j < NCOL / 16; ++j) tmem_ld16(row + 16 * j, v + 16 * j);
    tmem_ld_wait();
}

struct CloudArgs {
    const float* q;              // [B, n]
    const float* cloud;          // [B, P, pd]  global point cloud of every state (world frame), pd = 3 or 6
    const float* frames;         // [B, S, 12]  per sensor: origin (3) | rotation (9, row-major)
    const float* J_p;            // [B, S, 3, n] or nullptr   d origin / dq
    const float* J_R;            // [B, S, 9, n] or nullptr   d rotation / dq
    const int* sampled;          // [S, R]      cloud indices of the rays of every sensor (one draw for the whole batch)
    long long B, P;
    int n, S, R, pd, evals;      // evals = 1 (h only) or n + 1 (h and its one-sided differences)
    float delta;                 // look-ahead step d: evaluation k >= 1 is q + d e_{k-1}
};

// the 16-wide input row of point `p` (global index over [B * evals, S, R]): datax_to_x + batch_lookahead in registers
__device__ __forceinline__ void point_row(const CloudArgs& a, long long p, float (&v)[K1]) {
    const int r = (int)(p % a.R);
    const long long es = p / a.R;
    const int s = (int)(es % a.S);
    const long long e = es / a.S;
    const long long b = e / a.evals;
    const int k = (int)(e % a.evals);
    const float* f = a.frames + (b * a.S + s) * 12;
    float o[3], Rm[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = __ldg(f + i);
#pragma unroll
    for (int i = 0; i < 9; ++i) Rm[i] = __ldg(f + 3 + i);
    if (k > 0) {   // arm_lidar.py:250-288: first-order advance of the sensor frame along joint k - 1
        const float* jp = a.J_p + ((b * a.S + s) * 3) * a.n + (k - 1);
        const float* jr = a.J_R + ((b * a.S + s) * 9) * a.n + (k - 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) o[i] = fmaf(__ldg(jp + i * a.n), a.delta, o[i]);
#pragma unroll
        for (int i = 0; i < 9; ++i) Rm[i] = fmaf(__ldg(jr + i * a.n), a.delta, Rm[i]);
    }
    const int idx = __ldg(a.sampled + s * a.R + r);
    const float* pt = a.cloud + (b * a.P + idx) * a.pd;
    const float dx = __ldg(pt) - o[0], dy = __ldg(pt + 1) - o[1], dz = __ldg(pt + 2) - o[2];
#pragma unroll
    for (int i = 0; i < K1; ++i) v[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = fmaf(Rm[6 + i], dz, fmaf(Rm[3 + i], dy, Rm[i] * dx));          // R^T (x - p)
    int c0 = 3;
    if (a.pd == 6) {
        const float nx = __ldg(pt + 3), ny = __ldg(pt + 4), nz = __ldg(pt + 5);
#pragma unroll
        for (int i = 0; i < 3; ++i) v[3 + i] = fmaf(Rm[6 + i], nz, fmaf(Rm[3 + i], ny, Rm[i] * nx));   // R^T n
        c0 = 6;
    }
#pragma unroll
    for (int i = 0; i < MAX_S; ++i)
        if (c0 + i < K1) v[c0 + i] = (i == s) ? 1.f : 0.f;                                           // one-hot sensor id
}

// w1t / w2t / w3t: [hi | lo] tiles in the no-swizzle K-major layout W4[k/4][n] (w3t: 8 chunks of [hi | lo] 128 x 64),
// prepared by the host (cbfrrt.ops.pack_lidar_net); bias = [b1 | b2 | b3];
// out_max [n_eval * S, 512]: int keys (ordered image of the float), preset to INT_MIN
__global__ void __launch_bounds__(PT, 1) backbone_kernel(CloudArgs a, long long n_points, const float* __restrict__ w1t,
                                                         const float* __restrict__ w2t, const float* __restrict__ w3t,
                                                         const float* __restrict__ bias, int* __restrict__ out_max) {
    extern __shared__ __align__(128) float sm[];
    Tile c;
    c.sm = sm;
    c.t = threadIdx.x;
    c.phase = 0;
    for (int i = c.t; i < 2 * K1 * N1; i += PT) sm[S_W1 + i] = __ldg(w1t + i);
    for (int i = c.t; i < 2 * N1 * N2; i += PT) sm[S_W2 + i] = __ldg(w2t + i);
    for (int i = c.t; i < N1 + N2 + N3; i += PT) sm[S_B + i] = __ldg(bias + i);
    if (c.t == 0) {
        mbar_init1(reinterpret_cast<uint64_t*>(sm + S_MBAR));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (c.t < 32) tmem_alloc(reinterpret_cast<uint32_t*>(sm + S_TMEM), 512);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    c.tmem = *reinterpret_cast<uint32_t*>(sm + S_TMEM);
    c.lane = (uint32_t)(c.t & ~31) << 16;

    const long long n_tiles = (n_points + PT - 1) / PT;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long p = tile * PT + c.t;
        const bool active = p < n_points;                  // R % 32 == 0: uniform over a warp
        float v[N2];
        {   // layer 1: K = 16, N = 64
            float x[K1];
#pragma unroll
            for (int i = 0; i < K1; ++i) x[i] = 0.f;
            if (active) point_row(a, p, x);
            store_a<K1>(c, x);
            gemm<K1, N1>(c, sm + S_W1, sm + S_W1 + K1 * N1);
            load_d<N1>(c, v);
#pragma unroll
            for (int k = 0; k < N1; ++k) v[k] = leaky(v[k] + sm[S_B + k]);
        }
        {   // layer 2: K = 64, N = 128
            store_a<N1>(c, v);
            gemm<N1, N2>(c, sm + S_W2, sm + S_W2 + N1 * N2);
            load_d<N2>(c, v);
#pragma unroll
            for (int k = 0; k < N2; ++k) v[k] = leaky(v[k] + sm[S_B + N1 + k]);
        }
        store_a<N2>(c, v);                                  // A of layer 3 stays for all 8 chunks
        // the warp's 32 rows belong to one (evaluation, sensor) group: R % 32 == 0
        int* out = out_max + ((tile * PT + (c.t & ~31)) / a.R) * N3;
        for (int ch = 0; ch < CHUNKS; ++ch) {               // layer 3: K = 128, N = 64 per chunk
            __syncthreads();                                // previous chunk's MMAs are complete (gemm waited on them)
            const float4* src = reinterpret_cast<const float4*>(w3t + (long long)ch * 2 * N2 * NCH);
            float4* dst = reinterpret_cast<float4*>(sm + S_W3);
            for (int i = c.t; i < 2 * N2 * NCH / 4; i += PT) dst[i] = __ldg(src + i);
            fence_proxy_async();
            gemm<N2, NCH>(c, sm + S_W3, sm + S_W3 + N2 * NCH);
            float d[NCH];
            load_d<NCH>(c, d);
            if (active) {
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    const int key = __reduce_max_sync(0xffffffffu, ordered(leaky(d[k] + sm[S_B + N1 + N2 + ch * NCH + k])));
                    if ((c.t & 31) == (k & 31)) atomicMax(out + ch * NCH + k, key);
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (c.t < 32) tmem_dealloc(c.tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// backbone_kernel_v2: the same arithmetic, software-pipelined.  What changed against backbone_kernel (kept as the
// reference implementation, CBFRRT_LIDAR_V1=1):
//   * A_lo lives in TENSOR MEMORY next to A_hi (columns 256..383), so all three passes of the split are TS-form MMAs and
//     the 64 KB shared-memory A_lo tile is gone -- which pays for a SECOND layer-3 weight buffer;
//   * the layer-3 chunks (64 output columns = 64 KB of hi / lo tiles each) arrive by TMA bulk copies
//     (cp.async.bulk ... mbarrier::complete_tx) into the two buffers, issued as soon as the MMAs that read a buffer have
//     completed: the copy of chunk c + 1 overlaps the MMAs of chunk c and the epilogue of chunk c - 1;
//   * two accumulators (TMEM columns 0..63 / 64..127): the MMAs of chunk c run while all four warps reduce chunk c - 1;
//   * epilogue: max over the warp's rows FIRST (REDUX on the ordered-int image of the raw accumulator), then bias +
//     LeakyReLU once per (warp, column) -- both are monotone, so max(leaky(x + b)) = leaky(max(x) + b).
namespace v2 {
constexpr int COL_ALO = 256;
constexpr int CH_FLOATS = 2 * N2 * NCH;                     // one chunk: [hi | lo] 128 x 64
constexpr uint32_t CH_BYTES = CH_FLOATS * 4;
constexpr int S_W1 = 0;
constexpr int S_W2 = S_W1 + 2 * K1 * N1;
constexpr int S_W3A = S_W2 + 2 * N1 * N2;                   // chunk buffer 0
constexpr int S_W3B = S_W3A + CH_FLOATS;                    // chunk buffer 1
constexpr int S_B = S_W3B + CH_FLOATS;
constexpr int S_MBAR = S_B + N1 + N2 + N3;                  // 5 mbarriers: layer gemm, weights[2], accumulators[2]
constexpr int S_TMEM = S_MBAR + 12;
constexpr int S_TOTAL = S_TMEM + 4;

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

template <int K>
__device__ __forceinline__ void store_a2(const Tile& c, const float* v) {
    const uint32_t row = c.tmem + c.lane;
#pragma unroll
    for (int j = 0; j < K / 16; ++j) {
        float lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) lo[i] = v[16 * j + i] - __uint_as_float(__float_as_uint(v[16 * j + i]) & 0xffffe000u);
        tmem_st16(row + COLA + 16 * j, v + 16 * j);
        tmem_st16(row + COL_ALO + 16 * j, lo);
    }
    tmem_st_wait();
}

// issue the three passes of D[d_col .. d_col + N) = A . B^T (warp 0, elected lane inside the wrappers) and commit to bar
template <int K, int N>
__device__ __forceinline__ void issue_mmas(const Tile& c, uint32_t d_col, const float* w_hi, const float* w_lo, uint64_t* bar) {
    constexpr uint32_t idesc = instr_desc(N);
    constexpr uint32_t B_LBO = N * 16, B_SBO = 128, B_STEP = 2 * N * 16;
    const uint32_t b_hi = smem_addr(w_hi), b_lo = smem_addr(w_lo);
    const uint32_t a_hi = c.tmem + COLA, a_lo = c.tmem + COL_ALO, d = c.tmem + d_col;
#pragma unroll
    for (int ks = 0; ks < K / 8; ++ks)
        mma_tf32_ts(d, a_hi + 8 * ks, smem_desc(b_hi + ks * B_STEP, B_LBO, B_SBO), idesc, ks ? 1u : 0u);
#pragma unroll
    for (int ks = 0; ks < K / 8; ++ks)
        mma_tf32_ts(d, a_hi + 8 * ks, smem_desc(b_lo + ks * B_STEP, B_LBO, B_SBO), idesc, 1u);
#pragma unroll
    for (int ks = 0; ks < K / 8; ++ks)
        mma_tf32_ts(d, a_lo + 8 * ks, smem_desc(b_hi + ks * B_STEP, B_LBO, B_SBO), idesc, 1u);
    mma_commit(bar);
}

// synchronous GEMM of layers 1 and 2 into columns [0, N)
template <int K, int N>
__device__ __forceinline__ void gemm_sync(Tile& c, const float* w_hi, const float* w_lo, uint64_t* bar) {
    fence_before_sync();
    __syncthreads();
    if (c.t < 32) {
        fence_after_sync();
        issue_mmas<K, N>(c, COLD, w_hi, w_lo, bar);
    }
    mbar_wait_parity(bar, c.phase);
    c.phase ^= 1u;
    fence_after_sync();
}

template <int NCOL>
__device__ __forceinline__ void load_cols(const Tile& c, uint32_t col, float* v) {
    const uint32_t row = c.tmem + c.lane + col;
#pragma unroll
    for (int j = 0; j < NCOL / 16; ++j) tmem_ld16(row + 16 * j, v + 16 * j);
    tmem_ld_wait();
}

// reduce one 64-column chunk of the accumulator over the warp's 32 rows and merge it into the group's keys
__device__ __forceinline__ void chunk_epilogue(const Tile& c, uint32_t d_col, int ch, bool active, int* __restrict__ out) {
    float d[NCH];
    load_cols<NCH>(c, d_col, d);
    if (!active) return;
    int mine0 = 0, mine1 = 0;
    const int lane = c.t & 31;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int i = __float_as_int(d[k]);
        const int key = __reduce_max_sync(0xffffffffu, i ^ ((i >> 31) & 0x7fffffff));
        if (k == lane) mine0 = key;
        if (k == lane + 32) mine1 = key;
    }
    const float* b3 = c.sm + S_B + N1 + N2 + ch * NCH;
    atomicMax(out + ch * NCH + lane, ordered(leaky(unordered(mine0) + b3[lane])));
    atomicMax(out + ch * NCH + lane + 32, ordered(leaky(unordered(mine1) + b3[lane + 32])));
}

__global__ void __launch_bounds__(PT, 1) backbone_kernel_v2(CloudArgs a, long long n_points, const float* __restrict__ w1t,
                                                            const float* __restrict__ w2t, const float* __restrict__ w3t,
                                                            const float* __restrict__ bias, int* __restrict__ out_max) {
    extern __shared__ __align__(128) float sm[];
    Tile c;
    c.sm = sm;
    c.t = threadIdx.x;
    c.phase = 0;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S_MBAR);      // 0: layer gemm, 1..2: weights, 3..4: accumulators
    for (int i = c.t; i < 2 * K1 * N1; i += PT) sm[S_W1 + i] = __ldg(w1t + i);
    for (int i = c.t; i < 2 * N1 * N2; i += PT) sm[S_W2 + i] = __ldg(w2t + i);
    for (int i = c.t; i < N1 + N2 + N3; i += PT) sm[S_B + i] = __ldg(bias + i);
    if (c.t == 0) {
        for (int i = 0; i < 5; ++i) mbar_init1(bars + i);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (c.t < 32) tmem_alloc(reinterpret_cast<uint32_t*>(sm + S_TMEM), 512);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    c.tmem = *reinterpret_cast<uint32_t*>(sm + S_TMEM);
    c.lane = (uint32_t)(c.t & ~31) << 16;
    uint32_t ph_w[2] = {0u, 0u}, ph_d[2] = {0u, 0u};
    float* const buf[2] = {sm + S_W3A, sm + S_W3B};

    const long long n_tiles = (n_points + PT - 1) / PT;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long p = tile * PT + c.t;
        const bool active = p < n_points;                  // R % 32 == 0: uniform over a warp
        // both chunk buffers are free here (their last readers, the MMAs of chunks 6 and 7 of the previous tile, were
        // waited for): start the copies of chunks 0 and 1 under layers 1 and 2
        if (c.t == 0) {
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                mbar_expect_tx(bars + 1 + b, CH_BYTES);
                tma_g2s(buf[b], w3t + (long long)b * CH_FLOATS, CH_BYTES, bars + 1 + b);
            }
        }
        float v[N2];
        {   // layer 1: K = 16, N = 64
            float x[K1];
#pragma unroll
            for (int i = 0; i < K1; ++i) x[i] = 0.f;
            if (active) point_row(a, p, x);
            store_a2<K1>(c, x);
            gemm_sync<K1, N1>(c, sm + S_W1, sm + S_W1 + K1 * N1, bars);
            load_cols<N1>(c, COLD, v);
#pragma unroll
            for (int k = 0; k < N1; ++k) v[k] = leaky(v[k] + sm[S_B + k]);
        }
        {   // layer 2: K = 64, N = 128
            store_a2<N1>(c, v);
            gemm_sync<N1, N2>(c, sm + S_W2, sm + S_W2 + N1 * N2, bars);
            load_cols<N2>(c, COLD, v);
#pragma unroll
            for (int k = 0; k < N2; ++k) v[k] = leaky(v[k] + sm[S_B + N1 + k]);
        }
        store_a2<N2>(c, v);                                 // A of layer 3 (hi and lo) stays in TMEM for all 8 chunks
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        int* out = out_max + ((tile * PT + (c.t & ~31)) / a.R) * N3;
        for (int ch = 0; ch < CHUNKS; ++ch) {               // layer 3, pipelined: MMAs of chunk ch | epilogue of chunk ch - 1
            const int b = ch & 1;
            if (c.t < 32) {
                mbar_wait_parity(bars + 1 + b, ph_w[b]);    // weights of chunk ch have landed
                issue_mmas<N2, NCH>(c, (uint32_t)(b * NCH), buf[b], buf[b] + N2 * NCH, bars + 3 + b);
            }
            ph_w[b] ^= 1u;
            if (ch >= 1) {
                const int pb = b ^ 1;
                mbar_wait_parity(bars + 3 + pb, ph_d[pb]);  // MMAs of chunk ch - 1 complete: accumulator pb ready, buffer pb free
                ph_d[pb] ^= 1u;
                fence_after_sync();
                if (c.t == 0 && ch + 1 < CHUNKS) {
                    mbar_expect_tx(bars + 1 + pb, CH_BYTES);
                    tma_g2s(buf[pb], w3t + (long long)(ch + 1) * CH_FLOATS, CH_BYTES, bars + 1 + pb);
                }
                chunk_epilogue(c, (uint32_t)(pb * NCH), ch - 1, active, out);
                fence_before_sync();
                __syncthreads();                            // accumulator pb may be overwritten by chunk ch + 1
                fence_after_sync();
            }
        }
        {
            const int pb = (CHUNKS - 1) & 1;
            mbar_wait_parity(bars + 3 + pb, ph_d[pb]);
            ph_d[pb] ^= 1u;
            fence_after_sync();
            chunk_epilogue(c, (uint32_t)(pb * NCH), CHUNKS - 1, active, out);
            fence_before_sync();
            __syncthreads();
            fence_after_sync();
        }
    }
    fence_before_sync();
    __syncthreads();
    if (c.t < 32) tmem_dealloc(c.tmem, 512);
}
}  // namespace v2

static __global__ void preset_keys(int* __restrict__ keys, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        keys[i] = (int)0x80000000;
}

// ---- everything after the max-pool, FP32 on the CUDA cores: one CTA (256 threads, 8 warps) per evaluation.
// y[o] = act(b[o] + sum_k W[o][k] x[k]),  W row-major [n_out][n_in] (torch layout): one warp per output, lanes over k.
__device__ __forceinline__ void dense(const float* __restrict__ W, const float* __restrict__ b, const float* xin, float* yout,
                                      int n_in, int n_out, bool act) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o = warp; o < n_out; o += 8) {
        float acc = 0.f;
        for (int k = lane; k < n_in; k += 32) acc = fmaf(__ldg(W + (long long)o * n_in + k), xin[k], acc);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0) { const float s = acc + __ldg(b + o); yout[o] = act ? leaky(s) : s; }
    }
    __syncthreads();
}

static __global__ void __launch_bounds__(256) head_kernel(const int* __restrict__ max_keys, CloudArgs a, cbfrrt_lidar_net w,
                                                          float* __restrict__ h_eval) {
    __shared__ float bufa[512], bufb[512], feat[MAX_S * 64], hin[64 + MAX_N];
    const long long e = blockIdx.x;
    const int S = a.S, PF = w.per_feature_dim, FD = w.feature_dim, n = a.n;
    for (int s = 0; s < S; ++s) {
        for (int i = threadIdx.x; i < 512; i += 256) bufa[i] = unordered(max_keys[(e * S + s) * 512 + i]);
        __syncthreads();
        dense(w.fc2_w, w.fc2_b, bufa, bufb, 512, 256, true);
        dense(w.fc3_w, w.fc3_b, bufb, feat + s * PF, 256, PF, false);
    }
    dense(w.e1_w, w.e1_b, feat, bufa, S * PF, 512, true);
    dense(w.e2_w, w.e2_b, bufa, bufb, 512, 256, true);
    const long long b = e / a.evals;
    const int k = (int)(e % a.evals);
    for (int i = threadIdx.x; i < n; i += 256) hin[i] = __ldg(a.q + b * n + i) + ((k > 0 && i == k - 1) ? a.delta : 0.f);
    __syncthreads();
    dense(w.e4_w, w.e4_b, bufb, hin + n, 256, FD, false);
    dense(w.h_in_w, w.h_in_b, hin, bufa, n + FD, w.hidden, true);
    float *cur = bufa, *nxt = bufb;
    for (int l = 0; l < w.layers; ++l) {
        dense(w.h_hid_w + (long long)l * w.hidden * w.hidden, w.h_hid_b + (long long)l * w.hidden, cur, nxt, w.hidden, w.hidden, true);
        float* t = cur; cur = nxt; nxt = t;
    }
    dense(w.h_out_w, w.h_out_b, cur, nxt, w.hidden, 1, false);
    if (threadIdx.x == 0) h_eval[e] = nxt[0];
}

// h [B] = evaluation 0 of every state;  J [B, n] = (h_k - h_0) / (2 d)   (neural_lidar_cbf_controller.py:205-213)
static __global__ void finish_kernel(const float* __restrict__ h_eval, long long B, int n, int evals, float delta,
                                     float* __restrict__ h, float* __restrict__ J) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B * evals; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / evals;
        const int k = (int)(i % evals);
        const float h0 = h_eval[b * evals];
        if (k == 0) { if (h) h[b] = h0; }
        else if (J) J[b * n + k - 1] = (h_eval[i] - h0) / (2.f * delta);
    }
}

}  // namespace lidar
}  // namespace cbfrrt

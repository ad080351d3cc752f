// DRAFT -- NEVER RUN ON HARDWARE (written at the end of round 1 with no GPU minutes left; compile-checked for sm_100a only).
// First tcgen05 version of the PointNet backbone of the reference's LiDAR barrier:
//     per point  x[16] = [p (3) | n (3) | one-hot sensor (7) | 0 0 0]  ->  64 -> 128 -> 512, LeakyReLU(0.1) after every layer,
//     max over the R points of a (state evaluation, sensor) group            neural_cbf/controllers/utils/pointnet.py:20-66
// followed by a small FP32 kernel for 512 -> 256 -> 64 per sensor, 448 -> 512 -> 256 -> 32, and the 39-48-48-48-1 head
// (pointnet.py:42-48,78-97, neural_lidar_cbf_controller.py:66-95,115-139).
//
// Correctness-first layout, built from the PTX wrappers of csrc/mlp_tc.cuh (same descriptors, same 3-term TF32 split):
//   * one CTA = 128 threads = one tile of 128 points (M of the MMA), thread t <-> point t <-> TMEM lane t;
//   * TMEM: D in columns 0..127, A_hi in columns 128..255 (K <= 128);  A_lo: shared-memory tile [k/4][row] float4;
//   * B (weights) in shared memory as hi / lo tiles in the no-swizzle K-major layout W4[k/4][n], PRE-TILED by the host
//     (pointnet_tc_draft_check.py) so that staging is a linear copy; layer 3 is streamed in 8 chunks of 64 output columns;
//   * layer-3 epilogue: bias, LeakyReLU, then the max over the tile's 128 rows per column with REDUX on the ordered-int
//     image of the float and one atomicMax per (warp, column) into out_max[group][512] (int keys, preset to INT_MIN).
// Requires R % 128 == 0 (a tile never straddles two groups).  What round 2 should change first: double-buffer the layer-3
// chunks with TMA, two tiles per CTA, and the transposed (channels-on-lanes) product that needs no cross-row max at all.
#include "../../../cbf-inc-rrt_b200/csrc/mlp_tc.cuh"

using namespace cbfrrt;
using namespace cbfrrt::tc;

namespace pn {
constexpr int PT = 128;                                   // points per tile
constexpr int K1 = 16, N1 = 64, N2 = 128, N3 = 512, NCH = 64, CHUNKS = N3 / NCH;
constexpr int COLD = 0, COLA = 128;                       // TMEM columns of D and A_hi
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

// x [n_points, 16] (point-major, already in the sensor frame, one-hot appended, zero padded)
// w1t / w2t / w3t: pre-tiled [hi | lo] tiles (w3t: 8 chunks of [hi | lo] 128 x 64); bias = [b1 | b2 | b3]
// out_max [n_points / R, 512] int keys (ordered image of the float), preset to INT_MIN
__global__ void __launch_bounds__(PT, 1) backbone_kernel(const float* __restrict__ x, long long n_points, int R,
                                                         const float* __restrict__ w1t, const float* __restrict__ w2t,
                                                         const float* __restrict__ w3t, const float* __restrict__ bias,
                                                         int* __restrict__ out_max) {
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

    const long long n_tiles = n_points / PT;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long p = tile * PT + c.t;
        float v[N2];
        {   // layer 1: K = 16, N = 64
            const float4* xp = reinterpret_cast<const float4*>(x + p * K1);
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float4 q = __ldg(xp + j); v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w; }
            store_a<K1>(c, v);
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
        int* out = out_max + (tile * PT / R) * N3;
        for (int ch = 0; ch < CHUNKS; ++ch) {               // layer 3: K = 128, N = 64 per chunk
            __syncthreads();                                // previous chunk's MMAs are complete (gemm waited on them)
            const float4* src = reinterpret_cast<const float4*>(w3t + (long long)ch * 2 * N2 * NCH);
            float4* dst = reinterpret_cast<float4*>(sm + S_W3);
            for (int i = c.t; i < 2 * N2 * NCH / 4; i += PT) dst[i] = __ldg(src + i);
            fence_proxy_async();
            gemm<N2, NCH>(c, sm + S_W3, sm + S_W3 + N2 * NCH);
            float d[NCH];
            load_d<NCH>(c, d);
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int key = __reduce_max_sync(0xffffffffu, ordered(leaky(d[k] + sm[S_B + N1 + N2 + ch * NCH + k])));
                if ((c.t & 31) == (k & 31)) atomicMax(out + ch * NCH + k, key);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (c.t < 32) tmem_dealloc(c.tmem, 512);
}

// ---- everything after the max-pool, FP32 on the CUDA cores: one CTA (256 threads, 8 warps) per state evaluation.
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

struct HeadW {   // raw torch tensors (row-major [out][in]) on the device
    const float *fc2_w, *fc2_b, *fc3_w, *fc3_b;                    // backbone_fc 512->256->PF
    const float *e1_w, *e1_b, *e2_w, *e2_b, *e4_w, *e4_b;          // encoder S*PF->512->256->FD
    const float *h0_w, *h0_b, *h1_w, *h1_b, *h2_w, *h2_b, *ho_w, *ho_b;   // h_nn (n+FD)->48->48->48->1
};

__global__ void __launch_bounds__(256) head_kernel(const int* __restrict__ max_keys, const float* __restrict__ q, int n, int S,
                                                   int PF, int FD, HeadW w, float* __restrict__ h_out) {
    __shared__ float a[512], b[512], feat[7 * 64], hin[64];
    const long long e = blockIdx.x;
    for (int s = 0; s < S; ++s) {
        for (int i = threadIdx.x; i < 512; i += 256) {
            const int key = max_keys[(e * S + s) * 512 + i];
            a[i] = __int_as_float(key >= 0 ? key : key ^ 0x7fffffff);
        }
        __syncthreads();
        dense(w.fc2_w, w.fc2_b, a, b, 512, 256, true);
        dense(w.fc3_w, w.fc3_b, b, feat + s * PF, 256, PF, false);
    }
    dense(w.e1_w, w.e1_b, feat, a, S * PF, 512, true);
    dense(w.e2_w, w.e2_b, a, b, 512, 256, true);
    for (int i = threadIdx.x; i < n; i += 256) hin[i] = q[e * n + i];
    __syncthreads();
    dense(w.e4_w, w.e4_b, b, hin + n, 256, FD, false);
    dense(w.h0_w, w.h0_b, hin, a, n + FD, 48, true);
    dense(w.h1_w, w.h1_b, a, b, 48, 48, true);
    dense(w.h2_w, w.h2_b, b, a, 48, 48, true);
    dense(w.ho_w, w.ho_b, a, b, 48, 1, false);
    if (threadIdx.x == 0) h_out[e] = b[0];
}
}  // namespace pn

extern "C" int pn_backbone(const float* x, long long n_points, int R, const float* w1t, const float* w2t, const float* w3t,
                           const float* bias, int* out_max, void* stream) {
    if (n_points % pn::PT || R % pn::PT) return -2;
    const size_t smem = pn::S_TOTAL * sizeof(float);
    cudaFuncSetAttribute(pn::backbone_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = n_points / pn::PT;
    pn::backbone_kernel<<<(unsigned)(tiles < sms ? tiles : sms), pn::PT, smem, (cudaStream_t)stream>>>(x, n_points, R, w1t, w2t, w3t, bias, out_max);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

extern "C" int pn_head(const int* max_keys, const float* q, long long n_eval, int n, int S, int PF, int FD, const pn::HeadW* w,
                       float* h_out, void* stream) {
    if (S > 7 || PF > 64 || n + FD > 64) return -3;
    pn::head_kernel<<<(unsigned)n_eval, 256, 0, (cudaStream_t)stream>>>(max_keys, q, n, S, PF, FD, *w, h_out);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

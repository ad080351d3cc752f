// mlp_tc.cuh -- the barrier MLP (value + exact reverse-mode Jacobian) on the 5th-generation tensor cores.
//
// A group of 128 threads (4 warps) owns a tile of 128 states = the M dimension of one tcgen05.mma; a CTA runs
// GG such groups that share the staged weights, one CTA per SM, 128 TMEM columns per tile.  GG = 4 (16 warps per SM at
// 128 registers) for the kernels whose geometry stage lives in shared memory; GG = 3 for the broad-phase-grid
// variants, which need the L1 capacity that a fourth A_lo tile (24 KB) would take from the grid look-ups.
// Every layer of
//     h = w3 . relu(W2 relu(W1 relu(W0 x + b0) + b1) + b2) + b3        (neural_mindis_cbf_controller.py:41-53)
// and of its reverse-mode Jacobian (:93-96) is a [128 x K] . [K x N] GEMM with K in {8, 48}, N in {16, 48}:
//   * A operand (activations / back-propagated gradient): the "hi" part lives in TENSOR MEMORY -- thread t writes its
//     row with tcgen05.st 32x32b (lane t, one column per K element); the small "lo" part goes to a per-group
//     shared-memory tile in the same K-major core-matrix layout as the weights (a third TMEM operand region would cap
//     the CTA at 3 tiles = 3 warps per scheduler, and the kernel is issue/latency bound, not smem bound);
//   * B operand (weights) in shared memory, canonical no-swizzle K-major layout W4[k/4][n] (core matrix = 8 rows
//     x 16 bytes).  kind::tf32 returned zeros for MN-major (transposed) B in every descriptor variant probed
//     (tools/tc_probe.cu), so the backward GEMMs read a second, transposed copy of each weight matrix;
//   * accumulator D [128 lanes x 48 columns] fp32 in TMEM, read back with tcgen05.ld 32x32b (thread t <-> lane t),
//     bias / ReLU / ReLU-mask epilogue in registers;
//   * kind::tf32 with the 3-term split  a.w ~= a_hi.w_hi + a_lo.w_hi + a_hi.w_lo  (hi = cvt.rna.tf32, lo = a - hi,
//     exact in fp32): ~2^-22 relative per product, i.e. fp32-grade, which the 1e-5 tolerance on h / J / u needs
//     (a single TF32 pass is 2^-11: measured 5.6e-3 absolute on |D| ~ 8 in tools/tc_probe.cu);
//   * one thread per group issues the MMAs of a layer and commits them to the group's mbarrier.
// The 1 x 48 output layer and the seed of the backward pass stay on the CUDA cores (N = 1 is not an MMA shape).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mlp_qp.cuh"

namespace cbfrrt {
namespace tc {

constexpr int M_TILE = 128;
constexpr int G_MAX = 4;                   // tile groups per CTA: template parameter GG <= G_MAX (512 threads)
constexpr int H = 48;                      // hidden width (validated on the host: only 48 x 2 is built)
constexpr int KIN = 16;                    // padded N of the last backward GEMM (dh/dx has N_IN <= 8 real columns)
constexpr int COLS_PER_TILE = 128;         // D 0..47 | pad | A_hi 64..111 | pad
constexpr int COL_D = 0, COL_AHI = 64;
constexpr int ALO_TILE = M_TILE * H;       // floats of one group's A_lo tile: [k/4][row] float4
constexpr int TMEM_COLS = 512;

// ---- shared-memory carve-up of the tensor-core MLP region (floats, every sub-region 16-byte aligned)
template <int L, int GG>
struct Layout {
    static constexpr int MAT = H * H;                         // one 48 x 48 tile (hi or lo)
    static constexpr int W_F = 0;                             // forward tiles  [l][hi|lo]  W4[k/4][n]
    static constexpr int W_T = W_F + L * 2 * MAT;             // transposed tiles [l][hi|lo]  (W^T)4[j/4][k]
    static constexpr int WIN_F = W_T + L * 2 * MAT;           // [hi|lo] [2][48] float4   (K = 8, N = 48)
    static constexpr int WIN_T = WIN_F + 2 * 8 * H;           // [hi|lo] [12][16] float4  (K = 48, N = 16)
    static constexpr int BIAS = WIN_T + 2 * KIN * H;          // b_in[48], b_l[48] x L, w_out[48], b_out (padded to 4)
    static constexpr int MBAR = BIAS + (L + 2) * H + 4;       // GG 64-bit mbarriers
    static constexpr int TMEM_PTR = MBAR + 2 * GG + 2;
    static constexpr int A_LO = (TMEM_PTR + 4 + 31) & ~31;    // GG tiles of A_lo, 128-byte aligned
    static constexpr int TOTAL = A_LO + GG * ALO_TILE;
};

// ---- PTX wrappers
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // UMMA shared-memory descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), no swizzle
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t instr_desc(int n) {
    // c_format F32 [4,6) = 1, a_format TF32 [7,10) = 2, b_format TF32 [10,13) = 2, A/B K-major, N>>3 [17,23), M>>4 [24,29)
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M_TILE >> 4) << 24);
}
// D[tmem] (+)= A[tmem] . B[smem]^T
// Called by a WHOLE warp with warp-uniform operands; one elected lane issues.  Keeping the operands warp-uniform lets
// ptxas hold the descriptors in uniform registers (a divergent `if (tid == 0)` costs an ELECT/R2UR loop per MMA).
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T  (both operands through shared-memory descriptors)
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}\n" ::"r"(smem_addr(bar))
        : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_init1(uint64_t* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"   // %3: suspend-time hint (ns): sleep, don't spin
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity), "r"(20000u)
            : "memory");
    } while (!done);
}
// barrier over the 128 threads of one tile group (named barriers 1..G; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(M_TILE) : "memory"); }
// true if `pred` holds for any thread of the group
__device__ __forceinline__ bool group_any(int g, bool pred) {
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 p, %1, 0;\n"
        "bar.red.or.pred q, %2, %3, p;\n"
        "selp.u32 %0, 1, 0, q;\n"
        "}\n"
        : "=r"(r)
        : "r"((uint32_t)pred), "r"(g + 1), "r"(M_TILE)
        : "memory");
    return r != 0;
}

// ---- per-thread view of its tile group
template <int L, int GG>
struct Ctx {
    static constexpr int NT = GG * M_TILE;
    float* base;       // start of the TC region in shared memory
    uint32_t tmem;     // TMEM address of column 0 of this group's slice (lane field = 0)
    uint32_t lane;     // this thread's warp lane-quarter base, pre-shifted (<< 16)
    uint32_t phase;    // parity of the next MMA-completion wait
    int g, t;          // group index, row within the tile
    __device__ __forceinline__ float* w_f(int l, int lo) const { return base + Layout<L, GG>::W_F + (2 * l + lo) * Layout<L, GG>::MAT; }
    __device__ __forceinline__ float* w_t(int l, int lo) const { return base + Layout<L, GG>::W_T + (2 * l + lo) * Layout<L, GG>::MAT; }
    __device__ __forceinline__ float* win_f(int lo) const { return base + Layout<L, GG>::WIN_F + lo * 8 * H; }
    __device__ __forceinline__ float* win_t(int lo) const { return base + Layout<L, GG>::WIN_T + lo * KIN * H; }
    __device__ __forceinline__ const float* bias(int l) const { return base + Layout<L, GG>::BIAS + l * H; }  // 0: b_in, 1..L, L+1: w_out, L+2: b_out
    __device__ __forceinline__ float4* alo() const { return reinterpret_cast<float4*>(base + Layout<L, GG>::A_LO + g * ALO_TILE); }
    __device__ __forceinline__ uint64_t* mbar() const { return reinterpret_cast<uint64_t*>(base + Layout<L, GG>::MBAR) + g; }
    __device__ __forceinline__ uint32_t* tmem_slot() const { return reinterpret_cast<uint32_t*>(base + Layout<L, GG>::TMEM_PTR); }
};

// tile index of element (n, k) in the K-major no-swizzle layout with NROWS rows:  W4[k/4][n]
__device__ __forceinline__ int tile_idx(int n, int k, int nrows) { return ((k >> 2) * nrows + n) * 4 + (k & 3); }

// Stage the packed fp32 weights (global memory, layout of include/cbfrrt.h) as hi / lo TF32 tiles (forward and
// transposed), allocate TMEM, init the MMA mbarriers.  Called by all GG * 128 threads once; ends with a __syncthreads().
template <int N_IN, int L, int GG>
__device__ __forceinline__ void setup(Ctx<L, GG>& c, float* region, const float* __restrict__ wg) {
    using LY = NetLayout<N_IN, H, L>;
    const int tid = threadIdx.x;
    c.base = region;
    c.phase = 0;
    c.g = tid / M_TILE;
    c.t = tid % M_TILE;
    for (int i = tid; i < H * 8; i += GG * M_TILE) {               // input layer, forward: n = 0..47, k = 0..7
        const int n = i / 8, k = i % 8;
        const float w = (k < N_IN) ? __ldg(wg + LY::W_IN + n * LY::IN_PAD + k) : 0.f;
        const float hi = rna_tf32(w);
        c.win_f(0)[tile_idx(n, k, H)] = hi;
        c.win_f(1)[tile_idx(n, k, H)] = w - hi;
    }
    for (int i = tid; i < KIN * H; i += GG * M_TILE) {             // input layer, transposed: n' = 0..15 (input index), k' = 0..47
        const int np = i / H, kp = i % H;
        const float w = (np < N_IN) ? __ldg(wg + LY::W_IN + kp * LY::IN_PAD + np) : 0.f;
        const float hi = rna_tf32(w);
        c.win_t(0)[tile_idx(np, kp, KIN)] = hi;
        c.win_t(1)[tile_idx(np, kp, KIN)] = w - hi;
    }
    for (int l = 0; l < L; ++l) {
        const float* W = wg + LY::HID0 + l * LY::HID_STRIDE;
        for (int i = tid; i < H * H; i += GG * M_TILE) {
            const int n = i / H, k = i % H;
            const float w = __ldg(W + i);
            const float hi = rna_tf32(w), lo = w - hi;
            // forward tiles hold W / 2: the epilogue hands 2 * relu(s) = s + |s| to the next layer (one FADD on the
            // FMA pipe instead of an FMNMX on the half-rate ALU pipe); halving is exact, also for the hi / lo split
            c.w_f(l, 0)[tile_idx(n, k, H)] = 0.5f * hi;
            c.w_f(l, 1)[tile_idx(n, k, H)] = 0.5f * lo;
            c.w_t(l, 0)[tile_idx(k, n, H)] = hi;             // (W^T)[k][n] = W[n][k]
            c.w_t(l, 1)[tile_idx(k, n, H)] = lo;
        }
    }
    float* b = region + Layout<L, GG>::BIAS;
    for (int i = tid; i < H; i += GG * M_TILE) {
        b[i] = __ldg(wg + LY::B_IN + i);
        for (int l = 0; l < L; ++l) b[(l + 1) * H + i] = __ldg(wg + LY::HID0 + l * LY::HID_STRIDE + H * H + i);
        b[(L + 1) * H + i] = __ldg(wg + LY::W_OUT + i);
    }
    if (tid == 0) {
        b[(L + 2) * H] = __ldg(wg + LY::B_OUT);
        for (int g = 0; g < GG; ++g) mbar_init1(reinterpret_cast<uint64_t*>(region + Layout<L, GG>::MBAR) + g);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) tmem_alloc(c.tmem_slot(), TMEM_COLS);      // warp 0 allocates (and later frees)
    fence_proxy_async();                                     // weight tiles (generic proxy) -> tensor-core (async) proxy
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    c.tmem = *c.tmem_slot() + c.g * COLS_PER_TILE;
    c.lane = (uint32_t)(c.t & ~31) << 16;
}

template <int L, int GG>
__device__ __forceinline__ void teardown(Ctx<L, GG>& c) {
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(*c.tmem_slot(), TMEM_COLS);
}

// write this thread's K values as the hi row (tensor memory) and the lo row (shared-memory tile) of the A operand
// The "hi" operand is the raw fp32 value: the tensor core reads only the TF32 bits, i.e. hi = trunc_tf32(v), for free.
// lo = v - trunc_tf32(v) (exact in fp32; one LOP3 + one FADD).  With the weights split by round-to-nearest at staging
// time the 3-term product is accurate to ~2^-20 relative per term -- measured better than an fp32 FMA chain.
template <int K, int L, int GG>
__device__ __forceinline__ void store_a(const Ctx<L, GG>& c, const float* v) {
    static_assert(K == 8 || K % 16 == 0, "");
    const uint32_t row = c.tmem + c.lane;
    float4* alo = c.alo() + c.t;               // K-major tile [k/4][row]: consecutive rows are consecutive float4
#pragma unroll
    for (int kc = 0; kc < K / 4; ++kc) {
        float4 lo;
        lo.x = v[4 * kc] - __uint_as_float(__float_as_uint(v[4 * kc]) & 0xffffe000u);
        lo.y = v[4 * kc + 1] - __uint_as_float(__float_as_uint(v[4 * kc + 1]) & 0xffffe000u);
        lo.z = v[4 * kc + 2] - __uint_as_float(__float_as_uint(v[4 * kc + 2]) & 0xffffe000u);
        lo.w = v[4 * kc + 3] - __uint_as_float(__float_as_uint(v[4 * kc + 3]) & 0xffffe000u);
        alo[kc * M_TILE] = lo;
    }
    if constexpr (K == 8) {
        tmem_st8(row + COL_AHI, v);
    } else {
#pragma unroll
        for (int j = 0; j < K / 16; ++j) tmem_st16(row + COL_AHI + 16 * j, v + 16 * j);
    }
    fence_proxy_async();                       // A_lo tile (generic proxy) -> tensor core (async proxy)
    tmem_st_wait();
}

// D[128 x N] = A[128 x K] . B^T with the 3-term split.  B tile: K-major no-swizzle with NROWS (= N) rows:
// LBO = NROWS*16 B between K-chunks of 4, SBO = 128 B between 8-row groups, a K = 8 step advances 2 chunks.
template <int K, int N, int L, int GG>
__device__ __forceinline__ void gemm(Ctx<L, GG>& c, const float* w_hi, const float* w_lo) {
    fence_before_sync();
    group_sync(c.g);
    if (c.t < 32) {   // warp 0 of the group, warp-uniform: one elected lane issues (see mma_tf32_ts)
        fence_after_sync();
        constexpr uint32_t idesc = instr_desc(N);
        constexpr uint32_t B_LBO = N * 16, B_SBO = 128, B_STEP = 2 * N * 16;
        const uint32_t b_hi = smem_addr(w_hi), b_lo = smem_addr(w_lo);
        const uint32_t a_hi = c.tmem + COL_AHI, a_lo = smem_addr(c.alo());
        constexpr uint32_t A_LBO = M_TILE * 16, A_SBO = 128, A_STEP = 2 * M_TILE * 16;
#pragma unroll
        for (int ks = 0; ks < K / 8; ++ks)       // a_hi . w_hi
            mma_tf32_ts(c.tmem + COL_D, a_hi + 8 * ks, smem_desc(b_hi + ks * B_STEP, B_LBO, B_SBO), idesc, ks ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < K / 8; ++ks)       // a_hi . w_lo
            mma_tf32_ts(c.tmem + COL_D, a_hi + 8 * ks, smem_desc(b_lo + ks * B_STEP, B_LBO, B_SBO), idesc, 1u);
#pragma unroll
        for (int ks = 0; ks < K / 8; ++ks)       // a_lo . w_hi  (A from shared memory)
            mma_tf32_ss(c.tmem + COL_D, smem_desc(a_lo + ks * A_STEP, A_LBO, A_SBO), smem_desc(b_hi + ks * B_STEP, B_LBO, B_SBO), idesc, 1u);
        mma_commit(c.mbar());
        // only the issuing warp polls the mbarrier; the other three block in hardware at the named barrier below
        // instead of spinning (the try_wait loops of 16 warps were ~8 % of all issued instructions)
        mbar_wait_parity(c.mbar(), c.phase);
        fence_before_sync();
    }
    c.phase ^= 1u;
    group_sync(c.g);
    fence_after_sync();
}

// read this thread's row (lane) of the accumulator: NCOL columns (multiple of 16)
template <int NCOL, int L, int GG>
__device__ __forceinline__ void load_d(const Ctx<L, GG>& c, float* v) {
    const uint32_t row = c.tmem + c.lane + COL_D;
#pragma unroll
    for (int j = 0; j < NCOL / 16; ++j) tmem_ld16(row + 16 * j, v + 16 * j);
    tmem_ld_wait();
}

// epilogue of a forward layer: s = v + bias;  v <- 2 * relu(s) = s + |s|;  returns the ReLU mask (bit k = s_k > 0).
// The mask bit is the sign of (0 - s): exactly (s > 0), also for s = +-0; one FADD + one funnel shift per element.
template <int L, int GG>
__device__ __forceinline__ uint64_t bias_relu2(const Ctx<L, GG>& c, int layer, float (&v)[H]) {
    static_assert(H == 48, "three 16-bit mask chains");
    uint32_t m[3] = {0, 0, 0};   // independent funnel-shift chains (16 elements each) instead of one 48-long one
    const float4* b4 = reinterpret_cast<const float4*>(c.bias(layer));
#pragma unroll
    for (int i = 3; i >= 0; --i) {             // descending inside a chain: bit k lands at position k % 16
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            const int k4 = 4 * w + i;
            const float4 b = b4[k4];
            const float s3 = v[4 * k4 + 3] + b.w, s2 = v[4 * k4 + 2] + b.z, s1 = v[4 * k4 + 1] + b.y, s0 = v[4 * k4] + b.x;
            m[w] = __funnelshift_l(__float_as_uint(0.f - s3), m[w], 1); m[w] = __funnelshift_l(__float_as_uint(0.f - s2), m[w], 1);
            m[w] = __funnelshift_l(__float_as_uint(0.f - s1), m[w], 1); m[w] = __funnelshift_l(__float_as_uint(0.f - s0), m[w], 1);
            v[4 * k4] = s0 + fabsf(s0); v[4 * k4 + 1] = s1 + fabsf(s1); v[4 * k4 + 2] = s2 + fabsf(s2); v[4 * k4 + 3] = s3 + fabsf(s3);
        }
    }
    return ((uint64_t)m[2] << 32) | ((uint64_t)(m[1] << 16) | m[0]);
}

// value h and Jacobian dh/dx for this thread's x = [q, o]; every thread of the tile group must call it (idle
// rows pass zeros).  Same contract as mlp_h_jac (mlp_qp.cuh).
template <int N_IN, int L, int GG>
__device__ __forceinline__ void mlp_h_jac(Ctx<L, GG>& c, const float (&x)[N_IN], float& h_out, float (&jac)[N_IN]) {
    static_assert(N_IN <= 8, "first layer is one K = 8 MMA step");
    float v[H];
    uint64_t mask[L + 1];
    // ---- input layer: K = 8 (zero padded), N = 48
    {
        float xin[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) xin[k] = (k < N_IN) ? x[k < N_IN ? k : 0] : 0.f;
        store_a<8>(c, xin);
        gemm<8, H>(c, c.win_f(0), c.win_f(1));
        load_d<H>(c, v);
        mask[0] = bias_relu2(c, 0, v);
    }
    // ---- hidden layers
#pragma unroll
    for (int l = 0; l < L; ++l) {
        store_a<H>(c, v);
        gemm<H, H>(c, c.w_f(l, 0), c.w_f(l, 1));
        load_d<H>(c, v);
        mask[l + 1] = bias_relu2(c, l + 1, v);
    }
    // ---- output layer (N = 1) and seed of the backward pass on the CUDA cores
    {
        const float4* w4 = reinterpret_cast<const float4*>(c.bias(L + 1));
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        const uint64_t m = mask[L];
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
            const float4 w = w4[k4];
            s0 = fmaf(w.x, v[4 * k4], s0); s1 = fmaf(w.y, v[4 * k4 + 1], s1);
            s2 = fmaf(w.z, v[4 * k4 + 2], s2); s3 = fmaf(w.w, v[4 * k4 + 3], s3);
            v[4 * k4] = ((m >> (4 * k4)) & 1) ? w.x : 0.f;
            v[4 * k4 + 1] = ((m >> (4 * k4 + 1)) & 1) ? w.y : 0.f;
            v[4 * k4 + 2] = ((m >> (4 * k4 + 2)) & 1) ? w.z : 0.f;
            v[4 * k4 + 3] = ((m >> (4 * k4 + 3)) & 1) ? w.w : 0.f;
        }
        h_out = fmaf(0.5f, (s0 + s1) + (s2 + s3), c.bias(L + 2)[0]);   // v held 2 * relu(s)
    }
    // ---- backward through the hidden layers: g_prev = relu'_prev * (W^T g)
#pragma unroll
    for (int l = L - 1; l >= 0; --l) {
        store_a<H>(c, v);
        gemm<H, H>(c, c.w_t(l, 0), c.w_t(l, 1));
        load_d<H>(c, v);
        const uint64_t m = mask[l];
#pragma unroll
        for (int k = 0; k < H; ++k) v[k] = ((m >> k) & 1) ? v[k] : 0.f;
    }
    // ---- input layer backward: jac = W_in^T g   (N = KIN = 16 columns, the first N_IN are real)
    store_a<H>(c, v);
    gemm<H, KIN>(c, c.win_t(0), c.win_t(1));
    float j16[KIN];
    load_d<KIN>(c, j16);
#pragma unroll
    for (int k = 0; k < N_IN; ++k) jac[k] = j16[k];
}

}  // namespace tc
}  // namespace cbfrrt

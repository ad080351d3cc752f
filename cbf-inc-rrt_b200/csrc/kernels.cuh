// kernels.cuh -- sm_100a kernels and their launchers (templates).  Included by the per-family translation units
// (k_*.cu, compiled in parallel) and by abi.cu, which holds the device-pointer C-ABI of include/cbfrrt.h.
//
// Launch shape: CTAs of NT=128 threads, persistent grid (SMs x resident CTAs), grid-stride over tiles of NT
// states so that the per-CTA staging (scene + weights by TMA bulk copy, base-link distance) is paid once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "../../include/cbfrrt.h"
#include "core.cuh"
#include "mlp_qp.cuh"
#include "mlp_tc.cuh"

namespace cbfrrt {

constexpr int NT = 128;
#ifndef CBF_COLLIDE_MIN_CTAS
#define CBF_COLLIDE_MIN_CTAS 5   // <= 96 registers: 20 warps per SM.  The field path is bound by the latency of its L2 gathers,
                                 // not by registers: measured per 2^23 Panda states, 256 boxes (4 / 5 / 6 / 7 / 8 CTAs per SM):
                                 // masks 2.15 / 2.09 / 1.97 / 2.16 / 2.30 ms, masks + datax 3.23 / 3.14 / 3.16 / 3.13 / 3.28 ms,
                                 // 2^22 edges x 33 points 25.3 / 22.4 / 22.3 / 22.7 / 25.2 ms
#endif

// ------------------------------------------------------------------ TMA bulk copy + mbarrier (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// ------------------------------------------------------------------ shared-memory staging
struct SceneArgs {
    const float* boxes;
    const float* spheres;
    int n_boxes, n_spheres, floor;
    const uint16_t* grid;
    unsigned long long* stats;   // optional [4]: states, exactly swept spheres, exact pair evaluations, field look-ups
};
struct NetArgs {
    const float* weights;  // packed, NetLayout<>::TOTAL floats
    const float* K;
    const float* u_lo;
    const float* u_hi;
    const float* goal;
    long long goal_rows;
    float lam, penalty;
    const float* u_ref;  // optional [B, n]
};

// dynamic smem carve-up (all offsets in floats, every region 16-byte aligned):
//   [mbarrier 4][base_min reduce 12][boxes 8*n_boxes][spheres 4*n_spheres][weights][params][scratch H*NT]
struct Smem {
    uint64_t* bar;
    float* red;
    float* boxes;
    float* spheres;
    float* weights;
    float* params;
    float* scratch;
};
__host__ __device__ inline size_t smem_floats(int n_boxes, int n_spheres, int w_floats, int p_floats, int scr_floats) {
    return 32 + (size_t)8 * n_boxes + (size_t)4 * n_spheres + w_floats + p_floats + scr_floats;
}
__device__ __forceinline__ Smem carve(float* base, int n_boxes, int n_spheres, int w_floats, int p_floats) {
    Smem s;
    s.bar = reinterpret_cast<uint64_t*>(base);
    s.red = base + 4;          // one float per warp of the CTA (<= 28 warps)
    s.boxes = base + 32;
    s.spheres = s.boxes + 8 * n_boxes;
    s.weights = s.spheres + 4 * n_spheres;
    s.params = s.weights + w_floats;
    s.scratch = s.params + p_floats;
    return s;
}

// Stage scene (+ optional weights) with TMA bulk copies signalled on one mbarrier; small parameter vectors are
// copied by the threads.  Returns after everything is visible to the whole CTA.
template <int N, int NTH = NT>
__device__ __forceinline__ void stage(const Smem& s, const SceneArgs& sc, const NetArgs* net, int w_floats) {
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(s.bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bb = 32u * sc.n_boxes, sb = 16u * sc.n_spheres, wb = net ? 4u * w_floats : 0u;
        mbar_arrive_expect_tx(s.bar, bb + sb + wb);
        if (bb) tma_bulk_g2s(s.boxes, sc.boxes, bb, s.bar);
        if (sb) tma_bulk_g2s(s.spheres, sc.spheres, sb, s.bar);
        if (wb) tma_bulk_g2s(s.weights, net->weights, wb, s.bar);
    }
    if (net) {
        using PL = ParamLayout<N>;
        for (int i = tid; i < N * N; i += NTH) s.params[PL::K + i] = net->K ? net->K[i] : 0.f;
        for (int i = tid; i < N; i += NTH) {
            s.params[PL::LO + i] = net->u_lo[i];
            s.params[PL::HI + i] = net->u_hi[i];
        }
    }
    mbar_wait(s.bar, 0);
    __syncthreads();
}

template <class RB, int NTH = NT>
__device__ __forceinline__ float base_min_cta(const Smem& s, const SceneS& sc) {
    float m = base_min_partial<RB>(sc, threadIdx.x, NTH);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) s.red[threadIdx.x >> 5] = m;
    __syncthreads();
    float r = s.red[0];
#pragma unroll
    for (int w = 1; w < NTH / 32; ++w) r = fminf(r, s.red[w]);
    return r;
}

__device__ __forceinline__ SceneS scene_view(const Smem& s, const SceneArgs& sc, float thr_soft, float thr_hard) {
    SceneS v;
    v.cull_floor = fmaxf(thr_soft, thr_hard);
    v.thr_min = fminf(thr_soft, thr_hard);
    v.boxes = reinterpret_cast<const float4*>(s.boxes);
    v.spheres = reinterpret_cast<const float4*>(s.spheres);
    v.n_boxes = sc.n_boxes; v.n_spheres = sc.n_spheres; v.floor = sc.floor;
    v.grid = sc.grid;
    return v;
}

// ------------------------------------------------------------------ kernels
template <class RB>
__global__ void __launch_bounds__(NT) fk_kernel(const float* __restrict__ q, long long q_stride, long long B,
                                                float* __restrict__ pos, float* __restrict__ rot) {
    for (long long b = (long long)blockIdx.x * NT + threadIdx.x; b < B; b += (long long)gridDim.x * NT) {
        float qq[RB::N];
#pragma unroll
        for (int k = 0; k < RB::N; ++k) qq[k] = __ldg(q + b * q_stride + k);
        float* po = pos + b * RB::NLINKS * 3;
        float* ro = rot + b * RB::NLINKS * 9;
        fk_all<RB>(qq, [&](int l, const Frame& f) {
#pragma unroll
            for (int k = 0; k < 3; ++k) po[l * 3 + k] = f.p[k];
#pragma unroll
            for (int k = 0; k < 9; ++k) ro[l * 9 + k] = f.R[k];
        });
    }
}

struct CollideOutPtrs {
    uint8_t* safe;
    uint8_t* unsafe;
    float* datax;
    int* arg_link;
    int* arg_obs;
    float* mins;
};

template <class RB>
__device__ __forceinline__ void write_collide(const CollideOutPtrs& o, long long b, const float (&q)[RB::N],
                                              const CollideResult<RB>& r, float thr_soft, float thr_hard) {
    const bool unsafe = r.hard_min <= thr_hard;
    const bool safe = r.self_ok && (r.soft_min > thr_soft) && !unsafe;
    if (o.safe) o.safe[b] = safe;
    if (o.unsafe) o.unsafe[b] = unsafe;
    if (o.datax) {
        float* d = o.datax + b * (2 * RB::N + 1);
#pragma unroll
        for (int k = 0; k < RB::N; ++k) d[k] = q[k];
        d[RB::N] = r.o;
#pragma unroll
        for (int k = 0; k < RB::N; ++k) d[RB::N + 1 + k] = r.dodq[k];
    }
    if (o.arg_link) o.arg_link[b] = r.arg_sphere >= 0 ? RB::sph_link(r.arg_sphere) : -2;
    if (o.arg_obs) o.arg_obs[b] = r.arg_obs;
    if (o.mins) {
        o.mins[3 * b + 0] = r.self_min;
        o.mins[3 * b + 1] = r.soft_min;
        o.mins[3 * b + 2] = r.hard_min;
    }
}

// floats of the shared-memory sphere-centre scratch of the SCR instantiations (field path): 3 per non-base sphere and thread
template <class RB>
constexpr int scr_floats() { return 3 * (RB::NSPH - RB::sph_begin(0)) * NT; }

template <class RB, bool WANT_GRAD, bool WANT_SELFMIN, bool GRID, bool SCR = false>
__global__ void __launch_bounds__(NT, SCR ? 4 : CBF_COLLIDE_MIN_CTAS) collide_kernel(SceneArgs sc, const float* __restrict__ q, long long q_stride,
                                                     long long B, float thr_soft, float thr_hard, CollideOutPtrs out) {
    extern __shared__ __align__(16) float smem_raw[];
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, 0, 0);
    stage<RB::N>(s, sc, nullptr, 0);
    SceneS sv = scene_view(s, sc, thr_soft, thr_hard);
    if constexpr (SCR) { sv.scratch = s.scratch + threadIdx.x; sv.scratch_stride = NT; }
    const float base_min = base_min_cta<RB>(s, sv);
    unsigned long long n_states = 0, n_ref = 0, n_pairs = 0;
    for (long long b = (long long)blockIdx.x * NT + threadIdx.x; b < B; b += (long long)gridDim.x * NT) {
        float qq[RB::N];
#pragma unroll
        for (int k = 0; k < RB::N; ++k) qq[k] = __ldg(q + b * q_stride + k);
        CollideResult<RB> r;
        collide_state<RB, WANT_GRAD, WANT_SELFMIN, GRID, true, SCR>(qq, sv, base_min, r);
        write_collide<RB>(out, b, qq, r, thr_soft, thr_hard);
        n_states += 1; n_ref += r.n_refined; n_pairs += r.n_pairs;
    }
    if (sc.stats) {   // broad-phase statistics for the roofline accounting (bench.py); off (nullptr) in every other call
        atomicAdd(sc.stats + 0, n_states);
        atomicAdd(sc.stats + 1, n_ref);
        atomicAdd(sc.stats + 2, n_pairs);
        if (GRID && !WANT_SELFMIN) atomicAdd(sc.stats + 3, n_states * (unsigned long long)(RB::NSPH - RB::sph_begin(0)));
    }
}

// goal row (or, when the caller supplies u_ref, that row: cbf_filter_state then uses it verbatim)
template <int N>
__device__ __forceinline__ void load_goal(const NetArgs& net, long long b, float (&g)[N]) {
    const float* src = net.u_ref ? net.u_ref + b * N : net.goal + (net.goal_rows == 1 ? 0 : b * N);
#pragma unroll
    for (int k = 0; k < N; ++k) g[k] = __ldg(src + k);
}

// fused all-gather epilogue: peer-mapped replicas of the [G] mask and [G, n] controls (include/cbfrrt.h)
struct PeerOut {
    int n;
    int multicast;      // 1: valid[0] / u[0] are NVSwitch multicast addresses -> multimem.st (PTX ISA: multimem
                        //    addresses may only be accessed by multimem.*); rows of a warp are 32 consecutive, aligned rows
    long long row_offset;
    uint8_t* valid[CBFRRT_MAX_PEERS];
    float* u[CBFRRT_MAX_PEERS];
};

// multimem.st: one store, replicated by the switch into every rank's replica (NVLS).  ptxas lowers it to
// STG.E.STRONG.SYS on sm_100a -- the multicast happens in the fabric, by address -- so it shows as `multimem.st` in the
// PTX only (tests/test_cabi_cpu.py compiles this header to PTX and checks).
__device__ __forceinline__ void multimem_st_f32(float* addr, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void multimem_st_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int N>
__device__ __forceinline__ void write_peers(const PeerOut& po, long long b, bool safe, const float (&u)[N]) {
    const long long g = po.row_offset + b;
    if (po.multicast) {
        // mask bytes: multimem.st has no sub-word form, so the 32 rows of the warp are packed into 8 words (4 mask bytes
        // each) and lanes 0..7 store one word each.  The ABI guarantees full, 32-row-aligned warps in this mode.
        const unsigned bal = __ballot_sync(0xffffffffu, safe);
        const int lane = threadIdx.x & 31;
        if (lane < 8) {
            const unsigned nib = (bal >> (4 * lane)) & 0xfu;
            const unsigned word = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
            multimem_st_f32(reinterpret_cast<float*>(po.valid[0] + (g - lane) + 4 * lane), __uint_as_float(word));
        }
        float* dst = po.u[0] + g * N;
        if constexpr (N % 4 == 0) {
#pragma unroll
            for (int k = 0; k < N; k += 4) multimem_st_v4(dst + k, u[k], u[k + 1], u[k + 2], u[k + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < N; ++k) multimem_st_f32(dst + k, u[k]);
        }
        return;
    }
    for (int p = 0; p < po.n; ++p) {
        po.valid[p][g] = safe;
        float* dst = po.u[p] + g * N;
        if constexpr (N % 4 == 0) {
#pragma unroll
            for (int k = 0; k < N; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(u[k], u[k + 1], u[k + 2], u[k + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < N; ++k) dst[k] = u[k];
        }
    }
}

struct FilterOutPtrs {
    float* h;
    float* J;
    float* u;
    float* r;
};

template <int N>
__device__ __forceinline__ void write_filter(const FilterOutPtrs& o, long long b, float h, const float (&J)[N],
                                             const float (&u)[N], float r) {
    if (o.h) o.h[b] = h;
    if (o.r) o.r[b] = r;
    if (o.J) {
#pragma unroll
        for (int k = 0; k < N; ++k) o.J[b * N + k] = J[k];
    }
    if (o.u) {
#pragma unroll
        for (int k = 0; k < N; ++k) o.u[b * N + k] = u[k];
    }
}

template <int N, int H, int L>
__global__ void __launch_bounds__(NT) filter_kernel(NetArgs net, const float* __restrict__ datax, long long B,
                                                    FilterOutPtrs out) {
    extern __shared__ __align__(16) float smem_raw[];
    using LY = NetLayout<N + 1, H, L>;
    const Smem s = carve(smem_raw, 0, 0, LY::TOTAL, ParamLayout<N>::TOTAL);
    SceneArgs none{nullptr, nullptr, 0, 0, 0, nullptr, nullptr};
    stage<N>(s, none, &net, LY::TOTAL);
    float* scr = s.scratch + threadIdx.x;
    for (long long b = (long long)blockIdx.x * NT + threadIdx.x; b < B; b += (long long)gridDim.x * NT) {
        const float* d = datax + b * (2 * N + 1);
        float q[N], dodq[N], J[N], u[N], h, r;
#pragma unroll
        for (int k = 0; k < N; ++k) { q[k] = __ldg(d + k); dodq[k] = __ldg(d + N + 1 + k); }
        const float o = __ldg(d + N);
        float g[N];
        load_goal<N>(net, b, g);
        cbf_filter_state<N, H, L, NT>(s.weights, s.params, scr, q, o, dodq, g, net.u_ref != nullptr, net.lam,
                                      net.penalty, h, J, u, r);
        write_filter<N>(out, b, h, J, u, r);
    }
}

// Barrier networks of ANY shape (cbf_hidden_layers / cbf_hidden_size are hyper-parameters of the reference:
// neural_cbf/training/train_arm_mindis.py:176-177, neural_mindis_cbf_controller.py:41-53).  The specialised kernels are
// built for the shape every shipped entry point uses (48 x 2); everything else runs here: one state per thread, CTAs of
// 32 threads, weights read through the read-only path (every lane the same address: one broadcast per load), all layer
// activations of the thread parked in a conflict-free shared-memory column (they double as the ReLU masks of the reverse
// pass).  Same contract and packing as filter_kernel; O(L H^2) loads per state, so it is a compatibility path, not a fast
// one.  Dynamic shared memory: 32 * (L + 2) * H floats (+ parameters).
constexpr int GEN_NT = 32;
template <int N>
__global__ void __launch_bounds__(GEN_NT) filter_kernel_generic(NetArgs net, int H, int L, const float* __restrict__ datax,
                                                                long long B, FilterOutPtrs out) {
    extern __shared__ __align__(16) float smem_raw[];
    constexpr int N_IN = N + 1, IN_PAD = (N_IN + 3) & ~3;
    float* prm = smem_raw;                                     // ParamLayout<N>
    float* act = smem_raw + ParamLayout<N>::TOTAL;             // [(L + 1) * H][32] activations, then [H][32] gradient
    using PL = ParamLayout<N>;
    for (int i = threadIdx.x; i < N * N; i += GEN_NT) prm[PL::K + i] = net.K ? net.K[i] : 0.f;
    for (int i = threadIdx.x; i < N; i += GEN_NT) { prm[PL::LO + i] = net.u_lo[i]; prm[PL::HI + i] = net.u_hi[i]; }
    __syncthreads();
    const float* __restrict__ w = net.weights;
    const float* W_in = w;
    const float* b_in = w + H * IN_PAD;
    const float* hid = b_in + H;
    const long long hid_stride = (long long)H * H + H;
    const float* w_out = hid + L * hid_stride;
    const float* b_out = w_out + H;
    const int lane = threadIdx.x;
    float* grad = act + (L + 1) * H * GEN_NT;
    for (long long b = (long long)blockIdx.x * GEN_NT + lane; b < B; b += (long long)gridDim.x * GEN_NT) {
        const float* d = datax + b * (2 * N + 1);
        float x[N_IN], q[N], dodq[N];
#pragma unroll
        for (int k = 0; k < N; ++k) { q[k] = __ldg(d + k); x[k] = q[k]; dodq[k] = __ldg(d + N + 1 + k); }
        x[N] = __ldg(d + N);
        for (int j = 0; j < H; ++j) {                          // input layer
            float sacc = __ldg(b_in + j);
#pragma unroll
            for (int k = 0; k < N_IN; ++k) sacc = fmaf(__ldg(W_in + j * IN_PAD + k), x[k], sacc);
            act[j * GEN_NT + lane] = fmaxf(sacc, 0.f);
        }
        for (int l = 0; l < L; ++l) {                          // hidden layers
            const float* Wl = hid + l * hid_stride;
            const float* bl = Wl + (long long)H * H;
            const float* a_in = act + l * H * GEN_NT;
            float* a_out = act + (l + 1) * H * GEN_NT;
            for (int j = 0; j < H; ++j) {
                float s0 = __ldg(bl + j), s1 = 0.f;
                int k = 0;
                for (; k + 1 < H; k += 2) {
                    s0 = fmaf(__ldg(Wl + (long long)j * H + k), a_in[k * GEN_NT + lane], s0);
                    s1 = fmaf(__ldg(Wl + (long long)j * H + k + 1), a_in[(k + 1) * GEN_NT + lane], s1);
                }
                if (k < H) s0 = fmaf(__ldg(Wl + (long long)j * H + k), a_in[k * GEN_NT + lane], s0);
                a_out[j * GEN_NT + lane] = fmaxf(s0 + s1, 0.f);
            }
        }
        const float* a_last = act + L * H * GEN_NT;
        float h = __ldg(b_out);
        for (int k = 0; k < H; ++k) {
            const float a = a_last[k * GEN_NT + lane], wo = __ldg(w_out + k);
            h = fmaf(wo, a, h);
            grad[k * GEN_NT + lane] = a > 0.f ? wo : 0.f;      // seed of the reverse pass
        }
        for (int l = L - 1; l >= 0; --l) {                     // reverse mode through the hidden layers, in place
            const float* Wl = hid + l * hid_stride;
            float* a_prev = act + l * H * GEN_NT;              // relu outputs of layer l: mask, then overwritten by g_prev
            for (int k = 0; k < H; ++k) {
                float acc = 0.f;
                for (int j = 0; j < H; ++j) acc = fmaf(__ldg(Wl + (long long)j * H + k), grad[j * GEN_NT + lane], acc);
                a_prev[k * GEN_NT + lane] = a_prev[k * GEN_NT + lane] > 0.f ? acc : 0.f;
            }
            for (int k = 0; k < H; ++k) grad[k * GEN_NT + lane] = a_prev[k * GEN_NT + lane];
        }
        float jac[N_IN];
#pragma unroll
        for (int k = 0; k < N_IN; ++k) jac[k] = 0.f;
        for (int j = 0; j < H; ++j) {
            const float gj = grad[j * GEN_NT + lane];
#pragma unroll
            for (int k = 0; k < N_IN; ++k) jac[k] = fmaf(__ldg(W_in + j * IN_PAD + k), gj, jac[k]);
        }
        float g[N], J[N], u[N], r;
        load_goal<N>(net, b, g);
        cbf_filter_post<N>(prm, q, dodq, g, net.u_ref != nullptr, net.lam, net.penalty, h, jac, J, u, r);
        write_filter<N>(out, b, h, J, u, r);
    }
}

// stand-alone QP (CLFController.solve_CLF_QP with caller-computed Lie derivatives)
template <int N>
__global__ void __launch_bounds__(256) qp_kernel(const float* __restrict__ a, const float* __restrict__ c,
                                                 const float* __restrict__ u_ref, long long B,
                                                 const float* __restrict__ u_lo, const float* __restrict__ u_hi,
                                                 float penalty, float* __restrict__ u, float* __restrict__ r) {
    __shared__ float lim[2 * N];
    if (threadIdx.x < N) { lim[threadIdx.x] = u_lo[threadIdx.x]; lim[N + threadIdx.x] = u_hi[threadIdx.x]; }
    __syncthreads();
    for (long long b = (long long)blockIdx.x * 256 + threadIdx.x; b < B; b += (long long)gridDim.x * 256) {
        float aa[N], ur[N], uu[N], rr;
#pragma unroll
        for (int k = 0; k < N; ++k) { aa[k] = __ldg(a + b * N + k); ur[k] = __ldg(u_ref + b * N + k); }
        qp_solve<N>(aa, ur, lim, lim + N, __ldg(c + b), fminf(penalty, 1e6f), uu, rr);
#pragma unroll
        for (int k = 0; k < N; ++k) u[b * N + k] = uu[k];
        if (r) r[b] = rr;
    }
}

// reverse mode of the stand-alone QP: (a, c, u_ref, dL/du, dL/dr) -> (dL/da, dL/dc, dL/du_ref).  fp32 in / out, the
// sensitivities are evaluated in double (qp_backward).  HBM-bound: (3n+2) floats in, (2n+1) floats out per row.
template <int N>
__global__ void __launch_bounds__(256) qp_backward_kernel(const float* __restrict__ a, const float* __restrict__ c,
                                                          const float* __restrict__ u_ref, long long B,
                                                          const float* __restrict__ u_lo, const float* __restrict__ u_hi,
                                                          float penalty, const float* __restrict__ grad_u,
                                                          const float* __restrict__ grad_r, float* __restrict__ grad_a,
                                                          float* __restrict__ grad_c, float* __restrict__ grad_u_ref) {
    __shared__ double lim[2 * N];
    if (threadIdx.x < N) { lim[threadIdx.x] = u_lo[threadIdx.x]; lim[N + threadIdx.x] = u_hi[threadIdx.x]; }
    __syncthreads();
    for (long long b = (long long)blockIdx.x * 256 + threadIdx.x; b < B; b += (long long)gridDim.x * 256) {
        double aa[N], ur[N], gu[N], ga[N], gt[N], gc;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            aa[k] = __ldg(a + b * N + k); ur[k] = __ldg(u_ref + b * N + k); gu[k] = __ldg(grad_u + b * N + k);
        }
        qp_backward<N, double>(aa, ur, lim, lim + N, (double)__ldg(c + b), (double)fminf(penalty, 1e6f), gu,
                               grad_r ? (double)__ldg(grad_r + b) : 0.0, ga, gc, gt);
#pragma unroll
        for (int k = 0; k < N; ++k) {
            if (grad_a) grad_a[b * N + k] = (float)ga[k];
            if (grad_u_ref) grad_u_ref[b * N + k] = (float)gt[k];
        }
        if (grad_c) grad_c[b] = (float)gc;
    }
}

// multi-scenario QP: a [B,S,n], c [B,S], u_ref [B,n] -> u [B,n], r [B,S], iters [B] (optional).  fp32 in / out, the
// dual iteration runs in double (see qp_solve_multi).
constexpr int QP_MAX_CON = 4;
template <int N>
__global__ void __launch_bounds__(128) qp_multi_kernel(const float* __restrict__ a, const float* __restrict__ c,
                                                       const float* __restrict__ u_ref, long long B, int S,
                                                       const float* __restrict__ u_lo, const float* __restrict__ u_hi,
                                                       float penalty, int max_iter, float tol, float* __restrict__ u,
                                                       float* __restrict__ r, int* __restrict__ iters) {
    __shared__ double lim[2 * N];
    if (threadIdx.x < N) { lim[threadIdx.x] = u_lo[threadIdx.x]; lim[N + threadIdx.x] = u_hi[threadIdx.x]; }
    __syncthreads();
    for (long long b = (long long)blockIdx.x * 128 + threadIdx.x; b < B; b += (long long)gridDim.x * 128) {
        double aa[QP_MAX_CON][N], cc[QP_MAX_CON], ur[N], uu[N], rr[QP_MAX_CON];
#pragma unroll
        for (int i = 0; i < QP_MAX_CON; ++i) {
            cc[i] = i < S ? (double)__ldg(c + b * S + i) : 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) aa[i][k] = i < S ? (double)__ldg(a + (b * S + i) * N + k) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < N; ++k) ur[k] = (double)__ldg(u_ref + b * N + k);
        const int it = qp_solve_multi<N, QP_MAX_CON, double>(aa, ur, lim, lim + N, cc, S, (double)fminf(penalty, 1e6f),
                                                             max_iter, (double)tol, uu, rr);
#pragma unroll
        for (int k = 0; k < N; ++k) u[b * N + k] = (float)uu[k];
#pragma unroll
        for (int i = 0; i < QP_MAX_CON; ++i)
            if (r && i < S) r[b * S + i] = (float)rr[i];
        if (iters) iters[b] = it;
    }
}

// reverse mode of the multi-scenario QP: re-solves the row (stateless ABI: no multipliers are carried over from the
// forward call) and applies qp_multi_adjoint.  grad_a [B,S,n], grad_c [B,S], grad_u_ref [B,n]; all double inside.
template <int N>
__global__ void __launch_bounds__(128) qp_multi_backward_kernel(const float* __restrict__ a, const float* __restrict__ c,
                                                                const float* __restrict__ u_ref, long long B, int S,
                                                                const float* __restrict__ u_lo, const float* __restrict__ u_hi,
                                                                float penalty, int max_iter, float tol,
                                                                const float* __restrict__ grad_u, const float* __restrict__ grad_r,
                                                                float* __restrict__ grad_a, float* __restrict__ grad_c,
                                                                float* __restrict__ grad_u_ref) {
    __shared__ double lim[2 * N];
    if (threadIdx.x < N) { lim[threadIdx.x] = u_lo[threadIdx.x]; lim[N + threadIdx.x] = u_hi[threadIdx.x]; }
    __syncthreads();
    for (long long b = (long long)blockIdx.x * 128 + threadIdx.x; b < B; b += (long long)gridDim.x * 128) {
        double aa[QP_MAX_CON][N], cc[QP_MAX_CON], ur[N], uu[N], rr[QP_MAX_CON], gu[N], gr[QP_MAX_CON];
#pragma unroll
        for (int i = 0; i < QP_MAX_CON; ++i) {
            cc[i] = i < S ? (double)__ldg(c + b * S + i) : 0.0;
            gr[i] = (i < S && grad_r) ? (double)__ldg(grad_r + b * S + i) : 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) aa[i][k] = i < S ? (double)__ldg(a + (b * S + i) * N + k) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < N; ++k) { ur[k] = (double)__ldg(u_ref + b * N + k); gu[k] = (double)__ldg(grad_u + b * N + k); }
        const double p = (double)fminf(penalty, 1e6f);
        qp_solve_multi<N, QP_MAX_CON, double>(aa, ur, lim, lim + N, cc, S, p, max_iter, (double)tol, uu, rr);
        double ga[QP_MAX_CON][N], gc[QP_MAX_CON], gt[N];
        qp_multi_adjoint<N, QP_MAX_CON, double>(aa, ur, lim, lim + N, cc, S, p, uu, gu, gr, ga, gc, gt);
#pragma unroll
        for (int k = 0; k < N; ++k)
            if (grad_u_ref) grad_u_ref[b * N + k] = (float)gt[k];
#pragma unroll
        for (int i = 0; i < QP_MAX_CON; ++i) {
            if (i < S) {
                if (grad_c) grad_c[b * S + i] = (float)gc[i];
#pragma unroll
                for (int k = 0; k < N; ++k)
                    if (grad_a) grad_a[(b * S + i) * N + k] = (float)ga[i][k];
            }
        }
    }
}

template <class RB, int H, int L, bool GRID>
__global__ void __launch_bounds__(NT) safety_kernel(SceneArgs sc, NetArgs net, const float* __restrict__ q,
                                                    long long q_stride, long long B, float thr_soft, float thr_hard,
                                                    CollideOutPtrs cout, FilterOutPtrs fout, PeerOut po) {
    constexpr int N = RB::N;
    extern __shared__ __align__(16) float smem_raw[];
    using LY = NetLayout<N + 1, H, L>;
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, LY::TOTAL, ParamLayout<N>::TOTAL);
    stage<N>(s, sc, &net, LY::TOTAL);
    const SceneS sv = scene_view(s, sc, thr_soft, thr_hard);
    const float base_min = base_min_cta<RB>(s, sv);
    float* scr = s.scratch + threadIdx.x;
    for (long long b = (long long)blockIdx.x * NT + threadIdx.x; b < B; b += (long long)gridDim.x * NT) {
        float qq[N];
#pragma unroll
        for (int k = 0; k < N; ++k) qq[k] = __ldg(q + b * q_stride + k);
        CollideResult<RB> cr;
        collide_state<RB, true, false, GRID>(qq, sv, base_min, cr);
        write_collide<RB>(cout, b, qq, cr, thr_soft, thr_hard);
        float g[N], J[N], u[N], h, r;
        load_goal<N>(net, b, g);
        cbf_filter_state<N, H, L, NT>(s.weights, s.params, scr, qq, cr.o, cr.dodq, g, net.u_ref != nullptr, net.lam,
                                      net.penalty, h, J, u, r);
        write_filter<N>(fout, b, h, J, u, r);
        if (po.n) write_peers<N>(po, b, cr.self_ok && (cr.soft_min > thr_soft) && !(cr.hard_min <= thr_hard), u);
    }
}


// closed-loop rollouts (the steer loop): one thread per trajectory, the whole loop in one launch.
struct RolloutArgs {
    int steps, ctrl_every;
    float dt;
    int stuck_lag, stuck_after;
    float stuck_eps;
    int keep_going;
    int stuck_mode;   // 0: single steer (tsa.py:263), 1: batched steer (batch_tsa.py:225-229), see include/cbfrrt.h
};

// What happens to a trajectory after the Euler step t produced qn (unsafe = hard collision test of qn).  Shared by the
// FP32-FMA and the tcgen05 rollout kernels.  Returns true when the trajectory is finished.
//   stuck_mode 0 (tsa.py:251-266): an unsafe qn is not appended (ok = 0, stop unless keep_going); otherwise q <- qn is
//     appended and the stuck test compares the states appended at steps t and t - stuck_lag when t > stuck_after.
//   stuck_mode 1 (batch_tsa.py:225-229): the stuck test runs BEFORE the append, on the list as it stands (more than
//     stuck_after states stored: last vs the one stuck_lag before it); stuck OR unsafe clears ok and the row stops
//     appending for good, but x_current keeps integrating (q <- qn always; q_final is that running state).  Without
//     keep_going the row stops at its death instead.
template <int N>
__device__ __forceinline__ bool rollout_after_step(const RolloutArgs& ra, int t, bool unsafe, float (&q)[N],
                                                   const float (&qn)[N], float* __restrict__ tr, bool& ok, bool& dead,
                                                   int& done) {
    if (ra.stuck_mode == 1) {
        if (!dead) {
            bool stuck = false;
            if (ra.stuck_lag > 0 && done > ra.stuck_after) {
                const float* a = tr + (done - 1) * N;
                const float* b = tr + (done - 1 - ra.stuck_lag) * N;
                float dsq = 0.f;
#pragma unroll
                for (int k = 0; k < N; ++k) { const float d = __fsub_rn(a[k], b[k]); dsq = fmaf(d, d, dsq); }
                stuck = __fsqrt_rn(dsq) < ra.stuck_eps;
            }
            if (stuck || unsafe) { ok = false; dead = true; }
        }
#pragma unroll
        for (int k = 0; k < N; ++k) q[k] = qn[k];
        if (!dead) {
            if (tr) {
#pragma unroll
                for (int k = 0; k < N; ++k) tr[done * N + k] = q[k];
            }
            ++done;
        }
        return dead && !ra.keep_going;
    }
    if (unsafe) {  // tsa.py:251-253: colliding state is not appended
        ok = false;
        if (!ra.keep_going) return true;
    }
#pragma unroll
    for (int k = 0; k < N; ++k) q[k] = qn[k];
    if (tr) {
#pragma unroll
        for (int k = 0; k < N; ++k) tr[t * N + k] = q[k];
    }
    done = t + 1;
    if (ra.stuck_lag > 0 && t > ra.stuck_after) {  // tsa.py:263
        const float* old = tr + (t - ra.stuck_lag) * N;
        float dsq = 0.f;
#pragma unroll
        for (int k = 0; k < N; ++k) { const float d = __fsub_rn(q[k], old[k]); dsq = fmaf(d, d, dsq); }
        if (__fsqrt_rn(dsq) < ra.stuck_eps) return true;
    }
    return false;
}

// The same step logic with the stored states mirrored in a RING of the last STUCK_RING rows (shared memory of the warp
// kernels): the stuck test re-reads states written a few steps ago, and from the trajectory array that is an L2 round trip
// (~0.35 us) on the critical path of every step of a one-warp trajectory.  tr (global, optional) is only written.
// Requires ra.stuck_lag < STUCK_RING (the launcher checks).  Ring slot of stored state number d: d % STUCK_RING.
constexpr int STUCK_RING = 16;
template <int N>
__device__ __forceinline__ bool rollout_after_step_ring(const RolloutArgs& ra, int t, bool unsafe, float (&q)[N],
                                                        const float (&qn)[N], float* __restrict__ tr, float* __restrict__ ring,
                                                        bool& ok, bool& dead, int& done) {
    if (ra.stuck_mode == 1) {
        if (!dead) {
            bool stuck = false;
            if (ra.stuck_lag > 0 && done > ra.stuck_after) {
                const float* a = ring + ((done - 1) % STUCK_RING) * N;
                const float* b = ring + ((done - 1 - ra.stuck_lag) % STUCK_RING) * N;
                float dsq = 0.f;
#pragma unroll
                for (int k = 0; k < N; ++k) { const float d = __fsub_rn(a[k], b[k]); dsq = fmaf(d, d, dsq); }
                stuck = __fsqrt_rn(dsq) < ra.stuck_eps;
            }
            if (stuck || unsafe) { ok = false; dead = true; }
        }
#pragma unroll
        for (int k = 0; k < N; ++k) q[k] = qn[k];
        if (!dead) {
#pragma unroll
            for (int k = 0; k < N; ++k) ring[(done % STUCK_RING) * N + k] = q[k];
            if (tr) {
#pragma unroll
                for (int k = 0; k < N; ++k) tr[done * N + k] = q[k];
            }
            ++done;
        }
        return dead && !ra.keep_going;
    }
    if (unsafe) {
        ok = false;
        if (!ra.keep_going) return true;
    }
#pragma unroll
    for (int k = 0; k < N; ++k) q[k] = qn[k];
#pragma unroll
    for (int k = 0; k < N; ++k) ring[(t % STUCK_RING) * N + k] = q[k];
    if (tr) {
#pragma unroll
        for (int k = 0; k < N; ++k) tr[t * N + k] = q[k];
    }
    done = t + 1;
    if (ra.stuck_lag > 0 && t > ra.stuck_after) {
        const float* old = ring + ((t - ra.stuck_lag) % STUCK_RING) * N;
        float dsq = 0.f;
#pragma unroll
        for (int k = 0; k < N; ++k) { const float d = __fsub_rn(q[k], old[k]); dsq = fmaf(d, d, dsq); }
        if (__fsqrt_rn(dsq) < ra.stuck_eps) return true;
    }
    return false;
}

// ------------------------------------------------------------------ tensor-core (tcgen05) variants
// Same contracts as filter_kernel / safety_kernel; the MLP of every 128-state tile runs as TF32x3 GEMMs on the
// tensor cores (mlp_tc.cuh).  One CTA per SM of GG tile groups (see tc_groups) sharing the staged weights; all 128
// threads of a group take part in every tile of that group (idle rows are masked), because the MMAs, the TMEM
// traffic and the barriers are group-wide.
// tile groups per CTA: 4 (512 threads, 128 registers, 16 warps per SM).  (Round 1 ran the broad-phase-grid variants with
// 3 groups to leave L1 to the per-sphere candidate gathers; with the distance-bound field the geometry stage is
// latency bound and the fourth group wins: cfg4 valid+u 5.96 -> 5.33 ms per 2^23 states, profiles/r2c_*.)
constexpr int tc_groups(bool) { return 4; }

template <int N, int L, int GG>
__device__ __forceinline__ void filter_tile_tc(tc::Ctx<L, GG>& c, const float* prm, const float (&q)[N], float o,
                                               const float (&dodq)[N], const float (&goal)[N], bool goal_is_uref,
                                               float lam, float penalty, float& h, float (&J)[N], float (&u)[N], float& r) {
    float x[N + 1], jac[N + 1];
#pragma unroll
    for (int k = 0; k < N; ++k) x[k] = q[k];
    x[N] = o;
    tc::mlp_h_jac<N + 1, L, GG>(c, x, h, jac);
    cbf_filter_post<N>(prm, q, dodq, goal, goal_is_uref, lam, penalty, h, jac, J, u, r);
}

// ---- coalesced tile I/O of the tensor-core safety kernel.  A tile is 128 CONSECUTIVE rows, so every per-row array of
// the tile is one contiguous, 16-byte aligned block in global memory (128 x 28 B of q, 128 x 60 B of datax, ...).  Rows
// are exchanged through the group's A_lo tile, which is idle between the last GEMM of a tile and the first hidden-layer
// store of the next: thread t parks its row (odd row length: conflict-free), one group barrier, then the 128 threads
// stream the block as float4 -- full 128-byte lines instead of 28-byte-strided scalar accesses (round 1: 79 % of the L2
// sectors "excessive"), and 16-byte multimem.st to the NVSwitch multicast address in the fused exchange.
// Layout of the group's A_lo tile in floats: [0,1024) first-layer A_lo | [1024,1920) q rows | [2048,6144) output rows.
template <int W>
__device__ __forceinline__ void park_row(float* __restrict__ stg, int t, const float (&v)[W]) {
    if constexpr (W % 4 == 0) {
#pragma unroll
        for (int k = 0; k < W; k += 4) *reinterpret_cast<float4*>(stg + t * W + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < W; ++k) stg[t * W + k] = v[k];
    }
}
// stream 128 rows of W floats (or n16 16-byte words) from the staging block to dst; MC: multimem.st to a multicast address
template <bool MC>
__device__ __forceinline__ void flush_block(const float* __restrict__ stg, int t, float* __restrict__ dst, int n16) {
    const float4* s4 = reinterpret_cast<const float4*>(stg);
    for (int i = t; i < n16; i += tc::M_TILE) {
        const float4 v = s4[i];
        if constexpr (MC) multimem_st_v4(dst + 4 * i, v.x, v.y, v.z, v.w);
        else *reinterpret_cast<float4*>(dst + 4 * i) = v;
    }
}
// Per-WARP coalesced row I/O: the 32 rows of a warp are 32 x W consecutive floats in global memory, i.e. W full 128-byte
// lines.  Rows go through a warp-private staging block (conflict-free for odd W) with only a __syncwarp: instruction k
// then moves floats [32 k, 32 k + 32) of the block -- one full line per instruction -- instead of 32 scattered 4-byte
// pieces at a 4 W-byte stride (ncu, round 2: 2.9x the algorithmic DRAM write traffic with the strided stores).
template <int W>
__device__ __forceinline__ void warp_store_rows(float* __restrict__ stg, int lane, const float (&v)[W], float* __restrict__ dst) {
    __syncwarp();
#pragma unroll
    for (int k = 0; k < W; ++k) stg[lane * W + k] = v[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < W; ++k) dst[32 * k + lane] = stg[32 * k + lane];
}
template <int W>
__device__ __forceinline__ void warp_load_rows(float* __restrict__ stg, int lane, const float* __restrict__ src, float (&v)[W]) {
    __syncwarp();
#pragma unroll
    for (int k = 0; k < W; ++k) stg[32 * k + lane] = __ldg(src + 32 * k + lane);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < W; ++k) v[k] = stg[lane * W + k];
}
constexpr int VEC_Q = 1, VEC_OUT = 2, VEC_PEER = 4;   // which pointer sets are 16-byte aligned (checked on the host)

// datax -> (h, J, u, r) on the tensor cores.  PEERS: the barrier-network pass of the two-pass sweep (launch_safety) in the
// multi-GPU exchange: the row's control and the mask byte the geometry pass left in safe_in are stored into every rank's
// replica (fused all-gather, see safety_kernel_tc) -- whole 128-row tiles staged through the group's A_lo tile and
// streamed as 16-byte (multimem) stores, ragged tiles row by row.
template <int N, int L, bool PEERS>
__global__ void __launch_bounds__(4 * tc::M_TILE, 1) filter_kernel_tc(NetArgs net, const float* __restrict__ datax,
                                                                      long long B, FilterOutPtrs out, PeerOut po,
                                                                      const uint8_t* __restrict__ safe_in, int vec) {
    constexpr int GG = 4, NTC = GG * tc::M_TILE, MT = tc::M_TILE;
    constexpr int O_U = 2048, O_B = O_U + MT * N;          // staging offsets (floats) inside the group's A_lo tile
    static_assert(O_B + MT / 4 <= tc::ALO_TILE, "staging does not fit the A_lo tile");
    extern __shared__ __align__(16) float smem_raw[];
    const Smem s = carve(smem_raw, 0, 0, tc::Layout<L, GG>::TOTAL, ParamLayout<N>::TOTAL);
    SceneArgs none{nullptr, nullptr, 0, 0, 0, nullptr, nullptr};
    stage<N, NTC>(s, none, &net, 0);
    tc::Ctx<L, GG> c;
    tc::setup<N + 1, L, GG>(c, s.weights, net.weights);
    float* const alo = reinterpret_cast<float*>(c.alo());
    const long long n_tiles = (B + tc::M_TILE - 1) / tc::M_TILE;
    for (long long tile = (long long)blockIdx.x * GG + c.g; tile < n_tiles; tile += (long long)gridDim.x * GG) {
        const long long b0 = tile * tc::M_TILE, b = b0 + c.t;
        const bool active = b < B;
        float q[N], dodq[N], g[N], J[N], u[N], h, r, o = 0.f;
#pragma unroll
        for (int k = 0; k < N; ++k) { q[k] = 0.f; dodq[k] = 0.f; g[k] = 0.f; }
        if (active) {
            const float* d = datax + b * (2 * N + 1);
#pragma unroll
            for (int k = 0; k < N; ++k) { q[k] = __ldg(d + k); dodq[k] = __ldg(d + N + 1 + k); }
            o = __ldg(d + N);
            load_goal<N>(net, b, g);
        }
        filter_tile_tc<N, L, GG>(c, s.params, q, o, dodq, g, net.u_ref != nullptr, net.lam, net.penalty, h, J, u, r);
        if (active) write_filter<N>(out, b, h, J, u, r);
        if constexpr (PEERS) {
            const bool safe = active && safe_in[b] != 0;
            if (b0 + MT <= B && (vec & VEC_PEER)) {        // uniform over the group
                park_row<N>(alo + O_U, c.t, u);
                reinterpret_cast<uint8_t*>(alo + O_B)[c.t] = safe;
                tc::group_sync(c.g);
                const long long g0 = po.row_offset + b0;
                if (po.multicast) {
                    flush_block<true>(alo + O_U, c.t, po.u[0] + g0 * N, MT * N / 4);
                    flush_block<true>(alo + O_B, c.t, reinterpret_cast<float*>(po.valid[0] + g0), MT / 16);
                } else {
                    for (int p = 0; p < po.n; ++p) {
                        flush_block<false>(alo + O_U, c.t, po.u[p] + g0 * N, MT * N / 4);
                        flush_block<false>(alo + O_B, c.t, reinterpret_cast<float*>(po.valid[p] + g0), MT / 16);
                    }
                }
                tc::group_sync(c.g);                       // the staging block is the next tile's A_lo
            } else if (active) {
                write_peers<N>(po, b, safe, u);
            }
        }
    }
    tc::teardown<L, GG>(c);
}

template <class RB, int L, bool GRID, int GG>
__global__ void __launch_bounds__(GG * tc::M_TILE, 1) safety_kernel_tc(SceneArgs sc, NetArgs net, const float* __restrict__ q,
                                                           long long q_stride, long long B, float thr_soft, float thr_hard,
                                                           CollideOutPtrs cout, FilterOutPtrs fout, PeerOut po, int vec) {
    constexpr int N = RB::N, DX = 2 * N + 1, MT = tc::M_TILE;
    constexpr int NTC = GG * tc::M_TILE;
    // output staging offsets (floats) inside [2048, 6144) of the group's A_lo tile
    constexpr int O_DX = 2048, O_U = O_DX + MT * DX, O_J = O_U + MT * N, O_H = O_J + MT * N, O_R = O_H + MT, O_B = O_R + MT;
    static_assert(O_B + 2 * MT / 4 <= tc::ALO_TILE && 1024 + MT * N <= 2048, "staging does not fit the A_lo tile");
    extern __shared__ __align__(16) float smem_raw[];
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, tc::Layout<L, GG>::TOTAL, ParamLayout<N>::TOTAL);
    stage<N, NTC>(s, sc, &net, 0);
    const SceneS sv = scene_view(s, sc, thr_soft, thr_hard);
    const float base_min = base_min_cta<RB, NTC>(s, sv);
    tc::Ctx<L, GG> c;
    tc::setup<N + 1, L, GG>(c, s.weights, net.weights);
    float* const alo = reinterpret_cast<float*>(c.alo());
    const bool want_valid_stage = cout.safe != nullptr || po.n > 0;
    const long long n_tiles = (B + tc::M_TILE - 1) / tc::M_TILE;
    for (long long tile = (long long)blockIdx.x * GG + c.g; tile < n_tiles; tile += (long long)gridDim.x * GG) {
        const long long b0 = tile * tc::M_TILE, b = b0 + c.t;
        const bool active = b < B, full = b0 + tc::M_TILE <= B;      // full: uniform over the group
        const long long bw = b0 + (c.t & ~31);                        // first row of this warp
        const bool warp_full = bw + 32 <= B;                          // uniform over the warp
        float* const wstg = alo + 2048 + ((c.t >> 5) & 3) * 512;      // warp-private staging inside the output region
        float qq[N], g[N], J[N], u[N], h, r;
#pragma unroll
        for (int k = 0; k < N; ++k) { qq[k] = 0.f; g[k] = 0.f; }
        CollideResult<RB> cr;
        cr.o = 0.f; cr.soft_min = 0.f; cr.hard_min = 0.f; cr.self_ok = false;
#pragma unroll
        for (int k = 0; k < N; ++k) cr.dodq[k] = 0.f;
        if (full && (vec & VEC_Q)) {                                   // coalesced load of the tile's 128 x N floats
            const float4* src = reinterpret_cast<const float4*>(q + b0 * N);
            float4* dst = reinterpret_cast<float4*>(alo + 1024);
            for (int i = c.t; i < MT * N / 4; i += MT) dst[i] = __ldg(src + i);
            tc::group_sync(c.g);
#pragma unroll
            for (int k = 0; k < N; ++k) qq[k] = alo[1024 + c.t * N + k];
        } else if (warp_full && q_stride == N) {                       // the warp's 32 rows: N full lines
            warp_load_rows<N>(wstg, c.t & 31, q + bw * N, qq);
        } else if (active) {
#pragma unroll
            for (int k = 0; k < N; ++k) qq[k] = __ldg(q + b * q_stride + k);
        }
        if (active) {
            collide_state<RB, true, false, GRID>(qq, sv, base_min, cr);
            load_goal<N>(net, b, g);
        }
        filter_tile_tc<N, L, GG>(c, s.params, qq, cr.o, cr.dodq, g, net.u_ref != nullptr, net.lam, net.penalty, h, J, u, r);
        const bool unsafe = cr.hard_min <= thr_hard;
        const bool safe = cr.self_ok && (cr.soft_min > thr_soft) && !unsafe;
        const bool out_fast = full && (vec & VEC_OUT), peer_fast = full && po.n > 0 && (vec & VEC_PEER);
        if (out_fast || peer_fast) {
            if (cout.datax && out_fast) {
                float row[DX];
#pragma unroll
                for (int k = 0; k < N; ++k) { row[k] = qq[k]; row[N + 1 + k] = cr.dodq[k]; }
                row[N] = cr.o;
                park_row<DX>(alo + O_DX, c.t, row);
            }
            if ((fout.u && out_fast) || peer_fast) park_row<N>(alo + O_U, c.t, u);
            if (fout.J && out_fast) park_row<N>(alo + O_J, c.t, J);
            if (fout.h && out_fast) alo[O_H + c.t] = h;
            if (fout.r && out_fast) alo[O_R + c.t] = r;
            if (want_valid_stage) reinterpret_cast<uint8_t*>(alo + O_B)[c.t] = safe;
            if (cout.unsafe && out_fast) reinterpret_cast<uint8_t*>(alo + O_B)[MT + c.t] = unsafe;
            tc::group_sync(c.g);
            if (out_fast) {
                if (cout.datax) flush_block<false>(alo + O_DX, c.t, cout.datax + b0 * DX, MT * DX / 4);
                if (fout.u) flush_block<false>(alo + O_U, c.t, fout.u + b0 * N, MT * N / 4);
                if (fout.J) flush_block<false>(alo + O_J, c.t, fout.J + b0 * N, MT * N / 4);
                if (fout.h) flush_block<false>(alo + O_H, c.t, fout.h + b0, MT / 4);
                if (fout.r) flush_block<false>(alo + O_R, c.t, fout.r + b0, MT / 4);
                if (cout.safe) flush_block<false>(alo + O_B, c.t, reinterpret_cast<float*>(cout.safe + b0), MT / 16);
                if (cout.unsafe) flush_block<false>(alo + O_B + MT / 4, c.t, reinterpret_cast<float*>(cout.unsafe + b0), MT / 16);
            }
            if (peer_fast) {
                const long long g0 = po.row_offset + b0;
                if (po.multicast) {
                    flush_block<true>(alo + O_U, c.t, po.u[0] + g0 * N, MT * N / 4);
                    flush_block<true>(alo + O_B, c.t, reinterpret_cast<float*>(po.valid[0] + g0), MT / 16);
                } else {
                    for (int p = 0; p < po.n; ++p) {
                        flush_block<false>(alo + O_U, c.t, po.u[p] + g0 * N, MT * N / 4);
                        flush_block<false>(alo + O_B, c.t, reinterpret_cast<float*>(po.valid[p] + g0), MT / 16);
                    }
                }
            }
        }
        if (!out_fast && warp_full) {   // whole warp: rows through the warp's staging block, full-line stores
            const int lane = c.t & 31;
            if (cout.safe) cout.safe[b] = safe;
            if (cout.unsafe) cout.unsafe[b] = unsafe;
            if (cout.datax) {
                float row[DX];
#pragma unroll
                for (int k = 0; k < N; ++k) { row[k] = qq[k]; row[N + 1 + k] = cr.dodq[k]; }
                row[N] = cr.o;
                warp_store_rows<DX>(wstg, lane, row, cout.datax + bw * DX);
            }
            if (cout.arg_link) cout.arg_link[b] = cr.arg_sphere >= 0 ? RB::sph_link(cr.arg_sphere) : -2;
            if (cout.arg_obs) cout.arg_obs[b] = cr.arg_obs;
            if (fout.h) fout.h[b] = h;
            if (fout.r) fout.r[b] = r;
            if (fout.J) warp_store_rows<N>(wstg, lane, J, fout.J + bw * N);
            if (fout.u) warp_store_rows<N>(wstg, lane, u, fout.u + bw * N);
        } else if (active && !out_fast) {   // tail warp: per-row stores
            write_collide<RB>(cout, b, qq, cr, thr_soft, thr_hard);
            write_filter<N>(fout, b, h, J, u, r);
        }
        if (active && po.n && !peer_fast) write_peers<N>(po, b, safe, u);
    }
    tc::teardown<L, GG>(c);
}

// (The rollout kernels never read the self-collision predicate -- the steer loop tests unsafe_mask only, tsa.py:251 --
// so every collide_state call in them is instantiated without it: WANT_SELF = false.)
// steer loop on the tensor-core path: the time loop is uniform over a tile group (a finished trajectory idles,
// it does not leave), because the MLP rounds are group-wide; the group exits early once all its rows are finished.
template <class RB, int L, bool GRID>
__global__ void __launch_bounds__(tc_groups(GRID) * tc::M_TILE, 1) rollout_kernel_tc(SceneArgs sc, NetArgs net, const float* __restrict__ q0,
                                                            long long T, RolloutArgs ra, float thr_soft, float thr_hard,
                                                            float* __restrict__ q_final, uint8_t* __restrict__ alive,
                                                            int* __restrict__ steps_done, float* __restrict__ traj) {
    constexpr int N = RB::N;
    constexpr int GG = tc_groups(GRID), NTC = GG * tc::M_TILE;
    extern __shared__ __align__(16) float smem_raw[];
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, tc::Layout<L, GG>::TOTAL, ParamLayout<N>::TOTAL);
    stage<N, NTC>(s, sc, &net, 0);
    const SceneS sv = scene_view(s, sc, thr_soft, thr_hard);
    const float base_min = base_min_cta<RB, NTC>(s, sv);
    tc::Ctx<L, GG> c;
    tc::setup<N + 1, L, GG>(c, s.weights, net.weights);
    const long long n_tiles = (T + tc::M_TILE - 1) / tc::M_TILE;
    for (long long tile = (long long)blockIdx.x * GG + c.g; tile < n_tiles; tile += (long long)gridDim.x * GG) {
        const long long b = tile * tc::M_TILE + c.t;
        const bool active = b < T;
        float q[N], qn[N], g[N], u[N], dodq[N], o = 0.f;
#pragma unroll
        for (int k = 0; k < N; ++k) { q[k] = 0.f; g[k] = 0.f; u[k] = 0.f; dodq[k] = 0.f; }
        CollideResult<RB> cr;
        if (active) {
#pragma unroll
            for (int k = 0; k < N; ++k) q[k] = __ldg(q0 + b * N + k);
            load_goal<N>(net, b, g);
            collide_state<RB, true, false, GRID, false>(q, sv, base_min, cr);
            o = cr.o;
#pragma unroll
            for (int k = 0; k < N; ++k) dodq[k] = cr.dodq[k];
        }
        bool ok = true, finished = !active, dead = false;
        int done = 0;
        float* tr = (traj && active) ? traj + b * (long long)ra.steps * N : nullptr;
        for (int t = 0; t < ra.steps; ++t) {
            if (t % ra.ctrl_every == 0) {
                if (!tc::group_any(c.g, !finished)) break;
                float J[N], h, r, un[N];
                filter_tile_tc<N, L, GG>(c, s.params, q, o, dodq, g, false, net.lam, net.penalty, h, J, un, r);
                if (!finished) {
#pragma unroll
                    for (int k = 0; k < N; ++k) u[k] = un[k];
                }
            }
            if (finished) continue;
#pragma unroll
            for (int k = 0; k < N; ++k) qn[k] = fmaf(u[k], ra.dt, q[k]);  // arm_dynamics.py:508 (step = 1)
            if ((t + 1) % ra.ctrl_every == 0) {       // observation refresh: minimum + gradient needed
                collide_state<RB, true, false, GRID, false>(qn, sv, base_min, cr);
            } else {                                  // only the hard collision test is read: mask-only, no self test
                collide_state<RB, false, false, GRID, false>(qn, sv, base_min, cr);
            }
            if ((t + 1) % ra.ctrl_every == 0) {
                o = cr.o;
#pragma unroll
                for (int k = 0; k < N; ++k) dodq[k] = cr.dodq[k];
            }
            if (rollout_after_step<N>(ra, t, cr.hard_min <= thr_hard, q, qn, tr, ok, dead, done)) finished = true;
        }
        if (active) {
#pragma unroll
            for (int k = 0; k < N; ++k) q_final[b * N + k] = q[k];
            alive[b] = ok;
            steps_done[b] = done;
        }
    }
    tc::teardown<L, GG>(c);
}

// edge validity (tsa.py:272-286): one thread per (edge, point) over the concatenation of all edges' points -- every
// lane busy whatever the edge lengths (a CTA-owns-whole-edges layout idles 23 % of the lanes at 33 points per edge and
// measured 14 % slower on cfg2).  Edges have a common K (off == nullptr: e = p / (K+1)) or their own K_e = int(|q_to -
// q_from| / eps) (tsa.py:277-279: off [E+1] = exclusive prefix sum of K_e + 1).  Consecutive threads are consecutive
// points of one edge, so a warp's states are geometrically coherent.  first_bad must be preset to EDGE_SENTINEL
// (edge_var_init), the reduction is a global atomicMin, edge_var_finish writes valid / maps the sentinel.
constexpr int EDGE_SENTINEL = 0x7fffffff;
template <int = 0>
__global__ void edge_var_init(int* __restrict__ first_bad, long long E) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x)
        first_bad[e] = EDGE_SENTINEL;
}
template <int = 0>
__global__ void edge_var_finish(int* __restrict__ first_bad, uint8_t* __restrict__ valid, long long E) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
        const int bad = first_bad[e];
        valid[e] = bad == EDGE_SENTINEL;
        if (bad == EDGE_SENTINEL) first_bad[e] = -1;
    }
}
template <class RB, bool GRID, bool SCR = false>
__global__ void __launch_bounds__(NT, SCR ? 4 : CBF_COLLIDE_MIN_CTAS) edge_var_kernel(SceneArgs sc, const float* __restrict__ q_from,
                                                      const float* __restrict__ q_to, long long E,
                                                      const long long* __restrict__ off, int K_all, long long P,
                                                      float thr_soft, float thr_hard, int* __restrict__ first_bad) {
    constexpr int N = RB::N;
    extern __shared__ __align__(16) float smem_raw[];
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, 0, 0);
    stage<N>(s, sc, nullptr, 0);
    SceneS sv = scene_view(s, sc, thr_soft, thr_hard);
    if constexpr (SCR) { sv.scratch = s.scratch + threadIdx.x; sv.scratch_stride = NT; }
    const float base_min = base_min_cta<RB>(s, sv);
    for (long long p = (long long)blockIdx.x * NT + threadIdx.x; p < P; p += (long long)gridDim.x * NT) {
        long long e;
        int K, k;
        if (off == nullptr) {                               // common K
            K = K_all;
            e = p / (K + 1);
            k = (int)(p - e * (K + 1));
        } else {
            long long lo = 0, hi = E;                       // largest e with off[e] <= p
            while (hi - lo > 1) {
                const long long mid = (lo + hi) >> 1;
                if (__ldg(off + mid) <= p) lo = mid; else hi = mid;
            }
            e = lo;
            const long long o0 = __ldg(off + e);
            K = (int)(__ldg(off + e + 1) - o0) - 1;
            k = (int)(p - o0);
        }
        float c[N];
        if (k == K) {
#pragma unroll
            for (int i = 0; i < N; ++i) c[i] = __ldg(q_to + e * N + i);
        } else {
            const float t = __fdiv_rn((float)k, (float)K);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const float a = __ldg(q_from + e * N + i), bb = __ldg(q_to + e * N + i);
                c[i] = fmaf(t, __fsub_rn(bb, a), a);
            }
        }
        CollideResult<RB> r;
        collide_state<RB, false, false, GRID, true, SCR>(c, sv, base_min, r);
        const bool unsafe = r.hard_min <= thr_hard;
        const bool safe = r.self_ok && (r.soft_min > thr_soft) && !unsafe;
        if (!safe) atomicMin(first_bad + e, k);
    }
}

template <class RB, int H, int L, bool GRID>
__global__ void __launch_bounds__(NT) rollout_kernel(SceneArgs sc, NetArgs net, const float* __restrict__ q0,
                                                     long long T, RolloutArgs ra, float thr_soft, float thr_hard,
                                                     float* __restrict__ q_final, uint8_t* __restrict__ alive,
                                                     int* __restrict__ steps_done, float* __restrict__ traj) {
    constexpr int N = RB::N;
    extern __shared__ __align__(16) float smem_raw[];
    using LY = NetLayout<N + 1, H, L>;
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, LY::TOTAL, ParamLayout<N>::TOTAL);
    stage<N>(s, sc, &net, LY::TOTAL);
    const SceneS sv = scene_view(s, sc, thr_soft, thr_hard);
    const float base_min = base_min_cta<RB>(s, sv);
    float* scr = s.scratch + threadIdx.x;
    for (long long b = (long long)blockIdx.x * NT + threadIdx.x; b < T; b += (long long)gridDim.x * NT) {
        float q[N], qn[N], g[N], u[N], dodq[N], o;
#pragma unroll
        for (int k = 0; k < N; ++k) { q[k] = __ldg(q0 + b * N + k); u[k] = 0.f; }
        load_goal<N>(net, b, g);
        CollideResult<RB> cr;
        collide_state<RB, true, false, GRID, false>(q, sv, base_min, cr);
        o = cr.o;
#pragma unroll
        for (int k = 0; k < N; ++k) dodq[k] = cr.dodq[k];
        bool ok = true, dead = false;
        int done = 0;
        float* tr = traj ? traj + b * (long long)ra.steps * N : nullptr;
        for (int t = 0; t < ra.steps; ++t) {
            if (t % ra.ctrl_every == 0) {
                float J[N], h, r;
                cbf_filter_state<N, H, L, NT>(s.weights, s.params, scr, q, o, dodq, g, false, net.lam, net.penalty, h,
                                              J, u, r);
            }
#pragma unroll
            for (int k = 0; k < N; ++k) qn[k] = fmaf(u[k], ra.dt, q[k]);  // arm_dynamics.py:508 (step = 1)
            if ((t + 1) % ra.ctrl_every == 0) {       // observation refresh: minimum + gradient needed
                collide_state<RB, true, false, GRID, false>(qn, sv, base_min, cr);
            } else {                                  // only the hard collision test is read: mask-only, no self test
                collide_state<RB, false, false, GRID, false>(qn, sv, base_min, cr);
            }
            if ((t + 1) % ra.ctrl_every == 0) {
                o = cr.o;
#pragma unroll
                for (int k = 0; k < N; ++k) dodq[k] = cr.dodq[k];
            } else {
                // closed_loop_dynamics(update_observation=False) leaves o / do_dq at zero (arm_dynamics.py:491-493);
                // they are not read before the next refresh, the controller only runs at t % ctrl_every == 0.
            }
            if (rollout_after_step<N>(ra, t, cr.hard_min <= thr_hard, q, qn, tr, ok, dead, done)) break;
        }
#pragma unroll
        for (int k = 0; k < N; ++k) q_final[b * N + k] = q[k];
        alive[b] = ok;
        steps_done[b] = done;
    }
}

// CTA-cooperative barrier network for planner-sized batches: the (row, neuron) outputs of every layer are spread over
// the 128 threads of the CTA instead of one thread walking all ~20 k multiply-adds of its row (with ONE row per call --
// the reference's regime -- 127 threads used to idle).  Activations / back-propagated gradients of a chunk of COOP_ROWS
// rows live in shared memory; forward layers read a transposed copy of the hidden weights (lanes over output neurons:
// conflict-free), the reverse pass the packed row-major weights (lanes over input neurons: conflict-free).  FP32 FMAs,
// tolerance-checked like the other FP32 path.  All NT threads must call it; `x` / outputs are per row (thread b < T).
constexpr int COOP_ROWS = 32;
template <int N_IN, int H>
struct CoopLayout {   // floats
    static constexpr int IN_PAD = (N_IN + 3) & ~3;
    static constexpr int WT = 0;                                   // [2][H (k)][H (j)] transposed hidden weights
    static constexpr int X = WT + 2 * H * H;                       // [COOP_ROWS][IN_PAD]
    static constexpr int A = X + COOP_ROWS * IN_PAD;               // [3][COOP_ROWS][H]  relu outputs of the three layers
    static constexpr int G = A + 3 * COOP_ROWS * H;                // [2][COOP_ROWS][H]  gradients (ping-pong)
    static constexpr int HJ = G + 2 * COOP_ROWS * H;               // [COOP_ROWS][IN_PAD + 1]  h, jac
    static constexpr int TOTAL = HJ + COOP_ROWS * (IN_PAD + 4);
};
template <int N_IN, int H, int NTH>
__device__ __forceinline__ void coop_setup(const float* __restrict__ w, float* __restrict__ cb) {
    using LY = NetLayout<N_IN, H, 2>;
    using CL = CoopLayout<N_IN, H>;
    for (int l = 0; l < 2; ++l)
        for (int i = threadIdx.x; i < H * H; i += NTH) {
            const int j = i / H, k = i % H;
            cb[CL::WT + l * H * H + k * H + j] = w[LY::HID0 + l * LY::HID_STRIDE + i];
        }
    __syncthreads();
}
template <int N_IN, int H, int NTH>
__device__ __forceinline__ void coop_mlp_h_jac(const float* __restrict__ w, float* __restrict__ cb, int T, bool active,
                                               const float (&x)[N_IN], float& h_out, float (&jac)[N_IN]) {
    using LY = NetLayout<N_IN, H, 2>;
    using CL = CoopLayout<N_IN, H>;
    const int tid = threadIdx.x;
    for (int r0 = 0; r0 < T; r0 += COOP_ROWS) {
        const int rows = min(COOP_ROWS, T - r0);
        const bool mine = active && tid >= r0 && tid < r0 + rows;
        __syncthreads();                                           // previous chunk's readers are done
        if (mine) {
#pragma unroll
            for (int k = 0; k < N_IN; ++k) cb[CL::X + (tid - r0) * CL::IN_PAD + k] = x[k];
        }
        __syncthreads();
        float* A0 = cb + CL::A;
        float* A1 = A0 + COOP_ROWS * H;
        float* A2 = A1 + COOP_ROWS * H;
        float* G0 = cb + CL::G;
        float* G1 = G0 + COOP_ROWS * H;
        for (int idx = tid; idx < rows * H; idx += NTH) {          // input layer
            const int r = idx / H, j = idx % H;
            float sacc = w[LY::B_IN + j];
#pragma unroll
            for (int k = 0; k < N_IN; ++k) sacc = fmaf(w[LY::W_IN + j * LY::IN_PAD + k], cb[CL::X + r * CL::IN_PAD + k], sacc);
            A0[r * H + j] = fmaxf(sacc, 0.f);
        }
        __syncthreads();
#pragma unroll
        for (int l = 0; l < 2; ++l) {                              // hidden layers (transposed weights: lanes over j)
            const float* in = l ? A1 : A0;
            float* out = l ? A2 : A1;
            const float* wt = cb + CL::WT + l * H * H;
            const float* bias = w + LY::HID0 + l * LY::HID_STRIDE + H * H;
            for (int idx = tid; idx < rows * H; idx += NTH) {
                const int r = idx / H, j = idx % H;
                float s0 = bias[j], s1 = 0.f;
#pragma unroll 8
                for (int k = 0; k < H; k += 2) {
                    s0 = fmaf(wt[k * H + j], in[r * H + k], s0);
                    s1 = fmaf(wt[(k + 1) * H + j], in[r * H + k + 1], s1);
                }
                out[r * H + j] = fmaxf(s0 + s1, 0.f);
            }
            __syncthreads();
        }
        for (int idx = tid; idx < rows * H; idx += NTH) {          // seed of the reverse pass; h by the row's first lane
            const int r = idx / H, k = idx % H;
            G0[r * H + k] = A2[r * H + k] > 0.f ? w[LY::W_OUT + k] : 0.f;
        }
        for (int r = tid; r < rows; r += NTH) {
            float hh = w[LY::B_OUT];
            for (int k = 0; k < H; ++k) hh = fmaf(w[LY::W_OUT + k], A2[r * H + k], hh);
            cb[CL::HJ + r * (CL::IN_PAD + 4)] = hh;
        }
        __syncthreads();
#pragma unroll
        for (int l = 1; l >= 0; --l) {                             // reverse mode through the hidden layers (row-major: lanes over k)
            const float* gin = l ? G0 : G1;
            float* gout = l ? G1 : G0;
            const float* mask = l ? A1 : A0;
            const float* W = w + LY::HID0 + l * LY::HID_STRIDE;
            for (int idx = tid; idx < rows * H; idx += NTH) {
                const int r = idx / H, k = idx % H;
                float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
                for (int j = 0; j < H; j += 2) {
                    s0 = fmaf(W[j * H + k], gin[r * H + j], s0);
                    s1 = fmaf(W[(j + 1) * H + k], gin[r * H + j + 1], s1);
                }
                gout[r * H + k] = mask[r * H + k] > 0.f ? s0 + s1 : 0.f;
            }
            __syncthreads();
        }
        for (int idx = tid; idx < rows * N_IN; idx += NTH) {       // input layer, reverse: jac = W_in^T g   (g = G0)
            const int r = idx / N_IN, k = idx % N_IN;
            float acc = 0.f;
            for (int j = 0; j < H; ++j) acc = fmaf(w[LY::W_IN + j * LY::IN_PAD + k], G0[r * H + j], acc);
            cb[CL::HJ + r * (CL::IN_PAD + 4) + 1 + k] = acc;
        }
        __syncthreads();
        if (mine) {
            const float* o = cb + CL::HJ + (tid - r0) * (CL::IN_PAD + 4);
            h_out = o[0];
#pragma unroll
            for (int k = 0; k < N_IN; ++k) jac[k] = o[1 + k];
        }
    }
}

// A whole expansion of the batched planner in ONE launch: n_seg consecutive batched steer segments
// (batch_tsa.py:154-170 over :191-231) for ONE batch of T <= NT rows in a single CTA.  Per segment: fresh observation of
// every row, the mode-1 step logic of rollout_after_step (stuck rows count as collisions, dead rows keep integrating),
// and the reference's batch-wide early exit -- the segment's loop ends once EVERY row is dead (:229-230) -- which is why
// the rows of a batch must share a CTA (__syncthreads_or over the alive flags; the loops are uniform over the CTA).
// The planner's nine 20-step segments per expansion become one launch and one read-back instead of nine.
template <class RB, int H, int L, bool GRID>
__global__ void __launch_bounds__(NT) rollout_segments_kernel(SceneArgs sc, NetArgs net, const float* __restrict__ q0, int T,
                                                              int n_seg, int seg_len, RolloutArgs ra, float thr_soft,
                                                              float thr_hard, float* __restrict__ seg_q,
                                                              uint8_t* __restrict__ seg_alive, int* __restrict__ seg_done,
                                                              float* __restrict__ traj) {
    constexpr int N = RB::N;
    extern __shared__ __align__(16) float smem_raw[];
    using LY = NetLayout<N + 1, H, L>;
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, LY::TOTAL, ParamLayout<N>::TOTAL);
    stage<N>(s, sc, &net, LY::TOTAL);
    const SceneS sv = scene_view(s, sc, thr_soft, thr_hard);
    const float base_min = base_min_cta<RB>(s, sv);
    float* const coop = s.scratch;                     // CoopLayout (cooperative barrier network, above)
    coop_setup<N + 1, H, NT>(s.weights, coop);
    const int b = threadIdx.x;
    const bool active = b < T;
    float q[N], qn[N], g[N], u[N], dodq[N], o = 0.f;
#pragma unroll
    for (int k = 0; k < N; ++k) { q[k] = active ? __ldg(q0 + b * N + k) : 0.f; u[k] = 0.f; g[k] = 0.f; dodq[k] = 0.f; }
    if (active) load_goal<N>(net, b, g);
    CollideResult<RB> cr;
    const long long steps = (long long)n_seg * seg_len;
    for (int sg = 0; sg < n_seg; ++sg) {
        if (active) {                                  // complete_sample_with_observations at the start of the segment
            collide_state<RB, true, false, GRID, false>(q, sv, base_min, cr);
            o = cr.o;
#pragma unroll
            for (int k = 0; k < N; ++k) dodq[k] = cr.dodq[k];
        }
        bool ok = true, dead = !active;
        int done = 0;
        float* tr = traj + ((long long)b * steps + (long long)sg * seg_len) * N;
        for (int t = 0; t < seg_len; ++t) {
            if (t % ra.ctrl_every == 0) {              // uniform over the CTA: all threads work on the rows' networks
                float x[N + 1], jac[N + 1], h = 0.f;
#pragma unroll
                for (int k = 0; k < N; ++k) x[k] = q[k];
                x[N] = o;
                coop_mlp_h_jac<N + 1, H, NT>(s.weights, coop, T, active, x, h, jac);
                if (active) {
                    float J[N], r;
                    cbf_filter_post<N>(s.params, q, dodq, g, false, net.lam, net.penalty, h, jac, J, u, r);
                }
            }
            if (active) {
#pragma unroll
                for (int k = 0; k < N; ++k) qn[k] = fmaf(u[k], ra.dt, q[k]);
                if ((t + 1) % ra.ctrl_every == 0) {
                    collide_state<RB, true, false, GRID, false>(qn, sv, base_min, cr);
                    o = cr.o;
#pragma unroll
                    for (int k = 0; k < N; ++k) dodq[k] = cr.dodq[k];
                } else {
                    collide_state<RB, false, false, GRID, false>(qn, sv, base_min, cr);
                }
                rollout_after_step<N>(ra, t, cr.hard_min <= thr_hard, q, qn, tr, ok, dead, done);   // stuck_mode 1, keep_going 1
            }
            if (!__syncthreads_or(!dead)) break;       // batch_tsa.py:229-230: every row of the batch is dead
        }
        if (active) {
#pragma unroll
            for (int k = 0; k < N; ++k) seg_q[((long long)b * n_seg + sg) * N + k] = q[k];
            seg_alive[(long long)b * n_seg + sg] = ok;
            seg_done[(long long)b * n_seg + sg] = done;
        }
    }
}

}  // namespace cbfrrt
#include "warp.cuh"   // one state per warp: the planner-sized kernels
namespace cbfrrt {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
template <class RB>
__global__ void __launch_bounds__(256) sample_kernel(uint32_t seed, unsigned long long start_row, long long B,
                                                     float* __restrict__ q) {
    const long long total = B * RB::N;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int j = (int)(i % RB::N);
        const unsigned long long idx = start_row * RB::N + (unsigned long long)i;
        uint32_t x = mix32((uint32_t)idx ^ (seed * 0x9E3779B9U));
        x = mix32(x + (uint32_t)(idx >> 32) + 0x85ebca6bU);
        const float u = __fmul_rn((float)(x >> 8), 5.9604644775390625e-8f);
        float lo = 0.f, hi = 0.f;
        static_for<0, RB::N>([&](auto jc) {
            constexpr int jj = decltype(jc)::value;
            if (j == jj) { lo = RB::qlo(jj); hi = RB::qhi(jj); }
        });
        q[i] = fmaf(u, __fsub_rn(hi, lo), lo);
    }
}

// FP32 FMA throughput probe (roofline denominator; see include/cbfrrt.h)
template <int = 0>
__global__ void fma_probe_kernel(int iters, float* sink) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f,
          a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f + blockIdx.x * 1e-9f, c = 1e-3f;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
        a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
    const float s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456f) *sink = s;  // never true; keeps the chains alive
}

// ------------------------------------------------------------------ host side of the device-pointer ABI
extern std::atomic<long long> g_launches;
extern std::atomic<int> g_last_cuda;

inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
}

// SM count of the CURRENT device (one process may drive several GPUs through cbfrrt_safety_filter_host)
inline int num_sms() {
    static std::atomic<int> sms[64];
    const int dev = current_device() & 63;
    int v = sms[dev].load(std::memory_order_relaxed);
    if (!v) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (v <= 0) v = 148;
        sms[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

// resident CTAs per SM for (device, kernel, dynamic smem): the attribute / occupancy queries cost ~10 us of host time,
// so they are made once per distinct configuration.  The opt-in to > 48 KB of dynamic shared memory is a per-device
// function attribute, hence the device in the key.
struct CfgKey {
    int dev;
    const void* kern;
    size_t smem;
    bool operator<(const CfgKey& o) const {
        if (dev != o.dev) return dev < o.dev;
        if (kern != o.kern) return kern < o.kern;
        return smem < o.smem;
    }
};
extern std::mutex g_cfg_mu;
extern std::map<CfgKey, int> g_cfg;

// Opt-in to more than 48 KB of dynamic shared memory.  cudaFuncAttributeMaxDynamicSharedMemorySize is ONE value per
// (device, kernel): it is only ever raised -- a launch with a small scene after a launch with a large one must not lower
// it under the size a cached configuration still relies on.  Call with g_cfg_mu held.
template <class Kern>
static bool smem_optin(Kern kern, size_t smem) {
    if (smem <= 48 * 1024) return true;
    const CfgKey key{current_device(), reinterpret_cast<const void*>(kern), ~size_t(0)};
    auto it = g_cfg.find(key);
    if (it != g_cfg.end() && (size_t)it->second >= smem) return true;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    g_cfg[key] = (int)smem;
    return true;
}

template <class Kern>
static int grid_for(Kern kern, size_t smem, long long tiles, int threads = NT) {
    int per_sm = 0;
    {
        std::lock_guard<std::mutex> lock(g_cfg_mu);
        const CfgKey key{current_device(), reinterpret_cast<const void*>(kern), smem};
        auto it = g_cfg.find(key);
        if (it != g_cfg.end()) per_sm = it->second;
        else {
            if (!smem_optin(kern, smem)) return -1;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) {
                cudaGetLastError();
                return -1;
            }
            g_cfg[key] = per_sm;
        }
    }
    long long g = (long long)num_sms() * per_sm;
    if (g > tiles) g = tiles;
    if (g < 1) g = 1;
    return (int)g;
}

// tensor-core kernels: one 512-thread CTA per SM (it owns all 512 TMEM columns), 4 tiles of 128 rows per pass.
// Weights (84 KB) + the groups' A_lo tiles (96 KB) leave ~45 KB for the scene: scenes beyond that (> ~900 obstacles)
// take the FP32-FMA kernels.
constexpr size_t TC_SMEM_MAX = 227 * 1024;
template <class Kern>
static int grid_for_tc(Kern kern, size_t smem, long long rows, int groups) {
    {
        std::lock_guard<std::mutex> lock(g_cfg_mu);
        if (!smem_optin(kern, smem)) return -1;
    }
    const long long tiles = (rows + tc::M_TILE - 1) / tc::M_TILE;
    long long g = (tiles + groups - 1) / groups;
    if (g > num_sms()) g = num_sms();
    return (int)(g < 1 ? 1 : g);
}

inline int finish_launch() {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_cuda = (int)e;
        return CBFRRT_ERR_CUDA;
    }
    g_launches++;
    return CBFRRT_OK;
}

constexpr int GRID_MIN_OBSTACLES = 12;

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int check_scene(const cbfrrt_scene* sc, SceneArgs& a) {
    if (!sc || sc->n_boxes < 0 || sc->n_spheres < 0) return CBFRRT_ERR_BAD_SHAPE;
    if (sc->n_boxes + sc->n_spheres > CBFRRT_MAX_OBSTACLES) return CBFRRT_ERR_UNSUPPORTED;
    if ((sc->n_boxes && !sc->boxes) || (sc->n_spheres && !sc->spheres)) return CBFRRT_ERR_BAD_SHAPE;
    if (!aligned16(sc->boxes) || !aligned16(sc->spheres) || !aligned16(sc->grid)) return CBFRRT_ERR_ALIGNMENT;
    a = SceneArgs{sc->boxes, sc->spheres, sc->n_boxes, sc->n_spheres, sc->floor ? 1 : 0,
                  static_cast<const uint16_t*>(sc->grid), static_cast<unsigned long long*>(sc->stats)};
    return CBFRRT_OK;
}

// When to use the broad phase a caller supplied (results are bit-identical either way).  Measured on B200
// (profiles/r2c_configs_*.jsonl, r2j_configs*.jsonl): from about a dozen obstacles on the distance-bound field wins for
// single-shot kernels (16 boxes, Magician: masks 0.176 -> 0.119 ms, fused 0.416 -> 0.373 ms per 2^20 states; 256 boxes,
// Panda: masks 2.5x, valid+u 1.8x over the round-1 candidate grid).  Rollouts are long dependent chains of small sweeps:
// there the shared-memory sweep with ball culling (no global loads at all) wins until the scene is dense (8 boxes: 17 vs
// 35 ms), so they switch at GRID_MIN_OBSTACLES_ROLLOUT.  Planner-sized launches (<= SMALL_BATCH rows: cold instruction
// cache, one wave) are faster on the field path's much smaller instruction stream whatever the scene (4096 states,
// 1 box: 68 -> 42 us).  CBFRRT_GRID_MIN overrides both thresholds.
constexpr long long SMALL_BATCH = 16384;
constexpr int GRID_MIN_OBSTACLES_ROLLOUT = 40;
inline SceneArgs scene_for(const SceneArgs& sc, long long rows, bool is_rollout) {
    static const int grid_min_env = [] { const char* e = getenv("CBFRRT_GRID_MIN"); return e ? atoi(e) : -1; }();
    const int grid_min = grid_min_env >= 0 ? grid_min_env : (is_rollout ? GRID_MIN_OBSTACLES_ROLLOUT : GRID_MIN_OBSTACLES);
    SceneArgs a = sc;
    const int n_obs = sc.n_boxes + sc.n_spheres;
    const bool use = sc.grid && (n_obs >= grid_min || (!is_rollout && rows <= SMALL_BATCH && n_obs >= 1));
    if (!use) a.grid = nullptr;
    return a;
}

constexpr long long GEN_MAX_UNITS = 1536;   // (layers + 2) * hidden: 32 threads x 1536 floats = 192 KB of shared memory
inline bool net_is_fast(const cbfrrt_cbf_net* net) { return net->hidden == 48 && net->layers == 2; }

inline int check_net(int n, const cbfrrt_cbf_net* net, const cbfrrt_filter_params* fp, long long B, NetArgs& a) {
    if (!net || !fp || !net->weights || !fp->u_lo || !fp->u_hi) return CBFRRT_ERR_BAD_SHAPE;
    if (!fp->u_ref && (!fp->goal || !fp->K)) return CBFRRT_ERR_BAD_SHAPE;
    if (net->n_in != n + 1) return CBFRRT_ERR_BAD_SHAPE;
    if (net->hidden < 1 || net->layers < 0) return CBFRRT_ERR_BAD_SHAPE;
    // 48 x 2 has the specialised kernels; other shapes run filter_kernel_generic (cbf_filter, safety_filter) as long as
    // a 32-thread CTA's activations fit in shared memory; the fused rollout exists for 48 x 2 only
    if (!net_is_fast(net) && (long long)(net->layers + 2) * net->hidden > GEN_MAX_UNITS) return CBFRRT_ERR_UNSUPPORTED;
    if (!fp->u_ref && fp->goal_rows != 1 && fp->goal_rows != B) return CBFRRT_ERR_BAD_SHAPE;
    if (!aligned16(net->weights)) return CBFRRT_ERR_ALIGNMENT;
    a = NetArgs{net->weights, fp->K, fp->u_lo, fp->u_hi, fp->goal, (long long)fp->goal_rows, fp->cbf_lambda,
                fp->relaxation_penalty, fp->u_ref};
    return CBFRRT_OK;
}

// Batches of at most SMALL_BATCH rows (< 1 tile group per SM) are launch-latency bound: the tcgen05 kernels' per-CTA
// set-up (two weight tilings, TMEM allocation) costs more than it saves there, so they take the FP32-FMA kernels
// (measured on cfg0, 4096 rows: 59 us vs 75 us).  cbfrrt_set_mlp_path(CBFRRT_MLP_TC) forces the tensor-core kernels.
// MLP path: CBFRRT_MLP_AUTO (tcgen05 kernels above SMALL_BATCH rows, FP32-FMA kernels below), _FMA or _TC for every
// size.  Set through cbfrrt_set_mlp_path() (include/cbfrrt.h); the environment variable CBFRRT_MLP=fma|tc only gives
// the initial value.  Read at every launch, so one process can exercise both paths (the parity tests do).
extern std::atomic<int> g_mlp_path;
inline bool use_tc() { return g_mlp_path.load(std::memory_order_relaxed) != CBFRRT_MLP_FMA; }
inline bool force_tc() { return g_mlp_path.load(std::memory_order_relaxed) == CBFRRT_MLP_TC; }

// One state per warp (warp.cuh) for planner-sized calls.  Measured on B200 (profiles/r2r_planner_sized.jsonl, device time
// by CUDA-graph replay): a one-thread-per-state launch is one dependent chain per state -- fused filter, Panda, 8 boxes:
// 26 us for 1 state, 37-42 us from 64 to 16 384 states; 200-step rollouts 3.5 ms whatever the batch -- while the warp form
// is bound by its total instruction count: 8.8 us for 1 state, 11.3 us for 64 ... 1024, 22.5 us at 4096, 38.9 us at 8192
// (cross-over), rollouts 0.43 / 0.57 / 1.17 / 1.72 ms for 64 / 2048 / 4096 / 8192 trajectories.  CBFRRT_MLP_WARP forces it
// at every size, CBFRRT_MLP_FMA / _TC keep the one-thread-per-state kernels (the parity tests run every path).
constexpr long long WARP_MAX_ROWS = 4096;       // masks / observations / fused filter / edge points: rows per call
constexpr long long WARP_MAX_TRAJ = 8192;       // closed-loop rollouts: trajectories per call
constexpr int WARP_GRID_MIN_OBSTACLES = 48;     // below it a lane sweeps the whole staged scene (no global look-ups)
inline bool use_warp(long long rows, long long max_rows) {
    const int p = g_mlp_path.load(std::memory_order_relaxed);
    if (p == CBFRRT_MLP_WARP) return true;
    if (p != CBFRRT_MLP_AUTO) return false;
    static const long long env_max = [] { const char* e = getenv("CBFRRT_WARP_MAX"); return e ? atoll(e) : -1LL; }();
    return rows <= (env_max >= 0 ? env_max : max_rows);
}
inline SceneArgs scene_for_warp(const SceneArgs& sc) {
    SceneArgs a = sc;
    if (sc.n_boxes + sc.n_spheres < WARP_GRID_MIN_OBSTACLES) a.grid = nullptr;
    a.stats = nullptr;
    return a;
}

template <class RB>
static int launch_fk(const float* q, long long qs, long long B, float* pos, float* rot, cudaStream_t st) {
    const long long tiles = (B + NT - 1) / NT;
    int grid = (int)(tiles < (long long)num_sms() * 8 ? tiles : (long long)num_sms() * 8);
    fk_kernel<RB><<<grid, NT, 0, st>>>(q, qs, B, pos, rot);
    return finish_launch();
}

template <class RB>
static int launch_collide(const SceneArgs& sc_in, const float* q, long long qs, long long B, float ts, float th,
                          const CollideOutPtrs& out, cudaStream_t st) {
    const bool grad = out.datax || out.arg_link || out.arg_obs;
    int grid;
    if (use_warp(B, WARP_MAX_ROWS) && !sc_in.stats) {
        const SceneArgs sc = scene_for_warp(sc_in);
        const size_t smem = 4 * smem_floats(sc.n_boxes, sc.n_spheres, 0, 0, 0);
        const long long tiles = (B + wp::WARPS - 1) / wp::WARPS;
#define CBF_LAUNCH_COLLIDE_W(G_, S_)                                                       \
    {                                                                                      \
        auto k = wp::collide_kernel_warp<RB, G_, S_>;                                      \
        if ((grid = grid_for(k, smem, tiles, wp::WNT)) < 0) return CBFRRT_ERR_CUDA;        \
        k<<<grid, wp::WNT, smem, st>>>(sc, q, qs, B, ts, th, out);                         \
    }
        if (out.mins) CBF_LAUNCH_COLLIDE_W(true, true)
        else if (grad) CBF_LAUNCH_COLLIDE_W(true, false)
        else CBF_LAUNCH_COLLIDE_W(false, false)
#undef CBF_LAUNCH_COLLIDE_W
        return finish_launch();
    }
    const SceneArgs sc = scene_for(sc_in, B, false);
    const size_t smem = 4 * smem_floats(sc.n_boxes, sc.n_spheres, 0, 0, 0);
    const long long tiles = (B + NT - 1) / NT;
#define CBF_LAUNCH_COLLIDE(G_, S_, U_)                                             \
    {                                                                              \
        auto k = collide_kernel<RB, G_, S_, U_>;                                   \
        if ((grid = grid_for(k, smem, tiles)) < 0) return CBFRRT_ERR_CUDA;         \
        k<<<grid, NT, smem, st>>>(sc, q, qs, B, ts, th, out);                      \
    }
    const bool ug = sc.grid != nullptr;
    // Field path with the sphere-centre scratch in shared memory instead of local memory (CBFRRT_SMEM_SCRATCH=1; four CTAs
    // per SM must still fit).  Measured per 2^23 Panda states, 256 boxes (profiles/r3a_* vs r3b_*): DRAM traffic 108 instead
    // of 478 B per state (89 compulsory: the rest of the default's traffic is write-back of that scratch and of register
    // spills at 96 registers) but 3.44 instead of 3.19 ms -- five CTAs of latency hiding beat the cleaner memory picture
    // while DRAM runs at a fifth of its bandwidth, so the local-memory form stays the default.
    static const bool smem_scratch = [] { const char* e = getenv("CBFRRT_SMEM_SCRATCH"); return e && e[0] == '1'; }();
    const size_t smem_scr = smem + 4 * (size_t)scr_floats<RB>();
    const bool scr_ok = smem_scratch && ug && !out.mins && B > SMALL_BATCH && smem_scr + 1024 <= (227 * 1024) / 4;
#define CBF_LAUNCH_COLLIDE_SCR(G_)                                                 \
    {                                                                              \
        auto k = collide_kernel<RB, G_, false, true, true>;                        \
        if ((grid = grid_for(k, smem_scr, tiles)) < 0) return CBFRRT_ERR_CUDA;     \
        k<<<grid, NT, smem_scr, st>>>(sc, q, qs, B, ts, th, out);                  \
    }
    if (out.mins) { if (ug) CBF_LAUNCH_COLLIDE(true, true, true) else CBF_LAUNCH_COLLIDE(true, true, false) }
    else if (scr_ok) { if (grad) CBF_LAUNCH_COLLIDE_SCR(true) else CBF_LAUNCH_COLLIDE_SCR(false) }
    else if (grad) { if (ug) CBF_LAUNCH_COLLIDE(true, false, true) else CBF_LAUNCH_COLLIDE(true, false, false) }
    else { if (ug) CBF_LAUNCH_COLLIDE(false, false, true) else CBF_LAUNCH_COLLIDE(false, false, false) }
#undef CBF_LAUNCH_COLLIDE_SCR
#undef CBF_LAUNCH_COLLIDE
    return finish_launch();
}

template <int N>
static int launch_filter_generic(const NetArgs& net, int H, int L, const float* datax, long long B, const FilterOutPtrs& out,
                                 cudaStream_t st) {
    const size_t smem = 4 * ((size_t)ParamLayout<N>::TOTAL + (size_t)GEN_NT * (L + 2) * H);
    auto k = filter_kernel_generic<N>;
    {
        std::lock_guard<std::mutex> lock(g_cfg_mu);
        if (!smem_optin(k, smem)) return CBFRRT_ERR_CUDA;
    }
    const long long tiles = (B + GEN_NT - 1) / GEN_NT;
    const long long cap = (long long)num_sms() * 16;
    k<<<(int)(tiles < cap ? tiles : cap), GEN_NT, smem, st>>>(net, H, L, datax, B, out);
    return finish_launch();
}

template <int N>
static int launch_filter(const NetArgs& net, const float* datax, long long B, const FilterOutPtrs& out, cudaStream_t st) {
    if (use_warp(B, WARP_MAX_ROWS)) {
        using LYW = NetLayout<N + 1, 48, 2>;
        const size_t smem = 4 * smem_floats(0, 0, LYW::TOTAL, ParamLayout<N>::TOTAL, wp::WarpNet<N + 1, 48>::TOTAL);
        auto k = wp::filter_kernel_warp<N, 48>;
        const int grid = grid_for(k, smem, (B + wp::WARPS - 1) / wp::WARPS, wp::WNT);
        if (grid < 0) return CBFRRT_ERR_CUDA;
        k<<<grid, wp::WNT, smem, st>>>(net, datax, B, out);
        return finish_launch();
    }
    if (use_tc() && (B > SMALL_BATCH || force_tc())) {
        const size_t smem = 4 * smem_floats(0, 0, tc::Layout<2, 4>::TOTAL, ParamLayout<N>::TOTAL, 0);
        auto k = filter_kernel_tc<N, 2, false>;
        const int grid = grid_for_tc(k, smem, B, 4);
        if (grid < 0) return CBFRRT_ERR_CUDA;
        k<<<grid, 4 * tc::M_TILE, smem, st>>>(net, datax, B, out, PeerOut{}, nullptr, 0);
        return finish_launch();
    }
    using LY = NetLayout<N + 1, 48, 2>;
    const size_t smem = 4 * smem_floats(0, 0, LY::TOTAL, ParamLayout<N>::TOTAL, 48 * NT);
    auto k = filter_kernel<N, 48, 2>;
    const int grid = grid_for(k, smem, (B + NT - 1) / NT);
    if (grid < 0) return CBFRRT_ERR_CUDA;
    k<<<grid, NT, smem, st>>>(net, datax, B, out);
    return finish_launch();
}

// ---- per-robot launch wrappers: each is defined in its own translation unit (k_<family>_<robot>.cu) so that the
// big template instantiations compile in parallel
#define CBF_DECLARE_ROBOT(R)                                                                                          \
    int launch_collide_##R(const SceneArgs&, const float*, long long, long long, float, float, const CollideOutPtrs&,  \
                           cudaStream_t);                                                                              \
    int launch_safety_##R(const SceneArgs&, const NetArgs&, const float*, long long, long long, float, float,          \
                          const CollideOutPtrs&, const FilterOutPtrs&, const PeerOut&, cudaStream_t);                  \
    int launch_edges_##R(const SceneArgs&, const float*, const float*, long long, int, float, float, uint8_t*, int*,   \
                         cudaStream_t);                                                                                \
    int launch_edges_var_##R(const SceneArgs&, const float*, const float*, long long, const long long*, long long,     \
                             float, float, uint8_t*, int*, cudaStream_t);                                              \
    int launch_rollout_##R(const SceneArgs&, const NetArgs&, const float*, long long, const RolloutArgs&, float,       \
                           float, float*, uint8_t*, int*, float*, cudaStream_t);                                       \
    int launch_rollout_segments_##R(const SceneArgs&, const NetArgs&, const float*, int, int, int, const RolloutArgs&,  \
                                    float, float, float*, uint8_t*, int*, float*, cudaStream_t);
CBF_DECLARE_ROBOT(panda)
CBF_DECLARE_ROBOT(magician)
#undef CBF_DECLARE_ROBOT


// Two-pass form of the large sweeps.  The fused kernel runs the geometry at the occupancy the barrier network dictates
// (one 512-thread CTA per SM: TMEM, 180 KB of weights and A_lo tiles, 128 registers) and measures 5.65 ms per 2^23 Panda
// states on the 256-box scene; the same work as a geometry pass at its own occupancy (collide_kernel -> masks + datax)
// followed by a barrier-network pass through datax (filter_kernel_tc -> u) takes 3.2 + 1.5 = 4.7 ms although it moves
// 120 B per state more (profiles/r2w_split_pipeline_probe.json; the outputs are bit-identical: same device functions).
// It needs somewhere to put datax (60 B per row; + 1 B per row for the mask in the multi-GPU exchange): the caller's
// datax output when requested, else the scratch registered with cbfrrt_set_workspace for this device.
constexpr long long SPLIT_MIN_ROWS = 1 << 19;   // per-call times, Panda 256 boxes: 2^18 rows 0.183 (two-pass) vs 0.177 ms (fused), 2^19 rows 0.323 vs 0.353 ms
struct Workspace {
    void* ptr;
    long long bytes;
};
extern Workspace g_workspace[64];   // guarded by g_cfg_mu (set by cbfrrt_set_workspace, read at every large launch)
inline Workspace workspace_of_current_device() {
    std::lock_guard<std::mutex> lock(g_cfg_mu);
    return g_workspace[current_device() & 63];
}
extern std::atomic<int> g_split;   // 1: two-pass form allowed (default), 0: always the fused kernel (CBFRRT_SPLIT=0, tests)

// the barrier-network pass with the fused exchange: datax -> u, and (u, safe_in) stored into every rank's replica
template <int N>
static int launch_filter_peers(const NetArgs& net, const float* dx, const uint8_t* safe_loc, long long B, const FilterOutPtrs& fo,
                               const PeerOut& po, cudaStream_t st) {
    const size_t smem = 4 * smem_floats(0, 0, tc::Layout<2, 4>::TOTAL, ParamLayout<N>::TOTAL, 0);
    int vec = 0;
    if (po.row_offset % 16 == 0) {
        bool ok = true;
        for (int p = 0; p < po.n; ++p) ok = ok && aligned16(po.valid[p]) && aligned16(po.u[p]);
        if (ok) vec |= VEC_PEER;
    }
    auto k = filter_kernel_tc<N, 2, true>;
    const int grid = grid_for_tc(k, smem, B, 4);
    if (grid < 0) return CBFRRT_ERR_CUDA;
    k<<<grid, 4 * tc::M_TILE, smem, st>>>(net, dx, B, fo, po, safe_loc, vec);
    return finish_launch();
}

template <class RB>
static int launch_safety_two_pass(const SceneArgs& sc, const NetArgs& net, const float* q, long long qs, long long B, float ts,
                                  float th, const CollideOutPtrs& co, const FilterOutPtrs& fo, const PeerOut& po, float* dx,
                                  uint8_t* safe_loc, cudaStream_t st) {
    constexpr int N = RB::N;
    CollideOutPtrs a = co;
    a.datax = dx;
    if (po.n) a.safe = safe_loc;
    int rc = RB::ID == 0 ? launch_collide_panda(sc, q, qs, B, ts, th, a, st) : launch_collide_magician(sc, q, qs, B, ts, th, a, st);
    if (rc) return rc;
    if (!fo.h && !fo.J && !fo.u && !fo.r && !po.n) return CBFRRT_OK;
    if (po.n) return launch_filter_peers<N>(net, dx, safe_loc, B, fo, po, st);
    const size_t smem = 4 * smem_floats(0, 0, tc::Layout<2, 4>::TOTAL, ParamLayout<N>::TOTAL, 0);
    auto k = filter_kernel_tc<N, 2, false>;
    const int grid = grid_for_tc(k, smem, B, 4);
    if (grid < 0) return CBFRRT_ERR_CUDA;
    k<<<grid, 4 * tc::M_TILE, smem, st>>>(net, dx, B, fo, po, nullptr, 0);
    return finish_launch();
}

template <class RB>
static int launch_safety(const SceneArgs& sc_in, const NetArgs& net, const float* q, long long qs, long long B, float ts,
                         float th, const CollideOutPtrs& co, const FilterOutPtrs& fo, const PeerOut& po, cudaStream_t st) {
    if (B >= SPLIT_MIN_ROWS && use_tc() && !force_tc() && g_split.load(std::memory_order_relaxed) && !sc_in.stats) {
        constexpr long long DXB = 4 * (2 * RB::N + 1);
        float* dx = co.datax;
        uint8_t* safe_loc = co.safe;
        const Workspace ws = workspace_of_current_device();
        const long long need = (dx ? 0 : ((B * DXB + 15) & ~15LL)) + ((po.n && !safe_loc) ? B : 0);
        if (need > 0 && ws.ptr && ws.bytes >= need) {
            char* w = static_cast<char*>(ws.ptr);
            if (!dx) { dx = reinterpret_cast<float*>(w); w += (B * DXB + 15) & ~15LL; }
            if (po.n && !safe_loc) safe_loc = reinterpret_cast<uint8_t*>(w);
        }
        if (dx && (!po.n || safe_loc)) return launch_safety_two_pass<RB>(sc_in, net, q, qs, B, ts, th, co, fo, po, dx, safe_loc, st);
    }
    if (use_warp(B, WARP_MAX_ROWS) && po.n == 0 && !sc_in.stats) {
        const SceneArgs sc = scene_for_warp(sc_in);
        using LYW = NetLayout<RB::N + 1, 48, 2>;
        const size_t smem = 4 * smem_floats(sc.n_boxes, sc.n_spheres, LYW::TOTAL, ParamLayout<RB::N>::TOTAL,
                                            wp::WarpNet<RB::N + 1, 48>::TOTAL);
        auto k = wp::safety_kernel_warp<RB, 48>;
        const int grid = grid_for(k, smem, (B + wp::WARPS - 1) / wp::WARPS, wp::WNT);
        if (grid < 0) return CBFRRT_ERR_CUDA;
        k<<<grid, wp::WNT, smem, st>>>(sc, net, q, qs, B, ts, th, co, fo);
        return finish_launch();
    }
    const SceneArgs sc = scene_for(sc_in, B, false);
    const int gg = 4;
    const size_t smem_tc = 4 * smem_floats(sc.n_boxes, sc.n_spheres, tc::Layout<2, 4>::TOTAL, ParamLayout<RB::N>::TOTAL, 0);
    if (use_tc() && (B > SMALL_BATCH || force_tc()) && smem_tc <= TC_SMEM_MAX) {
        const size_t smem = smem_tc;
        auto k = sc.grid ? safety_kernel_tc<RB, 2, true, 4> : safety_kernel_tc<RB, 2, false, 4>;
        const int grid = grid_for_tc(k, smem, B, gg);
        if (grid < 0) return CBFRRT_ERR_CUDA;
        // which pointer sets allow 16-byte tile accesses (see the kernel): rows are contiguous and every base aligned
        // Measured (profiles/r2e_configs.jsonl vs r2c_*_gg4): for LOCAL outputs the staged form is 5 % slower than the
        // per-row stores (cfg4 valid+u 5.33 -> 5.62 ms per 2^23 states: two more group barriers per tile, and the L2 absorbs
        // the partial lines), so it is used only where it pays -- the NVLink stores of the fused exchange, where a
        // 4-byte store is a whole fabric packet.  CBFRRT_TILE_IO=1 switches it on for q / local outputs too (A/B runs).
        static const bool tile_io = [] { const char* e = getenv("CBFRRT_TILE_IO"); return e && e[0] == '1'; }();
        int vec = 0;
        if (tile_io && qs == RB::N && aligned16(q)) vec |= VEC_Q;
        if (tile_io && aligned16(co.safe) && aligned16(co.unsafe) && aligned16(co.datax) && aligned16(fo.h) && aligned16(fo.J) &&
            aligned16(fo.u) && aligned16(fo.r) && !co.arg_link && !co.arg_obs && !co.mins) vec |= VEC_OUT;
        if (po.n > 0 && po.row_offset % 16 == 0) {
            bool ok = true;
            for (int p = 0; p < po.n; ++p) ok = ok && aligned16(po.valid[p]) && aligned16(po.u[p]);
            if (ok) vec |= VEC_PEER;
        }
        k<<<grid, gg * tc::M_TILE, smem, st>>>(sc, net, q, qs, B, ts, th, co, fo, po, vec);
        return finish_launch();
    }
    using LY = NetLayout<RB::N + 1, 48, 2>;
    const size_t smem = 4 * smem_floats(sc.n_boxes, sc.n_spheres, LY::TOTAL, ParamLayout<RB::N>::TOTAL, 48 * NT);
    auto k = sc.grid ? safety_kernel<RB, 48, 2, true> : safety_kernel<RB, 48, 2, false>;
    const int grid = grid_for(k, smem, (B + NT - 1) / NT);
    if (grid < 0) return CBFRRT_ERR_CUDA;
    k<<<grid, NT, smem, st>>>(sc, net, q, qs, B, ts, th, co, fo, po);
    return finish_launch();
}

template <class RB>
static int launch_edges_var(const SceneArgs& sc, const float* qf, const float* qt, long long E, const long long* off,
                            int K_all, long long P, float ts, float th, uint8_t* valid, int* first_bad, cudaStream_t st);

template <class RB>
static int launch_edges(const SceneArgs& sc, const float* qf, const float* qt, long long E, int K, float ts, float th,
                        uint8_t* valid, int* first_bad, cudaStream_t st) {
    return launch_edges_var<RB>(sc, qf, qt, E, nullptr, K, E * (long long)(K + 1), ts, th, valid, first_bad, st);
}

template <class RB>
static int launch_edges_var(const SceneArgs& sc_in, const float* qf, const float* qt, long long E, const long long* off,
                            int K_all, long long P, float ts, float th, uint8_t* valid, int* first_bad, cudaStream_t st) {
    const bool warp = use_warp(P, WARP_MAX_ROWS) && !sc_in.stats;
    const SceneArgs sc = warp ? scene_for_warp(sc_in) : scene_for(sc_in, P, false);
    const size_t smem = 4 * smem_floats(sc.n_boxes, sc.n_spheres, 0, 0, 0);
    const int eg = (int)((E + 255) / 256 < 1184 ? (E + 255) / 256 : 1184);
    edge_var_init<0><<<eg, 256, 0, st>>>(first_bad, E);
    if (finish_launch()) return CBFRRT_ERR_CUDA;
    if (P > 0 && warp) {
        auto k = wp::edge_var_kernel_warp<RB>;
        const int grid = grid_for(k, smem, (P + wp::WARPS - 1) / wp::WARPS, wp::WNT);
        if (grid < 0) return CBFRRT_ERR_CUDA;
        k<<<grid, wp::WNT, smem, st>>>(sc, qf, qt, E, off, K_all, P, ts, th, first_bad);
        if (finish_launch()) return CBFRRT_ERR_CUDA;
    } else if (P > 0) {
        const size_t smem_scr = smem + 4 * (size_t)scr_floats<RB>();
        static const bool smem_scratch = [] { const char* e = getenv("CBFRRT_SMEM_SCRATCH"); return e && e[0] == '1'; }();
        const bool scr_ok = smem_scratch && sc.grid && P > SMALL_BATCH && smem_scr + 1024 <= (227 * 1024) / 4;
        auto k = scr_ok ? edge_var_kernel<RB, true, true> : (sc.grid ? edge_var_kernel<RB, true> : edge_var_kernel<RB, false>);
        const size_t smem_k = scr_ok ? smem_scr : smem;
        const int grid = grid_for(k, smem_k, (P + NT - 1) / NT);
        if (grid < 0) return CBFRRT_ERR_CUDA;
        k<<<grid, NT, smem_k, st>>>(sc, qf, qt, E, off, K_all, P, ts, th, first_bad);
        if (finish_launch()) return CBFRRT_ERR_CUDA;
    }
    edge_var_finish<0><<<eg, 256, 0, st>>>(first_bad, valid, E);
    return finish_launch();
}

template <class RB>
static int launch_rollout(const SceneArgs& sc_in, const NetArgs& net, const float* q0, long long T, const RolloutArgs& ra,
                          float ts, float th, float* qf, uint8_t* alive, int* done, float* traj, cudaStream_t st) {
    if (use_warp(T, WARP_MAX_TRAJ)) {
        const SceneArgs sc = scene_for_warp(sc_in);
        using LYW = NetLayout<RB::N + 1, 48, 2>;
        const size_t smem = 4 * smem_floats(sc.n_boxes, sc.n_spheres, LYW::TOTAL, ParamLayout<RB::N>::TOTAL,
                                            wp::WarpNet<RB::N + 1, 48>::TOTAL + wp::WARPS * STUCK_RING * RB::N);
        auto k = wp::rollout_kernel_warp<RB, 48>;
        const int grid = grid_for(k, smem, (T + wp::WARPS - 1) / wp::WARPS, wp::WNT);
        if (grid < 0) return CBFRRT_ERR_CUDA;
        k<<<grid, wp::WNT, smem, st>>>(sc, net, q0, T, ra, ts, th, qf, alive, done, traj);
        return finish_launch();
    }
    const SceneArgs sc = scene_for(sc_in, T, true);
    const int gg = 4;
    const size_t smem_tc = 4 * smem_floats(sc.n_boxes, sc.n_spheres, tc::Layout<2, 4>::TOTAL, ParamLayout<RB::N>::TOTAL, 0);
    if (use_tc() && (T > SMALL_BATCH || force_tc()) && smem_tc <= TC_SMEM_MAX) {
        const size_t smem = smem_tc;
        auto k = sc.grid ? rollout_kernel_tc<RB, 2, true> : rollout_kernel_tc<RB, 2, false>;
        const int grid = grid_for_tc(k, smem, T, gg);
        if (grid < 0) return CBFRRT_ERR_CUDA;
        k<<<grid, gg * tc::M_TILE, smem, st>>>(sc, net, q0, T, ra, ts, th, qf, alive, done, traj);
        return finish_launch();
    }
    using LY = NetLayout<RB::N + 1, 48, 2>;
    const size_t smem = 4 * smem_floats(sc.n_boxes, sc.n_spheres, LY::TOTAL, ParamLayout<RB::N>::TOTAL, 48 * NT);
    auto k = sc.grid ? rollout_kernel<RB, 48, 2, true> : rollout_kernel<RB, 48, 2, false>;
    const int grid = grid_for(k, smem, (T + NT - 1) / NT);
    if (grid < 0) return CBFRRT_ERR_CUDA;
    k<<<grid, NT, smem, st>>>(sc, net, q0, T, ra, ts, th, qf, alive, done, traj);
    return finish_launch();
}

template <class RB>
static int launch_rollout_segments(const SceneArgs& sc_in, const NetArgs& net, const float* q0, int T, int n_seg, int seg_len,
                                   const RolloutArgs& ra, float ts, float th, float* seg_q, uint8_t* seg_alive, int* seg_done,
                                   float* traj, cudaStream_t st) {
    using LYW = NetLayout<RB::N + 1, 48, 2>;
    const size_t smem_w = 4 * smem_floats(sc_in.n_boxes, sc_in.n_spheres, LYW::TOTAL, ParamLayout<RB::N>::TOTAL,
                                          wp::seg_scratch_floats<RB, 48>(seg_len));
    if (use_warp(T, wp::SEG_MAX_ROWS) && smem_w <= TC_SMEM_MAX) {
        // one row per warp over a cluster of CTAs (two rows per warp beyond 64 rows)
        const SceneArgs sc = scene_for_warp(sc_in);
        const size_t smem = smem_w;
        auto k = wp::rollout_segments_kernel_warp<RB, 48>;
        if (grid_for(k, smem, 1, wp::WNT) < 0) return CBFRRT_ERR_CUDA;
        int n_cta = (T + wp::WARPS - 1) / wp::WARPS;
        if (n_cta > wp::SEG_MAX_CLUSTER) n_cta = wp::SEG_MAX_CLUSTER;
        if (n_cta < 1) n_cta = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(n_cta);
        cfg.blockDim = dim3(wp::WNT);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = n_cta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, k, sc, net, q0, T, n_seg, seg_len, ra, ts, th, seg_q, seg_alive, seg_done, traj) != cudaSuccess) {
            g_last_cuda = (int)cudaGetLastError();
            return CBFRRT_ERR_CUDA;
        }
        return finish_launch();
    }
    const SceneArgs sc = scene_for(sc_in, T, true);
    using LY = NetLayout<RB::N + 1, 48, 2>;
    const size_t smem = 4 * smem_floats(sc.n_boxes, sc.n_spheres, LY::TOTAL, ParamLayout<RB::N>::TOTAL, CoopLayout<RB::N + 1, 48>::TOTAL);
    auto k = sc.grid ? rollout_segments_kernel<RB, 48, 2, true> : rollout_segments_kernel<RB, 48, 2, false>;
    if (grid_for(k, smem, 1) < 0) return CBFRRT_ERR_CUDA;
    k<<<1, NT, smem, st>>>(sc, net, q0, T, n_seg, seg_len, ra, ts, th, seg_q, seg_alive, seg_done, traj);
    return finish_launch();
}

}  // namespace cbfrrt

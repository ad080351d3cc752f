// kernels.cu -- sm_100a kernels + the device-pointer C-ABI of libcbfrrt (include/cbfrrt.h).
//
// Launch shape: CTAs of NT=128 threads, persistent grid (SMs x resident CTAs), grid-stride over tiles of NT
// states so that the per-CTA staging (scene + weights by TMA bulk copy, base-link distance) is paid once.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdlib>

#include "../../include/cbfrrt.h"
#include "core.cuh"
#include "mlp_qp.cuh"
#include "mlp_tc.cuh"

namespace cbfrrt {

constexpr int NT = 128;

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
    return 4 + 12 + (size_t)8 * n_boxes + (size_t)4 * n_spheres + w_floats + p_floats + scr_floats;
}
__device__ __forceinline__ Smem carve(float* base, int n_boxes, int n_spheres, int w_floats, int p_floats) {
    Smem s;
    s.bar = reinterpret_cast<uint64_t*>(base);
    s.red = base + 4;
    s.boxes = base + 16;
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

__device__ __forceinline__ SceneS scene_view(const Smem& s, const SceneArgs& sc) {
    SceneS v;
    v.boxes = reinterpret_cast<const float4*>(s.boxes);
    v.spheres = reinterpret_cast<const float4*>(s.spheres);
    v.n_boxes = sc.n_boxes; v.n_spheres = sc.n_spheres; v.floor = sc.floor;
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

template <class RB, bool WANT_GRAD, bool WANT_SELFMIN>
__global__ void __launch_bounds__(NT) collide_kernel(SceneArgs sc, const float* __restrict__ q, long long q_stride,
                                                     long long B, float thr_soft, float thr_hard, CollideOutPtrs out) {
    extern __shared__ __align__(16) float smem_raw[];
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, 0, 0);
    stage<RB::N>(s, sc, nullptr, 0);
    const SceneS sv = scene_view(s, sc);
    const float base_min = base_min_cta<RB>(s, sv);
    for (long long b = (long long)blockIdx.x * NT + threadIdx.x; b < B; b += (long long)gridDim.x * NT) {
        float qq[RB::N];
#pragma unroll
        for (int k = 0; k < RB::N; ++k) qq[k] = __ldg(q + b * q_stride + k);
        CollideResult<RB> r;
        collide_state<RB, WANT_GRAD, WANT_SELFMIN>(qq, sv, base_min, r);
        write_collide<RB>(out, b, qq, r, thr_soft, thr_hard);
    }
}

// goal row (or, when the caller supplies u_ref, that row: cbf_filter_state then uses it verbatim)
template <int N>
__device__ __forceinline__ void load_goal(const NetArgs& net, long long b, float (&g)[N]) {
    const float* src = net.u_ref ? net.u_ref + b * N : net.goal + (net.goal_rows == 1 ? 0 : b * N);
#pragma unroll
    for (int k = 0; k < N; ++k) g[k] = __ldg(src + k);
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
    SceneArgs none{nullptr, nullptr, 0, 0, 0};
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

template <class RB, int H, int L>
__global__ void __launch_bounds__(NT) safety_kernel(SceneArgs sc, NetArgs net, const float* __restrict__ q,
                                                    long long q_stride, long long B, float thr_soft, float thr_hard,
                                                    CollideOutPtrs cout, FilterOutPtrs fout) {
    constexpr int N = RB::N;
    extern __shared__ __align__(16) float smem_raw[];
    using LY = NetLayout<N + 1, H, L>;
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, LY::TOTAL, ParamLayout<N>::TOTAL);
    stage<N>(s, sc, &net, LY::TOTAL);
    const SceneS sv = scene_view(s, sc);
    const float base_min = base_min_cta<RB>(s, sv);
    float* scr = s.scratch + threadIdx.x;
    for (long long b = (long long)blockIdx.x * NT + threadIdx.x; b < B; b += (long long)gridDim.x * NT) {
        float qq[N];
#pragma unroll
        for (int k = 0; k < N; ++k) qq[k] = __ldg(q + b * q_stride + k);
        CollideResult<RB> cr;
        collide_state<RB, true, false>(qq, sv, base_min, cr);
        write_collide<RB>(cout, b, qq, cr, thr_soft, thr_hard);
        float g[N], J[N], u[N], h, r;
        load_goal<N>(net, b, g);
        cbf_filter_state<N, H, L, NT>(s.weights, s.params, scr, qq, cr.o, cr.dodq, g, net.u_ref != nullptr, net.lam,
                                      net.penalty, h, J, u, r);
        write_filter<N>(fout, b, h, J, u, r);
    }
}


// ------------------------------------------------------------------ tensor-core (tcgen05) variants
// Same contracts as filter_kernel / safety_kernel; the MLP of every 128-state tile runs as TF32x3 GEMMs on the
// tensor cores (mlp_tc.cuh).  One CTA of 384 threads per SM = 3 tile groups sharing the staged weights; all 128
// threads of a group take part in every tile of that group (idle rows are masked), because the MMAs, the TMEM
// traffic and the barriers are group-wide.
constexpr int NTC = tc::NT_TC;

template <int N, int L>
__device__ __forceinline__ void filter_tile_tc(tc::Ctx<L>& c, const float* prm, const float (&q)[N], float o,
                                               const float (&dodq)[N], const float (&goal)[N], bool goal_is_uref,
                                               float lam, float penalty, float& h, float (&J)[N], float (&u)[N], float& r) {
    float x[N + 1], jac[N + 1];
#pragma unroll
    for (int k = 0; k < N; ++k) x[k] = q[k];
    x[N] = o;
    tc::mlp_h_jac<N + 1, L>(c, x, h, jac);
    cbf_filter_post<N>(prm, q, dodq, goal, goal_is_uref, lam, penalty, h, jac, J, u, r);
}

template <int N, int L>
__global__ void __launch_bounds__(NTC, 1) filter_kernel_tc(NetArgs net, const float* __restrict__ datax, long long B,
                                                           FilterOutPtrs out) {
    extern __shared__ __align__(16) float smem_raw[];
    const Smem s = carve(smem_raw, 0, 0, tc::Layout<L>::TOTAL, ParamLayout<N>::TOTAL);
    SceneArgs none{nullptr, nullptr, 0, 0, 0};
    stage<N, NTC>(s, none, &net, 0);
    tc::Ctx<L> c;
    tc::setup<N + 1, L>(c, s.weights, net.weights);
    const long long n_tiles = (B + tc::M_TILE - 1) / tc::M_TILE;
    for (long long tile = (long long)blockIdx.x * tc::G + c.g; tile < n_tiles; tile += (long long)gridDim.x * tc::G) {
        const long long b = tile * tc::M_TILE + c.t;
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
        filter_tile_tc<N, L>(c, s.params, q, o, dodq, g, net.u_ref != nullptr, net.lam, net.penalty, h, J, u, r);
        if (active) write_filter<N>(out, b, h, J, u, r);
    }
    tc::teardown<L>(c);
}

template <class RB, int L>
__global__ void __launch_bounds__(NTC, 1) safety_kernel_tc(SceneArgs sc, NetArgs net, const float* __restrict__ q,
                                                           long long q_stride, long long B, float thr_soft, float thr_hard,
                                                           CollideOutPtrs cout, FilterOutPtrs fout) {
    constexpr int N = RB::N;
    extern __shared__ __align__(16) float smem_raw[];
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, tc::Layout<L>::TOTAL, ParamLayout<N>::TOTAL);
    stage<N, NTC>(s, sc, &net, 0);
    const SceneS sv = scene_view(s, sc);
    const float base_min = base_min_cta<RB, NTC>(s, sv);
    tc::Ctx<L> c;
    tc::setup<N + 1, L>(c, s.weights, net.weights);
    const long long n_tiles = (B + tc::M_TILE - 1) / tc::M_TILE;
    for (long long tile = (long long)blockIdx.x * tc::G + c.g; tile < n_tiles; tile += (long long)gridDim.x * tc::G) {
        const long long b = tile * tc::M_TILE + c.t;
        const bool active = b < B;
        float qq[N], g[N], J[N], u[N], h, r;
#pragma unroll
        for (int k = 0; k < N; ++k) { qq[k] = 0.f; g[k] = 0.f; }
        CollideResult<RB> cr;
        cr.o = 0.f;
#pragma unroll
        for (int k = 0; k < N; ++k) cr.dodq[k] = 0.f;
        if (active) {
#pragma unroll
            for (int k = 0; k < N; ++k) qq[k] = __ldg(q + b * q_stride + k);
            collide_state<RB, true, false>(qq, sv, base_min, cr);
            write_collide<RB>(cout, b, qq, cr, thr_soft, thr_hard);
            load_goal<N>(net, b, g);
        }
        filter_tile_tc<N, L>(c, s.params, qq, cr.o, cr.dodq, g, net.u_ref != nullptr, net.lam, net.penalty, h, J, u, r);
        if (active) write_filter<N>(fout, b, h, J, u, r);
    }
    tc::teardown<L>(c);
}

// edge validity: one thread per (edge, point); point k in [0, K] (K = end point).  A CTA owns EPB whole edges
// (EPB * (K+1) <= NT threads active) and reduces "first failing index" with a shared-memory atomicMin.
template <class RB>
__global__ void __launch_bounds__(NT) edge_kernel(SceneArgs sc, const float* __restrict__ q_from,
                                                  const float* __restrict__ q_to, long long E, int K, int epb,
                                                  float thr_soft, float thr_hard, uint8_t* __restrict__ valid,
                                                  int* __restrict__ first_bad) {
    constexpr int N = RB::N;
    extern __shared__ __align__(16) float smem_raw[];
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, 0, 0);
    __shared__ int s_bad[NT];
    stage<N>(s, sc, nullptr, 0);
    const SceneS sv = scene_view(s, sc);
    const float base_min = base_min_cta<RB>(s, sv);
    const int pts = K + 1;
    const int le = threadIdx.x / pts, k = threadIdx.x - le * pts;
    const bool active = le < epb;
    const long long n_groups = (E + epb - 1) / epb;
    for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const long long e = grp * epb + le;
        if (threadIdx.x < epb) s_bad[threadIdx.x] = 0x7fffffff;
        __syncthreads();
        if (active && e < E) {
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
            collide_state<RB, false, false>(c, sv, base_min, r);
            const bool unsafe = r.hard_min <= thr_hard;
            const bool safe = r.self_ok && (r.soft_min > thr_soft) && !unsafe;
            if (!safe) atomicMin(&s_bad[le], k);
        }
        __syncthreads();
        if (threadIdx.x < epb) {
            const long long ee = grp * epb + threadIdx.x;
            if (ee < E) {
                const int bad = s_bad[threadIdx.x];
                valid[ee] = bad == 0x7fffffff;
                first_bad[ee] = bad == 0x7fffffff ? -1 : bad;
            }
        }
        __syncthreads();
    }
}

// long edges (K + 1 > NT): one warp-strided thread loop per edge is not needed by any reference caller; they
// are handled by the thread-per-edge fallback below (sequential over k with early exit, tsa.py:281-285).
template <class RB>
__global__ void __launch_bounds__(NT) edge_seq_kernel(SceneArgs sc, const float* __restrict__ q_from,
                                                      const float* __restrict__ q_to, long long E, int K,
                                                      float thr_soft, float thr_hard, uint8_t* __restrict__ valid,
                                                      int* __restrict__ first_bad) {
    constexpr int N = RB::N;
    extern __shared__ __align__(16) float smem_raw[];
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, 0, 0);
    stage<N>(s, sc, nullptr, 0);
    const SceneS sv = scene_view(s, sc);
    const float base_min = base_min_cta<RB>(s, sv);
    for (long long e = (long long)blockIdx.x * NT + threadIdx.x; e < E; e += (long long)gridDim.x * NT) {
        float a[N], d[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { a[i] = __ldg(q_from + e * N + i); d[i] = __fsub_rn(__ldg(q_to + e * N + i), a[i]); }
        int bad = -1;
        for (int k = 0; k <= K && bad < 0; ++k) {
            float c[N];
            if (k == K) {
#pragma unroll
                for (int i = 0; i < N; ++i) c[i] = __ldg(q_to + e * N + i);
            } else {
                const float t = __fdiv_rn((float)k, (float)K);
#pragma unroll
                for (int i = 0; i < N; ++i) c[i] = fmaf(t, d[i], a[i]);
            }
            CollideResult<RB> r;
            collide_state<RB, false, false>(c, sv, base_min, r);
            const bool unsafe = r.hard_min <= thr_hard;
            if (!(r.self_ok && (r.soft_min > thr_soft) && !unsafe)) bad = k;
        }
        valid[e] = bad < 0;
        first_bad[e] = bad;
    }
}

// closed-loop rollouts (the steer loop): one thread per trajectory, the whole loop in one launch.
struct RolloutArgs {
    int steps, ctrl_every;
    float dt;
    int stuck_lag, stuck_after;
    float stuck_eps;
    int keep_going;
};
template <class RB, int H, int L>
__global__ void __launch_bounds__(NT) rollout_kernel(SceneArgs sc, NetArgs net, const float* __restrict__ q0,
                                                     long long T, RolloutArgs ra, float thr_soft, float thr_hard,
                                                     float* __restrict__ q_final, uint8_t* __restrict__ alive,
                                                     int* __restrict__ steps_done, float* __restrict__ traj) {
    constexpr int N = RB::N;
    extern __shared__ __align__(16) float smem_raw[];
    using LY = NetLayout<N + 1, H, L>;
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, LY::TOTAL, ParamLayout<N>::TOTAL);
    stage<N>(s, sc, &net, LY::TOTAL);
    const SceneS sv = scene_view(s, sc);
    const float base_min = base_min_cta<RB>(s, sv);
    float* scr = s.scratch + threadIdx.x;
    for (long long b = (long long)blockIdx.x * NT + threadIdx.x; b < T; b += (long long)gridDim.x * NT) {
        float q[N], qn[N], g[N], u[N], dodq[N], o;
#pragma unroll
        for (int k = 0; k < N; ++k) { q[k] = __ldg(q0 + b * N + k); u[k] = 0.f; }
        load_goal<N>(net, b, g);
        CollideResult<RB> cr;
        collide_state<RB, true, false>(q, sv, base_min, cr);
        o = cr.o;
#pragma unroll
        for (int k = 0; k < N; ++k) dodq[k] = cr.dodq[k];
        bool ok = true;
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
            collide_state<RB, true, false>(qn, sv, base_min, cr);
            if ((t + 1) % ra.ctrl_every == 0) {
                o = cr.o;
#pragma unroll
                for (int k = 0; k < N; ++k) dodq[k] = cr.dodq[k];
            } else {
                // closed_loop_dynamics(update_observation=False) leaves o / do_dq at zero (arm_dynamics.py:491-493);
                // they are not read before the next refresh, the controller only runs at t % ctrl_every == 0.
            }
            if (cr.hard_min <= thr_hard) {  // tsa.py:251-253: colliding state is not appended
                ok = false;
                if (!ra.keep_going) break;
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
                if (__fsqrt_rn(dsq) < ra.stuck_eps) break;
            }
        }
#pragma unroll
        for (int k = 0; k < N; ++k) q_final[b * N + k] = q[k];
        alive[b] = ok;
        steps_done[b] = done;
    }
}

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
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_last_cuda{0};

static int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

template <class Kern>
static int grid_for(Kern kern, size_t smem, long long tiles) {
    int per_sm = 0;
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem) != cudaSuccess || per_sm < 1) return -1;
    long long g = (long long)num_sms() * per_sm;
    if (g > tiles) g = tiles;
    if (g < 1) g = 1;
    return (int)g;
}

// tensor-core kernels: one 384-thread CTA per SM (it owns all 512 TMEM columns), 3 tiles of 128 rows per pass
template <class Kern>
static int grid_for_tc(Kern kern, size_t smem, long long rows) {
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    }
    const long long tiles = (rows + tc::M_TILE - 1) / tc::M_TILE;
    long long g = (tiles + tc::G - 1) / tc::G;
    if (g > num_sms()) g = num_sms();
    return (int)(g < 1 ? 1 : g);
}

static int finish_launch() {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_cuda = (int)e;
        return CBFRRT_ERR_CUDA;
    }
    g_launches++;
    return CBFRRT_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int check_scene(const cbfrrt_scene* sc, SceneArgs& a) {
    if (!sc || sc->n_boxes < 0 || sc->n_spheres < 0) return CBFRRT_ERR_BAD_SHAPE;
    if (sc->n_boxes + sc->n_spheres > CBFRRT_MAX_OBSTACLES) return CBFRRT_ERR_UNSUPPORTED;
    if ((sc->n_boxes && !sc->boxes) || (sc->n_spheres && !sc->spheres)) return CBFRRT_ERR_BAD_SHAPE;
    if (!aligned16(sc->boxes) || !aligned16(sc->spheres)) return CBFRRT_ERR_ALIGNMENT;
    a = SceneArgs{sc->boxes, sc->spheres, sc->n_boxes, sc->n_spheres, sc->floor ? 1 : 0};
    return CBFRRT_OK;
}

static int check_net(int n, const cbfrrt_cbf_net* net, const cbfrrt_filter_params* fp, long long B, NetArgs& a) {
    if (!net || !fp || !net->weights || !fp->u_lo || !fp->u_hi) return CBFRRT_ERR_BAD_SHAPE;
    if (!fp->u_ref && (!fp->goal || !fp->K)) return CBFRRT_ERR_BAD_SHAPE;
    if (net->n_in != n + 1) return CBFRRT_ERR_BAD_SHAPE;
    if (net->hidden != 48 || net->layers != 2) return CBFRRT_ERR_UNSUPPORTED;
    if (!fp->u_ref && fp->goal_rows != 1 && fp->goal_rows != B) return CBFRRT_ERR_BAD_SHAPE;
    if (!aligned16(net->weights)) return CBFRRT_ERR_ALIGNMENT;
    a = NetArgs{net->weights, fp->K, fp->u_lo, fp->u_hi, fp->goal, (long long)fp->goal_rows, fp->cbf_lambda,
                fp->relaxation_penalty, fp->u_ref};
    return CBFRRT_OK;
}

template <class RB>
static int launch_fk(const float* q, long long qs, long long B, float* pos, float* rot, cudaStream_t st) {
    const long long tiles = (B + NT - 1) / NT;
    int grid = (int)(tiles < (long long)num_sms() * 8 ? tiles : (long long)num_sms() * 8);
    fk_kernel<RB><<<grid, NT, 0, st>>>(q, qs, B, pos, rot);
    return finish_launch();
}

template <class RB>
static int launch_collide(const SceneArgs& sc, const float* q, long long qs, long long B, float ts, float th,
                          const CollideOutPtrs& out, cudaStream_t st) {
    const size_t smem = 4 * smem_floats(sc.n_boxes, sc.n_spheres, 0, 0, 0);
    const long long tiles = (B + NT - 1) / NT;
    const bool grad = out.datax || out.arg_link || out.arg_obs;
    int grid;
    if (out.mins) {
        auto k = collide_kernel<RB, true, true>;
        if ((grid = grid_for(k, smem, tiles)) < 0) return CBFRRT_ERR_CUDA;
        k<<<grid, NT, smem, st>>>(sc, q, qs, B, ts, th, out);
    } else if (grad) {
        auto k = collide_kernel<RB, true, false>;
        if ((grid = grid_for(k, smem, tiles)) < 0) return CBFRRT_ERR_CUDA;
        k<<<grid, NT, smem, st>>>(sc, q, qs, B, ts, th, out);
    } else {
        auto k = collide_kernel<RB, false, false>;
        if ((grid = grid_for(k, smem, tiles)) < 0) return CBFRRT_ERR_CUDA;
        k<<<grid, NT, smem, st>>>(sc, q, qs, B, ts, th, out);
    }
    return finish_launch();
}

// CBFRRT_MLP=fma selects the FP32-FMA MLP kernels (A/B comparisons); default is the tcgen05 path
static bool use_tc() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("CBFRRT_MLP");
        v = (e && e[0] == 'f') ? 0 : 1;
    }
    return v == 1;
}

template <int N>
static int launch_filter(const NetArgs& net, const float* datax, long long B, const FilterOutPtrs& out, cudaStream_t st) {
    if (use_tc()) {
        const size_t smem = 4 * smem_floats(0, 0, tc::Layout<2>::TOTAL, ParamLayout<N>::TOTAL, 0);
        auto k = filter_kernel_tc<N, 2>;
        const int grid = grid_for_tc(k, smem, B);
        if (grid < 0) return CBFRRT_ERR_CUDA;
        k<<<grid, NTC, smem, st>>>(net, datax, B, out);
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

template <class RB>
static int launch_safety(const SceneArgs& sc, const NetArgs& net, const float* q, long long qs, long long B, float ts,
                         float th, const CollideOutPtrs& co, const FilterOutPtrs& fo, cudaStream_t st) {
    if (use_tc()) {
        const size_t smem = 4 * smem_floats(sc.n_boxes, sc.n_spheres, tc::Layout<2>::TOTAL, ParamLayout<RB::N>::TOTAL, 0);
        auto k = safety_kernel_tc<RB, 2>;
        const int grid = grid_for_tc(k, smem, B);
        if (grid < 0) return CBFRRT_ERR_CUDA;
        k<<<grid, NTC, smem, st>>>(sc, net, q, qs, B, ts, th, co, fo);
        return finish_launch();
    }
    using LY = NetLayout<RB::N + 1, 48, 2>;
    const size_t smem = 4 * smem_floats(sc.n_boxes, sc.n_spheres, LY::TOTAL, ParamLayout<RB::N>::TOTAL, 48 * NT);
    auto k = safety_kernel<RB, 48, 2>;
    const int grid = grid_for(k, smem, (B + NT - 1) / NT);
    if (grid < 0) return CBFRRT_ERR_CUDA;
    k<<<grid, NT, smem, st>>>(sc, net, q, qs, B, ts, th, co, fo);
    return finish_launch();
}

template <class RB>
static int launch_edges(const SceneArgs& sc, const float* qf, const float* qt, long long E, int K, float ts, float th,
                        uint8_t* valid, int* first_bad, cudaStream_t st) {
    const size_t smem = 4 * smem_floats(sc.n_boxes, sc.n_spheres, 0, 0, 0);
    if (K + 1 <= NT) {
        const int epb = NT / (K + 1);
        auto k = edge_kernel<RB>;
        const int grid = grid_for(k, smem, (E + epb - 1) / epb);
        if (grid < 0) return CBFRRT_ERR_CUDA;
        k<<<grid, NT, smem, st>>>(sc, qf, qt, E, K, epb, ts, th, valid, first_bad);
    } else {
        auto k = edge_seq_kernel<RB>;
        const int grid = grid_for(k, smem, (E + NT - 1) / NT);
        if (grid < 0) return CBFRRT_ERR_CUDA;
        k<<<grid, NT, smem, st>>>(sc, qf, qt, E, K, ts, th, valid, first_bad);
    }
    return finish_launch();
}

template <class RB>
static int launch_rollout(const SceneArgs& sc, const NetArgs& net, const float* q0, long long T, const RolloutArgs& ra,
                          float ts, float th, float* qf, uint8_t* alive, int* done, float* traj, cudaStream_t st) {
    using LY = NetLayout<RB::N + 1, 48, 2>;
    const size_t smem = 4 * smem_floats(sc.n_boxes, sc.n_spheres, LY::TOTAL, ParamLayout<RB::N>::TOTAL, 48 * NT);
    auto k = rollout_kernel<RB, 48, 2>;
    const int grid = grid_for(k, smem, (T + NT - 1) / NT);
    if (grid < 0) return CBFRRT_ERR_CUDA;
    k<<<grid, NT, smem, st>>>(sc, net, q0, T, ra, ts, th, qf, alive, done, traj);
    return finish_launch();
}

}  // namespace cbfrrt

using namespace cbfrrt;

extern "C" {

int cbfrrt_abi_version(void) { return CBFRRT_ABI_VERSION; }

const char* cbfrrt_status_string(int s) {
    switch (s) {
        case CBFRRT_OK: return "ok";
        case CBFRRT_ERR_BAD_ROBOT: return "unknown robot id";
        case CBFRRT_ERR_BAD_SHAPE: return "bad shape or missing pointer";
        case CBFRRT_ERR_UNSUPPORTED: return "unsupported network shape or scene size";
        case CBFRRT_ERR_CUDA: return "CUDA launch/runtime failure";
        case CBFRRT_ERR_ALIGNMENT: return "scene/weight buffer not 16-byte aligned";
        default: return "unknown status";
    }
}

int cbfrrt_last_cuda_error(void) { return g_last_cuda.load(); }
int64_t cbfrrt_launch_count(void) { return g_launches.load(); }

int cbfrrt_robot_info(int robot, int32_t* n, int32_t* n_links, int32_t* n_spheres) {
    if (robot == CBFRRT_ROBOT_PANDA) { *n = Panda::N; *n_links = Panda::NLINKS; *n_spheres = Panda::NSPH; return CBFRRT_OK; }
    if (robot == CBFRRT_ROBOT_MAGICIAN) { *n = Magician::N; *n_links = Magician::NLINKS; *n_spheres = Magician::NSPH; return CBFRRT_OK; }
    return CBFRRT_ERR_BAD_ROBOT;
}

int cbfrrt_fk(int robot, const float* q, int64_t q_stride, int64_t B, float* pos, float* rot, void* stream) {
    if (B < 0 || !q || !pos || !rot) return CBFRRT_ERR_BAD_SHAPE;
    if (B == 0) return CBFRRT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (robot == CBFRRT_ROBOT_PANDA) return q_stride < Panda::N ? CBFRRT_ERR_BAD_SHAPE : launch_fk<Panda>(q, q_stride, B, pos, rot, st);
    if (robot == CBFRRT_ROBOT_MAGICIAN) return q_stride < Magician::N ? CBFRRT_ERR_BAD_SHAPE : launch_fk<Magician>(q, q_stride, B, pos, rot, st);
    return CBFRRT_ERR_BAD_ROBOT;
}

int cbfrrt_collide(int robot, const cbfrrt_scene* scene, const float* q, int64_t q_stride, int64_t B, float thr_soft,
                   float thr_hard, uint8_t* safe, uint8_t* unsafe, float* datax, int32_t* arg_link, int32_t* arg_obs,
                   float* mins, void* stream) {
    if (robot != CBFRRT_ROBOT_PANDA && robot != CBFRRT_ROBOT_MAGICIAN) return CBFRRT_ERR_BAD_ROBOT;
    if (B < 0 || !q) return CBFRRT_ERR_BAD_SHAPE;
    SceneArgs sc;
    int rc = check_scene(scene, sc);
    if (rc) return rc;
    if (B == 0) return CBFRRT_OK;
    CollideOutPtrs out{safe, unsafe, datax, arg_link, arg_obs, mins};
    cudaStream_t st = (cudaStream_t)stream;
    if (robot == CBFRRT_ROBOT_PANDA) return q_stride < Panda::N ? CBFRRT_ERR_BAD_SHAPE : launch_collide<Panda>(sc, q, q_stride, B, thr_soft, thr_hard, out, st);
    return q_stride < Magician::N ? CBFRRT_ERR_BAD_SHAPE : launch_collide<Magician>(sc, q, q_stride, B, thr_soft, thr_hard, out, st);
}

int cbfrrt_cbf_filter(int n, const cbfrrt_cbf_net* net, const float* datax, int64_t B, const cbfrrt_filter_params* fp,
                      float* h, float* J, float* u, float* r, void* stream) {
    if (B < 0 || !datax) return CBFRRT_ERR_BAD_SHAPE;
    if (n != Panda::N && n != Magician::N) return CBFRRT_ERR_UNSUPPORTED;
    NetArgs na;
    int rc = check_net(n, net, fp, B, na);
    if (rc) return rc;
    if (B == 0) return CBFRRT_OK;
    FilterOutPtrs out{h, J, u, r};
    cudaStream_t st = (cudaStream_t)stream;
    return n == Panda::N ? launch_filter<Panda::N>(na, datax, B, out, st) : launch_filter<Magician::N>(na, datax, B, out, st);
}

int cbfrrt_qp(int n, const float* a, const float* c, const float* u_ref, int64_t B, const float* u_lo, const float* u_hi,
              float relaxation_penalty, float* u, float* r, void* stream) {
    if (B < 0 || !a || !c || !u_ref || !u_lo || !u_hi || !u) return CBFRRT_ERR_BAD_SHAPE;
    if (n < 1 || n > 8) return CBFRRT_ERR_UNSUPPORTED;
    if (B == 0) return CBFRRT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long blocks = (B + 255) / 256;
    const int grid = (int)(blocks < (long long)num_sms() * 8 ? blocks : (long long)num_sms() * 8);
    switch (n) {
#define CBF_QP_CASE(NN) case NN: qp_kernel<NN><<<grid, 256, 0, st>>>(a, c, u_ref, B, u_lo, u_hi, relaxation_penalty, u, r); break;
        CBF_QP_CASE(1) CBF_QP_CASE(2) CBF_QP_CASE(3) CBF_QP_CASE(4) CBF_QP_CASE(5) CBF_QP_CASE(6) CBF_QP_CASE(7) CBF_QP_CASE(8)
#undef CBF_QP_CASE
    }
    return finish_launch();
}

int cbfrrt_safety_filter(int robot, const cbfrrt_scene* scene, const cbfrrt_cbf_net* net, const float* q,
                         int64_t q_stride, int64_t B, const cbfrrt_filter_params* fp, float thr_soft, float thr_hard,
                         uint8_t* safe, uint8_t* unsafe, float* datax, float* h, float* J, float* u, float* r,
                         void* stream) {
    if (robot != CBFRRT_ROBOT_PANDA && robot != CBFRRT_ROBOT_MAGICIAN) return CBFRRT_ERR_BAD_ROBOT;
    if (B < 0 || !q) return CBFRRT_ERR_BAD_SHAPE;
    const int n = robot == CBFRRT_ROBOT_PANDA ? Panda::N : Magician::N;
    if (q_stride < n) return CBFRRT_ERR_BAD_SHAPE;
    SceneArgs sc;
    NetArgs na;
    int rc = check_scene(scene, sc);
    if (rc) return rc;
    rc = check_net(n, net, fp, B, na);
    if (rc) return rc;
    if (B == 0) return CBFRRT_OK;
    CollideOutPtrs co{safe, unsafe, datax, nullptr, nullptr, nullptr};
    FilterOutPtrs fo{h, J, u, r};
    cudaStream_t st = (cudaStream_t)stream;
    return robot == CBFRRT_ROBOT_PANDA ? launch_safety<Panda>(sc, na, q, q_stride, B, thr_soft, thr_hard, co, fo, st)
                                       : launch_safety<Magician>(sc, na, q, q_stride, B, thr_soft, thr_hard, co, fo, st);
}

int cbfrrt_edge_validity(int robot, const cbfrrt_scene* scene, const float* q_from, const float* q_to, int64_t E,
                         int32_t n_interp, float thr_soft, float thr_hard, uint8_t* valid, int32_t* first_bad,
                         void* stream) {
    if (robot != CBFRRT_ROBOT_PANDA && robot != CBFRRT_ROBOT_MAGICIAN) return CBFRRT_ERR_BAD_ROBOT;
    if (E < 0 || n_interp < 1 || !q_from || !q_to || !valid || !first_bad) return CBFRRT_ERR_BAD_SHAPE;
    SceneArgs sc;
    int rc = check_scene(scene, sc);
    if (rc) return rc;
    if (E == 0) return CBFRRT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return robot == CBFRRT_ROBOT_PANDA ? launch_edges<Panda>(sc, q_from, q_to, E, n_interp, thr_soft, thr_hard, valid, first_bad, st)
                                       : launch_edges<Magician>(sc, q_from, q_to, E, n_interp, thr_soft, thr_hard, valid, first_bad, st);
}

int cbfrrt_rollout(int robot, const cbfrrt_scene* scene, const cbfrrt_cbf_net* net, const float* q0, int64_t T,
                   const cbfrrt_filter_params* fp, const cbfrrt_rollout_params* rp, float thr_soft, float thr_hard,
                   float* q_final, uint8_t* alive, int32_t* steps_done, float* traj, void* stream) {
    if (robot != CBFRRT_ROBOT_PANDA && robot != CBFRRT_ROBOT_MAGICIAN) return CBFRRT_ERR_BAD_ROBOT;
    if (T < 0 || !rp || rp->steps < 0 || rp->ctrl_every < 1 || !q0 || !q_final || !alive || !steps_done)
        return CBFRRT_ERR_BAD_SHAPE;
    if (rp->stuck_lag < 0 || (rp->stuck_lag > 0 && (!traj || rp->stuck_after < rp->stuck_lag))) return CBFRRT_ERR_BAD_SHAPE;
    if (fp && fp->u_ref) return CBFRRT_ERR_UNSUPPORTED;  // the loop needs the nominal controller
    const int n = robot == CBFRRT_ROBOT_PANDA ? Panda::N : Magician::N;
    SceneArgs sc;
    NetArgs na;
    int rc = check_scene(scene, sc);
    if (rc) return rc;
    rc = check_net(n, net, fp, T, na);
    if (rc) return rc;
    if (T == 0) return CBFRRT_OK;
    const RolloutArgs ra{rp->steps, rp->ctrl_every, rp->dt, rp->stuck_lag, rp->stuck_after, rp->stuck_eps, rp->keep_going};
    cudaStream_t st = (cudaStream_t)stream;
    return robot == CBFRRT_ROBOT_PANDA
               ? launch_rollout<Panda>(sc, na, q0, T, ra, thr_soft, thr_hard, q_final, alive, steps_done, traj, st)
               : launch_rollout<Magician>(sc, na, q0, T, ra, thr_soft, thr_hard, q_final, alive, steps_done, traj, st);
}

int cbfrrt_fma_probe(int32_t grid, int32_t block, int32_t iters, float* sink, void* stream) {
    if (grid < 1 || block < 32 || block > 1024 || iters < 1 || !sink) return CBFRRT_ERR_BAD_SHAPE;
    fma_probe_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(iters, sink);
    return finish_launch();
}

int cbfrrt_sample_states(int robot, uint32_t seed, uint64_t start_row, int64_t B, float* q, void* stream) {
    if (B < 0 || !q) return CBFRRT_ERR_BAD_SHAPE;
    if (B == 0) return CBFRRT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long blocks = (B * 8 + 255) / 256;
    const int grid = (int)(blocks < (long long)num_sms() * 16 ? blocks : (long long)num_sms() * 16);
    if (robot == CBFRRT_ROBOT_PANDA) sample_kernel<Panda><<<grid, 256, 0, st>>>(seed, start_row, B, q);
    else if (robot == CBFRRT_ROBOT_MAGICIAN) sample_kernel<Magician><<<grid, 256, 0, st>>>(seed, start_row, B, q);
    else return CBFRRT_ERR_BAD_ROBOT;
    return finish_launch();
}

}  // extern "C"

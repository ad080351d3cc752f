// warp.cuh -- ONE STATE PER WARP: the planner-sized form of the hot path (sm_100a).
//
// The reference's planners call the path with a handful of states at a time (batch = 1 is the default of
// motion_planning/evaluation/eval_batchrrt_mindis_cbfsteer.py:66; tsa.py:217-269 steers one trajectory).  With one thread
// per state such a call is ONE dependent instruction chain of ~8 k instructions (26 us of device time for a single Panda
// state through the fused filter, 18 us per closed-loop step) while 147 SMs idle.  Here the 32 lanes of a warp share one state:
//   * the seven sin/cos pairs are evaluated by seven lanes at once, the (short) frame chain redundantly by every lane;
//   * lane i owns link sphere i: its centre, its exact obstacle term min_j e_j(C_i) - r_i, its floor candidate;
//   * minima by warp shuffles, the arg-min sphere as the lowest lane that attains the minimum, the arg-min obstacle by a
//     ballot over obstacle-parallel exact re-evaluations -- the oracle's scan order (sphere-major, floor candidate first,
//     strict '<') is "first position that attains the minimum", which is what lowest lane / lowest obstacle index gives;
//   * the self-collision pairs are spread over the lanes (centres fetched from their owners by shuffle);
//   * the barrier network runs lanes-over-neurons (forward: transposed weights, reverse: packed weights; both
//     conflict-free in shared memory).
// Every value that feeds a boolean or an index is produced by the same single IEEE operations as in core.cuh (the same
// device functions are called), and min is order-independent, so masks, minima and arg-min indices are bit-identical to
// the one-thread-per-state kernels and to the oracle.  All 32 lanes of a warp must call these functions converged; the
// results are uniform over the warp.
#pragma once
#include <cooperative_groups.h>

namespace cbfrrt {
namespace wp {
namespace cg = cooperative_groups;

constexpr int WNT = 256, WARPS = WNT / 32;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, d));
    return v;
}
// butterfly sum: every lane adds the same pairs in the same tree (addition is commutative), so the result is uniform
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = __fadd_rn(v, __shfl_xor_sync(FULL, v, d));
    return v;
}

// exact min_j e_j(c) over ALL obstacles of the staged scene (restructuring (1) of core.cuh, as sweep_scene's brute arm)
__device__ __forceinline__ float scene_min_brute(const SceneS& sc, float cx, float cy, float cz) {
    float D = CBF_INF, M = CBF_INF, E = CBF_INF;
    for (int j = 0; j < sc.n_boxes; ++j) acc_box(cx, cy, cz, sc.boxes[2 * j], sc.boxes[2 * j + 1], D, M);
    for (int j = 0; j < sc.n_spheres; ++j) acc_sphere(cx, cy, cz, sc.spheres[j], E);
    const float e_box = __fadd_rn(__fmul_rn(__fsqrt_rn(D), 0.5f), fminf(M, 0.f));   // D holds 4 * dsq
    return fminf(e_box, E);
}

template <class RB, bool WANT_GRAD, bool WANT_SELF, bool WANT_SELFMIN = false>
__device__ __forceinline__ void collide_state_warp(const float (&q)[RB::N], const SceneS& sc, float base_min,
                                                   CollideResult<RB>& res) {
    constexpr int N = RB::N, NS = RB::NSPH;
    static_assert(NS <= 32 && N <= 32, "one lane per link sphere");
    const int lane = threadIdx.x & 31;

    // joint angles -> sin / cos, one joint per lane
    float ql = 0.f, s_l, c_l;
#pragma unroll
    for (int k = 0; k < N; ++k) ql = (lane == k) ? q[k] : ql;
    det_sincosf(ql, s_l, c_l);

    const int sph = lane < NS ? lane : NS - 1;
    const int mylink = lane < NS ? RB::sph_link(sph) : -1;
    const bool act = mylink >= 0;                          // spheres of the static base link: base_min (per scene)
    float P[3] = {0.f, 0.f, 0.f}, R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    float jz[WANT_GRAD ? N : 1][3], jp[WANT_GRAD ? N : 1][3];
    Frame fr[RB::NLINKS];
    static_for<0, RB::NLINKS>([&](auto lc) {
        constexpr int L = decltype(lc)::value;
        constexpr int PL = RB::parent(L);
        float sq = 0.f, cq = 1.f;
        if constexpr (RB::qidx(L) >= 0) {
            sq = __shfl_sync(FULL, s_l, RB::qidx(L));
            cq = __shfl_sync(FULL, c_l, RB::qidx(L));
        }
        if constexpr (PL < 0) {
            const Frame base = identity_frame();
            child_frame<RB, L>(base, true, sq, cq, fr[L]);
        } else {
            child_frame<RB, L>(fr[PL], false, sq, cq, fr[L]);
        }
        if constexpr (WANT_GRAD && RB::qidx(L) >= 0) {
            constexpr int J = RB::qidx(L);
            jz[J][0] = fr[L].R[2]; jz[J][1] = fr[L].R[5]; jz[J][2] = fr[L].R[8];
            jp[J][0] = fr[L].p[0]; jp[J][1] = fr[L].p[1]; jp[J][2] = fr[L].p[2];
        }
        if constexpr (RB::sph_end(L) > RB::sph_begin(L)) {
            if (mylink == L) {
#pragma unroll
                for (int k = 0; k < 3; ++k) P[k] = fr[L].p[k];
#pragma unroll
                for (int k = 0; k < 9; ++k) R[k] = fr[L].R[k];
            }
        }
    });

    // this lane's sphere: centre, obstacle term, floor candidates (same exemptions as collide_state)
    const float c0 = RB::sph(sph, 0), c1 = RB::sph(sph, 1), c2 = RB::sph(sph, 2), rad = RB::sph(sph, 3);
    float C[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) C[r] = fmaf(R[3 * r + 2], c2, fmaf(R[3 * r + 1], c1, fmaf(R[3 * r + 0], c0, P[r])));
    float m = CBF_INF, f_o = CBF_INF, f_soft = CBF_INF, f_hard = CBF_INF;
    if (act) {
        unsigned n_pairs = 0u;
        const float e = sc.grid ? scene_min_exact(sc, C[0], C[1], C[2], n_pairs) : scene_min_brute(sc, C[0], C[1], C[2]);
        m = __fsub_rn(e, rad);
        if (sc.floor) {
            if (mylink == 0) {
                if (lane == RB::sph_begin(0)) { f_o = RB::LINK0_FLOOR; f_hard = RB::LINK0_FLOOR; }   // soft: floor-exempt
            } else {
                const float d = __fsub_rn(C[2], rad);
                f_o = d; f_soft = d;
                if (mylink != 1) f_hard = d;                                                          // arm_env.py:84
            }
        }
    }
    const float v_o = fminf(f_o, m);
    const float o = warp_min(v_o);
    res.soft_min = fminf(base_min, warp_min(fminf(f_soft, m)));
    res.hard_min = fminf(base_min, warp_min(fminf(f_hard, m)));
    res.n_refined = 0u; res.n_pairs = 0u;

    // self collision: the sphere pairs of the non-adjacent link pairs, 32 at a time
    bool self_ok = true;
    float self_min = CBF_INF;
    if constexpr (WANT_SELF || WANT_SELFMIN) {
#pragma unroll 1
        for (int p0 = 0; p0 < RB::NSELFSPH; p0 += 32) {
            const int p = min(p0 + lane, RB::NSELFSPH - 1);
            const int i = RB::self_i(p), j = RB::self_j(p);
            const float dx = __fsub_rn(__shfl_sync(FULL, C[0], i), __shfl_sync(FULL, C[0], j));
            const float dy = __fsub_rn(__shfl_sync(FULL, C[1], i), __shfl_sync(FULL, C[1], j));
            const float dz = __fsub_rn(__shfl_sync(FULL, C[2], i), __shfl_sync(FULL, C[2], j));
            const float dsq = fmaf(dz, dz, fmaf(dy, dy, __fmul_rn(dx, dx)));
            self_ok = self_ok && (dsq > RB::self_x(p));
            if constexpr (WANT_SELFMIN)
                self_min = fminf(self_min, __fsub_rn(__fsub_rn(__fsqrt_rn(dsq), RB::sph(i, 3)), RB::sph(j, 3)));
        }
        self_ok = __all_sync(FULL, self_ok);
        if constexpr (WANT_SELFMIN) self_min = warp_min(self_min);
    }
    res.self_ok = self_ok; res.self_min = self_min;

    // first position of the oracle's scan that attains the minimum: lowest sphere index
    const unsigned hit = __ballot_sync(FULL, act && v_o == o);
    const int arg_s = (o < CBF_INF && hit) ? __ffs(hit) - 1 : -1;
    res.o = o; res.arg_sphere = arg_s; res.arg_obs = -1;
#pragma unroll
    for (int k = 0; k < N; ++k) res.dodq[k] = 0.f;
    if constexpr (WANT_GRAD) {
        if (arg_s < 0) return;                             // uniform
        const float argC[3] = {__shfl_sync(FULL, C[0], arg_s), __shfl_sync(FULL, C[1], arg_s), __shfl_sync(FULL, C[2], arg_s)};
        const int link = RB::sph_link(arg_s);
        const float r = RB::sph(arg_s, 3);
        const int off = sc.floor ? 1 : 0;
        float nrm[3] = {0.f, 0.f, 1.f};
        int arg_o = -1;
        if (sc.floor) {
            const float d = (link == 0) ? RB::LINK0_FLOOR : __fsub_rn(argC[2], r);
            if (d == o) arg_o = 0;
        }
        // first obstacle (boxes, then spheres) whose exact pair distance equals o: 32 obstacles per ballot
        for (int j0 = 0; arg_o < 0 && j0 < sc.n_boxes; j0 += 32) {
            const int j = min(j0 + lane, sc.n_boxes - 1);
            const bool eq = (j0 + lane < sc.n_boxes) &&
                            sd_box_exact(argC[0], argC[1], argC[2], sc.boxes[2 * j], sc.boxes[2 * j + 1], r) == o;
            const unsigned b = __ballot_sync(FULL, eq);
            if (b) {
                const int jj = j0 + __ffs(b) - 1;
                arg_o = off + jj;
                normal_box(argC[0], argC[1], argC[2], sc.boxes[2 * jj], sc.boxes[2 * jj + 1], nrm);
            }
        }
        for (int j0 = 0; arg_o < 0 && j0 < sc.n_spheres; j0 += 32) {
            const int j = min(j0 + lane, sc.n_spheres - 1);
            const bool eq = (j0 + lane < sc.n_spheres) && sd_sphere_exact(argC[0], argC[1], argC[2], sc.spheres[j], r) == o;
            const unsigned b = __ballot_sync(FULL, eq);
            if (b) {
                const int jj = j0 + __ffs(b) - 1;
                arg_o = off + sc.n_boxes + jj;
                normal_sphere(argC[0], argC[1], argC[2], sc.spheres[jj], nrm);
            }
        }
        res.arg_obs = arg_o;
        const int mask = RB::anc_mask(link);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const float rx = __fsub_rn(argC[0], jp[j][0]), ry = __fsub_rn(argC[1], jp[j][1]), rz = __fsub_rn(argC[2], jp[j][2]);
            const float cx = fmaf(jz[j][1], rz, -__fmul_rn(jz[j][2], ry));
            const float cy = fmaf(jz[j][2], rx, -__fmul_rn(jz[j][0], rz));
            const float cz = fmaf(jz[j][0], ry, -__fmul_rn(jz[j][1], rx));
            const float g = fmaf(nrm[2], cz, fmaf(nrm[1], cy, __fmul_rn(nrm[0], cx)));
            res.dodq[j] = ((mask >> j) & 1) ? g : 0.f;
        }
    }
}

// ------------------------------------------------------------------ barrier network, lanes over neurons
// shared memory of a CTA: [transposed hidden weights 2 x H x (H + 1)] [per warp: 5 x H activations / gradients]
template <int N_IN, int H>
struct WarpNet {
    static constexpr int HP = H + 1;                       // odd row pitch: the transposing copy is conflict-free
    static constexpr int WT = 2 * H * HP;
    static constexpr int SCR = 5 * H;
    static constexpr int TOTAL = WT + WARPS * SCR;
};
template <int N_IN, int H, int NTH>
__device__ __forceinline__ void warp_net_setup(const float* __restrict__ w, float* __restrict__ wt) {
    using LY = NetLayout<N_IN, H, 2>;
    using WN = WarpNet<N_IN, H>;
    for (int l = 0; l < 2; ++l)
        for (int i = threadIdx.x; i < H * H; i += NTH) {
            const int j = i / H, k = i % H;
            wt[l * H * WN::HP + k * WN::HP + j] = w[LY::HID0 + l * LY::HID_STRIDE + i];
        }
    __syncthreads();
}
// h and dh/dx of the 2-hidden-layer ReLU network (neural_mindis_cbf_controller.py:41-102) for ONE input x shared by the
// warp.  FP32 FMAs; the summation order differs from the oracle's, results are tolerance-checked like the other MLP paths.
template <int N_IN, int H>
__device__ __forceinline__ void warp_mlp_h_jac(const float* __restrict__ w, const float* __restrict__ wt, float* __restrict__ ws,
                                               const float (&x)[N_IN], float& h_out, float (&jac)[N_IN]) {
    using LY = NetLayout<N_IN, H, 2>;
    using WN = WarpNet<N_IN, H>;
    static_assert(H > 32 && H <= 64 && H % 2 == 0, "two neurons per lane");
    static_assert(32 % LY::IN_PAD == 0, "input columns x partial sums must tile the warp");
    const int lane = threadIdx.x & 31;
    const int j0 = lane, j1 = lane + 32;
    const bool two = j1 < H;
    const int j1c = two ? j1 : j0;
    float* A0 = ws;
    float* A1 = ws + H;
    float* A2 = ws + 2 * H;
    float* G0 = ws + 3 * H;
    float* G1 = ws + 4 * H;
    __syncwarp();                                          // the previous call's readers are done with ws
    {
        float s0 = w[LY::B_IN + j0], s1 = w[LY::B_IN + j1c];
        const float4* r0 = reinterpret_cast<const float4*>(w + LY::W_IN + j0 * LY::IN_PAD);
        const float4* r1 = reinterpret_cast<const float4*>(w + LY::W_IN + j1c * LY::IN_PAD);
#pragma unroll
        for (int k4 = 0; k4 < LY::IN_PAD / 4; ++k4) {
            const float4 a = r0[k4], b = r1[k4];
            const float wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (4 * k4 + e < N_IN) {
                    s0 = fmaf(wa[e], x[4 * k4 + e < N_IN ? 4 * k4 + e : 0], s0);
                    s1 = fmaf(wb[e], x[4 * k4 + e < N_IN ? 4 * k4 + e : 0], s1);
                }
        }
        A0[j0] = fmaxf(s0, 0.f);
        if (two) A0[j1] = fmaxf(s1, 0.f);
    }
    __syncwarp();
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        const float* in = l ? A1 : A0;
        float* out = l ? A2 : A1;
        const float* wtl = wt + l * H * WN::HP;
        const float* bias = w + LY::HID0 + l * LY::HID_STRIDE + H * H;
        float s00 = bias[j0], s01 = 0.f, s10 = bias[j1c], s11 = 0.f;
#pragma unroll 8
        for (int k = 0; k < H; k += 2) {
            const float a0 = in[k], a1 = in[k + 1];
            s00 = fmaf(wtl[k * WN::HP + j0], a0, s00);
            s01 = fmaf(wtl[(k + 1) * WN::HP + j0], a1, s01);
            s10 = fmaf(wtl[k * WN::HP + j1c], a0, s10);
            s11 = fmaf(wtl[(k + 1) * WN::HP + j1c], a1, s11);
        }
        out[j0] = fmaxf(s00 + s01, 0.f);
        if (two) out[j1] = fmaxf(s10 + s11, 0.f);
        __syncwarp();
    }
    {
        float p = __fmul_rn(w[LY::W_OUT + j0], A2[j0]);
        if (two) p = fmaf(w[LY::W_OUT + j1], A2[j1], p);
        h_out = __fadd_rn(w[LY::B_OUT], warp_sum(p));
        G0[j0] = A2[j0] > 0.f ? w[LY::W_OUT + j0] : 0.f;
        if (two) G0[j1] = A2[j1] > 0.f ? w[LY::W_OUT + j1] : 0.f;
    }
    __syncwarp();
#pragma unroll
    for (int l = 1; l >= 0; --l) {                         // reverse mode: g_prev[k] = relu'(z_prev[k]) sum_j W[j][k] g[j]
        const float* gin = l ? G0 : G1;
        float* gout = l ? G1 : G0;
        const float* mask = l ? A1 : A0;
        const float* W = w + LY::HID0 + l * LY::HID_STRIDE;
        float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
#pragma unroll 8
        for (int j = 0; j < H; j += 2) {
            const float g0 = gin[j], g1 = gin[j + 1];
            s00 = fmaf(W[j * H + j0], g0, s00);
            s01 = fmaf(W[(j + 1) * H + j0], g1, s01);
            s10 = fmaf(W[j * H + j1c], g0, s10);
            s11 = fmaf(W[(j + 1) * H + j1c], g1, s11);
        }
        gout[j0] = mask[j0] > 0.f ? s00 + s01 : 0.f;
        if (two) gout[j1] = mask[j1] > 0.f ? s10 + s11 : 0.f;
        __syncwarp();
    }
    {   // input layer, reverse: jac[k] = sum_j W_in[j][k] g[j]; lane = (partial sum, column), j interleaved over the parts
        constexpr int PARTS = 32 / LY::IN_PAD;
        const int k = lane % LY::IN_PAD, part = lane / LY::IN_PAD;
        float acc = 0.f;
        for (int j = part; j < H; j += PARTS) acc = fmaf(w[LY::W_IN + j * LY::IN_PAD + k], G0[j], acc);
#pragma unroll
        for (int d = LY::IN_PAD; d < 32; d <<= 1) acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, d));
#pragma unroll
        for (int kk = 0; kk < N_IN; ++kk) jac[kk] = __shfl_sync(FULL, acc, kk);
    }
}

// q row of one state, the same for every lane of the warp (one broadcast transaction per element)
template <int N>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&v)[N]) {
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = __ldg(p + k);
}

// ------------------------------------------------------------------ kernels: same contracts as their one-thread-per-state
// namesakes in kernels.cuh (collide_kernel, safety_kernel, edge_var_kernel, rollout_kernel, rollout_segments_kernel)
template <class RB, bool WANT_GRAD, bool WANT_SELFMIN>
__global__ void __launch_bounds__(WNT) collide_kernel_warp(SceneArgs sc, const float* __restrict__ q, long long q_stride,
                                                           long long B, float thr_soft, float thr_hard, CollideOutPtrs out) {
    extern __shared__ __align__(16) float smem_raw[];
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, 0, 0);
    stage<RB::N, WNT>(s, sc, nullptr, 0);
    const SceneS sv = scene_view(s, sc, thr_soft, thr_hard);
    const float base_min = base_min_cta<RB, WNT>(s, sv);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long b = (long long)blockIdx.x * WARPS + warp; b < B; b += (long long)gridDim.x * WARPS) {
        float qq[RB::N];
        load_row<RB::N>(q + b * q_stride, qq);
        CollideResult<RB> r;
        collide_state_warp<RB, WANT_GRAD, true, WANT_SELFMIN>(qq, sv, base_min, r);
        if (lane == 0) write_collide<RB>(out, b, qq, r, thr_soft, thr_hard);
    }
}

template <class RB, int H>
__global__ void __launch_bounds__(WNT) safety_kernel_warp(SceneArgs sc, NetArgs net, const float* __restrict__ q,
                                                          long long q_stride, long long B, float thr_soft, float thr_hard,
                                                          CollideOutPtrs cout, FilterOutPtrs fout) {
    constexpr int N = RB::N;
    extern __shared__ __align__(16) float smem_raw[];
    using LY = NetLayout<N + 1, H, 2>;
    using WN = WarpNet<N + 1, H>;
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, LY::TOTAL, ParamLayout<N>::TOTAL);
    stage<N, WNT>(s, sc, &net, LY::TOTAL);
    const SceneS sv = scene_view(s, sc, thr_soft, thr_hard);
    const float base_min = base_min_cta<RB, WNT>(s, sv);
    warp_net_setup<N + 1, H, WNT>(s.weights, s.scratch);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* ws = s.scratch + WN::WT + warp * WN::SCR;
    for (long long b = (long long)blockIdx.x * WARPS + warp; b < B; b += (long long)gridDim.x * WARPS) {
        float qq[N], g[N], J[N], u[N], x[N + 1], jac[N + 1], h, r;
        load_row<N>(q + b * q_stride, qq);
        CollideResult<RB> cr;
        collide_state_warp<RB, true, true>(qq, sv, base_min, cr);
        if (lane == 0) write_collide<RB>(cout, b, qq, cr, thr_soft, thr_hard);
        load_goal<N>(net, b, g);
#pragma unroll
        for (int k = 0; k < N; ++k) x[k] = qq[k];
        x[N] = cr.o;
        warp_mlp_h_jac<N + 1, H>(s.weights, s.scratch, ws, x, h, jac);
        cbf_filter_post<N>(s.params, qq, cr.dodq, g, net.u_ref != nullptr, net.lam, net.penalty, h, jac, J, u, r);
        if (lane == 0) write_filter<N>(fout, b, h, J, u, r);
    }
}

template <int N, int H>
__global__ void __launch_bounds__(WNT) filter_kernel_warp(NetArgs net, const float* __restrict__ datax, long long B,
                                                          FilterOutPtrs out) {
    extern __shared__ __align__(16) float smem_raw[];
    using LY = NetLayout<N + 1, H, 2>;
    using WN = WarpNet<N + 1, H>;
    const Smem s = carve(smem_raw, 0, 0, LY::TOTAL, ParamLayout<N>::TOTAL);
    SceneArgs none{nullptr, nullptr, 0, 0, 0, nullptr, nullptr};
    stage<N, WNT>(s, none, &net, LY::TOTAL);
    warp_net_setup<N + 1, H, WNT>(s.weights, s.scratch);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* ws = s.scratch + WN::WT + warp * WN::SCR;
    for (long long b = (long long)blockIdx.x * WARPS + warp; b < B; b += (long long)gridDim.x * WARPS) {
        const float* d = datax + b * (2 * N + 1);
        float q[N], dodq[N], g[N], J[N], u[N], x[N + 1], jac[N + 1], h, r;
#pragma unroll
        for (int k = 0; k < N; ++k) { q[k] = __ldg(d + k); dodq[k] = __ldg(d + N + 1 + k); x[k] = q[k]; }
        x[N] = __ldg(d + N);
        load_goal<N>(net, b, g);
        warp_mlp_h_jac<N + 1, H>(s.weights, s.scratch, ws, x, h, jac);
        cbf_filter_post<N>(s.params, q, dodq, g, net.u_ref != nullptr, net.lam, net.penalty, h, jac, J, u, r);
        if (lane == 0) write_filter<N>(out, b, h, J, u, r);
    }
}

template <class RB>
__global__ void __launch_bounds__(WNT) edge_var_kernel_warp(SceneArgs sc, const float* __restrict__ q_from,
                                                            const float* __restrict__ q_to, long long E,
                                                            const long long* __restrict__ off, int K_all, long long P,
                                                            float thr_soft, float thr_hard, int* __restrict__ first_bad) {
    constexpr int N = RB::N;
    extern __shared__ __align__(16) float smem_raw[];
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, 0, 0);
    stage<N, WNT>(s, sc, nullptr, 0);
    const SceneS sv = scene_view(s, sc, thr_soft, thr_hard);
    const float base_min = base_min_cta<RB, WNT>(s, sv);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long p = (long long)blockIdx.x * WARPS + warp; p < P; p += (long long)gridDim.x * WARPS) {
        long long e;
        int K, k;
        if (off == nullptr) {
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
            load_row<N>(q_to + e * N, c);
        } else {
            const float t = __fdiv_rn((float)k, (float)K);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const float a = __ldg(q_from + e * N + i), bb = __ldg(q_to + e * N + i);
                c[i] = fmaf(t, __fsub_rn(bb, a), a);
            }
        }
        CollideResult<RB> r;
        collide_state_warp<RB, false, true>(c, sv, base_min, r);
        const bool unsafe = r.hard_min <= thr_hard;
        const bool safe = r.self_ok && (r.soft_min > thr_soft) && !unsafe;
        if (!safe && lane == 0) atomicMin(first_bad + e, k);
    }
}

// one closed-loop trajectory per warp (tsa.py:217-269 / batch_tsa.py:191-231 through rollout_after_step); the trajectory
// rows are written (and re-read by the stuck test) by all lanes alike -- same address, same value
template <class RB, int H>
__global__ void __launch_bounds__(WNT) rollout_kernel_warp(SceneArgs sc, NetArgs net, const float* __restrict__ q0,
                                                           long long T, RolloutArgs ra, float thr_soft, float thr_hard,
                                                           float* __restrict__ q_final, uint8_t* __restrict__ alive,
                                                           int* __restrict__ steps_done, float* __restrict__ traj) {
    constexpr int N = RB::N;
    extern __shared__ __align__(16) float smem_raw[];
    using LY = NetLayout<N + 1, H, 2>;
    using WN = WarpNet<N + 1, H>;
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, LY::TOTAL, ParamLayout<N>::TOTAL);
    stage<N, WNT>(s, sc, &net, LY::TOTAL);
    const SceneS sv = scene_view(s, sc, thr_soft, thr_hard);
    const float base_min = base_min_cta<RB, WNT>(s, sv);
    warp_net_setup<N + 1, H, WNT>(s.weights, s.scratch);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* ws = s.scratch + WN::WT + warp * WN::SCR;
    float* ring = s.scratch + WN::TOTAL + warp * STUCK_RING * N;   // last stored states (stuck test)
    const bool use_ring = ra.stuck_lag < STUCK_RING;
    for (long long b = (long long)blockIdx.x * WARPS + warp; b < T; b += (long long)gridDim.x * WARPS) {
        float q[N], qn[N], g[N], u[N], dodq[N], o;
        load_row<N>(q0 + b * N, q);
#pragma unroll
        for (int k = 0; k < N; ++k) u[k] = 0.f;
        load_goal<N>(net, b, g);
        CollideResult<RB> cr;
        collide_state_warp<RB, true, false>(q, sv, base_min, cr);
        o = cr.o;
#pragma unroll
        for (int k = 0; k < N; ++k) dodq[k] = cr.dodq[k];
        bool ok = true, dead = false;
        int done = 0;
        float* tr = traj ? traj + b * (long long)ra.steps * N : nullptr;
        for (int t = 0; t < ra.steps; ++t) {
            if (t % ra.ctrl_every == 0) {
                float x[N + 1], jac[N + 1], J[N], h, r;
#pragma unroll
                for (int k = 0; k < N; ++k) x[k] = q[k];
                x[N] = o;
                warp_mlp_h_jac<N + 1, H>(s.weights, s.scratch, ws, x, h, jac);
                cbf_filter_post<N>(s.params, q, dodq, g, false, net.lam, net.penalty, h, jac, J, u, r);
            }
#pragma unroll
            for (int k = 0; k < N; ++k) qn[k] = fmaf(u[k], ra.dt, q[k]);
            if ((t + 1) % ra.ctrl_every == 0) {
                collide_state_warp<RB, true, false>(qn, sv, base_min, cr);
                o = cr.o;
#pragma unroll
                for (int k = 0; k < N; ++k) dodq[k] = cr.dodq[k];
            } else {
                collide_state_warp<RB, false, false>(qn, sv, base_min, cr);
            }
            __syncwarp();                                  // stored states of earlier steps are visible to every lane
            const bool fin = use_ring ? rollout_after_step_ring<N>(ra, t, cr.hard_min <= thr_hard, q, qn, tr, ring, ok, dead, done)
                                      : rollout_after_step<N>(ra, t, cr.hard_min <= thr_hard, q, qn, tr, ok, dead, done);
            if (fin) break;
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < N; ++k) q_final[b * N + k] = q[k];
            alive[b] = ok;
            steps_done[b] = done;
        }
    }
}

// A whole expansion of the batched planner (see rollout_segments_kernel): one row per warp, the rows of the batch spread
// over a thread-block CLUSTER of up to 8 CTAs (64 warps; up to two rows per warp for T <= 128).
//
// The reference's batch-wide early exit (batch_tsa.py:229-230: the segment's loop ends after the first step at which EVERY
// row is dead, and x_last is the running state of every row after THAT step) couples the rows.  A barrier per step would
// put a cluster round trip (~1.5 us) on a 2 us step, so the rows run DECOUPLED within a segment and the exit step is
// reconstructed exactly:
//   * a row that dies at step d publishes d into every CTA's shared memory (distributed shared memory stores) -- a row dies
//     at most once per segment, so the entry is final once written;
//   * after each of its steps a warp polls its CTA's copy: when every row has a death step, t* = max d is THE exit step; the
//     warp keeps stepping while t < t* and stops at t >= t*.  Polling a stale copy only delays the exit;
//   * a dead row saves its running state after every step (it can run past t* before it learns t*); after the ONE cluster
//     barrier at the end of the segment every CTA computes the same exact t* and rows that ran past it restore the state
//     after step t*.  Appended states, counts and flags stop changing at a row's death, so nothing else needs a roll-back.
// Two death vectors alternate between segments; the next segment's vector is reset before the barrier of this one.
constexpr int SEG_SLOTS = 2, SEG_MAX_CLUSTER = 8, SEG_MAX_ROWS = SEG_SLOTS * SEG_MAX_CLUSTER * WARPS, SEG_ALIVE = 0x7fffffff;
template <int N>
struct RowState {   // floats per (warp, slot)
    static constexpr int Q = 0, U = N, G = 2 * N, DODQ = 3 * N, O = 4 * N, TOTAL = 4 * N + 4;
};
// floats of shared memory behind the staged scene / weights / parameters
template <class RB, int H>
__host__ __device__ constexpr int seg_scratch_floats(int seg_len) {
    return WarpNet<RB::N + 1, H>::TOTAL + WARPS * SEG_SLOTS * (RowState<RB::N>::TOTAL + STUCK_RING * RB::N + seg_len * RB::N) +
           2 * SEG_MAX_ROWS;
}
template <class RB, int H>
__global__ void __launch_bounds__(WNT) rollout_segments_kernel_warp(SceneArgs sc, NetArgs net, const float* __restrict__ q0,
                                                                    int T, int n_seg, int seg_len, RolloutArgs ra,
                                                                    float thr_soft, float thr_hard, float* __restrict__ seg_q,
                                                                    uint8_t* __restrict__ seg_alive, int* __restrict__ seg_done,
                                                                    float* __restrict__ traj) {
    constexpr int N = RB::N;
    using RS = RowState<N>;
    extern __shared__ __align__(16) float smem_raw[];
    using LY = NetLayout<N + 1, H, 2>;
    using WN = WarpNet<N + 1, H>;
    cg::cluster_group cluster = cg::this_cluster();
    const int n_cta = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const Smem s = carve(smem_raw, sc.n_boxes, sc.n_spheres, LY::TOTAL, ParamLayout<N>::TOTAL);
    stage<N, WNT>(s, sc, &net, LY::TOTAL);
    const SceneS sv = scene_view(s, sc, thr_soft, thr_hard);
    const float base_min = base_min_cta<RB, WNT>(s, sv);
    warp_net_setup<N + 1, H, WNT>(s.weights, s.scratch);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* ws = s.scratch + WN::WT + warp * WN::SCR;
    float* p_rows = s.scratch + WN::TOTAL;
    float* p_rings = p_rows + WARPS * SEG_SLOTS * RS::TOTAL;
    float* p_hist = p_rings + WARPS * SEG_SLOTS * STUCK_RING * N;
    int* death = reinterpret_cast<int*>(p_hist + WARPS * SEG_SLOTS * seg_len * N);   // [2][SEG_MAX_ROWS]
    float* rows = p_rows + warp * SEG_SLOTS * RS::TOTAL;
    float* rings = p_rings + warp * SEG_SLOTS * STUCK_RING * N;
    float* hist = p_hist + warp * SEG_SLOTS * seg_len * N;
    const bool use_ring = ra.stuck_lag < STUCK_RING;
    const int total_warps = n_cta * WARPS, gw = rank * WARPS + warp;
    const long long steps = (long long)n_seg * seg_len;
    bool ok[SEG_SLOTS], dead[SEG_SLOTS];
    int done[SEG_SLOTS];
    for (int i = threadIdx.x; i < 2 * SEG_MAX_ROWS; i += WNT) death[i] = SEG_ALIVE;
#pragma unroll
    for (int sl = 0; sl < SEG_SLOTS; ++sl) {
        const int b = gw + sl * total_warps;
        float* st = rows + sl * RS::TOTAL;
        if (b < T && lane == 0) {
#pragma unroll
            for (int k = 0; k < N; ++k) { st[RS::Q + k] = __ldg(q0 + (long long)b * N + k); st[RS::U + k] = 0.f; }
            float g[N];
            load_goal<N>(net, b, g);
#pragma unroll
            for (int k = 0; k < N; ++k) st[RS::G + k] = g[k];
        }
    }
    __syncwarp();
    cluster.sync();                                        // every CTA runs and has initialised its death vectors
    for (int sg = 0; sg < n_seg; ++sg) {
        int* dv = death + (sg & 1) * SEG_MAX_ROWS;
#pragma unroll
        for (int sl = 0; sl < SEG_SLOTS; ++sl) {           // complete_sample_with_observations at the start of the segment
            const int b = gw + sl * total_warps;
            ok[sl] = true; dead[sl] = !(b < T); done[sl] = 0;
            if (b < T) {
                float* st = rows + sl * RS::TOTAL;
                float q[N];
#pragma unroll
                for (int k = 0; k < N; ++k) q[k] = st[RS::Q + k];
                CollideResult<RB> cr;
                collide_state_warp<RB, true, false>(q, sv, base_min, cr);
                __syncwarp();
                if (lane == 0) {
                    st[RS::O] = cr.o;
#pragma unroll
                    for (int k = 0; k < N; ++k) st[RS::DODQ + k] = cr.dodq[k];
                }
                __syncwarp();
            }
        }
        int last = -1;                                     // last step this warp executed
        for (int t = 0; t < seg_len; ++t) {
#pragma unroll
            for (int sl = 0; sl < SEG_SLOTS; ++sl) {
                const int b = gw + sl * total_warps;
                if (b < T) {                               // uniform over the warp
                    float* st = rows + sl * RS::TOTAL;
                    float q[N], qn[N], u[N];
#pragma unroll
                    for (int k = 0; k < N; ++k) { q[k] = st[RS::Q + k]; u[k] = st[RS::U + k]; }
                    if (t % ra.ctrl_every == 0) {
                        float x[N + 1], jac[N + 1], J[N], g[N], dodq[N], h, r;
#pragma unroll
                        for (int k = 0; k < N; ++k) { x[k] = q[k]; g[k] = st[RS::G + k]; dodq[k] = st[RS::DODQ + k]; }
                        x[N] = st[RS::O];
                        warp_mlp_h_jac<N + 1, H>(s.weights, s.scratch, ws, x, h, jac);
                        cbf_filter_post<N>(s.params, q, dodq, g, false, net.lam, net.penalty, h, jac, J, u, r);
                    }
#pragma unroll
                    for (int k = 0; k < N; ++k) qn[k] = fmaf(u[k], ra.dt, q[k]);
                    CollideResult<RB> cr;
                    const bool refresh = (t + 1) % ra.ctrl_every == 0;
                    if (refresh) collide_state_warp<RB, true, false>(qn, sv, base_min, cr);
                    else collide_state_warp<RB, false, false>(qn, sv, base_min, cr);
                    float* tr = traj + ((long long)b * steps + (long long)sg * seg_len) * N;
                    const bool was_dead = dead[sl];
                    __syncwarp();                          // earlier stored states / row state: visible, reads done
                    if (use_ring)                          // mode 1, keep_going
                        rollout_after_step_ring<N>(ra, t, cr.hard_min <= thr_hard, q, qn, tr, rings + sl * STUCK_RING * N, ok[sl], dead[sl], done[sl]);
                    else
                        rollout_after_step<N>(ra, t, cr.hard_min <= thr_hard, q, qn, tr, ok[sl], dead[sl], done[sl]);
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < N; ++k) { st[RS::Q + k] = q[k]; st[RS::U + k] = u[k]; }
                        if (refresh) {
                            st[RS::O] = cr.o;
#pragma unroll
                            for (int k = 0; k < N; ++k) st[RS::DODQ + k] = cr.dodq[k];
                        }
                        if (dead[sl]) {                    // running state after step t of a dead row (roll-back target)
#pragma unroll
                            for (int k = 0; k < N; ++k) hist[(sl * seg_len + t) * N + k] = q[k];
                        }
                    }
                    if (dead[sl] && !was_dead && lane < n_cta) *cluster.map_shared_rank(dv + b, lane) = t;   // died at step t
                    __syncwarp();
                }
            }
            last = t;
            // every row has a death step?  then t* = the largest is the reference's exit step (batch_tsa.py:229-230)
            int mx = -1;
            bool fin = true;
            for (int r = lane; r < T; r += 32) {
                const int d = *reinterpret_cast<volatile int*>(dv + r);
                fin = fin && d != SEG_ALIVE;
                mx = max(mx, d);
            }
            fin = __all_sync(FULL, fin);
            mx = __reduce_max_sync(FULL, mx);
            if (fin && t >= mx) break;
        }
        for (int i = threadIdx.x; i < SEG_MAX_ROWS; i += WNT) death[((sg + 1) & 1) * SEG_MAX_ROWS + i] = SEG_ALIVE;
        cluster.sync();                                    // every death step of this segment is published
        int t_exit = -1;
        bool all_dead = true;
        for (int r = lane; r < T; r += 32) {
            const int d = dv[r];
            all_dead = all_dead && d != SEG_ALIVE;
            t_exit = max(t_exit, d);
        }
        all_dead = __all_sync(FULL, all_dead);
        t_exit = __reduce_max_sync(FULL, t_exit);
#pragma unroll
        for (int sl = 0; sl < SEG_SLOTS; ++sl) {
            const int b = gw + sl * total_warps;
            if (b < T && lane == 0) {
                float* st = rows + sl * RS::TOTAL;
                if (all_dead && t_exit < last) {           // ran past the exit step: the state after step t_exit
#pragma unroll
                    for (int k = 0; k < N; ++k) st[RS::Q + k] = hist[(sl * seg_len + t_exit) * N + k];
                }
#pragma unroll
                for (int k = 0; k < N; ++k) seg_q[((long long)b * n_seg + sg) * N + k] = st[RS::Q + k];
                seg_alive[(long long)b * n_seg + sg] = ok[sl];
                seg_done[(long long)b * n_seg + sg] = done[sl];
            }
        }
        __syncwarp();
    }
}

}  // namespace wp
}  // namespace cbfrrt

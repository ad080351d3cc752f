// mlp_qp.cuh -- barrier MLP (value + exact reverse-mode Jacobian), nominal control and the closed-form
// single-constraint CBF-QP, one state per thread (sm_100a).
//
// Reference semantics:
//   h = W3 relu(W2 relu(W1 relu(W0 [q,o] + b0) + b1) + b2) + b3      neural_mindis_cbf_controller.py:41-53,71-81
//   Jh = dh/d[q,o] (autograd == analytic backward with the ReLU masks) :83-98
//   J  = Jh_q + Jh_o * do/dq                                           :99-101
//   Lf_V = J f = 0, Lg_V = J g = J  (f = 0, g = I)                     clf_controller.py:169-218, arm_dynamics.py:327-363
//   u_ref = clamp(-K (q - goal), lo, hi)                               arm_dynamics.py:124-161
//   QP: min ||u-u_ref||^2 + p r  s.t. Lf + Lg.u + lam h - r <= 0, r >= 0, lo <= u <= hi   clf_controller.py:82-139
//
// Layout: the packed weights live in shared memory (staged once per CTA by a TMA bulk copy); every lane of a
// warp reads the same weight at the same time, so each LDS.128 is a single broadcast wavefront feeding 4 FFMAs.
// The activation vector of the thread lives in registers for the dot products of a layer and is parked in a
// per-thread shared-memory column (scratch[k * NT + tid], conflict-free) between layers, which keeps the layer
// loops rolled (small code, ~80 registers) instead of a 200 KB fully unrolled body.
// These values are tolerance-checked (1e-5), not bit-exact, so plain fmaf ordering is free here.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cbfrrt {

template <int N_IN, int H, int L>
struct NetLayout {
    static constexpr int IN_PAD = (N_IN + 3) & ~3;
    static constexpr int W_IN = 0;
    static constexpr int B_IN = H * IN_PAD;
    static constexpr int HID0 = B_IN + H;              // first hidden block: [W HxH][b H]
    static constexpr int HID_STRIDE = H * H + H;
    static constexpr int W_OUT = HID0 + L * HID_STRIDE;
    static constexpr int B_OUT = W_OUT + H;
    static constexpr int TOTAL = B_OUT + 4;            // b_out padded to 4 floats
};

// dot(W_row[0..H), a[0..H)) + bias with 4 partial sums (breaks the 48-long dependent FFMA chain)
template <int H>
__device__ __forceinline__ float dot_row(const float* __restrict__ wrow, const float (&a)[H], float bias) {
    float s0 = bias, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const float4* w4 = reinterpret_cast<const float4*>(wrow);
#pragma unroll
    for (int k = 0; k < H / 4; ++k) {
        const float4 w = w4[k];
        s0 = fmaf(w.x, a[4 * k + 0], s0);
        s1 = fmaf(w.y, a[4 * k + 1], s1);
        s2 = fmaf(w.z, a[4 * k + 2], s2);
        s3 = fmaf(w.w, a[4 * k + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}

// value h and Jacobian jac = dh/dx for x = [q, o]  (N_IN = n + 1).
//   w: packed weights in shared memory;  scr: this thread's scratch column (stride NT floats between entries)
template <int N_IN, int H, int L, int NT>
__device__ __forceinline__ void mlp_h_jac(const float* __restrict__ w, float* __restrict__ scr, const float (&x)[N_IN],
                                          float& h_out, float (&jac)[N_IN]) {
    using LY = NetLayout<N_IN, H, L>;
    static_assert(H % 4 == 0 && H <= 64, "hidden size must be a multiple of 4, at most 64 (ReLU masks are 64-bit)");
    float a[H];
    uint64_t mask[L + 1];
    // ---- input layer
    {
        uint64_t m = 0;
#pragma unroll 4
        for (int j = 0; j < H; ++j) {
            const float4* w4 = reinterpret_cast<const float4*>(w + LY::W_IN + j * LY::IN_PAD);
            float s = w[LY::B_IN + j];
#pragma unroll
            for (int k4 = 0; k4 < LY::IN_PAD / 4; ++k4) {
                const float4 ww = w4[k4];
                if (4 * k4 + 0 < N_IN) s = fmaf(ww.x, x[4 * k4 + 0 < N_IN ? 4 * k4 + 0 : 0], s);
                if (4 * k4 + 1 < N_IN) s = fmaf(ww.y, x[4 * k4 + 1 < N_IN ? 4 * k4 + 1 : 0], s);
                if (4 * k4 + 2 < N_IN) s = fmaf(ww.z, x[4 * k4 + 2 < N_IN ? 4 * k4 + 2 : 0], s);
                if (4 * k4 + 3 < N_IN) s = fmaf(ww.w, x[4 * k4 + 3 < N_IN ? 4 * k4 + 3 : 0], s);
            }
            const bool on = s > 0.f;
            m |= (uint64_t)on << j;
            scr[j * NT] = on ? s : 0.f;
        }
        mask[0] = m;
    }
    // ---- hidden layers
#pragma unroll
    for (int l = 0; l < L; ++l) {
#pragma unroll
        for (int k = 0; k < H; ++k) a[k] = scr[k * NT];
        const float* W = w + LY::HID0 + l * LY::HID_STRIDE;
        const float* bia = W + H * H;
        uint64_t m = 0;
#pragma unroll 2
        for (int j = 0; j < H; ++j) {
            const float s = dot_row<H>(W + j * H, a, bia[j]);
            const bool on = s > 0.f;
            m |= (uint64_t)on << j;
            scr[j * NT] = on ? s : 0.f;
        }
        mask[l + 1] = m;
    }
    // ---- output layer + seed of the backward pass
#pragma unroll
    for (int k = 0; k < H; ++k) a[k] = scr[k * NT];
    h_out = dot_row<H>(w + LY::W_OUT, a, w[LY::B_OUT]);
    {
        const float4* w4 = reinterpret_cast<const float4*>(w + LY::W_OUT);
        const uint64_t m = mask[L];
#pragma unroll
        for (int k = 0; k < H / 4; ++k) {
            const float4 ww = w4[k];
            scr[(4 * k + 0) * NT] = ((m >> (4 * k + 0)) & 1) ? ww.x : 0.f;
            scr[(4 * k + 1) * NT] = ((m >> (4 * k + 1)) & 1) ? ww.y : 0.f;
            scr[(4 * k + 2) * NT] = ((m >> (4 * k + 2)) & 1) ? ww.z : 0.f;
            scr[(4 * k + 3) * NT] = ((m >> (4 * k + 3)) & 1) ? ww.w : 0.f;
        }
    }
    // ---- backward through the hidden layers: g_prev[k] = relu'_prev[k] * sum_j W[j][k] g[j]
#pragma unroll
    for (int l = L - 1; l >= 0; --l) {
        const float* W = w + LY::HID0 + l * LY::HID_STRIDE;
#pragma unroll
        for (int k = 0; k < H; ++k) a[k] = 0.f;
#pragma unroll 2
        for (int j = 0; j < H; ++j) {
            const float g = scr[j * NT];
            const float4* w4 = reinterpret_cast<const float4*>(W + j * H);
#pragma unroll
            for (int k = 0; k < H / 4; ++k) {
                const float4 ww = w4[k];
                a[4 * k + 0] = fmaf(ww.x, g, a[4 * k + 0]);
                a[4 * k + 1] = fmaf(ww.y, g, a[4 * k + 1]);
                a[4 * k + 2] = fmaf(ww.z, g, a[4 * k + 2]);
                a[4 * k + 3] = fmaf(ww.w, g, a[4 * k + 3]);
            }
        }
        const uint64_t m = mask[l];
#pragma unroll
        for (int k = 0; k < H; ++k) scr[k * NT] = ((m >> k) & 1) ? a[k] : 0.f;
    }
    // ---- input layer backward: jac[k] = sum_j W_in[j][k] g[j]
#pragma unroll
    for (int k = 0; k < N_IN; ++k) jac[k] = 0.f;
#pragma unroll 4
    for (int j = 0; j < H; ++j) {
        const float g = scr[j * NT];
        const float4* w4 = reinterpret_cast<const float4*>(w + LY::W_IN + j * LY::IN_PAD);
#pragma unroll
        for (int k4 = 0; k4 < LY::IN_PAD / 4; ++k4) {
            const float4 ww = w4[k4];
            if (4 * k4 + 0 < N_IN) jac[4 * k4 + 0 < N_IN ? 4 * k4 + 0 : 0] = fmaf(ww.x, g, jac[4 * k4 + 0 < N_IN ? 4 * k4 + 0 : 0]);
            if (4 * k4 + 1 < N_IN) jac[4 * k4 + 1 < N_IN ? 4 * k4 + 1 : 0] = fmaf(ww.y, g, jac[4 * k4 + 1 < N_IN ? 4 * k4 + 1 : 0]);
            if (4 * k4 + 2 < N_IN) jac[4 * k4 + 2 < N_IN ? 4 * k4 + 2 : 0] = fmaf(ww.z, g, jac[4 * k4 + 2 < N_IN ? 4 * k4 + 2 : 0]);
            if (4 * k4 + 3 < N_IN) jac[4 * k4 + 3 < N_IN ? 4 * k4 + 3 : 0] = fmaf(ww.w, g, jac[4 * k4 + 3 < N_IN ? 4 * k4 + 3 : 0]);
        }
    }
}

// filter parameters staged in shared memory next to the weights: [K n*n][u_lo n][u_hi n]
template <int N>
struct ParamLayout {
    static constexpr int K = 0, LO = N * N, HI = N * N + N, TOTAL = ((N * N + 2 * N) + 3) & ~3;
};

template <int N, typename T>
__host__ __device__ __forceinline__ T qp_g(const T (&a)[N], const T (&uref)[N], const T* __restrict__ lo,
                                      const T* __restrict__ hi, T c, T mu) {
    T g = c;
    const T hm = T(0.5) * mu;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const T v = fmin(fmax(fma(-hm, a[i], uref[i]), lo[i]), hi[i]);
        g = fma(a[i], v, g);
    }
    return g;
}

// Root of the dual constraint of ONE relaxed half-space constraint  c + a.u - r <= 0  (penalty p on r) for the
// box-constrained projection  u(mu) = clip(t - mu a / 2, lo, hi):  g(mu) = c + a.u(mu) is non-increasing and
// piecewise linear with at most 2N breakpoints (a coordinate can leave one bound and reach the other when the
// target t lies outside the box).  Returns mu* = 0 if g(0) <= 0, p if g(p) > 0 (constraint relaxed), else the root,
// found by evaluating g at every breakpoint (no sort: 2N * N is small) and interpolating inside the bracketing
// segment (exact: g is linear there).
template <int N, typename T>
__host__ __device__ __forceinline__ T qp_mu(const T (&a)[N], const T (&t)[N], const T* __restrict__ lo,
                                       const T* __restrict__ hi, T c, T p) {
    const T g0 = qp_g<N>(a, t, lo, hi, c, T(0.));
    if (!(g0 > T(0.))) return T(0.);
    T mu_lo = T(0.), g_lo = g0;
    T mu_hi = p, g_hi = qp_g<N>(a, t, lo, hi, c, p);
    if (g_hi > T(0.)) return p;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (a[i] != T(0.)) {
            const T inv = T(2.) / a[i];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const T m = (t[i] - (s ? hi[i] : lo[i])) * inv;
                if (m > mu_lo && m < mu_hi) {
                    const T gm = qp_g<N>(a, t, lo, hi, c, m);
                    if (gm > T(0.)) { mu_lo = m; g_lo = gm; }
                    else          { mu_hi = m; g_hi = gm; }
                }
            }
        }
    }
    return (g_lo > g_hi) ? fma(mu_hi - mu_lo, g_lo / (g_lo - g_hi), mu_lo) : mu_hi;
}

// Exact KKT solution of the single-constraint CBF-QP
//   min ||u-u_ref||^2 + p r   s.t.  c + a.u - r <= 0, r >= 0, lo <= u <= hi      (clf_controller.py:82-139)
template <int N>
__host__ __device__ __forceinline__ void qp_solve(const float (&a)[N], const float (&uref)[N], const float* __restrict__ lo,
                                         const float* __restrict__ hi, float c, float p, float (&u)[N], float& r) {
    const float mu = qp_mu<N>(a, uref, lo, hi, c, p);
    const float hm = 0.5f * mu;
    float g = c;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        u[i] = fminf(fmaxf(fmaf(-hm, a[i], uref[i]), lo[i]), hi[i]);
        g = fmaf(a[i], u[i], g);
    }
    r = (mu >= p) ? fmaxf(g, 0.f) : 0.f;
}

// qp_mu in double at (nearly) float cost: the breakpoint search runs in float and only proposes an active set; the
// multiplier is then recomputed in double from that set (mu = 2 (c + a_F.t_F + a_C.b_C) / |a_F|^2) and accepted only if
// it reproduces the set and lies in (0, p) -- by monotonicity of g that is THE root.  Anything else (a float
// misclassification next to a kink, |a_F| = 0) falls back to the full double search.
template <int N>
__host__ __device__ __forceinline__ double qp_mu_refined(const double (&a)[N], const double (&t)[N],
                                                     const double* __restrict__ lo, const double* __restrict__ hi,
                                                     double c, double p) {
    if (!(qp_g<N>(a, t, lo, hi, c, 0.) > 0.)) return 0.;
    if (qp_g<N>(a, t, lo, hi, c, p) > 0.) return p;
    float af[N], tf[N], lf[N], hf[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { af[i] = (float)a[i]; tf[i] = (float)t[i]; lf[i] = (float)lo[i]; hf[i] = (float)hi[i]; }
    double mu = (double)qp_mu<N, float>(af, tf, lf, hf, (float)c, (float)p);
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
        double Nn = c, S = 0.;
        unsigned set = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double v = fma(-0.5 * mu, a[i], t[i]);
            if (v > lo[i] && v < hi[i]) { set |= 1u << i; Nn = fma(a[i], t[i], Nn); S = fma(a[i], a[i], S); }
            else Nn = fma(a[i], v <= lo[i] ? lo[i] : hi[i], Nn);
        }
        if (!(S > 0.)) break;
        const double m2 = 2. * Nn / S;
        unsigned set2 = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double v = fma(-0.5 * m2, a[i], t[i]);
            if (v > lo[i] && v < hi[i]) set2 |= 1u << i;
            else if ((v <= lo[i]) != (fma(-0.5 * mu, a[i], t[i]) <= lo[i])) set2 |= 1u << 31;   // other bound: reject
        }
        if (set2 == set && m2 > 0. && m2 < p) return m2;
        mu = m2;
        if (!(m2 > 0. && m2 < p)) break;
    }
    return qp_mu<N, double>(a, t, lo, hi, c, p);
}
template <int N>
__host__ __device__ __forceinline__ float qp_mu_refined(const float (&a)[N], const float (&t)[N], const float* __restrict__ lo,
                                                    const float* __restrict__ hi, float c, float p) {
    return qp_mu<N, float>(a, t, lo, hi, c, p);
}

// Reverse mode of the single-constraint CBF-QP: what backpropagation through the reference's CvxpyLayer returns
// (clf_controller.py:82-139, params Lf_V / Lg_V / V / u_ref; used with requires_grad=True, :461-528), from the
// closed-form sensitivities of the KKT solution instead of diffcp's differentiated cone program.
// With F = {i : lo_i < u_ref_i - mu a_i / 2 < hi_i} (unclipped coordinates), the solution is one of three pieces:
//   mu = 0          (constraint inactive)   u_F = u_ref_F, r = 0
//   0 < mu < p      (constraint tight)      mu = 2 (c + a_F.u_ref_F + a_C.b_C) / |a_F|^2, u_F = u_ref_F - mu a_F / 2, r = 0
//   mu = p          (constraint relaxed)    u_F = u_ref_F - p a_F / 2, r = c + a.u
// and each piece is differentiated in place (gu = dL/du, gr = dL/dr  ->  ga = dL/da, gc = dL/dc, gt = dL/du_ref).
// At a kink (measure zero) the clipped / inactive side is taken.  T = double: the sensitivities divide by |a_F|^2.
template <int N, typename T>
__host__ __device__ __forceinline__ void qp_backward(const T (&a)[N], const T (&uref)[N], const T* __restrict__ lo,
                                                const T* __restrict__ hi, T c, T p, const T (&gu)[N], T gr,
                                                T (&ga)[N], T& gc, T (&gt)[N]) {
    const T mu = qp_mu_refined<N>(a, uref, lo, hi, c, p);
    const T hm = T(0.5) * mu;
    bool fr[N];
    T u[N], S = T(0.), G = T(0.), g = c;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const T v = fma(-hm, a[i], uref[i]);
        fr[i] = v > lo[i] && v < hi[i];
        u[i] = fmin(fmax(v, lo[i]), hi[i]);
        g = fma(a[i], u[i], g);
        if (fr[i]) { S = fma(a[i], a[i], S); G = fma(gu[i], a[i], G); }
    }
    if (!(mu > T(0.))) {                        // inactive
        gc = T(0.);
#pragma unroll
        for (int i = 0; i < N; ++i) { ga[i] = T(0.); gt[i] = fr[i] ? gu[i] : T(0.); }
    } else if (mu >= p) {                       // relaxed: r = g(p) (r = 0 and no r-gradient if g(p) <= 0)
        const T grr = g > T(0.) ? gr : T(0.);
        gc = grr;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const T gi = fr[i] ? fma(grr, a[i], gu[i]) : T(0.);
            gt[i] = gi;
            ga[i] = fma(grr, u[i], -hm * gi);
        }
    } else {                                    // tight: mu is a function of (a, c, u_ref)
        const T k = S > T(0.) ? G / S : T(0.);  // = -dL/dmu * 2 / |a_F|^2
        gc = -k;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            gt[i] = fr[i] ? fma(-k, a[i], gu[i]) : T(0.);
            ga[i] = fr[i] ? fma(-k, u[i] - hm * a[i], -hm * gu[i]) : -k * u[i];
        }
    }
}

// S relaxed constraints (multi-scenario CLF/CBF-QP, clf_controller.py:86-139 with len(scenarios) > 1):
//   min ||u-u_ref||^2 + p sum_i r_i   s.t.  c_i + a_i.u - r_i <= 0, r_i >= 0, lo <= u <= hi.
// Batched dual active-set solver.  The dual is concave and piecewise quadratic in mu in [0,p]^S with
// u(mu) = clip(u_ref - sum_i mu_i a_i / 2).  One iteration =
//   (1) a cyclic sweep of exact coordinate maximisations -- each the single-constraint root find qp_mu on the target
//       shifted by the other multipliers;
//   (2) up to S + 2 Newton steps on the current active set (free multipliers W, unclipped coordinates F):
//       (A_WF A_WF^T / 2 + eps I) d = g_W, followed by an EXACT line search: along d the directional derivative is
//       d.c + w.clip(t - alpha w / 2) with w = A^T d, i.e. qp_mu again with (a, c, p) := (w, d.c, alpha_max).
// Both moves increase the dual monotonically; (1) guarantees convergence, (2) terminates finitely once the active set
// is identified (2-5 iterations on random problems, including infeasible systems relaxed at p = 1e6).  Stops when the
// control-space movement of an iteration is <= tol + (rounding floor of t = 32 eps |dual shift|).  One constraint is
// one sweep.
// S is the compile-time capacity, n_con <= S the constraints present (the rest zero-filled).  Returns iterations.
// T = double in the shipped kernel: with saturated multipliers (mu = p up to 1e6) t = u_ref - A^T mu / 2 cancels
// ~6 digits, which fp32 cannot afford at 1e-5; B200's FP64 pipe makes this tiny kernel cost nothing.
template <int N, int S, typename T>
__host__ __device__ __forceinline__ int qp_solve_multi(const T (&a)[S][N], const T (&uref)[N], const T* __restrict__ lo,
                                              const T* __restrict__ hi, const T (&c)[S], int n_con, T p,
                                              int max_iter, T tol, T (&u)[N], T (&r)[S]) {
    const T KINK = sizeof(T) == 8 ? T(1e-9) : T(1e-6), REG = sizeof(T) == 8 ? T(1e-9) : T(1e-6);       // coordinates sitting on a kink of the dual count as free in the Newton matrix
    T mu[S], t[N], amx[S];
#pragma unroll
    for (int i = 0; i < S; ++i) {
        mu[i] = T(0.);
        amx[i] = T(0.);
#pragma unroll
        for (int k = 0; k < N; ++k) amx[i] = fmax(amx[i], fabs(a[i][k]));
    }
#pragma unroll
    for (int k = 0; k < N; ++k) t[k] = uref[k];
    int it = 0;
    while (it < max_iter) {
        ++it;
        T moved = T(0.), shift = T(0.);
#pragma unroll
        for (int i = 0; i < S; ++i) {
            if (i < n_con) {
                T ti[N];
#pragma unroll
                for (int k = 0; k < N; ++k) ti[k] = fma(T(0.5) * mu[i], a[i][k], t[k]);   // remove constraint i's shift
                const T m = qp_mu<N>(a[i], ti, lo, hi, c[i], p);
                moved = fmax(moved, fabs(m - mu[i]) * T(0.5) * amx[i]);
                shift = fmax(shift, T(0.5) * m * amx[i]);
                mu[i] = m;
#pragma unroll
                for (int k = 0; k < N; ++k) t[k] = fma(-T(0.5) * m, a[i][k], ti[k]);
            }
        }
        const T stop = fma(sizeof(T) == 8 ? T(4e-15) : T(2e-6), shift, tol);   // tol + rounding floor of t
        if (moved <= stop) break;
        if (n_con < 2) continue;
        for (int rep = 0; rep < S + 2; ++rep) {
            T g[S], d[S], M[S][S];
            bool inW[S], isF[N];
            int nW = 0;
#pragma unroll
            for (int k = 0; k < N; ++k) isF[k] = t[k] >= lo[k] - KINK && t[k] <= hi[k] + KINK;
#pragma unroll
            for (int i = 0; i < S; ++i) {
                g[i] = c[i];
#pragma unroll
                for (int k = 0; k < N; ++k) g[i] = fma(a[i][k], fmin(fmax(t[k], lo[k]), hi[k]), g[i]);
                inW[i] = (mu[i] > T(0.) && mu[i] < p) || (mu[i] <= T(0.) && g[i] > T(0.)) || (mu[i] >= p && g[i] < T(0.));
                nW += inW[i];
            }
            if (nW < 2) break;
            T tr = T(0.);
#pragma unroll
            for (int i = 0; i < S; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    T acc = T(0.);
#pragma unroll
                    for (int k = 0; k < N; ++k) acc = fma(isF[k] ? a[i][k] : T(0.), a[j][k], acc);
                    const T v = (inW[i] && inW[j]) ? T(0.5) * acc : (i == j ? T(1.) : T(0.));
                    M[i][j] = v;
                    M[j][i] = v;
                    if (i == j && inW[i]) tr += v;
                }
#pragma unroll
            for (int i = 0; i < S; ++i) {
                if (inW[i]) M[i][i] += fma(REG, tr, T(1e-30));
                d[i] = inW[i] ? g[i] : T(0.);
            }
            bool ok = true;
#pragma unroll
            for (int i = 0; i < S; ++i) {            // Gaussian elimination (SPD, no pivoting)
                ok = ok && M[i][i] > T(0.);
                const T inv = T(1.) / M[i][i];
#pragma unroll
                for (int j = i + 1; j < S; ++j) {
                    const T f = M[j][i] * inv;
#pragma unroll
                    for (int k = i + 1; k < S; ++k) M[j][k] = fma(-f, M[i][k], M[j][k]);
                    d[j] = fma(-f, d[i], d[j]);
                }
            }
            if (!ok) break;
#pragma unroll
            for (int i = S - 1; i >= 0; --i) {
#pragma unroll
                for (int j = i + 1; j < S; ++j) d[i] = fma(-M[i][j], d[j], d[i]);
                d[i] = d[i] / M[i][i];
            }
            T amax = T(3e38), cc = T(0.);
            bool any = false;
#pragma unroll
            for (int i = 0; i < S; ++i) {
                if ((mu[i] <= T(0.) && d[i] < T(0.)) || (mu[i] >= p && d[i] > T(0.)) || !(fabs(d[i]) < T(1e30))) d[i] = T(0.);
                if (d[i] > T(0.)) { amax = fmin(amax, (p - mu[i]) / d[i]); any = true; }
                if (d[i] < T(0.)) { amax = fmin(amax, -mu[i] / d[i]); any = true; }
                cc = fma(d[i], c[i], cc);
            }
            if (!any) break;
            T w[N], wmax = T(0.);
#pragma unroll
            for (int k = 0; k < N; ++k) {
                w[k] = T(0.);
#pragma unroll
                for (int i = 0; i < S; ++i) w[k] = fma(a[i][k], d[i], w[k]);
                wmax = fmax(wmax, fabs(w[k]));
            }
            const T al = qp_mu<N>(w, t, lo, hi, cc, amax);
#pragma unroll
            for (int i = 0; i < S; ++i) mu[i] = fmin(fmax(fma(al, d[i], mu[i]), T(0.)), p);
#pragma unroll
            for (int k = 0; k < N; ++k) t[k] = fma(-T(0.5) * al, w[k], t[k]);
            // a step cut short by a multiplier bound changes the active set: go on, however small it was in u
            if (al < amax && !(al * wmax > stop)) break;
        }
    }
#pragma unroll
    for (int k = 0; k < N; ++k) u[k] = fmin(fmax(t[k], lo[k]), hi[k]);
#pragma unroll
    for (int i = 0; i < S; ++i) {
        T g = c[i];
#pragma unroll
        for (int k = 0; k < N; ++k) g = fma(a[i][k], u[k], g);
        r[i] = fmax(g, T(0.));       // optimal relaxation for this u
    }
    return it;
}

// Reverse mode of the S-constraint QP at its solution u (from qp_solve_multi).  Active sets are read off the solution:
// F = unclipped coordinates; per constraint g_i = c_i + a_i.u:  relaxed P (g_i > 0, mu_i = p, r_i = g_i), tight W
// (g_i = 0, 0 < mu_i < p), inactive Z (g_i < 0).  On that set
//     u_F = t_F - (A_WF^T mu_W + p sum_P a_i,F) / 2,        A_W u + c_W = 0,
// and with M = A_WF A_WF^T, cotangent w_F = gu_F + sum_P gr_i a_i,F:
//     lambda = -M^-1 A_WF w_F,  v_F = w_F + A_WF^T lambda,  mu_W = M^-1 A_WF (2 (t_F - u_F) - p sum_P a_i,F)
//     dL/dt_F = v_F;   dL/dc_W = lambda, dL/dc_P = gr_P;   dL/da_i = lambda_i u - mu_i v_F / 2  (W),  gr_i u - p v_F / 2  (P).
// S = 1 reduces to qp_backward.  M is regularised like the Newton matrix of the solver (duplicate tight constraints).
template <int N, int S, typename T>
__host__ __device__ __forceinline__ void qp_multi_adjoint(const T (&a)[S][N], const T (&uref)[N], const T* __restrict__ lo,
                                                     const T* __restrict__ hi, const T (&c)[S], int n_con, T p,
                                                     const T (&u)[N], const T (&gu)[N], const T (&gr)[S],
                                                     T (&ga)[S][N], T (&gc)[S], T (&gt)[N]) {
    bool fr[N];
    int kind[S];                                 // 0 inactive, 1 tight, 2 relaxed
    T w[N], sp[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { fr[k] = u[k] > lo[k] && u[k] < hi[k]; w[k] = fr[k] ? gu[k] : T(0.); sp[k] = T(0.); }
#pragma unroll
    for (int i = 0; i < S; ++i) {
        T g = c[i], sc = T(1.) + fabs(c[i]);
#pragma unroll
        for (int k = 0; k < N; ++k) { g = fma(a[i][k], u[k], g); sc += fabs(a[i][k]); }
        sc *= T(1e-6);
        kind[i] = i < n_con ? (g > sc ? 2 : (g >= -sc ? 1 : 0)) : 0;
        if (kind[i] == 2) {
#pragma unroll
            for (int k = 0; k < N; ++k)
                if (fr[k]) { w[k] = fma(gr[i], a[i][k], w[k]); sp[k] += a[i][k]; }
        }
    }
    T M[S][S], lam[S], mu[S], tr = T(0.);
#pragma unroll
    for (int i = 0; i < S; ++i) {
        T r1 = T(0.), r2 = T(0.);
#pragma unroll
        for (int k = 0; k < N; ++k)
            if (fr[k]) { r1 = fma(-a[i][k], w[k], r1); r2 = fma(a[i][k], T(2.) * (uref[k] - u[k]) - p * sp[k], r2); }
        lam[i] = kind[i] == 1 ? r1 : T(0.);
        mu[i] = kind[i] == 1 ? r2 : T(0.);
#pragma unroll
        for (int j = 0; j < S; ++j) {
            T acc = T(0.);
#pragma unroll
            for (int k = 0; k < N; ++k)
                if (fr[k]) acc = fma(a[i][k], a[j][k], acc);
            M[i][j] = (kind[i] == 1 && kind[j] == 1) ? acc : (i == j ? T(1.) : T(0.));
        }
        if (kind[i] == 1) tr += M[i][i];
    }
#pragma unroll
    for (int i = 0; i < S; ++i)
        if (kind[i] == 1) M[i][i] += T(1e-12) * tr + T(1e-300);
#pragma unroll
    for (int i = 0; i < S; ++i) {                // Gaussian elimination (SPD), two right-hand sides
        const T inv = T(1.) / M[i][i];
#pragma unroll
        for (int j = i + 1; j < S; ++j) {
            const T f = M[j][i] * inv;
#pragma unroll
            for (int k = i; k < S; ++k) M[j][k] = fma(-f, M[i][k], M[j][k]);
            lam[j] = fma(-f, lam[i], lam[j]);
            mu[j] = fma(-f, mu[i], mu[j]);
        }
    }
#pragma unroll
    for (int i = S - 1; i >= 0; --i) {
#pragma unroll
        for (int j = i + 1; j < S; ++j) { lam[i] = fma(-M[i][j], lam[j], lam[i]); mu[i] = fma(-M[i][j], mu[j], mu[i]); }
        lam[i] /= M[i][i];
        mu[i] /= M[i][i];
    }
    T v[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        v[k] = w[k];
#pragma unroll
        for (int i = 0; i < S; ++i)
            if (kind[i] == 1 && fr[k]) v[k] = fma(a[i][k], lam[i], v[k]);
        gt[k] = v[k];
    }
#pragma unroll
    for (int i = 0; i < S; ++i) {
        const T head = kind[i] == 1 ? lam[i] : (kind[i] == 2 ? gr[i] : T(0.));
        const T half = kind[i] == 1 ? T(0.5) * mu[i] : (kind[i] == 2 ? T(0.5) * p : T(0.));
        gc[i] = head;
#pragma unroll
        for (int k = 0; k < N; ++k) ga[i][k] = fma(head, u[k], -half * v[k]);
    }
}

// everything after the MLP: J = Jh_q + Jh_o do/dq, nominal control, QP
template <int N>
__device__ __forceinline__ void cbf_filter_post(const float* __restrict__ prm, const float (&q)[N], const float (&dodq)[N],
                                                const float (&goal)[N], bool goal_is_uref, float lam, float penalty,
                                                float h, const float (&jac)[N + 1], float (&J)[N], float (&u)[N],
                                                float& r) {
#pragma unroll
    for (int k = 0; k < N; ++k) J[k] = fmaf(jac[N], dodq[k], jac[k]);
    using PL = ParamLayout<N>;
    float uref[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < N; ++k) acc = fmaf(prm[PL::K + i * N + k], q[k] - goal[k], acc);
        uref[i] = fminf(fmaxf(-acc, prm[PL::LO + i]), prm[PL::HI + i]);
        if (goal_is_uref) uref[i] = goal[i];  // caller-supplied u_ref is used verbatim (clf_controller.py:497-504)
    }
    qp_solve<N>(J, uref, prm + PL::LO, prm + PL::HI, lam * h, fminf(penalty, 1e6f), u, r);
}

// datax = [q | o | do/dq]  ->  h, J, u, r   (one controller.u call of the reference); FP32-FMA MLP
template <int N, int H, int L, int NT>
__device__ __forceinline__ void cbf_filter_state(const float* __restrict__ w, const float* __restrict__ prm,
                                                 float* __restrict__ scr, const float (&q)[N], float o,
                                                 const float (&dodq)[N], const float (&goal)[N], bool goal_is_uref,
                                                 float lam, float penalty, float& h, float (&J)[N], float (&u)[N],
                                                 float& r) {
    float x[N + 1], jac[N + 1];
#pragma unroll
    for (int k = 0; k < N; ++k) x[k] = q[k];
    x[N] = o;
    mlp_h_jac<N + 1, H, L, NT>(w, scr, x, h, jac);
    cbf_filter_post<N>(prm, q, dodq, goal, goal_is_uref, lam, penalty, h, jac, J, u, r);
}

}  // namespace cbfrrt

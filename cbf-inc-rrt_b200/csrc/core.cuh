// core.cuh -- per-state device code of the CBF-INC-RRT hot path (sm_100a).
//
// One thread owns one manipulator state.  The chain is walked once with everything in registers:
//   FK frame of link l  ->  world centres of that link's spheres  ->  distance sweep over the scene that the CTA
//   staged in shared memory (TMA bulk copy, see kernels.cu)  ->  running minima  ->  self-collision predicate
//   ->  arg-min recovery + geometric-Jacobian gradient.
//
// DETERMINISTIC-MATH CONTRACT: every value that feeds a boolean or an index is produced by the
// same sequence of single IEEE binary32 operations as oracle/cbfrrt_oracle.c -- written here with explicit
// __fadd_rn/__fsub_rn/__fmul_rn/fmaf so that neither -fmad nor instruction selection can re-associate or
// contract anything.  Two algebraic restructurings are used, both EXACT (bit-identical results):
//   (1) box sweep without sqrt: for one sphere, sd_j = fl(fl(fl(sqrt(dsq_j)) + mi_j) - r) where at most one of
//       the two inner terms is non-zero, and correctly rounded sqrt / subtraction are monotone, hence
//       min_j sd_j = fl(fl(sqrt(min_j dsq_j)) + min(min_j max3(a_j), 0)) - r);
//   (2) self-collision by squared-distance thresholds: the oracle predicate fl(fl(sqrt(dsq)-ri)-rj) > 0.01f is
//       monotone in dsq, so it equals dsq > X_ij with X_ij tabulated per sphere pair (robot_tables.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "robot_tables.cuh"
#include "grid.cuh"

namespace cbfrrt {

// ------------------------------------------------------------------ robot traits over the generated tables
struct Panda {
    static constexpr int ID = 0;
    static constexpr int N = PANDA_N, NLINKS = PANDA_NLINKS, NSPH = PANDA_NSPHERES, NSELFSPH = PANDA_NSELFSPH;
    static constexpr float LINK0_FLOOR = PANDA_LINK0_FLOOR_DIST;
    __device__ static constexpr int parent(int l) { return PANDA_PARENT[l]; }
    __device__ static constexpr int qidx(int l) { return PANDA_QIDX[l]; }
    __device__ static constexpr float t(int l, int k) { return PANDA_T[l][k]; }
    __device__ static constexpr float rfix(int l, int k) { return PANDA_R[l][k]; }
    __device__ static constexpr int sph_begin(int l) { return PANDA_SPH_BEGIN[l + 1]; }
    __device__ static constexpr int sph_end(int l) { return PANDA_SPH_BEGIN[l + 2]; }
    __device__ static constexpr float sph(int i, int k) { return PANDA_SPH[i][k]; }
    __device__ static constexpr int sph_link(int i) { return PANDA_SPH_LINK[i]; }
    __device__ static float rad_dyn(int i) { return PANDA_SPH[i][3]; }           // run-time index: a global-memory read
    __device__ static constexpr int anc_mask(int l) { return PANDA_ANC_MASK[l]; }
    __device__ static constexpr float ball(int l, int k) { return PANDA_LINK_BALL[l + 1][k]; }
    __device__ static constexpr int self_i(int p) { return PANDA_SELFSPH_I[p]; }
    __device__ static constexpr int self_j(int p) { return PANDA_SELFSPH_J[p]; }
    __device__ static constexpr float self_x(int p) { return PANDA_SELFSPH_X[p]; }
    __device__ static constexpr float qlo(int j) { return PANDA_QLO[j]; }
    __device__ static constexpr float qhi(int j) { return PANDA_QHI[j]; }
};
struct Magician {
    static constexpr int ID = 1;
    static constexpr int N = MAGICIAN_N, NLINKS = MAGICIAN_NLINKS, NSPH = MAGICIAN_NSPHERES,
                         NSELFSPH = MAGICIAN_NSELFSPH;
    static constexpr float LINK0_FLOOR = MAGICIAN_LINK0_FLOOR_DIST;
    __device__ static constexpr int parent(int l) { return MAGICIAN_PARENT[l]; }
    __device__ static constexpr int qidx(int l) { return MAGICIAN_QIDX[l]; }
    __device__ static constexpr float t(int l, int k) { return MAGICIAN_T[l][k]; }
    __device__ static constexpr float rfix(int l, int k) { return MAGICIAN_R[l][k]; }
    __device__ static constexpr int sph_begin(int l) { return MAGICIAN_SPH_BEGIN[l + 1]; }
    __device__ static constexpr int sph_end(int l) { return MAGICIAN_SPH_BEGIN[l + 2]; }
    __device__ static constexpr float sph(int i, int k) { return MAGICIAN_SPH[i][k]; }
    __device__ static constexpr int sph_link(int i) { return MAGICIAN_SPH_LINK[i]; }
    __device__ static float rad_dyn(int i) { return MAGICIAN_SPH[i][3]; }
    __device__ static constexpr int anc_mask(int l) { return MAGICIAN_ANC_MASK[l]; }
    __device__ static constexpr float ball(int l, int k) { return MAGICIAN_LINK_BALL[l + 1][k]; }
    __device__ static constexpr int self_i(int p) { return MAGICIAN_SELFSPH_I[p]; }
    __device__ static constexpr int self_j(int p) { return MAGICIAN_SELFSPH_J[p]; }
    __device__ static constexpr float self_x(int p) { return MAGICIAN_SELFSPH_X[p]; }
    __device__ static constexpr float qlo(int j) { return MAGICIAN_QLO[j]; }
    __device__ static constexpr float qhi(int j) { return MAGICIAN_QHI[j]; }
};

// compile-time loop: f(std::integral_constant<int, I>) for I in [B, E)
template <int I>
struct IC {
    static constexpr int value = I;
};
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(IC<B>{});
        static_for<B + 1, E>(f);
    }
}

#define CBF_INF __int_as_float(0x7f800000)

// ------------------------------------------------------------------ contract sincos (oracle: det_sincosf)
__device__ __forceinline__ void det_sincosf(float x, float& so, float& co) {
    const float k = rintf(__fmul_rn(x, 0.636619772f));
    float r = fmaf(-k, 1.57079637050628662109375f, x);
    r = fmaf(-k, -4.37113900018624283e-8f, r);
    const float z = __fmul_rn(r, r);
    float sp = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fmaf(z, sp, -1.6666654611e-1f);
    sp = __fmul_rn(sp, z);
    const float s = fmaf(sp, r, r);
    float cp = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fmaf(z, cp, 4.166664568298827e-2f);
    cp = fmaf(z, cp, -0.5f);
    const float c = fmaf(z, cp, 1.0f);
    const int n = ((int)k) & 3;
    const float a = (n & 1) ? c : s;  // sin-slot magnitude
    const float b = (n & 1) ? s : c;  // cos-slot magnitude
    so = (n & 2) ? -a : a;            // n: 0 -> s, 1 -> c, 2 -> -s, 3 -> -c
    co = ((n + 1) & 2) ? -b : b;      // n: 0 -> c, 1 -> -s, 2 -> -c, 3 -> s
}

// ------------------------------------------------------------------ frames
struct Frame {
    float p[3];
    float R[9];  // row-major
};

// x*F + acc with F a compile-time table constant: multiplications by 0 / +-1 are dropped.  Dropping them is
// exact: fmaf(x, 0, acc) == acc, x*(+-1) == +-x and fmaf(x, +-1, 0-product) == +-x for finite x; only the sign
// of an exact zero can differ from the oracle's general formula, which no comparison can observe.
template <class RB, int L>
__device__ __forceinline__ void child_frame(const Frame& par, bool par_is_base, float sq, float cq, Frame& out) {
    // position: p = fma(Rp[r][2], t2, fma(Rp[r][1], t1, fma(Rp[r][0], t0, pp[r])))
    constexpr float t0 = RB::t(L, 0), t1 = RB::t(L, 1), t2 = RB::t(L, 2);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = par.p[r];
        if constexpr (t0 != 0.f) acc = fmaf(par.R[3 * r + 0], t0, acc);
        if constexpr (t1 != 0.f) acc = fmaf(par.R[3 * r + 1], t1, acc);
        if constexpr (t2 != 0.f) acc = fmaf(par.R[3 * r + 2], t2, acc);
        out.p[r] = acc;
    }
    // A = Rp * F : A[r][c] = fma(Rp[r][2], F[2][c], fma(Rp[r][1], F[1][c], Rp[r][0]*F[0][c]))
    float A[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        static_for<0, 3>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            constexpr float f0 = RB::rfix(L, c), f1 = RB::rfix(L, 3 + c), f2 = RB::rfix(L, 6 + c);
            float acc = 0.f;
            if constexpr (f0 == 1.f) acc = par.R[3 * r + 0];
            else if constexpr (f0 == -1.f) acc = -par.R[3 * r + 0];
            else if constexpr (f0 != 0.f) acc = __fmul_rn(par.R[3 * r + 0], f0);
            if constexpr (f1 != 0.f) {
                if constexpr (f0 == 0.f && f1 == 1.f) acc = par.R[3 * r + 1];
                else if constexpr (f0 == 0.f && f1 == -1.f) acc = -par.R[3 * r + 1];
                else acc = fmaf(par.R[3 * r + 1], f1, acc);
            }
            if constexpr (f2 != 0.f) {
                if constexpr (f0 == 0.f && f1 == 0.f && f2 == 1.f) acc = par.R[3 * r + 2];
                else if constexpr (f0 == 0.f && f1 == 0.f && f2 == -1.f) acc = -par.R[3 * r + 2];
                else acc = fmaf(par.R[3 * r + 2], f2, acc);
            }
            A[3 * r + c] = acc;
        });
    }
    if constexpr (RB::qidx(L) >= 0) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float a0 = A[3 * r + 0], a1 = A[3 * r + 1];
            out.R[3 * r + 0] = fmaf(a1, sq, __fmul_rn(a0, cq));
            out.R[3 * r + 1] = fmaf(a1, cq, -__fmul_rn(a0, sq));
            out.R[3 * r + 2] = A[3 * r + 2];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) out.R[k] = A[k];
    }
}

__device__ __forceinline__ Frame identity_frame() {
    Frame f;
    f.p[0] = f.p[1] = f.p[2] = 0.f;
    f.R[0] = 1.f; f.R[1] = 0.f; f.R[2] = 0.f;
    f.R[3] = 0.f; f.R[4] = 1.f; f.R[5] = 0.f;
    f.R[6] = 0.f; f.R[7] = 0.f; f.R[8] = 1.f;
    return f;
}

// ------------------------------------------------------------------ scene view in shared memory
struct SceneS {
    const float4* boxes;    // 2 x float4 per box: (cx cy cz hx) (hy hz 0 0)
    const float4* spheres;  // (cx cy cz r)
    int n_boxes, n_spheres, floor;
    const uint16_t* grid;   // nearest-candidate grid in global memory (grid.cuh) or nullptr = brute force
    float cull_floor;       // max(thr_soft, thr_hard): pair distances above max(this, running min) cannot matter
    float thr_min;          // min(thr_soft, thr_hard): a pair distance at or below it decides both masks
    float* scratch = nullptr;   // SCR instantiations of the field path: this thread's column of the CTA's shared-memory
    int scratch_stride = 0;     // sphere-centre scratch -- component c of sphere i at scratch[(3 i + c) * scratch_stride]
};

// exact per-pair formulas (oracle: sd_box / sd_sphere / normal_box / normal_sphere)
__device__ __forceinline__ float sd_box_exact(float cx, float cy, float cz, float4 b0, float4 b1, float r) {
    const float dx = __fsub_rn(cx, b0.x), dy = __fsub_rn(cy, b0.y), dz = __fsub_rn(cz, b0.z);
    const float ax = __fsub_rn(fabsf(dx), b0.w), ay = __fsub_rn(fabsf(dy), b1.x), az = __fsub_rn(fabsf(dz), b1.y);
    const float qx = fmaxf(ax, 0.f), qy = fmaxf(ay, 0.f), qz = fmaxf(az, 0.f);
    const float dsq = fmaf(qz, qz, fmaf(qy, qy, __fmul_rn(qx, qx)));
    const float mi = fminf(fmaxf(ax, fmaxf(ay, az)), 0.f);
    return __fsub_rn(__fadd_rn(__fsqrt_rn(dsq), mi), r);
}
__device__ __forceinline__ float sd_sphere_exact(float cx, float cy, float cz, float4 s, float r) {
    const float dx = __fsub_rn(cx, s.x), dy = __fsub_rn(cy, s.y), dz = __fsub_rn(cz, s.z);
    const float dsq = fmaf(dz, dz, fmaf(dy, dy, __fmul_rn(dx, dx)));
    return __fsub_rn(__fsub_rn(__fsqrt_rn(dsq), s.w), r);
}
__device__ __forceinline__ void normal_box(float cx, float cy, float cz, float4 b0, float4 b1, float* n) {
    const float dx = __fsub_rn(cx, b0.x), dy = __fsub_rn(cy, b0.y), dz = __fsub_rn(cz, b0.z);
    const float ax = __fsub_rn(fabsf(dx), b0.w), ay = __fsub_rn(fabsf(dy), b1.x), az = __fsub_rn(fabsf(dz), b1.y);
    const float qx = fmaxf(ax, 0.f), qy = fmaxf(ay, 0.f), qz = fmaxf(az, 0.f);
    const float dsq = fmaf(qz, qz, fmaf(qy, qy, __fmul_rn(qx, qx)));
    if (dsq > 0.f) {
        const float d = __fsqrt_rn(dsq);
        n[0] = __fdiv_rn(copysignf(qx, dx), d);
        n[1] = __fdiv_rn(copysignf(qy, dy), d);
        n[2] = __fdiv_rn(copysignf(qz, dz), d);
    } else {
        n[0] = n[1] = n[2] = 0.f;
        if (ax >= ay && ax >= az) n[0] = copysignf(1.f, dx);
        else if (ay >= az) n[1] = copysignf(1.f, dy);
        else n[2] = copysignf(1.f, dz);
    }
}
__device__ __forceinline__ void normal_sphere(float cx, float cy, float cz, float4 s, float* n) {
    const float dx = __fsub_rn(cx, s.x), dy = __fsub_rn(cy, s.y), dz = __fsub_rn(cz, s.z);
    const float dsq = fmaf(dz, dz, fmaf(dy, dy, __fmul_rn(dx, dx)));
    if (dsq > 0.f) {
        const float d = __fsqrt_rn(dsq);
        n[0] = __fdiv_rn(dx, d); n[1] = __fdiv_rn(dy, d); n[2] = __fdiv_rn(dz, d);
    } else {
        n[0] = 0.f; n[1] = 0.f; n[2] = 1.f;
    }
}

// one box / one sphere obstacle against one link sphere: running (min dsq, min max3) resp. min (dist - rs)
// D accumulates 4 * dsq: max(a, 0) is taken as (a + |a|) = 2 max(a, 0) -- an FADD on the full-rate FMA pipe instead
// of an FMNMX on the half-rate ALU pipe (the sweep is ALU-pipe bound, profiles/r1b_*).  Scaling by 2 / 4 is exact in
// binary floating point, and sqrt(4 x) = 2 sqrt(x) exactly, so 0.5 * sqrt(D) below is bit-identical to sqrt(dsq).
__device__ __forceinline__ void acc_box(float cx, float cy, float cz, float4 b0, float4 b1, float& D, float& M) {
    const float ax = __fsub_rn(fabsf(__fsub_rn(cx, b0.x)), b0.w);
    const float ay = __fsub_rn(fabsf(__fsub_rn(cy, b0.y)), b1.x);
    const float az = __fsub_rn(fabsf(__fsub_rn(cz, b0.z)), b1.y);
    const float tx = __fadd_rn(ax, fabsf(ax)), ty = __fadd_rn(ay, fabsf(ay)), tz = __fadd_rn(az, fabsf(az));
    D = fminf(D, fmaf(tz, tz, fmaf(ty, ty, __fmul_rn(tx, tx))));
    M = fminf(M, fmaxf(ax, fmaxf(ay, az)));
}
__device__ __forceinline__ void acc_sphere(float cx, float cy, float cz, float4 s, float& E) {
    const float dx = __fsub_rn(cx, s.x), dy = __fsub_rn(cy, s.y), dz = __fsub_rn(cz, s.z);
    E = fminf(E, __fsub_rn(__fsqrt_rn(fmaf(dz, dz, fmaf(dy, dy, __fmul_rn(dx, dx)))), s.w));
}

// min over the scene's boxes and spheres of the signed distance of K spheres (centres C[k], radii r[k]) that
// belong to one link -- restructuring (1) above.  out[k] = min_j sd(k, j)  (floor not included).
// With a nearest-candidate grid only the candidates of each centre's cell are swept (exact, see grid.cuh); a link
// with a centre outside the grid region or in an overflowed cell takes the brute-force sweep.
//
// CULL (exact branch-and-bound): the caller passes a ball (centre B, radius included in `reach`) that contains the K
// spheres and reach = max(T, 0) + R_ball + slack, where T >= every threshold the results are compared with and >= the
// running minimum.  An obstacle whose distance to B exceeds reach has pair distances > T for all K spheres: it can
// change neither the running minimum (nor its arg-min, found later by re-evaluation) nor a mask, so it is skipped.
// The test sweep over obstacles is warp-uniform (broadcast LDS) and leaves a per-lane bit mask; each lane then visits
// only its own survivors (typically 5-10 % of the obstacles).
constexpr int CULL_MIN_OBSTACLES = 4;
template <int K, bool GRID, bool CULL>
__device__ __forceinline__ void sweep_scene(const SceneS& sc, const float (&C)[K][3], const float (&rad)[K],
                                            float (&out)[K], float Bx = 0.f, float By = 0.f, float Bz = 0.f,
                                            float reach = 0.f) {
    float D[K], M[K], E[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { D[k] = CBF_INF; M[k] = CBF_INF; E[k] = CBF_INF; }
    bool brute = true;
    if constexpr (GRID) {
        const uint16_t* rec[K];
        brute = false;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            rec[k] = grid::lookup(sc.grid, C[k][0], C[k][1], C[k][2]);
            brute = brute || (rec[k] == nullptr);
        }
        if (!brute) {
            // first 16 bytes of a cell record = header + the first 6 candidates: one LDG.128 per sphere, issued for
            // all K spheres before any of them is consumed; longer lists (rare) are read element by element
            uint4 r0[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                r0[k] = __ldg(reinterpret_cast<const uint4*>(rec[k]));
                brute = brute || ((r0[k].x & 0xffffu) == grid::OVERFLOW);
            }
            if (!brute) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int nb = r0[k].x & 0xffffu, nt = nb + (int)(r0[k].x >> 16);
                    const uint32_t w[3] = {r0[k].y, r0[k].z, r0[k].w};
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        if (i < nt) {
                            const int j = (i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xffffu);
                            if (i < nb) acc_box(C[k][0], C[k][1], C[k][2], sc.boxes[2 * j], sc.boxes[2 * j + 1], D[k], M[k]);
                            else acc_sphere(C[k][0], C[k][1], C[k][2], sc.spheres[j], E[k]);
                        }
                    }
                    for (int i = 6; i < nt; ++i) {
                        const int j = __ldg(rec[k] + 2 + i);
                        if (i < nb) acc_box(C[k][0], C[k][1], C[k][2], sc.boxes[2 * j], sc.boxes[2 * j + 1], D[k], M[k]);
                        else acc_sphere(C[k][0], C[k][1], C[k][2], sc.spheres[j], E[k]);
                    }
                }
            }
        }
    }
    if (brute && CULL) {
        const float r2 = __fmul_rn(reach, reach);
        for (int j0 = 0; j0 < sc.n_boxes; j0 += 32) {
            const int je = min(32, sc.n_boxes - j0);
            unsigned keep = 0u;
            for (int jj = 0; jj < je; ++jj) {
                const float4 b0 = sc.boxes[2 * (j0 + jj)], b1 = sc.boxes[2 * (j0 + jj) + 1];
                const float ax = fabsf(Bx - b0.x) - b0.w, ay = fabsf(By - b0.y) - b1.x, az = fabsf(Bz - b0.z) - b1.y;
                const float qx = fmaxf(ax, 0.f), qy = fmaxf(ay, 0.f), qz = fmaxf(az, 0.f);
                const float dsq = fmaf(qz, qz, fmaf(qy, qy, qx * qx));
                keep |= (dsq > r2) ? 0u : (1u << jj);
            }
            while (keep) {
                const int j = j0 + __ffs(keep) - 1;
                keep &= keep - 1;
                const float4 b0 = sc.boxes[2 * j], b1 = sc.boxes[2 * j + 1];
#pragma unroll
                for (int k = 0; k < K; ++k) acc_box(C[k][0], C[k][1], C[k][2], b0, b1, D[k], M[k]);
            }
        }
        for (int j0 = 0; j0 < sc.n_spheres; j0 += 32) {
            const int je = min(32, sc.n_spheres - j0);
            unsigned keep = 0u;
            for (int jj = 0; jj < je; ++jj) {
                const float4 s = sc.spheres[j0 + jj];
                const float dx = Bx - s.x, dy = By - s.y, dz = Bz - s.z, rr = reach + s.w;   // rr > 0: reach > 0, s.w >= 0
                keep |= (fmaf(dz, dz, fmaf(dy, dy, dx * dx)) > rr * rr) ? 0u : (1u << jj);
            }
            while (keep) {
                const int j = j0 + __ffs(keep) - 1;
                keep &= keep - 1;
                const float4 s = sc.spheres[j];
#pragma unroll
                for (int k = 0; k < K; ++k) acc_sphere(C[k][0], C[k][1], C[k][2], s, E[k]);
            }
        }
    } else if (brute) {
        for (int j = 0; j < sc.n_boxes; ++j) {
            const float4 b0 = sc.boxes[2 * j], b1 = sc.boxes[2 * j + 1];
#pragma unroll
            for (int k = 0; k < K; ++k) acc_box(C[k][0], C[k][1], C[k][2], b0, b1, D[k], M[k]);
        }
        for (int j = 0; j < sc.n_spheres; ++j) {
            const float4 s = sc.spheres[j];
#pragma unroll
            for (int k = 0; k < K; ++k) acc_sphere(C[k][0], C[k][1], C[k][2], s, E[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        // boxes: e = sqrt(min dsq) + min(min max3, 0); with no boxes D = M = +inf -> e = +inf
        const float e_box = __fadd_rn(__fmul_rn(__fsqrt_rn(D[k]), 0.5f), fminf(M[k], 0.f));   // D holds 4 * dsq
        out[k] = __fsub_rn(fminf(e_box, E[k]), rad[k]);
    }
}

// ------------------------------------------------------------------ per-state result
template <class RB>
struct CollideResult {
    float o;         // min signed distance over non-base links x (floor, boxes, spheres)  (arm_mindis.py:71-90)
    float soft_min;  // as checked by check_collision_free_soft  (arm_dynamics.py:169-182), before self test
    float hard_min;  // as checked by check_collision_free_hard  (arm_dynamics.py:184-187, arm_env.py:82-84)
    bool self_ok;    // no non-adjacent link pair within 0.01 m    (basic_robot.py:89-98)
    float self_min;  // only when WANT_SELFMIN
    int arg_sphere, arg_obs;
    float dodq[RB::N];
    unsigned n_refined, n_pairs;   // broad-phase statistics (field path): spheres swept exactly, pair evaluations
};

template <class RB, bool WANT_GRAD, bool WANT_SELF, bool SCR>
__device__ __forceinline__ void collide_state_field(const float (&q)[RB::N], const SceneS& sc, float base_min,
                                                    CollideResult<RB>& res);   // distance-bound field path, below

// base link (-1) is static: its distance to every obstacle is a per-scene constant that the CTA computes once
// (kernels.cu: base_min_cta) and passes in.  The floor is exempt for the base in both checks.
// GRID: the scene carries a nearest-candidate grid (sc.grid != nullptr); grid-free instantiations keep the small-
// scene kernels free of the look-up code and its registers.
// WANT_SELF = false skips the self-collision predicate (res.self_ok stays true): the unsafe test of the steer loop
// (tsa.py:251, hard check only) does not read it.
// SCR: the field path keeps the sphere centres of pass A in the CTA's shared memory (sc.scratch) instead of a per-thread
// local-memory array: kernels that have the room (no network weights) avoid the write-back of that scratch to DRAM.
template <class RB, bool WANT_GRAD, bool WANT_SELFMIN, bool GRID, bool WANT_SELF = true, bool SCR = false>
__device__ __forceinline__ void collide_state(const float (&q)[RB::N], const SceneS& sc, float base_min,
                                              CollideResult<RB>& res) {
    constexpr int N = RB::N;
    res.n_refined = 0u; res.n_pairs = 0u;
    if constexpr (GRID && !WANT_SELFMIN) {           // distance-bound field: exact sweeps only where they can matter
        collide_state_field<RB, WANT_GRAD, WANT_SELF, SCR>(q, sc, base_min, res);
        return;
    }
    float o = CBF_INF, soft = base_min, hard = base_min;
    int arg_s = -1;
    float argC[3] = {0.f, 0.f, 0.f};
    // joint frames needed by the gradient: z axis + origin of every revolute joint frame
    float jz[WANT_GRAD ? N : 1][3], jp[WANT_GRAD ? N : 1][3];
    // sphere centres kept for the self-collision predicate (only spheres that appear in a self pair)
    float SC[RB::NSPH][3];
    bool self_ok = true;
    float self_min = CBF_INF;

    Frame fr[RB::NLINKS];  // fully unrolled: only live frames stay in registers
    static_for<0, RB::NLINKS>([&](auto lc) {
        constexpr int L = decltype(lc)::value;
        constexpr int P = RB::parent(L);
        float sq = 0.f, cq = 1.f;
        if constexpr (RB::qidx(L) >= 0) det_sincosf(q[RB::qidx(L)], sq, cq);
        if constexpr (P < 0) {
            const Frame base = identity_frame();
            child_frame<RB, L>(base, true, sq, cq, fr[L]);
        } else {
            child_frame<RB, L>(fr[P], false, sq, cq, fr[L]);
        }
        if constexpr (WANT_GRAD && RB::qidx(L) >= 0) {
            constexpr int J = RB::qidx(L);
            jz[J][0] = fr[L].R[2]; jz[J][1] = fr[L].R[5]; jz[J][2] = fr[L].R[8];
            jp[J][0] = fr[L].p[0]; jp[J][1] = fr[L].p[1]; jp[J][2] = fr[L].p[2];
        }
        constexpr int S0 = RB::sph_begin(L), S1 = RB::sph_end(L), K = S1 - S0;
        if constexpr (K > 0) {
            float C[K][3], rad[K], m[K];
            static_for<0, K>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                constexpr int i = S0 + k;
                constexpr float c0 = RB::sph(i, 0), c1 = RB::sph(i, 1), c2 = RB::sph(i, 2);
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    C[k][r] = fmaf(fr[L].R[3 * r + 2], c2, fmaf(fr[L].R[3 * r + 1], c1, fmaf(fr[L].R[3 * r + 0], c0, fr[L].p[r])));
                rad[k] = RB::sph(i, 3);
                SC[i][0] = C[k][0]; SC[i][1] = C[k][1]; SC[i][2] = C[k][2];
            });
            // self collision, restructuring (2): pairs whose second sphere belongs to this link are tested now, so
            // only the spheres of the (few) "first" links of a pair stay live along the chain
            static_for<0, RB::NSELFSPH>([&](auto pc) {
                constexpr int p = decltype(pc)::value;
                constexpr int i = RB::self_i(p), j = RB::self_j(p);
                if constexpr (WANT_SELF && j >= S0 && j < S1) {
                    const float dx = __fsub_rn(SC[i][0], SC[j][0]), dy = __fsub_rn(SC[i][1], SC[j][1]), dz = __fsub_rn(SC[i][2], SC[j][2]);
                    const float dsq = fmaf(dz, dz, fmaf(dy, dy, __fmul_rn(dx, dx)));
                    self_ok = self_ok && (dsq > RB::self_x(p));
                    if constexpr (WANT_SELFMIN)
                        self_min = fminf(self_min, __fsub_rn(__fsub_rn(__fsqrt_rn(dsq), RB::sph(i, 3)), RB::sph(j, 3)));
                }
            });
            // culling pays when a ball test (about one pair evaluation) can save K >= 3 of them and there is more than a
            // handful of obstacles; CULL_MIN_OBSTACLES is warp-uniform
            if (!WANT_SELFMIN && !GRID && K >= 3 && sc.n_boxes + sc.n_spheres >= CULL_MIN_OBSTACLES) {
                // bound for the exact culling test: running minimum incl. this link's floor candidates (merged
                // for real, in oracle order, below), never below the thresholds the masks compare with
                // (mask-only callers -- WANT_GRAD false: edges, masks, the non-refresh steps of a rollout -- never
                // read the minimum itself, so for them the bound is just the mask thresholds)
                float tb = -CBF_INF;
                if constexpr (WANT_GRAD) {
                    tb = o;
                    if (sc.floor) {
                        if constexpr (L == 0) tb = fminf(tb, RB::LINK0_FLOOR);
                        else {
#pragma unroll
                            for (int k = 0; k < K; ++k) tb = fminf(tb, C[k][2] - rad[k]);
                        }
                    }
                }
                constexpr float b0 = RB::ball(L, 0), b1 = RB::ball(L, 1), b2 = RB::ball(L, 2);
                float Bc[3];
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    Bc[r] = fmaf(fr[L].R[3 * r + 2], b2, fmaf(fr[L].R[3 * r + 1], b1, fmaf(fr[L].R[3 * r + 0], b0, fr[L].p[r])));
                const float reach = fmaxf(fmaxf(tb, sc.cull_floor), 0.f) + (RB::ball(L, 3) + 1e-4f);
                sweep_scene<K, false, true>(sc, C, rad, m, Bc[0], Bc[1], Bc[2], reach);
            } else {
                sweep_scene<K, GRID, false>(sc, C, rad, m);
            }
            // merge in sphere order, floor candidate first (oracle iteration order; strict '<')
            static_for<0, K>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                constexpr int i = S0 + k;
                if (sc.floor) {
                    if constexpr (L == 0) {
                        if constexpr (k == 0) {  // kinematic constant attached to the link's first sphere
                            const float d = RB::LINK0_FLOOR;
                            if (d < o) { o = d; arg_s = i; argC[0] = C[k][0]; argC[1] = C[k][1]; argC[2] = C[k][2]; }
                            hard = fminf(hard, d);  // soft: link 0 is floor-exempt (arm_dynamics.py:180)
                        }
                    } else {
                        const float d = __fsub_rn(C[k][2], rad[k]);
                        if (d < o) { o = d; arg_s = i; argC[0] = C[k][0]; argC[1] = C[k][1]; argC[2] = C[k][2]; }
                        soft = fminf(soft, d);
                        if constexpr (L != 1) hard = fminf(hard, d);  // link 1 is filtered vs plane (arm_env.py:84)
                    }
                }
                if (m[k] < o) { o = m[k]; arg_s = i; argC[0] = C[k][0]; argC[1] = C[k][1]; argC[2] = C[k][2]; }
                soft = fminf(soft, m[k]);
                hard = fminf(hard, m[k]);
            });
        }
    });

    res.o = o; res.soft_min = soft; res.hard_min = hard; res.self_ok = self_ok; res.self_min = self_min;
    res.arg_sphere = arg_s; res.arg_obs = -1;
#pragma unroll
    for (int k = 0; k < N; ++k) res.dodq[k] = 0.f;
    if constexpr (WANT_GRAD) {
        if (arg_s < 0) return;
        // recover the obstacle: first j (floor, boxes, spheres) whose exact pair distance equals o
        const int link = RB::sph_link(arg_s);
        const float r = RB::sph(arg_s, 3);
        const int off = sc.floor ? 1 : 0;
        float nrm[3] = {0.f, 0.f, 1.f};
        int arg_o = -1;
        if (sc.floor) {
            const float d = (link == 0) ? RB::LINK0_FLOOR : __fsub_rn(argC[2], r);
            if (d == o) arg_o = 0;
        }
        // candidates of the winning sphere's cell (ascending indices, exact -- grid.cuh) or every obstacle
        const uint16_t* rec = GRID ? grid::lookup(sc.grid, argC[0], argC[1], argC[2]) : nullptr;
        int nb = sc.n_boxes, ns = sc.n_spheres;
        if (rec) {
            const uint32_t hdr = __ldg(reinterpret_cast<const unsigned int*>(rec));
            if ((hdr & 0xffffu) == grid::OVERFLOW) rec = nullptr;
            else { nb = hdr & 0xffffu; ns = hdr >> 16; }
        }
        if (arg_o < 0) {
            for (int i = 0; i < nb; ++i) {
                const int j = rec ? (int)__ldg(rec + 2 + i) : i;
                const float4 b0 = sc.boxes[2 * j], b1 = sc.boxes[2 * j + 1];
                if (sd_box_exact(argC[0], argC[1], argC[2], b0, b1, r) == o) {
                    arg_o = off + j;
                    normal_box(argC[0], argC[1], argC[2], b0, b1, nrm);
                    break;
                }
            }
        }
        if (arg_o < 0) {
            for (int i = 0; i < ns; ++i) {
                const int j = rec ? (int)__ldg(rec + 2 + nb + i) : i;
                const float4 s = sc.spheres[j];
                if (sd_sphere_exact(argC[0], argC[1], argC[2], s, r) == o) {
                    arg_o = off + sc.n_boxes + j;
                    normal_sphere(argC[0], argC[1], argC[2], s, nrm);
                    break;
                }
            }
        }
        res.arg_obs = arg_o;
        // n^T (z_j x (C - p_j)) for the joints that move the arg-min link (arm_mindis.py:56-69)
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

// ------------------------------------------------------------------ distance-bound field path (grid.cuh)
// exact  min_j e_j(c)  over the scene for ONE sphere centre: the candidates of the centre's 5 cm cell (ascending
// obstacle order, exact per-pair formulas), brute force when the centre is outside the region or its cell overflowed.
// Same restructuring (1) as sweep_scene: running (min 4 dsq, min max3) and one sqrt at the end.
__device__ __forceinline__ float scene_min_exact(const SceneS& sc, float cx, float cy, float cz, unsigned& n_pairs) {
    float D = CBF_INF, M = CBF_INF, E = CBF_INF;
    const uint16_t* rec = grid::lookup(sc.grid, cx, cy, cz);
    uint4 r0 = make_uint4(grid::OVERFLOW, 0u, 0u, 0u);
    if (rec) r0 = __ldg(reinterpret_cast<const uint4*>(rec));
    if ((r0.x & 0xffffu) != grid::OVERFLOW) {
        const int nb = r0.x & 0xffffu, nt = nb + (int)(r0.x >> 16);
        n_pairs += nt;
        const uint32_t w[3] = {r0.y, r0.z, r0.w};
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i < nt) {
                const int j = (i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xffffu);
                if (i < nb) acc_box(cx, cy, cz, sc.boxes[2 * j], sc.boxes[2 * j + 1], D, M);
                else acc_sphere(cx, cy, cz, sc.spheres[j], E);
            }
        }
        for (int i = 6; i < nt; ++i) {
            const int j = __ldg(rec + 2 + i);
            if (i < nb) acc_box(cx, cy, cz, sc.boxes[2 * j], sc.boxes[2 * j + 1], D, M);
            else acc_sphere(cx, cy, cz, sc.spheres[j], E);
        }
    } else {
        n_pairs += sc.n_boxes + sc.n_spheres;
        for (int j = 0; j < sc.n_boxes; ++j) acc_box(cx, cy, cz, sc.boxes[2 * j], sc.boxes[2 * j + 1], D, M);
        for (int j = 0; j < sc.n_spheres; ++j) acc_sphere(cx, cy, cz, sc.spheres[j], E);
    }
    const float e_box = __fadd_rn(__fmul_rn(__fsqrt_rn(D), 0.5f), fminf(M, 0.f));   // D holds 4 * dsq
    return fminf(e_box, E);
}

// collide_state on the distance-bound field (sc.grid != nullptr).  Pass A walks the chain once: frames, sphere centres
// (kept in a per-thread indexed array for pass B), the exact floor candidates in oracle order, the self-collision
// predicate, and ONE 4-byte field look-up per sphere.  From the look-ups every sphere's obstacle term
// m_i = min_j e_j(C_i) - r_i is bracketed:  lo_i <= m_i <= hi_i.  Pass B sweeps exactly (scene_min_exact) only the
// spheres with lo_i <= B, in ascending sphere order, where
//     B = min(min_k hi_k, max(o_floor, thr_soft, thr_hard))        observation callers (WANT_GRAD)
//     B = min(min_k hi_k, max(thr_soft, thr_hard))                 mask-only callers
// Why that is exact.  M* = min_i m_i <= min_k hi_k, so with B = min_k hi_k every sphere that attains M* (ties included)
// is swept and a skipped sphere is strictly above M*.  When B is the second term, a skipped sphere has
// m_i > max(o_floor, thresholds): it can neither win or tie the observation minimum (which is <= o_floor) nor move a
// mask.  Mask-only callers stop even earlier: min_k hi_k <= min(thr_soft, thr_hard) already proves unsafe and not safe.
// The merge reproduces the oracle's scan order (sphere-major, floor candidate before the obstacle term, strict '<').
template <class RB, bool WANT_GRAD, bool WANT_SELF, bool SCR>
__device__ __forceinline__ void collide_state_field(const float (&q)[RB::N], const SceneS& sc, float base_min,
                                                    CollideResult<RB>& res) {
    constexpr int N = RB::N, NS = RB::NSPH;
    constexpr float WIDEN = grid::FIELD_HALF_DIAG + grid::SLACK;
    const float* __restrict__ field = grid::field_of(sc.grid);
    float Cx[SCR ? 1 : NS], Cy[SCR ? 1 : NS], Cz[SCR ? 1 : NS];   // indexed dynamically in pass B -> local memory, or (SCR) shared
    float* const scr = sc.scratch;
    const int sst = sc.scratch_stride;
    constexpr int SFIRST = RB::sph_begin(0);  // scratch rows start at the first non-base sphere

    float fc[NS];                            // field samples (static indices: registers)
    float o_fl = CBF_INF, soft_fl = CBF_INF, hard_fl = CBF_INF;
    int arg_fl = -1;
    float jz[WANT_GRAD ? N : 1][3], jp[WANT_GRAD ? N : 1][3];
    float SC[NS][3];
    bool self_ok = true;
    unsigned outside = 0u;

    Frame fr[RB::NLINKS];
    static_for<0, RB::NLINKS>([&](auto lc) {
        constexpr int L = decltype(lc)::value;
        constexpr int P = RB::parent(L);
        float sq = 0.f, cq = 1.f;
        if constexpr (RB::qidx(L) >= 0) det_sincosf(q[RB::qidx(L)], sq, cq);
        if constexpr (P < 0) {
            const Frame base = identity_frame();
            child_frame<RB, L>(base, true, sq, cq, fr[L]);
        } else {
            child_frame<RB, L>(fr[P], false, sq, cq, fr[L]);
        }
        if constexpr (WANT_GRAD && RB::qidx(L) >= 0) {
            constexpr int J = RB::qidx(L);
            jz[J][0] = fr[L].R[2]; jz[J][1] = fr[L].R[5]; jz[J][2] = fr[L].R[8];
            jp[J][0] = fr[L].p[0]; jp[J][1] = fr[L].p[1]; jp[J][2] = fr[L].p[2];
        }
        constexpr int S0 = RB::sph_begin(L), S1 = RB::sph_end(L), K = S1 - S0;
        if constexpr (K > 0) {
            static_for<0, K>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                constexpr int i = S0 + k;
                constexpr float c0 = RB::sph(i, 0), c1 = RB::sph(i, 1), c2 = RB::sph(i, 2), rad = RB::sph(i, 3);
                float C[3];
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    C[r] = fmaf(fr[L].R[3 * r + 2], c2, fmaf(fr[L].R[3 * r + 1], c1, fmaf(fr[L].R[3 * r + 0], c0, fr[L].p[r])));
                SC[i][0] = C[0]; SC[i][1] = C[1]; SC[i][2] = C[2];
                if constexpr (SCR) {
                    scr[(3 * (i - SFIRST) + 0) * sst] = C[0]; scr[(3 * (i - SFIRST) + 1) * sst] = C[1]; scr[(3 * (i - SFIRST) + 2) * sst] = C[2];
                } else {
                    Cx[i] = C[0]; Cy[i] = C[1]; Cz[i] = C[2];
                }
                const int cell = grid::field_index(C[0], C[1], C[2]);
                fc[i] = __ldg(field + max(cell, 0));
                if (cell < 0) outside |= 1u << i;
                // floor candidates, exact and in oracle order (same exemptions as collide_state)
                if (sc.floor) {
                    if constexpr (L == 0) {
                        if constexpr (k == 0) {
                            const float d = RB::LINK0_FLOOR;
                            if (d < o_fl) { o_fl = d; arg_fl = i; }
                            hard_fl = fminf(hard_fl, d);
                        }
                    } else {
                        const float d = __fsub_rn(C[2], rad);
                        if (d < o_fl) { o_fl = d; arg_fl = i; }
                        soft_fl = fminf(soft_fl, d);
                        if constexpr (L != 1) hard_fl = fminf(hard_fl, d);
                    }
                }
            });
            static_for<0, RB::NSELFSPH>([&](auto pc) {
                constexpr int p = decltype(pc)::value;
                constexpr int i = RB::self_i(p), j = RB::self_j(p);
                if constexpr (WANT_SELF && j >= S0 && j < S1) {
                    const float dx = __fsub_rn(SC[i][0], SC[j][0]), dy = __fsub_rn(SC[i][1], SC[j][1]), dz = __fsub_rn(SC[i][2], SC[j][2]);
                    const float dsq = fmaf(dz, dz, fmaf(dy, dy, __fmul_rn(dx, dx)));
                    self_ok = self_ok && (dsq > RB::self_x(p));
                }
            });
        }
    });

    auto centre = [&](int i, int c) -> float {       // pass B / arg-min recovery: the centres of pass A
        if constexpr (SCR) return scr[(3 * (i - SFIRST) + c) * sst];
        else return c == 0 ? Cx[i] : (c == 1 ? Cy[i] : Cz[i]);
    };
    // brackets of the obstacle terms and the pruning bound
    constexpr int FIRST = RB::sph_begin(0);          // spheres of the static base link are not part of the sweep
    float best_hi = CBF_INF;
    static_for<FIRST, NS>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        const float hi = ((outside >> i) & 1u) ? CBF_INF : fc[i] + (WIDEN - RB::sph(i, 3));
        best_hi = fminf(best_hi, hi);
    });
    float M = CBF_INF;
    int argM = -1;
    unsigned need = 0u;
    bool decided = false;
    if constexpr (!WANT_GRAD) decided = best_hi <= sc.thr_min;       // some pair is at or below both thresholds
    if (decided) {
        M = best_hi;
    } else {
        const float B = WANT_GRAD ? fminf(best_hi, fmaxf(o_fl, sc.cull_floor)) : fminf(best_hi, sc.cull_floor);
        static_for<FIRST, NS>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            const float lo = fc[i] - (WIDEN + RB::sph(i, 3));
            if (lo <= B || ((outside >> i) & 1u)) need |= 1u << i;
        });
    }
    unsigned n_ref = __popc(need), n_pairs = 0u;
    while (need) {                                   // ascending sphere order: strict '<' keeps the first minimum
        const int i = __ffs(need) - 1;
        need &= need - 1;
        const float m = __fsub_rn(scene_min_exact(sc, centre(i, 0), centre(i, 1), centre(i, 2), n_pairs), RB::rad_dyn(i));
        if (m < M) { M = m; argM = i; }
    }
    res.n_refined = n_ref; res.n_pairs = n_pairs;

    res.soft_min = fminf(fminf(base_min, soft_fl), M);
    res.hard_min = fminf(fminf(base_min, hard_fl), M);
    res.self_ok = self_ok; res.self_min = CBF_INF;
    // observation minimum: first element of the oracle's scan that attains it
    const bool obs_wins = (M < o_fl) || (M == o_fl && argM >= 0 && argM < arg_fl);
    const float o = obs_wins ? M : o_fl;
    const int arg_s = obs_wins ? argM : arg_fl;
    res.o = o; res.arg_sphere = arg_s; res.arg_obs = -1;
#pragma unroll
    for (int k = 0; k < N; ++k) res.dodq[k] = 0.f;
    if constexpr (WANT_GRAD) {
        if (arg_s < 0) return;
        const float argC[3] = {centre(arg_s, 0), centre(arg_s, 1), centre(arg_s, 2)};
        const int link = RB::sph_link(arg_s);
        const float r = RB::rad_dyn(arg_s);
        const int off = sc.floor ? 1 : 0;
        float nrm[3] = {0.f, 0.f, 1.f};
        int arg_o = obs_wins ? -1 : 0;               // the floor is obstacle 0
        if (arg_o < 0) {
            const uint16_t* rec = grid::lookup(sc.grid, argC[0], argC[1], argC[2]);
            int nb = sc.n_boxes, ns = sc.n_spheres;
            if (rec) {
                const uint32_t hdr = __ldg(reinterpret_cast<const unsigned int*>(rec));
                if ((hdr & 0xffffu) == grid::OVERFLOW) rec = nullptr;
                else { nb = hdr & 0xffffu; ns = hdr >> 16; }
            }
            for (int t = 0; t < nb; ++t) {
                const int j = rec ? (int)__ldg(rec + 2 + t) : t;
                const float4 b0 = sc.boxes[2 * j], b1 = sc.boxes[2 * j + 1];
                if (sd_box_exact(argC[0], argC[1], argC[2], b0, b1, r) == o) {
                    arg_o = off + j;
                    normal_box(argC[0], argC[1], argC[2], b0, b1, nrm);
                    break;
                }
            }
            if (arg_o < 0) {
                for (int t = 0; t < ns; ++t) {
                    const int j = rec ? (int)__ldg(rec + 2 + nb + t) : t;
                    const float4 sp = sc.spheres[j];
                    if (sd_sphere_exact(argC[0], argC[1], argC[2], sp, r) == o) {
                        arg_o = off + sc.n_boxes + j;
                        normal_sphere(argC[0], argC[1], argC[2], sp, nrm);
                        break;
                    }
                }
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

// FK only: all link frames (basic_robot.py:55-62)
template <class RB, class Sink>
__device__ __forceinline__ void fk_all(const float (&q)[RB::N], Sink&& sink) {
    Frame fr[RB::NLINKS];
    static_for<0, RB::NLINKS>([&](auto lc) {
        constexpr int L = decltype(lc)::value;
        constexpr int P = RB::parent(L);
        float sq = 0.f, cq = 1.f;
        if constexpr (RB::qidx(L) >= 0) det_sincosf(q[RB::qidx(L)], sq, cq);
        if constexpr (P < 0) {
            const Frame base = identity_frame();
            child_frame<RB, L>(base, true, sq, cq, fr[L]);
        } else {
            child_frame<RB, L>(fr[P], false, sq, cq, fr[L]);
        }
        sink(L, fr[L]);
    });
}

// distance of the static base link to the scene: every thread of the CTA takes a slice of the obstacles,
// then a block-wide min.  Order-independent (min of identical values), so bit-exact with the oracle.
template <class RB>
__device__ __forceinline__ float base_min_partial(const SceneS& sc, int tid, int nthreads) {
    float m = CBF_INF;
    constexpr int S0 = RB::sph_begin(-1), S1 = RB::sph_end(-1);
    for (int j = tid; j < sc.n_boxes; j += nthreads) {
        const float4 b0 = sc.boxes[2 * j], b1 = sc.boxes[2 * j + 1];
        static_for<S0, S1>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            m = fminf(m, sd_box_exact(RB::sph(i, 0), RB::sph(i, 1), RB::sph(i, 2), b0, b1, RB::sph(i, 3)));
        });
    }
    for (int j = tid; j < sc.n_spheres; j += nthreads) {
        const float4 s = sc.spheres[j];
        static_for<S0, S1>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            m = fminf(m, sd_sphere_exact(RB::sph(i, 0), RB::sph(i, 1), RB::sph(i, 2), s, RB::sph(i, 3)));
        });
    }
    return m;
}

}  // namespace cbfrrt

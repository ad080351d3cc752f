/*
 * TEST INFRASTRUCTURE -- scalar C restatement ("oracle") of the CBF-INC-RRT hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  The product (cbf-inc-rrt_b200/) never links, imports or executes it.
 *
 * PARITY STATUS: "parity unpinned" at the pybullet / SCS boundaries: the reference's FK,
 * closest-point and contact arithmetic lives in pybullet 3.2.5 and its QP in cvxpylayers 0.1.5 -> diffcp
 * 1.0.19 -> SCS 3.2.0; none is installed here and the reference ships no golden vectors.  Pinned pieces:
 * URDF constants, FK at q=0/q0 (hand-computed), LQR gain K (scipy DARE), and the MLP + Jacobian +
 * Lie-derivative + nominal-control code, which IS checked against the reference's own Python executed in
 * the dev container (tests/golden/make_golden.py), and the reference's loops AROUND the physics engine (link ranges,
 * exemptions, arg-min, do/dq assembly), which are run unmodified on a stand-in client backed by this repo's sphere
 * model and compared with this file (tests/test_reference_systems_cpu.py).  Unpinned: Bullet's own geometry.
 *
 * Reference code followed (file:line under /root/reference):
 *   FK of the URDF chain                 environment/basic_robot.py:55-68, utils/robot/{franka_panda/panda,magician/urdf/demo}.urdf
 *   safe / unsafe masks                  neural_cbf/systems/arm_dynamics.py:169-232, environment/basic_robot.py:89-98,
 *                                        environment/arm_env.py:82-84
 *   observation o = min distance, do/dq  neural_cbf/systems/arm_mindis.py:56-90
 *   barrier MLP h and its Jacobian       neural_cbf/controllers/neural_mindis_cbf_controller.py:41-53,71-102
 *   Lie derivatives (f=0, g=I)           neural_cbf/controllers/clf_controller.py:169-218, arm_dynamics.py:327-363
 *   nominal control                      neural_cbf/systems/arm_dynamics.py:124-161
 *   CBF-QP                               neural_cbf/controllers/clf_controller.py:82-139,354-404
 *   Euler step / steer loop              neural_cbf/systems/arm_dynamics.py:474-523, motion_planning/baseline/tsa.py:217-269
 *   edge check                           motion_planning/baseline/tsa.py:272-286 (fixed K and per-edge K = int(d/eps))
 *   multi-scenario QP (S > 1)            neural_cbf/controllers/clf_controller.py:86-139 (float64 dual active-set solve,
 *                                        certified by KKT residuals and SLSQP in tests/test_oracle_cpu.py)
 *   nearest / k-nearest / radius         motion_planning/baseline/batch_tsa.py:135-145, tsa.py:151-153, bit_star.py:196-216
 *                                        (checked against the reference's numpy expressions in tests/test_oracle_cpu.py)
 *
 * DETERMINISTIC FLOAT32 CONTRACT (DESIGN.md "Deterministic-math contract"): every geometric quantity that
 * feeds a boolean or an index is computed with exactly the operation sequence written below -- single
 * IEEE-754 binary32 operations (add, sub, mul, fmaf, sqrtf, div, rintf), no contraction (compile with
 * -ffp-contract=off), own sincos.  The CUDA kernels follow the same sequence with __fmul_rn/__fadd_rn/fmaf,
 * so masks and arg-min indices are bit-exact between the two.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "robot_tables.h"

#define MAXN 8
#define MAXLINKS 16
#define MAXSPH 64

typedef struct {
    int n, n_links, n_spheres, n_self;
    const int *parent, *qidx;
    const float (*T)[3];
    const float (*R)[9];
    const int *sph_link;
    const float (*sph)[4];
    const int (*self_pairs)[2];
    float link0_floor;
} Robot;

static const Robot ROBOTS[2] = {
    {PANDA_N, PANDA_NLINKS, PANDA_NSPHERES, PANDA_NSELFPAIRS, PANDA_PARENT, PANDA_QIDX, PANDA_T, PANDA_R,
     PANDA_SPH_LINK, PANDA_SPH, PANDA_SELF_PAIRS, PANDA_LINK0_FLOOR_DIST},
    {MAGICIAN_N, MAGICIAN_NLINKS, MAGICIAN_NSPHERES, MAGICIAN_NSELFPAIRS, MAGICIAN_PARENT, MAGICIAN_QIDX, MAGICIAN_T,
     MAGICIAN_R, MAGICIAN_SPH_LINK, MAGICIAN_SPH, MAGICIAN_SELF_PAIRS, MAGICIAN_LINK0_FLOOR_DIST},
};

/* exemptions (same for both robots): arm_dynamics.py:180 (soft: links -1,0 vs floor) and
 * arm_env.py:83-84 (hard: collision filter for links -1 and 1 vs plane) */
static inline int soft_floor_exempt(int link) { return link == -1 || link == 0; }
static inline int hard_floor_exempt(int link) { return link == -1 || link == 1; }

/* ------------------------------------------------------------------ contract sincos */
static inline void det_sincosf(float x, float *s_out, float *c_out) {
    const float TWO_OVER_PI = 0.636619772f;
    const float PIO2_HI = 1.57079637050628662109375f;
    const float PIO2_LO = -4.37113900018624283e-8f;
    float k = rintf(x * TWO_OVER_PI);
    float r = fmaf(-k, PIO2_HI, x);
    r = fmaf(-k, PIO2_LO, r);
    float z = r * r;
    /* sin(r) = r + r*z*(S1 + z*(S2 + z*S3)),  cos(r) = 1 + z*(-1/2 + z*(C1 + z*(C2 + z*C3))) on |r|<=pi/4 */
    float sp = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fmaf(z, sp, -1.6666654611e-1f);
    sp = sp * z;
    float s = fmaf(sp, r, r);
    float cp = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fmaf(z, cp, 4.166664568298827e-2f);
    cp = fmaf(z, cp, -0.5f);
    float c = fmaf(z, cp, 1.0f);
    int n = ((int)k) & 3;
    float so, co;
    switch (n) {
        case 0: so = s; co = c; break;
        case 1: so = c; co = -s; break;
        case 2: so = -s; co = -c; break;
        default: so = -c; co = s; break;
    }
    *s_out = so;
    *c_out = co;
}

/* ------------------------------------------------------------------ FK (contract order) */
typedef struct {
    float p[3];
    float R[9]; /* row-major */
} Frame;

static void fk_chain(const Robot *rb, const float *q, Frame *fr /* [n_links] */) {
    for (int i = 0; i < rb->n_links; ++i) {
        float pp[3] = {0.f, 0.f, 0.f};
        float Rp[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
        if (rb->parent[i] >= 0) {
            memcpy(pp, fr[rb->parent[i]].p, sizeof pp);
            memcpy(Rp, fr[rb->parent[i]].R, sizeof Rp);
        }
        const float *t = rb->T[i];
        const float *F = rb->R[i];
        Frame *o = &fr[i];
        for (int r = 0; r < 3; ++r)
            o->p[r] = fmaf(Rp[3 * r + 2], t[2], fmaf(Rp[3 * r + 1], t[1], fmaf(Rp[3 * r + 0], t[0], pp[r])));
        float A[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                A[3 * r + c] = fmaf(Rp[3 * r + 2], F[6 + c], fmaf(Rp[3 * r + 1], F[3 + c], Rp[3 * r + 0] * F[c]));
        if (rb->qidx[i] >= 0) {
            float s, c;
            det_sincosf(q[rb->qidx[i]], &s, &c);
            for (int r = 0; r < 3; ++r) {
                float a0 = A[3 * r + 0], a1 = A[3 * r + 1];
                o->R[3 * r + 0] = fmaf(a1, s, a0 * c);
                o->R[3 * r + 1] = fmaf(a1, c, -(a0 * s));
                o->R[3 * r + 2] = A[3 * r + 2];
            }
        } else {
            memcpy(o->R, A, sizeof A);
        }
    }
}

static inline void sphere_centre(const Robot *rb, const Frame *fr, int i, float *C) {
    int link = rb->sph_link[i];
    const float *c = rb->sph[i];
    if (link < 0) { /* base frame = identity at the origin: centre is the table value */
        C[0] = c[0]; C[1] = c[1]; C[2] = c[2];
        return;
    }
    const Frame *f = &fr[link];
    for (int r = 0; r < 3; ++r)
        C[r] = fmaf(f->R[3 * r + 2], c[2], fmaf(f->R[3 * r + 1], c[1], fmaf(f->R[3 * r + 0], c[0], f->p[r])));
}

/* ------------------------------------------------------------------ pair distances (contract order) */
static inline float sd_box(const float *C, const float *b /* cx cy cz hx hy hz */, float r) {
    float dx = C[0] - b[0], dy = C[1] - b[1], dz = C[2] - b[2];
    float ax = fabsf(dx) - b[3], ay = fabsf(dy) - b[4], az = fabsf(dz) - b[5];
    float qx = fmaxf(ax, 0.f), qy = fmaxf(ay, 0.f), qz = fmaxf(az, 0.f);
    float dsq = fmaf(qz, qz, fmaf(qy, qy, qx * qx));
    float mi = fminf(fmaxf(ax, fmaxf(ay, az)), 0.f);
    return (sqrtf(dsq) + mi) - r;
}

static inline void normal_box(const float *C, const float *b, float *n) {
    float dx = C[0] - b[0], dy = C[1] - b[1], dz = C[2] - b[2];
    float ax = fabsf(dx) - b[3], ay = fabsf(dy) - b[4], az = fabsf(dz) - b[5];
    float qx = fmaxf(ax, 0.f), qy = fmaxf(ay, 0.f), qz = fmaxf(az, 0.f);
    float dsq = fmaf(qz, qz, fmaf(qy, qy, qx * qx));
    if (dsq > 0.f) {
        float d = sqrtf(dsq);
        n[0] = copysignf(qx, dx) / d;
        n[1] = copysignf(qy, dy) / d;
        n[2] = copysignf(qz, dz) / d;
    } else {
        n[0] = n[1] = n[2] = 0.f;
        if (ax >= ay && ax >= az) n[0] = copysignf(1.f, dx);
        else if (ay >= az) n[1] = copysignf(1.f, dy);
        else n[2] = copysignf(1.f, dz);
    }
}

static inline float sd_sphere(const float *C, const float *s /* cx cy cz r */, float r) {
    float dx = C[0] - s[0], dy = C[1] - s[1], dz = C[2] - s[2];
    float dsq = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    return (sqrtf(dsq) - s[3]) - r;
}

static inline void normal_sphere(const float *C, const float *s, float *n) {
    float dx = C[0] - s[0], dy = C[1] - s[1], dz = C[2] - s[2];
    float dsq = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    if (dsq > 0.f) {
        float d = sqrtf(dsq);
        n[0] = dx / d; n[1] = dy / d; n[2] = dz / d;
    } else {
        n[0] = 0.f; n[1] = 0.f; n[2] = 1.f;
    }
}

typedef struct {
    int n_boxes, n_spheres, floor;
    const float *boxes;   /* [n_boxes, 8]: cx cy cz hx hy hz 0 0  (device layout, see include/cbfrrt.h) */
    const float *spheres; /* [n_spheres, 4]: cx cy cz r */
} Scene;

typedef struct {
    uint8_t safe, unsafe;
    float o, self_min, soft_min, hard_min;
    int arg_link, arg_obs, arg_sphere;
    float dodq[MAXN];
} CollideOut;

/* One state: masks + observation + gradient.  Iteration order (sphere-major; floor, boxes, spheres) with
 * strict '<' defines the arg-min tie-break: lowest (sphere, obstacle) index wins. */
static void collide_state(const Robot *rb, const Scene *sc, const float *q, float thr_soft, float thr_hard,
                          CollideOut *out) {
    Frame fr[MAXLINKS];
    float C[MAXSPH][3];
    fk_chain(rb, q, fr);
    for (int i = 0; i < rb->n_spheres; ++i) sphere_centre(rb, fr, i, C[i]);

    const int off = sc->floor ? 1 : 0;
    float o = INFINITY, soft = INFINITY, hard = INFINITY;
    int arg_s = -1, arg_o = -1;
    int prev_link = -2;
    for (int i = 0; i < rb->n_spheres; ++i) {
        int link = rb->sph_link[i];
        float r = rb->sph[i][3];
        int first_of_link = (link != prev_link);
        prev_link = link;
        if (sc->floor) {
            int have = 1;
            float d;
            if (link == 0) { /* kinematic constant, attached to the link's first sphere (spec floor_const) */
                have = first_of_link;
                d = rb->link0_floor;
            } else {
                d = C[i][2] - r;
            }
            if (have) {
                if (link >= 0 && d < o) { o = d; arg_s = i; arg_o = 0; }
                if (!soft_floor_exempt(link) && d < soft) soft = d;
                if (!hard_floor_exempt(link) && d < hard) hard = d;
            }
        }
        for (int j = 0; j < sc->n_boxes; ++j) {
            float d = sd_box(C[i], sc->boxes + 8 * j, r);
            if (link >= 0 && d < o) { o = d; arg_s = i; arg_o = off + j; }
            if (d < soft) soft = d;
            if (d < hard) hard = d;
        }
        for (int j = 0; j < sc->n_spheres; ++j) {
            float d = sd_sphere(C[i], sc->spheres + 4 * j, r);
            if (link >= 0 && d < o) { o = d; arg_s = i; arg_o = off + sc->n_boxes + j; }
            if (d < soft) soft = d;
            if (d < hard) hard = d;
        }
    }
    /* self collision, basic_robot.py:89-98 (threshold 0.01) */
    float self_min = INFINITY;
    for (int p = 0; p < rb->n_self; ++p) {
        int a = rb->self_pairs[p][0], b = rb->self_pairs[p][1];
        for (int i = 0; i < rb->n_spheres; ++i) {
            if (rb->sph_link[i] != a) continue;
            for (int j = 0; j < rb->n_spheres; ++j) {
                if (rb->sph_link[j] != b) continue;
                float dx = C[i][0] - C[j][0], dy = C[i][1] - C[j][1], dz = C[i][2] - C[j][2];
                float dsq = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                float d = (sqrtf(dsq) - rb->sph[i][3]) - rb->sph[j][3];
                if (d < self_min) self_min = d;
            }
        }
    }
    out->unsafe = (uint8_t)(hard <= thr_hard);
    out->safe = (uint8_t)((self_min > 0.01f) && (soft > thr_soft) && !(hard <= thr_hard));
    out->o = o;
    out->self_min = self_min;
    out->soft_min = soft;
    out->hard_min = hard;
    out->arg_sphere = arg_s;
    out->arg_obs = arg_o;
    out->arg_link = arg_s >= 0 ? rb->sph_link[arg_s] : -2;
    for (int k = 0; k < rb->n; ++k) out->dodq[k] = 0.f;
    if (arg_s < 0) return;
    /* gradient: n^T J_lin(link, closest point)[:, body_joints] (arm_mindis.py:56-69); for a sphere the r*n
     * offset of the surface point is orthogonal to z x n, so the centre gives the same value. */
    float nrm[3];
    if (sc->floor && arg_o == 0) { nrm[0] = 0.f; nrm[1] = 0.f; nrm[2] = 1.f; }
    else if (arg_o - off < sc->n_boxes) normal_box(C[arg_s], sc->boxes + 8 * (arg_o - off), nrm);
    else normal_sphere(C[arg_s], sc->spheres + 4 * (arg_o - off - sc->n_boxes), nrm);
    for (int l = out->arg_link; l >= 0; l = rb->parent[l]) {
        int qi = rb->qidx[l];
        if (qi < 0) continue;
        const Frame *f = &fr[l];
        float zx = f->R[2], zy = f->R[5], zz = f->R[8];
        float rx = C[arg_s][0] - f->p[0], ry = C[arg_s][1] - f->p[1], rz = C[arg_s][2] - f->p[2];
        float cx = fmaf(zy, rz, -(zz * ry));
        float cy = fmaf(zz, rx, -(zx * rz));
        float cz = fmaf(zx, ry, -(zy * rx));
        out->dodq[qi] = fmaf(nrm[2], cz, fmaf(nrm[1], cy, nrm[0] * cx));
    }
}

/* ------------------------------------------------------------------ barrier MLP + Jacobian (fp32, tolerance-checked)
 * Packed weights (include/cbfrrt.h "CBF weight packing"): IN_PAD = round_up(n+1,4), H hidden, L hidden->hidden
 * layers: [W_in H x IN_PAD][b_in H] { [W_l H x H][b_l H] } x L [w_out H][b_out 1 (+3 pad)] */
typedef struct {
    int n_in, hidden, layers;
    const float *w;
} Net;

#define MAXH 128
#define MAXL 4

static void mlp_h_jac(const Net *net, const float *x /* n_in */, float *h_out, float *jac /* n_in */) {
    const int H = net->hidden, L = net->layers, IN = net->n_in, INP = (IN + 3) & ~3;
    float act[MAXL + 1][MAXH];
    const float *w = net->w;
    const float *W = w, *b = w + H * INP;
    for (int j = 0; j < H; ++j) {
        float a = b[j];
        for (int k = 0; k < IN; ++k) a = fmaf(W[j * INP + k], x[k], a);
        act[0][j] = a > 0.f ? a : 0.f;
    }
    const float *wl = w + H * INP + H;
    for (int l = 0; l < L; ++l) {
        const float *Wl = wl + (size_t)l * (H * H + H), *bl = Wl + H * H;
        for (int j = 0; j < H; ++j) {
            float a = bl[j];
            for (int k = 0; k < H; ++k) a = fmaf(Wl[j * H + k], act[l][k], a);
            act[l + 1][j] = a > 0.f ? a : 0.f;
        }
    }
    const float *wo = wl + (size_t)L * (H * H + H), *bo = wo + H;
    float h = bo[0];
    for (int k = 0; k < H; ++k) h = fmaf(wo[k], act[L][k], h);
    *h_out = h;
    /* reverse mode with the ReLU masks (== autograd, neural_mindis_cbf_controller.py:93-96) */
    float g[MAXH], gp[MAXH];
    for (int k = 0; k < H; ++k) g[k] = act[L][k] > 0.f ? wo[k] : 0.f;
    for (int l = L - 1; l >= 0; --l) {
        const float *Wl = wl + (size_t)l * (H * H + H);
        for (int k = 0; k < H; ++k) gp[k] = 0.f;
        for (int j = 0; j < H; ++j)
            for (int k = 0; k < H; ++k) gp[k] = fmaf(Wl[j * H + k], g[j], gp[k]);
        for (int k = 0; k < H; ++k) g[k] = act[l][k] > 0.f ? gp[k] : 0.f;
    }
    for (int k = 0; k < IN; ++k) {
        float a = 0.f;
        for (int j = 0; j < H; ++j) a = fmaf(W[j * INP + k], g[j], a);
        jac[k] = a;
    }
}

/* ------------------------------------------------------------------ exact single-constraint CBF-QP (double)
 *   min ||u-u_ref||^2 + p r   s.t.  c + a.u - r <= 0, r >= 0, lo <= u <= hi      (clf_controller.py:82-139)
 * KKT: u(mu) = clip(u_ref - mu a / 2, lo, hi), 0 <= mu <= p; g(mu) = c + a.u(mu) is non-increasing and
 * piecewise linear.  mu = 0 if g(0) <= 0; else the root of g on (0,p]; if g(p) > 0: mu = p, r = g(p). */
static double qp_g(int n, const double *a, const double *uref, const double *lo, const double *hi, double c,
                   double mu, double *u) {
    double g = c;
    for (int i = 0; i < n; ++i) {
        double v = uref[i] - 0.5 * mu * a[i];
        v = v < lo[i] ? lo[i] : (v > hi[i] ? hi[i] : v);
        if (u) u[i] = v;
        g += a[i] * v;
    }
    return g;
}

/* multiplier of one relaxed constraint for the projection target `t` (may lie outside the box): 2n breakpoints */
static double qp_mu_exact(int n, const double *a, const double *t, const double *lo, const double *hi, double c, double p) {
    double g0 = qp_g(n, a, t, lo, hi, c, 0.0, 0);
    if (!(g0 > 0.0)) return 0.0;
    double bp[2 * MAXN + 2];
    int nb = 0;
    bp[nb++] = 0.0;
    for (int i = 0; i < n; ++i) {
        if (a[i] == 0.0) continue;
        double m1 = 2.0 * (t[i] - lo[i]) / a[i], m2 = 2.0 * (t[i] - hi[i]) / a[i];
        if (m1 > 0.0 && m1 < p) bp[nb++] = m1;
        if (m2 > 0.0 && m2 < p) bp[nb++] = m2;
    }
    bp[nb++] = p;
    for (int i = 1; i < nb; ++i) { /* insertion sort */
        double v = bp[i];
        int j = i - 1;
        while (j >= 0 && bp[j] > v) { bp[j + 1] = bp[j]; --j; }
        bp[j + 1] = v;
    }
    double mu_lo = 0.0, g_lo = g0;
    for (int k = 1; k < nb; ++k) {
        double g_hi = qp_g(n, a, t, lo, hi, c, bp[k], 0);
        if (g_hi <= 0.0) return (g_lo > g_hi) ? mu_lo + (bp[k] - mu_lo) * (g_lo / (g_lo - g_hi)) : bp[k];
        mu_lo = bp[k];
        g_lo = g_hi;
    }
    return p; /* g(p) > 0: constraint relaxed */
}

static void qp_exact(int n, const double *a, const double *uref, const double *lo, const double *hi, double c,
                     double p, double *u, double *r_out) {
    double mu = qp_mu_exact(n, a, uref, lo, hi, c, p);
    double g = qp_g(n, a, uref, lo, hi, c, mu, u);
    *r_out = (mu >= p && g > 0.0) ? g : 0.0;
}

/* S relaxed constraints: maximise the concave piecewise-quadratic dual over mu in [0,p]^S (float64).
 * Every iteration = one cyclic sweep of exact coordinate maximisations (each a single-constraint root find on the
 * target shifted by the other multipliers) followed by one Newton step on the current active set (free multipliers
 * W, unclipped coordinates F:  (A_WF A_WF^T / 2 + eps I) d = g_W ) with an EXACT line search -- along d the
 * directional derivative is again c' + w.clip(t - alpha w / 2) with w = A^T d, i.e. the same root find.  Both moves
 * are monotone in the dual, so the iteration inherits the convergence of coordinate ascent and the finite
 * termination of Newton once the active set is identified.  Returns the iterations used (max_iter = not converged:
 * only seen on infeasible systems with a huge penalty, where mu creeps along a flat dual ridge). */
#define MAXS 8
#define KINK_TOL 1e-9 /* coordinates sitting on a kink of the dual count as free: the step must not leave the kink */
#ifndef NEWTON_REPS
#define NEWTON_REPS 6
#endif
static int qp_multi_exact(int n, int S, const double *a /* [S][n] */, const double *uref, const double *lo,
                          const double *hi, const double *c, double p, int max_iter, double tol, double *u, double *r,
                          double *mu_out) {
    double mu[MAXS] = {0}, t[MAXN], ti[MAXN];
    for (int k = 0; k < n; ++k) t[k] = uref[k];
    int it = 0;
    while (it < max_iter) {
        ++it;
        double moved = 0.0, shift = 0.0; /* control-space movement of this iteration, size of the dual shift of t */
        for (int i = 0; i < S; ++i) {
            double amax_i = 0.0;
            for (int k = 0; k < n; ++k) {
                ti[k] = t[k] + 0.5 * mu[i] * a[i * n + k];
                if (fabs(a[i * n + k]) > amax_i) amax_i = fabs(a[i * n + k]);
            }
            double m = qp_mu_exact(n, a + i * n, ti, lo, hi, c[i], p);
            double mv = fabs(m - mu[i]) * 0.5 * amax_i;
            if (mv > moved) moved = mv;
            if (0.5 * m * amax_i > shift) shift = 0.5 * m * amax_i;
            mu[i] = m;
            for (int k = 0; k < n; ++k) t[k] = ti[k] - 0.5 * m * a[i * n + k];
        }
        const double stop = tol + 4e-15 * shift; /* t = u_ref - A^T mu / 2 carries rounding of size eps * shift */
        if (moved <= stop) break;
        if (S < 2) continue;
        /* Newton steps on the active set (a step that ends on a kink changes the set: try again, a few times) */
        for (int rep = 0; rep < NEWTON_REPS; ++rep) {
        double g[MAXS], d[MAXS], M[MAXS][MAXS], w[MAXN];
        int inW[MAXS], nW = 0;
        for (int i = 0; i < S; ++i) {
            g[i] = c[i];
            for (int k = 0; k < n; ++k) g[i] += a[i * n + k] * (t[k] < lo[k] ? lo[k] : (t[k] > hi[k] ? hi[k] : t[k]));
            inW[i] = (mu[i] > 0.0 && mu[i] < p) || (mu[i] <= 0.0 && g[i] > 0.0) || (mu[i] >= p && g[i] < 0.0);
            nW += inW[i];
        }
        if (nW < 2) break;
        double tr = 0.0;
        for (int i = 0; i < S; ++i)
            for (int j = 0; j < S; ++j) {
                double acc = 0.0;
                if (inW[i] && inW[j])
                    for (int k = 0; k < n; ++k)
                        if (t[k] >= lo[k] - KINK_TOL && t[k] <= hi[k] + KINK_TOL) acc += a[i * n + k] * a[j * n + k];
                M[i][j] = (inW[i] && inW[j]) ? 0.5 * acc : (i == j ? 1.0 : 0.0);
                if (i == j && inW[i]) tr += M[i][i];
            }
        for (int i = 0; i < S; ++i) {
            if (inW[i]) M[i][i] += 1e-9 * tr + 1e-300;
            d[i] = inW[i] ? g[i] : 0.0;
        }
        int ok = 1;
        for (int i = 0; i < S && ok; ++i) { /* Gaussian elimination, SPD */
            if (!(M[i][i] > 0.0)) { ok = 0; break; }
            for (int j = i + 1; j < S; ++j) {
                double f = M[j][i] / M[i][i];
                for (int k = i; k < S; ++k) M[j][k] -= f * M[i][k];
                d[j] -= f * d[i];
            }
        }
        if (!ok) break;
        for (int i = S - 1; i >= 0; --i) {
            for (int j = i + 1; j < S; ++j) d[i] -= M[i][j] * d[j];
            d[i] /= M[i][i];
        }
        double amax = 1e300, cc = 0.0;
        int any = 0;
        for (int i = 0; i < S; ++i) {
            if ((mu[i] <= 0.0 && d[i] < 0.0) || (mu[i] >= p && d[i] > 0.0)) d[i] = 0.0;
            if (d[i] > 0.0) { double v = (p - mu[i]) / d[i]; if (v < amax) amax = v; any = 1; }
            if (d[i] < 0.0) { double v = -mu[i] / d[i]; if (v < amax) amax = v; any = 1; }
            cc += d[i] * c[i];
        }
        if (!any) break;
        double wmax = 0.0;
        for (int k = 0; k < n; ++k) {
            w[k] = 0.0;
            for (int i = 0; i < S; ++i) w[k] += a[i * n + k] * d[i];
            if (fabs(w[k]) > wmax) wmax = fabs(w[k]);
        }
        double al = qp_mu_exact(n, w, t, lo, hi, cc, amax);
        for (int i = 0; i < S; ++i) {
            mu[i] += al * d[i];
            mu[i] = mu[i] < 0.0 ? 0.0 : (mu[i] > p ? p : mu[i]);
        }
        for (int k = 0; k < n; ++k) t[k] -= 0.5 * al * w[k];
        if (al < amax && !(al * wmax > stop)) break; /* a step cut short by a multiplier bound changed the active set */
        }
    }
    for (int k = 0; k < n; ++k) u[k] = t[k] < lo[k] ? lo[k] : (t[k] > hi[k] ? hi[k] : t[k]);
    for (int i = 0; i < S; ++i) {
        double g = c[i];
        for (int k = 0; k < n; ++k) g += a[i * n + k] * u[k];
        r[i] = g > 0.0 ? g : 0.0; /* optimal relaxation for this u */
        if (mu_out) mu_out[i] = mu[i];
    }
    return it;
}

/* ------------------------------------------------------------------ exported batch API (ctypes) */
/* ------------------------------------------------------------------ tiny pthread parallel-for (no libgomp here) */
#include <pthread.h>
static int g_threads = 1;
typedef void (*range_fn)(long lo, long hi, void *ctx);
typedef struct { range_fn fn; void *ctx; long lo, hi; } Job;
static void *job_main(void *p) { Job *j = (Job *)p; j->fn(j->lo, j->hi, j->ctx); return 0; }
static void par_for(long B, range_fn fn, void *ctx) {
    int T = g_threads;
    if (T > B) T = (int)(B > 0 ? B : 1);
    if (T <= 1) { fn(0, B, ctx); return; }
    pthread_t th[256]; Job jobs[256];
    if (T > 256) T = 256;
    long chunk = (B + T - 1) / T;
    for (int t = 0; t < T; ++t) {
        long lo = t * chunk, hi = lo + chunk < B ? lo + chunk : B;
        if (lo > hi) lo = hi;
        jobs[t] = (Job){fn, ctx, lo, hi};
        pthread_create(&th[t], 0, job_main, &jobs[t]);
    }
    for (int t = 0; t < T; ++t) pthread_join(th[t], 0);
}
/* the range workers are cloned for FMA-capable CPUs (fmaf -> vfmadd; same correctly-rounded result) */
#define CLONES __attribute__((target_clones("default", "fma"), flatten))
#define API __attribute__((visibility("default")))
API void oracle_set_threads(int t) { g_threads = t < 1 ? 1 : t; }

/* reverse mode of the S-constraint QP (double).  Active sets from the MULTIPLIERS of the converged dual solve
 * (inactive mu_i = 0, tight 0 < mu_i < p, relaxed mu_i = p) and from the unclipped point v = u_ref - A^T mu / 2
 * (free coordinates lo < v < hi).  Adjoint of the KKT system on that set:
 *     stationarity  u_F + A_WF^T mu_W / 2 = u_ref_F - p sum_P a_i,F / 2        (adjoint variable  v_F)
 *     tightness     A_W u + c_W = 0                                            (adjoint variable  lambda)
 *  => (A_WF A_WF^T) lambda = -A_WF w_F,  w_F = gu_F + sum_P gr_i a_i,F,  v_F = w_F + A_WF^T lambda.
 * margin: distance of the row from the nearest change of active set. */
API int oracle_qp_multi_backward(int n, int S, const float *a, const float *c, const float *u_ref, long B,
                                 const float *u_lo, const float *u_hi, float penalty, const float *grad_u,
                                 const float *grad_r, float *grad_a, float *grad_c, float *grad_u_ref, float *margin) {
    if (n > MAXN || S > MAXS) return -1;
    const double p = penalty < 1e6f ? (double)penalty : 1e6;
    for (long b = 0; b < B; ++b) {
        double ad[MAXS * MAXN], cd[MAXS], ur[MAXN], lo[MAXN], hi[MAXN], u[MAXN], r[MAXS], mu[MAXS], gu[MAXN], gr[MAXS];
        for (int k = 0; k < S * n; ++k) ad[k] = a[b * S * n + k];
        for (int i = 0; i < S; ++i) { cd[i] = c[b * S + i]; gr[i] = grad_r ? (double)grad_r[b * S + i] : 0.0; }
        for (int k = 0; k < n; ++k) { ur[k] = u_ref[b * n + k]; lo[k] = u_lo[k]; hi[k] = u_hi[k]; gu[k] = grad_u[b * n + k]; }
        qp_multi_exact(n, S, ad, ur, lo, hi, cd, p, 5000, 1e-13, u, r, mu);
        int fr[MAXN], W[MAXS], nW = 0, P[MAXS];
        double w[MAXN], mg = 1e30;
        for (int k = 0; k < n; ++k) {
            double v = ur[k];
            for (int i = 0; i < S; ++i) v -= 0.5 * mu[i] * ad[i * n + k];
            fr[k] = v > lo[k] && v < hi[k];
            w[k] = fr[k] ? gu[k] : 0.0;
            if (fabs(v - lo[k]) < mg) mg = fabs(v - lo[k]);
            if (fabs(v - hi[k]) < mg) mg = fabs(v - hi[k]);
        }
        for (int i = 0; i < S; ++i) {
            double g = cd[i];
            for (int k = 0; k < n; ++k) g += ad[i * n + k] * u[k];
            P[i] = mu[i] >= p;
            if (mu[i] > 0.0 && mu[i] < p) { W[nW++] = i; if (mu[i] < mg) mg = mu[i]; if (p - mu[i] < mg) mg = p - mu[i]; }
            else if (fabs(g) < mg) mg = fabs(g);
            if (P[i])
                for (int k = 0; k < n; ++k) if (fr[k]) w[k] += gr[i] * ad[i * n + k];
        }
        if (margin) margin[b] = (float)mg;
        double M[MAXS][MAXS], lam[MAXS];
        for (int x = 0; x < nW; ++x) {
            lam[x] = 0.0;
            for (int k = 0; k < n; ++k) if (fr[k]) lam[x] -= ad[W[x] * n + k] * w[k];
            for (int y = 0; y < nW; ++y) {
                M[x][y] = 0.0;
                for (int k = 0; k < n; ++k) if (fr[k]) M[x][y] += ad[W[x] * n + k] * ad[W[y] * n + k];
            }
        }
        for (int x = 0; x < nW; ++x) { /* elimination without pivoting: M is a Gram matrix */
            if (!(M[x][x] > 0.0)) M[x][x] = 1e-300;
            for (int y = x + 1; y < nW; ++y) {
                double f = M[y][x] / M[x][x];
                for (int z = x; z < nW; ++z) M[y][z] -= f * M[x][z];
                lam[y] -= f * lam[x];
            }
        }
        for (int x = nW - 1; x >= 0; --x) {
            for (int y = x + 1; y < nW; ++y) lam[x] -= M[x][y] * lam[y];
            lam[x] /= M[x][x];
        }
        double v[MAXN];
        for (int k = 0; k < n; ++k) {
            v[k] = w[k];
            if (fr[k]) for (int x = 0; x < nW; ++x) v[k] += ad[W[x] * n + k] * lam[x];
            if (grad_u_ref) grad_u_ref[b * n + k] = (float)v[k];
        }
        for (int i = 0; i < S; ++i) {
            double head = 0.0, half = 0.0;
            for (int x = 0; x < nW; ++x) if (W[x] == i) { head = lam[x]; half = 0.5 * mu[i]; }
            if (P[i]) { head = r[i] > 0.0 ? gr[i] : 0.0; half = 0.5 * p; }
            if (grad_c) grad_c[b * S + i] = (float)head;
            for (int k = 0; k < n; ++k)
                if (grad_a) grad_a[(b * S + i) * n + k] = (float)(head * u[k] - half * v[k]);
        }
    }
    return 0;
}

API int oracle_robot_info(int robot, int *n, int *n_links, int *n_spheres) {
    if (robot < 0 || robot > 1) return -1;
    *n = ROBOTS[robot].n; *n_links = ROBOTS[robot].n_links; *n_spheres = ROBOTS[robot].n_spheres;
    return 0;
}

API void oracle_sincos(const float *x, long B, float *s, float *c) {
    for (long i = 0; i < B; ++i) det_sincosf(x[i], &s[i], &c[i]);
}

/* one context struct for every batch entry point (unused fields stay 0) */
typedef struct {
    const Robot *rb; Scene sc; Net net; int n;
    const float *q, *q2, *goal, *K, *u_lo, *u_hi, *datax_in, *u_ref;
    long q_stride, goal_rows;
    float thr_soft, thr_hard, lam, penalty, dt;
    int Kint, steps, ctrl_every, stuck_lag, stuck_after, keep_going, stuck_mode;
    const int *Kvar;
    float stuck_eps, *traj;
    uint8_t *safe, *unsafe, *alive;
    float *datax, *h, *J, *u, *r, *mins, *pos, *rot, *q_final;
    int *arg_link, *arg_obs, *first_bad, *steps_done;
} Ctx;

/* FK: pos [B, L, 3], rot [B, L, 9] for L = n_links pybullet links (basic_robot.py:55-62) */
CLONES static void fk_range(long lo, long hi, void *p) {
    Ctx *c = (Ctx *)p; const Robot *rb = c->rb;
    for (long b = lo; b < hi; ++b) {
        Frame fr[MAXLINKS];
        fk_chain(rb, c->q + b * c->q_stride, fr);
        for (int l = 0; l < rb->n_links; ++l) {
            memcpy(c->pos + (b * rb->n_links + l) * 3, fr[l].p, 12);
            memcpy(c->rot + (b * rb->n_links + l) * 9, fr[l].R, 36);
        }
    }
}
API int oracle_fk(int robot, const float *q, long q_stride, long B, float *pos, float *rot) {
    Ctx c; memset(&c, 0, sizeof c);
    c.rb = &ROBOTS[robot]; c.q = q; c.q_stride = q_stride; c.pos = pos; c.rot = rot;
    par_for(B, fk_range, &c);
    return 0;
}

/* collide: masks (arm_dynamics.py:189-232), observation + gradient packed as datax = [q | o | do/dq]
 * (arm_dynamics.py:267-279), plus the arg-min pair and the three minima for diagnostics. */
CLONES static void collide_range(long lo, long hi, void *p) {
    Ctx *c = (Ctx *)p; const int n = c->n;
    for (long b = lo; b < hi; ++b) {
        CollideOut co;
        const float *q = c->q + b * c->q_stride;
        collide_state(c->rb, &c->sc, q, c->thr_soft, c->thr_hard, &co);
        if (c->safe) c->safe[b] = co.safe;
        if (c->unsafe) c->unsafe[b] = co.unsafe;
        if (c->datax) {
            float *d = c->datax + b * (2 * n + 1);
            for (int k = 0; k < n; ++k) d[k] = q[k];
            d[n] = co.o;
            for (int k = 0; k < n; ++k) d[n + 1 + k] = co.dodq[k];
        }
        if (c->arg_link) c->arg_link[b] = co.arg_link;
        if (c->arg_obs) c->arg_obs[b] = co.arg_obs;
        if (c->mins) { c->mins[3 * b] = co.self_min; c->mins[3 * b + 1] = co.soft_min; c->mins[3 * b + 2] = co.hard_min; }
    }
}
API int oracle_collide(int robot, int n_boxes, const float *boxes, int n_sph, const float *spheres, int floor,
                       const float *q, long q_stride, long B, float thr_soft, float thr_hard, uint8_t *safe,
                       uint8_t *unsafe, float *datax, int *arg_link, int *arg_obs, float *mins /* [B,3] or NULL */) {
    Ctx c; memset(&c, 0, sizeof c);
    c.rb = &ROBOTS[robot]; c.n = c.rb->n; c.sc = (Scene){n_boxes, n_sph, floor, boxes, spheres};
    c.q = q; c.q_stride = q_stride; c.thr_soft = thr_soft; c.thr_hard = thr_hard;
    c.safe = safe; c.unsafe = unsafe; c.datax = datax; c.arg_link = arg_link; c.arg_obs = arg_obs; c.mins = mins;
    par_for(B, collide_range, &c);
    return 0;
}

/* cbf_filter: datax [B, 2n+1] -> h, J (bs x n), u, r    (controller.u)
 *   goal [goal_rows, n] with goal_rows in {1, B}; K [n, n]; u_lo/u_hi [n] */
static void filter_state(int n, const Net *net, const float *d, const float *goal, const float *uref_in, const float *K,
                         const float *u_lo, const float *u_hi, float lam, float penalty, float *h_out, float *J,
                         float *u, float *r) {
    float x[MAXN + 1], jac[MAXN + 1], h;
    for (int k = 0; k <= n; ++k) x[k] = d[k];
    mlp_h_jac(net, x, &h, jac);
    /* J = Jh_q + Jh_o * do/dq   (neural_mindis_cbf_controller.py:99-101) */
    double a[MAXN], uref[MAXN], lo[MAXN], hi[MAXN], ud[MAXN], rd;
    for (int k = 0; k < n; ++k) {
        float j = fmaf(jac[n], d[n + 1 + k], jac[k]);
        J[k] = j;
        a[k] = (double)j; /* Lg_V = J g = J  (g = I), Lf_V = 0 */
    }
    /* u_ref = clamp(-K (q - goal), lo, hi)   (arm_dynamics.py:124-161) */
    for (int i = 0; i < n; ++i) {
        float v;
        if (uref_in) { /* solve_CLF_QP(u_ref=...) bypasses the nominal controller, clf_controller.py:497-504 */
            v = uref_in[i];
        } else {
            float acc = 0.f;
            for (int k = 0; k < n; ++k) acc = fmaf(K[i * n + k], d[k] - goal[k], acc);
            v = -acc;
            v = v < u_lo[i] ? u_lo[i] : (v > u_hi[i] ? u_hi[i] : v);
        }
        uref[i] = (double)v; lo[i] = (double)u_lo[i]; hi[i] = (double)u_hi[i];
    }
    double p = penalty < 1e6f ? (double)penalty : 1e6; /* clf_controller.py:380 */
    qp_exact(n, a, uref, lo, hi, (double)lam * (double)h, p, ud, &rd);
    *h_out = h;
    for (int k = 0; k < n; ++k) u[k] = (float)ud[k];
    *r = (float)rd;
}

CLONES static void filter_range(long lo, long hi, void *p) {
    Ctx *c = (Ctx *)p; const int n = c->n;
    for (long b = lo; b < hi; ++b) {
        const float *g = c->goal ? c->goal + (c->goal_rows == 1 ? 0 : b * n) : 0;
        filter_state(n, &c->net, c->datax_in + b * (2 * n + 1), g, c->u_ref ? c->u_ref + b * n : 0, c->K, c->u_lo,
                     c->u_hi, c->lam, c->penalty, c->h + b, c->J + b * n, c->u + b * n, c->r + b);
    }
}
API int oracle_cbf_filter(int n, int hidden, int layers, const float *weights, const float *datax, long B,
                          const float *goal, long goal_rows, const float *u_ref, const float *K, const float *u_lo,
                          const float *u_hi, float lam, float penalty, float *h, float *J, float *u, float *r) {
    if (hidden > MAXH || layers > MAXL || n > MAXN) return -1;
    Ctx c; memset(&c, 0, sizeof c);
    c.n = n; c.net = (Net){n + 1, hidden, layers, weights}; c.datax_in = datax; c.goal = goal; c.goal_rows = goal_rows;
    c.u_ref = u_ref; c.K = K; c.u_lo = u_lo; c.u_hi = u_hi; c.lam = lam; c.penalty = penalty; c.h = h; c.J = J; c.u = u; c.r = r;
    par_for(B, filter_range, &c);
    return 0;
}

/* ------------------------------------------------------------------ float64 END-TO-END truth of the filter stage
 * datax (the float32 inputs, taken as exact) -> h, J, u, r with the network, its reverse-mode Jacobian, J = Jh_q +
 * Jh_o do/dq, the nominal control and the exact KKT solution ALL in double.  Besides the values it returns what a
 * tolerance statement needs:
 *   margin [B]   smallest |pre-activation| over all hidden neurons: the distance (in pre-activation units) to the
 *                nearest ReLU kink.  J is DISCONTINUOUS across a kink, so two correct fp32 evaluations whose
 *                pre-activations differ by more than `margin` may return Jacobians of different linear pieces;
 *   kappa_c, kappa_a [B]   first-order conditioning of the QP solution:  |du|_inf <= kappa_c |dc| + kappa_a |da|_inf
 *                for perturbations of c = lambda h and a = J on the current active set (F = unclipped coordinates,
 *                S = |a_F|^2):  inactive (mu = 0): 0, 0;  relaxed (mu = p): 0, p/2;
 *                tight: kappa_c = |a_F|_inf / S,  kappa_a = (|a_F|_inf / S)(sum_F |t_k| + sum_C |b_k| + mu |a_F|_1) + mu/2;
 *   regime [B]   0 inactive, 1 tight, 2 relaxed.
 * (neural_mindis_cbf_controller.py:71-102, clf_controller.py:82-139,461-535, arm_dynamics.py:124-161) */
typedef struct {
    int n; Net net; const float *datax, *goal, *u_ref, *K, *u_lo, *u_hi; long goal_rows; double lam, p;
    double *h, *J, *u, *r, *margin, *kc, *ka; int *regime;
} Ctx64;
static void filter64_range(long lo_, long hi_, void *pp) {
    Ctx64 *c = (Ctx64 *)pp; const int n = c->n, H = c->net.hidden, L = c->net.layers, IN = n + 1, INP = (IN + 3) & ~3;
    const float *w = c->net.w;
    for (long b = lo_; b < hi_; ++b) {
        const float *d = c->datax + b * (2 * n + 1);
        double act[MAXL + 1][MAXH], margin = INFINITY;
        const float *W = w, *bi = w + H * INP;
        for (int j = 0; j < H; ++j) {
            double a = bi[j];
            for (int k = 0; k < IN; ++k) a += (double)W[j * INP + k] * (double)d[k];
            if (fabs(a) < margin) margin = fabs(a);
            act[0][j] = a > 0. ? a : 0.;
        }
        const float *wl = w + H * INP + H;
        for (int l = 0; l < L; ++l) {
            const float *Wl = wl + (size_t)l * (H * H + H), *bl = Wl + H * H;
            for (int j = 0; j < H; ++j) {
                double a = bl[j];
                for (int k = 0; k < H; ++k) a += (double)Wl[j * H + k] * act[l][k];
                if (fabs(a) < margin) margin = fabs(a);
                act[l + 1][j] = a > 0. ? a : 0.;
            }
        }
        const float *wo = wl + (size_t)L * (H * H + H), *bo = wo + H;
        double h = bo[0];
        for (int k = 0; k < H; ++k) h += (double)wo[k] * act[L][k];
        double g[MAXH], gp[MAXH], jac[MAXN + 1];
        for (int k = 0; k < H; ++k) g[k] = act[L][k] > 0. ? (double)wo[k] : 0.;
        for (int l = L - 1; l >= 0; --l) {
            const float *Wl = wl + (size_t)l * (H * H + H);
            for (int k = 0; k < H; ++k) gp[k] = 0.;
            for (int j = 0; j < H; ++j)
                for (int k = 0; k < H; ++k) gp[k] += (double)Wl[j * H + k] * g[j];
            for (int k = 0; k < H; ++k) g[k] = act[l][k] > 0. ? gp[k] : 0.;
        }
        for (int k = 0; k < IN; ++k) {
            double a = 0.;
            for (int j = 0; j < H; ++j) a += (double)W[j * INP + k] * g[j];
            jac[k] = a;
        }
        double a[MAXN], t[MAXN], lo[MAXN], hi[MAXN], u[MAXN], r;
        for (int k = 0; k < n; ++k) a[k] = jac[k] + jac[n] * (double)d[n + 1 + k];
        const float *goal = c->goal ? c->goal + (c->goal_rows == 1 ? 0 : b * n) : 0;
        for (int i = 0; i < n; ++i) {
            lo[i] = c->u_lo[i]; hi[i] = c->u_hi[i];
            if (c->u_ref) t[i] = c->u_ref[b * n + i];
            else {
                double acc = 0.;
                for (int k = 0; k < n; ++k) acc += (double)c->K[i * n + k] * ((double)d[k] - (double)goal[k]);
                t[i] = -acc < lo[i] ? lo[i] : (-acc > hi[i] ? hi[i] : -acc);
            }
        }
        const double cc = c->lam * h;
        const double mu = qp_mu_exact(n, a, t, lo, hi, cc, c->p);
        const double gg = qp_g(n, a, t, lo, hi, cc, mu, u);
        r = (mu >= c->p && gg > 0.) ? gg : 0.;
        double S = 0., ainf = 0., a1 = 0., st = 0.;
        for (int k = 0; k < n; ++k) {
            const double v = t[k] - 0.5 * mu * a[k];
            if (v > lo[k] && v < hi[k]) { S += a[k] * a[k]; a1 += fabs(a[k]); st += fabs(t[k]); if (fabs(a[k]) > ainf) ainf = fabs(a[k]); }
            else st += fabs(u[k]);
        }
        int regime = mu <= 0. ? 0 : (mu >= c->p ? 2 : 1);
        double kc = 0., ka = 0.;
        if (regime == 2) ka = 0.5 * c->p;
        else if (regime == 1 && S > 0.) { kc = ainf / S; ka = (ainf / S) * (st + mu * a1) + 0.5 * mu; }
        c->h[b] = h; c->r[b] = r; c->margin[b] = margin; c->kc[b] = kc; c->ka[b] = ka; c->regime[b] = regime;
        for (int k = 0; k < n; ++k) { c->J[b * n + k] = a[k]; c->u[b * n + k] = u[k]; }
    }
}
API int oracle_cbf_filter_f64(int n, int hidden, int layers, const float *weights, const float *datax, long B,
                              const float *goal, long goal_rows, const float *u_ref, const float *K, const float *u_lo,
                              const float *u_hi, float lam, float penalty, double *h, double *J, double *u, double *r,
                              double *margin, double *kappa_c, double *kappa_a, int *regime) {
    if (hidden > MAXH || layers > MAXL || n > MAXN) return -1;
    Ctx64 c; memset(&c, 0, sizeof c);
    c.n = n; c.net = (Net){n + 1, hidden, layers, weights}; c.datax = datax; c.goal = goal; c.goal_rows = goal_rows;
    c.u_ref = u_ref; c.K = K; c.u_lo = u_lo; c.u_hi = u_hi; c.lam = lam; c.p = penalty < 1e6f ? (double)penalty : 1e6;
    c.h = h; c.J = J; c.u = u; c.r = r; c.margin = margin; c.kc = kappa_c; c.ka = kappa_a; c.regime = regime;
    par_for(B, filter64_range, &c);
    return 0;
}

/* the QP stage alone on caller-supplied data (double precision): a = Lg_V [B,n], c = Lf_V + lambda*V [B].
 * Used to check the CUDA QP on exactly the (h, J) the CUDA MLP produced, which separates the QP's own error
 * from the amplification of MLP rounding differences by an ill-conditioned QP (tiny |a_i| on a free axis). */
API int oracle_qp(int n, const float *a, const float *c, const float *u_ref, long B, const float *u_lo,
                  const float *u_hi, float penalty, float *u, float *r) {
    if (n > MAXN) return -1;
    for (long b = 0; b < B; ++b) {
        double ad[MAXN], ur[MAXN], lo[MAXN], hi[MAXN], ud[MAXN], rd;
        for (int k = 0; k < n; ++k) { ad[k] = a[b * n + k]; ur[k] = u_ref[b * n + k]; lo[k] = u_lo[k]; hi[k] = u_hi[k]; }
        qp_exact(n, ad, ur, lo, hi, (double)c[b], penalty < 1e6f ? (double)penalty : 1e6, ud, &rd);
        for (int k = 0; k < n; ++k) u[b * n + k] = (float)ud[k];
        r[b] = (float)rd;
    }
    return 0;
}

/* reverse mode of the single-constraint QP (double): what backpropagation through the reference's CvxpyLayer yields
 * (clf_controller.py:82-139, solve_CLF_QP(requires_grad=True) :461-528), written from the KKT system on the active
 * set.  F = unclipped coordinates at the solution.
 *   inactive (mu = 0):    du_F = du_ref_F
 *   tight (0 < mu < p):   a.u + c = 0 and u_F = u_ref_F - mu a_F / 2  =>  differentiate both, eliminate dmu
 *   relaxed (mu = p):     u_F = u_ref_F - p a_F / 2,  r = c + a.u
 * margin [B] (optional): distance of the row from the nearest kink of the solution map (min over coordinates of the
 * distance of the unclipped value to its bounds, |g(0)|, |g(p)|, mu, p - mu) -- tests skip rows sitting on a kink. */
API int oracle_qp_backward(int n, const float *a, const float *c, const float *u_ref, long B, const float *u_lo,
                           const float *u_hi, float penalty, const float *grad_u, const float *grad_r, float *grad_a,
                           float *grad_c, float *grad_u_ref, float *margin) {
    if (n > MAXN) return -1;
    const double p = penalty < 1e6f ? (double)penalty : 1e6;
    for (long b = 0; b < B; ++b) {
        double ad[MAXN], ur[MAXN], lo[MAXN], hi[MAXN], u[MAXN], gu[MAXN];
        int fr[MAXN];
        for (int k = 0; k < n; ++k) {
            ad[k] = a[b * n + k]; ur[k] = u_ref[b * n + k]; lo[k] = u_lo[k]; hi[k] = u_hi[k]; gu[k] = grad_u[b * n + k];
        }
        const double cc = c[b], gr = grad_r ? (double)grad_r[b] : 0.0;
        const double mu = qp_mu_exact(n, ad, ur, lo, hi, cc, p);
        const double g = qp_g(n, ad, ur, lo, hi, cc, mu, u);
        double S = 0.0, G = 0.0, mg = 1e30;
        for (int k = 0; k < n; ++k) {
            const double v = ur[k] - 0.5 * mu * ad[k];
            fr[k] = v > lo[k] && v < hi[k];
            if (fr[k]) { S += ad[k] * ad[k]; G += gu[k] * ad[k]; }
            if (fabs(v - lo[k]) < mg) mg = fabs(v - lo[k]);
            if (fabs(v - hi[k]) < mg) mg = fabs(v - hi[k]);
        }
        const double g0 = qp_g(n, ad, ur, lo, hi, cc, 0.0, 0), gp = qp_g(n, ad, ur, lo, hi, cc, p, 0);
        if (fabs(g0) < mg) mg = fabs(g0);
        if (fabs(gp) < mg) mg = fabs(gp);
        if (margin) margin[b] = (float)mg;
        double ga[MAXN], gt[MAXN], gc;
        if (!(mu > 0.0)) {
            gc = 0.0;
            for (int k = 0; k < n; ++k) { ga[k] = 0.0; gt[k] = fr[k] ? gu[k] : 0.0; }
        } else if (mu >= p) {
            const double grr = g > 0.0 ? gr : 0.0;
            gc = grr;
            for (int k = 0; k < n; ++k) {
                const double gk = fr[k] ? gu[k] + grr * ad[k] : 0.0;
                gt[k] = gk;
                ga[k] = grr * u[k] - 0.5 * p * gk;
            }
        } else {
            /* unknowns (du_F, dmu):  du_F + a_F dmu / 2 = du_ref_F - mu da_F / 2 ;  a_F.du_F = -dc - u.da
             * adjoint: lambda solves the transposed system for the seed gu_F */
            const double lam = S > 0.0 ? G / S : 0.0;            /* multiplier of the tightness equation */
            gc = -lam;
            for (int k = 0; k < n; ++k) {
                const double w = fr[k] ? gu[k] - lam * ad[k] : 0.0; /* adjoint of the stationarity rows */
                gt[k] = w;
                ga[k] = -lam * u[k] - 0.5 * mu * w;
            }
        }
        for (int k = 0; k < n; ++k) {
            if (grad_a) grad_a[b * n + k] = (float)ga[k];
            if (grad_u_ref) grad_u_ref[b * n + k] = (float)gt[k];
        }
        if (grad_c) grad_c[b] = (float)gc;
    }
    return 0;
}

/* multi-scenario QP on caller-supplied data: a [B,S,n], c [B,S], u_ref [B,n] -> u [B,n], r [B,S] */
API int oracle_qp_multi(int n, int S, const float *a, const float *c, const float *u_ref, long B, const float *u_lo,
                        const float *u_hi, float penalty, int max_iter, double tol, float *u, float *r, int *iters) {
    if (n > MAXN || S > MAXS) return -1;
    for (long b = 0; b < B; ++b) {
        double ad[MAXS * MAXN], cd[MAXS], ur[MAXN], lo[MAXN], hi[MAXN], ud[MAXN], rd[MAXS];
        for (int k = 0; k < S * n; ++k) ad[k] = a[b * S * n + k];
        for (int i = 0; i < S; ++i) cd[i] = c[b * S + i];
        for (int k = 0; k < n; ++k) { ur[k] = u_ref[b * n + k]; lo[k] = u_lo[k]; hi[k] = u_hi[k]; }
        int it = qp_multi_exact(n, S, ad, ur, lo, hi, cd, penalty < 1e6f ? (double)penalty : 1e6, max_iter, tol, ud, rd, 0);
        if (iters) iters[b] = it;
        for (int k = 0; k < n; ++k) u[b * n + k] = (float)ud[k];
        for (int i = 0; i < S; ++i) r[b * S + i] = (float)rd[i];
    }
    return 0;
}

/* fused q -> everything (the "safety filter" of BASELINE configs 1, 2, 5) */
CLONES static void safety_range(long lo, long hi, void *p) {
    Ctx *c = (Ctx *)p; const int n = c->n;
    for (long b = lo; b < hi; ++b) {
        CollideOut co;
        const float *q = c->q + b * c->q_stride;
        collide_state(c->rb, &c->sc, q, c->thr_soft, c->thr_hard, &co);
        float d[2 * MAXN + 1];
        for (int k = 0; k < n; ++k) d[k] = q[k];
        d[n] = co.o;
        for (int k = 0; k < n; ++k) d[n + 1 + k] = co.dodq[k];
        if (c->safe) c->safe[b] = co.safe;
        if (c->unsafe) c->unsafe[b] = co.unsafe;
        if (c->datax) memcpy(c->datax + b * (2 * n + 1), d, sizeof(float) * (2 * n + 1));
        float hh, JJ[MAXN], uu[MAXN], rr;
        const float *g = c->goal + (c->goal_rows == 1 ? 0 : b * n);
        filter_state(n, &c->net, d, g, 0, c->K, c->u_lo, c->u_hi, c->lam, c->penalty, &hh, JJ, uu, &rr);
        if (c->h) c->h[b] = hh;
        if (c->J) memcpy(c->J + b * n, JJ, sizeof(float) * n);
        if (c->u) memcpy(c->u + b * n, uu, sizeof(float) * n);
        if (c->r) c->r[b] = rr;
    }
}
API int oracle_safety_filter(int robot, int n_boxes, const float *boxes, int n_sph, const float *spheres, int floor,
                             int hidden, int layers, const float *weights, const float *q, long q_stride, long B,
                             const float *goal, long goal_rows, const float *K, const float *u_lo, const float *u_hi,
                             float thr_soft, float thr_hard, float lam, float penalty, uint8_t *safe, uint8_t *unsafe,
                             float *datax, float *h, float *J, float *u, float *r) {
    if (hidden > MAXH || layers > MAXL) return -1;
    Ctx c; memset(&c, 0, sizeof c);
    c.rb = &ROBOTS[robot]; c.n = c.rb->n; c.sc = (Scene){n_boxes, n_sph, floor, boxes, spheres};
    c.net = (Net){c.n + 1, hidden, layers, weights};
    c.q = q; c.q_stride = q_stride; c.goal = goal; c.goal_rows = goal_rows; c.K = K; c.u_lo = u_lo; c.u_hi = u_hi;
    c.thr_soft = thr_soft; c.thr_hard = thr_hard; c.lam = lam; c.penalty = penalty;
    c.safe = safe; c.unsafe = unsafe; c.datax = datax; c.h = h; c.J = J; c.u = u; c.r = r;
    par_for(B, safety_range, &c);
    return 0;
}

/* edge validity, tsa.py:272-286 with a fixed number K of interpolation points:
 *   valid <=> safe(q_to) and safe(q_from + (k/K)(q_to - q_from)) for k = 0..K-1
 * first_bad = smallest along-edge index that fails (k for interpolation points, K for the end point), -1 if valid.
 * Interpolation contract: t = (float)k / (float)K (one division), c_i = fmaf(t, q_to_i - q_from_i, q_from_i). */
CLONES static void edge_range(long lo, long hi, void *p) {
    Ctx *c = (Ctx *)p; const int n = c->n;
    for (long e = lo; e < hi; ++e) {
        const int K = c->Kvar ? c->Kvar[e] : c->Kint; /* per-edge K (tsa.py:277-279) or one K for all */
        int bad = -1;
        for (int k = 0; k <= K && bad < 0; ++k) {
            float x[MAXN];
            if (k == K) {
                for (int i = 0; i < n; ++i) x[i] = c->q2[e * n + i];
            } else {
                float t = (float)k / (float)K;
                for (int i = 0; i < n; ++i) x[i] = fmaf(t, c->q2[e * n + i] - c->q[e * n + i], c->q[e * n + i]);
            }
            CollideOut co;
            collide_state(c->rb, &c->sc, x, c->thr_soft, c->thr_hard, &co);
            if (!co.safe) bad = k;
        }
        c->safe[e] = (uint8_t)(bad < 0);
        c->first_bad[e] = bad;
    }
}
API int oracle_edge_validity(int robot, int n_boxes, const float *boxes, int n_sph, const float *spheres, int floor,
                             const float *q_from, const float *q_to, long E, int K, float thr_soft, float thr_hard,
                             uint8_t *valid, int *first_bad) {
    Ctx c; memset(&c, 0, sizeof c);
    c.rb = &ROBOTS[robot]; c.n = c.rb->n; c.sc = (Scene){n_boxes, n_sph, floor, boxes, spheres};
    c.q = q_from; c.q2 = q_to; c.Kint = K; c.thr_soft = thr_soft; c.thr_hard = thr_hard; c.safe = valid;
    c.first_bad = first_bad;
    par_for(E, edge_range, &c);
    return 0;
}

API int oracle_edge_validity_var(int robot, int n_boxes, const float *boxes, int n_sph, const float *spheres, int floor,
                                 const float *q_from, const float *q_to, long E, const int *K, float thr_soft,
                                 float thr_hard, uint8_t *valid, int *first_bad) {
    Ctx c; memset(&c, 0, sizeof c);
    c.rb = &ROBOTS[robot]; c.n = c.rb->n; c.sc = (Scene){n_boxes, n_sph, floor, boxes, spheres};
    c.q = q_from; c.q2 = q_to; c.Kvar = K; c.thr_soft = thr_soft; c.thr_hard = thr_hard; c.safe = valid;
    c.first_bad = first_bad;
    par_for(E, edge_range, &c);
    return 0;
}

/* planner front-end (batch_tsa.py:135-145, tsa.py:151-153, bit_star.py:196-216): k nearest tree rows per query in
 * ascending (squared distance, row) order; contract: d_k = q_k - t_k, dsq = fma chain starting from d_0*d_0 */
static float knn_dsq(int n, const float *q, const float *t) {
    float d = q[0] - t[0];
    float acc = d * d;
    for (int k = 1; k < n; ++k) { d = q[k] - t[k]; acc = fmaf(d, d, acc); }
    return acc;
}
API int oracle_knn(int n, const float *queries, const float *tree, long M, long Nt, int k, int *idx, float *dist) {
    for (long m = 0; m < M; ++m) {
        float prev_d = -1.0f;
        long prev_i = -1;
        for (int r = 0; r < k; ++r) {
            float bd = 0.0f;
            long bi = -1;
            for (long j = 0; j < Nt; ++j) {
                float d = knn_dsq(n, queries + m * n, tree + j * n);
                int after_prev = (d > prev_d) || (d == prev_d && j > prev_i);
                if (after_prev && (bi < 0 || d < bd)) { bd = d; bi = j; }
            }
            idx[m * k + r] = (int)bi;
            dist[m * k + r] = bi < 0 ? INFINITY : sqrtf(bd);
            if (bi < 0) { prev_d = INFINITY; prev_i = Nt; } else { prev_d = bd; prev_i = bi; }
        }
    }
    return 0;
}
API int oracle_radius_mask(int n, const float *queries, const float *set, long M, long Ns, float radius, uint8_t *within) {
    for (long m = 0; m < M; ++m)
        for (long j = 0; j < Ns; ++j) within[m * Ns + j] = sqrtf(knn_dsq(n, queries + m * n, set + j * n)) <= radius;
    return 0;
}

/* closed-loop rollout = the steer loop (tsa.py:217-269; eval_obs_only_env.py:121-148):
 *   every ctrl_every-th step: u = CBF-QP(x);  every step: q' = q + u*dt (arm_dynamics.py:497-510), then the
 *   unsafe check on q' (tsa.py:251: a colliding state is NOT appended, no_collision = False, loop ends);
 *   the observation (o, do/dq) is refreshed when (t+1) % ctrl_every == 0 (tsa.py:242-245);
 *   stuck test (tsa.py:263): stop when t > stuck_after and ||x_t - x_{t-stuck_lag}|| < stuck_eps.
 * q_final = x_list[-1], alive = no_collision, steps_done = number of appended states, traj = x_list[1:].
 * stuck_mode 1 = the batched steer of batch_tsa.py:191-231 (see the branch below). */
CLONES static void rollout_range(long lo, long hi, void *p) {
    Ctx *c = (Ctx *)p; const int n = c->n;
    for (long b = lo; b < hi; ++b) {
        float d[2 * MAXN + 1], qn[MAXN], u[MAXN] = {0};
        CollideOut co;
        for (int k = 0; k < n; ++k) d[k] = c->q[b * n + k];
        collide_state(c->rb, &c->sc, d, c->thr_soft, c->thr_hard, &co);
        d[n] = co.o;
        for (int k = 0; k < n; ++k) d[n + 1 + k] = co.dodq[k];
        int ok = 1, done = 0, dead = 0;
        float *tr = c->traj ? c->traj + b * (long)c->steps * n : 0;
        for (int t = 0; t < c->steps; ++t) {
            if (t % c->ctrl_every == 0) {
                float hh, JJ[MAXN], rr;
                const float *g = c->goal + (c->goal_rows == 1 ? 0 : b * n);
                filter_state(n, &c->net, d, g, 0, c->K, c->u_lo, c->u_hi, c->lam, c->penalty, &hh, JJ, u, &rr);
            }
            for (int k = 0; k < n; ++k) qn[k] = fmaf(u[k], c->dt, d[k]);
            collide_state(c->rb, &c->sc, qn, c->thr_soft, c->thr_hard, &co);
            if ((t + 1) % c->ctrl_every == 0) {
                d[n] = co.o;
                for (int k = 0; k < n; ++k) d[n + 1 + k] = co.dodq[k];
            }
            if (c->stuck_mode == 1) {
                /* batched steer (batch_tsa.py:225-229): stuck test BEFORE the append on the stored list (more than
                 * stuck_after states: the last against the one stuck_lag before it); stuck OR unsafe kills the row for
                 * good (no_collisions[i] = False, nothing appended any more) while x_current keeps integrating */
                if (!dead) {
                    int stuck = 0;
                    if (c->stuck_lag > 0 && done > c->stuck_after) {
                        const float *a = tr + (done - 1) * n, *o2 = tr + (done - 1 - c->stuck_lag) * n;
                        float dsq = 0.f;
                        for (int k = 0; k < n; ++k) { float e = a[k] - o2[k]; dsq = fmaf(e, e, dsq); }
                        stuck = sqrtf(dsq) < c->stuck_eps;
                    }
                    if (stuck || co.unsafe) { ok = 0; dead = 1; }
                }
                for (int k = 0; k < n; ++k) d[k] = qn[k];
                if (!dead) {
                    if (tr) for (int k = 0; k < n; ++k) tr[done * n + k] = d[k];
                    ++done;
                } else if (!c->keep_going) break;
                continue;
            }
            if (co.unsafe) { ok = 0; if (!c->keep_going) break; }
            for (int k = 0; k < n; ++k) d[k] = qn[k];
            if (tr) for (int k = 0; k < n; ++k) tr[t * n + k] = d[k];
            done = t + 1;
            if (c->stuck_lag > 0 && t > c->stuck_after) {
                const float *old = tr + (t - c->stuck_lag) * n;
                float dsq = 0.f;
                for (int k = 0; k < n; ++k) { float e = d[k] - old[k]; dsq = fmaf(e, e, dsq); }
                if (sqrtf(dsq) < c->stuck_eps) break;
            }
        }
        for (int k = 0; k < n; ++k) c->q_final[b * n + k] = d[k];
        c->alive[b] = (uint8_t)ok;
        c->steps_done[b] = done;
    }
}
API int oracle_rollout(int robot, int n_boxes, const float *boxes, int n_sph, const float *spheres, int floor,
                       int hidden, int layers, const float *weights, const float *q0, const float *goal, long goal_rows,
                       long T, int steps, int ctrl_every, float dt, int stuck_lag, int stuck_after, float stuck_eps,
                       int keep_going, const float *K, const float *u_lo, const float *u_hi, float thr_soft, float thr_hard, float lam,
                       float penalty, float *q_final, uint8_t *alive, int *steps_done, float *traj, int stuck_mode) {
    if (hidden > MAXH || layers > MAXL) return -1;
    if (stuck_lag > 0 && !traj) return -1;
    Ctx c; memset(&c, 0, sizeof c);
    c.rb = &ROBOTS[robot]; c.n = c.rb->n; c.sc = (Scene){n_boxes, n_sph, floor, boxes, spheres};
    c.net = (Net){c.n + 1, hidden, layers, weights};
    c.q = q0; c.goal = goal; c.goal_rows = goal_rows; c.steps = steps; c.ctrl_every = ctrl_every; c.dt = dt;
    c.stuck_lag = stuck_lag; c.stuck_after = stuck_after; c.stuck_eps = stuck_eps; c.traj = traj; c.keep_going = keep_going; c.stuck_mode = stuck_mode;
    c.K = K; c.u_lo = u_lo; c.u_hi = u_hi; c.thr_soft = thr_soft; c.thr_hard = thr_hard; c.lam = lam; c.penalty = penalty;
    c.q_final = q_final; c.alive = alive; c.steps_done = steps_done;
    par_for(T, rollout_range, &c);
    return 0;
}

/* A whole expansion of the batched planner in one call: n_seg consecutive batched steer segments
 * (motion_planning/baseline/batch_tsa.py:154-170: `for _ in range(RRT_PARAM // 20): new_states, no_collisions, ... =
 * steer(..., current, RRT_step=20); current = new_states`), each with the semantics of batch_tsa.py:191-231:
 *   - a segment starts with a fresh observation of every row's current state (complete_sample_with_observations);
 *   - control every ctrl_every steps, Euler every step, observation refresh when (t+1) % ctrl_every == 0;
 *   - per row: stuck (more than stuck_after states stored: last vs the one stuck_lag before, < stuck_eps) OR unsafe new
 *     state -> no_collision = False for good, nothing appended any more, but the row keeps integrating;
 *   - the segment's loop ends early once EVERY row of the batch is dead (:229-230); x_current of that step is what the
 *     next segment starts from.
 * Outputs per segment: seg_q [T, n_seg, n] = x_current at the end of the segment, seg_alive [T, n_seg], seg_done
 * [T, n_seg] = stored states, traj [T, n_seg * seg_len, n] with segment k's stored states from row k * seg_len.
 * Rows of one call form ONE batch (single-threaded here: the batch-wide break couples the rows). */
API int oracle_rollout_segments(int robot, int n_boxes, const float *boxes, int n_sph, const float *spheres, int floor,
                                int hidden, int layers, const float *weights, const float *q0, const float *goal,
                                long goal_rows, long T, int n_seg, int seg_len, int ctrl_every, float dt, int stuck_lag,
                                int stuck_after, float stuck_eps, const float *K, const float *u_lo, const float *u_hi,
                                float thr_soft, float thr_hard, float lam, float penalty, float *seg_q, uint8_t *seg_alive,
                                int *seg_done, float *traj) {
    if (hidden > MAXH || layers > MAXL || T > 4096) return -1;
    const Robot *rb = &ROBOTS[robot];
    const int n = rb->n;
    Scene sc = {n_boxes, n_sph, floor, boxes, spheres};
    Net net = {n + 1, hidden, layers, weights};
    float (*d)[2 * MAXN + 1] = malloc(sizeof(float[2 * MAXN + 1]) * T);
    float (*u)[MAXN] = calloc(T, sizeof(float[MAXN]));
    int *ok = malloc(sizeof(int) * T), *dead = malloc(sizeof(int) * T), *done = malloc(sizeof(int) * T);
    CollideOut co;
    for (long b = 0; b < T; ++b)
        for (int k = 0; k < n; ++k) d[b][k] = q0[b * n + k];
    const long steps = (long)n_seg * seg_len;
    for (int sg = 0; sg < n_seg; ++sg) {
        for (long b = 0; b < T; ++b) {       /* fresh observation at the start of the segment */
            collide_state(rb, &sc, d[b], thr_soft, thr_hard, &co);
            d[b][n] = co.o;
            for (int k = 0; k < n; ++k) d[b][n + 1 + k] = co.dodq[k];
            ok[b] = 1; dead[b] = 0; done[b] = 0;
        }
        for (int t = 0; t < seg_len; ++t) {
            int any_alive = 0;
            for (long b = 0; b < T; ++b) {
                float qn[MAXN];
                if (t % ctrl_every == 0) {
                    float hh, JJ[MAXN], rr;
                    const float *g = goal + (goal_rows == 1 ? 0 : b * n);
                    filter_state(n, &net, d[b], g, 0, K, u_lo, u_hi, lam, penalty, &hh, JJ, u[b], &rr);
                }
                for (int k = 0; k < n; ++k) qn[k] = fmaf(u[b][k], dt, d[b][k]);
                collide_state(rb, &sc, qn, thr_soft, thr_hard, &co);
                if ((t + 1) % ctrl_every == 0) {
                    d[b][n] = co.o;
                    for (int k = 0; k < n; ++k) d[b][n + 1 + k] = co.dodq[k];
                }
                float *tr = traj + (b * steps + (long)sg * seg_len) * n;
                if (!dead[b]) {
                    int stuck = 0;
                    if (stuck_lag > 0 && done[b] > stuck_after) {
                        const float *a = tr + (done[b] - 1) * n, *o2 = tr + (done[b] - 1 - stuck_lag) * n;
                        float dsq = 0.f;
                        for (int k = 0; k < n; ++k) { float e = a[k] - o2[k]; dsq = fmaf(e, e, dsq); }
                        stuck = sqrtf(dsq) < stuck_eps;
                    }
                    if (stuck || co.unsafe) { ok[b] = 0; dead[b] = 1; }
                }
                for (int k = 0; k < n; ++k) d[b][k] = qn[k];
                if (!dead[b]) {
                    for (int k = 0; k < n; ++k) tr[done[b] * n + k] = d[b][k];
                    ++done[b];
                    any_alive = 1;
                }
            }
            if (!any_alive) break;           /* batch_tsa.py:229-230 */
        }
        for (long b = 0; b < T; ++b) {
            for (int k = 0; k < n; ++k) seg_q[(b * n_seg + sg) * n + k] = d[b][k];
            seg_alive[b * n_seg + sg] = (uint8_t)ok[b];
            seg_done[b * n_seg + sg] = done[b];
        }
    }
    free(d); free(u); free(ok); free(dead); free(done);
    return 0;
}

/* counter-based state sampler (BASELINE config 5: results must not depend on the shard count):
 *   x = hash(seed, global_row * n + j)  ->  u = (x >> 8) * 2^-24 in [0,1)  ->  q = fmaf(u, hi-lo, lo)
 * hash = two rounds of the 32-bit "lowbias32" mix over (index ^ seed*0x9E3779B9). */
static inline uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
API int oracle_sample_states(int robot, uint32_t seed, uint64_t start_row, long B, float *q) {
    const Robot *rb = &ROBOTS[robot];
    const float *lo = robot == 0 ? PANDA_QLO : MAGICIAN_QLO;
    const float *hi = robot == 0 ? PANDA_QHI : MAGICIAN_QHI;
    const int n = rb->n;
    for (long b = 0; b < B; ++b)
        for (int j = 0; j < n; ++j) {
            uint64_t idx = (start_row + (uint64_t)b) * (uint64_t)n + (uint64_t)j;
            uint32_t x = mix32((uint32_t)idx ^ (seed * 0x9E3779B9U));
            x = mix32(x + (uint32_t)(idx >> 32) + 0x85ebca6bU);
            float u = (float)(x >> 8) * 5.9604644775390625e-8f;
            q[b * n + j] = fmaf(u, hi[j] - lo[j], lo[j]);
        }
    return 0;
}

/* TOOL (dev container only; not part of the product or of the tests' oracle) -- exact distance between a convex
 * polytope (the convex hull of a link's collision mesh, what Bullet checks in the reference: arm_dynamics.py:169-187,
 * basic_robot.py:89-98) and an axis-aligned box / another hull, by GJK in double precision.  Used by
 * tools/geometry/fidelity.py to measure how far the sphere model of spec/robot_models.json is from those hulls.
 *
 * Convention: returns the Euclidean distance between the two sets, 0 when they intersect (penetration depth is not
 * computed; the masks only need the sign and the positive range).
 *
 * gcc -O2 -shared -fPIC -o _build/libgjk.so gjk_hull.c -lm -pthread */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double x, y, z; } v3;
static inline v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul(v3 a, double s) { return V(a.x * s, a.y * s, a.z * s); }
static inline double dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 cross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

/* a convex set given by its vertices in a local frame + pose (R row-major, p), or a box (nv == 0: centre p, half h) */
typedef struct { const double *verts; int nv; double R[9]; v3 p; v3 h; } Shape;

static v3 support(const Shape *s, v3 d) {
    if (s->nv == 0) return V(s->p.x + (d.x >= 0 ? s->h.x : -s->h.x), s->p.y + (d.y >= 0 ? s->h.y : -s->h.y),
                             s->p.z + (d.z >= 0 ? s->h.z : -s->h.z));
    /* local direction = R^T d */
    const double *R = s->R;
    const double lx = R[0] * d.x + R[3] * d.y + R[6] * d.z, ly = R[1] * d.x + R[4] * d.y + R[7] * d.z,
                 lz = R[2] * d.x + R[5] * d.y + R[8] * d.z;
    int best = 0; double bv = -1e300;
    for (int i = 0; i < s->nv; ++i) {
        const double v = s->verts[3 * i] * lx + s->verts[3 * i + 1] * ly + s->verts[3 * i + 2] * lz;
        if (v > bv) { bv = v; best = i; }
    }
    const double *q = s->verts + 3 * best;
    return V(R[0] * q[0] + R[1] * q[1] + R[2] * q[2] + s->p.x, R[3] * q[0] + R[4] * q[1] + R[5] * q[2] + s->p.y,
             R[6] * q[0] + R[7] * q[1] + R[8] * q[2] + s->p.z);
}

/* closest point to the origin on a simplex of 1..4 points (Ericson, Real-Time Collision Detection 5.1); reduces the
 * simplex to the supporting sub-simplex.  Returns the closest point. */
static v3 closest_segment(v3 *s, int *n) {
    v3 a = s[0], b = s[1], ab = sub(b, a);
    double t = -dot(a, ab), den = dot(ab, ab);
    if (t <= 0 || den <= 0) { *n = 1; return a; }
    if (t >= den) { s[0] = b; *n = 1; return b; }
    return add(a, mul(ab, t / den));
}
static v3 closest_triangle(v3 *s, int *n) {
    v3 a = s[0], b = s[1], c = s[2];
    v3 ab = sub(b, a), ac = sub(c, a), ap = mul(a, -1);
    double d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) { *n = 1; return a; }
    v3 bp = mul(b, -1);
    double d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) { s[0] = b; *n = 1; return b; }
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) { *n = 2; return add(a, mul(ab, d1 / (d1 - d3))); }
    v3 cp = mul(c, -1);
    double d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) { s[0] = c; *n = 1; return c; }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) { s[1] = c; *n = 2; return add(a, mul(ac, d2 / (d2 - d6))); }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        s[0] = b; s[1] = c; *n = 2;
        return add(b, mul(sub(c, b), w));
    }
    double den = 1.0 / (va + vb + vc), v = vb * den, w = vc * den;
    return add(a, add(mul(ab, v), mul(ac, w)));
}
static int origin_outside(v3 a, v3 b, v3 c, v3 d) {   /* origin and d on opposite sides of plane abc */
    v3 nrm = cross(sub(b, a), sub(c, a));
    double so = dot(mul(a, -1), nrm), sd = dot(sub(d, a), nrm);
    return so * sd < 0;
}
static v3 closest_tetra(v3 *s, int *n) {
    v3 a = s[0], b = s[1], c = s[2], d = s[3];
    double best = 1e300; v3 bp = V(0, 0, 0); v3 bs[3]; int bn = 0, inside = 1;
    v3 faces[4][4] = {{a, b, c, d}, {a, c, d, b}, {a, d, b, c}, {b, d, c, a}};
    for (int f = 0; f < 4; ++f) {
        if (origin_outside(faces[f][0], faces[f][1], faces[f][2], faces[f][3])) {
            inside = 0;
            v3 t[3] = {faces[f][0], faces[f][1], faces[f][2]}; int tn = 3;
            v3 q = closest_triangle(t, &tn);
            double dd = dot(q, q);
            if (dd < best) { best = dd; bp = q; bn = tn; memcpy(bs, t, sizeof t); }
        }
    }
    if (inside) return V(0, 0, 0);
    memcpy(s, bs, sizeof bs); *n = bn;
    return bp;
}

static double gjk_distance(const Shape *A, const Shape *B) {
    v3 s[4]; int n = 0;
    v3 v = sub(support(A, V(1, 0, 0)), support(B, V(-1, 0, 0)));
    s[0] = v; n = 1;
    for (int it = 0; it < 64; ++it) {
        double vv = dot(v, v);
        if (vv < 1e-20) return 0.0;
        v3 d = mul(v, -1);
        v3 w = sub(support(A, d), support(B, v));
        if (vv - dot(v, w) <= 1e-12 * vv + 1e-18) return sqrt(vv);   /* no progress possible: v is the closest point */
        s[n++] = w;
        /* newest point last; the sub-algorithms expect it anywhere */
        if (n == 2) v = closest_segment(s, &n);
        else if (n == 3) v = closest_triangle(s, &n);
        else { v = closest_tetra(s, &n); if (n == 4) return 0.0; }
        if (n == 4) return 0.0;
    }
    return sqrt(dot(v, v));
}

/* ---- batched entry points (pthreads over states) */
typedef struct {
    const double *verts; const int *voff; int n_links;          /* hull vertices of every link, concatenated */
    const double *frames;                                        /* [B][n_links][12]: R (9, row-major) then p (3) */
    const double *boxes; int n_boxes;                            /* [n_boxes][6] centre, half extents */
    const int *pairs; int n_pairs;                               /* self pairs (link a, link b) */
    double *d_box;                                               /* out [B][n_links][n_boxes] */
    double *d_self;                                              /* out [B][n_pairs] */
    long B; int n_threads;
} Job;
typedef struct { Job *j; int tid; } Arg;

static void *worker(void *p) {
    Arg *a = (Arg *)p; Job *j = a->j;
    for (long b = a->tid; b < j->B; b += j->n_threads) {
        for (int l = 0; l < j->n_links; ++l) {
            const int nv = j->voff[l + 1] - j->voff[l];
            if (nv == 0) { for (int k = 0; k < j->n_boxes; ++k) j->d_box[(b * j->n_links + l) * j->n_boxes + k] = INFINITY; continue; }
            Shape A; A.verts = j->verts + 3 * j->voff[l]; A.nv = nv;
            const double *f = j->frames + (b * j->n_links + l) * 12;
            memcpy(A.R, f, 9 * sizeof(double)); A.p = V(f[9], f[10], f[11]);
            for (int k = 0; k < j->n_boxes; ++k) {
                Shape Bx; Bx.nv = 0; Bx.verts = 0;
                Bx.p = V(j->boxes[6 * k], j->boxes[6 * k + 1], j->boxes[6 * k + 2]);
                Bx.h = V(j->boxes[6 * k + 3], j->boxes[6 * k + 4], j->boxes[6 * k + 5]);
                j->d_box[(b * j->n_links + l) * j->n_boxes + k] = gjk_distance(&A, &Bx);
            }
        }
        for (int q = 0; q < j->n_pairs; ++q) {
            const int la = j->pairs[2 * q], lb = j->pairs[2 * q + 1];
            Shape A, Bs;
            A.verts = j->verts + 3 * j->voff[la]; A.nv = j->voff[la + 1] - j->voff[la];
            Bs.verts = j->verts + 3 * j->voff[lb]; Bs.nv = j->voff[lb + 1] - j->voff[lb];
            if (A.nv == 0 || Bs.nv == 0) { j->d_self[b * j->n_pairs + q] = INFINITY; continue; }
            const double *fa = j->frames + (b * j->n_links + la) * 12, *fb = j->frames + (b * j->n_links + lb) * 12;
            memcpy(A.R, fa, 9 * sizeof(double)); A.p = V(fa[9], fa[10], fa[11]);
            memcpy(Bs.R, fb, 9 * sizeof(double)); Bs.p = V(fb[9], fb[10], fb[11]);
            j->d_self[b * j->n_pairs + q] = gjk_distance(&A, &Bs);
        }
    }
    return 0;
}

__attribute__((visibility("default")))
int hull_distances(const double *verts, const int *voff, int n_links, const double *frames, long B, const double *boxes,
                   int n_boxes, const int *pairs, int n_pairs, double *d_box, double *d_self, int n_threads) {
    Job j = {verts, voff, n_links, frames, boxes, n_boxes, pairs, n_pairs, d_box, d_self, B, n_threads};
    pthread_t th[64]; Arg args[64];
    if (n_threads > 64) n_threads = j.n_threads = 64;
    for (int t = 0; t < n_threads; ++t) { args[t].j = &j; args[t].tid = t; pthread_create(&th[t], 0, worker, &args[t]); }
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
    return 0;
}

// host-side build of the fp32 device solver for debugging: reads a,c,ur from a binary file, writes u,r,iters
#include "../../cbf-inc-rrt_b200/csrc/mlp_qp.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace cbfrrt;
#ifndef REAL
#define REAL double
#endif
template <int N> void run(int S, long B, const float* a, const float* c, const float* ur, float p, int maxit, float tol, float* u, float* r, int* it) {
    REAL lo[N], hi[N];
    for (int k = 0; k < N; ++k) { lo[k] = -1; hi[k] = 1; }
    for (long b = 0; b < B; ++b) {
        REAL aa[4][N], cc[4], uu[N], rr[4], urr[N];
        for (int i = 0; i < 4; ++i) { cc[i] = i < S ? c[b * S + i] : 0.f; for (int k = 0; k < N; ++k) aa[i][k] = i < S ? a[(b * S + i) * N + k] : 0.f; }
        for (int k = 0; k < N; ++k) urr[k] = ur[b * N + k];
        it[b] = qp_solve_multi<N, 4, REAL>(aa, urr, lo, hi, cc, S, (REAL)fminf(p, 1e6f), maxit, (REAL)tol, uu, rr);
        for (int k = 0; k < N; ++k) u[b * N + k] = uu[k];
        for (int i = 0; i < S; ++i) r[b * S + i] = rr[i];
    }
}
int main(int argc, char** argv) {
    int n = atoi(argv[1]), S = atoi(argv[2]); long B = atol(argv[3]); float p = atof(argv[4]); int maxit = atoi(argv[5]); float tol = atof(argv[6]);
    std::vector<float> a(B * S * n), c(B * S), ur(B * n), u(B * n), r(B * S); std::vector<int> it(B);
    FILE* f = fopen(argv[7], "rb"); fread(a.data(), 4, a.size(), f); fread(c.data(), 4, c.size(), f); fread(ur.data(), 4, ur.size(), f); fclose(f);
    if (n == 4) run<4>(S, B, a.data(), c.data(), ur.data(), p, maxit, tol, u.data(), r.data(), it.data());
    else run<7>(S, B, a.data(), c.data(), ur.data(), p, maxit, tol, u.data(), r.data(), it.data());
    f = fopen(argv[8], "wb"); fwrite(u.data(), 4, u.size(), f); fwrite(r.data(), 4, r.size(), f); fwrite(it.data(), 4, it.size(), f); fclose(f);
    return 0;
}

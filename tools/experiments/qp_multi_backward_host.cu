// host-side build of the device reverse mode of the multi-scenario QP (mlp_qp.cuh qp_solve_multi + qp_multi_adjoint,
// double): reads a[B,S,n], c[B,S], ur[B,n], gu[B,n], gr[B,S], lo[n], hi[n]; writes ga[B,S,n], gc[B,S], gt[B,n].
#include "../../cbf-inc-rrt_b200/csrc/mlp_qp.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace cbfrrt;
constexpr int SM = 4;
template <int N> void run(int S, long B, const float* a, const float* c, const float* ur, const float* gu, const float* gr, const float* lo_,
                          const float* hi_, float p_, float* ga, float* gc, float* gt) {
    double lo[N], hi[N];
    for (int k = 0; k < N; ++k) { lo[k] = lo_[k]; hi[k] = hi_[k]; }
    const double p = fminf(p_, 1e6f);
    for (long b = 0; b < B; ++b) {
        double aa[SM][N], cc[SM], urr[N], uu[N], rr[SM], g[N], grr[SM], oa[SM][N], oc[SM], ot[N];
        for (int i = 0; i < SM; ++i) {
            cc[i] = i < S ? c[b * S + i] : 0.0; grr[i] = i < S ? gr[b * S + i] : 0.0;
            for (int k = 0; k < N; ++k) aa[i][k] = i < S ? a[(b * S + i) * N + k] : 0.0;
        }
        for (int k = 0; k < N; ++k) { urr[k] = ur[b * N + k]; g[k] = gu[b * N + k]; }
        qp_solve_multi<N, SM, double>(aa, urr, lo, hi, cc, S, p, 100, 1e-9, uu, rr);
        qp_multi_adjoint<N, SM, double>(aa, urr, lo, hi, cc, S, p, uu, g, grr, oa, oc, ot);
        for (int k = 0; k < N; ++k) gt[b * N + k] = (float)ot[k];
        for (int i = 0; i < S; ++i) { gc[b * S + i] = (float)oc[i]; for (int k = 0; k < N; ++k) ga[(b * S + i) * N + k] = (float)oa[i][k]; }
    }
}
int main(int argc, char** argv) {
    int n = atoi(argv[1]), S = atoi(argv[2]); long B = atol(argv[3]); float p = atof(argv[4]);
    std::vector<float> a(B * S * n), c(B * S), ur(B * n), gu(B * n), gr(B * S), lo(n), hi(n), ga(B * S * n), gc(B * S), gt(B * n);
    FILE* f = fopen(argv[5], "rb");
    size_t k = fread(a.data(), 4, a.size(), f) + fread(c.data(), 4, c.size(), f) + fread(ur.data(), 4, ur.size(), f) + fread(gu.data(), 4, gu.size(), f) +
               fread(gr.data(), 4, gr.size(), f) + fread(lo.data(), 4, n, f) + fread(hi.data(), 4, n, f);
    fclose(f);
    if (k != a.size() + 2 * c.size() + 2 * ur.size() + 2 * (size_t)n) return 2;
#define CASE(NN) case NN: run<NN>(S, B, a.data(), c.data(), ur.data(), gu.data(), gr.data(), lo.data(), hi.data(), p, ga.data(), gc.data(), gt.data()); break;
    switch (n) { CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) default: return 3; }
    f = fopen(argv[6], "wb"); fwrite(ga.data(), 4, ga.size(), f); fwrite(gc.data(), 4, gc.size(), f); fwrite(gt.data(), 4, gt.size(), f); fclose(f);
    return 0;
}

// host-side build of the device reverse-mode QP (mlp_qp.cuh qp_backward<N,double>, incl. the float-guided root find):
// reads a,c,ur,gu,gr,lo,hi from a binary file, writes ga,gc,gt.  Checked against the oracle by qp_backward_host_check.py
#include "../../cbf-inc-rrt_b200/csrc/mlp_qp.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace cbfrrt;
template <int N> void run(long B, const float* a, const float* c, const float* ur, const float* gu, const float* gr, const float* lo_, const float* hi_,
                          float p, float* ga, float* gc, float* gt) {
    double lo[N], hi[N];
    for (int k = 0; k < N; ++k) { lo[k] = lo_[k]; hi[k] = hi_[k]; }
    for (long b = 0; b < B; ++b) {
        double aa[N], urr[N], g[N], oa[N], ot[N], oc;
        for (int k = 0; k < N; ++k) { aa[k] = a[b * N + k]; urr[k] = ur[b * N + k]; g[k] = gu[b * N + k]; }
        qp_backward<N, double>(aa, urr, lo, hi, (double)c[b], (double)fminf(p, 1e6f), g, (double)gr[b], oa, oc, ot);
        for (int k = 0; k < N; ++k) { ga[b * N + k] = (float)oa[k]; gt[b * N + k] = (float)ot[k]; }
        gc[b] = (float)oc;
    }
}
int main(int argc, char** argv) {
    int n = atoi(argv[1]); long B = atol(argv[2]); float p = atof(argv[3]);
    std::vector<float> a(B * n), c(B), ur(B * n), gu(B * n), gr(B), lo(n), hi(n), ga(B * n), gc(B), gt(B * n);
    FILE* f = fopen(argv[4], "rb");
    size_t k = fread(a.data(), 4, a.size(), f) + fread(c.data(), 4, c.size(), f) + fread(ur.data(), 4, ur.size(), f) + fread(gu.data(), 4, gu.size(), f) +
               fread(gr.data(), 4, gr.size(), f) + fread(lo.data(), 4, n, f) + fread(hi.data(), 4, n, f);
    fclose(f);
    if (k != a.size() * 3 + 2 * (size_t)B + 2 * n) return 2;
#define CASE(NN) case NN: run<NN>(B, a.data(), c.data(), ur.data(), gu.data(), gr.data(), lo.data(), hi.data(), p, ga.data(), gc.data(), gt.data()); break;
    switch (n) { CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) default: return 3; }
    f = fopen(argv[5], "wb"); fwrite(ga.data(), 4, ga.size(), f); fwrite(gc.data(), 4, gc.size(), f); fwrite(gt.data(), 4, gt.size(), f); fclose(f);
    return 0;
}

// k_edges_magician.cu -- instantiates the edges kernels for the Magician (see kernels.cuh)
#include "kernels.cuh"

namespace cbfrrt {
int launch_edges_magician(const SceneArgs& sc, const float* qf, const float* qt, long long E, int K, float ts, float th, uint8_t* valid, int* first_bad, cudaStream_t st) { return launch_edges<Magician>(sc, qf, qt, E, K, ts, th, valid, first_bad, st); }
int launch_edges_var_magician(const SceneArgs& sc, const float* qf, const float* qt, long long E, const long long* off, long long P, float ts, float th, uint8_t* valid, int* first_bad, cudaStream_t st) { return launch_edges_var<Magician>(sc, qf, qt, E, off, 0, P, ts, th, valid, first_bad, st); }
}  // namespace cbfrrt

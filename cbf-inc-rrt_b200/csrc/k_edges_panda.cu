// k_edges_panda.cu -- instantiates the edges kernels for the Panda (see kernels.cuh)
#include "kernels.cuh"

namespace cbfrrt {
int launch_edges_panda(const SceneArgs& sc, const float* qf, const float* qt, long long E, int K, float ts, float th, uint8_t* valid, int* first_bad, cudaStream_t st) { return launch_edges<Panda>(sc, qf, qt, E, K, ts, th, valid, first_bad, st); }
int launch_edges_var_panda(const SceneArgs& sc, const float* qf, const float* qt, long long E, const long long* off, long long P, float ts, float th, uint8_t* valid, int* first_bad, cudaStream_t st) { return launch_edges_var<Panda>(sc, qf, qt, E, off, 0, P, ts, th, valid, first_bad, st); }
}  // namespace cbfrrt

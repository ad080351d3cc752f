// k_collide_magician.cu -- instantiates the collide kernels for the Magician (see kernels.cuh)
#include "kernels.cuh"

namespace cbfrrt {
int launch_collide_magician(const SceneArgs& sc, const float* q, long long qs, long long B, float ts, float th, const CollideOutPtrs& out, cudaStream_t st) { return launch_collide<Magician>(sc, q, qs, B, ts, th, out, st); }
}  // namespace cbfrrt

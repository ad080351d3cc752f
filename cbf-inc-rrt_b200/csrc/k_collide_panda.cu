// k_collide_panda.cu -- instantiates the collide kernels for the Panda (see kernels.cuh)
#include "kernels.cuh"

namespace cbfrrt {
int launch_collide_panda(const SceneArgs& sc, const float* q, long long qs, long long B, float ts, float th, const CollideOutPtrs& out, cudaStream_t st) { return launch_collide<Panda>(sc, q, qs, B, ts, th, out, st); }
}  // namespace cbfrrt

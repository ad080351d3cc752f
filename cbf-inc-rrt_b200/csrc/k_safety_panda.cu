// k_safety_panda.cu -- instantiates the safety kernels for the Panda (see kernels.cuh)
#include "kernels.cuh"

namespace cbfrrt {
int launch_safety_panda(const SceneArgs& sc, const NetArgs& net, const float* q, long long qs, long long B, float ts, float th, const CollideOutPtrs& co, const FilterOutPtrs& fo, const PeerOut& po, cudaStream_t st) { return launch_safety<Panda>(sc, net, q, qs, B, ts, th, co, fo, po, st); }
}  // namespace cbfrrt

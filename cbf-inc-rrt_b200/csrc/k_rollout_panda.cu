// k_rollout_panda.cu -- instantiates the rollout kernels for the Panda (see kernels.cuh)
#include "kernels.cuh"

namespace cbfrrt {
int launch_rollout_panda(const SceneArgs& sc, const NetArgs& net, const float* q0, long long T, const RolloutArgs& ra, float ts, float th, float* qf, uint8_t* alive, int* done, float* traj, cudaStream_t st) { return launch_rollout<Panda>(sc, net, q0, T, ra, ts, th, qf, alive, done, traj, st); }
}  // namespace cbfrrt

// k_rollout_magician.cu -- instantiates the rollout kernels for the Magician (see kernels.cuh)
#include "kernels.cuh"

namespace cbfrrt {
int launch_rollout_magician(const SceneArgs& sc, const NetArgs& net, const float* q0, long long T, const RolloutArgs& ra, float ts, float th, float* qf, uint8_t* alive, int* done, float* traj, cudaStream_t st) { return launch_rollout<Magician>(sc, net, q0, T, ra, ts, th, qf, alive, done, traj, st); }
int launch_rollout_segments_magician(const SceneArgs& sc, const NetArgs& net, const float* q0, int T, int n_seg, int seg_len, const RolloutArgs& ra, float ts, float th, float* seg_q, uint8_t* seg_alive, int* seg_done, float* traj, cudaStream_t st) { return launch_rollout_segments<Magician>(sc, net, q0, T, n_seg, seg_len, ra, ts, th, seg_q, seg_alive, seg_done, traj, st); }
}  // namespace cbfrrt

#!/usr/bin/env python3
"""Where does one unbatched planner expansion (batch = 1, nine 20-step segments) spend its time?  Device time of
cbfrrt_rollout_segments (CUDA events), the call + read-back, and the rest of global_explore (host Python)."""
import cProfile
import json
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "cbf-inc-rrt_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from cbfrrt import ops, workloads as W  # noqa: E402
from cbfrrt.environment import ArmEnv  # noqa: E402
from cbfrrt.neural_cbf.systems import ArmMindis  # noqa: E402
from cbfrrt.neural_cbf.controllers import NeuralMindisCBFController  # noqa: E402
from cbfrrt.motion_planning.baseline import batch_RRT_plan, steer_segments_fused  # noqa: E402

env = ArmEnv(["panda"], config_file="", include_floor=True)
b = W.scene(3)[0]
env.reset_env((b[:, :3], b[:, 3:]))
dyn = ArmMindis({"m": 1}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=env, robot=env.robot_list[0])
torch.manual_seed(1)
ctrl = NeuralMindisCBFController(dyn, [{"m": 1}], None, None, cbf_alpha=0.1, cbf_relaxation_penalty=500, safe_level=0.01,
                                 unsafe_level=0.01, cbf_hidden_layers=2, cbf_hidden_size=48)
x = dyn.sample_state_space(3000)
xs = x[dyn.safe_mask(x)][:, :7].cpu().numpy()
out = {}
for B in (1, 10, 50):
    starts, targets = xs[:B], xs[B:2 * B]
    for _ in range(3):
        steer_segments_fused(ctrl, dyn, targets, starts, 9)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        steer_segments_fused(ctrl, dyn, targets, starts, 9)
    wall = (time.perf_counter() - t0) / 50
    dyn.set_intermediate_goals(targets)
    q0 = torch.as_tensor(starts, dtype=torch.float32, device="cuda")
    bb, ss, fl, gr = env.scene
    lo, hi = ctrl._limits()
    args = (q0, dyn.goal_tensor("cuda"), 9, 20, 4, 1 / 120, 9, 10, 3e-2, bb, ss, fl, gr, ctrl._weights(), dyn.K_on("cuda"), lo, hi, 0.1, 500.0,
            0.02, 0.0, 0, 48, 2)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, e in ev:
        torch.cuda.synchronize(); a.record(); ops.rollout_segments(*args); e.record()
    torch.cuda.synchronize()
    out[f"batch_{B}"] = {"steer_segments_fused_wall_ms": wall * 1e3, "kernel_plus_launch_ms_events": float(np.median([a.elapsed_time(e) for a, e in ev]))}
print(json.dumps(out))
dyn.set_goal(xs[5])
pr = cProfile.Profile()
pr.enable()
res = batch_RRT_plan(env, dyn, xs[0], xs[5], RRT_PARAM=180, controller=ctrl, T=60, model_eps=0.2, steer_type="cbf", batch=1,
                     stop_when_success=False, rng=np.random.RandomState(0))
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(18)
print("explored", res["explored_nodes"], "total", res["time_dict"]["total_time"])

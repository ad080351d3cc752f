#!/usr/bin/env python3
"""Golden vectors produced by the REFERENCE'S OWN ArmEnv / BasicRobot / ArmMindis (loaded unmodified from
/root/reference) running on the stand-in physics client of tests/fake_bullet.py, i.e. the reference's loops evaluated
on this repo's sphere model in float64 (dev container only -- the GPU box has no reference tree).

  safe_mask / unsafe_mask                      neural_cbf/systems/arm_dynamics.py:169-232, environment/basic_robot.py:89-98
  complete_sample_with_observations            neural_cbf/systems/arm_dynamics.py:267-279, arm_mindis.py:56-90

Output: tests/golden/ref_systems_{panda,magician}.npz (committed).  Consumers: tests/test_oracle_cpu.py (C oracle) and
tests/test_gpu_parity.py (CUDA kernels): masks equal away from the thresholds, o within 2e-6, do/dq within 1e-4.
    python tests/golden/make_ref_systems_golden.py"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CBFRRT_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cbf-inc-rrt_b200"))


def main():
    import make_golden as mg
    from fake_bullet import FakeBulletClient
    sys.meta_path.insert(0, mg._StubFinder())
    sys.path.insert(0, REF)
    import pybullet_utils.bullet_client as bc
    import pybullet
    import pybullet_data
    bc.BulletClient = FakeBulletClient
    pybullet.DIRECT, pybullet.GUI = FakeBulletClient.DIRECT, FakeBulletClient.GUI
    pybullet_data.getDataPath = lambda: ""
    from environment import ArmEnv              # reference code
    from neural_cbf.systems import ArmMindis    # reference code
    from cbfrrt._robot_tables import ROBOT_MODELS

    for robot, seed in (("panda", 201), ("magician", 202)):
        rng = np.random.default_rng(seed)
        m = ROBOT_MODELS[robot]
        lo, hi = np.array(m["q_lower"], np.float32), np.array(m["q_upper"], np.float32)
        q = (lo + (hi - lo) * rng.random((400, m["n"]))).astype(np.float32)
        positions, sizes = np.zeros((1, 6, 3)), rng.uniform(0.04, 0.14, size=(1, 6, 3))
        for k in range(6):            # keep the boxes off the base column so that the masks are not trivially constant
            while True:
                c = rng.uniform([-0.5, -0.5, 0.05], [0.5, 0.5, 0.9])
                if np.hypot(c[0], c[1]) > 0.3 + sizes[0, k, :2].max():
                    positions[0, k] = c
                    break
        out = dict(q=q, positions=positions[0], sizes=sizes[0])
        with tempfile.TemporaryDirectory() as d:
            cfg = os.path.join(d, "env.npz")
            np.savez(cfg, obstacle_positions=positions, obstacle_sizes=sizes)
            for floor in (1, 0):
                env = ArmEnv([robot], config_file=cfg, GUI=False, include_floor=bool(floor))
                env.reset_env(env.get_env_config(0), enable_object=False)
                dyn = ArmMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=env, robot=env.robot_list[0])
                x = torch.tensor(q)
                out[f"safe_floor{floor}"] = dyn.safe_mask(x).numpy()
                out[f"unsafe_floor{floor}"] = dyn.unsafe_mask(x).numpy()
                out[f"datax_floor{floor}"] = dyn.complete_sample_with_observations(x, len(q)).numpy()
        np.savez_compressed(os.path.join(HERE, f"ref_systems_{robot}.npz"), **out)
        print(robot, "safe", out["safe_floor1"].mean(), "unsafe", out["unsafe_floor1"].mean())


if __name__ == "__main__":
    main()

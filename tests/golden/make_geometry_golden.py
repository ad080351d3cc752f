#!/usr/bin/env python3
"""Regression fixture of THIS repo's geometry contract (DESIGN.md sections 2-3): seeded states and scenes with the C
oracle's masks / minimum distance / arg-min / gradient.  It does not pin anything against pybullet (parity there is
unpinned); it pins the sphere model, the exemption rules and the float32 operation order across
rounds, for the oracle (CPU suite) and the CUDA kernels (GPU suite) alike.
    python tests/golden/make_geometry_golden.py      # rewrites tests/golden/geometry_contract_{panda,magician}.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cbf-inc-rrt_b200"))
from oracle import c_oracle  # noqa: E402


def main():
    for robot, n in (("panda", 7), ("magician", 4)):
        rng = np.random.default_rng({"panda": 101, "magician": 102}[robot])
        from cbfrrt._robot_tables import ROBOT_MODELS
        m = ROBOT_MODELS[robot]
        lo, hi = np.array(m["q_lower"], np.float32), np.array(m["q_upper"], np.float32)
        q = (lo + (hi - lo) * rng.random((1200, n))).astype(np.float32)
        boxes = np.concatenate([rng.uniform([-0.5, -0.5, 0.0], [0.5, 0.5, 1.0], (6, 3)), rng.uniform(0.03, 0.15, (6, 3))], 1).astype(np.float32)
        spheres = np.concatenate([rng.uniform([-0.5, -0.5, 0.1], [0.5, 0.5, 1.0], (3, 3)), rng.uniform(0.03, 0.1, (3, 1))], 1).astype(np.float32)
        out = {}
        for floor in (1, 0):
            o = c_oracle.collide(robot, q, boxes, spheres, bool(floor), 0.02, 0.0)
            for k in ("safe", "unsafe", "datax", "arg_link", "arg_obs"):
                out[f"{k}_floor{floor}"] = o[k]
        pos, rot = c_oracle.fk(robot, q[:50])
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"geometry_contract_{robot}.npz"), q=q, boxes=boxes,
                            spheres=spheres, fk_pos=pos, fk_rot=rot, **out)
        print(robot, "safe", out["safe_floor1"].mean(), "unsafe", out["unsafe_floor1"].mean())


if __name__ == "__main__":
    main()

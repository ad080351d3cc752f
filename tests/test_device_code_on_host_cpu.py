"""CPU suite: the DEVICE geometry code itself (csrc/core.cuh + csrc/grid.cuh: FK, floor / self / obstacle terms, the
brute-force and culled sweeps, the nearest-candidate grid and the distance-bound field path with its pruning argument)
compiled for the host through tests/host_shim/cuda_runtime.h and compared with the C oracle -- bit for bit on masks,
observation, arg-min indices; 1e-5 on do/dq.  What the GPU suite proves on the B200 is proved here first on the same
source lines (only the launch scaffolding differs)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel_err

SHIM = os.path.join(ROOT, "tests", "host_shim")
SO = os.path.join(SHIM, "_build", "libcollide_host.so")
SRCS = [os.path.join(SHIM, "collide_host.cpp"), os.path.join(SHIM, "cuda_runtime.h")] + [
    os.path.join(ROOT, "cbf-inc-rrt_b200", "csrc", f) for f in ("core.cuh", "grid.cuh", "robot_tables.cuh")]


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(s) for s in SRCS):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I", SHIM, "-o", SO, SRCS[0]])
    lib = C.CDLL(SO)
    lib.host_grid_bytes.restype = C.c_long
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def run_host(lib, robot, q, boxes6, spheres, floor, use_grid, want_grad, ts=0.02, th=0.0):
    from oracle import c_oracle
    b8, s4 = c_oracle.pack_scene(boxes6, spheres)
    B, n = q.shape
    grid = None
    if use_grid:
        grid = np.zeros(lib.host_grid_bytes(), np.uint8)
        lib.host_grid_build(_p(b8), b8.shape[0], _p(s4), s4.shape[0], _p(grid))
    out = dict(safe=np.zeros(B, np.uint8), unsafe=np.zeros(B, np.uint8), o=np.zeros(B, np.float32),
               arg_link=np.zeros(B, np.int32), arg_obs=np.zeros(B, np.int32), dodq=np.zeros((B, n), np.float32),
               soft_hard=np.zeros((B, 2), np.float32), stats=np.zeros((B, 2), np.uint32))
    rc = lib.host_collide(c_oracle.ROBOT_ID[robot], _p(b8), b8.shape[0], _p(s4), s4.shape[0], int(floor), _p(grid) if use_grid else None,
                          _p(np.ascontiguousarray(q, np.float32)), C.c_long(B), C.c_float(ts), C.c_float(th), int(want_grad),
                          _p(out["safe"]), _p(out["unsafe"]), _p(out["o"]), _p(out["arg_link"]), _p(out["arg_obs"]),
                          _p(out["dodq"]), _p(out["soft_hard"]), _p(out["stats"]))
    assert rc == 0
    return out


def scenes():
    from cbfrrt import workloads as W
    return {
        "cfg4-256box": (W.scene(4)[0], None, True),
        "cfg2-32box-32sph": (W.scene(2)[0], W.scene(2)[1], True),
        "cfg3-8box": (W.scene(3)[0], None, True),
        "dense-1m3-256box": (W.random_boxes(256, 2, 0.01, 0.05), None, True),
        "default3": (np.array([[0.3, 0.17, 0.3, 0.05, 0.05, 0.1], [-0.4, 0.23, 0.4, 0.05, 0.05, 0.1],
                               [0.0, -0.23, 0.6, 0.05, 0.05, 0.1]], np.float32), None, True),
        "spheres-nofloor": (None, np.array([[0.3, -0.3, 0.5, 0.1], [0.0, 0.4, 0.7, 0.15]], np.float32), False),
        "far-box-only": (np.array([[0.0, 0.0, 2.5, 0.2, 0.2, 0.2]], np.float32), None, True),
        "duplicates-zero-size": (np.array([[0.35, 0.0, 0.45, 0.06, 0.06, 0.06], [0.35, 0.0, 0.45, 0.06, 0.06, 0.06],
                                           [0.2, 0.3, 0.6, 0.0, 0.0, 0.0], [0.0, 0.0, 0.5, 0.6, 0.6, 0.1]], np.float32),
                                 np.array([[0.3, 0.3, 0.3, 0.0]], np.float32), True),
    }


@pytest.mark.parametrize("robot", ["panda", "magician"])
@pytest.mark.parametrize("scene", list(scenes()))
def test_field_path_is_bit_exact(host, coracle, robot, scene):
    """distance-bound field path (observation + gradient caller) == brute-force oracle"""
    from cbfrrt import workloads as W
    boxes, sph, floor = scenes()[scene]
    q = W.sample_states(robot, 3000, 7)
    q[:40] *= 3.0                                  # far outside the joint limits: sphere centres leave the grid region
    o = coracle.collide(robot, q, boxes, sph, floor, 0.02, 0.0)
    n = q.shape[1]
    for use_grid in (True, False):
        h = run_host(host, robot, q, boxes, sph, floor, use_grid, want_grad=True)
        assert np.array_equal(h["safe"], o["safe"]) and np.array_equal(h["unsafe"], o["unsafe"]), (scene, use_grid)
        assert np.array_equal(h["o"], o["datax"][:, n]), (scene, use_grid)
        assert np.array_equal(h["arg_link"], o["arg_link"]) and np.array_equal(h["arg_obs"], o["arg_obs"]), (scene, use_grid)
        assert rel_err(h["dodq"], o["datax"][:, n + 1:]) < 1e-5
    st = h if False else run_host(host, robot, q[40:], boxes, sph, floor, True, want_grad=True)["stats"]
    n_sph = {"panda": 28, "magician": 20}[robot]
    print(f"[field path] {robot} {scene}: exactly swept spheres / state {st[:, 0].mean():.2f} of {n_sph}, "
          f"pair evaluations / state {st[:, 1].mean():.1f}")
    if boxes is not None and len(boxes) >= 8 and "far" not in scene:
        assert st[:, 0].mean() < 0.6 * n_sph           # the pruning does prune


@pytest.mark.parametrize("robot", ["panda", "magician"])
@pytest.mark.parametrize("thr", [(0.02, 0.0), (0.05, 0.0), (0.02, 0.02), (0.0, -0.01)])
def test_field_path_masks_only(host, coracle, robot, thr):
    """mask-only caller (edges, masks, the inner steps of a rollout): same booleans as the oracle for several threshold
    pairs, including thr_hard >= thr_soft and negative thresholds"""
    from cbfrrt import workloads as W
    for scene in ("cfg4-256box", "cfg2-32box-32sph", "cfg3-8box", "duplicates-zero-size"):
        boxes, sph, floor = scenes()[scene]
        q = W.sample_states(robot, 3000, 11)
        o = coracle.collide(robot, q, boxes, sph, floor, thr[0], thr[1])
        h = run_host(host, robot, q, boxes, sph, floor, True, want_grad=False, ts=thr[0], th=thr[1])
        assert np.array_equal(h["safe"], o["safe"]) and np.array_equal(h["unsafe"], o["unsafe"]), (scene, thr)
        assert 0 < o["unsafe"].mean() < 1 or scene == "duplicates-zero-size"

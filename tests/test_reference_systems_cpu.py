"""Dev-container-only (skipped where /root/reference is absent): the reference's OWN environment / dynamics classes
(environment/arm_env.py, basic_robot.py, franka_panda.py, magician.py, neural_cbf/systems/arm_dynamics.py,
arm_mindis.py -- loaded unmodified) run on a stand-in physics client that answers the geometric primitives from this
repo's sphere model (tests/fake_bullet.py).  Everything AROUND the primitives is then the reference's code: the link
ranges, the floor exemptions and collision filters, the self-collision pair loop, the dropped first record, the
arg-min and its ties, and the do/dq assembly n^T J.  Their results must equal this repo's restatement (the C oracle,
to which the CUDA kernels are bit-exact)."""
import os
import sys

import numpy as np
import pytest
import torch

REF = os.environ.get("CBFRRT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "environment", "arm_env.py")),
                                reason="reference tree not available")


@pytest.fixture(scope="module")
def reference_classes():
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden as mg                      # inert stubs for the absent third-party modules
    from fake_bullet import FakeBulletClient
    finder = mg._StubFinder()
    sys.meta_path.insert(0, finder)
    sys.path.insert(0, REF)
    saved = {k: sys.modules.get(k) for k in ("environment", "neural_cbf")}
    try:
        import pybullet_utils.bullet_client as bc
        import pybullet
        import pybullet_data
        bc.BulletClient = FakeBulletClient
        pybullet.DIRECT, pybullet.GUI = FakeBulletClient.DIRECT, FakeBulletClient.GUI
        pybullet_data.getDataPath = lambda: ""
        from environment import ArmEnv              # noqa: E402  (reference code)
        from neural_cbf.systems import ArmMindis    # noqa: E402  (reference code)
        yield ArmEnv, ArmMindis
    finally:
        sys.meta_path.remove(finder)
        sys.path.remove(REF)
        for k in list(sys.modules):
            if k.split(".")[0] in ("environment", "neural_cbf", "pybullet", "pybullet_utils", "pybullet_data", "pinocchio",
                                   "cvxpy", "cvxpylayers", "pytorch_lightning", "matplotlib", "seaborn", "gurobipy", "qpsolvers"):
                del sys.modules[k]


@pytest.mark.parametrize("robot,floor", [("panda", True), ("magician", True), ("panda", False)])
def test_reference_masks_and_observation_loops(reference_classes, coracle, tmp_path, robot, floor):
    ArmEnv, ArmMindis = reference_classes
    from cbfrrt import workloads as W
    rng = np.random.default_rng(4)
    positions = rng.uniform([-0.5, -0.5, 0.05], [0.5, 0.5, 0.9], size=(1, 5, 3))
    sizes = rng.uniform(0.04, 0.14, size=(1, 5, 3))
    cfg = tmp_path / "env.npz"
    np.savez(cfg, obstacle_positions=positions, obstacle_sizes=sizes)
    env = ArmEnv([robot], config_file=str(cfg), GUI=False, include_floor=floor)
    env.reset_env(env.get_env_config(0), enable_object=False)
    rb = env.robot_list[0]
    dyn = ArmMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=env, robot=rb)
    n = rb.body_dim
    assert str(rb) in ("Franka_Panda", "Magician") and len(env.obstacle_ids) == 5 + int(floor)

    q = W.sample_states(robot, 150, 9)
    x = torch.tensor(q)
    ref_safe = dyn.safe_mask(x).numpy()                       # arm_dynamics.py:169-210 + basic_robot.py:89-98
    ref_unsafe = dyn.unsafe_mask(x).numpy()                   # arm_dynamics.py:184-187,212-232 + arm_env.py:82-84
    ref_datax = dyn.complete_sample_with_observations(x, len(q)).numpy()      # arm_mindis.py:56-90

    boxes = np.concatenate([positions[0], sizes[0]], 1).astype(np.float32)
    o = coracle.collide(robot, q, boxes, None, floor, 0.02, 0.0)
    self_min, soft_min, hard_min = o["mins"][:, 0], o["mins"][:, 1], o["mins"][:, 2]
    # the fake answers in float64, the oracle computes in float32: only states within 1e-5 of a threshold may differ
    clear = (np.abs(soft_min - 0.02) > 1e-5) & (np.abs(hard_min) > 1e-5) & (np.abs(self_min - 0.01) > 1e-5)
    assert clear.mean() > 0.95
    assert np.array_equal(ref_safe[clear], o["safe"][clear].astype(bool))
    assert np.array_equal(ref_unsafe[clear], o["unsafe"][clear].astype(bool))
    assert 0.03 < ref_safe.mean() < 0.97 and 0.03 < ref_unsafe.mean() < 0.97
    assert ref_datax.shape == (len(q), 2 * n + 1) and np.array_equal(ref_datax[:, :n], q)
    assert np.abs(ref_datax[:, n] - o["datax"][:, n]).max() < 2e-6                      # o: min distance
    scale = np.maximum(np.abs(o["datax"][:, n + 1:]), 0.1)
    assert (np.abs(ref_datax[:, n + 1:] - o["datax"][:, n + 1:]) / scale).max() < 1e-4  # do/dq = n^T J
    # the reference's own limits / sizes on the stand-in
    up, lo = dyn.state_limits
    assert np.allclose(up.numpy(), W.ROBOT_MODELS[robot]["q_upper"]) and np.allclose(lo.numpy(), W.ROBOT_MODELS[robot]["q_lower"])
    cu, cl = dyn.control_limits
    assert np.allclose(cu.numpy(), W.ROBOT_MODELS[robot]["v_max"]) and torch.equal(cl, -cu)

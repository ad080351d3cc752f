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
        from neural_cbf.controllers.neural_mindis_cbf_controller import NeuralMindisCBFController  # noqa: E402
        yield ArmEnv, ArmMindis, NeuralMindisCBFController, mg
    finally:
        sys.meta_path.remove(finder)
        sys.path.remove(REF)
        for k in list(sys.modules):
            if k.split(".")[0] in ("environment", "neural_cbf", "pybullet", "pybullet_utils", "pybullet_data", "pinocchio",
                                   "cvxpy", "cvxpylayers", "pytorch_lightning", "matplotlib", "seaborn", "gurobipy", "qpsolvers"):
                del sys.modules[k]


@pytest.mark.parametrize("robot,floor", [("panda", True), ("magician", True), ("panda", False)])
def test_reference_masks_and_observation_loops(reference_classes, coracle, tmp_path, robot, floor):
    ArmEnv, ArmMindis = reference_classes[:2]
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
    # goal / boundary masks and one closed-loop step with observation refresh (arm_dynamics.py:234-262,474-523,
    # control_affine_system.py:254-268) against this package's mirrors on the oracle-backed ops
    import fake_ops

    class _MP:
        def setattr(self, obj, name, val):
            setattr(obj, name, val)
    from cbfrrt import ops as _ops
    saved = {k: getattr(_ops, k) for k in dir(_ops) if not k.startswith("__")}
    try:
        fake_ops.install(_MP())
        from cbfrrt.environment import ArmEnv as MyEnv
        from cbfrrt.neural_cbf.systems import ArmMindis as MyMindis
        m_env = MyEnv([robot], config_file="", include_floor=floor, device="cpu")
        m_env.reset_env((positions[0], sizes[0]))
        m_dyn = MyMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=m_env, robot=m_env.robot_list[0])
        goal = q[np.argmax(ref_safe)].astype(np.float64)
        dyn.set_goal(goal)
        m_dyn.set_goal(goal)
        xg = torch.tensor(np.concatenate([q[:40], goal[None].astype(np.float32) + 0.01], 0))
        assert torch.equal(dyn.goal_mask(xg), m_dyn.goal_mask(xg)) and bool(dyn.goal_mask(xg)[-1])
        ok = torch.tensor(clear[:40])
        assert torch.equal(dyn.boundary_mask(x[:40])[ok], m_dyn.boundary_mask(x[:40])[ok])
        u = torch.tensor(np.random.default_rng(3).uniform(-1, 1, size=(40, n)).astype(np.float32))
        r_next = dyn.closed_loop_dynamics(torch.tensor(ref_datax[:40]), u)
        m_next = m_dyn.closed_loop_dynamics(torch.tensor(ref_datax[:40]), u)
        assert r_next.shape == m_next.shape == (40, 2 * n + 1)
        assert torch.equal(r_next[:, :n], m_next[:, :n]) and (r_next[:, n] - m_next[:, n]).abs().max() < 2e-6
        sc = torch.clamp(r_next[:, n + 1:].abs(), min=0.1)
        assert ((r_next[:, n + 1:] - m_next[:, n + 1:]).abs() / sc).max() < 1e-4
    finally:
        for k, v in saved.items():
            setattr(_ops, k, v)


def test_whole_reference_stack_steers_like_this_package(reference_classes, coracle, tmp_path, monkeypatch):
    """The reference's planner `steer` on the reference's dynamics + controller (stand-in physics client, the QP layer
    replaced by a float64 SLSQP solve) against this package's one-launch steer on its own classes (oracle-backed ops):
    same barrier weights, same scene -> the same trajectory, collision flag and length."""
    ArmEnv, ArmMindis, RefController, mg = reference_classes
    import importlib.util
    import fake_ops
    fake_ops.install(monkeypatch)
    spec = importlib.util.spec_from_file_location("ref_tsa_for_stack", os.path.join(REF, "motion_planning", "baseline", "tsa.py"),
                                                  submodule_search_locations=None)
    pkg = __import__("types").ModuleType("ref_stack_pkg")
    pkg.__path__ = [os.path.join(REF, "motion_planning", "baseline")]
    sys.modules["ref_stack_pkg"] = pkg
    spec = importlib.util.spec_from_file_location("ref_stack_pkg.tsa", os.path.join(REF, "motion_planning", "baseline", "tsa.py"))
    ref_tsa = importlib.util.module_from_spec(spec)
    sys.modules["ref_stack_pkg.tsa"] = ref_tsa
    spec.loader.exec_module(ref_tsa)

    positions = np.array([[[0.35, 0.2, 0.35], [-0.3, 0.3, 0.5], [0.1, -0.4, 0.25], [0.3, -0.25, 0.7]]])
    sizes = np.array([[[0.06, 0.05, 0.1], [0.05, 0.08, 0.05], [0.07, 0.05, 0.12], [0.04, 0.06, 0.05]]])
    cfg = tmp_path / "env.npz"
    np.savez(cfg, obstacle_positions=positions, obstacle_sizes=sizes)
    # ---- the reference stack
    r_env = ArmEnv(["magician"], config_file=str(cfg), GUI=False, include_floor=True)
    r_env.reset_env(r_env.get_env_config(0), enable_object=False)
    r_dyn = ArmMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=r_env, robot=r_env.robot_list[0])
    up, lo = r_dyn.control_limits
    mg.QP_LIMITS.update(lam=0.1, lo=lo.double().numpy(), hi=up.double().numpy())
    torch.manual_seed(1)
    r_ctrl = RefController(r_dyn, [{"m1": 5.76}], None, None, cbf_alpha=0.1, cbf_relaxation_penalty=5000.0, safe_level=0.1,
                           unsafe_level=0.1, cbf_hidden_layers=2, cbf_hidden_size=48, use_neural_actor=False)
    # ---- this package's stack with the same weights and scene
    from cbfrrt.environment import ArmEnv as MyEnv
    from cbfrrt.neural_cbf.systems import ArmMindis as MyMindis
    from cbfrrt.neural_cbf.controllers import NeuralMindisCBFController as MyController
    from cbfrrt.motion_planning.baseline import steer_fused
    m_env = MyEnv(["magician"], config_file="", include_floor=True, device="cpu")
    m_env.reset_env((positions[0], sizes[0]))
    m_dyn = MyMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=m_env, robot=m_env.robot_list[0])
    m_ctrl = MyController(m_dyn, [{"m1": 5.76}], None, None, cbf_alpha=0.1, cbf_relaxation_penalty=5000.0, safe_level=0.1,
                          unsafe_level=0.1, cbf_hidden_layers=2, cbf_hidden_size=48, use_neural_actor=False)
    m_ctrl.load_weights({"h_nn." + k: v for k, v in r_ctrl.h_nn.state_dict().items()})
    x = m_dyn.sample_state_space(200)
    xs = x[m_dyn.safe_mask(x)][:, :4].numpy()
    compared = 0
    for i in range(5):
        start, target = xs[i].astype(np.float32), xs[i + 15].astype(np.float32)
        r_last, r_ok, _, r_list = ref_tsa.steer(r_ctrl.u, r_dyn, target, start, RRT_step=20, device="cpu")
        m_last, m_ok, _, m_list = steer_fused(m_ctrl, m_dyn, target, start, RRT_step=20)
        assert r_ok == m_ok and len(r_list) == len(m_list), (i, r_ok, m_ok, len(r_list), len(m_list))
        if len(r_list) > 1:
            a = np.asarray([np.asarray(p)[:4] for p in r_list[1:]], dtype=np.float64)
            b = np.asarray([np.asarray(p)[:4] for p in m_list[1:]], dtype=np.float64)
            assert np.abs(a - b).max() < 2e-4, (i, np.abs(a - b).max())
            compared += len(a)
    assert compared >= 40

"""Dev-container-only (skipped where /root/reference is absent, e.g. on the GPU box): the REFERENCE'S OWN planner code
(motion_planning/baseline/tsa.py, batch_tsa.py, search_tree.py -- pure Python over the dynamics-model / controller
interface) is loaded unmodified and run on THIS package's objects (CUDA ops replaced by the oracle-backed fakes).
That is the drop-in claim of SURVEY.md 8b exercised literally: steer, edge_checking, the batched steer and a whole
RRT_plan run call exactly the attributes and methods the reference uses."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

import fake_ops

REF = os.environ.get("CBFRRT_REFERENCE", "/root/reference")
BASE = os.path.join(REF, "motion_planning", "baseline")
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(BASE, "tsa.py")), reason="reference tree not available")


def _load_reference_planner():
    """import the three pure-Python files as a throw-away package (their package __init__ pulls in pybullet code)"""
    pkg = types.ModuleType("ref_baseline")
    pkg.__path__ = [BASE]
    sys.modules["ref_baseline"] = pkg
    mods = {}
    for name in ("search_tree", "tsa", "batch_tsa", "bit_star"):
        spec = importlib.util.spec_from_file_location(f"ref_baseline.{name}", os.path.join(BASE, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        if name == "bit_star":       # it imports its siblings by their absolute names
            for sib in ("tsa", "search_tree"):
                sys.modules[f"motion_planning.baseline.{sib}"] = mods[sib]
            sys.modules.setdefault("motion_planning", types.ModuleType("motion_planning"))
            sys.modules.setdefault("motion_planning.baseline", types.ModuleType("motion_planning.baseline"))
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods


@pytest.fixture()
def world(monkeypatch, coracle):
    fake_ops.install(monkeypatch)
    from cbfrrt.environment import ArmEnv
    from cbfrrt.neural_cbf.systems import ArmMindis
    from cbfrrt.neural_cbf.controllers import NeuralMindisCBFController
    env = ArmEnv(["magician"], config_file="", include_floor=True, device="cpu")
    robot = env.robot_list[0]
    dyn = ArmMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=env, robot=robot)
    torch.manual_seed(1)
    ctrl = NeuralMindisCBFController(dyn, [{"m1": 5.76}], None, None, cbf_alpha=0.1, cbf_relaxation_penalty=5000.0,
                                     safe_level=0.1, unsafe_level=0.1, cbf_hidden_layers=2, cbf_hidden_size=48,
                                     use_neural_actor=False)
    x = dyn.sample_state_space(300)
    xs = x[dyn.safe_mask(x)][:, :4].numpy()
    return env, dyn, ctrl, xs, _load_reference_planner()


def test_reference_steer_and_edge_checking_run_on_our_objects(world):
    env, dyn, ctrl, xs, ref = world
    from cbfrrt.motion_planning.baseline import steer, steer_fused, edge_checking, edge_checking_reference_loop
    for i in range(4):
        start, target = xs[i], xs[i + 20]
        r_last, r_ok, r_td, r_list = ref["tsa"].steer(ctrl.u, dyn, target, start, RRT_step=30, device="cpu")
        m_last, m_ok, m_td, m_list = steer(ctrl.u, dyn, target, start, RRT_step=30, device="cpu")
        assert r_ok == m_ok and len(r_list) == len(m_list)
        assert np.array_equal(np.asarray(r_list[1:], dtype=np.float32), np.asarray(m_list[1:], dtype=np.float32))
        f_last, f_ok, _, f_list = steer_fused(ctrl, dyn, target, start, RRT_step=30)
        assert f_ok == r_ok and len(f_list) == len(r_list)
        assert np.abs(np.asarray(f_list[1:])[:, :4] - np.asarray(r_list[1:])[:, :4]).max() < 1e-4
        assert set(r_td) >= {"complete_sample", "cl_dynamics", "collision_checking"}
    agree = 0
    for i in range(12):
        a, b = xs[i], xs[i] + 0.5 * (xs[i + 30] - xs[i])
        r = ref["tsa"].edge_checking(a, b, dyn, 0.05)
        assert r == edge_checking_reference_loop(a, b, dyn, 0.05)
        agree += (r == edge_checking(a, b, dyn, 0.05))
    assert agree >= 11        # one launch interpolates in float32, the reference in float64: a borderline point may flip


def test_reference_batched_steer_runs_on_our_objects(world):
    env, dyn, ctrl, xs, ref = world
    from cbfrrt.motion_planning.baseline import batch_steer
    starts, targets = xs[:6], xs[40:46]
    r_x, r_ok, _, r_lists = ref["batch_tsa"].steer(ctrl.u, dyn, targets, starts, RRT_step=20, device="cpu")
    m_x, m_ok, _, m_lists = batch_steer(ctrl.u, dyn, targets, starts, RRT_step=20, device="cpu")
    assert r_ok == m_ok and [len(l) for l in r_lists] == [len(l) for l in m_lists]
    assert np.allclose(np.asarray(r_x), np.asarray(m_x), atol=1e-6)


def test_reference_rrt_plan_runs_on_our_objects(world):
    env, dyn, ctrl, xs, ref = world
    init = xs[0]
    goal = init + 0.25 * (xs[1] - init) / np.linalg.norm(xs[1] - init)
    dyn.set_goal(goal)
    np.random.seed(0)
    res = ref["tsa"].RRT_plan(env=env, dynamics_model=dyn, init_state=init, goal_state=goal, RRT_PARAM=30, controller=ctrl,
                              T=40, model_eps=0.4, steer_type="line", stop_when_success=True)
    assert set(res) >= {"success", "path", "edge_states", "explored_nodes", "all_vertex", "all_edge"}
    assert res["success"] and np.allclose(res["path"][0], init) and np.linalg.norm(res["path"][-1][:4] - goal) <= 0.2 + 1e-6
    np.random.seed(0)
    res2 = ref["tsa"].RRT_plan(env=env, dynamics_model=dyn, init_state=init, goal_state=goal, RRT_PARAM=30, controller=ctrl,
                               T=15, model_eps=0.4, steer_type="cbf", stop_when_success=True)
    pts = np.array([np.asarray(p)[:4] for e in res2["all_edge"][1:] for p in e])
    if len(pts):
        assert not dyn.unsafe_mask(torch.as_tensor(pts, dtype=torch.float32)).any()


def test_reference_bit_star_runs_on_our_objects(world):
    """BITStarPlanner (bit_star.py:12-305): is_point_free -> safe_mask, is_edge_free -> edge_checking, sampling bounds from
    state_limits; a found path is re-validated edge by edge with this package's one-launch edge_checking"""
    env, dyn, ctrl, xs, ref = world
    from cbfrrt.motion_planning.baseline import edge_checking
    init = xs[2]
    goal = init + 0.4 * (xs[3] - init) / np.linalg.norm(xs[3] - init)
    if not bool(dyn.safe_mask(torch.as_tensor(goal[None], dtype=torch.float32)).item()):
        goal = xs[3]
    np.random.seed(1)
    planner = ref["bit_star"].BITStarPlanner(dyn, init, goal, batch_size=40, T=400, edge_eps=0.05)
    vertices, edges, cost, T, dt = planner.plan(time_budget=120)
    assert planner.dimension == 4 and planner.n_free_points > 2 and len(vertices) >= 1
    assert cost < float("inf"), "a straight 0.4 rad move in free space must be found"
    path = planner.get_best_path()
    assert np.allclose(path[0], init) and np.allclose(path[-1], goal)
    for a, b in zip(path[:-1], path[1:]):
        assert edge_checking(np.array(a), np.array(b), dyn, 0.05)
    # the module-level wrapper also executes the path with the nominal steer (bit_star.py:308-351)
    np.random.seed(1)
    res = ref["bit_star"].BITStar(env, dyn, init, goal, batch_size=40, T=400, EPS=0.05, time_budget=120)
    assert res["success"] and len(res["path"]) >= 2 and res["explored_nodes"] >= 40


def test_reference_datamodule_sampling_runs_on_our_objects(world, monkeypatch):
    """EpisodicDataModule.sample_fixed (neural_cbf/datamodules/episodic_datamodule.py:146-194), the caller of the
    sampled-state sweeps (sample_safe / sample_unsafe / sample_boundary + the four masks), loaded unmodified"""
    env, dyn, ctrl, xs, ref = world
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as mg
    finder = mg._StubFinder()
    sys.meta_path.insert(0, finder)
    stub_sys = types.ModuleType("neural_cbf.systems")
    stub_sys.ControlAffineSystem = object
    monkeypatch.setitem(sys.modules, "neural_cbf", types.ModuleType("neural_cbf"))
    monkeypatch.setitem(sys.modules, "neural_cbf.systems", stub_sys)
    try:
        spec = importlib.util.spec_from_file_location(
            "ref_datamodule", os.path.join(REF, "neural_cbf", "datamodules", "episodic_datamodule.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.meta_path.remove(finder)
        for k in list(sys.modules):
            if k.split(".")[0] in ("pytorch_lightning", "matplotlib"):
                del sys.modules[k]
    up, lo = dyn.state_limits
    dm = mod.EpisodicDataModule(dyn, [(float(a), float(b)) for a, b in zip(lo, up)], max_episode=1, fixed_samples=200,
                                quotas={"safe": 0.3, "unsafe": 0.3, "boundary": 0.2}, batch_size=16)
    dyn.set_goal(xs[0])
    samples, masks = dm.sample_fixed()
    assert samples.shape == (160, 2 * 4 + 1) and set(masks) == {"safe", "unsafe", "goal", "boundary"}
    assert masks["safe"][:60].all() and masks["unsafe"][60:120].all() and masks["boundary"][120:].all()
    assert not (masks["safe"] & masks["unsafe"]).any()
    assert torch.equal(masks["boundary"], ~(masks["safe"] | masks["unsafe"]))

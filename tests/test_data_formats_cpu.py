"""the reference's on-disk formats and the data-collection callers of the path -- the training
dataset dict written by EpisodicDataModule.prepare_data (neural_cbf/datamodules/episodic_datamodule.py:196-260), the
planning-problem pickles (motion_planning/evaluation/generate_problem_env.py:51-56, collect_problems.py,
eval_rrt_all.py:74-104) and Lightning checkpoints (eval_rrt_mindis_cbfsteer.py:33-50).  CUDA ops are replaced by the
oracle-backed fakes (host logic only).  Where /root/reference exists the reference's own EpisodicDataModule is loaded
unmodified and must draw the same samples, trajectories and labels under the same seeds."""
import importlib.util
import os
import pickle
import random
import sys
import types

import numpy as np
import pytest
import torch

import fake_ops

REF = os.environ.get("CBFRRT_REFERENCE", "/root/reference")
HAVE_REF = os.path.isfile(os.path.join(REF, "neural_cbf", "datamodules", "episodic_datamodule.py"))
KW = dict(cbf_alpha=0.1, cbf_relaxation_penalty=5000.0, safe_level=0.1, unsafe_level=0.1, cbf_hidden_layers=2,
          cbf_hidden_size=48, use_neural_actor=False)


@pytest.fixture()
def stack(monkeypatch, coracle):
    fake_ops.install(monkeypatch)
    from cbfrrt.environment import ArmEnv
    from cbfrrt.neural_cbf.systems import ArmMindis

    def make(robot_name="magician", config_file=""):
        np.random.seed(5)                      # the training scenes are drawn at construction (arm_env.py:44-53)
        env = ArmEnv([robot_name], config_file=config_file, include_floor=True, device="cpu")
        dyn = ArmMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.05, env=env,
                        robot=env.robot_list[0])
        return env, dyn
    return make


def _datamodule(cls, dyn, **kw):
    up, lo = dyn.state_limits
    args = dict(max_episode=2, trajectories_per_episode=6, trajectory_length=5, fixed_samples=100, val_split=0.1,
                batch_size=16, noise_level=0.1, quotas={"safe": 0.3, "unsafe": 0.3, "boundary": 0.2, "goal": 0.1})
    args.update(kw)
    return cls(dyn, [(float(a), float(b)) for a, b in zip(lo, up)], **args)


def _seed(s):
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


def _load_reference_datamodule(monkeypatch):
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
    return mod


# ---------------------------------------------------------------------------------------------------- datasets
def test_dataset_dict_format_and_roundtrip(stack, tmp_path):
    from cbfrrt.neural_cbf.datamodules import EpisodicDataModule
    env, dyn = stack()
    _seed(0)
    dm = _datamodule(EpisodicDataModule, dyn, data_root=str(tmp_path))
    path = dm.dataset_path()
    # episodic_datamodule.py:203-218: <thr*100>_<obstacles>_<noise*100>_<episodes>_<traj>_<len>_<fixed>.pt
    assert path == os.path.join(str(tmp_path), "models", "neural_cbf", str(dyn), "data", "5_4_10_2_6_5_100.pt")
    assert "mindis" in str(dyn)
    dm.prepare_data()
    assert os.path.isfile(path)
    data = torch.load(path)
    assert set(data) == {"x_training", "x_validation", "x_training_mask", "x_validation_mask"}      # :249-252
    n_rows = 2 * 90 + 2 * 6 * 5                         # fixed quota (90 of 100 per episode) + trajectories
    n_val = int(n_rows * 0.1)
    assert data["x_training"].shape == (n_rows - n_val, 9) and data["x_validation"].shape == (n_val, 9)
    assert data["x_training"].dtype == torch.float32
    for part in ("x_training_mask", "x_validation_mask"):
        assert set(data[part]) == {"safe", "unsafe", "goal", "boundary"}
        for v in data[part].values():
            assert v.dtype == torch.bool and v.shape == (data[part.replace("_mask", "")].shape[0],)
        safe, unsafe, bnd = data[part]["safe"], data[part]["unsafe"], data[part]["boundary"]
        assert not (safe & unsafe).any() and torch.equal(bnd, ~(safe | unsafe))
        assert not (data[part]["goal"] & ~safe).any()
    # the (x, goal, safe, unsafe) tensor datasets of :306-319 and their loaders
    xb, g, s, u = next(iter(dm.train_dataloader()))
    assert xb.shape == (16, 9) and g.dtype == s.dtype == u.dtype == torch.bool
    assert len(dm.validation_data) == n_val
    # second module: finds the file and loads it instead of sampling again (:219-224)
    dm2 = _datamodule(EpisodicDataModule, dyn, data_root=str(tmp_path))
    dm2.sample_fixed = dm2.sample_trajectories = None
    dm2.prepare_data()
    assert torch.equal(dm2.x_training, dm.x_training) and torch.equal(dm2.x_validation_mask["boundary"],
                                                                      dm.x_validation_mask["boundary"])


def test_stored_rows_are_consistent_with_the_path(stack, tmp_path, coracle):
    """every stored row is [q | o | do/dq] of its own q in the scene it was drawn in, and its labels are that scene's"""
    from cbfrrt.neural_cbf.datamodules import EpisodicDataModule
    env, dyn = stack()
    _seed(3)
    dm = _datamodule(EpisodicDataModule, dyn, max_episode=1, data_root=str(tmp_path))
    samples, masks = dm.sample_fixed()             # scene = get_env_config(0) (or the next one that could be filled)
    assert samples.shape == (90, 9)
    assert masks["safe"][:30].all() and masks["unsafe"][30:60].all() and masks["boundary"][60:80].all()
    again = dyn.complete_sample_with_observations(samples[:, :4], samples.shape[0])
    assert torch.equal(again, samples)
    assert torch.equal(dyn.safe_mask(samples), masks["safe"]) and torch.equal(dyn.goal_mask(samples), masks["goal"])
    # goal-quota rows lie within eps = 0.02 of the goal in every joint (arm_dynamics.py:293-305)
    assert (samples[80:, :4] - torch.tensor(dyn.goal_state).float()).abs().max() <= 0.02 + 1e-6


def test_noisy_simulator_semantics(stack):
    """arm_dynamics.py:382-471: (bs, steps, 2n+1); 10 Euler sub-steps per stored state; control = clamp(nominal +
    noise * upper limit) re-drawn every controller_period / dt steps; motor control is refused (needs physics)."""
    env, dyn = stack()
    _seed(1)
    dyn.set_goal(dyn.sample_state_space(1)[0, :4].numpy())
    x0 = dyn.sample_safe(5)[:, :4]
    _seed(2)
    traj = dyn.noisy_simulator(x0, 4, collect_dataset=True, noise_level=0.0)
    assert traj.shape == (5, 4, 9)
    assert torch.equal(traj[:, 0], dyn.complete_sample_with_observations(x0, 5))
    u = dyn.u_nominal(traj[:, 0])
    assert torch.allclose(traj[:, 1, :4], x0 + u * dyn.dt * 10, atol=1e-6)
    # controller_period defaults to dt (:437-439) -> the control is recomputed for every stored state
    assert torch.allclose(traj[:, 2, :4], traj[:, 1, :4] + dyn.u_nominal(traj[:, 1]) * dyn.dt * 10, atol=1e-6)
    _seed(2)
    noisy = dyn.noisy_simulator(x0, 3, collect_dataset=True, noise_level=0.5)
    up, lo = dyn.control_limits
    du = (noisy[:, 1, :4] - x0) / (dyn.dt * 10)
    assert (du <= up + 1e-4).all() and (du >= lo - 1e-4).all() and not torch.allclose(du, u, atol=1e-3)
    with pytest.raises(AssertionError):
        dyn.noisy_simulator(x0, 3, collect_dataset=False)
    with pytest.raises(NotImplementedError):
        dyn.noisy_simulator(x0, 3, collect_dataset=True, use_motor_control=True)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not available")
def test_reference_datamodule_draws_the_same_dataset(stack, monkeypatch):
    """the reference's EpisodicDataModule (loaded unmodified) and the mirror, same seeds, same dynamics model:
    identical fixed samples, trajectories and labels"""
    from cbfrrt.neural_cbf.datamodules import EpisodicDataModule
    ref = _load_reference_datamodule(monkeypatch)
    env, dyn = stack()
    # the reference flips a coin for motor control (:120); keep its coin on "kinematic" -- this path has no physics step
    real_random = random.random
    out = {}
    for name, cls in (("ref", ref.EpisodicDataModule), ("ours", EpisodicDataModule)):
        dm = _datamodule(cls, dyn)
        dyn.set_goal(np.array(dyn.robot.q0))
        _seed(11)
        fixed = dm.sample_fixed()
        if name == "ref":
            monkeypatch.setattr(random, "random", lambda: (real_random(), 0.99)[1])   # draws, then says "kinematic"
        _seed(12)
        sim = dm.sample_trajectories(dyn.noisy_simulator)
        monkeypatch.setattr(random, "random", real_random)
        out[name] = (fixed, sim)
    for part in (0, 1):
        (xr, mr), (xo, mo) = out["ref"][part], out["ours"][part]
        assert xr.shape == xo.shape and torch.equal(xr, xo)
        for k in ("safe", "unsafe", "goal", "boundary"):
            assert torch.equal(mr[k], mo[k]), k
    assert out["ours"][1][0].shape == (2 * 6 * 5, 9)


# ---------------------------------------------------------------------------------------------------- problem sets
def test_problem_pickles_roundtrip_and_feed_the_env(stack, tmp_path):
    from cbfrrt.motion_planning.evaluation import (propose_test_cases, generate_refined_test_cases,
                                                   read_problem_stream, collect_problems, save_collected,
                                                   get_test_cases)
    env, dyn = stack()
    _seed(4)
    pos, siz, start, goal = propose_test_cases(dyn, obstacle_num=5)
    assert pos.shape == siz.shape == (5, 3) and start.shape == goal.shape == (4,)
    assert (siz >= 0.05).all() and (siz <= 0.25).all() and (np.abs(pos[:, :2]) <= 0.5).all()
    assert np.linalg.norm(start - goal) > 0.7                                   # generate_problem_env.py:89
    env.reset_env(obs_configs=(pos, siz), enable_object=False)
    both = torch.tensor(np.stack([start, goal]), dtype=torch.float32)
    assert dyn.safe_mask(both).all()
    low = pos[:, 2] - siz[:, 2] < 0.3                                            # :79-81 "far from base"
    near = (np.abs(np.abs(pos[:, 0]) - siz[:, 0]) < 0.1) | (np.abs(np.abs(pos[:, 1]) - siz[:, 1]) < 0.1)
    assert not (low & near).any()

    times = iter([1.0, 0.0, 3.0, 0.7, 1.5, 1.2])

    def solver(env_, init_state, goal_state, dynamics_model):   # stands for BIT* (generate_problem_env.py:46-48)
        t = next(times)
        return {"success": t > 0, "path": [init_state, goal_state], "explored_nodes": 7, "time_dict": {"total_time": t}}

    raw = str(tmp_path / "magician_5_5_vsupersafe_1234_refined.pkl")
    n = generate_refined_test_cases(dyn, solver, raw, problem_num=5, obstacle_num=5, seed=1234)
    assert n == 5
    records = read_problem_stream(raw)
    assert len(records) == 5 and [r["time_dict"]["total_time"] for r in records] == [1.0, 3.0, 0.7, 1.5, 1.2]
    for r in records:                                                             # :51-56
        assert {"obstacle_positions", "obstacle_sizes", "init_config", "goal_config", "success", "path"} <= set(r)
        assert r["obstacle_positions"].shape == (5, 3) and r["init_config"].shape == (4,)
    # the stream is plain consecutive pickles: the reference's reader (collect_problems.py:12-18) reads it
    with open(raw, "rb") as fh:
        first = pickle.load(fh)
    assert np.array_equal(first["goal_config"], records[0]["goal_config"])
    collected = collect_problems(records, 0.5, 2.0)                               # collect_problems.py:20-30
    assert len(collected["init_config"]) == 4 and set(collected) == set(records[0])
    refined = str(tmp_path / "magician_easy.pkl")
    save_collected(refined, collected)
    p, s, a, b = get_test_cases(refined, start_idx=1, end_idx=3)                  # eval_rrt_all.py:74-104
    assert len(p) == len(s) == len(a) == len(b) == 2 and np.array_equal(a[0], records[2]["init_config"])
    with pytest.raises(ValueError):
        get_test_cases(str(tmp_path / "missing.pkl"))
    # plural key spelling (eval_rrt_all.py:85-87) and the .npz form of the same keys
    np.savez(str(tmp_path / "alt.npz"), obstacle_positions=np.stack(collected["obstacle_positions"]),
             obstacle_sizes=np.stack(collected["obstacle_sizes"]), init_configs=np.stack(collected["init_config"]),
             goal_configs=np.stack(collected["goal_config"]))
    p2, _, a2, _ = get_test_cases(str(tmp_path / "alt.npz"), 0, 4)
    assert np.array_equal(a2[3], collected["init_config"][3]) and p2.shape == (4, 5, 3)
    # ArmEnv(config_file=<refined pickle>) serves the scenes by index (arm_env.py:55,103-111)
    env2, dyn2 = stack(config_file=refined)
    q_pos, q_siz = env2.get_env_config(2)
    assert np.array_equal(q_pos, collected["obstacle_positions"][2]) and np.array_equal(q_siz, collected["obstacle_sizes"][2])
    env2.reset_env(obs_configs=(q_pos, q_siz), enable_object=False)
    assert env2.obs_positions.shape == (5, 3) and len(env2.obstacle_ids) == 6


# ---------------------------------------------------------------------------------------------------- checkpoints
def test_lightning_checkpoint_loads(stack, tmp_path):
    from cbfrrt.neural_cbf.controllers import NeuralMindisCBFController
    env, dyn = stack("panda")
    torch.manual_seed(7)
    src = NeuralMindisCBFController(dyn, [{"m1": 5.76}], None, None, **KW)
    # what Lightning 1.3 writes: state_dict with the module's parameter names + hyper_parameters (+ bookkeeping)
    ckpt = {"epoch": 3, "global_step": 100, "pytorch-lightning_version": "1.3.4",
            "state_dict": {k: v.clone() for k, v in src.state_dict().items()},
            "hyper_parameters": {"cbf_hidden_layers": 2, "cbf_hidden_size": 48, "cbf_alpha": 1.0,
                                 "cbf_relaxation_penalty": 500.0, "learning_rate": 1e-3, "batch_size": 64}}
    assert set(ckpt["state_dict"]) == {f"h_nn.{n}.{p}" for n in ("input_linear", "layer_0_linear", "layer_1_linear",
                                                                  "output_linear") for p in ("weight", "bias")}
    path = str(tmp_path / "epoch=3.ckpt")
    torch.save(ckpt, path)
    torch.manual_seed(8)
    ctrl = NeuralMindisCBFController.load_from_checkpoint(      # the call of eval_rrt_mindis_cbfsteer.py:38-50
        path, dynamics_model=dyn, scenarios=[{"m1": 5.76}], datamodule=None, experiment_suite=None,
        cbf_relaxation_penalty=5000, cbf_alpha=0.1, safe_level=0.1, unsafe_level=0.1, cbf_hidden_layers=2,
        cbf_hidden_size=48, use_neural_actor=False)
    ctrl.to("cpu")
    assert ctrl.cbf_alpha == 0.1 and ctrl.clf_relaxation_penalty == 5000      # keyword arguments win
    for (k, a), (_, b) in zip(src.state_dict().items(), ctrl.state_dict().items()):
        assert torch.equal(a, b), k
    x = dyn.sample_safe(8)
    assert torch.equal(ctrl.h(x), src.h(x))
    u1, u2 = ctrl.u(x)[0], src.u(x)[0]
    assert torch.equal(u1, u2)
    # hyper-parameters stored in the checkpoint are used when the caller does not override them
    ctrl2 = NeuralMindisCBFController.load_from_checkpoint(path, dynamics_model=dyn, scenarios=[{"m1": 5.76}])
    assert ctrl2.cbf_alpha == 1.0 and ctrl2.clf_relaxation_penalty == 500.0 and ctrl2.h_hidden_size == 48
    torch.save({"weights": 1}, str(tmp_path / "bad.ckpt"))
    with pytest.raises(KeyError):
        NeuralMindisCBFController.load_from_checkpoint(str(tmp_path / "bad.ckpt"), dynamics_model=dyn,
                                                       scenarios=[{"m1": 5.76}], **KW)


# ---------------------------------------------------------------------------------------------------- evaluation
def test_planner_evaluation_over_a_problem_file(stack, tmp_path):
    """eval_rrt_all.py:20-71 / eval_batchrrt_vanilla.py / eval_batchrrt_mindis_cbfsteer.py on a refined problem pickle
    written by this package: per-problem protocol, result list with the reference's keys, summary, result pickle"""
    from cbfrrt.motion_planning.evaluation import (generate_refined_test_cases, read_problem_stream, collect_problems,
                                                   save_collected, eval_rrt_all, eval_batchrrt_mindis, summarize)
    from cbfrrt.motion_planning.baseline import batch_RRT_plan
    from cbfrrt.neural_cbf.controllers import NeuralMindisCBFController
    env, dyn = stack()

    def solver(env_, init_state, goal_state, dynamics_model):          # the filter of generate_problem_env.py:46-48
        return batch_RRT_plan(env_, dynamics_model, init_state, goal_state, RRT_PARAM=20, T=200, model_eps=0.3,
                              steer_type="line", batch=10, stop_when_success=True, rng=np.random.RandomState(0))

    raw = str(tmp_path / "raw.pkl")
    n = generate_refined_test_cases(dyn, lambda *a: {k: v for k, v in solver(*a).items() if k != "search_tree"}, raw,
                                    problem_num=3, obstacle_num=4, seed=7, max_proposals=12)
    assert n >= 1
    refined = str(tmp_path / "magician_easy.pkl")
    save_collected(refined, collect_problems(read_problem_stream(raw), 0.0, 1e9))
    out = str(tmp_path / "time_magician_easy_23_20_20_0_100.pkl")
    result, summary = eval_rrt_all("magician", "batchrrt_vanilla", refined, out_file=out, seed=23, device="cpu",
                                   simulation_dt=1 / 120, controller_period=1 / 30, RRT_STEP=20, goal_biasing=0.3,
                                   max_node=1500, start_idx=0, end_idx=100)
    assert len(result) == n and summary["problems"] == n and 0.0 <= summary["success_rate"] <= 1.0
    for sol in result:
        assert set(sol) == {"success", "path", "edge_states", "nominal_path", "path_cost", "explored_nodes", "time_dict"}
        assert sol["explored_nodes"] >= 1 and sol["time_dict"]["total_time"] > 0
    assert summary["success_rate"] > 0                 # problems the line-steer planner solved once are solved again
    import pickle
    with open(out, "rb") as fh:
        assert len(pickle.load(fh)) == n
    # CBF steer with a caller-built controller (no checkpoint ships with the reference) and from a checkpoint directory
    env2, dyn2 = stack(config_file=refined)
    torch.manual_seed(1)
    ctrl = NeuralMindisCBFController(dyn2, [{"name": 1}], None, None, **KW)
    from cbfrrt.motion_planning.evaluation import get_test_cases
    p, s, a, b = get_test_cases(refined, 0, 1)
    kw = dict(simulation_dt=1 / 120, controller_period=1 / 30, RRT_STEP=20, goal_biasing=0.3, max_node=40, start_idx=0,
              end_idx=1, cbf_alpha=0.1)
    res = eval_batchrrt_mindis(3, env2, env2.robot_list[0], None, p, s, a, b, controller=ctrl, batch=5, **kw)
    assert len(res) == 1 and res[0]["explored_nodes"] >= 5 and summarize(res)["problems"] == 1
    os.makedirs(tmp_path / "ckpt")
    torch.save({"state_dict": ctrl.state_dict()}, str(tmp_path / "ckpt" / "epoch=1.ckpt"))
    res2 = eval_batchrrt_mindis(3, env2, env2.robot_list[0], str(tmp_path / "ckpt"), p, s, a, b, batch=5, **kw)
    assert res2[0]["explored_nodes"] == res[0]["explored_nodes"] and res2[0]["success"] == res[0]["success"]

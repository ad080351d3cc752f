"""Dev-container-only (skipped where /root/reference is absent): the reference's own experiment classes -- callers of the
path named by SURVEY.md 8b (neural_cbf/experiments/rollout_state_space_experiment.py:155,213, bf_contour_experiment.py) and
wired up by neural_cbf/evaluation/eval_arm_mindis.py:24-86 -- are loaded unmodified and run on THIS package's dynamics
model and controller (CUDA ops replaced by the oracle-backed fakes).  plot() needs matplotlib / seaborn and is not run."""
import os
import sys

import numpy as np
import pytest
import torch

import fake_ops

REF = os.environ.get("CBFRRT_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "neural_cbf", "experiments", "experiment.py")),
                                reason="reference tree not available")


@pytest.fixture()
def ref_experiments():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as mg
    finder = mg._StubFinder()
    before = set(sys.modules)
    sys.meta_path.insert(0, finder)
    sys.path.insert(0, REF)
    try:
        from neural_cbf.experiments import RolloutStateSpaceExperiment, BFContourExperiment, ExperimentSuite
        yield RolloutStateSpaceExperiment, BFContourExperiment, ExperimentSuite
    finally:
        sys.meta_path.remove(finder)
        sys.path.remove(REF)
        for k in set(sys.modules) - before:       # drop the reference's packages and the stubs again
            if k.split(".")[0] in ("neural_cbf", "environment", "motion_planning", "utils") or \
                    any(k == r or k.startswith(r + ".") for r in mg.STUB_ROOTS):
                del sys.modules[k]


def test_reference_experiments_run_on_our_objects(monkeypatch, coracle, ref_experiments):
    Rollout, Contour, Suite = ref_experiments
    fake_ops.install(monkeypatch)
    from cbfrrt.environment import ArmEnv
    from cbfrrt.neural_cbf.systems import ArmMindis
    from cbfrrt.neural_cbf.controllers import NeuralMindisCBFController
    env = ArmEnv(["magician"], config_file="", include_floor=True, device="cpu")
    robot = env.robot_list[0]
    dyn = ArmMindis({"m1": 5.76}, dt=1 / 120, controller_dt=1 / 30, dis_threshold=0.02, env=env, robot=robot)
    dyn.set_goal(torch.tensor(robot.q0).float())                                   # eval_arm_mindis.py:42-43
    ul, ll = dyn.state_limits
    start_x = torch.lerp(ll.double(), ul.double(), 0.2 * torch.ones(ll.shape[-1]).double()).reshape(1, -1).float()
    start_x = dyn.complete_sample_with_observations(start_x, num_samples=1)         # :45-49
    rollout = Rollout("Rollout", start_x, 2, "theta_2", 0, "theta_0", scenarios=[{"m1": 5.76}],
                      n_sims_per_start=2, t_sim=0.1)
    contour = Contour("h_Contour", domain=[tuple(robot.body_range[2]), tuple(robot.body_range[0])], n_grid=4,
                      x_axis_index=2, y_axis_index=0, x_axis_label="a", y_axis_label="b",
                      default_state=dyn.sample_boundary(1).squeeze(), plot_unsafe_region=True)
    suite = Suite([rollout, contour])
    torch.manual_seed(1)
    ctrl = NeuralMindisCBFController(dyn, [{"m1": 5.76}], None, suite, cbf_alpha=0.1, cbf_relaxation_penalty=5000.0,
                                     safe_level=0.1, unsafe_level=0.1, cbf_hidden_layers=2, cbf_hidden_size=48)
    df = rollout.run(ctrl)
    steps = int(0.1 // dyn.dt)
    assert len(df) == 2 * steps and {"t", "Simulation", "state", "h", "m1"} <= set(df.columns)
    # the logged states are the closed loop of THIS package: control every controller_dt / dt steps, Euler in between
    x = start_x.clone()
    for t in range(steps):
        if t % int(ctrl.controller_period / dyn.dt) == 0:
            u = ctrl.u(x)[0]
        row = df[(df["Simulation"] == "0") & (np.isclose(df["t"], t * dyn.dt))].iloc[0]
        # (the logged "state" arrays alias the experiment's working tensor on the CPU, so only the scalars are compared)
        assert abs(row["theta_2"] - float(x[0, 2])) < 1e-6 and abs(row["theta_0"] - float(x[0, 0])) < 1e-6
        assert abs(row["h"] - float(ctrl.h(x))) < 1e-6
        x = dyn.closed_loop_dynamics(x, u)
    cdf = contour.run(ctrl)
    assert len(cdf) == 16 and np.isfinite(np.asarray(cdf.iloc[:, -1], dtype=float)).all()

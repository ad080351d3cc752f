"""Synthetic workloads of BASELINE.json (`configs[0..4]`), as fixed by SURVEY.md section 8d.

Everything is generated from seeds with numpy / torch on the host (the reference ships no checkpoints, problem
sets or datasets); config 5 samples its states on the device with the counter-based sampler so that results do
not depend on the shard count.  Returned dicts hold plain numpy arrays so that the CUDA path, the oracle and the
host-buffer C-ABI can all consume the same inputs.  This module touches neither the CUDA library nor torch.cuda
(bench.py's --impl reference arm loads it without mapping libcbfrrt.so).

Deviations from the reference (stated in every result): sphere obstacles, 64/256-obstacle scenes, batches
>= 4096, seeded random MLP weights, a fixed number of interpolation points per edge.

Obstacle distributions (`OBSTACLES` below; also printed in bench.py's `config`).  Box half extents follow SURVEY 8d /
generate_problem_env.py:62-78 (U(.05,.25) m; spheres r ~ U(.05,.2)) wherever the scene stays usable.  The PLACEMENT
region is the reference's 1 m^3 box U([-.5,-.5,0],[.5,.5,1]) only for its own problem size (config 4: 8 boxes); with
16 / 64 / 256 obstacles that box leaves 0.03 % / 0.06 % / 0.00 % of the sampled configurations collision-free
(measured with the oracle on 20 000 states), so those scenes are spread over the arm's whole reach, and the
256-box scene uses half extents U(.02,.08).  Free-space fractions (safe_mask true, 20 000 states): cfg1 22 %,
cfg2 17 % of the states / 8 % of the edges, cfg3 12 %, cfg4 18 %.
"""
from collections import OrderedDict

import numpy as np
import torch

from ._robot_tables import ROBOT_MODELS

DT, CONTROLLER_DT = 1.0 / 120.0, 1.0 / 30.0   # eval_rrt_mindis_cbfsteer.py:24-31
THR_SOFT, THR_HARD = 0.02, 0.0                 # dis_threshold (eval_rrt_mindis_cbfsteer.py:27)
CBF_ALPHA, PENALTY = 0.1, 5000.0               # eval_rrt_all.py:128, eval_rrt_mindis_cbfsteer.py:43
HIDDEN, LAYERS = 48, 2                         # eval_rrt_mindis_cbfsteer.py:47-48


def make_state_dict(n: int, seed: int = 1, hidden: int = HIDDEN, layers: int = LAYERS):
    """nn.Linear default init of the reference's h_nn (neural_mindis_cbf_controller.py:44-53) under manual_seed."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    mods = OrderedDict()
    mods["input_linear"] = torch.nn.Linear(n + 1, hidden)
    mods["input_activation"] = torch.nn.ReLU()
    for i in range(layers):
        mods[f"layer_{i}_linear"] = torch.nn.Linear(hidden, hidden)
        mods[f"layer_{i}_activation"] = torch.nn.ReLU()
    mods["output_linear"] = torch.nn.Linear(hidden, 1)
    net = torch.nn.Sequential(mods)
    torch.random.set_rng_state(g)
    return {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}, net


def scalar_dare_gain(b: float) -> float:
    """LQR gain of x+ = x + b u, Q = R = 1 (control_affine_system.py:125-157, utils.py:18-49): same closed form as
    cbfrrt.neural_cbf.systems.utils.scalar_dare_gain, repeated here so that this module stays free of the package's
    CUDA-backed imports."""
    b2 = b * b
    x = (b2 + np.sqrt(b2 * b2 + 4.0 * b2)) / (2.0 * b2)
    return b * x / (1.0 + b2 * x)


def packed_weight_count(n: int, hidden: int, layers: int) -> int:
    in_pad = (n + 1 + 3) & ~3
    return hidden * in_pad + hidden + layers * (hidden * hidden + hidden) + hidden + 4


def pack_weights_np(sd, n: int, hidden: int = HIDDEN, layers: int = LAYERS) -> np.ndarray:
    """h_nn state_dict (keys input_linear / layer_{i}_linear / output_linear, with or without the ``h_nn.`` prefix;
    neural_mindis_cbf_controller.py:44-53) -> the packed float32 buffer described in include/cbfrrt.h."""
    def get(name):
        for k in (name, "h_nn." + name):
            if k in sd:
                v = sd[k]
                return v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        raise KeyError(name)

    in_pad = (n + 1 + 3) & ~3
    out = np.zeros(packed_weight_count(n, hidden, layers), np.float32)
    w_in = np.zeros((hidden, in_pad), np.float32)
    w_in[:, :n + 1] = get("input_linear.weight")
    o = 0
    out[o:o + hidden * in_pad] = w_in.ravel(); o += hidden * in_pad
    out[o:o + hidden] = get("input_linear.bias"); o += hidden
    for i in range(layers):
        out[o:o + hidden * hidden] = get(f"layer_{i}_linear.weight").ravel(); o += hidden * hidden
        out[o:o + hidden] = get(f"layer_{i}_linear.bias"); o += hidden
    out[o:o + hidden] = get("output_linear.weight").ravel(); o += hidden
    out[o] = float(np.asarray(get("output_linear.bias")).ravel()[0])
    return out


def sample_states(robot: str, B: int, seed: int) -> np.ndarray:
    """U(lower, upper) per joint like ControlAffineSystem.sample_state_space (control_affine_system.py:291-300)."""
    m = ROBOT_MODELS[robot]
    lo, hi = np.array(m["q_lower"], np.float32), np.array(m["q_upper"], np.float32)
    rng = np.random.default_rng(seed)
    return (lo + (hi - lo) * rng.random((B, m["n"]), dtype=np.float32)).astype(np.float32)


REF_REGION = ((-0.5, -0.5, 0.0), (0.5, 0.5, 1.0))      # generate_problem_env.py:65-78
PANDA_REACH = ((-1.2, -1.2, 0.0), (1.2, 1.2, 1.5))     # shoulder at z = 0.333 m + 0.855 m reach + hand
MAGICIAN_REACH = ((-0.8, -0.8, 0.0), (0.8, 0.8, 1.0))  # 2 x scaled Dobot (environment/magician.py:20)

# per BASELINE config: (n_boxes, box half-extent range, n_spheres, sphere radius range, placement region)
OBSTACLES = {
    1: dict(n_boxes=16, h=(0.05, 0.25), n_spheres=0, r=(0.0, 0.0), region=MAGICIAN_REACH),
    2: dict(n_boxes=32, h=(0.05, 0.25), n_spheres=32, r=(0.05, 0.2), region=PANDA_REACH),
    3: dict(n_boxes=8, h=(0.05, 0.25), n_spheres=0, r=(0.0, 0.0), region=REF_REGION),
    4: dict(n_boxes=256, h=(0.02, 0.08), n_spheres=0, r=(0.0, 0.0), region=PANDA_REACH),
}


def obstacle_note(idx: int) -> str:
    if idx not in OBSTACLES:
        return "floor + first default box of arm_env.py:113-114"
    o = OBSTACLES[idx]
    s = f"{o['n_boxes']} boxes, half extents U{o['h']} m"
    if o["n_spheres"]:
        s += f" + {o['n_spheres']} spheres, r U{o['r']} m"
    return s + f", centres U({o['region'][0]}, {o['region'][1]}), obstacles overlapping the base column rejected"


def random_boxes(n_boxes: int, seed: int, h_lo: float, h_hi: float, region=REF_REGION) -> np.ndarray:
    """centres U(region) (default: generate_problem_env.py:65-78); half extents U(h_lo, h_hi); boxes that
    would swallow the robot base (a 0.2 m column up to z = 0.4) are rejected.  -> (n_boxes, 6)"""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n_boxes:
        c = rng.uniform(region[0], region[1])
        h = rng.uniform(h_lo, h_hi, size=3)
        if abs(c[0]) - h[0] < 0.2 and abs(c[1]) - h[1] < 0.2 and c[2] - h[2] < 0.4:
            continue
        out.append(np.concatenate([c, h]))
    return np.array(out, dtype=np.float32).reshape(-1, 6)


def random_spheres(n_sph: int, seed: int, r_lo: float, r_hi: float, region=REF_REGION) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n_sph:
        c = rng.uniform(region[0], region[1])
        r = rng.uniform(r_lo, r_hi)
        if np.hypot(c[0], c[1]) - r < 0.2 and c[2] - r < 0.4:
            continue
        out.append(np.concatenate([c, [r]]))
    return np.array(out, dtype=np.float32).reshape(-1, 4)


def scene(idx: int):
    """(boxes (Ob,6), spheres (Os,4)) of BASELINE config idx (seeds 2 / 5 as in SURVEY 8d)"""
    o = OBSTACLES[idx]
    boxes = random_boxes(o["n_boxes"], 2, o["h"][0], o["h"][1], o["region"])
    sph = random_spheres(o["n_spheres"], 5, o["r"][0], o["r"][1], o["region"]) if o["n_spheres"] else np.zeros((0, 4), np.float32)
    return boxes, sph


def controller_constants(robot: str):
    m = ROBOT_MODELS[robot]
    n = m["n"]
    K = (scalar_dare_gain(CONTROLLER_DT) * np.eye(n)).astype(np.float32)
    v = np.array(m["v_max"], np.float32)
    return dict(K=K, u_lo=-v, u_hi=v, lam=CBF_ALPHA, penalty=PENALTY, thr_soft=THR_SOFT, thr_hard=THR_HARD,
                hidden=HIDDEN, layers=LAYERS, dt=DT, ctrl_every=int(CONTROLLER_DT // DT))


def config(idx: int, scale: float = 1.0):
    """BASELINE.json configs[idx]; ``scale`` shrinks the number of units (tests / CPU baselines)."""
    if idx == 0:
        robot, B = "panda", max(1, int(4096 * scale))
        w = dict(robot=robot, name="cfg0_panda_1box_4096", q=sample_states(robot, B, 0),
                 boxes=np.array([[0.3, 0.17, 0.3, 0.05, 0.05, 0.1]], np.float32), spheres=np.zeros((0, 4), np.float32),
                 floor=True, goal=np.array(ROBOT_MODELS[robot]["q0"], np.float32).reshape(1, -1))
    elif idx == 1:
        robot, B = "magician", max(1, int((1 << 20) * scale))
        w = dict(robot=robot, name="cfg1_magician_16obs_1M", q=sample_states(robot, B, 0),
                 boxes=scene(1)[0], spheres=scene(1)[1], floor=True,
                 goal=sample_states(robot, 1, 3))
    elif idx == 2:
        robot, E = "panda", max(1, int((1 << 22) * scale))
        q_from = sample_states(robot, E, 0)
        m = ROBOT_MODELS[robot]
        lo, hi = np.array(m["q_lower"], np.float32), np.array(m["q_upper"], np.float32)
        step = np.random.default_rng(1).uniform(-0.35, 0.35, size=q_from.shape).astype(np.float32)
        w = dict(robot=robot, name="cfg2_panda_64obs_4Medges_x32", q=q_from, q_to=np.clip(q_from + step, lo, hi),
                 n_interp=32, boxes=scene(2)[0], spheres=scene(2)[1],
                 floor=True, goal=np.array(m["q0"], np.float32).reshape(1, -1))
    elif idx == 3:
        robot, T = "panda", max(1, int((1 << 18) * scale))
        w = dict(robot=robot, name="cfg3_panda_rollout_262k_x200", q=sample_states(robot, T, 0),
                 goal=sample_states(robot, T, 1), steps=200, boxes=scene(3)[0], spheres=scene(3)[1], floor=True)
    elif idx == 4:
        robot, B = "panda", max(1, int((1 << 26) * scale))
        w = dict(robot=robot, name="cfg4_panda_256obs_64M", q=None, n_states=B, sample_seed=0,
                 boxes=scene(4)[0], spheres=scene(4)[1], floor=True,
                 goal=np.array(ROBOT_MODELS[robot]["q0"], np.float32).reshape(1, -1))
    else:
        raise ValueError(idx)
    n = ROBOT_MODELS[w["robot"]]["n"]
    sd, _ = make_state_dict(n, seed=1)
    w.update(controller_constants(w["robot"]))
    w["state_dict"] = sd
    w["weights"] = pack_weights_np(sd, n)
    w["n"] = n
    w["obstacles"] = obstacle_note(idx)
    return w

"""PyTorch custom ops (namespace ``cbfrrt::``) over the C-ABI of libcbfrrt.so.

CUDA-only registration: there is no CPU kernel, no Triton path and no dispatch table -- calling an op with CPU
tensors raises NotImplementedError.  PyTorch is plumbing here (device memory, current stream); every op is one call
into include/cbfrrt.h with raw device pointers.

Every op is registered with torch.library (``torch.ops.cbfrrt.<name>``, also ``ops.registered[<name>]``); the
module-level functions of the same names call the implementations DIRECTLY, because the dispatcher costs 20-30 us
per call -- most of the host time of a planner-sized (1-50 state) call.

Reference call sites each op replaces are listed in include/cbfrrt.h.
"""
import ctypes as C
import functools
from typing import Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib
from ._lib import lib, check, Scene, CbfNet, FilterParams, RolloutParams, PeerOut, LidarNet

ROBOT_ID = _lib.ROBOT_ID
ROBOT_N = {0: 7, 1: 4}
ROBOT_LINKS = {0: 12, 1: 4}


# ----------------------------------------------------------------------------- helpers
def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else C.c_void_p(0)


def _rows(q: Tensor, n: int):
    """float32 rows with unit inner stride; returns (tensor, row_stride) without copying views like x[:, :n]."""
    if q.dtype != torch.float32:
        q = q.float()
    if q.dim() != 2 or q.shape[1] < n:
        raise ValueError(f"expected a (B, >={n}) tensor, got {tuple(q.shape)}")
    if q.stride(1) != 1 or (q.shape[0] > 1 and q.stride(0) < n):
        q = q.contiguous()
    return q, (q.stride(0) if q.shape[0] > 1 else max(q.shape[1], n))


def _f32c(t: Tensor) -> Tensor:
    return t.to(dtype=torch.float32).contiguous()


# scratch of the two-pass sweeps (include/cbfrrt.h cbfrrt_set_workspace): one tensor per device, grown on demand.  The
# library never allocates device memory itself; this layer owns the buffer and registers it.
SPLIT_MIN_ROWS = 1 << 19
_WORKSPACE = {}


def _ensure_workspace(dev: torch.device, rows: int, n: int, hidden: int, layers: int) -> None:
    """room for the observation rows (+ mask bytes) of a sweep of ``rows`` states between its two passes"""
    if rows < SPLIT_MIN_ROWS or (hidden, layers) != (48, 2):
        return
    need = 16 + (4 * (2 * n + 1) + 1) * rows + 16
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    ws = _WORKSPACE.get(idx)
    if ws is None or ws.numel() < need:
        ws = None
        _WORKSPACE.pop(idx, None)
        with torch.cuda.device(dev):
            check(lib.cbfrrt_set_workspace(None, 0), "cbfrrt_set_workspace")
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            check(lib.cbfrrt_set_workspace(C.c_void_p(ws.data_ptr()), ws.numel()), "cbfrrt_set_workspace")
        _WORKSPACE[idx] = ws


def set_two_pass(enabled: bool) -> None:
    """False: large sweeps always run the single fused kernel (A/B measurements, tests); True (default): the two-pass form
    from 524 288 rows on (include/cbfrrt.h cbfrrt_set_two_pass)"""
    check(lib.cbfrrt_set_two_pass(int(bool(enabled))), "cbfrrt_set_two_pass")


def build_grid(boxes8: Tensor, spheres: Tensor) -> Tensor:
    """Nearest-candidate grid of a packed scene (include/cbfrrt.h cbfrrt_grid_build).  Returns a uint8 device
    tensor to be passed as ``grid`` to the scene-consuming ops; an EMPTY tensor means brute force.  Results are
    bit-identical either way."""
    dev = boxes8.device
    g = torch.empty(int(lib.cbfrrt_grid_bytes()), dtype=torch.uint8, device=dev)
    sc = _scene(boxes8, spheres, 0, torch.empty(0, dtype=torch.uint8, device=dev))
    with torch.cuda.device(dev):
        check(lib.cbfrrt_grid_build(C.byref(sc), _ptr(g), _stream()), "cbfrrt_grid_build")
    return g


def no_grid(device) -> Tensor:
    return torch.empty(0, dtype=torch.uint8, device=device)


def _scene(boxes8: Tensor, spheres: Tensor, floor: int, grid: Tensor, stats: Tensor = None) -> Scene:
    if boxes8.numel() and (boxes8.dim() != 2 or boxes8.shape[1] != 8):
        raise ValueError("boxes must be packed as (Ob, 8): cx cy cz hx hy hz 0 0 (use pack_scene)")
    if spheres.numel() and (spheres.dim() != 2 or spheres.shape[1] != 4):
        raise ValueError("spheres must be (Os, 4): cx cy cz r")
    if grid.numel() and (grid.dtype != torch.uint8 or grid.numel() != int(lib.cbfrrt_grid_bytes())):
        raise ValueError("grid must be the uint8 tensor returned by build_grid (or empty)")
    return Scene(int(boxes8.shape[0]) if boxes8.numel() else 0, int(spheres.shape[0]) if spheres.numel() else 0,
                 int(bool(floor)), 0, _ptr(boxes8).value, _ptr(spheres).value, _ptr(grid).value,
                 _ptr(stats).value if stats is not None else None)


def pack_scene(positions, sizes, spheres=None, device="cuda"):
    """(positions (O,3), half-extents (O,3)) as held by ArmEnv.obs_positions / obs_sizes (arm_env.py:154-182)
    -> device tensors in the library's layout: boxes (O,8), spheres (P,4)."""
    positions = np.asarray(positions, dtype=np.float32).reshape(-1, 3)
    sizes = np.asarray(sizes, dtype=np.float32).reshape(-1, 3)
    b8 = np.zeros((positions.shape[0], 8), np.float32)
    b8[:, :3] = positions
    b8[:, 3:6] = sizes
    sp = np.zeros((0, 4), np.float32) if spheres is None else np.asarray(spheres, np.float32).reshape(-1, 4)
    return torch.from_numpy(b8).to(device), torch.from_numpy(np.ascontiguousarray(sp)).to(device)


def packed_weight_count(n: int, hidden: int, layers: int) -> int:
    in_pad = (n + 1 + 3) & ~3
    return hidden * in_pad + hidden + layers * (hidden * hidden + hidden) + hidden + 4


def pack_weights(state_dict, n: int, hidden: int, layers: int, device="cuda") -> Tensor:
    """h_nn state_dict (keys input_linear / layer_{i}_linear / output_linear, with or without the ``h_nn.``
    prefix; neural_mindis_cbf_controller.py:44-53) -> packed float32 buffer described in include/cbfrrt.h."""
    from .workloads import pack_weights_np
    return torch.from_numpy(pack_weights_np(state_dict, n, hidden, layers)).to(device)


def _net_params(n, W, K, lo, hi, goal, B, lam, penalty, hidden, layers, u_ref=None):
    W, K, lo, hi, goal = _f32c(W), _f32c(K), _f32c(lo), _f32c(hi), _f32c(goal).reshape(-1, n)
    if u_ref is not None and u_ref.numel() > 0:
        u_ref = _f32c(u_ref)
        if tuple(u_ref.shape) != (B, n):
            raise ValueError(f"u_ref must be ({B}, {n})")
    else:
        u_ref = None
    if W.numel() != packed_weight_count(n, hidden, layers):
        raise ValueError(f"packed weights have {W.numel()} floats, expected {packed_weight_count(n, hidden, layers)}")
    if K.numel() != n * n or lo.numel() != n or hi.numel() != n:
        raise ValueError("K must be (n,n); control limits (n,)")
    if goal.shape[0] not in (1, B):
        raise ValueError(f"goal must have 1 or {B} rows, got {goal.shape[0]}")
    net = CbfNet(n + 1, hidden, layers, 0, _ptr(W).value)
    fp = FilterParams(_ptr(goal).value, goal.shape[0], _ptr(K).value, _ptr(lo).value, _ptr(hi).value, float(lam),
                      float(penalty), _ptr(u_ref).value if u_ref is not None else None)
    return net, fp, (W, K, lo, hi, goal, u_ref)  # keep tensors alive until the call returns


# ----------------------------------------------------------------------------- ops
@torch.library.custom_op("cbfrrt::fk", mutates_args=(), device_types="cuda")
def fk(q: Tensor, robot_id: int) -> Tuple[Tensor, Tensor]:
    n, L = ROBOT_N[robot_id], ROBOT_LINKS[robot_id]
    q, qs = _rows(q, n)
    B = q.shape[0]
    pos = torch.empty((B, L, 3), dtype=torch.float32, device=q.device)
    rot = torch.empty((B, L, 3, 3), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        check(lib.cbfrrt_fk(robot_id, _ptr(q), qs, B, _ptr(pos), _ptr(rot), _stream()), "cbfrrt_fk")
    return pos, rot


@torch.library.custom_op("cbfrrt::collide", mutates_args=(), device_types="cuda")
def collide(q: Tensor, boxes: Tensor, spheres: Tensor, floor: int, grid: Tensor, thr_soft: float, thr_hard: float,
            robot_id: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> safe u8[B], unsafe u8[B], datax f32[B,2n+1], arg_link i32[B], arg_obs i32[B], mins f32[B,3]"""
    n = ROBOT_N[robot_id]
    q, qs = _rows(q, n)
    B, dev = q.shape[0], q.device
    safe = torch.empty(B, dtype=torch.uint8, device=dev)
    unsafe = torch.empty(B, dtype=torch.uint8, device=dev)
    datax = torch.empty((B, 2 * n + 1), dtype=torch.float32, device=dev)
    arg_link = torch.empty(B, dtype=torch.int32, device=dev)
    arg_obs = torch.empty(B, dtype=torch.int32, device=dev)
    mins = torch.empty((B, 3), dtype=torch.float32, device=dev)
    sc = _scene(boxes, spheres, floor, grid)
    with torch.cuda.device(dev):
        check(lib.cbfrrt_collide(robot_id, C.byref(sc), _ptr(q), qs, B, thr_soft, thr_hard, _ptr(safe), _ptr(unsafe),
                                 _ptr(datax), _ptr(arg_link), _ptr(arg_obs), _ptr(mins), _stream()), "cbfrrt_collide")
    return safe, unsafe, datax, arg_link, arg_obs, mins


@torch.library.custom_op("cbfrrt::observe", mutates_args=(), device_types="cuda")
def observe(q: Tensor, boxes: Tensor, spheres: Tensor, floor: int, grid: Tensor, thr_soft: float, thr_hard: float,
            robot_id: int) -> Tuple[Tensor, Tensor, Tensor]:
    """-> safe u8[B], unsafe u8[B], datax f32[B,2n+1]   (complete_sample_with_observations + masks)"""
    n = ROBOT_N[robot_id]
    q, qs = _rows(q, n)
    B, dev = q.shape[0], q.device
    safe = torch.empty(B, dtype=torch.uint8, device=dev)
    unsafe = torch.empty(B, dtype=torch.uint8, device=dev)
    datax = torch.empty((B, 2 * n + 1), dtype=torch.float32, device=dev)
    sc = _scene(boxes, spheres, floor, grid)
    with torch.cuda.device(dev):
        check(lib.cbfrrt_collide(robot_id, C.byref(sc), _ptr(q), qs, B, thr_soft, thr_hard, _ptr(safe), _ptr(unsafe),
                                 _ptr(datax), None, None, None, _stream()), "cbfrrt_collide")
    return safe, unsafe, datax


@torch.library.custom_op("cbfrrt::masks", mutates_args=(), device_types="cuda")
def masks(q: Tensor, boxes: Tensor, spheres: Tensor, floor: int, grid: Tensor, thr_soft: float, thr_hard: float,
          robot_id: int) -> Tuple[Tensor, Tensor]:
    """-> safe u8[B], unsafe u8[B]   (safe_mask / unsafe_mask only: no arg-min, no gradient)"""
    n = ROBOT_N[robot_id]
    q, qs = _rows(q, n)
    B, dev = q.shape[0], q.device
    safe = torch.empty(B, dtype=torch.uint8, device=dev)
    unsafe = torch.empty(B, dtype=torch.uint8, device=dev)
    sc = _scene(boxes, spheres, floor, grid)
    with torch.cuda.device(dev):
        check(lib.cbfrrt_collide(robot_id, C.byref(sc), _ptr(q), qs, B, thr_soft, thr_hard, _ptr(safe), _ptr(unsafe),
                                 None, None, None, None, _stream()), "cbfrrt_collide")
    return safe, unsafe


@torch.library.custom_op("cbfrrt::cbf_filter", mutates_args=(), device_types="cuda")
def cbf_filter(datax: Tensor, goal: Tensor, weights: Tensor, K: Tensor, u_lo: Tensor, u_hi: Tensor, lam: float,
               penalty: float, hidden: int, layers: int, u_ref: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """datax f32[B,2n+1] -> h f32[B], J f32[B,n], u f32[B,n], r f32[B].  u_ref: (B,n) caller-supplied reference
    control, or an empty tensor to use the nominal controller -K(q-goal)."""
    datax = _f32c(datax)
    if datax.dim() != 2 or datax.shape[1] % 2 != 1:
        raise ValueError("datax must be (B, 2n+1)")
    B, n, dev = datax.shape[0], (datax.shape[1] - 1) // 2, datax.device
    net, fp, keep = _net_params(n, weights, K, u_lo, u_hi, goal, B, lam, penalty, hidden, layers, u_ref)
    h = torch.empty(B, dtype=torch.float32, device=dev)
    J = torch.empty((B, n), dtype=torch.float32, device=dev)
    u = torch.empty((B, n), dtype=torch.float32, device=dev)
    r = torch.empty(B, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.cbfrrt_cbf_filter(n, C.byref(net), _ptr(datax), B, C.byref(fp), _ptr(h), _ptr(J), _ptr(u), _ptr(r),
                                    _stream()), "cbfrrt_cbf_filter")
    del keep
    return h, J, u, r


def cbf_filter_u_into(datax: Tensor, goal: Tensor, weights: Tensor, K: Tensor, u_lo: Tensor, u_hi: Tensor, lam: float,
                      penalty: float, hidden: int, layers: int, u_out: Tensor) -> Tensor:
    """the barrier-network pass alone with only the control written (what the second pass of a valid + u sweep does):
    datax f32[B,2n+1] -> u_out f32[B,n] (caller-owned).  bench.py times it for the per-pass roofline."""
    datax = _f32c(datax)
    B, n, dev = datax.shape[0], (datax.shape[1] - 1) // 2, datax.device
    if tuple(u_out.shape) != (B, n) or u_out.dtype != torch.float32 or not u_out.is_contiguous() or u_out.device != dev:
        raise ValueError("cbf_filter_u_into: u_out must be a contiguous (B, n) float32 tensor on datax's device")
    net, fp, keep = _net_params(n, weights, K, u_lo, u_hi, goal, B, lam, penalty, hidden, layers, None)
    with torch.cuda.device(dev):
        check(lib.cbfrrt_cbf_filter(n, C.byref(net), _ptr(datax), B, C.byref(fp), None, None, _ptr(u_out), None, _stream()),
              "cbfrrt_cbf_filter")
    del keep
    return u_out


@torch.library.custom_op("cbfrrt::qp", mutates_args=(), device_types="cuda")
def qp(a: Tensor, c: Tensor, u_ref: Tensor, u_lo: Tensor, u_hi: Tensor, penalty: float) -> Tuple[Tensor, Tensor]:
    """a = Lg_V f32[B,n], c = Lf_V + lam*V f32[B], u_ref f32[B,n] -> u f32[B,n], r f32[B]"""
    a, c, u_ref, u_lo, u_hi = _f32c(a), _f32c(c).reshape(-1), _f32c(u_ref), _f32c(u_lo), _f32c(u_hi)
    B, n = a.shape
    if u_ref.shape != a.shape or c.numel() != B or u_lo.numel() != n or u_hi.numel() != n:
        raise ValueError("qp: inconsistent shapes")
    u = torch.empty((B, n), dtype=torch.float32, device=a.device)
    r = torch.empty(B, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib.cbfrrt_qp(n, _ptr(a), _ptr(c), _ptr(u_ref), B, _ptr(u_lo), _ptr(u_hi), penalty, _ptr(u), _ptr(r),
                            _stream()), "cbfrrt_qp")
    return u, r


@torch.library.custom_op("cbfrrt::qp_backward", mutates_args=(), device_types="cuda")
def qp_backward(a: Tensor, c: Tensor, u_ref: Tensor, u_lo: Tensor, u_hi: Tensor, penalty: float, grad_u: Tensor,
                grad_r: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """reverse mode of ``qp``: (dL/du f32[B,n], dL/dr f32[B] or empty) -> dL/da f32[B,n], dL/dc f32[B], dL/du_ref f32[B,n]"""
    a, c, u_ref, u_lo, u_hi = _f32c(a), _f32c(c).reshape(-1), _f32c(u_ref), _f32c(u_lo), _f32c(u_hi)
    B, n = a.shape
    grad_u = _f32c(grad_u)
    grad_r = _f32c(grad_r).reshape(-1)
    if u_ref.shape != a.shape or grad_u.shape != a.shape or c.numel() != B or grad_r.numel() not in (0, B) \
            or u_lo.numel() != n or u_hi.numel() != n:
        raise ValueError("qp_backward: inconsistent shapes")
    ga = torch.empty((B, n), dtype=torch.float32, device=a.device)
    gc = torch.empty(B, dtype=torch.float32, device=a.device)
    gt = torch.empty((B, n), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib.cbfrrt_qp_backward(n, _ptr(a), _ptr(c), _ptr(u_ref), B, _ptr(u_lo), _ptr(u_hi), penalty, _ptr(grad_u),
                                     _ptr(grad_r), _ptr(ga), _ptr(gc), _ptr(gt), _stream()), "cbfrrt_qp_backward")
    return ga, gc, gt


@torch.library.custom_op("cbfrrt::qp_multi", mutates_args=(), device_types="cuda")
def qp_multi(a: Tensor, c: Tensor, u_ref: Tensor, u_lo: Tensor, u_hi: Tensor, penalty: float, max_sweeps: int,
             tol: float) -> Tuple[Tensor, Tensor, Tensor]:
    """a f32[B,S,n], c f32[B,S], u_ref f32[B,n] -> u f32[B,n], r f32[B,S], sweeps i32[B]   (S <= 4 scenarios)"""
    a, c, u_ref, u_lo, u_hi = _f32c(a), _f32c(c), _f32c(u_ref), _f32c(u_lo), _f32c(u_hi)
    if a.dim() != 3:
        raise ValueError("qp_multi: a must be [B, S, n]")
    B, S, n = a.shape
    if u_ref.shape != (B, n) or c.shape != (B, S) or u_lo.numel() != n or u_hi.numel() != n:
        raise ValueError("qp_multi: inconsistent shapes")
    u = torch.empty((B, n), dtype=torch.float32, device=a.device)
    r = torch.empty((B, S), dtype=torch.float32, device=a.device)
    sweeps = torch.empty(B, dtype=torch.int32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib.cbfrrt_qp_multi(n, S, _ptr(a), _ptr(c), _ptr(u_ref), B, _ptr(u_lo), _ptr(u_hi), penalty, max_sweeps,
                                  tol, _ptr(u), _ptr(r), _ptr(sweeps), _stream()), "cbfrrt_qp_multi")
    return u, r, sweeps


@torch.library.custom_op("cbfrrt::qp_multi_backward", mutates_args=(), device_types="cuda")
def qp_multi_backward(a: Tensor, c: Tensor, u_ref: Tensor, u_lo: Tensor, u_hi: Tensor, penalty: float, max_sweeps: int,
                      tol: float, grad_u: Tensor, grad_r: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """reverse mode of ``qp_multi``: (dL/du f32[B,n], dL/dr f32[B,S] or empty) -> dL/da [B,S,n], dL/dc [B,S], dL/du_ref [B,n]"""
    a, c, u_ref, u_lo, u_hi, grad_u, grad_r = (_f32c(x) for x in (a, c, u_ref, u_lo, u_hi, grad_u, grad_r))
    if a.dim() != 3:
        raise ValueError("qp_multi_backward: a must be [B, S, n]")
    B, S, n = a.shape
    if u_ref.shape != (B, n) or grad_u.shape != (B, n) or c.shape != (B, S) or grad_r.numel() not in (0, B * S) \
            or u_lo.numel() != n or u_hi.numel() != n:
        raise ValueError("qp_multi_backward: inconsistent shapes")
    ga = torch.empty((B, S, n), dtype=torch.float32, device=a.device)
    gc = torch.empty((B, S), dtype=torch.float32, device=a.device)
    gt = torch.empty((B, n), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib.cbfrrt_qp_multi_backward(n, S, _ptr(a), _ptr(c), _ptr(u_ref), B, _ptr(u_lo), _ptr(u_hi), penalty,
                                           max_sweeps, tol, _ptr(grad_u), _ptr(grad_r), _ptr(ga), _ptr(gc), _ptr(gt),
                                           _stream()), "cbfrrt_qp_multi_backward")
    return ga, gc, gt


@torch.library.custom_op("cbfrrt::safety_filter", mutates_args=(), device_types="cuda")
def safety_filter(q: Tensor, goal: Tensor, boxes: Tensor, spheres: Tensor, floor: int, grid: Tensor, weights: Tensor, K: Tensor,
                  u_lo: Tensor, u_hi: Tensor, lam: float, penalty: float, thr_soft: float, thr_hard: float,
                  robot_id: int, hidden: int, layers: int
                  ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """fused q -> safe, unsafe, datax, h, J, u, r"""
    n = ROBOT_N[robot_id]
    q, qs = _rows(q, n)
    B, dev = q.shape[0], q.device
    net, fp, keep = _net_params(n, weights, K, u_lo, u_hi, goal, B, lam, penalty, hidden, layers)
    sc = _scene(boxes, spheres, floor, grid)
    safe = torch.empty(B, dtype=torch.uint8, device=dev)
    unsafe = torch.empty(B, dtype=torch.uint8, device=dev)
    datax = torch.empty((B, 2 * n + 1), dtype=torch.float32, device=dev)
    h = torch.empty(B, dtype=torch.float32, device=dev)
    J = torch.empty((B, n), dtype=torch.float32, device=dev)
    u = torch.empty((B, n), dtype=torch.float32, device=dev)
    r = torch.empty(B, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.cbfrrt_safety_filter(robot_id, C.byref(sc), C.byref(net), _ptr(q), qs, B, C.byref(fp), thr_soft,
                                       thr_hard, _ptr(safe), _ptr(unsafe), _ptr(datax), _ptr(h), _ptr(J), _ptr(u),
                                       _ptr(r), _stream()), "cbfrrt_safety_filter")
    del keep
    return safe, unsafe, datax, h, J, u, r


@torch.library.custom_op("cbfrrt::valid_control", mutates_args=(), device_types="cuda")
def valid_control(q: Tensor, goal: Tensor, boxes: Tensor, spheres: Tensor, floor: int, grid: Tensor, weights: Tensor, K: Tensor,
                  u_lo: Tensor, u_hi: Tensor, lam: float, penalty: float, thr_soft: float, thr_hard: float,
                  robot_id: int, hidden: int, layers: int) -> Tuple[Tensor, Tensor]:
    """fused q -> (valid u8[B], u f32[B,n]): the sweep outputs that are all-gathered across GPUs"""
    n = ROBOT_N[robot_id]
    q, qs = _rows(q, n)
    B, dev = q.shape[0], q.device
    net, fp, keep = _net_params(n, weights, K, u_lo, u_hi, goal, B, lam, penalty, hidden, layers)
    sc = _scene(boxes, spheres, floor, grid)
    safe = torch.empty(B, dtype=torch.uint8, device=dev)
    u = torch.empty((B, n), dtype=torch.float32, device=dev)
    # network shapes other than 48 x 2 run as two launches and need the observation rows in between
    datax = None if (hidden, layers) == (48, 2) else torch.empty((B, 2 * n + 1), dtype=torch.float32, device=dev)
    _ensure_workspace(dev, B, n, hidden, layers)
    with torch.cuda.device(dev):
        check(lib.cbfrrt_safety_filter(robot_id, C.byref(sc), C.byref(net), _ptr(q), qs, B, C.byref(fp), thr_soft,
                                       thr_hard, _ptr(safe), None, _ptr(datax), None, None, _ptr(u), None, _stream()),
              "cbfrrt_safety_filter")
    del keep, datax
    return safe, u


def valid_control_into(q: Tensor, goal: Tensor, boxes: Tensor, spheres: Tensor, floor: int, grid: Tensor, weights: Tensor,
                       K: Tensor, u_lo: Tensor, u_hi: Tensor, lam: float, penalty: float, thr_soft: float, thr_hard: float,
                       robot_id: int, hidden: int, layers: int, valid_out: Tensor, u_out: Tensor) -> Tuple[Tensor, Tensor]:
    """``valid_control`` into caller-owned outputs (valid u8[B], u f32[B,n]): sweeps of tens of millions of states
    reuse their 2 GB result buffers instead of allocating per call"""
    n = ROBOT_N[robot_id]
    q, qs = _rows(q, n)
    B, dev = q.shape[0], q.device
    if (valid_out.dtype != torch.uint8 or valid_out.numel() != B or u_out.dtype != torch.float32 or tuple(u_out.shape) != (B, n)
            or not valid_out.is_contiguous() or not u_out.is_contiguous() or valid_out.device != dev or u_out.device != dev):
        raise ValueError("valid_control_into: outputs must be contiguous (B,) uint8 and (B, n) float32 on q's device")
    net, fp, keep = _net_params(n, weights, K, u_lo, u_hi, goal, B, lam, penalty, hidden, layers)
    sc = _scene(boxes, spheres, floor, grid)
    _ensure_workspace(dev, B, n, hidden, layers)
    with torch.cuda.device(dev):
        check(lib.cbfrrt_safety_filter(robot_id, C.byref(sc), C.byref(net), _ptr(q), qs, B, C.byref(fp), thr_soft,
                                       thr_hard, _ptr(valid_out), None, None, None, None, _ptr(u_out), None, _stream()),
              "cbfrrt_safety_filter")
    del keep
    return valid_out, u_out


def broadphase_stats(q: Tensor, boxes: Tensor, spheres: Tensor, floor: int, grid: Tensor, thr_soft: float, thr_hard: float,
                     robot_id: int, want_grad: bool = True) -> dict:
    """Measured work of the broad phase on these states (cbfrrt_scene.stats): per state, the link spheres swept exactly,
    the exact sphere-obstacle pair evaluations and the field look-ups.  want_grad: the observation caller (masks + datax)
    or the mask-only caller.  Used by bench.py for the per-pipe roofline."""
    n = ROBOT_N[robot_id]
    q, qs = _rows(q, n)
    B, dev = q.shape[0], q.device
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    safe = torch.empty(B, dtype=torch.uint8, device=dev)
    unsafe = torch.empty(B, dtype=torch.uint8, device=dev)
    datax = torch.empty((B, 2 * n + 1), dtype=torch.float32, device=dev) if want_grad else None
    sc = _scene(boxes, spheres, floor, grid, stats)
    with torch.cuda.device(dev):
        check(lib.cbfrrt_collide(robot_id, C.byref(sc), _ptr(q), qs, B, thr_soft, thr_hard, _ptr(safe), _ptr(unsafe),
                                 _ptr(datax) if want_grad else None, None, None, None, _stream()), "cbfrrt_collide")
    st = stats.cpu().numpy().astype(np.float64)
    states = max(st[0], 1.0)
    n_obs = (int(boxes.shape[0]) if boxes.numel() else 0) + (int(spheres.shape[0]) if spheres.numel() else 0)
    field = st[3] > 0
    return {"grid": bool(field), "states": int(st[0]), "spheres_swept_per_state": st[1] / states if field else None,
            "pairs_per_state": st[2] / states if field else None, "lookups_per_state": st[3] / states,
            "obstacles": n_obs}


def valid_control_scatter(q: Tensor, goal: Tensor, boxes: Tensor, spheres: Tensor, floor: int, grid: Tensor, weights: Tensor,
                          K: Tensor, u_lo: Tensor, u_hi: Tensor, lam: float, penalty: float, thr_soft: float,
                          thr_hard: float, robot_id: int, hidden: int, layers: int, valid_ptrs, u_ptrs,
                          row_offset: int, multicast: bool = False) -> None:
    """Fused sweep + all-gather (include/cbfrrt.h cbfrrt_safety_filter_scatter): the (valid, u) rows of this rank
    are stored at global row ``row_offset + b`` of every peer buffer in ``valid_ptrs`` / ``u_ptrs`` (raw device
    pointers of NVLink peer mappings, e.g. torch symmetric memory ``buffer_ptrs``).  multicast=True: the single pointer
    pair is the buffers' NVSwitch multicast address and the kernel stores with multimem.st (B and row_offset must be
    multiples of 128).  No output tensors."""
    n = ROBOT_N[robot_id]
    q, qs = _rows(q, n)
    B, dev = q.shape[0], q.device
    net, fp, keep = _net_params(n, weights, K, u_lo, u_hi, goal, B, lam, penalty, hidden, layers)
    sc = _scene(boxes, spheres, floor, grid)
    po = PeerOut()
    po.n_peers, po.multicast, po.row_offset = len(valid_ptrs), int(bool(multicast)), int(row_offset)
    for i, (v, u) in enumerate(zip(valid_ptrs, u_ptrs)):
        po.valid[i], po.u[i] = int(v), int(u)
    _ensure_workspace(dev, B, n, hidden, layers)
    with torch.cuda.device(dev):
        check(lib.cbfrrt_safety_filter_scatter(robot_id, C.byref(sc), C.byref(net), _ptr(q), qs, B, C.byref(fp), thr_soft,
                                               thr_hard, C.byref(po), _stream()), "cbfrrt_safety_filter_scatter")
    del keep


def observe_into(q: Tensor, boxes: Tensor, spheres: Tensor, floor: int, grid: Tensor, thr_soft: float, thr_hard: float,
                 robot_id: int, safe_out: Tensor, datax_out: Tensor) -> None:
    """the geometry pass into caller-owned buffers: q -> safe u8[B], datax f32[B,2n+1] (cbfrrt_collide)"""
    n = ROBOT_N[robot_id]
    q, qs = _rows(q, n)
    B, dev = q.shape[0], q.device
    if safe_out.numel() < B or datax_out.numel() < B * (2 * n + 1) or safe_out.dtype != torch.uint8 or datax_out.dtype != torch.float32:
        raise ValueError("observe_into: safe_out u8[>=B], datax_out f32[>=B*(2n+1)]")
    sc = _scene(boxes, spheres, floor, grid)
    with torch.cuda.device(dev):
        check(lib.cbfrrt_collide(robot_id, C.byref(sc), _ptr(q), qs, B, thr_soft, thr_hard, _ptr(safe_out), None, _ptr(datax_out),
                                 None, None, None, _stream()), "cbfrrt_collide")


def cbf_filter_scatter(datax: Tensor, valid_in: Tensor, rows: int, goal: Tensor, weights: Tensor, K: Tensor, u_lo: Tensor,
                       u_hi: Tensor, lam: float, penalty: float, hidden: int, layers: int, n: int, valid_ptrs, u_ptrs,
                       row_offset: int, multicast: bool = False) -> None:
    """the barrier-network pass with the fused exchange (include/cbfrrt.h cbfrrt_cbf_filter_scatter): the first ``rows`` rows of
    datax / valid_in -> (valid, u) stored at global row ``row_offset + b`` of every peer buffer"""
    dev = datax.device
    net, fp, keep = _net_params(n, weights, K, u_lo, u_hi, goal, rows, lam, penalty, hidden, layers)
    po = PeerOut()
    po.n_peers, po.multicast, po.row_offset = len(valid_ptrs), int(bool(multicast)), int(row_offset)
    for i, (v, u) in enumerate(zip(valid_ptrs, u_ptrs)):
        po.valid[i], po.u[i] = int(v), int(u)
    with torch.cuda.device(dev):
        check(lib.cbfrrt_cbf_filter_scatter(n, C.byref(net), _ptr(datax), _ptr(valid_in), rows, C.byref(fp), C.byref(po), _stream()),
              "cbfrrt_cbf_filter_scatter")
    del keep


@torch.library.custom_op("cbfrrt::edge_validity", mutates_args=(), device_types="cuda")
def edge_validity(q_from: Tensor, q_to: Tensor, n_interp: int, boxes: Tensor, spheres: Tensor, floor: int,
                  grid: Tensor, thr_soft: float, thr_hard: float, robot_id: int) -> Tuple[Tensor, Tensor]:
    """-> valid u8[E], first_bad i32[E]"""
    n = ROBOT_N[robot_id]
    q_from, q_to = _f32c(q_from), _f32c(q_to)
    if q_from.shape != q_to.shape or q_from.dim() != 2 or q_from.shape[1] != n:
        raise ValueError(f"q_from / q_to must both be (E, {n})")
    E, dev = q_from.shape[0], q_from.device
    valid = torch.empty(E, dtype=torch.uint8, device=dev)
    first_bad = torch.empty(E, dtype=torch.int32, device=dev)
    sc = _scene(boxes, spheres, floor, grid)
    with torch.cuda.device(dev):
        check(lib.cbfrrt_edge_validity(robot_id, C.byref(sc), _ptr(q_from), _ptr(q_to), E, n_interp, thr_soft, thr_hard,
                                       _ptr(valid), _ptr(first_bad), _stream()), "cbfrrt_edge_validity")
    return valid, first_bad


@torch.library.custom_op("cbfrrt::edge_validity_var", mutates_args=(), device_types="cuda")
def edge_validity_var(q_from: Tensor, q_to: Tensor, n_interp: Tensor, boxes: Tensor, spheres: Tensor, floor: int,
                      grid: Tensor, thr_soft: float, thr_hard: float, robot_id: int) -> Tuple[Tensor, Tensor]:
    """edges with their own number of interpolation points n_interp i32/i64[E] (>= 0; tsa.py:277-279)
    -> valid u8[E], first_bad i32[E]"""
    n = ROBOT_N[robot_id]
    q_from, q_to = _f32c(q_from), _f32c(q_to)
    if q_from.shape != q_to.shape or q_from.dim() != 2 or q_from.shape[1] != n:
        raise ValueError(f"q_from / q_to must both be (E, {n})")
    E, dev = q_from.shape[0], q_from.device
    if n_interp.numel() != E:
        raise ValueError("n_interp must have one entry per edge")
    k_host = n_interp.detach().reshape(-1).to("cpu", torch.int64)    # sizes the launch: the total must be known on the host
    if E and int(k_host.min()) < 0:
        raise ValueError("n_interp must be >= 0")
    off_host = torch.zeros(E + 1, dtype=torch.int64)
    torch.cumsum(k_host + 1, 0, out=off_host[1:])
    off = off_host.to(dev, non_blocking=True)
    valid = torch.empty(E, dtype=torch.uint8, device=dev)
    first_bad = torch.empty(E, dtype=torch.int32, device=dev)
    sc = _scene(boxes, spheres, floor, grid)
    with torch.cuda.device(dev):
        check(lib.cbfrrt_edge_validity_var(robot_id, C.byref(sc), _ptr(q_from), _ptr(q_to), E, _ptr(off), int(off_host[-1]),
                                           thr_soft, thr_hard, _ptr(valid), _ptr(first_bad), _stream()),
              "cbfrrt_edge_validity_var")
    del off
    return valid, first_bad


@torch.library.custom_op("cbfrrt::knn", mutates_args=(), device_types="cuda")
def knn(queries: Tensor, tree: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """queries f32[M,n], tree f32[N,n] -> idx i32[M,k], dist f32[M,k]: k nearest rows, ascending (distance, row)"""
    queries, tree = _f32c(queries), _f32c(tree)
    if queries.dim() != 2 or tree.dim() != 2 or queries.shape[1] != tree.shape[1]:
        raise ValueError("knn: queries [M,n] and tree [N,n] must share n")
    M, n = queries.shape
    idx = torch.empty((M, k), dtype=torch.int32, device=queries.device)
    dist = torch.empty((M, k), dtype=torch.float32, device=queries.device)
    ws = torch.empty(int(lib.cbfrrt_knn_workspace_bytes(M, tree.shape[0], k)) // 8, dtype=torch.int64, device=queries.device)
    with torch.cuda.device(queries.device):
        check(lib.cbfrrt_knn(n, _ptr(queries), _ptr(tree), M, tree.shape[0], k, _ptr(idx), _ptr(dist), _ptr(ws), _stream()),
              "cbfrrt_knn")
    del ws
    return idx, dist


@torch.library.custom_op("cbfrrt::radius_mask", mutates_args=(), device_types="cuda")
def radius_mask(queries: Tensor, points: Tensor, radius: float) -> Tensor:
    """within u8[M,N] = ||queries[m] - points[j]|| <= radius   (bit_star.py:196-216)"""
    queries, points = _f32c(queries), _f32c(points)
    if queries.dim() != 2 or points.dim() != 2 or queries.shape[1] != points.shape[1]:
        raise ValueError("radius_mask: queries [M,n] and points [N,n] must share n")
    M, n = queries.shape
    out = torch.empty((M, points.shape[0]), dtype=torch.uint8, device=queries.device)
    with torch.cuda.device(queries.device):
        check(lib.cbfrrt_radius_mask(n, _ptr(queries), _ptr(points), M, points.shape[0], radius, _ptr(out), _stream()),
              "cbfrrt_radius_mask")
    return out


@torch.library.custom_op("cbfrrt::rollout", mutates_args=(), device_types="cuda")
def rollout(q0: Tensor, goal: Tensor, steps: int, ctrl_every: int, dt: float, stuck_lag: int, stuck_after: int,
            stuck_eps: float, want_traj: bool, keep_going: bool, boxes: Tensor, spheres: Tensor, floor: int, grid: Tensor, weights: Tensor, K: Tensor,
            u_lo: Tensor, u_hi: Tensor, lam: float, penalty: float, thr_soft: float, thr_hard: float, robot_id: int,
            hidden: int, layers: int, stuck_mode: int = 0) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """the steer loop in one launch -> q_final f32[T,n], alive u8[T], steps_done i32[T], traj f32[T,steps,n]
    (traj is (0,) unless want_traj or the stuck test is on).  stuck_mode 0: tsa.py steer, 1: batch_tsa.py steer
    (include/cbfrrt.h cbfrrt_rollout_params)"""
    n = ROBOT_N[robot_id]
    q0 = _f32c(q0)
    if q0.dim() != 2 or q0.shape[1] != n:
        raise ValueError(f"q0 must be (T, {n})")
    T, dev = q0.shape[0], q0.device
    net, fp, keep = _net_params(n, weights, K, u_lo, u_hi, goal, T, lam, penalty, hidden, layers)
    sc = _scene(boxes, spheres, floor, grid)
    rp = RolloutParams(steps, ctrl_every, dt, stuck_lag, stuck_after, stuck_eps, int(keep_going), int(stuck_mode))
    qf = torch.empty((T, n), dtype=torch.float32, device=dev)
    alive = torch.empty(T, dtype=torch.uint8, device=dev)
    done = torch.empty(T, dtype=torch.int32, device=dev)
    traj = (torch.zeros((T, steps, n), dtype=torch.float32, device=dev) if (want_traj or stuck_lag > 0)
            else torch.empty(0, dtype=torch.float32, device=dev))
    with torch.cuda.device(dev):
        check(lib.cbfrrt_rollout(robot_id, C.byref(sc), C.byref(net), _ptr(q0), T, C.byref(fp), C.byref(rp), thr_soft,
                                 thr_hard, _ptr(qf), _ptr(alive), _ptr(done), _ptr(traj) if traj.numel() else None,
                                 _stream()), "cbfrrt_rollout")
    del keep
    return qf, alive, done, traj


def rollout_segments(q0: Tensor, goal: Tensor, n_segments: int, seg_len: int, ctrl_every: int, dt: float, stuck_lag: int,
                     stuck_after: int, stuck_eps: float, boxes: Tensor, spheres: Tensor, floor: int, grid: Tensor, weights: Tensor,
                     K: Tensor, u_lo: Tensor, u_hi: Tensor, lam: float, penalty: float, thr_soft: float, thr_hard: float,
                     robot_id: int, hidden: int, layers: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """n_segments consecutive batched steer segments of ONE batch (T <= 128 rows) in one launch
    (include/cbfrrt.h cbfrrt_rollout_segments; batch_tsa.py:154-170 over :191-231)
    -> seg_q f32[T,n_seg,n], seg_alive u8[T,n_seg], seg_done i32[T,n_seg], traj f32[T,n_seg*seg_len,n]"""
    n = ROBOT_N[robot_id]
    q0 = _f32c(q0)
    if q0.dim() != 2 or q0.shape[1] != n or not q0.is_cuda:
        raise ValueError(f"q0 must be a CUDA tensor (T, {n})")
    T, dev = q0.shape[0], q0.device
    net, fp, keep = _net_params(n, weights, K, u_lo, u_hi, goal, T, lam, penalty, hidden, layers)
    sc = _scene(boxes, spheres, floor, grid)
    rp = RolloutParams(seg_len, ctrl_every, dt, stuck_lag, stuck_after, stuck_eps, 1, 1)
    seg_q = torch.empty((T, n_segments, n), dtype=torch.float32, device=dev)
    seg_alive = torch.empty((T, n_segments), dtype=torch.uint8, device=dev)
    seg_done = torch.empty((T, n_segments), dtype=torch.int32, device=dev)
    traj = torch.empty((T, n_segments * seg_len, n), dtype=torch.float32, device=dev)   # rows beyond seg_done are never read
    with torch.cuda.device(dev):
        check(lib.cbfrrt_rollout_segments(robot_id, C.byref(sc), C.byref(net), _ptr(q0), T, C.byref(fp), C.byref(rp), n_segments,
                                          thr_soft, thr_hard, _ptr(seg_q), _ptr(seg_alive), _ptr(seg_done), _ptr(traj),
                                          _stream()), "cbfrrt_rollout_segments")
    del keep
    return seg_q, seg_alive, seg_done, traj


# ----------------------------------------------------------------------------- direct entry points
registered = {}


def _direct(op, cpu_ok=()):
    impl = op._init_fn

    @functools.wraps(impl)
    def call(*args, **kwargs):
        for i, a in enumerate(args):
            if isinstance(a, Tensor) and not a.is_cuda and i not in cpu_ok:
                raise NotImplementedError(f"cbfrrt::{impl.__name__} needs CUDA tensors (argument {i} is on {a.device}); "
                                          "there is no CPU path")
        return impl(*args, **kwargs)

    call.op = op
    return call


for _name in ("fk", "collide", "observe", "masks", "cbf_filter", "qp", "qp_backward", "qp_multi", "qp_multi_backward", "safety_filter", "valid_control",
              "edge_validity", "edge_validity_var", "knn", "radius_mask", "rollout"):
    registered[_name] = globals()[_name]
    globals()[_name] = _direct(registered[_name], cpu_ok=(2,) if _name == "edge_validity_var" else ())
del _name


class _DifferentiableQP(torch.autograd.Function):
    """``qp`` with the closed-form reverse mode ``qp_backward`` (no gradient for the limits and the penalty)."""

    @staticmethod
    def forward(ctx, a, c, u_ref, u_lo, u_hi, penalty):
        ctx.save_for_backward(a, c, u_ref, u_lo, u_hi)
        ctx.penalty = float(penalty)
        return qp(a, c, u_ref, u_lo, u_hi, float(penalty))

    @staticmethod
    def backward(ctx, grad_u, grad_r):
        a, c, u_ref, u_lo, u_hi = ctx.saved_tensors
        empty = torch.empty(0, dtype=torch.float32, device=a.device)
        gu = torch.zeros_like(a, dtype=torch.float32) if grad_u is None else grad_u
        ga, gc, gt = qp_backward(a, c, u_ref, u_lo, u_hi, ctx.penalty, gu, empty if grad_r is None else grad_r)
        return ga.to(a.dtype).reshape(a.shape), gc.to(c.dtype).reshape(c.shape), gt.to(u_ref.dtype), None, None, None


class _DifferentiableMultiQP(torch.autograd.Function):
    """``qp_multi`` with the reverse mode ``qp_multi_backward``"""

    @staticmethod
    def forward(ctx, a, c, u_ref, u_lo, u_hi, penalty, max_sweeps, tol):
        ctx.save_for_backward(a, c, u_ref, u_lo, u_hi)
        ctx.args = (float(penalty), int(max_sweeps), float(tol))
        u, r, _ = qp_multi(a, c, u_ref, u_lo, u_hi, *ctx.args)
        return u, r

    @staticmethod
    def backward(ctx, grad_u, grad_r):
        a, c, u_ref, u_lo, u_hi = ctx.saved_tensors
        empty = torch.empty(0, dtype=torch.float32, device=a.device)
        gu = torch.zeros_like(u_ref, dtype=torch.float32) if grad_u is None else grad_u
        ga, gc, gt = qp_multi_backward(a, c, u_ref, u_lo, u_hi, *ctx.args, gu, empty if grad_r is None else grad_r)
        return ga.to(a.dtype), gc.to(c.dtype), gt.to(u_ref.dtype), None, None, None, None, None


def qp_multi_diff(a: Tensor, c: Tensor, u_ref: Tensor, u_lo: Tensor, u_hi: Tensor, penalty: float, max_sweeps: int = 100,
                  tol: float = 1e-9) -> Tuple[Tensor, Tensor]:
    """differentiable multi-scenario QP: (u [B,n], r [B,S]) with gradients w.r.t. a [B,S,n], c [B,S], u_ref [B,n]"""
    return _DifferentiableMultiQP.apply(a, c, u_ref, u_lo, u_hi, penalty, max_sweeps, tol)


def qp_diff(a: Tensor, c: Tensor, u_ref: Tensor, u_lo: Tensor, u_hi: Tensor, penalty: float) -> Tuple[Tensor, Tensor]:
    """differentiable single-constraint QP: (u [B,n], r [B]) with gradients w.r.t. a = Lg_V, c = Lf_V + lam V, u_ref
    -- what the reference gets from its CvxpyLayer (clf_controller.py:354-404, requires_grad=True)."""
    return _DifferentiableQP.apply(a, c, u_ref, u_lo, u_hi, penalty)


def sample_states(B: int, seed: int, start_row: int, robot_id: int, device="cuda") -> Tensor:
    """counter-based uniform sampler over the joint limits; row g depends only on (seed, g)"""
    n = ROBOT_N[robot_id]
    q = torch.empty((B, n), dtype=torch.float32, device=device)
    with torch.cuda.device(q.device):
        check(lib.cbfrrt_sample_states(robot_id, seed & 0xFFFFFFFF, start_row, B, _ptr(q), _stream()),
              "cbfrrt_sample_states")
    return q


def safety_filter_host(robot: str, q: np.ndarray, boxes6: np.ndarray, spheres: np.ndarray, floor: bool,
                       weights: np.ndarray, hidden: int, layers: int, goal: np.ndarray, K: np.ndarray, u_lo, u_hi,
                       lam: float, penalty: float, thr_soft: float, thr_hard: float, out: dict, device: int = 0):
    """Host-buffer entry point (numpy in / numpy out; include/cbfrrt.h cbfrrt_safety_filter_host).
    ``out`` maps any of safe, unsafe, datax, h, J, u, r to preallocated (ideally pinned) numpy arrays."""
    n = ROBOT_N[ROBOT_ID[robot]]
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    q, goal, K, u_lo, u_hi, weights = f(q), f(goal).reshape(-1, n), f(K), f(u_lo), f(u_hi), f(weights)
    boxes6 = f(boxes6).reshape(-1, 6)
    spheres = f(spheres).reshape(-1, 4) if spheres is not None else np.zeros((0, 4), np.float32)
    hs = _lib.HostScene(boxes6.shape[0], spheres.shape[0], int(bool(floor)), 0, boxes6.ctypes.data, spheres.ctypes.data)
    p = lambda k: C.c_void_p(out[k].ctypes.data) if out.get(k) is not None else None
    check(lib.cbfrrt_safety_filter_host(device, ROBOT_ID[robot], C.byref(hs), hidden, layers, weights.ctypes.data,
                                        weights.size, q.ctypes.data, q.shape[0], goal.ctypes.data, goal.shape[0],
                                        K.ctypes.data, u_lo.ctypes.data, u_hi.ctypes.data, lam, penalty, thr_soft,
                                        thr_hard, p("safe"), p("unsafe"), p("datax"), p("h"), p("J"), p("u"), p("r")),
          "cbfrrt_safety_filter_host")
    return out


def rollout_segments_host(robot: str, q0: np.ndarray, goal: np.ndarray, n_segments: int, seg_len: int, ctrl_every: int,
                          dt: float, stuck_lag: int, stuck_after: int, stuck_eps: float, boxes6: np.ndarray, spheres,
                          floor: bool, weights: np.ndarray, hidden: int, layers: int, K: np.ndarray, u_lo, u_hi, lam: float,
                          penalty: float, thr_soft: float, thr_hard: float, device: int = 0):
    """numpy in / numpy out form of rollout_segments (include/cbfrrt.h cbfrrt_rollout_segments_host): one pinned H2D copy,
    the launch, one D2H copy.  -> seg_q f32[T,n_seg,n], seg_alive bool[T,n_seg], seg_done i32[T,n_seg],
    traj f32[T,n_seg*seg_len,n].  ``weights``: the packed buffer (workloads.pack_weights_np)."""
    n = ROBOT_N[ROBOT_ID[robot]]
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    q0, goal, K, u_lo, u_hi, weights = f(q0).reshape(-1, n), f(goal).reshape(-1, n), f(K), f(u_lo), f(u_hi), f(weights)
    T = q0.shape[0]
    if goal.shape[0] != T:
        raise ValueError(f"goal must hold one row per trajectory ({T}), got {goal.shape[0]}")
    boxes6 = f(boxes6).reshape(-1, 6)
    spheres = f(spheres).reshape(-1, 4) if spheres is not None else np.zeros((0, 4), np.float32)
    hs = _lib.HostScene(boxes6.shape[0], spheres.shape[0], int(bool(floor)), 0, boxes6.ctypes.data, spheres.ctypes.data)
    rp = RolloutParams(seg_len, ctrl_every, dt, stuck_lag, stuck_after, stuck_eps, 1, 1)
    seg_q = np.empty((T, n_segments, n), np.float32)
    seg_alive = np.empty((T, n_segments), np.uint8)
    seg_done = np.empty((T, n_segments), np.int32)
    traj = np.empty((T, n_segments * seg_len, n), np.float32)
    check(lib.cbfrrt_rollout_segments_host(device, ROBOT_ID[robot], C.byref(hs), hidden, layers, weights.ctypes.data, weights.size,
                                           q0.ctypes.data, T, goal.ctypes.data, K.ctypes.data, u_lo.ctypes.data, u_hi.ctypes.data,
                                           lam, penalty, C.byref(rp), n_segments, thr_soft, thr_hard, seg_q.ctypes.data,
                                           seg_alive.ctypes.data, seg_done.ctypes.data, traj.ctypes.data),
          "cbfrrt_rollout_segments_host")
    return seg_q, seg_alive.astype(bool), seg_done, traj


def edge_validity_host(robot: str, q_from: np.ndarray, q_to: np.ndarray, n_interp: np.ndarray, boxes6: np.ndarray, spheres,
                       floor: bool, thr_soft: float, thr_hard: float, device: int = 0):
    """numpy in / numpy out form of edge_validity_var (include/cbfrrt.h cbfrrt_edge_validity_host)
    -> valid bool[E], first_bad i32[E]"""
    n = ROBOT_N[ROBOT_ID[robot]]
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    q_from, q_to = f(q_from).reshape(-1, n), f(q_to).reshape(-1, n)
    K = np.ascontiguousarray(n_interp, dtype=np.int32).reshape(-1)
    E = q_from.shape[0]
    if q_to.shape[0] != E or K.shape[0] != E:
        raise ValueError("q_from, q_to and n_interp must describe the same number of edges")
    boxes6 = f(boxes6).reshape(-1, 6)
    spheres = f(spheres).reshape(-1, 4) if spheres is not None else np.zeros((0, 4), np.float32)
    hs = _lib.HostScene(boxes6.shape[0], spheres.shape[0], int(bool(floor)), 0, boxes6.ctypes.data, spheres.ctypes.data)
    valid, first_bad = np.empty(E, np.uint8), np.empty(E, np.int32)
    check(lib.cbfrrt_edge_validity_host(device, ROBOT_ID[robot], C.byref(hs), q_from.ctypes.data, q_to.ctypes.data, E, K.ctypes.data,
                                        thr_soft, thr_hard, valid.ctypes.data, first_bad.ctypes.data), "cbfrrt_edge_validity_host")
    return valid.astype(bool), first_bad


def knn_host(queries: np.ndarray, tree: np.ndarray, k: int, device: int = 0):
    """numpy in / numpy out form of knn (include/cbfrrt.h cbfrrt_knn_host) -> idx i32[M,k], dist f32[M,k]"""
    queries, tree = np.ascontiguousarray(queries, np.float32), np.ascontiguousarray(tree, np.float32)
    if queries.ndim != 2 or tree.ndim != 2 or queries.shape[1] != tree.shape[1]:
        raise ValueError("knn_host: queries [M,n] and tree [N,n] must share n")
    M, n = queries.shape
    idx, dist = np.empty((M, k), np.int32), np.empty((M, k), np.float32)
    check(lib.cbfrrt_knn_host(device, n, queries.ctypes.data, tree.ctypes.data, M, tree.shape[0], k, idx.ctypes.data,
                              dist.ctypes.data), "cbfrrt_knn_host")
    return idx, dist


# ----------------------------------------------------------------------------- LiDAR / PointNet barrier
def _rna_tf32(w: np.ndarray) -> np.ndarray:
    """cvt.rna.tf32.f32: round to nearest (ties away from zero) onto 10 mantissa bits"""
    u = np.ascontiguousarray(w, np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def _tf32_tiles(w: np.ndarray) -> np.ndarray:
    """W [n_out, K] (torch layout, K % 4 == 0) -> [hi | lo] tiles in the tensor core's no-swizzle K-major layout
    W4[k/4][n] (core matrix = 8 rows x 16 bytes): hi = rna_tf32(W), lo = W - hi (exact in float32)"""
    w = np.ascontiguousarray(w, np.float32)
    hi = _rna_tf32(w)
    lo = (w - hi).astype(np.float32)
    t = lambda m: m.reshape(m.shape[0], m.shape[1] // 4, 4).transpose(1, 0, 2).reshape(-1)
    return np.concatenate([t(hi), t(lo)])


class PackedLidarNet:
    """device copy of the LiDAR barrier's networks in the layout of include/cbfrrt.h (cbfrrt_lidar_net)"""

    def __init__(self, struct, tensors, meta):
        self.struct, self._tensors, self.meta = struct, tensors, meta


def pack_lidar_net(pc_head: dict, encoder: dict, h_nn: dict, n: int, n_sensors: int, point_dims: int = 6,
                   device="cuda") -> PackedLidarNet:
    """state dicts of the reference's modules -> PackedLidarNet.
    pc_head: PointNetfeat (pointnet.py:20-48; keys backbone_nn.fc11 / fc2 / fc3, backbone_fc.fc2 / fc3), encoder:
    PointNetVanillaEncoder (pointnet.py:78-90; encoder_nn.fc1 / fc2 / fc4), h_nn: input_linear / layer_i_linear /
    output_linear (neural_lidar_cbf_controller.py:66-95)."""
    g = lambda d, k: np.ascontiguousarray(d[k].detach().cpu().numpy() if isinstance(d[k], torch.Tensor) else d[k], np.float32)
    w1 = g(pc_head, "backbone_nn.fc11.weight")
    if w1.shape != (64, point_dims + n_sensors) or point_dims + n_sensors > 16:
        raise ValueError(f"backbone_nn.fc11 must be (64, {point_dims + n_sensors}), got {w1.shape}")
    w1p = np.zeros((64, 16), np.float32)
    w1p[:, :w1.shape[1]] = w1
    w2, w3 = g(pc_head, "backbone_nn.fc2.weight"), g(pc_head, "backbone_nn.fc3.weight")
    if w2.shape != (128, 64) or w3.shape != (512, 128):
        raise ValueError("backbone_nn must be 64 -> 128 -> 512 (pointnet.py:24-26)")
    layers = sum(1 for k in h_nn if k.endswith("_linear.weight") and k.startswith("layer_"))
    hidden = g(h_nn, "input_linear.weight").shape[0]
    arrs = {
        "w1t": _tf32_tiles(w1p), "w2t": _tf32_tiles(w2),
        "w3t": np.concatenate([_tf32_tiles(w3[c * 64:(c + 1) * 64]) for c in range(8)]),
        "bias": np.concatenate([g(pc_head, "backbone_nn.fc11.bias"), g(pc_head, "backbone_nn.fc2.bias"), g(pc_head, "backbone_nn.fc3.bias")]),
        "fc2_w": g(pc_head, "backbone_fc.fc2.weight"), "fc2_b": g(pc_head, "backbone_fc.fc2.bias"),
        "fc3_w": g(pc_head, "backbone_fc.fc3.weight"), "fc3_b": g(pc_head, "backbone_fc.fc3.bias"),
        "e1_w": g(encoder, "encoder_nn.fc1.weight"), "e1_b": g(encoder, "encoder_nn.fc1.bias"),
        "e2_w": g(encoder, "encoder_nn.fc2.weight"), "e2_b": g(encoder, "encoder_nn.fc2.bias"),
        "e4_w": g(encoder, "encoder_nn.fc4.weight"), "e4_b": g(encoder, "encoder_nn.fc4.bias"),
        "h_in_w": g(h_nn, "input_linear.weight"), "h_in_b": g(h_nn, "input_linear.bias"),
        "h_hid_w": (np.concatenate([g(h_nn, f"layer_{i}_linear.weight").ravel() for i in range(layers)]) if layers else np.zeros(4, np.float32)),
        "h_hid_b": (np.concatenate([g(h_nn, f"layer_{i}_linear.bias").ravel() for i in range(layers)]) if layers else np.zeros(4, np.float32)),
        "h_out_w": g(h_nn, "output_linear.weight"), "h_out_b": g(h_nn, "output_linear.bias"),
    }
    pf, fd = arrs["fc3_w"].shape[0], arrs["e4_w"].shape[0]
    if arrs["e1_w"].shape[1] != n_sensors * pf or arrs["h_in_w"].shape[1] != n + fd:
        raise ValueError("encoder / h_nn input widths do not match n_sensors * per_feature_dim / n + feature_dim")
    tensors = {k: torch.from_numpy(np.ascontiguousarray(v.ravel())).to(device) for k, v in arrs.items()}
    st = LidarNet(n, n_sensors, point_dims, pf, fd, hidden, layers, 0, *[tensors[k].data_ptr() for k in (
        "w1t", "w2t", "w3t", "bias", "fc2_w", "fc2_b", "fc3_w", "fc3_b", "e1_w", "e1_b", "e2_w", "e2_b", "e4_w", "e4_b",
        "h_in_w", "h_in_b", "h_hid_w", "h_hid_b", "h_out_w", "h_out_b")])
    return PackedLidarNet(st, tensors, dict(n=n, n_sensors=n_sensors, point_dims=point_dims, per_feature_dim=pf, feature_dim=fd,
                                            hidden=hidden, layers=layers))


def lidar_cbf(net: PackedLidarNet, q: Tensor, cloud: Tensor, frames: Tensor, sampled_index: Tensor, J_p: Tensor = None,
              J_R: Tensor = None, delta: float = 0.0):
    """LiDAR barrier value (and, with the frame Jacobians, its one-sided numerical Jacobian) for B states in one call
    (include/cbfrrt.h cbfrrt_lidar_cbf).  q [B,n], cloud [B,P,pd], frames [B,S,12], sampled_index i32 [S,R]
    -> h [B]  or  (h [B], J [B,n]) when J_p [B,S,3,n], J_R [B,S,9,n] and delta are given."""
    m = net.meta
    q, cloud, frames = _f32c(q), _f32c(cloud), _f32c(frames)
    for t in (q, cloud, frames, sampled_index):
        if not t.is_cuda:
            raise NotImplementedError("cbfrrt::lidar_cbf needs CUDA tensors; there is no CPU path")
    B, dev = q.shape[0], q.device
    sampled_index = sampled_index.to(torch.int32).contiguous()
    if (q.shape != (B, m["n"]) or cloud.dim() != 3 or cloud.shape[0] != B or cloud.shape[2] != m["point_dims"]
            or tuple(frames.shape) != (B, m["n_sensors"], 12) or sampled_index.dim() != 2 or sampled_index.shape[0] != m["n_sensors"]):
        raise ValueError("lidar_cbf: inconsistent shapes")
    want_j = J_p is not None
    if want_j:
        J_p, J_R = _f32c(J_p).reshape(B, m["n_sensors"], 3, m["n"]), _f32c(J_R).reshape(B, m["n_sensors"], 9, m["n"])
    ws = torch.empty(int(lib.cbfrrt_lidar_workspace_bytes(C.byref(net.struct), B, int(want_j))), dtype=torch.uint8, device=dev)
    h = torch.empty(B, dtype=torch.float32, device=dev)
    J = torch.empty((B, m["n"]), dtype=torch.float32, device=dev) if want_j else None
    with torch.cuda.device(dev):
        check(lib.cbfrrt_lidar_cbf(C.byref(net.struct), _ptr(q), _ptr(cloud), cloud.shape[1], _ptr(frames),
                                   _ptr(J_p) if want_j else None, _ptr(J_R) if want_j else None, _ptr(sampled_index),
                                   int(sampled_index.shape[1]), B, float(delta), _ptr(ws), _ptr(h), _ptr(J) if want_j else None,
                                   _stream()), "cbfrrt_lidar_cbf")
    del ws
    return (h, J) if want_j else h


MLP_PATHS = {"auto": 0, "fma": 1, "tc": 2, "warp": 3}


def set_mlp_path(path: str) -> None:
    """'auto' (one state per warp up to 4096 rows, FP32-FMA kernels up to 16 384 rows, tcgen05 kernels above), or 'fma' /
    'tc' / 'warp' for every size (include/cbfrrt.h cbfrrt_set_mlp_path).  Process-wide."""
    check(lib.cbfrrt_set_mlp_path(MLP_PATHS[path]), "cbfrrt_set_mlp_path")


def get_mlp_path() -> str:
    return {v: k for k, v in MLP_PATHS.items()}[int(lib.cbfrrt_get_mlp_path())]


def launch_count() -> int:
    return int(lib.cbfrrt_launch_count())

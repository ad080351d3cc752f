"""bench.py prints ONE JSON line with the keys the driver's contract names: the reference arm (CPU port of the path)
is checked on the CPU suite, the GPU arm on the GPU suite (reduced workload so that it takes seconds)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().split("\n") if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--scale", "0.01")
    assert BASE <= set(r) and r["impl"] == "reference" and r["unit"] == "states/s" and r["value"] > 0
    assert r["cpu_baseline"]["kind"] == "port" and r["cpu_baseline"]["cores"] >= 1 and r["cpu_baseline"]["value"] == r["value"]
    assert r["e2e"] == {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["config"]["workload"].startswith("cfg4") and r["vs_baseline"] is None and r["scaling"] == "strong"
    assert "obstacles" in r["config"] and r["config"]["boxes"] == 256


def test_reference_arm_does_not_load_the_cuda_library():
    """the reference arm computes through the oracle only; it must not even map libcbfrrt.so"""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--scale', '0.001'];"
            "runpy.run_path(%r, run_name='__main__');"
            "maps = open('/proc/self/maps').read(); assert 'libcbfrrt_oracle' in maps; assert 'libcbfrrt.so' not in maps, 'CUDA library mapped'"
            % os.path.join(ROOT, "bench.py"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]


def test_reference_arm_under_torchrun_prints_once():
    """N > 1: rank 0 alone runs the reference arm and prints it, the other ranks exit 0 without output"""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
                          os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                          "--scale", "0.01"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().split("\n") if l.startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["impl"] == "reference" and json.loads(lines[0])["n_gpus"] == 2


@pytest.mark.gpu
def test_gpu_arm_line():
    r = _run("--steps", "3", "--warmup", "3", "--workload", "cfg1", "--scale", "0.125")
    assert BASE | {"clocks", "roofline"} <= set(r) and "impl" not in r
    assert r["n_gpus"] == 1 and r["steps"] == 3 and r["warmup"] >= 3 and r["higher_is_better"] and r["vs_baseline"] is None
    # structure only: three timed steps on a shared box are too noisy for performance thresholds
    assert r["gpu_launches"] == 3 and r["value"] > 0 and r["dtype"] == "f32" and r["data"] == "synthetic", r
    e = r["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 131072 * 16 and e["d2h_bytes_per_step"] == 131072 * 17, e
    rf = r["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "pipes"} <= set(rf) and set(rf["pipes"]) == {"fp32", "tensor", "hbm"}, rf
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and rf["peak"] > 0
    assert rf["frac"] == max(p["frac"] for p in rf["pipes"].values()) and rf["bound"] in rf["pipes"]
    c = r["cpu_baseline"]
    assert c["kind"] == "port" and c["value"] > 0 and "sample" in c
    assert set(r["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


@pytest.mark.gpu
def test_gpu_arm_default_workload_line():
    """the default (headline) workload: BASELINE configs[4], strong scaling, valid + u; shrunk 1/64 here"""
    r = _run("--steps", "3", "--warmup", "3", "--scale", str(1 / 64), "--no-other-configs")
    assert r["config"]["workload"].startswith("cfg4") and r["scaling"] == "strong" and r["config"]["outputs"] == "valid+u"
    # 2^20 rows: the two-pass sweep (geometry pass + barrier-network pass) -> two launches per step, each with its roofline
    assert r["config"]["states"] == 1 << 20 and r["gpu_launches"] == 6 and r["launches_per_step"] == 2 and r["value"] > 0
    ps = r["roofline"]["passes"]
    assert len(ps) == 2 and all(set(p["pipes"]) == {"fp32", "tensor", "hbm"} and p["ms"] > 0 for p in ps)
    assert abs(r["roofline"]["passes_ms_sum"] - r["roofline"]["step_ms"]) < 0.35 * r["roofline"]["step_ms"]
    assert 0.05 < r["config"]["valid_fraction"] < 0.5
    assert r["e2e"]["h2d_bytes_per_step"] == (1 << 20) * 28 and r["e2e"]["d2h_bytes_per_step"] == (1 << 20) * 29
    bp = r["roofline"]["broad_phase"]
    assert bp["grid"] and 0 < bp["spheres_swept_per_state"] < 28 and bp["pairs_per_state"] < 100
    assert set(r["targets"]) == {"collision_checked_franka_configs_per_s", "cbf_qp_filtered_control_steps_per_s"}

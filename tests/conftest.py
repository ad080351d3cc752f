import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cbf-inc-rrt_b200"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def models():
    from oracle import oracle_f64
    return oracle_f64.load_models()


@pytest.fixture(scope="session")
def coracle():
    from oracle import c_oracle
    c_oracle.build()
    c_oracle.set_threads(min(8, os.cpu_count() or 1))
    return c_oracle


@pytest.fixture(scope="session")
def golden():
    out = {}
    for name in ("panda", "magician"):
        out[name] = dict(np.load(os.path.join(ROOT, "tests", "golden", f"ref_controller_{name}.npz")))
    return out


def golden_state_dict(g):
    return {k[3:]: v for k, v in g.items() if k.startswith("sd_")}


def rel_err(a, b, floor=1.0):
    """max |a-b| / max(|b|, floor): the repo's reading of "1e-5 relative" (floor = natural scale of the
    quantity, so that values crossing zero do not blow the ratio up)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import gnn_cpp_b200  # noqa: E402,F401  (registers the package under its importable name)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def directed_problem():
    """Same construction as tests/golden/make_golden.py:directed_problem."""
    from gnn_cpp_b200 import synth
    cfg = synth.Config("directed", 97, 700, [12, 9, 6, 4], False, 93)
    p = synth.make_problem(cfg)
    h = synth.hash3(cfg.seed, 7, np.arange(cfg.E, dtype=np.uint64))
    p.src = (h % np.uint64(cfg.N)).astype(np.int32)
    p.dst = ((h >> np.uint64(20)) % np.uint64(cfg.N)).astype(np.int32)
    p.src[:5] = p.dst[:5]
    p.src[5:25] = p.src[25:45]
    p.dst[5:25] = p.dst[25:45]
    return p


def load_problem(name):
    from gnn_cpp_b200 import synth
    if name == "directed":
        return directed_problem()
    return synth.make_problem(synth.CONFIGS[name])


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_err(a, ref):
    """max|a-ref| / max|ref| — the norm-wise relative error the 1e-5 FP32 bar is stated in."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    denom = max(float(np.abs(ref).max()), 1e-30)
    return float(np.abs(a - ref).max()) / denom


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.lib()
    return orc

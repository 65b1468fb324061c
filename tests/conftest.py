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


def err_stats(a, ref, sample=4_000_000):
    """Norm-wise AND element-wise error of `a` against `ref`:
        norm      max|a-ref| / max|ref|                                  (the 1e-5 bar, see rel_err)
        elem_max  max of |a-ref| / (|ref| + 1e-5 max|ref|)                 element-wise, small entries included
        elem_p999 99.9th percentile of the same quantity (over a fixed random sample of <= `sample` entries when the
                  array is larger: the exact maximum is always taken over every entry)
    Processed in row chunks so that a 2.45 M x 256 activation does not need fp64 copies of the whole matrix."""
    a = np.asarray(a); ref = np.asarray(ref)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    a2 = a.reshape(-1); r2 = ref.reshape(-1)
    n = a2.size
    step = 1 << 24
    m = 0.0
    for s0 in range(0, n, step):
        m = max(m, float(np.abs(r2[s0:s0 + step]).max(initial=0.0)))
    m = max(m, 1e-30)
    dmax, emax = 0.0, 0.0
    for s0 in range(0, n, step):
        r = r2[s0:s0 + step].astype(np.float64)
        d = np.abs(a2[s0:s0 + step].astype(np.float64) - r)
        dmax = max(dmax, float(d.max(initial=0.0)))
        emax = max(emax, float((d / (np.abs(r) + 1e-5 * m)).max(initial=0.0)))
    if n > sample:
        pick = np.random.default_rng(12345).integers(0, n, sample)
        r = r2[pick].astype(np.float64); d = np.abs(a2[pick].astype(np.float64) - r)
    else:
        r = r2.astype(np.float64); d = np.abs(a2.astype(np.float64) - r)
    p999 = float(np.quantile(d / (np.abs(r) + 1e-5 * m), 0.999)) if n else 0.0
    return {"norm": dmax / m, "elem_max": emax, "elem_p999": p999, "max_ref": m, "n": int(n)}


def parity_report(record):
    """Append one JSON line to gpurun_out/parity_report.jsonl (merged back from the GPU box; summaries of it are kept
    under profiles/) and echo it, so the element-wise error distribution of every full-size comparison is recorded."""
    import json
    line = json.dumps(record, sort_keys=True)
    print("[parity] " + line)
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
            f.write(line + "\n")
    except OSError:
        pass


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.lib()
    return orc

#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference (oracle/_ref/ref_gcn, built by
oracle/build_ref.sh from /root/reference) on deterministic synthetic problems.

Run in the build container only (needs /root/reference to build ref_gcn); the fixtures it writes are
committed so the GPU box and CI never need the reference.  Inputs are NOT stored: they are regenerated
from gnn.cpp_b200/synth.py (same seed) by the tests.

    python tests/golden/make_golden.py            # toy, tiny, tiny_pl, small directed, cora
    python tests/golden/make_golden.py --pubmed   # adds the Pubmed-shaped step (~8 min, ~8 GB RSS)
"""
import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gnn_cpp_b200  # noqa: E402
from gnn_cpp_b200 import problem_io, synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_gcn")
OUT = os.path.dirname(os.path.abspath(__file__))


def directed_problem():
    """A small DIRECTED graph with duplicates and self loops: exercises A != A^T (CSC != CSR)."""
    cfg = synth.Config("directed", 97, 700, [12, 9, 6, 4], False, 93)
    p = synth.make_problem(cfg)
    h = synth.hash3(cfg.seed, 7, np.arange(cfg.E, dtype=np.uint64))
    p.src = (h % np.uint64(cfg.N)).astype(np.int32)
    p.dst = ((h >> np.uint64(20)) % np.uint64(cfg.N)).astype(np.int32)
    p.src[:5] = p.dst[:5]          # self loops
    p.src[5:25] = p.src[25:45]     # duplicates
    p.dst[5:25] = p.dst[25:45]
    return p


def run(prob, name, step=True, structure=True, subsample=None):
    with tempfile.TemporaryDirectory() as td:
        pin = os.path.join(td, "p.gcnp")
        problem_io.write_problem(pin, prob)
        res = {}
        if structure:
            subprocess.check_call([REF, "structure", pin, os.path.join(td, "s.gcno")])
            res.update({"s_" + k: v for k, v in problem_io.read_results(os.path.join(td, "s.gcno")).items()})
        if step:
            subprocess.check_call([REF, "step", pin, os.path.join(td, "o.gcno")])
            res.update(problem_io.read_results(os.path.join(td, "o.gcno")))
    if subsample:  # keep big activations small: every k-th row
        for k in list(res):
            if k.startswith("Z") and res[k].shape[0] > subsample:
                stride = res[k].shape[0] // subsample
                res[k + "_rows"] = np.arange(0, res[k].shape[0], stride, dtype=np.int64)
                res[k] = res[k][::stride].copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
    print(name, {k: v.shape for k, v in res.items()})


def run_aswritten(prob, name, subsample=256):
    """one graph::GCNConv::forward exactly as written (ref_gcn aswritten): lin / BatchNorm / layer output"""
    with tempfile.TemporaryDirectory() as td:
        pin = os.path.join(td, "p.gcnp")
        problem_io.write_problem(pin, prob)
        subprocess.check_call([REF, "aswritten", pin, os.path.join(td, "o.gcno")])
        res = problem_io.read_results(os.path.join(td, "o.gcno"))
    n = res["aw_Z"].shape[0]
    if n > subsample:
        rows = np.arange(0, n, n // subsample, dtype=np.int64)
        res = {k: v[rows].copy() for k, v in res.items()}
        res["rows"] = rows
    np.savez_compressed(os.path.join(OUT, "aswritten_" + name + ".npz"), **res)
    print("aswritten", name, {k: v.shape for k, v in res.items()})


def run_mlp(prob, name, subsample=256):
    """nn::MLP forward + nn::tanh through the reference (ref_gcn mlp)"""
    with tempfile.TemporaryDirectory() as td:
        pin = os.path.join(td, "p.gcnp")
        problem_io.write_problem(pin, prob)
        subprocess.check_call([REF, "mlp", pin, os.path.join(td, "o.gcno")])
        res = problem_io.read_results(os.path.join(td, "o.gcno"))
    n = res["mlp_out"].shape[0]
    if n > subsample:
        rows = np.arange(0, n, n // subsample, dtype=np.int64)
        res = {k: v[rows].copy() for k, v in res.items()}
        res["rows"] = rows
    np.savez_compressed(os.path.join(OUT, "mlp_" + name + ".npz"), **res)
    print("mlp", name, {k: v.shape for k, v in res.items()})


def run_weighted(prob, name):
    """weighted adjacency + normalisation through the reference (ref_gcn structure_w)"""
    with tempfile.TemporaryDirectory() as td:
        pin = os.path.join(td, "p.gcnp")
        problem_io.write_problem(pin, prob)
        subprocess.check_call([REF, "structure_w", pin, os.path.join(td, "o.gcno")])
        res = problem_io.read_results(os.path.join(td, "o.gcno"))
    np.savez_compressed(os.path.join(OUT, "weighted_" + name + ".npz"), **res)
    print("weighted", name, {k: v.shape for k, v in res.items()})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pubmed", action="store_true")
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    if not os.path.exists(REF):
        subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh")])
    jobs = {
        "toy": lambda: run(synth.make_problem(synth.CONFIGS["toy"]), "toy"),
        "tiny": lambda: run(synth.make_problem(synth.CONFIGS["tiny"]), "tiny"),
        "directed": lambda: run(directed_problem(), "directed"),
        "tiny_pl": lambda: run(synth.make_problem(synth.CONFIGS["tiny_pl"]), "tiny_pl", subsample=256),
        "cora": lambda: run(synth.make_problem(synth.CONFIGS["cora"]), "cora", subsample=256),
    }
    if a.pubmed:
        jobs["pubmed"] = lambda: run(synth.make_problem(synth.CONFIGS["pubmed"]), "pubmed", structure=False, subsample=512)
    for k in ("toy", "tiny", "directed", "tiny_pl", "cora"):
        prob = directed_problem() if k == "directed" else synth.make_problem(synth.CONFIGS[k])
        jobs["aswritten_" + k] = (lambda prob=prob, k=k: run_aswritten(prob, k))
    for k in ("toy", "tiny", "directed", "tiny_pl"):
        prob = directed_problem() if k == "directed" else synth.make_problem(synth.CONFIGS[k])
        jobs["weighted_" + k] = (lambda prob=prob, k=k: run_weighted(prob, k))
        jobs["mlp_" + k] = (lambda prob=prob, k=k: run_mlp(prob, k))
    for k, fn in jobs.items():
        if a.only is None or a.only == k or (a.only in ("aswritten", "weighted", "mlp") and k.startswith(a.only + "_")):
            fn()


if __name__ == "__main__":
    main()

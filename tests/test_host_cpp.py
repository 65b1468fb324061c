"""The reference-shaped C++ API (gnn.cpp_b200/host: cyg::tensor, autograd nodes, nn::, graph::GCNConv, main.cpp)
on the GPU: its own unit tests, and the main.cpp training driver against the REAL reference's outputs."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden, load_problem, rel_err

HOST = os.path.join(ROOT, "gnn.cpp_b200", "host")


def test_host_binaries_built_and_fail_loudly_without_gpu():
    import torch
    for b in ("gcn_main", "host_tests"):
        assert os.path.exists(os.path.join(HOST, b)), "run __graft_entry__.build()"
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([os.path.join(HOST, "gcn_main"), "--config", "tiny", "--epochs", "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_host_unit_tests():
    r = subprocess.run([os.path.join(HOST, "host_tests")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "0 failed" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["toy", "tiny", "directed", "tiny_pl", "cora"])
def test_main_driver_matches_reference_golden(name, tmp_path):
    """One epoch of main.cpp's loop (GCNConv stack -> cross_entropy_loss -> backward) vs reference mode-B outputs."""
    from gnn_cpp_b200 import problem_io
    p, g = load_problem(name), load_golden(name)
    pin, pout = str(tmp_path / "p.gcnp"), str(tmp_path / "o.gcno")
    problem_io.write_problem(pin, p)
    r = subprocess.run([os.path.join(HOST, "gcn_main"), "--problem", pin, "--epochs", "1", "--lr", "0", "--dump", pout],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = problem_io.read_results(pout)
    L = len(p.cfg.dims) - 1
    assert abs(float(out["loss"][0]) - float(g["loss"][0])) <= 1e-5 * abs(float(g["loss"][0]))
    for l in range(1, L + 1):
        Z, A = g["Z%d" % l], out["A%d" % l]
        if "Z%d_rows" % l in g.files:
            A = A[g["Z%d_rows" % l]]
        assert rel_err(A, np.maximum(Z, 0) if l < L else Z) <= 1e-5
        assert rel_err(out["dW%d" % l], g["dW%d" % l]) <= 1e-5
        assert rel_err(out["db%d" % l], g["db%d" % l]) <= 1e-5


@pytest.mark.gpu
def test_main_driver_trains(tmp_path):
    r = subprocess.run([os.path.join(HOST, "gcn_main"), "--config", "tiny_pl", "--epochs", "30", "--lr", "0.5"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    losses = [float(l.split()[3]) for l in r.stdout.splitlines() if l.startswith("epoch")]
    assert len(losses) == 30 and losses[-1] < losses[0]


@pytest.mark.gpu
def test_reference_model_shape_trains(tmp_path):
    """the reference's own Model (src/main.cpp:10-30: pre MLP, GCNConv as written + tanh, post MLP, dropout 0.1) built
    from the C++ mirror and trained with Adam: finite, decreasing loss; two runs are bit-identical (seeded RNG, B6)."""
    cmd = [os.path.join(HOST, "gcn_main"), "--config", "tiny_pl", "--model", "reference", "--epochs", "40", "--lr", "0.01"]
    runs = []
    for _ in range(2):
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        runs.append([float(l.split()[3]) for l in r.stdout.splitlines() if l.startswith("epoch")])
    losses = runs[0]
    assert len(losses) == 40 and all(l == l and abs(l) < 1e6 for l in losses) and min(losses[-5:]) < losses[0]
    assert runs[0] == runs[1]


@pytest.mark.gpu
@pytest.mark.parametrize("name,world", [("tiny_pl", 2), ("directed", 2), ("tiny_pl", 3), ("cora", 4)])
def test_main_driver_multi_gpu_matches_single(name, world, tmp_path):
    """`gcn_main --gpus N` (one process per GPU; graph::Data::partitioned, the SpMM node exchanges aggregation inputs,
    gradients all-reduced) against the same binary on one GPU: loss, rank 0's rows of every activation, every gradient."""
    import torch
    from gnn_cpp_b200 import problem_io
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    p = load_problem(name)
    pin = str(tmp_path / "p.gcnp")
    problem_io.write_problem(pin, p)
    outs = []
    for n in (1, world):
        pout = str(tmp_path / ("o%d.gcno" % n))
        r = subprocess.run([os.path.join(HOST, "gcn_main"), "--gpus", str(n), "--problem", pin, "--epochs", "1", "--lr", "0", "--dump", pout],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(problem_io.read_results(pout))
    one, many = outs
    L = len(p.cfg.dims) - 1
    assert abs(float(one["loss"][0]) - float(many["loss"][0])) <= 1e-5 * abs(float(one["loss"][0]))
    for l in range(1, L + 1):
        n_loc = many["A%d" % l].shape[0]
        assert n_loc == (p.cfg.N + world - 1) // world
        assert rel_err(many["A%d" % l], one["A%d" % l][:n_loc]) <= 1e-5
        assert rel_err(many["dW%d" % l], one["dW%d" % l]) <= 1e-5
        assert rel_err(many["db%d" % l], one["db%d" % l]) <= 1e-5

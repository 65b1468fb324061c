"""Pins oracle/gcn_oracle.c against outputs of the REAL reference (tests/golden/*.npz, produced by
oracle/_ref/ref_gcn via tests/golden/make_golden.py).  CPU only.

The restatement follows the reference's accumulation order, so everything is required BIT-EXACT here
(far inside the 1e-5 bar the GPU path is held to)."""
import numpy as np
import pytest

import os

from conftest import GOLDEN, load_golden, load_problem

SMALL = ["toy", "tiny", "directed", "tiny_pl", "cora"]


@pytest.mark.parametrize("name", SMALL)
def test_structure_bit_exact(oracle, name):
    p, g = load_problem(name), load_golden(name)
    N = p.cfg.N
    for fill in (0, 1):  # add_self_loops(fillValue=0) removes loops, fillValue=1 adds them (graph.cpp:68-75)
        rp, ci = oracle.csr_build(p.src, p.dst, N, fill)
        rows = np.repeat(np.arange(N, dtype=np.int32), np.diff(rp))
        coo = g["s_coo_fill%d" % fill]
        assert np.array_equal(coo[0], rows)
        assert np.array_equal(coo[1], ci)
    G = oracle.Graph(p.src, p.dst, N)
    assert np.array_equal(G.deg.astype(np.float32), g["s_deg"])
    assert np.array_equal(G.dinv, g["s_dinv"])
    assert np.array_equal(G.val, g["s_ahat_val"])


def test_toy_graph_known_answer(oracle):
    """The reference's own 8-edge fixture (tests/graph.test.cpp:19-20), N=5; SURVEY.md §8c lists the expected COO."""
    src = np.array([1, 2, 3, 0, 4, 1, 2, 3], dtype=np.int32)
    dst = np.array([1, 2, 0, 1, 2, 2, 1, 1], dtype=np.int32)
    rp, ci = oracle.csr_build(src, dst, 5, 2)
    assert list(np.repeat(np.arange(5), np.diff(rp))) == [0, 1, 1, 2, 2, 3, 3, 4] and list(ci) == [1, 1, 2, 1, 2, 0, 1, 2]
    rp, ci = oracle.csr_build(src, dst, 5, 0)
    assert list(np.repeat(np.arange(5), np.diff(rp))) == [0, 1, 2, 3, 3, 4] and list(ci) == [1, 2, 1, 0, 1, 2]
    rp, ci = oracle.csr_build(src, dst, 5, 1)
    assert rp[5] == 11


@pytest.mark.parametrize("name", SMALL + ["pubmed"])
def test_train_step_bit_exact(oracle, name):
    import os
    from conftest import GOLDEN
    if not os.path.exists(os.path.join(GOLDEN, name + ".npz")):
        pytest.skip("fixture %s.npz not generated" % name)
    p, g = load_problem(name), load_golden(name)
    G = oracle.Graph(p.src, p.dst, p.cfg.N)
    W = [w.copy() for w in p.W]
    b = [x.copy() for x in p.b]
    out = oracle.train_step(G, p.cfg.dims, p.X, p.y, W, b, lr=0.0, order=0)
    assert out["loss"] == float(g["loss"][0])
    for k in g.files:
        if k.startswith("s_") or k == "loss" or k.endswith("_rows"):
            continue
        mine = out[k]
        if k + "_rows" in g.files:
            mine = mine[g[k + "_rows"]]
        assert np.array_equal(mine, g[k]), k


@pytest.mark.parametrize("name", ["tiny", "directed", "cora"])
def test_fp64_order_close_to_reference_order(oracle, name):
    """order=1 (fp64 accumulate) is the large-size checker; it must agree with the reference order."""
    from conftest import rel_err
    p = load_problem(name)
    G = oracle.Graph(p.src, p.dst, p.cfg.N)
    o0 = oracle.train_step(G, p.cfg.dims, p.X, p.y, [w.copy() for w in p.W], [x.copy() for x in p.b], order=0)
    o1 = oracle.train_step(G, p.cfg.dims, p.X, p.y, [w.copy() for w in p.W], [x.copy() for x in p.b], order=1)
    for k in o0:
        if k == "loss":
            assert abs(o0[k] - o1[k]) <= 1e-5 * abs(o0[k])
        else:
            assert rel_err(o1[k], o0[k]) <= 1e-5, k


def test_csc_is_transpose(oracle):
    p = load_problem("directed")
    G = oracle.Graph(p.src, p.dst, p.cfg.N)
    N = p.cfg.N
    A = np.zeros((N, N), dtype=np.float32)
    rows = np.repeat(np.arange(N), np.diff(G.rowptr))
    A[rows, G.colidx] = G.val
    At = np.zeros((N, N), dtype=np.float32)
    cols = np.repeat(np.arange(N), np.diff(G.colptr))
    At[cols, G.rowidx] = G.valT
    assert np.array_equal(A.T, At)
    assert np.all(np.diff(G.rowidx.astype(np.int64))[np.diff(cols) == 0] > 0)  # rows ascending inside a column
    assert np.array_equal(G.colidx[G.perm], cols)


def test_sgd_matches_torch(oracle):
    """nn::SGD's documented intent is torch.optim.SGD (include/nn.h:165-167); the reference body is broken (B4)."""
    import torch
    rng = np.random.default_rng(0)
    for kw in [dict(), dict(momentum=0.9), dict(momentum=0.9, dampening=0.1, weight_decay=1e-2),
               dict(momentum=0.8, nesterov=True, weight_decay=1e-3)]:
        p0 = rng.standard_normal(257).astype(np.float32)
        tp = torch.tensor(p0.copy(), requires_grad=True)
        opt = torch.optim.SGD([tp], lr=0.05, **kw)
        mine = p0.copy()
        vel = np.zeros_like(mine)
        for it in range(4):
            g = rng.standard_normal(257).astype(np.float32)
            tp.grad = torch.tensor(g)
            opt.step()
            oracle.sgd_step(mine, g, vel, lr=0.05, first=(it == 0), **kw)
            np.testing.assert_allclose(mine, tp.detach().numpy(), rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["tiny", "tiny_pl", "cora"])
def test_composed_step_equals_c_step(oracle, name):
    """The primitive-by-primitive composition (used by the kink-aware gradient checks) is bit-identical to
    orc_gcn_train_step, with and without an explicit ReLU-mask override equal to the natural mask."""
    p = load_problem(name)
    G = oracle.Graph(p.src, p.dst, p.cfg.N)
    L = len(p.cfg.dims) - 1
    for order in (0, 1):
        a = oracle.train_step(G, p.cfg.dims, p.X, p.y, [w.copy() for w in p.W], [x.copy() for x in p.b], order=order)
        c = oracle.train_step_composed(G, p.cfg.dims, p.X, p.y, p.W, p.b, order=order)
        masks = [c["Z%d" % l] > 0 for l in range(1, L)]
        d = oracle.train_step_composed(G, p.cfg.dims, p.X, p.y, p.W, p.b, order=order, masks=masks)
        for k, v in a.items():
            if k == "loss":
                assert v == c[k] == d[k]
            else:
                assert np.array_equal(v, c[k]) and np.array_equal(v, d[k]), k


def test_adam_matches_torch(oracle):
    """nn::Adam's intent is torch.optim.Adam (include/nn.h:180-188); the reference body (nn.cpp:419-441) is broken."""
    import torch
    rng = np.random.default_rng(1)
    for kw in [dict(), dict(weight_decay=1e-2), dict(betas=(0.8, 0.95), eps=1e-6)]:
        p0 = rng.standard_normal(301).astype(np.float32)
        tp = torch.tensor(p0.copy(), requires_grad=True)
        opt = torch.optim.Adam([tp], lr=0.01, **kw)
        mine = p0.copy(); m = np.zeros_like(mine); v = np.zeros_like(mine)
        b1, b2 = kw.get("betas", (0.9, 0.999))
        for it in range(5):
            g = rng.standard_normal(301).astype(np.float32)
            tp.grad = torch.tensor(g)
            opt.step()
            oracle.adam_step(mine, g, m, v, lr=0.01, beta1=b1, beta2=b2, eps=kw.get("eps", 1e-8),
                             weight_decay=kw.get("weight_decay", 0.0), step=it + 1)
            np.testing.assert_allclose(mine, tp.detach().numpy(), rtol=3e-6, atol=2e-7)


def test_masked_loss_and_accuracy_match_torch(oracle):
    """masked loss == cross_entropy over the sliced rows (how Data::set_mask masks are meant to be used); argmax
    accuracy with first-maximum tie breaking (tensor::argmax, tensor.h:645-648)."""
    import torch
    rng = np.random.default_rng(2)
    N, C = 500, 7
    Z = rng.standard_normal((N, C)).astype(np.float32); y = rng.integers(0, C, N).astype(np.int32)
    Z[3, 1] = Z[3, 4] = Z[3].max() + 1.0; y[3] = 1                    # tie: the first maximum wins
    mask = rng.random(N) < 0.3; mask[3] = True
    loss, dZ, nsel = oracle.softmax_xent_masked(Z, y, mask)
    zt = torch.tensor(Z, requires_grad=True)
    lt = torch.nn.functional.cross_entropy(zt[torch.tensor(mask)], torch.tensor(y[mask]).long())
    lt.backward()
    assert nsel == int(mask.sum())
    assert abs(loss - float(lt)) <= 1e-6 * abs(float(lt))
    np.testing.assert_allclose(dZ, zt.grad.numpy(), rtol=1e-5, atol=1e-8)
    assert oracle.argmax_correct(Z, y, mask) == int((torch.tensor(Z).argmax(1).numpy() == y)[mask].sum())
    assert oracle.argmax_correct(Z, y) == int((torch.tensor(Z).argmax(1).numpy() == y).sum())
    # with an all-true mask it is the plain loss of the restatement
    full, _ = oracle.softmax_xent(Z, y, order=1)
    assert abs(oracle.softmax_xent_masked(Z, y, np.ones(N, bool))[0] - full) <= 1e-6 * abs(full)


AW_TOL = 0.0    # the as-written layer, nn::MLP and nn::tanh restatements are BIT-EXACT against the real reference


@pytest.mark.parametrize("name", SMALL)
def test_gcnconv_as_written_matches_reference(oracle, name):
    """graph::GCNConv::forward exactly as written (loops removed, Linear -> BatchNorm -> ReLU,
    factorised norm, bias) restated from primitives vs outputs of the REAL reference (`ref_gcn aswritten`)."""
    p = load_problem(name)
    g = np.load(os.path.join(GOLDEN, "aswritten_%s.npz" % name))
    b = p.b[0]
    o = oracle.gcnconv_as_written(p.src, p.dst, p.cfg.N, p.X, p.W[0], b, 1 + 0.5 * b, 0.25 * b, order=0)
    rows = g["rows"] if "rows" in g.files else slice(None)
    assert np.array_equal(o["lin"][rows], g["aw_lin"])
    for k, ko in [("aw_bn", "bn"), ("aw_Z", "Z")]:
        assert np.array_equal(o[ko][rows], g[k]), k      # incl. the descending _Expr::sum of functional::var
    o1 = oracle.gcnconv_as_written(p.src, p.dst, p.cfg.N, p.X, p.W[0], b, 1 + 0.5 * b, 0.25 * b, order=1)
    assert np.abs(o1["Z"][rows] - g["aw_Z"]).max() <= 1e-5 * np.abs(g["aw_Z"]).max()


def test_batchnorm_backward_matches_torch(oracle):
    """the reference autograd loses BatchNorm's fan-out gradients (bug B2); the restated backward is pinned to torch"""
    import torch
    rng = np.random.default_rng(4)
    N, F = 300, 9
    X = rng.standard_normal((N, F)).astype(np.float32) * 2 + 1
    gamma = rng.uniform(0.5, 1.5, F).astype(np.float32); beta = rng.uniform(-0.5, 0.5, F).astype(np.float32)
    dY = rng.standard_normal((N, F)).astype(np.float32)
    for relu in (False, True):
        Y, mean, var = oracle.batchnorm_fwd(X, gamma, beta, relu=relu, order=1)
        xt = torch.tensor(X, dtype=torch.float64, requires_grad=True)
        gt = torch.tensor(gamma, dtype=torch.float64, requires_grad=True); bt = torch.tensor(beta, dtype=torch.float64, requires_grad=True)
        yt = torch.nn.functional.batch_norm(xt, None, None, gt, bt, training=True, eps=1e-5)
        if relu:
            yt = torch.relu(yt)
        np.testing.assert_allclose(Y, yt.detach().numpy(), rtol=2e-5, atol=2e-6)
        yt.backward(torch.tensor(dY, dtype=torch.float64))
        dX, dg, db = oracle.batchnorm_bwd(X, mean, var, gamma, dY, relu_out=Y if relu else None)
        np.testing.assert_allclose(dX, xt.grad.numpy(), rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(dg, gt.grad.numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(db, bt.grad.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["toy", "tiny", "directed", "tiny_pl"])
def test_weighted_adjacency_bit_exact(oracle, name):
    """edge_attr weights through edge_to_adj_mat (last write wins, graph.cpp:38-40) into the CSR
    values, diagonal forced to 1, weighted degree / D^-1/2 / A_hat — every array BIT-EXACT against the real reference
    (tests/golden/weighted_*.npz from `ref_gcn structure_w`)."""
    p = load_problem(name)
    g = np.load(os.path.join(GOLDEN, "weighted_%s.npz" % name))
    w = oracle.edge_weights(len(p.src))
    G = oracle.Graph(p.src, p.dst, p.cfg.N, w=w)
    rows = np.repeat(np.arange(p.cfg.N, dtype=np.int32), np.diff(G.rowptr))
    assert np.array_equal(rows, g["w_rows"]) and np.array_equal(G.colidx, g["w_cols"])
    assert np.array_equal(G.val0, g["w_raw_val"])
    assert np.array_equal(G.degf, g["w_deg"]) and np.array_equal(G.dinv, g["w_dinv"])
    assert np.array_equal(G.val, g["w_ahat_val"])
    # structure equals the unweighted build (weights never change which entries exist)
    G1 = oracle.Graph(p.src, p.dst, p.cfg.N)
    assert np.array_equal(G.rowptr, G1.rowptr) and np.array_equal(G.colidx, G1.colidx)


def _mlp_params(p):
    """LayerNorm on every MLP layer whose width differs from the last width (include/nn.h:201); gammas/betas as in
    oracle/ref_driver.cpp:cmd_mlp"""
    last = p.cfg.dims[-1]
    gam = [(1 + 0.5 * b).astype(np.float32) if d != last else None for b, d in zip(p.b, p.cfg.dims[1:])]
    bet = [(0.25 * b).astype(np.float32) if d != last else None for b, d in zip(p.b, p.cfg.dims[1:])]
    return gam, bet


@pytest.mark.parametrize("name", ["toy", "tiny", "directed", "tiny_pl"])
def test_mlp_layernorm_tanh_match_reference(oracle, name):
    """nn::MLP (Linear -> LayerNorm -> ReLU -> Dropout(0) chain, include/nn.h:193-214) and nn::tanh
    (src/nn.cpp:355-364) restated vs outputs of the REAL reference modules (`ref_gcn mlp`)."""
    p = load_problem(name)
    g = np.load(os.path.join(GOLDEN, "mlp_%s.npz" % name))
    rows = g["rows"] if "rows" in g.files else slice(None)
    gam, bet = _mlp_params(p)
    out = oracle.mlp_fwd(p.X, p.W, p.b, gam, bet, order=0)[-1]
    assert np.array_equal(out[rows], g["mlp_out"])
    t = oracle.tanh_fwd(out)
    assert np.array_equal(t[rows], g["tanh_out"])
    assert np.abs(np.tanh(out.astype(np.float64)) - t).max() <= 1e-6          # the as-written formula is tanh


def test_layernorm_backward_and_dropout(oracle):
    """LayerNorm backward pinned to torch (the reference loses the fan-out terms, bug B2); dropout: kept fraction,
    1/(1-p) scaling, determinism in the seed (the reference's mask is time-seeded, bug B6: unpinned)"""
    import torch
    rng = np.random.default_rng(6)
    N, F = 64, 19
    X = (rng.standard_normal((N, F)) * 3 + 0.5).astype(np.float32)
    gamma = rng.uniform(0.5, 1.5, F).astype(np.float32); beta = rng.uniform(-0.5, 0.5, F).astype(np.float32)
    dY = rng.standard_normal((N, F)).astype(np.float32)
    for relu in (False, True):
        Y, mean, rstd = oracle.layernorm_fwd(X, gamma, beta, relu=relu, order=1)
        xt = torch.tensor(X, dtype=torch.float64, requires_grad=True)
        gt = torch.tensor(gamma, dtype=torch.float64, requires_grad=True); bt = torch.tensor(beta, dtype=torch.float64, requires_grad=True)
        yt = torch.nn.functional.layer_norm(xt, (F,), gt, bt, eps=1e-5)
        if relu:
            yt = torch.relu(yt)
        np.testing.assert_allclose(Y, yt.detach().numpy(), rtol=2e-5, atol=2e-6)
        yt.backward(torch.tensor(dY, dtype=torch.float64))
        dX, dg, db = oracle.layernorm_bwd(X, mean, rstd, gamma, dY, relu_out=Y if relu else None)
        np.testing.assert_allclose(dX, xt.grad.numpy(), rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(dg, gt.grad.numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(db, bt.grad.numpy(), rtol=1e-5, atol=1e-5)
    x = np.ones(200000, np.float32)
    y = oracle.dropout_fwd(x, 0.3, 5)
    kept = y != 0
    assert abs(kept.mean() - 0.7) < 0.005 and np.allclose(y[kept], 1 / 0.7)
    assert np.array_equal(y, oracle.dropout_fwd(x, 0.3, 5)) and not np.array_equal(y, oracle.dropout_fwd(x, 0.3, 6))
    assert np.array_equal(oracle.dropout_fwd(x, 0.0, 5), x)

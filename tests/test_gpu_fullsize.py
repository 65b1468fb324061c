"""Full-size checks (BASELINE.json configs[3] and [4] shapes) through size-independent properties: the CPU oracle
would need minutes per call at these sizes, so the CUDA path is checked against invariants of the domain instead —
structure invariants of the CSR/CSC, the eigenvector identity A_hat sqrt(deg) = sqrt(deg), linearity and the
adjoint identity <A_hat P, Q> = <P, A_hat^T Q> of the aggregation kernels, agreement of the kernel variants,
bit-reproducibility, and a central finite difference of the loss along the gradient for the whole train step."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def ctx():
    from gnn_cpp_b200 import host
    c = host.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module", params=["products", "reddit"])
def big(request, ctx):
    import torch
    from gnn_cpp_b200 import host, synth
    cfg = synth.CONFIGS[request.param]
    src, dst = synth.edges(cfg.seed, cfg.E, cfg.N, cfg.powerlaw)
    g = host.Graph.build(ctx, torch.from_numpy(src).to(ctx.device), torch.from_numpy(dst).to(ctx.device), cfg.N)
    yield cfg, src, dst, g
    g.close()
    torch.cuda.empty_cache()


def test_structure_invariants_full_size(ctx, big):
    cfg, src, dst, g = big
    N = cfg.N
    e = g.export()
    rowptr, colidx, val, deg, dinv = e["rowptr"].astype(np.int64), e["colidx"], e["val"], e["deg"], e["dinv"]
    nnz = int(rowptr[-1])
    assert rowptr[0] == 0 and nnz == g.nnz and bool((np.diff(rowptr) >= 1).all())
    # exact entry count: distinct off-diagonal (src, dst) pairs + one diagonal per node (graph.cpp:21-75 semantics)
    key = src.astype(np.int64) * N + dst
    key = key[src != dst]
    exact = cfg.E < 100_000_000            # the 114.6 M-edge sort alone costs a minute of host time
    ukey = np.unique(key) if exact else None
    assert not exact or nnz == len(ukey) + N
    rows = np.repeat(np.arange(N, dtype=np.int32), np.diff(rowptr))
    inner = np.ones(nnz, bool); inner[rowptr[:-1]] = False
    assert bool((np.diff(colidx.astype(np.int64))[inner[1:]] > 0).all()), "columns strictly ascending inside a row"
    assert int((colidx == rows).sum()) == N, "diagonal present once per row"
    assert bool((colidx >= 0).all()) and bool((colidx < N).all())
    assert np.array_equal(deg, np.diff(rowptr).astype(np.int32)), "deg = rowsum(A0) + 1"
    want_dinv = (1.0 / np.sqrt(deg.astype(np.float64))).astype(np.float32)
    assert np.array_equal(dinv, want_dinv)
    assert np.array_equal(val, dinv[rows] * dinv[colidx]), "val = dinv[r] * dinv[c] in fp32"
    assert g.symmetric       # the generator symmetrises, so the CSC aliases the CSR
    # every stored entry is an input edge or a diagonal (spot check of 200k entries, exact)
    if exact:
        pick = np.random.default_rng(0).integers(0, nnz, 200000)
        k = rows[pick].astype(np.int64) * N + colidx[pick]
        pos = np.minimum(np.searchsorted(ukey, k), len(ukey) - 1)
        assert bool(((ukey[pos] == k) | (rows[pick] == colidx[pick])).all())


@pytest.mark.parametrize("variant", [1, 2])
def test_aggregation_identities_full_size(ctx, big, variant):
    import torch
    from gnn_cpp_b200 import capi
    cfg, src, dst, g = big
    N = cfg.N
    deg = torch.from_numpy(g.export()["deg"]).to(ctx.device).float()
    widths = sorted({min(a, b) for a, b in zip(cfg.dims[:-1], cfg.dims[1:])})
    try:
        capi.call("gnn_set_spmm_variant", ctx.h, variant)
        for F in widths:
            ld = (F + 3) // 4 * 4
            # eigenvector: A_hat sqrt(deg) = D^-1/2 (A+I) 1 = D^-1/2 deg = sqrt(deg), in every column
            P = torch.zeros((N, ld), device=ctx.device)
            P[:, :F] = torch.sqrt(deg)[:, None] * torch.linspace(0.5, 1.5, F, device=ctx.device)[None, :]
            Y = torch.zeros((N, ld), device=ctx.device)
            g.spmm_fwd(P[:, :F], out=Y[:, :F])
            # This identity is against the EXACT value, not against the reference's arithmetic: a hub row is a sum of n
            # (nearly) equal terms, and a sequential fp32 sum of equal terms rounds with a systematic bias of up to
            # n ulp/2 — the rows kernel (one lane group per row, like the reference's sequential dot product) shows
            # 3.6e-5 on the 6.8 K-nonzero hub row of the products-shaped graph; the merge kernel never chains more
            # than one 1,024-nonzero chunk.
            longest = int(deg.max().item()) if variant == 1 else 2048
            tol_id = max(TOL, 2e-8 * longest)
            err = float(((Y[:, :F] - P[:, :F]).abs().max() / P[:, :F].abs().max()).item())
            assert err <= tol_id, ("eigenvector", F, err)
            g.spmm_bwd(P[:, :F], out=Y[:, :F])
            assert float(((Y[:, :F] - P[:, :F]).abs().max() / P[:, :F].abs().max()).item()) <= tol_id
            # linearity and adjoint identity on random inputs (fp64 reductions of fp32 results)
            gen = torch.Generator(device=ctx.device); gen.manual_seed(F)
            A = torch.rand((N, ld), device=ctx.device, generator=gen) - 0.5
            B = torch.rand((N, ld), device=ctx.device, generator=gen) - 0.5
            A[:, F:] = 0; B[:, F:] = 0
            YA = g.spmm_fwd(A[:, :F]).clone(); YB = g.spmm_fwd(B[:, :F]).clone()
            YC = g.spmm_fwd((2.0 * A - 3.0 * B)[:, :F])
            lin = float(((YC - (2.0 * YA - 3.0 * YB)).abs().max() / YC.abs().max()).item())
            assert lin <= TOL, ("linearity", F, lin)
            ZB = g.spmm_bwd(B[:, :F])
            lhs = float((YA.double() * B[:, :F].double()).sum().item())
            rhs = float((A[:, :F].double() * ZB.double()).sum().item())
            assert abs(lhs - rhs) <= TOL * max(abs(lhs), abs(rhs), 1.0), ("adjoint", F, lhs, rhs)
            # reproducible bits
            assert torch.equal(g.spmm_fwd(A[:, :F]), YA)
            del P, Y, A, B, YA, YB, YC, ZB
    finally:
        capi.call("gnn_set_spmm_variant", ctx.h, 0)


def test_train_step_full_size_properties(ctx, big):
    """whole train step at full size: variants agree, bits reproduce, loss falls, and the loss moves along the
    computed gradient as a central finite difference predicts."""
    import torch
    from gnn_cpp_b200 import capi, host, synth
    cfg, src, dst, g = big
    L = len(cfg.dims) - 1
    X = torch.from_numpy(synth.uniform(cfg.seed, synth.STREAM_X, cfg.N * cfg.dims[0], -1.0, 1.0).reshape(cfg.N, cfg.dims[0])).to(ctx.device)
    y = torch.from_numpy(synth.labels(cfg.seed, synth.STREAM_Y, cfg.N, cfg.dims[-1])).to(ctx.device)
    W, b = synth.weights(cfg)

    import os
    from conftest import GOLDEN
    gold = np.load(os.path.join(GOLDEN, "bench_parity_%s.npz" % cfg.name))   # the exact oracle's first-step gradients + ReLU tie list

    def run(precision=1, variant=0, Ws=None, lr=0.0, steps=1, ties=False):
        capi.call("gnn_set_spmm_variant", ctx.h, variant)
        m = host.GCN(ctx, g, cfg.dims)
        m.set_option("precision", precision)
        m.set_params(Ws if Ws is not None else W, b)
        if ties:   # hidden pre-activations within 1e-5 max|Z| of zero take the oracle's side of `Z > 0` (gnn_gcn_set_relu_overrides)
            for l in range(1, L):
                m.set_relu_overrides(l, gold["kink%d_rows" % l], gold["kink%d_cols" % l].astype(np.int32), gold["kink%d_pos" % l])
        losses = [float(m.train_step(X, y, lr).cpu()[0]) for _ in range(steps)]
        grads = [m.grads(l) for l in range(1, L + 1)]
        m.close()
        capi.call("gnn_set_spmm_variant", ctx.h, 0)
        return losses, grads

    (l1,), g1 = run()
    (l1b,), g1b = run()
    assert l1 == l1b and all(np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1]) for a, c in zip(g1, g1b)), "bit-reproducible"
    # every code path — tcgen05 3xTF32 and FP32-FMA transforms, the automatic / rows / merge aggregation kernels —
    # against the EXACT oracle's gradients of this config (tests/golden/bench_parity_*.npz) at the plain 1e-5.  (Round 1
    # compared the paths with each other at 1e-3: the differences were ReLU ties — a pre-activation within rounding of
    # zero flips `Z > 0` and moves dW_1 by up to 1e-4 of its tiny largest entry — not accumulation error; with the ties
    # taken from the oracle every path is within 5e-6.)
    for kw in (dict(), dict(precision=0), dict(variant=1), dict(variant=2)):
        (lk,), gk = run(ties=True, **kw)
        assert abs(lk - float(gold["loss"][0])) <= TOL * abs(float(gold["loss"][0])), kw
        for l in range(L):
            assert rel_err(gk[l][0], gold["dW%d" % (l + 1)]) <= TOL and rel_err(gk[l][1], gold["db%d" % (l + 1)]) <= TOL, (kw, l)
    # loss falls over SGD steps
    ls, _ = run(lr=0.05, steps=4)
    assert all(np.isfinite(ls)) and ls[-1] < ls[0]
    # central difference along the (normalised) gradient direction of all weights
    gn = np.sqrt(sum(float((gw.astype(np.float64) ** 2).sum()) for gw, _ in g1))
    eps = 0.05
    Wp = [(w + eps * gw / gn).astype(np.float32) for w, (gw, _) in zip(W, g1)]
    Wm = [(w - eps * gw / gn).astype(np.float32) for w, (gw, _) in zip(W, g1)]
    (lp,), _ = run(Ws=Wp)
    (lm,), _ = run(Ws=Wm)
    fd = (lp - lm) / (2 * eps)
    assert abs(fd - gn) <= 2e-2 * gn, ("finite difference vs |dW|", fd, gn)


@pytest.mark.parametrize("precision", [0, 1])
def test_dense_transforms_full_size_vs_fp64(ctx, precision):
    """the three GEMM forms at products size (M = 2.45 M rows) against fp64 products of the same fp32 inputs computed
    by torch on the device (checker only): NT/NN on every row, TN over the full reduction."""
    import torch
    from gnn_cpp_b200 import host
    M, K, N = 2450000, 256, 100
    gen = torch.Generator(device=ctx.device); gen.manual_seed(1)
    A = torch.rand((M, K), device=ctx.device, generator=gen) - 0.5
    W = torch.rand((N, K), device=ctx.device, generator=gen) - 0.5
    G = torch.rand((M, N), device=ctx.device, generator=gen) - 0.5

    def err(x, ref):
        return float(((x.double() - ref).abs().max() / ref.abs().max()).item())

    out = host.gemm_nt(ctx, A, W, precision=precision)
    assert err(out, A.double() @ W.double().t()) <= TOL
    out = host.gemm_nn(ctx, G, W, precision=precision)          # [M, N] x [N, K]
    assert err(out, G.double() @ W.double()) <= TOL
    out = host.gemm_tn(ctx, G, A, precision=precision)          # G^T A: reduction over 2.45 M rows
    assert err(out, G.double().t() @ A.double()) <= TOL

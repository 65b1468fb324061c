"""A/B probe (not a test) of the shared-memory-staged merge kernel (GNN_SPMM_ASYNC = FIFO depth) against the register-gather
merge kernel: ms per gnn_spmm_fwd launch at the given widths and a BIT-EXACT comparison of the outputs (both kernels sum in the
same order).  python tools/spmm_async_probe.py [config] [widths...]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnn_cpp_b200  # noqa: E402,F401
from gnn_cpp_b200 import capi, host, synth  # noqa: E402

cfg_name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].isdigit() else "products"
widths = [int(a) for a in sys.argv[1:] if a.isdigit()] or [16, 32, 47, 64, 100, 128]
depths = [int(x) for x in os.environ.get("PROBE_DEPTHS", "8,12").split(",")]
cfg = synth.CONFIGS[cfg_name]
src, dst = synth.edges(cfg.seed, cfg.E, cfg.N, cfg.powerlaw)
os.environ.pop("GNN_SPMM_ASYNC", None)
ctxs = {0: host.Context(0)}
for d in depths:
    os.environ["GNN_SPMM_ASYNC"] = str(d)
    ctxs[d] = host.Context(0)
os.environ.pop("GNN_SPMM_ASYNC", None)
for c in ctxs.values():
    capi.call("gnn_set_spmm_variant", c.h, 2)  # the nonzero-balanced walk in every context
ctx = ctxs[0]
N = cfg.N
g = host.Graph.build(ctx, torch.from_numpy(src).to(ctx.device), torch.from_numpy(dst).to(ctx.device), N)
nnz = g.nnz


def p(x):
    return C.c_void_p(x.data_ptr()) if x is not None else None


def spmm(c, P, Y, F, bias=None, relu=0, mask=None):
    capi.call("gnn_spmm_fwd", c.h, g.h, p(P), P.stride(0), F, p(Y), Y.stride(0), p(bias), relu, p(mask),
              mask.stride(0) if mask is not None else 0, 1)


ok = True
for F in widths:
    ld = (F + 3) // 4 * 4
    P = torch.rand((N, ld), device=ctx.device) - 0.5
    balg = 4 * (N + 1) + nnz * (8 + 4 * F) + 4 * N * F
    ref = None
    for d, c in ctxs.items():
        Y = torch.full((N, ld), 7.0, device=ctx.device)
        for _ in range(2):
            spmm(c, P, Y, F)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            spmm(c, P, Y, F)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        same = ""
        if ref is None:
            ref = Y
        else:
            eq = torch.equal(ref[:, :F], Y[:, :F])
            ok &= eq
            same = "  bit-identical to depth 0: %s" % eq
            if not eq:
                diff = (ref[:, :F] - Y[:, :F]).abs()
                print("   max |diff| %.3e in %d rows" % (float(diff.max()), int((diff.amax(dim=1) > 0).sum())))
        print("%s F=%d depth=%d: %.3f ms  %.0f GB/s alg (%.1f%% of 6543.7)%s" %
              (cfg_name, F, d, ms, balg / ms / 1e6, balg / ms / 1e6 / 65.437, same), flush=True)
    # epilogue (bias + ReLU + mask) through both kernels
    bias = torch.rand(F, device=ctx.device) - 0.5
    mask = torch.rand((N, ld), device=ctx.device) - 0.3
    Y0 = torch.empty((N, ld), device=ctx.device)
    spmm(ctxs[0], P, Y0, F, bias, 1, mask)
    for d in depths:
        Y1 = torch.empty((N, ld), device=ctx.device)
        spmm(ctxs[d], P, Y1, F, bias, 1, mask)
        eq = torch.equal(Y0[:, :F], Y1[:, :F])
        ok &= eq
        if not eq:
            print("   epilogue F=%d depth=%d differs" % (F, d))
    del P, mask, Y0, Y1, ref
print("ALL BIT-IDENTICAL" if ok else "MISMATCH")
sys.exit(0 if ok else 1)

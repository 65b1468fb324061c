"""SpMM tuning probe (not a test): products-/reddit-shaped graph on one GPU, times gnn_spmm_fwd per width for a list
of gnn_set_spmm_variant codes.  python tools/spmm_probe.py [config] code code ..."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnn_cpp_b200  # noqa: E402,F401
from gnn_cpp_b200 import capi, host, synth  # noqa: E402

cfg_name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].isdigit() else "products"
codes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 2]
cfg = synth.CONFIGS[cfg_name]
src, dst = synth.edges(cfg.seed, cfg.E, cfg.N, cfg.powerlaw)
ctx = host.Context(0)
g = host.Graph.build(ctx, torch.from_numpy(src).to(ctx.device), torch.from_numpy(dst).to(ctx.device), cfg.N)
N, nnz = cfg.N, g.nnz
widths = sorted({min(a, b) for a, b in zip(cfg.dims[:-1], cfg.dims[1:])})
for F in widths:
    ld = (F + 3) // 4 * 4
    P = torch.rand((N, ld), device=ctx.device)
    Y = torch.empty((N, ld), device=ctx.device)
    balg = 4 * (N + 1) + nnz * (8 + 4 * F) + 4 * N * F
    for code in codes:
        capi.call("gnn_set_spmm_variant", ctx.h, code)
        for _ in range(2):
            g.spmm_fwd(P[:, :F], out=Y[:, :F])
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            g.spmm_fwd(P[:, :F], out=Y[:, :F])
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print("%s F=%d code=%d: %.3f ms  %.0f GB/s alg (%.1f%% of 6543.7)" % (cfg_name, F, code, ms, balg / ms / 1e6, balg / ms / 1e6 / 65.437), flush=True)

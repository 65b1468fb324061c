"""SpMM width / locality probe (not a test): one GPU, products- or reddit-shaped graph.
  (1) gnn_spmm_fwd at widths 4..256 with ld = round_up(F,4): what a feature-column partition (F/P columns per rank)
      would see per rank;
  (2) the same after renumbering the nodes by descending degree (hot rows contiguous -> L2 reuse), which is the
      experiment behind VERDICT r1 item 5.
python tools/spmm_width_probe.py [config] [widths...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnn_cpp_b200  # noqa: E402,F401
from gnn_cpp_b200 import capi, host, synth  # noqa: E402

cfg_name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].isdigit() else "products"
widths = [int(a) for a in sys.argv[1:] if a.isdigit()] or [4, 8, 12, 16, 24, 32, 48, 64, 100, 128, 256]
cfg = synth.CONFIGS[cfg_name]
src, dst = synth.edges(cfg.seed, cfg.E, cfg.N, cfg.powerlaw)
ctx = host.Context(0)
if os.environ.get("PROBE_VARIANT"):          # 1 = one row per lane group, 2 = nonzero-balanced merge kernel, 0 = automatic choice
    capi.call("gnn_set_spmm_variant", ctx.h, int(os.environ["PROBE_VARIANT"]))
N = cfg.N


def run(tag, s, d):
    g = host.Graph.build(ctx, torch.from_numpy(s).to(ctx.device), torch.from_numpy(d).to(ctx.device), N)
    nnz = g.nnz
    for F in widths:
        ld = (F + 3) // 4 * 4
        P = torch.rand((N, ld), device=ctx.device)
        Y = torch.empty((N, ld), device=ctx.device)
        balg = 4 * (N + 1) + nnz * (8 + 4 * F) + 4 * N * F
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
        print("%s %s F=%d: %.3f ms  %.0f GB/s alg (%.1f%% of 6543.7)  %.2f ns/nnz" %
              (cfg_name, tag, F, ms, balg / ms / 1e6, balg / ms / 1e6 / 65.437, ms * 1e6 / nnz), flush=True)
        del P, Y
    e = g.export(csc=False)
    g.close()
    return e["deg"]


deg = run("natural" + os.environ.get("PROBE_TAG", ""), src, dst)
if os.environ.get("PROBE_SORTED", "1") == "0":
    ctx.close()
    sys.exit(0)
order = np.argsort(-deg.astype(np.int64), kind="stable")          # new id -> old id
inv = np.empty(N, np.int32); inv[order] = np.arange(N, dtype=np.int32)
run("degree-sorted", inv[src], inv[dst])
ctx.close()

"""Generates tests/golden/bench_parity_<config>.npz: the CPU oracle's EXACT (order = 1: fp32 inputs, fp64
accumulation, reference op order) loss and weight/bias gradients of the first train step of a bench.py config.
bench.py compares the gradient slab the product computes at ANY GPU count with these (its `parity` object), so the
driver's SCALE lines carry oracle parity at 1/2/4/8 GPUs without the oracle having to run on the GPU box.

    python tests/golden/make_bench_parity.py products reddit arxiv pubmed cora
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gnn_cpp_b200  # noqa: E402,F401
from gnn_cpp_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

for name in sys.argv[1:] or ["products"]:
    cfg = synth.CONFIGS[name]
    t0 = time.time()
    p = synth.make_problem(cfg)
    G = orc.Graph(p.src, p.dst, cfg.N)
    orc.set_threads(os.cpu_count() or 1)
    ref = orc.train_step(G, cfg.dims, p.X, p.y, [w.copy() for w in p.W], [b.copy() for b in p.b], lr=0.0, order=1)
    out = {"loss": np.array([ref["loss"]], np.float64), "nnz": np.array([G.nnz], np.int64),
           "dims": np.asarray(cfg.dims, np.int32)}
    L = len(cfg.dims) - 1
    for l in range(1, L + 1):
        out["dW%d" % l] = ref["dW%d" % l]
        out["db%d" % l] = ref["db%d" % l]
    # ReLU tie-break list: hidden pre-activations within 1e-5 max|Z_l| of zero, where `Z > 0` may legitimately come out
    # on the other side in another correct fp32 implementation (gnn_gcn_set_relu_overrides takes the oracle's decision)
    for l in range(1, L):
        Z = ref["Z%d" % l]
        thr = 1e-5 * float(np.abs(Z).max())
        r, c = np.nonzero(np.abs(Z) <= thr)
        out["kink%d_rows" % l] = r.astype(np.int32)
        out["kink%d_cols" % l] = c.astype(np.int16)
        out["kink%d_pos" % l] = (Z[r, c] > 0).astype(np.uint8)
        print("  layer %d: %d of %d pre-activations within %.3g of zero" % (l, len(r), Z.size, thr), flush=True)
    # a fixed sample of logits rows (row ids + values) so the forward is pinned too
    rows = np.unique(np.random.default_rng(cfg.seed).integers(0, cfg.N, 4096)).astype(np.int64)
    out["logit_rows"] = rows
    out["logits"] = ref["Z%d" % L][rows]
    path = os.path.join(ROOT, "tests", "golden", "bench_parity_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%s: loss %.9f, nnz %d, %.1f s -> %s (%d KB)" % (name, ref["loss"], G.nnz, time.time() - t0, path, os.path.getsize(path) // 1024), flush=True)

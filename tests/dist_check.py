"""Multi-GPU parity check, launched by torchrun (one rank per GPU):
row-partitioned train step (NCCL all-gather of aggregation inputs, gradient all-reduce) vs the CPU oracle's
single-process step.  Exit code 0 = parity within 1e-5.  Used by tests/test_multi_gpu.py and by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnn_cpp_b200  # noqa: E402,F401
from gnn_cpp_b200 import dist_plan, host, synth  # noqa: E402


def parse_grid(world):
    """GNN_GRID=PrxPc selects the 2-D partition (csrc/trainer_grid.cu), GNN_GRID=row (or unset) the 1-D row partition."""
    e = os.environ.get("GNN_GRID", "row")
    if e in ("", "row"):
        return None
    if e == "auto":
        return dist_plan.default_grid(world)       # (bench.py's choose_grid also looks at the edge list)
    pr, pc = (int(x) for x in e.lower().split("x"))
    assert pr * pc == world, "GNN_GRID=%s does not match world %d" % (e, world)
    return pr, pc


def check_big(ctx, m, p, cfg, X, yb, lo, hi, rank, world, grid):
    """>= 200 k nodes: one train step (lr = 0) against the exact oracle at 1e-5, ReLU ties taken from the oracle
    (gnn_gcn_set_relu_overrides) exactly as bench.py's parity probe does."""
    from oracle import oracle as orc
    orc.set_threads(max(1, (os.cpu_count() or world) // world))
    L = len(cfg.dims) - 1
    G = orc.Graph(p.src, p.dst, cfg.N)
    ref = orc.train_step(G, cfg.dims, p.X, p.y, [w.copy() for w in p.W], [b.copy() for b in p.b], lr=0.0, order=1)
    ties = 0
    for l in range(1, L):
        Z = ref["Z%d" % l]
        r, c = np.nonzero(np.abs(Z[lo:hi]) <= 1e-5 * float(np.abs(Z).max()))
        m.set_relu_overrides(l, r.astype(np.int32), c.astype(np.int32), (Z[lo:hi][r, c] > 0))
        ties += len(r)
    loss = float(m.train_step(X, yb, 0.0).cpu()[0])
    errs = [abs(loss - ref["loss"]) / abs(ref["loss"])]
    for l in range(1, L + 1):
        dW, db = m.grads(l)
        errs.append(np.abs(dW - ref["dW%d" % l]).max() / np.abs(ref["dW%d" % l]).max())
        errs.append(np.abs(db - ref["db%d" % l]).max() / np.abs(ref["db%d" % l]).max())
    errs.append(np.abs(m.activation(L) - ref["Z%d" % L][lo:hi]).max() / np.abs(ref["Z%d" % L]).max())
    t = torch.tensor([max(errs)], device=ctx.device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)            # logits are checked on every rank's rows
    if rank == 0:
        print("[dist_check] %s (N=%d) world=%d grid=%s mode=%d loss=%.6f ref=%.6f max_rel_err=%.2e (tol 1e-5, %d relu ties on rank 0)" %
              (cfg.name, cfg.N, world, grid, m.exchange_mode(), loss, ref["loss"], float(t.item()), ties), flush=True)
    return float(t.item()) <= 1e-5


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = host.Context(local)
    ctx.init_comm_from_torch()
    ok = True
    grid = parse_grid(world)
    big = synth.Config("mid_pl", 250000, 5000000, [40, 96, 64, 19], True, 96)     # >= 200 k nodes
    cfgs = [synth.CONFIGS["tiny_pl"], synth.Config("odd", 1237, 9000, [20, 33, 12, 6], True, 95),
            synth.Config("widen", 1500, 12000, [12, 8, 24, 5], True, 97)]      # layer 2 aggregates first with l > 1
    # a graph with locality: every rank's structure block needs only a band of remote rows, so the halo-only exchange
    # (send lists) and the interior/boundary split of csrc/trainer_grid.cu are what runs (2-D partition modes)
    cfgs.append(synth.Config("local", 40000, 600000, [24, 40, 16, 6], False, 99, 1500))
    if os.environ.get("GNN_DIST_BIG", "1") != "0":
        cfgs.append(big)
    for cfg in cfgs:
        for mask in (None, 0, 0xFF):
            if cfg.name == "local" and mask == 0xFF:
                continue
            if cfg.name == "mid_pl" and mask is not None:
                continue                      # the big graph runs the automatic layer order only
            p = synth.make_problem(cfg)
            gfull = host.Graph.build(ctx, p.src, p.dst, cfg.N)
            chunk = (cfg.N + world - 1) // world
            lo, hi = min(cfg.N, rank * chunk), min(cfg.N, (rank + 1) * chunk)
            if grid is None:
                g = gfull.slice_rows(lo, hi)
            else:
                (rlo, rhi), (glo, ghi) = dist_plan.grid_partition(cfg.N, world, grid[1], rank)
                assert (rlo, rhi) == (lo, hi)
                g = gfull.slice_rows(glo, ghi)
            ld0 = (cfg.dims[0] + 3) // 4 * 4
            Xb = torch.zeros((chunk, ld0), device=ctx.device)
            X = Xb[:hi - lo, :cfg.dims[0]]
            X.copy_(torch.from_numpy(p.X[lo:hi]))
            yb = torch.zeros(chunk, dtype=torch.int32, device=ctx.device)
            yb[:hi - lo].copy_(torch.from_numpy(p.y[lo:hi]))
            m = host.GCN(ctx, g, cfg.dims) if grid is None else host.GCN(ctx, g, cfg.dims, grid=grid, n_loc=hi - lo)
            if mask is not None:
                m.set_option("agg_first_mask", mask)
            m.set_params(p.W, p.b)
            if cfg.name == "local" and grid is not None:
                st = m.exchange_stats()
                if rank == 0:
                    print("[dist_check] local: %s" % st, flush=True)
                if os.environ.get("GNN_HALO") is None and os.environ.get("GNN_SPLIT") is None and grid[0] > 1:
                    # with ONE row group (1 x P) every rank's structure block spans all rows and needs every row of its
                    # column slice: halo fraction 1.0 by construction, nothing to assert
                    ok &= st["halo_only_exchange"] and st["halo_fraction"] < 0.9   # (2 row groups of 4 ranks: 0.58; rows x 1: 0.03-0.09)
                if os.environ.get("GNN_HALO") is None and os.environ.get("GNN_SPLIT") is None:
                    ok &= st["interior_boundary_split"] == (st["interior_fraction"] >= 0.1)
            if cfg.name == "mid_pl":
                ok &= check_big(ctx, m, p, cfg, X, yb, lo, hi, rank, world, grid)
                m.close(); g.close(); gfull.close()
                continue
            # three SGD steps, each held to the plain 1e-5 against the oracle stepping from the SAME parameters (the oracle
            # continues from the product's parameters, so per-step differences do not compound)
            L = len(cfg.dims) - 1
            from oracle import oracle as orc
            G = orc.Graph(p.src, p.dst, cfg.N) if rank == 0 else None
            W = [w.copy() for w in p.W]; b = [x.copy() for x in p.b]
            errs, losses = [], []
            for it in range(3):
                losses.append(float(m.train_step(X, yb, 0.05).cpu()[0]))
                grads = [m.grads(l) for l in range(1, L + 1)]
                params = [m.params(l) for l in range(1, L + 1)]
                logits = m.activation(L)
                if rank == 0:
                    ref = orc.train_step(G, cfg.dims, p.X, p.y, W, b, lr=0.05, order=1)
                    errs.append(abs(losses[it] - ref["loss"]) / abs(ref["loss"]))
                    for l in range(L):
                        errs.append(np.abs(grads[l][0] - ref["dW%d" % (l + 1)]).max() / np.abs(ref["dW%d" % (l + 1)]).max())
                        errs.append(np.abs(grads[l][1] - ref["db%d" % (l + 1)]).max() / np.abs(ref["db%d" % (l + 1)]).max())
                        errs.append(np.abs(params[l][0] - W[l]).max() / np.abs(W[l]).max())
                        W[l][...] = params[l][0]; b[l][...] = params[l][1]
                    errs.append(np.abs(logits - ref["Z%d" % L][lo:hi]).max() / np.abs(ref["Z%d" % L]).max())
            if rank == 0:
                ok &= max(errs) <= 1e-5
                print("[dist_check] %s world=%d grid=%s mode=%d mask=%s loss=%.6f ref=%.6f max_rel_err=%.2e" %
                      (cfg.name, world, grid, m.exchange_mode(), mask, losses[-1], ref["loss"], max(errs)), flush=True)
            m.close(); g.close(); gfull.close()
    flag = torch.tensor([1 if ok else 0], device=ctx.device)
    dist.broadcast(flag, 0)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

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
from gnn_cpp_b200 import host, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = host.Context(local)
    ctx.init_comm_from_torch()
    ok = True
    for cfg in [synth.CONFIGS["tiny_pl"], synth.Config("odd", 1237, 9000, [20, 33, 12, 6], True, 95)]:
        for mask in (None, 0, 0xFF):
            p = synth.make_problem(cfg)
            gfull = host.Graph.build(ctx, p.src, p.dst, cfg.N)
            chunk = (cfg.N + world - 1) // world
            lo, hi = min(cfg.N, rank * chunk), min(cfg.N, (rank + 1) * chunk)
            g = gfull.slice_rows(lo, hi)
            ld0 = (cfg.dims[0] + 3) // 4 * 4
            Xb = torch.zeros((chunk, ld0), device=ctx.device)
            X = Xb[:hi - lo, :cfg.dims[0]]
            X.copy_(torch.from_numpy(p.X[lo:hi]))
            yb = torch.zeros(chunk, dtype=torch.int32, device=ctx.device)
            yb[:hi - lo].copy_(torch.from_numpy(p.y[lo:hi]))
            m = host.GCN(ctx, g, cfg.dims)
            if mask is not None:
                m.set_option("agg_first_mask", mask)
            m.set_params(p.W, p.b)
            losses = [float(m.train_step(X, yb, 0.05).cpu()[0]) for _ in range(3)]
            L = len(cfg.dims) - 1
            grads = [m.grads(l) for l in range(1, L + 1)]
            params = [m.params(l) for l in range(1, L + 1)]
            logits = m.activation(L)
            if rank == 0:
                from oracle import oracle as orc
                G = orc.Graph(p.src, p.dst, cfg.N)
                W = [w.copy() for w in p.W]; b = [x.copy() for x in p.b]
                for it in range(3):
                    ref = orc.train_step(G, cfg.dims, p.X, p.y, W, b, lr=0.05, order=1)
                    e = abs(losses[it] - ref["loss"]) / abs(ref["loss"])
                    ok &= e <= 2e-5
                errs = []
                for l in range(L):
                    errs.append(np.abs(grads[l][0] - ref["dW%d" % (l + 1)]).max() / np.abs(ref["dW%d" % (l + 1)]).max())
                    errs.append(np.abs(grads[l][1] - ref["db%d" % (l + 1)]).max() / np.abs(ref["db%d" % (l + 1)]).max())
                    errs.append(np.abs(params[l][0] - W[l]).max() / np.abs(W[l]).max())
                errs.append(np.abs(logits - ref["Z%d" % L][lo:hi]).max() / np.abs(ref["Z%d" % L]).max())
                ok &= max(errs) <= 2e-5
                print("[dist_check] %s world=%d mask=%s loss=%.6f ref=%.6f max_rel_err=%.2e" %
                      (cfg.name, world, mask, losses[-1], ref["loss"], max(errs)), flush=True)
            m.close(); g.close(); gfull.close()
    flag = torch.tensor([1 if ok else 0], device=ctx.device)
    dist.broadcast(flag, 0)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

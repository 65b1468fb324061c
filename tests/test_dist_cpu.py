"""N > 1 on CPU (world_size 2, gloo): the row-partition arithmetic and the exchange schedule of the multi-GPU train
step, executed with the oracle's kernels per rank and torch.distributed collectives, must reproduce the
single-process oracle step.  (The CUDA/NCCL version of the same schedule is checked by tests/dist_check.py.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, rel_err


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _rank_main(rank, world, port, name, q):
    import sys
    sys.path.insert(0, ROOT)
    import gnn_cpp_b200  # noqa: F401
    from gnn_cpp_b200 import dist_plan, synth
    from oracle import oracle as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc.set_threads(1)
    cfg = synth.CONFIGS[name] if name in synth.CONFIGS else synth.Config("odd", 1237, 9000, [20, 33, 12, 6], True, 95)
    p = synth.make_problem(cfg)
    N, dims, L = cfg.N, cfg.dims, len(cfg.dims) - 1
    G = orc.Graph(p.src, p.dst, N)
    ptr = dist_plan.partition(N, world)
    assert np.array_equal(ptr, orc.partition_ptr(N, world))
    lo, hi = int(ptr[rank]), int(ptr[rank + 1])
    chunk = dist_plan.chunk_rows(N, world)
    fr, fc, fv = orc.partition_rows(G.rowptr, G.colidx, G.val, lo, hi)        # rows of A_hat
    br, bc, bv = orc.partition_rows(G.colptr, np.ascontiguousarray(G.rowidx), G.valT, lo, hi)  # rows of A_hat^T
    n_loc = hi - lo

    def gather(local):                       # all-gather equal (padded) chunks into global row order
        buf = np.zeros((chunk, local.shape[1]), np.float32); buf[:n_loc] = local
        outs = [torch.zeros(chunk, local.shape[1]) for _ in range(world)]
        dist.all_gather(outs, torch.from_numpy(buf))
        return np.concatenate([o.numpy() for o in outs])[:N]

    def spmm_f(X):
        return orc.spmm(n_loc, fr, fc, fv, gather(X), order=1)

    def spmm_b(X):
        return orc.spmm(n_loc, br, bc, bv, gather(X), order=1)

    af = dist_plan.layer_order(dims)
    n_gathers = 0
    H, M = [p.X[lo:hi]], [None] * (L + 1)
    for l in range(1, L + 1):
        W, b = p.W[l - 1], p.b[l - 1]
        if af[l - 1]:
            M[l] = spmm_f(H[l - 1]); n_gathers += 1
            Z = orc.gemm_nt(M[l], W, order=1) + b
        else:
            Z = spmm_f(orc.gemm_nt(H[l - 1], W, order=1)) + b; n_gathers += 1
        H.append(np.maximum(Z, 0) if l < L else Z)
    loss_loc, dZ = orc.softmax_xent(H[L], p.y[lo:hi], order=1)
    dZ = dZ * (n_loc / N)                                     # the kernel divides by the GLOBAL node count
    loss_sum = torch.tensor([loss_loc * n_loc / N]); dist.all_reduce(loss_sum)
    grads = []
    for l in range(L, 0, -1):
        W = p.W[l - 1]
        db = orc.bias_grad(dZ, order=1)
        if af[l - 1]:
            dW = orc.gemm_tn(dZ, M[l], order=1)
            if l > 1:
                dZn = spmm_b(orc.gemm_nn(dZ, W, order=1)) * (H[l - 1] > 0); n_gathers += 1
        else:
            dP = spmm_b(dZ); n_gathers += 1
            dW = orc.gemm_tn(dP, H[l - 1], order=1)
            if l > 1:
                dZn = orc.gemm_nn(dP, W, order=1) * (H[l - 1] > 0)
        for g in (dW, db):
            t = torch.from_numpy(np.ascontiguousarray(g)); dist.all_reduce(t)
        grads.append((dW, db))
        if l > 1:
            dZ = dZn.astype(np.float32)
    assert n_gathers == len(dist_plan.exchange_schedule(dims))
    if rank == 0:
        ref = orc.train_step(G, dims, p.X, p.y, [w.copy() for w in p.W], [x.copy() for x in p.b], order=1)
        errs = [abs(float(loss_sum) - ref["loss"]) / abs(ref["loss"])]
        for i, l in enumerate(range(L, 0, -1)):
            errs.append(rel_err(grads[i][0], ref["dW%d" % l])); errs.append(rel_err(grads[i][1], ref["db%d" % l]))
        q.put(max(errs))
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["tiny_pl", "odd"])
def test_row_partitioned_schedule_world2_gloo(name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, name, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=300)
        assert pr.exitcode == 0
    assert q.get(timeout=5) <= 1e-5


def test_plan_matches_trainer_counts():
    from gnn_cpp_b200 import dist_plan
    dims = [100, 256, 256, 47]
    assert dist_plan.layer_order(dims) == [True, False, False]
    sched = dist_plan.exchange_schedule(dims)
    assert [w for *_, w in sched] == [100, 256, 48, 48, 256]          # fwd L1..L3, bwd L3, L2 (L1 needs none)
    assert dist_plan.comm_bytes_per_step(2450000, dims, 8) > 5e9
    ptr = dist_plan.partition(10, 3)
    assert list(ptr) == [0, 4, 8, 10]

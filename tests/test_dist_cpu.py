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


def _rank_main(rank, world, port, name, q, panel_cols=8, grid=None):
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

    def aggregate(X, ptr_, idx_, val_):
        """the trainer's exchange (csrc/trainer.cu): the rank's rows are written panel-major into its block of a gather
        region [panel][world][chunk, w]; every panel is exchanged on its own (each rank's panel is one contiguous tile
        at dist_plan.tile_offsets) and aggregated as soon as it is complete — SpMM acts on columns independently."""
        F = X.shape[1]
        ldw = dist_plan.padded(F)
        region = np.zeros(world * chunk * ldw, np.float32)
        tiles = dist_plan.tile_offsets(ldw, panel_cols, world, chunk, rank)
        Y = np.empty((n_loc, F), np.float32)
        for pi, ((c0, w), (off, n)) in enumerate(zip(dist_plan.panels(ldw, panel_cols), tiles)):
            f = min(w, F - c0)
            own = region[off:off + n].reshape(chunk, w)
            own[:n_loc, :f] = X[:, c0:c0 + f]                                  # produced in place by the rank
            outs = [torch.zeros(chunk * w) for _ in range(world)]
            dist.all_gather(outs, torch.from_numpy(region[off:off + n].copy()))
            for q, o in enumerate(outs):                                       # lands at rank q's tile offset
                qoff = dist_plan.tile_offsets(ldw, panel_cols, world, chunk, q)[pi][0]
                region[qoff:qoff + n] = o.numpy()
            panel = region[off - rank * n: off - rank * n + world * n].reshape(world * chunk, w)   # global row order
            Y[:, c0:c0 + f] = orc.spmm(n_loc, ptr_, idx_, val_, panel[:N, :f], order=1)
        return Y

    if grid is not None:
        # 2-D partition (csrc/trainer_grid.cu): world = Pr x Pc; the rank aggregates the structure rows of its row group
        # over its column slice of ALL nodes' rows, and every output row goes to the rank that owns it
        Pr, Pc = grid
        assert Pr * Pc == world
        (rlo, rhi), (glo, ghi) = dist_plan.grid_partition(N, world, Pc, rank)
        assert (rlo, rhi) == (lo, hi)
        gi, gj = rank // Pc, rank % Pc
        gfr, gfc, gfv = orc.partition_rows(G.rowptr, G.colidx, G.val, glo, ghi)
        gbr, gbc, gbv = orc.partition_rows(G.colptr, np.ascontiguousarray(G.rowidx), G.valT, glo, ghi)

        def aggregate_grid(X, ptr_, idx_, val_):
            F = X.shape[1]
            ldw = dist_plan.padded(F)
            slices = [dist_plan.col_slice(ldw, Pc, j) for j in range(Pc)]
            assert slices[0][0] == 0 and sum(w for _, w in slices) == ldw and all(w % 4 == 0 for _, w in slices)
            assert all(slices[j][0] + slices[j][1] == slices[j + 1][0] for j in range(Pc - 1))
            # 1. rows -> columns: rank q's rows land at rows [q chunk, ...) of every rank's gathered slice matrix
            mine = np.zeros((chunk, ldw), np.float32); mine[:n_loc, :F] = X
            blocks = [torch.zeros(chunk * ldw) for _ in range(world)]
            dist.all_gather(blocks, torch.from_numpy(mine.reshape(-1)))
            c0, w = slices[gj]
            f = max(0, min(w, F - c0))
            PC = np.zeros((world * chunk, max(w, 1)), np.float32)
            for q_, blk in enumerate(blocks):
                PC[q_ * chunk:(q_ + 1) * chunk, :w] = blk.numpy().reshape(chunk, ldw)[:, c0:c0 + w]
            # 2. aggregation over the row group's structure rows at the slice width
            part = np.zeros((Pc * chunk, ldw), np.float32)
            if f > 0 and ghi > glo:
                part[:ghi - glo, c0:c0 + f] = orc.spmm(ghi - glo, ptr_, idx_, val_, np.ascontiguousarray(PC[:N, :f]), order=1)
            # 3. columns -> rows: output row r of the group belongs to rank gi Pc + r // chunk, local row r % chunk
            parts = [torch.zeros(Pc * chunk * ldw) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(part.reshape(-1)))
            Y = np.zeros((n_loc, F), np.float32)
            for j in range(Pc):
                src_rank = gi * Pc + j
                cj, wj = slices[j]
                fj = max(0, min(wj, F - cj))
                blk = parts[src_rank].numpy().reshape(Pc * chunk, ldw)
                Y[:, cj:cj + fj] = blk[gj * chunk: gj * chunk + n_loc, cj:cj + fj]
            return Y

    def spmm_f(X):
        X = np.ascontiguousarray(X, dtype=np.float32)
        return aggregate_grid(X, gfr, gfc, gfv) if grid is not None else aggregate(X, fr, fc, fv)

    def spmm_b(X):
        X = np.ascontiguousarray(X, dtype=np.float32)
        return aggregate_grid(X, gbr, gbc, gbv) if grid is not None else aggregate(X, br, bc, bv)

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


@pytest.mark.parametrize("name,world", [("tiny_pl", 2), ("odd", 2), ("odd", 3)])
def test_row_partitioned_schedule_world2_gloo(name, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, name, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=300)
        assert pr.exitcode == 0
    assert q.get(timeout=5) <= 1e-5


@pytest.mark.parametrize("name,world,grid", [("odd", 2, (1, 2)), ("odd", 4, (2, 2)), ("tiny_pl", 3, (1, 3)), ("odd", 3, (3, 1))])
def test_grid_partitioned_schedule_gloo(name, world, grid):
    """the 2-D (row groups x column groups) schedule of csrc/trainer_grid.cu — column-slice scatter, aggregation over the
    row group's structure at the slice width, row exchange — with the library's own partition rules."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, name, q, 8, grid)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=300)
        assert pr.exitcode == 0
    assert q.get(timeout=5) <= 1e-5


def test_grid_partition_rules():
    from gnn_cpp_b200 import dist_plan
    assert [dist_plan.col_slice(256, 4, j) for j in range(4)] == [(0, 64), (64, 64), (128, 64), (192, 64)]
    assert [dist_plan.col_slice(100, 4, j) for j in range(4)] == [(0, 28), (28, 24), (52, 24), (76, 24)]
    assert [dist_plan.col_slice(48, 8, j) for j in range(8)] == [(0, 8), (8, 8), (16, 8), (24, 8), (32, 4), (36, 4), (40, 4), (44, 4)]
    assert [dist_plan.col_slice(4, 2, j) for j in range(2)] == [(0, 4), (4, 0)]          # narrower than the group count
    (rows, grp) = dist_plan.grid_partition(2450000, 8, 4, 5)
    assert rows == (5 * 306250, 6 * 306250) and grp == (4 * 306250, 2450000)
    (rows, grp) = dist_plan.grid_partition(10, 4, 2, 3)
    assert rows == (9, 10) and grp == (6, 10)
    # the automatic choice: power-law endpoints -> 2 row groups x column groups; a band graph -> rows x 1 (halo-only exchange)
    from gnn_cpp_b200 import synth
    s1, d1 = synth.edges(5, 200000, 50000, True)
    s2, d2 = synth.edges(5, 200000, 50000, False, band=500)
    assert dist_plan.choose_grid(50000, 8, s1, d1) == (2, 4) and dist_plan.choose_grid(50000, 4, s1, d1) == (2, 2)
    assert dist_plan.choose_grid(50000, 2, s1, d1) == (1, 2) and dist_plan.choose_grid(50000, 1, s1, d1) is None
    assert dist_plan.choose_grid(50000, 8, s2, d2) == (8, 1) and dist_plan.choose_grid(50000, 4, s2, d2) == (4, 1)
    # received bytes per rank and step, products-shaped: the 2 x 4 grid moves ~2.7x less than the all-gather
    dims = [100, 256, 256, 47]
    ag = dist_plan.comm_bytes_per_step(2450000, dims, 8)
    gr = dist_plan.comm_bytes_per_step_grid(2450000, dims, 8, 4)
    assert 2.3 < ag / gr < 3.2, (ag, gr)


def test_panel_tiling_rule():
    """column panels of a gather region (the library's rule through gnn_partition_panels_h): multiples of 4, at most 4
    panels, exact cover; the per-rank tiles of all ranks tile the region without gaps or overlap."""
    from gnn_cpp_b200 import dist_plan
    assert dist_plan.panels(256, 128) == [(0, 128), (128, 128)]
    assert dist_plan.panels(100, 128) == [(0, 100)]
    assert dist_plan.panels(48, 64) == [(0, 48)]
    assert dist_plan.panels(256, 64) == [(0, 64), (64, 64), (128, 64), (192, 64)]
    assert dist_plan.panels(604, 64) == [(0, 152), (152, 152), (304, 152), (456, 148)]     # wider than 4 x 64: 4 equal panels
    for ldw in (4, 48, 100, 128, 132, 256, 604, 1024):
        for pc in (16, 64, 128, 1 << 20):
            ps = dist_plan.panels(ldw, pc)
            assert 1 <= len(ps) <= 4 and ps[0][0] == 0 and sum(w for _, w in ps) == ldw
            assert all(w % 4 == 0 and w > 0 for _, w in ps) and all(ps[i][0] + ps[i][1] == ps[i + 1][0] for i in range(len(ps) - 1))
            world, chunk = 3, 10
            cover = np.zeros(world * chunk * ldw, np.int32)
            for r in range(world):
                for off, n in dist_plan.tile_offsets(ldw, pc, world, chunk, r):
                    cover[off:off + n] += 1
            assert bool((cover == 1).all())


def test_plan_matches_trainer_counts():
    from gnn_cpp_b200 import dist_plan
    dims = [100, 256, 256, 47]
    assert dist_plan.layer_order(dims) == [True, False, False]
    sched = dist_plan.exchange_schedule(dims)
    assert [w for *_, w in sched] == [100, 256, 48, 48, 256]          # fwd L1..L3, bwd L3, L2 (L1 needs none)
    assert dist_plan.comm_bytes_per_step(2450000, dims, 8) > 5e9
    ptr = dist_plan.partition(10, 3)
    assert list(ptr) == [0, 4, 8, 10]

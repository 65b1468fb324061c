"""Host-side description of the row-partitioned (multi-GPU) train step: who owns which rows, which tensors are
all-gathered when, and how many bytes cross NVLink per step.  Mirrors the schedule csrc/trainer.cu executes
(same layer-order rule, same padded widths) so the CPU tests and bench.py can reason about it without a GPU."""
import ctypes

import numpy as np

from . import capi


def partition(N, world):
    """part_ptr[p] = min(N, p*ceil(N/world)) — computed by the library (gnn_partition_ptr_h)."""
    out = np.empty(world + 1, dtype=np.int64)
    capi.call("gnn_partition_ptr_h", int(N), int(world), out.ctypes.data_as(ctypes.c_void_p))
    return out


def chunk_rows(N, world):
    return (N + world - 1) // world


def padded(F):
    return (F + 3) // 4 * 4


def panels(ldw, panel_cols=128):
    """[(first column, width)] of the column panels of a gathered matrix of padded width ldw — the library's rule
    (gnn_partition_panels_h, the one csrc/trainer.cu uses)."""
    c0 = np.zeros(8, np.int32); w = np.zeros(8, np.int32); n = ctypes.c_int32(0)
    capi.call("gnn_partition_panels_h", int(ldw), int(panel_cols), c0.ctypes.data_as(ctypes.c_void_p),
              w.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n))
    return [(int(c0[i]), int(w[i])) for i in range(n.value)]


def tile_offsets(ldw, panel_cols, world, chunk, rank):
    """float offsets of every rank-`rank` panel inside a gather region [panel][world][chunk, w]: (offset, floats)"""
    out, base = [], 0
    for c0, w in panels(ldw, panel_cols):
        assert base == world * chunk * c0
        out.append((base + rank * chunk * w, chunk * w))
        base += world * chunk * w
    return out


def layer_order(dims):
    """True = aggregate first (A_hat H, then the GEMM): chosen when the input is narrower than the output."""
    return [dims[l - 1] < dims[l] for l in range(1, len(dims))]


def exchange_schedule(dims):
    """[(phase, layer, tensor, width)] of every all-gather in one train step, in execution order."""
    L = len(dims) - 1
    af = layer_order(dims)
    sched = []
    for l in range(1, L + 1):
        if af[l - 1]:
            sched.append(("fwd", l, "H%d" % (l - 1), padded(dims[l - 1])))
        else:
            sched.append(("fwd", l, "P%d" % l, padded(dims[l])))
    for l in range(L, 0, -1):
        if af[l - 1]:
            if l > 1:
                sched.append(("bwd", l, "dM%d" % l, padded(dims[l - 1])))
        else:
            sched.append(("bwd", l, "dZ%d" % l, padded(dims[l])))
    return sched


def comm_bytes_per_step(N, dims, world):
    """bytes RECEIVED per rank per step: all-gathers of (world-1) remote chunks + the gradient all-reduce."""
    c = chunk_rows(N, world)
    gathers = sum(w for _, _, _, w in exchange_schedule(dims)) * 4 * c * (world - 1)
    n_params = sum(dims[l] * dims[l - 1] + dims[l] for l in range(1, len(dims)))
    return gathers + (2 * 4 * n_params * (world - 1)) // max(world, 1)


# ---- 2-D partition (gnn_gcn_create_grid, csrc/trainer_grid.cu) ---------------------------------------------------
def col_slice(ldw, Pc, j):
    """(first column, width) of the column slice group j of Pc receives of a matrix of padded width ldw — the library's
    rule (gnn_partition_col_slice_h)."""
    c0 = ctypes.c_int32(0); w = ctypes.c_int32(0)
    capi.call("gnn_partition_col_slice_h", int(ldw), int(Pc), int(j), ctypes.byref(c0), ctypes.byref(w))
    return int(c0.value), int(w.value)


def grid_partition(N, world, Pc, rank):
    """((rows_lo, rows_hi), (group_lo, group_hi)): activation rows the rank owns and the structure rows of its row group."""
    v = [ctypes.c_int64(0) for _ in range(4)]
    capi.call("gnn_partition_grid_h", int(N), int(world), int(Pc), int(rank), *[ctypes.byref(x) for x in v])
    return (int(v[0].value), int(v[1].value)), (int(v[2].value), int(v[3].value))


def default_grid(world):
    """(Pr, Pc) for a graph WITHOUT locality (every rank needs practically every remote row, as on the synthetic power-law
    graphs).  Measured on the products-shaped step, ms (profiles/r2_scaling_products.md):
        2 GPUs: 1-D rows 29.0 | 1x2 27.5          4 GPUs: rows 17.6 | 2x2 16.7 | 1x4 16.9          8 GPUs: rows 14.8 | 2x4 10.2 | 4x2 11.7 | 1x8 12.5
    Two row groups bound the loss of SpMM efficiency from narrow column slices while cutting the received bytes."""
    return (1, 2) if world == 2 else (2, world // 2)


def choose_grid(N, world, src, dst, sample=2_000_000):
    """Partition of the aggregation for this graph: when most edges stay inside a rank's contiguous row block (a graph
    WITH locality) the rows x 1 partition with the halo-only exchange and the interior/boundary overlap moves almost
    nothing (products-sized band graph, 4 GPUs: 8.3 ms vs 9.8 ms for 2x2 and 13.6 ms for the all-gather); otherwise
    default_grid.  Decided from a sample of the edge list: fraction of edges whose endpoints live on different ranks."""
    if world <= 1:
        return None
    c = chunk_rows(N, world)
    E = len(src)
    if E == 0:
        return (world, 1)
    step = max(1, E // sample)
    remote = float(np.mean((np.asarray(src[::step], dtype=np.int64) // c) != (np.asarray(dst[::step], dtype=np.int64) // c)))
    return (world, 1) if remote < 0.25 else default_grid(world)


def comm_bytes_per_step_grid(N, dims, world, Pc):
    """bytes RECEIVED per rank per step under the 2-D partition (widest column group): N F / Pc minus the own rows for the
    column-slice scatter + c F (Pc-1)/Pc for the row exchange fused into the aggregation, + the gradient all-reduce."""
    c = chunk_rows(N, world)
    total = 0
    for _, _, _, w in exchange_schedule(dims):
        wj = max(col_slice(w, Pc, j)[1] for j in range(Pc))
        total += 4 * (world - 1) * c * wj + 4 * c * (w - min(col_slice(w, Pc, j)[1] for j in range(Pc)))
    n_params = sum(dims[l] * dims[l - 1] + dims[l] for l in range(1, len(dims)))
    return total + (2 * 4 * n_params * (world - 1)) // max(world, 1)

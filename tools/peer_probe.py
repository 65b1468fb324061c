"""Peer-arena exchange probe (not a test): torchrun --nproc-per-node N tools/peer_probe.py
Times gnn_peer_gather_begin + wait of one aggregation-sized block per rank (env GNN_PEER_COPY=ce|sm, GNN_PEER_CTAS)."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnn_cpp_b200  # noqa: E402,F401
from gnn_cpp_b200 import capi, host  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.pop("NCCL_DEBUG", None)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = host.Context(local)
ctx.init_comm_from_torch()
rows = (2450000 + world - 1) // world
for F in (256, 128, 48):
    block = rows * F * 4
    arena = C.c_void_p()
    capi.call("gnn_peer_arena_create", ctx.h, world * block, C.byref(arena))
    bar = torch.zeros(1, device=ctx.device)
    ts = []
    for it in range(7):
        dist.all_reduce(bar)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        capi.call("gnn_peer_gather_begin", ctx.h, arena, 0, rank * block, block)
        capi.call("gnn_peer_gather_wait", ctx.h, arena, 0, 1)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts[2:])[len(ts[2:]) // 2]
    t = torch.tensor([ms], device=ctx.device); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gb = (world - 1) * block / 1e9
    if rank == 0:
        print("peer gather world=%d F=%d copy=%s ctas=%s: %.3f ms, recv %.2f GB/rank -> %.1f GB/s per rank" %
              (world, F, os.environ.get("GNN_PEER_COPY", "sm"), os.environ.get("GNN_PEER_CTAS", "32"), float(t.item()), gb, gb / float(t.item()) * 1e3), flush=True)
    capi.call("gnn_peer_arena_destroy", ctx.h, arena)
ctx.close()
dist.destroy_process_group()

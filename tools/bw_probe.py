"""Multi-GPU link probe (not a test): NCCL all-gather bandwidth of one aggregation-sized block and, with --p2p,
copy-engine peer copy bandwidth inside one process.  torchrun --nproc-per-node N tools/bw_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist


def p2p():
    n = torch.cuda.device_count()
    a = torch.empty(256 << 20, dtype=torch.float32, device="cuda:0")
    for d in range(1, min(n, 3)):
        b = torch.empty_like(a, device="cuda:%d" % d)
        for _ in range(2):
            b.copy_(a)
        torch.cuda.synchronize(0); torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for _ in range(5):
            b.copy_(a)
        torch.cuda.synchronize(0); torch.cuda.synchronize(d)
        dt = (time.perf_counter() - t0) / 5
        print("p2p copy cuda:0 -> cuda:%d  %.1f GB/s" % (d, a.numel() * 4 / dt / 1e9), flush=True)
    if n >= 2:  # bidirectional
        b = torch.empty_like(a, device="cuda:1"); c = torch.empty_like(a, device="cuda:1"); d0 = torch.empty_like(a)
        s0 = torch.cuda.Stream(device=0); s1 = torch.cuda.Stream(device=1)
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        t0 = time.perf_counter()
        for _ in range(5):
            with torch.cuda.stream(s0):
                b.copy_(a, non_blocking=True)
            with torch.cuda.stream(s1):
                d0.copy_(c, non_blocking=True)
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        dt = (time.perf_counter() - t0) / 5
        print("p2p bidirectional per direction %.1f GB/s" % (a.numel() * 4 / dt / 1e9), flush=True)


def nccl():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    rows = (2450000 + world - 1) // world
    for F in (256, 128, 48):
        send = torch.randn(rows * F, device="cuda")
        recv = torch.empty(world * rows * F, device="cuda")
        for _ in range(3):
            dist.all_gather_into_tensor(recv, send)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            dist.all_gather_into_tensor(recv, send)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        gb = (world - 1) * rows * F * 4 / 1e9
        if rank == 0:
            print("nccl all-gather world=%d F=%d: %.3f ms, recv %.2f GB/rank -> %.1f GB/s per rank (env %s)" %
                  (world, F, ms, gb, gb / ms * 1e3, {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    if "--p2p" in sys.argv:
        p2p()
    else:
        nccl()

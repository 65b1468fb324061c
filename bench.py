#!/usr/bin/env python
"""bench.py — full-batch GCN train step (fwd + loss + bwd + SGD) on a synthetic BASELINE.json config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config products] [--impl ours|reference]

One JSON line on stdout (rank 0).  Metric: GCN train-step ms (BASELINE.json), lower is better, with the SpMM
roofline (algorithmic GB/s of all aggregation launches vs measured HBM peak), the CPU baseline timed on the same
box, and the end-to-end number through the host-buffer C ABI entry point.

N > 1 (launched by torchrun, one rank per GPU): nodes are 1-D row-partitioned; every aggregation all-gathers its
input rows over NCCL; gradients are all-reduced.  Total work is fixed => "scaling": "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import gnn_cpp_b200  # noqa: E402,F401
from gnn_cpp_b200 import synth  # noqa: E402

LR = 0.01
CPU_SAMPLE_DIV = {"products": 16, "reddit": 32, "arxiv": 1, "pubmed": 1, "cora": 1, "tiny_pl": 1}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


class NvlinkCounter:
    """NVLink payload bytes this GPU sent / received (NVML throughput counters, summed over its links), read before and
    after the timed region: the measured exchange volume next to the algorithmic one of the partition plan."""

    def __init__(self, device):
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.ids = [pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX]
        except Exception:  # noqa: BLE001
            self.h = None

    def _field(self, fid, scope):
        v = self.nv.nvmlDeviceGetFieldValues(self.h, [(fid, scope)])[0]
        if v.nvmlReturn != 0:
            return None
        t = v.valueType   # NVML_VALUE_TYPE_*: 0 double, 1 unsigned int, 2 unsigned long, 3 unsigned long long, 4 signed long long
        return int({0: v.value.dVal, 1: v.value.uiVal, 2: v.value.ulVal, 3: v.value.ullVal, 4: v.value.sllVal}.get(t, v.value.ullVal))

    def read(self):
        """(tx_bytes, rx_bytes) or None"""
        if self.h is None:
            return None
        try:
            vals = []
            for fid in self.ids:
                v = self._field(fid, 0xFFFFFFFF)            # scope UINT_MAX = all links of the device
                if v is None:                               # older drivers: per-link scopes only
                    per = [self._field(fid, l) for l in range(18)]
                    per = [x for x in per if x is not None]
                    if not per:
                        return None
                    v = sum(per)
                vals.append(v * 1024)                       # the counters are in KiB
            return tuple(vals)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            return None


# ------------------------------------------------------------------------------------------------ CPU arms
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _port_step_ms(cfg, p, steps, warmup):
    """`steps` timed train steps (after `warmup`) of the oracle port on problem p, all host threads."""
    from oracle import oracle as orc
    orc.set_threads(host_threads())      # explicit: torchrun exports OMP_NUM_THREADS=1, which must not throttle this arm
    G = orc.Graph(p.src, p.dst, cfg.N)
    W = [w.copy() for w in p.W]; b = [x.copy() for x in p.b]
    ts = []
    for i in range(warmup + max(1, steps)):
        t0 = time.perf_counter()
        orc.train_step(G, cfg.dims, p.X, p.y, W, b, lr=LR, order=0)
        if i >= warmup:
            ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.mean(ts)), G.nnz, orc.max_threads()


def cpu_step_ms(cfg_name, steps=1, warmup=0, budget_s=240.0, problem=None):
    """The reference's algorithm on the host cores, ON THE BENCH CONFIG ITSELF.
    cora: the real reference binary (oracle/_ref/ref_gcn: 1 thread, dense N x N, the only config it can run besides
    Pubmed-shaped).  Every other config: the oracle port (sparse restatement in the reference's operation order, fp32,
    OpenMP over all host threads) — the FULL workload; a 1/16-scale probe step predicts its cost first and only when
    the full step would not fit `budget_s` does the leg fall back to the scaled sample (labelled as such)."""
    cfg = synth.CONFIGS[cfg_name]
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "ref_gcn")
    if cfg_name == "cora" and os.path.exists(ref_bin):
        from gnn_cpp_b200 import problem_io
        p = problem if problem is not None else synth.make_problem(cfg)
        with tempfile.TemporaryDirectory() as td:
            pin = os.path.join(td, "p.gcnp")
            problem_io.write_problem(pin, p)
            out = subprocess.check_output([ref_bin, "time", pin, str(max(1, steps))], text=True)
        r = json.loads(out.strip().splitlines()[-1])
        return r["ms_per_step"], {"kind": "reference", "cores": 1, "same_config": True,
                                  "sample": "full Cora-shaped step through the patched reference binary (mode B, dense A_hat rebuilt per step), %d timed step(s)" % max(1, steps)}
    div = CPU_SAMPLE_DIV.get(cfg_name, 1)
    predicted = None
    if div > 1:
        scfg = synth.Config(cfg.name + "_s", max(cfg.N // div, 64), max(cfg.E // div, 2), cfg.dims, cfg.powerlaw, cfg.config_id)
        sms, snnz, cores = _port_step_ms(scfg, synth.make_problem(scfg), 1, 0)
        predicted = sms * div / 1e3
        if predicted * (warmup + max(1, steps)) + 30.0 > budget_s and predicted + 30.0 <= budget_s:
            steps, warmup = 1, 0           # one full step fits, the requested number does not
        if predicted * (warmup + max(1, steps)) + 30.0 > budget_s:
            return sms * div, {"kind": "port", "cores": cores, "same_config": False,
                               "sample": "FALLBACK (a full step is predicted at %.0f s, budget %.0f s): oracle port on a 1/%d-scale sample "
                                         "(N=%d, E=%d, nnz=%d, same dims/generator), time x%d" % (predicted, budget_s, div, scfg.N, scfg.E, snnz, div)}
    p = problem if problem is not None else synth.make_problem(cfg)
    ms, nnz, cores = _port_step_ms(cfg, p, steps, warmup)
    return ms, {"kind": "port", "cores": cores, "same_config": True,
                "sample": "oracle port (sparse CSR restatement, reference op order, fp32, OpenMP over %d threads), FULL %s-shaped workload "
                          "(N=%d, nnz=%d): %d timed step(s) after %d warm-up%s" % (cores, cfg.name, cfg.N, nnz, max(1, steps), warmup,
                          "" if predicted is None else "; a 1/%d-scale probe predicted %.1f s/step" % (div, predicted))}


def workload_desc(cfg, nnz=None):
    d = {"workload": "%s-shaped synthetic graph, %d-layer GCN full-batch train step (fwd+loss+bwd+SGD)" % (cfg.name, len(cfg.dims) - 1),
         "nodes": cfg.N, "edges": cfg.E, "dims": cfg.dims, "graph": "power-law (Chung-Lu)" if cfg.powerlaw else "uniform (Erdos-Renyi)",
         "lr": LR, "optimizer": "SGD"}
    if nnz is not None:
        d["nnz_with_self_loops"] = int(nnz)
    return d


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = synth.CONFIGS[args.config]
    # bounded: the real reference binary runs up to 20 steps of the Cora-shaped config (~2 s each); the sparse port runs
    # the FULL config, as many of the requested steps as fit the time budget (products-shaped: one ~1 min step)
    if args.config == "cora":
        ms, info = cpu_step_ms(args.config, steps=min(args.steps, 20), warmup=0)
    else:
        ms, info = cpu_step_ms(args.config, steps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1), budget_s=args.cpu_budget)
    cfgd = workload_desc(cfg)
    cfgd["parallelism"] = "host CPU, %d thread(s)" % info["cores"]
    line = {"impl": "reference", "metric": "gcn_train_step_ms", "value": ms, "unit": "ms", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfgd,
            "cpu_baseline": dict(info, value=ms, unit="ms"),
            "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from gnn_cpp_b200 import host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG")          # the image's default prints a banner on stdout; keep stdout = one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = synth.CONFIGS[args.config]
    L = len(cfg.dims) - 1
    t0 = time.time()
    p = synth.make_problem(cfg)
    if rank == 0:
        log("[bench] generated %s: N=%d E=%d in %.1fs" % (cfg.name, cfg.N, cfg.E, time.time() - t0))

    ctx = host.Context(local_rank)
    if args.spmm_variant:
        from gnn_cpp_b200 import capi
        capi.call("gnn_set_spmm_variant", ctx.h, args.spmm_variant)
    if world > 1:
        ctx.init_comm_from_torch()
    t0 = time.time()
    src_d = torch.from_numpy(p.src).to(ctx.device)
    dst_d = torch.from_numpy(p.dst).to(ctx.device)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    gfull = host.Graph.build(ctx, src_d, dst_d, cfg.N)
    ev1.record()
    torch.cuda.synchronize()
    build_ms = ev0.elapsed_time(ev1)
    nnz = gfull.nnz
    # second build: the stream-ordered pool now holds the temporaries, so this is the kernels' own time
    ev0.record()
    g2 = host.Graph.build(ctx, src_d, dst_d, cfg.N)
    ev1.record()
    torch.cuda.synchronize()
    build_warm_ms = ev0.elapsed_time(ev1)
    g2.close(); del g2
    del src_d, dst_d
    grid = None
    if world > 1:
        from gnn_cpp_b200 import dist_plan
        chunk = (cfg.N + world - 1) // world
        lo, hi = min(cfg.N, rank * chunk), min(cfg.N, (rank + 1) * chunk)
        # partition of the aggregation: 1-D rows ("row") or Pr x Pc row groups x feature-column groups (csrc/trainer_grid.cu)
        ge = os.environ.get("GNN_GRID", "auto")
        grid = dist_plan.choose_grid(cfg.N, world, p.src, p.dst) if ge == "auto" else (None if ge in ("", "row") else tuple(int(x) for x in ge.lower().split("x")))
        if grid is not None and os.environ.get("GNN_COMM") == "nccl":
            grid = None                     # the ncclAllGather ablation is a row-partition schedule
        if grid is None:
            g = gfull.slice_rows(lo, hi)
        else:
            (_, _), (glo, ghi) = dist_plan.grid_partition(cfg.N, world, grid[1], rank)
            g = gfull.slice_rows(glo, ghi)
        gfull.close()
        rows_alloc = chunk
    else:
        lo, hi, g, rows_alloc = 0, cfg.N, gfull, cfg.N
    torch.cuda.empty_cache()
    n_loc = hi - lo
    if rank == 0:
        log("[bench] structure build %.1f ms on device (nnz=%d, symmetric=%s), wall %.1fs" % (build_ms, nnz, gfull.symmetric if world == 1 else "n/a", time.time() - t0))

    ld0 = (cfg.dims[0] + 3) // 4 * 4
    Xbuf = torch.zeros((rows_alloc, ld0), dtype=torch.float32, device=ctx.device)
    X = Xbuf[:n_loc, :cfg.dims[0]]
    X.copy_(torch.from_numpy(p.X[lo:hi]))
    ybuf = torch.zeros(rows_alloc, dtype=torch.int32, device=ctx.device)
    ybuf[:n_loc].copy_(torch.from_numpy(p.y[lo:hi]))
    if grid is not None:
        try:
            model = host.GCN(ctx, g, cfg.dims, grid=grid, n_loc=n_loc)
        except Exception as e:  # noqa: BLE001  (CUDA IPC peer mapping unavailable on this box: every rank gets the same status)
            if rank == 0:
                log("[bench] 2-D partition unavailable (%s): falling back to the 1-D row partition" % e)
            g.close()
            gfull = host.Graph.build(ctx, torch.from_numpy(p.src).to(ctx.device), torch.from_numpy(p.dst).to(ctx.device), cfg.N)
            g = gfull.slice_rows(lo, hi)
            gfull.close()
            grid = None
    if grid is None:
        model = host.GCN(ctx, g, cfg.dims)
    model.set_option("precision", args.precision)
    model.set_params(p.W, p.b)
    loss_d = torch.zeros(1, dtype=torch.float32, device=ctx.device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=ctx.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    step = lambda: model.train_step(X, ybuf, LR, loss_d)  # noqa: E731
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    nvl = NvlinkCounter(local_rank) if world > 1 else None
    nv0 = nvl.read() if nvl else None
    total_ms = timed(step, args.steps)
    nv1 = nvl.read() if nvl else None
    launches = ctx.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    final_loss = float(loss_d.cpu()[0])

    # per-class device time (CUDA events around every launch group inside the trainer), separate pass
    model.set_option("profile", 1)
    bd_acc = None
    nprof = max(2, min(args.steps, 5))
    per_width = {}   # F -> [sum ms, sum algorithmic bytes, launches]: CUDA events around each aggregation launch
    for _ in range(nprof):
        step()
        bd = model.breakdown()
        bd_acc = bd if bd_acc is None else {k: bd_acc[k] + bd[k] for k in bd}
        for ms, by, F in model.spmm_spans():
            a = per_width.setdefault(F, [0.0, 0.0, 0])
            a[0] += ms; a[1] += by; a[2] += 1
    model.set_option("profile", 0)
    bd = {k: v / nprof for k, v in bd_acc.items()}
    st = model.stats()
    peak, peak_src = peaks()
    spmm_gbs = st["spmm_alg_bytes"] / (bd["spmm"] * 1e-3) / 1e9 if bd["spmm"] > 0 else 0.0
    # dominant kernel = the aggregation width that takes the most time in a step; its roofline numbers are per launch
    dom_F = max(per_width, key=lambda f: per_width[f][0])
    dom_ms, dom_bytes, dom_n = per_width[dom_F]
    dom_gbs = dom_bytes / (dom_ms * 1e-3) / 1e9
    # DRAM bytes per launch of the dominant kernel come from an `ncu --set full` capture (dram__bytes_read.sum +
    # dram__bytes_write.sum), not from this run: the value is static and labelled with the capture it was read from
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "spmm_traffic.json")
    if os.path.exists(tpath) and world == 1:
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj.get(args.config, {}).get(str(dom_F))
        if traffic is not None:
            traffic_src = "static: %s" % tj.get("_note", "ncu capture under profiles/")
    launches_detail = {str(f): {"launches_per_step": v[2] / nprof, "ms_per_launch": v[0] / v[2], "alg_bytes_per_launch": v[1] / v[2],
                                "alg_gbs": v[1] / (v[0] * 1e-3) / 1e9, "frac": v[1] / (v[0] * 1e-3) / 1e9 / peak}
                       for f, v in sorted(per_width.items())}

    # end to end through the host-buffer entry point: H2D of the step's inputs (pinned) + step + D2H of the loss
    Xh = torch.from_numpy(np.ascontiguousarray(p.X[lo:hi])).pin_memory()
    yh = torch.from_numpy(np.ascontiguousarray(p.y[lo:hi])).pin_memory()
    # Pipelined: every call runs the step on the batch uploaded during the previous call and starts the upload of
    # the batch it is handed (copy stream, double-buffered) — each step's inputs cross PCIe exactly once, inside
    # the timed region, overlapped with the previous step's kernels.  The loss comes back to the host every step.
    e2e_fn = lambda: model.train_step_host(Xh, yh, LR)  # noqa: E731
    e2e_fn()                       # unpipelined first call (also allocates the staging slots)
    model.prefetch_host(Xh, yh)    # prime the pipeline
    e2e_fn()
    e2e_steps = max(2, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_fn()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    model.train_step_host(None, None, LR)   # drain the last prefetched batch
    if world > 1:
        t = torch.tensor([e2e_ms], device=ctx.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())

    # ---- parity, visible to the driver at every GPU count: gradients of the FIRST train step (initial parameters,
    # lr = 0) against the CPU oracle's exact result for this config (tests/golden/bench_parity_<config>.npz, written by
    # tests/golden/make_bench_parity.py).  Gradients are all-reduced, so every rank holds the whole slab.
    parity = None
    gpath = os.path.join(ROOT, "tests", "golden", "bench_parity_%s.npz" % args.config)
    if os.path.exists(gpath):
        gold = np.load(gpath)
        model.set_params(p.W, p.b)
        # ReLU tie-breaks: hidden pre-activations the oracle has within 1e-5 max|Z| of zero take the oracle's side of the
        # discontinuous `Z > 0` (gnn_gcn_set_relu_overrides); without this, ~1e-7 of the entries land on the other side in
        # ANY other fp32 evaluation order and each one moves dW_1 by up to ~1e-4 of its (tiny) largest entry
        n_ties = 0
        for l in range(1, L):
            if "kink%d_rows" % l in gold.files:
                kr = gold["kink%d_rows" % l].astype(np.int64)
                sel = (kr >= lo) & (kr < hi)
                model.set_relu_overrides(l, (kr[sel] - lo).astype(np.int32), gold["kink%d_cols" % l][sel].astype(np.int32),
                                         gold["kink%d_pos" % l][sel])
                n_ties += int(len(kr))
        loss0 = float(model.train_step(X, ybuf, 0.0, loss_d).cpu()[0])
        per, checksum = {}, 0.0
        for l in range(1, L + 1):
            dW, db = model.grads(l)
            checksum += float(dW.astype(np.float64).sum()) + float(db.astype(np.float64).sum())
            for nm, a in (("dW%d" % l, dW), ("db%d" % l, db)):
                r = gold[nm].astype(np.float64)
                per[nm] = float(np.abs(a.astype(np.float64) - r).max() / max(np.abs(r).max(), 1e-30))
        rows = gold["logit_rows"]
        rows = rows[(rows >= lo) & (rows < hi)]
        if len(rows):
            Z = model.activation(L)[rows - lo].astype(np.float64)
            r = gold["logits"][np.searchsorted(gold["logit_rows"], rows)].astype(np.float64)
            per["logits_sample"] = float(np.abs(Z - r).max() / max(np.abs(gold["logits"]).max(), 1e-30))
        lref = float(gold["loss"][0])
        tol = 1e-5
        worst = max(list(per.values()) + [abs(loss0 - lref) / abs(lref)])
        parity = {"against": "CPU oracle, exact arithmetic (fp32 inputs, fp64 accumulation, reference op order): tests/golden/bench_parity_%s.npz" % args.config,
                  "what": "first train step from the initial parameters (lr = 0): loss, every dW/db (whole arrays, after the gradient all-reduce), %d sampled logits rows of rank 0" % len(rows),
                  "loss": loss0, "loss_oracle": lref, "loss_rel_err": abs(loss0 - lref) / abs(lref),
                  "rel_err": per, "max_rel_err": worst, "grad_checksum": checksum, "tol": tol, "norm": "max|a-ref|/max|ref|",
                  "relu_tie_breaks": n_ties, "hidden_entries": int(cfg.N * sum(cfg.dims[1:-1])),
                  "nnz_matches": int(gold["nnz"][0]) == int(nnz), "ok": bool(worst <= tol and int(gold["nnz"][0]) == int(nnz))}
        for l in range(1, L):
            model.set_relu_overrides(l, [], [], [])

    # dense transforms: roofline = max(HBM time of operands read once + outputs written once, 3 x TF32 MMA time at the
    # MEASURED dense TF32 peak of this GPU) summed over the step's GEMM launches
    import ctypes
    from gnn_cpp_b200 import capi as _capi
    tf32 = ctypes.c_double(0.0)
    _capi.call("gnn_tf32_peak_probe", ctx.h, ctypes.byref(tf32))
    gemm_roof = None
    if bd["gemm"] > 0:
        af = [cfg.dims[l - 1] < cfg.dims[l] for l in range(1, L + 1)]
        hbm_ms = mma_ms = floor_ms = 0.0
        for l in range(1, L + 1):
            fi, fo = cfg.dims[l - 1], cfg.dims[l]
            shapes = [(fi, fo, 0)]                                  # forward: read [n, fi], write [n, fo]
            shapes.append((fi + fo, 0, 0))                          # dW: read both operands
            if l > 1:
                shapes.append((fo, fi, fi))                         # dH: read dP, write dH, read the ReLU mask
            for rd, wr, mask in shapes:
                by = 4.0 * n_loc * (rd + wr + mask)
                fl = 2.0 * n_loc * fi * fo * 3                      # three TF32 MMAs per product
                h, c = by / (peak * 1e9) * 1e3, fl / (tf32.value * 1e12) * 1e3
                hbm_ms += h; mma_ms += c; floor_ms += max(h, c)
        gemm_roof = {"kernels": "tc_rows_kernel / tc_tn_kernel (tcgen05 3xTF32; bias gradients fused into the NN epilogues on 1 GPU), %d launches per step" % (3 * L - 1),
                     "ms": bd["gemm"], "floor_ms": floor_ms, "frac": floor_ms / bd["gemm"], "hbm_floor_ms": hbm_ms,
                     "mma_floor_ms": mma_ms, "tf32_peak_tflops_measured": tf32.value,
                     "how": "floor = sum over launches of max(operand+output bytes / HBM peak, 3 * 2MNK / measured TF32 peak)"}

    nvlink = None
    if world > 1:
        from gnn_cpp_b200 import dist_plan as _dp
        alg = (_dp.comm_bytes_per_step(cfg.N, cfg.dims, world) if grid is None
               else _dp.comm_bytes_per_step_grid(cfg.N, cfg.dims, world, grid[1]))
        nvlink = {"algorithmic_rx_bytes_per_step": int(alg),
                  "what": "bytes rank 0 receives per step by the partition plan (every needed row once; full halo) vs NVML NVLink payload counters of GPU 0 over the timed region"}
        if not (nv0 and nv1):
            nvlink["measured"] = "unavailable: NVML reports the NVLink throughput fields (NVML_FI_DEV_NVLINK_THROUGHPUT_*) as NOT_SUPPORTED on this box"
        if nv0 and nv1:
            nvlink.update({"measured_tx_bytes_per_step": (nv1[0] - nv0[0]) / args.steps, "measured_rx_bytes_per_step": (nv1[1] - nv0[1]) / args.steps,
                           "rx_gbs_over_step": (nv1[1] - nv0[1]) / args.steps / (ms_per_step * 1e-3) / 1e9})

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            cms, info = cpu_step_ms(args.config, steps=1, warmup=0, budget_s=args.cpu_budget, problem=p)
            cpu = dict(info, value=cms, unit="ms")
        cfgd = workload_desc(cfg, nnz)
        cfgd.update({"parallelism": ("%s x%d, %s + NCCL grad all-reduce" % ("1-D row partition" if grid is None else "activations 1-D row-partitioned, aggregation %d x %d" % grid, world, model.exchange_desc())) if world > 1 else "single GPU",
                     "l2_policy": "working set (feature matrices %.1f GB) larger than L2; no flush needed" % (cfg.N * max(cfg.dims) * 4 / 1e9)
                     if cfg.N * max(cfg.dims) * 4 > 2.5e8 else "small working set (L2-resident): launch-bound config",
                     "gemm_precision": "fp32 FMA" if args.precision == 0 else "3xTF32 tcgen05 (CTA pairs, cta_group::2, on the wide products; truncating hi/lo split of the streamed operand)",
                     "exchange": model.exchange_stats() if grid is not None else None, "spmm_launches_per_step": st["n_spmm"], "structure_build_ms": build_ms, "structure_build_warm_ms": build_warm_ms, "final_loss": final_loss})
        line = {"metric": "gcn_train_step_ms", "value": ms_per_step, "unit": "ms", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfgd,
                "roofline": {"bound": "hbm",
                             "kernel": "%s, F=%d aggregation (%.0f launches per step, %.0f%% of the step's aggregation time)"
                                       % ("spmm_merge_kernel (nonzero-balanced; chosen by degree skew)" if _capi.load().gnn_graph_spmm_variant(ctx.h, g.h, 0) == 2
                                          else "spmm_rows_kernel (one row per lane group; chosen by degree skew)", dom_F, dom_n / nprof,
                                          100.0 * dom_ms / nprof / bd["spmm"]),
                             "achieved": dom_gbs, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                             "frac": dom_gbs / peak, "traffic": traffic, "traffic_source": traffic_src,
                             "alg_bytes_per_launch": dom_bytes / dom_n, "ms_per_launch": dom_ms / dom_n,
                             # when frac > 1 on B_alg (L2 serves reuse) quote the touch-once bound too
                             "b_min_bytes_per_launch": 4.0 * (n_loc + 1) + 8.0 * nnz / world + 8.0 * n_loc * dom_F,
                             "all_aggregations_of_a_step": {"achieved": spmm_gbs, "frac": spmm_gbs / peak,
                                                            "alg_bytes": st["spmm_alg_bytes"], "ms": bd["spmm"],
                                                            "by_width": launches_detail}},
                "breakdown_ms": bd,
                "gemm_tflops": st["gemm_flops"] / (bd["gemm"] * 1e-3) / 1e12 if bd["gemm"] > 0 else None,
                "gemm_roofline": gemm_roof,
                "cpu_baseline": cpu, "parity": parity, "nvlink": nvlink,
                "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(Xh.numel() * 4 + yh.numel() * 4),
                        "d2h_bytes_per_step": 4, "api": "gnn_gcn_train_step_h (pinned host buffers)", "pipelined": True,
                        "pipelining": "the upload of step i+1's inputs (copy stream, double-buffered) overlaps step i's kernels; every step's inputs cross PCIe once inside the timed region and its loss is read back"},
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    model.close(); g.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="products", choices=sorted(synth.CONFIGS))
    ap.add_argument("--precision", type=int, default=1, help="dense transforms: 1 = 3xTF32 on tcgen05 (default), 0 = FP32 FMA")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=240.0, help="seconds the CPU leg may take before it falls back to a scaled sample")
    ap.add_argument("--spmm-variant", type=int, default=0, help="0 auto (by degree skew), 1 rows kernel, 2 merge kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

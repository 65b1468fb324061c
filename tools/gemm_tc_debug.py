"""Bring-up / diagnosis script for the tcgen05 3xTF32 GEMMs (run on a B200: python tools/gemm_tc_debug.py [case...]).
Not collected by pytest; the parity tests proper are in test_gpu_parity.py."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import gnn_cpp_b200  # noqa: E402,F401  (registers the package)
from gnn_cpp_b200 import host  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b)) / max(np.max(np.abs(b)), 1e-30))


def report(name, got, ref):
    e = rel(got, ref)
    bad = np.argwhere(np.abs(got.astype(np.float64) - ref) > 1e-4 * max(np.max(np.abs(ref)), 1e-30))
    print(f"{name}: rel_err={e:.3e} bad={len(bad)}/{ref.size}", flush=True)
    if len(bad):
        rows = np.unique(bad[:, 0]); cols = np.unique(bad[:, 1])
        print("   bad rows", rows[:16], "... n=", len(rows), " bad cols", cols[:16], "... n=", len(cols))
        for r, c in bad[:6]:
            print(f"   [{r},{c}] got={got[r, c]:.6f} ref={ref[r, c]:.6f}")
    return e


def main():
    ctx = host.Context(0)
    dev = ctx.device
    # (257, 200, 96): second 256-row tile whose odd CTA is entirely out of range, ragged output halves;
    # (70000, 64, 64): weights resident in shared memory under CTA pairs
    cases = [(128, 32, 32), (128, 16, 32), (256, 64, 64), (257, 200, 96), (1000, 256, 256), (4099, 256, 100),
             (3001, 48, 256), (3001, 47, 256), (2708, 16, 1433), (70000, 64, 64), (300000, 256, 256)]
    if os.environ.get("DEBUG_NO_TIMING"):
        cases = cases[:-1] + [(150000, 256, 256)]
    which = sys.argv[1:] or ["nt", "nn", "tn"]
    worst = 0.0
    for (M, N, K) in cases:
        rng = np.random.default_rng(M + N + K)
        ldk = (K + 3) // 4 * 4
        ldn = (N + 3) // 4 * 4
        A = np.zeros((M, ldk), np.float32); A[:, :K] = rng.uniform(-1, 1, (M, K))
        W = rng.uniform(-1, 1, (N, K)).astype(np.float32)
        dP = np.zeros((M, ldn), np.float32); dP[:, :N] = rng.uniform(-1, 1, (M, N))
        bias = rng.uniform(-1, 1, N).astype(np.float32)
        Ad = torch.from_numpy(A).to(dev); Wd = torch.from_numpy(W).to(dev); dPd = torch.from_numpy(dP).to(dev)
        A64 = A[:, :K].astype(np.float64); W64 = W.astype(np.float64); dP64 = dP[:, :N].astype(np.float64)
        if "nt" in which:
            out = torch.full((M, ldn), 7.0, device=dev)
            host.gemm_nt(ctx, Ad[:, :K], Wd, precision=1, out=out[:, :N])
            torch.cuda.synchronize()
            worst = max(worst, report(f"NT  M={M} N={N} K={K}", out.cpu().numpy()[:, :N], A64 @ W64.T))
            if ldn != N:
                assert float(out[:, N:].min()) == 7.0 and float(out[:, N:].max()) == 7.0, "pad columns were written"
            out2 = torch.empty((M, ldn), device=dev)
            host.gemm_nt(ctx, Ad[:, :K], Wd, bias=torch.from_numpy(bias).to(dev), relu=True, precision=1, out=out2[:, :N])
            torch.cuda.synchronize()
            worst = max(worst, report("    +bias+relu", out2.cpu().numpy()[:, :N], np.maximum(A64 @ W64.T + bias, 0)))
        if "nn" in which:
            out = torch.empty((M, ldk), device=dev)
            host.gemm_nn(ctx, dPd[:, :N], Wd, precision=1, out=out[:, :K])
            torch.cuda.synchronize()
            ref = dP64 @ W64
            worst = max(worst, report(f"NN  M={M} N={K} K={N}", out.cpu().numpy()[:, :K], ref))
            host.gemm_nn(ctx, dPd[:, :N], Wd, mask=Ad[:, :K], precision=1, out=out[:, :K])
            torch.cuda.synchronize()
            worst = max(worst, report("    +mask", out.cpu().numpy()[:, :K], np.where(A[:, :K] > 0, ref, 0)))
        if "tn" in which:
            got = host.gemm_tn(ctx, dPd[:, :N], Ad[:, :K], precision=1)
            torch.cuda.synchronize()
            worst = max(worst, report(f"TN  M={M} K1={N} K2={K}", got.cpu().numpy(), dP64.T @ A64))
            got2 = host.gemm_tn(ctx, dPd[:, :N], Ad[:, :K], precision=1)
            assert torch.equal(got, got2), "TN not deterministic"
    # timing at the products-shaped sizes
    M = 2449029
    for (N, K) in ([] if os.environ.get("DEBUG_NO_TIMING") else [(256, 256), (256, 100), (47, 256)]):
        ldk = (K + 3) // 4 * 4; ldn = (N + 3) // 4 * 4
        A = torch.rand((M, ldk), device=dev) - 0.5
        W = torch.rand((N, K), device=dev) - 0.5
        dP = torch.rand((M, ldn), device=dev) - 0.5
        out = torch.empty((M, ldn), device=dev); outk = torch.empty((M, ldk), device=dev)
        for prec in (1, 0):
            for name, fn in (("nt", lambda: host.gemm_nt(ctx, A[:, :K], W, precision=prec, out=out[:, :N])),
                             ("nn", lambda: host.gemm_nn(ctx, dP[:, :N], W, mask=A[:, :K], precision=prec, out=outk[:, :K])),
                             ("tn", lambda: host.gemm_tn(ctx, dP[:, :N], A[:, :K], precision=prec))):
                fn(); torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    fn()
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 3
                print(f"time {name} prec={prec} M={M} N={N} K={K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s",
                      flush=True)
    print("WORST", worst)
    ctx.close()
    return 0 if worst <= 2e-5 else 1  # bring-up gate (protocol errors are O(1)); the 1e-5 contract is enforced by tests/


if __name__ == "__main__":
    t0 = time.time()
    rc = main()
    print("elapsed", time.time() - t0)
    sys.exit(rc)

"""Dense-transform probe (not a test): times the eight GEMM shapes of one products-shaped train step through the C ABI.
python tools/gemm_probe.py   (env GNN_GEMM_BK, GNN_GEMM_NO_RESIDENT, GNN_GEMM_DEBUG select experiment variants)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnn_cpp_b200  # noqa: E402,F401
from gnn_cpp_b200 import capi, host  # noqa: E402

ctx = host.Context(0)
M = 2450000
dev = ctx.device


def t(fn, n=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def p(x):
    import ctypes as C
    return C.c_void_p(x.data_ptr()) if x is not None else None


A256 = torch.rand((M, 256), device=dev); B256 = torch.rand((M, 256), device=dev); C256 = torch.empty((M, 256), device=dev)
A100 = torch.rand((M, 100), device=dev); A48 = torch.rand((M, 48), device=dev); C48 = torch.empty((M, 48), device=dev)
shapes = [("NT1 K=100 N=256", "nt", A100, 100, 256, C256, None), ("NT2 K=256 N=256", "nt", A256, 256, 256, C256, None),
          ("NT3 K=256 N=47", "nt", A256, 256, 47, C48, None), ("NN3 K=47 N=256 mask", "nn", A48, 47, 256, C256, B256),
          ("NN2 K=256 N=256 mask", "nn", A256, 256, 256, C256, B256), ("NN2 K=256 N=256 nomask", "nn", A256, 256, 256, C256, None),
          ("TN3 K1=47 K2=256", "tn", A48, 47, 256, None, B256), ("TN2 K1=256 K2=256", "tn", A256, 256, 256, None, B256),
          ("TN1 K1=256 K2=100", "tn", A256, 256, 100, None, A100)]
tot = 0.0
for name, kind, A, K, N, Cm, extra in shapes:
    if kind == "nt":
        W = torch.rand((N, K), device=dev); bias = torch.rand(N, device=dev)
        fn = lambda: capi.call("gnn_gemm_nt", ctx.h, M, N, K, p(A), A.stride(0), p(W), K, p(Cm), Cm.stride(0), p(bias), 1, 1)
        gb = (M * K + M * N) * 4 / 1e9
    elif kind == "nn":
        W = torch.rand((K, N), device=dev)
        fn = lambda: capi.call("gnn_gemm_nn", ctx.h, M, N, K, p(A), A.stride(0), p(W), N, p(Cm), Cm.stride(0), p(extra), extra.stride(0) if extra is not None else 0, 1)
        gb = (M * K + M * N * (2 if extra is not None else 1)) * 4 / 1e9
    else:
        out = torch.empty((K, N), device=dev)
        fn = lambda: capi.call("gnn_gemm_tn", ctx.h, M, K, N, p(A), A.stride(0), p(extra), extra.stride(0), p(out), N, 1)
        gb = (M * K + M * N) * 4 / 1e9
    ms = t(fn)
    if "nomask" not in name:
        tot += ms
    print("%-26s %.3f ms  %.2f GB -> %.0f GB/s  %.0f TF/s(x1)" % (name, ms, gb, gb / ms * 1e3, 2.0 * M * K * N / ms / 1e9), flush=True)
print("total (step shapes) %.2f ms  env=%s" % (tot, {k: v for k, v in os.environ.items() if k.startswith("GNN_")}), flush=True)

"""ncu target (not a test): the three 256 x 256 dense transforms of the products-shaped step, twice each (1st = warm-up).
ncu --set full --import-source on -k regex:"tc_rows_kernel|tc_tn_kernel" -s 3 -c 3 -o gpurun_out/<name> python tools/gemm_ncu_target.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnn_cpp_b200  # noqa: E402,F401
from gnn_cpp_b200 import capi, host  # noqa: E402

ctx = host.Context(0)
M = int(os.environ.get("TARGET_M", 2450000))
dev = ctx.device
p = lambda x: C.c_void_p(x.data_ptr())
A = torch.rand((M, 256), device=dev) - 0.5
B = torch.rand((M, 256), device=dev) - 0.5
Cm = torch.empty((M, 256), device=dev)
W = torch.rand((256, 256), device=dev) - 0.5
out = torch.empty((256, 256), device=dev)
for _ in range(2):
    capi.call("gnn_gemm_nt", ctx.h, M, 256, 256, p(A), 256, p(W), 256, p(Cm), 256, None, 0, 1)
    capi.call("gnn_gemm_nn", ctx.h, M, 256, 256, p(A), 256, p(W), 256, p(Cm), 256, p(B), 256, 1)
    capi.call("gnn_gemm_tn", ctx.h, M, 256, 256, p(A), 256, p(B), 256, p(out), 256, 1)
torch.cuda.synchronize()
print("done")

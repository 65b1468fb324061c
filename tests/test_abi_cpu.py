"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/gnn_c.h
declares, the Python binding covers them all, and the product path fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gnn_c.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"GNN_API\s+[\w\s\*]+?\b(gnn_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    assert len(syms) >= 55
    for must in ["gnn_graph_build", "gnn_graph_build_csc", "gnn_graph_normalize", "gnn_spmm_fwd", "gnn_spmm_bwd",
                 "gnn_gemm_nt", "gnn_gemm_tn", "gnn_gemm_nn", "gnn_softmax_xent", "gnn_sgd_step",
                 "gnn_gcn_train_step", "gnn_gcn_train_step_h", "gnn_allgather_rows", "gnn_partition_ptr_h"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from gnn_cpp_b200 import capi
    assert os.path.exists(capi.LIB_PATH), "libgnn_b200.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(capi.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(capi.SIGNATURES) == declared_symbols()       # the binding covers exactly the header
    assert capi.load().gnn_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device the product path must raise, not fall back (and never touch oracle/)."""
    import torch
    from gnn_cpp_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = capi.load().gnn_ctx_create(0, None, ctypes.byref(h))
    assert rc != 0 and b"no CUDA device" in capi.load().gnn_last_error()
    from gnn_cpp_b200 import host
    with pytest.raises(capi.GnnError):
        host.Context(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gnn.cpp_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "liboracle" not in src and "from oracle" not in src and "import oracle" not in src, fn


def test_partition_ptr_matches_oracle(oracle):
    import numpy as np
    from gnn_cpp_b200 import capi
    for N, P in [(10, 1), (10, 3), (2449029, 8), (7, 8), (232965, 4)]:
        mine = np.empty(P + 1, np.int64)
        capi.call("gnn_partition_ptr_h", N, P, mine.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(mine, oracle.partition_ptr(N, P))

"""ctypes binding of the C ABI in include/gnn_c.h (libgnn_b200.so).

This is plumbing for the Python harness (tests, bench.py): PyTorch supplies device memory and streams, every
computation goes through the C ABI.  There is no fallback: a missing library or a failing call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GNN_LIB") or os.path.join(_HERE, "libgnn_b200.so")   # GNN_LIB: experiment builds (tools/)
_lib = None

vp, i32, i64, f32, f64, cp, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_char_p, C.c_size_t
pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); int-returning functions are status codes and get error checking
SIGNATURES = {
    "gnn_version": (C.c_int, []),
    "gnn_last_error": (cp, []),
    "gnn_ctx_create": (C.c_int, [C.c_int, vp, pp]),
    "gnn_ctx_destroy": (C.c_int, [vp]),
    "gnn_ctx_sync": (C.c_int, [vp]),
    "gnn_ctx_stream": (vp, [vp]),
    "gnn_ctx_sm_count": (C.c_int, [vp]),
    "gnn_ctx_launch_count": (i64, [vp]),
    "gnn_malloc": (C.c_int, [vp, pp, sz]),
    "gnn_free": (C.c_int, [vp, vp]),
    "gnn_memset": (C.c_int, [vp, vp, C.c_int, sz]),
    "gnn_memcpy_h2d": (C.c_int, [vp, vp, vp, sz]),
    "gnn_memcpy_d2h": (C.c_int, [vp, vp, vp, sz]),
    "gnn_memcpy_d2d": (C.c_int, [vp, vp, vp, sz]),
    "gnn_fill_f32": (C.c_int, [vp, vp, f32, i64]),
    "gnn_graph_build": (C.c_int, [vp, vp, vp, i64, i32, C.c_int, pp]),
    "gnn_graph_build_h": (C.c_int, [vp, vp, vp, i64, i32, C.c_int, pp]),
    "gnn_graph_build_weighted": (C.c_int, [vp, vp, vp, vp, i64, i32, C.c_int, pp]),
    "gnn_graph_export_weights_h": (C.c_int, [vp, vp, vp]),
    "gnn_graph_from_csr": (C.c_int, [vp, i32, i32, vp, vp, vp, pp]),
    "gnn_graph_build_csc": (C.c_int, [vp, vp]),
    "gnn_graph_normalize": (C.c_int, [vp, vp]),
    "gnn_graph_normalize_as_written": (C.c_int, [vp, vp, vp]),
    "gnn_graph_destroy": (C.c_int, [vp, vp]),
    "gnn_graph_nnz": (i64, [vp]),
    "gnn_graph_rows": (i32, [vp]),
    "gnn_graph_cols": (i32, [vp]),
    "gnn_graph_is_symmetric": (C.c_int, [vp]),
    "gnn_graph_export_h": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "gnn_graph_to_dense": (C.c_int, [vp, vp, C.c_int, vp, i64]),
    "gnn_dense_to_coo": (C.c_int, [vp, vp, i64, i64, i64, vp, vp, vp, i64, vp]),
    "gnn_spmm_fwd": (C.c_int, [vp, vp, vp, i64, i32, vp, i64, vp, C.c_int, vp, i64, C.c_int]),
    "gnn_spmm_bwd": (C.c_int, [vp, vp, vp, i64, i32, vp, i64, vp, i64, C.c_int]),
    "gnn_set_spmm_variant": (C.c_int, [vp, C.c_int]),
    "gnn_graph_spmm_variant": (C.c_int, [vp, vp, C.c_int]),
    "gnn_gemm_nt": (C.c_int, [vp, i64, i32, i32, vp, i64, vp, i64, vp, i64, vp, C.c_int, C.c_int]),
    "gnn_gemm_nn": (C.c_int, [vp, i64, i32, i32, vp, i64, vp, i64, vp, i64, vp, i64, C.c_int]),
    "gnn_gemm_tn": (C.c_int, [vp, i64, i32, i32, vp, i64, vp, i64, vp, i64, C.c_int]),
    "gnn_bias_relu_fwd": (C.c_int, [vp, i64, i32, vp, i64, vp, C.c_int, vp, i64]),
    "gnn_relu_bwd": (C.c_int, [vp, i64, i32, vp, i64, vp, i64, vp, i64]),
    "gnn_bias_grad": (C.c_int, [vp, i64, i32, vp, i64, vp]),
    "gnn_softmax_xent": (C.c_int, [vp, i64, i32, vp, i64, vp, i64, vp, vp, i64]),
    "gnn_sgd_step": (C.c_int, [vp, i64, vp, vp, vp, f32, f32, f32, f32, C.c_int, C.c_int]),
    "gnn_batchnorm_fwd": (C.c_int, [vp, i64, i32, vp, i64, vp, vp, f32, C.c_int, vp, i64, vp, vp]),
    "gnn_batchnorm_bwd": (C.c_int, [vp, i64, i32, vp, i64, vp, vp, vp, f32, vp, i64, vp, i64, vp, i64, vp, vp]),
    "gnn_layernorm_fwd": (C.c_int, [vp, i64, i32, vp, i64, vp, vp, f32, C.c_int, vp, i64, vp, vp]),
    "gnn_layernorm_bwd": (C.c_int, [vp, i64, i32, vp, i64, vp, vp, vp, vp, i64, vp, i64, vp, i64, vp, vp]),
    "gnn_tanh_fwd": (C.c_int, [vp, i64, vp, vp]),
    "gnn_tanh_bwd": (C.c_int, [vp, i64, vp, vp, vp]),
    "gnn_dropout": (C.c_int, [vp, i64, vp, f32, C.c_uint64, vp]),
    "gnn_adam_step": (C.c_int, [vp, i64, vp, vp, vp, vp, f32, f32, f32, f32, f32, i64]),
    "gnn_softmax_xent_masked": (C.c_int, [vp, i64, i32, vp, i64, vp, vp, i64, vp, vp, i64]),
    "gnn_argmax_correct": (C.c_int, [vp, i64, i32, vp, i64, vp, vp, vp]),
    "gnn_binary_f32": (C.c_int, [vp, C.c_int, i64, i64, vp, i64, i64, vp, i64, i64, vp]),
    "gnn_unary_f32": (C.c_int, [vp, C.c_int, i64, vp, vp]),
    "gnn_where_f32": (C.c_int, [vp, i64, vp, vp, vp, vp]),
    "gnn_sum_f32": (C.c_int, [vp, i64, i64, vp, C.c_int, vp]),
    "gnn_transpose_f32": (C.c_int, [vp, i64, i64, vp, vp]),
    "gnn_gather_cols_f32": (C.c_int, [vp, i64, i64, vp, vp, vp]),
    "gnn_gcn_create": (C.c_int, [vp, vp, i32, vp, pp]),
    "gnn_gcn_destroy": (C.c_int, [vp, vp]),
    "gnn_gcn_set_params_h": (C.c_int, [vp, vp, i32, vp, vp]),
    "gnn_gcn_get_params_h": (C.c_int, [vp, vp, i32, vp, vp]),
    "gnn_gcn_get_grads_h": (C.c_int, [vp, vp, i32, vp, vp]),
    "gnn_gcn_get_activation_h": (C.c_int, [vp, vp, i32, vp]),
    "gnn_gcn_get_dlogits_h": (C.c_int, [vp, vp, vp]),
    "gnn_gcn_set_option": (C.c_int, [vp, cp, f64]),
    "gnn_gcn_train_step": (C.c_int, [vp, vp, vp, i64, vp, f32, vp]),
    "gnn_gcn_forward": (C.c_int, [vp, vp, vp, i64]),
    "gnn_gcn_train_step_h": (C.c_int, [vp, vp, vp, vp, f32, vp]),
    "gnn_gcn_prefetch_h": (C.c_int, [vp, vp, vp, vp]),
    "gnn_gcn_set_train_mask": (C.c_int, [vp, vp, vp, i64]),
    "gnn_gcn_set_relu_overrides": (C.c_int, [vp, vp, i32, vp, vp, vp, i64]),
    "gnn_gcn_accuracy": (C.c_int, [vp, vp, vp, vp, vp]),
    "gnn_gcn_last_breakdown": (C.c_int, [vp, vp, C.c_int]),
    "gnn_gcn_last_spmm_spans": (C.c_int, [vp, vp, vp, vp, C.c_int, vp]),
    "gnn_gcn_spmm_stats": (C.c_int, [vp, vp, vp, vp]),
    "gnn_gcn_exchange_mode": (C.c_int, [vp]),
    "gnn_gcn_create_grid": (C.c_int, [vp, vp, i32, vp, i32, i32, pp]),
    "gnn_gcn_exchange_stats": (C.c_int, [vp, vp, vp, vp, vp]),
    "gnn_partition_col_slice_h": (C.c_int, [i32, i32, i32, vp, vp]),
    "gnn_partition_grid_h": (C.c_int, [i64, i32, i32, i32, vp, vp, vp, vp]),
    "gnn_tf32_peak_probe": (C.c_int, [vp, vp]),
    "gnn_partition_ptr_h": (C.c_int, [i64, i32, vp]),
    "gnn_partition_panels_h": (C.c_int, [i32, i32, vp, vp, vp]),
    "gnn_graph_slice_rows": (C.c_int, [vp, vp, i64, i64, pp]),
    "gnn_partition_build": (C.c_int, [vp, vp, i64, i64, C.c_int, pp]),
    "gnn_partition_halo_count": (i64, [vp]),
    "gnn_partition_interior_count": (i64, [vp]),
    "gnn_partition_nnz": (i64, [vp]),
    "gnn_partition_export_h": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "gnn_partition_destroy": (C.c_int, [vp, vp]),
    "gnn_comm_unique_id_h": (C.c_int, [vp]),
    "gnn_comm_init": (C.c_int, [vp, vp, C.c_int, C.c_int]),
    "gnn_comm_destroy": (C.c_int, [vp]),
    "gnn_allgather_rows": (C.c_int, [vp, vp, vp, i64, i32]),
    "gnn_allreduce_sum": (C.c_int, [vp, vp, i64]),
    "gnn_peer_arena_create": (C.c_int, [vp, sz, pp]),
    "gnn_peer_arena_destroy": (C.c_int, [vp, vp]),
    "gnn_peer_arena_local": (vp, [vp]),
    "gnn_peer_gather_begin": (C.c_int, [vp, vp, C.c_int, sz, sz]),
    "gnn_peer_gather_wait": (C.c_int, [vp, vp, C.c_int, C.c_int]),
}
# int-returning functions that are NOT status codes
_PLAIN_INT = {"gnn_version", "gnn_ctx_sm_count", "gnn_graph_is_symmetric", "gnn_gcn_exchange_mode", "gnn_graph_spmm_variant"}


class GnnError(RuntimeError):
    pass


def load():
    """Load libgnn_b200.so (built by gnn.cpp_b200/csrc/Makefile or __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GnnError("libgnn_b200.so is missing at %s — run `python -c 'import __graft_entry__ as g; g.build()'`; "
                       "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        if os.environ.get("GNN_LIB") and not hasattr(lib, name):
            continue             # an older experiment build (A/B runs under tools/) may lack newer entry points
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what=""):
    if status != 0:
        raise GnnError("%s failed (%d): %s" % (what, status, load().gnn_last_error().decode(errors="replace")))


def call(name, *args):
    """Call a status-returning entry point and raise GnnError on failure."""
    lib = load()
    r = getattr(lib, name)(*args)
    if SIGNATURES[name][0] is C.c_int and name not in _PLAIN_INT:
        check(r, name)
    return r

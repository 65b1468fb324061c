"""ctypes front end of the CPU oracle (oracle/gcn_oracle.c).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and bench.py's CPU legs.
The product package (gnn.cpp_b200/) must never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "gcn_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_max_threads.restype = C.c_int
        L.orc_hash3.restype = C.c_uint64
        L.orc_hash3.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_synth_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_float, C.c_float, _f32p]
        L.orc_synth_labels.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_int32, _i32p]
        L.orc_synth_edges.argtypes = [C.c_uint64, C.c_int64, C.c_int32, C.c_int, _i32p, _i32p]
        L.orc_csr_build.restype = C.c_int64
        L.orc_csr_build.argtypes = [_i32p, _i32p, C.c_int64, C.c_int32, C.c_int, _i64p, _i32p]
        L.orc_csc_from_csr.argtypes = [C.c_int32, _i64p, _i32p, _i64p, _i32p, _i64p]
        L.orc_degree_norm.argtypes = [C.c_int32, _i64p, _i32p, _i32p, _f32p, _f32p]
        L.orc_spmm.argtypes = [C.c_int32, _i64p, _i32p, _f32p, _f32p, C.c_int64, C.c_int32, _f32p, C.c_int64, C.c_int]
        for fn in (L.orc_gemm_nt, L.orc_gemm_nn):
            fn.argtypes = [C.c_int64, C.c_int32, C.c_int32, _f32p, C.c_int64, _f32p, C.c_int64, _f32p, C.c_int64, C.c_int]
        L.orc_gemm_tn.argtypes = [C.c_int64, C.c_int32, C.c_int32, _f32p, C.c_int64, _f32p, C.c_int64, _f32p, C.c_int64, C.c_int]
        L.orc_bias_relu.argtypes = [C.c_int64, C.c_int32, _f32p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        L.orc_relu_bwd.argtypes = [C.c_int64, C.c_int32, _f32p, C.c_int64, _f32p, C.c_int64, _f32p, C.c_int64]
        L.orc_bias_grad.argtypes = [C.c_int64, C.c_int32, _f32p, C.c_int64, _f32p, C.c_int]
        L.orc_softmax_xent.restype = C.c_float
        L.orc_softmax_xent.argtypes = [C.c_int64, C.c_int32, _f32p, C.c_int64, _i32p, C.c_void_p, C.c_int64, C.c_int]
        L.orc_sgd_step.argtypes = [C.c_int64, _f32p, _f32p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]
        L.orc_adam_step.argtypes = [C.c_int64, _f32p, _f32p, _f32p, _f32p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64]
        L.orc_softmax_xent_masked.restype = C.c_float
        L.orc_softmax_xent_masked.argtypes = [C.c_int64, C.c_int32, _f32p, C.c_int64, _i32p, _u8p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_argmax_correct.restype = C.c_int64
        L.orc_argmax_correct.argtypes = [C.c_int64, C.c_int32, _f32p, C.c_int64, _i32p, C.c_void_p]
        L.orc_batchnorm_fwd.argtypes = [C.c_int64, C.c_int32, _f32p, C.c_int64, _f32p, C.c_void_p, C.c_float, C.c_int, _f32p,
                                        C.c_int64, _f32p, _f32p, C.c_int]
        L.orc_batchnorm_bwd.argtypes = [C.c_int64, C.c_int32, _f32p, C.c_int64, _f32p, _f32p, _f32p, C.c_float, C.c_int,
                                        C.c_void_p, C.c_int64, _f32p, C.c_int64, _f32p, C.c_int64, _f32p, _f32p]
        L.orc_aswritten_norm.argtypes = [C.c_int32, _i64p, _i32p, _f32p, _f32p]
        L.orc_csr_build_weighted.restype = C.c_int64
        L.orc_csr_build_weighted.argtypes = [_i32p, _i32p, _f32p, C.c_int64, C.c_int32, C.c_int, _i64p, _i32p, _f32p]
        L.orc_degree_norm_weighted.argtypes = [C.c_int32, _i64p, _i32p, _f32p, _f32p, _f32p, _f32p]
        L.orc_layernorm_fwd.argtypes = [C.c_int64, C.c_int32, _f32p, C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_int, _f32p,
                                        C.c_int64, _f32p, _f32p, C.c_int]
        L.orc_layernorm_bwd.argtypes = [C.c_int64, C.c_int32, _f32p, C.c_int64, _f32p, _f32p, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_int64, _f32p, C.c_int64, _f32p, C.c_int64, _f32p, _f32p]
        L.orc_tanh_fwd.argtypes = [C.c_int64, _f32p, _f32p]
        L.orc_dropout_fwd.argtypes = [C.c_int64, _f32p, C.c_float, C.c_uint64, _f32p]
        L.orc_partition_ptr.argtypes = [C.c_int64, C.c_int32, _i64p]
        L.orc_partition_rows.restype = C.c_int64
        L.orc_partition_rows.argtypes = [_i64p, _i32p, C.c_void_p, C.c_int64, C.c_int64, _i64p, _i32p, C.c_void_p]
        L.orc_partition_interior.restype = C.c_int64
        L.orc_partition_interior.argtypes = [_i64p, _i32p, C.c_int64, C.c_int64, C.c_int64, _u8p]
        L.orc_partition_halo.restype = C.c_int64
        L.orc_partition_halo.argtypes = [_i64p, _i32p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _i32p, _i32p]
        L.orc_gcn_train_step.restype = C.c_float
        _LIB = L
    return _LIB


def max_threads():
    return lib().orc_max_threads()


def set_threads(n):
    lib().orc_set_threads(int(n))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- synthetic inputs (must equal gnn.cpp_b200/synth.py bit for bit) ---------------------------------
def synth_uniform(seed, stream, n, lo, hi):
    out = np.empty(n, dtype=np.float32)
    lib().orc_synth_uniform(seed, stream, n, lo, hi, out)
    return out


def synth_labels(seed, stream, n, Cn):
    out = np.empty(n, dtype=np.int32)
    lib().orc_synth_labels(seed, stream, n, Cn, out)
    return out


def synth_edges(seed, E, N, powerlaw):
    src = np.empty(E, dtype=np.int32)
    dst = np.empty(E, dtype=np.int32)
    lib().orc_synth_edges(seed, E, N, int(powerlaw), src, dst)
    return src, dst


# ---- structure -------------------------------------------------------------------------------------
def csr_build(src, dst, N, fill_mode=1):
    src = np.ascontiguousarray(src, dtype=np.int32)
    dst = np.ascontiguousarray(dst, dtype=np.int32)
    E = len(src)
    rowptr = np.zeros(N + 1, dtype=np.int64)
    colidx = np.empty(max(E + N, 1), dtype=np.int32)
    nnz = lib().orc_csr_build(src, dst, E, N, fill_mode, rowptr, colidx)
    return rowptr, colidx[:nnz].copy()


def csc_from_csr(N, rowptr, colidx):
    nnz = int(rowptr[N])
    colptr = np.zeros(N + 1, dtype=np.int64)
    rowidx = np.empty(max(nnz, 1), dtype=np.int32)
    perm = np.empty(max(nnz, 1), dtype=np.int64)
    lib().orc_csc_from_csr(N, rowptr, colidx, colptr, rowidx, perm)
    return colptr, rowidx[:nnz], perm[:nnz]


def degree_norm(N, rowptr, colidx):
    deg = np.empty(N, dtype=np.int32)
    dinv = np.empty(N, dtype=np.float32)
    val = np.empty(max(int(rowptr[N]), 1), dtype=np.float32)
    lib().orc_degree_norm(N, rowptr, colidx, deg, dinv, val)
    return deg, dinv, val[: int(rowptr[N])]


def edge_weights(E):
    """deterministic positive test weights, exact in fp32 on every platform: w_i = 0.25 + ((7919 i) mod 1024) / 512"""
    i = np.arange(E, dtype=np.int64)
    return (np.float32(0.25) + ((i * 7919) % 1024).astype(np.float32) / np.float32(512.0)).astype(np.float32)


def csr_build_weighted(src, dst, w, N, fill_mode=1):
    src = np.ascontiguousarray(src, dtype=np.int32); dst = np.ascontiguousarray(dst, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    E = len(src)
    rowptr = np.zeros(N + 1, dtype=np.int64)
    colidx = np.empty(max(E + N, 1), dtype=np.int32); val0 = np.empty(max(E + N, 1), dtype=np.float32)
    nnz = lib().orc_csr_build_weighted(src, dst, w, E, N, fill_mode, rowptr, colidx, val0)
    return rowptr, colidx[:nnz].copy(), val0[:nnz].copy()


def degree_norm_weighted(N, rowptr, colidx, val0):
    degf = np.empty(N, np.float32); dinv = np.empty(N, np.float32); val = np.empty(max(len(val0), 1), np.float32)
    lib().orc_degree_norm_weighted(N, rowptr, np.ascontiguousarray(colidx), np.ascontiguousarray(val0), degf, dinv, val)
    return degf, dinv, val[: len(val0)]


class Graph:
    """CSR/CSC/normalisation of A_hat = D^-1/2 (A0 + I) D^-1/2 built by the oracle (A0 weighted when w is given)."""

    def __init__(self, src, dst, N, w=None):
        self.N = N
        if w is not None:
            self.rowptr, self.colidx, self.val0 = csr_build_weighted(src, dst, w, N, 1)
            self.nnz = int(self.rowptr[N])
            self.degf, self.dinv, self.val = degree_norm_weighted(N, self.rowptr, self.colidx, self.val0)
            self.deg = np.diff(self.rowptr).astype(np.int32)
            self.colptr, self.rowidx, self.perm = csc_from_csr(N, self.rowptr, self.colidx)
            self.valT = self.val[self.perm].copy()
            return
        self.rowptr, self.colidx = csr_build(src, dst, N, 1)
        self.nnz = int(self.rowptr[N])
        self.deg, self.dinv, self.val = degree_norm(N, self.rowptr, self.colidx)
        self.colptr, self.rowidx, self.perm = csc_from_csr(N, self.rowptr, self.colidx)
        self.valT = self.val[self.perm].copy()


# ---- compute ---------------------------------------------------------------------------------------
def spmm(N, ptr, idx, val, P, order=0):
    P = np.ascontiguousarray(P, dtype=np.float32)
    F = P.shape[1]
    Y = np.empty((N, F), dtype=np.float32)
    lib().orc_spmm(N, ptr, np.ascontiguousarray(idx), np.ascontiguousarray(val), P, F, F, Y, F, order)
    return Y


def gemm_nt(A, B, order=0):
    A = np.ascontiguousarray(A, dtype=np.float32); B = np.ascontiguousarray(B, dtype=np.float32)
    M, K = A.shape; Nn = B.shape[0]
    Cm = np.empty((M, Nn), dtype=np.float32)
    lib().orc_gemm_nt(M, Nn, K, A, K, B, K, Cm, Nn, order)
    return Cm


def gemm_nn(A, B, order=0):
    A = np.ascontiguousarray(A, dtype=np.float32); B = np.ascontiguousarray(B, dtype=np.float32)
    M, K = A.shape; Nn = B.shape[1]
    Cm = np.empty((M, Nn), dtype=np.float32)
    lib().orc_gemm_nn(M, Nn, K, A, K, B, Nn, Cm, Nn, order)
    return Cm


def gemm_tn(A, B, order=0):
    A = np.ascontiguousarray(A, dtype=np.float32); B = np.ascontiguousarray(B, dtype=np.float32)
    M, K1 = A.shape; K2 = B.shape[1]
    Cm = np.empty((K1, K2), dtype=np.float32)
    lib().orc_gemm_tn(M, K1, K2, A, K1, B, K2, Cm, K2, order)
    return Cm


def bias_relu(Y, bias):
    Y = np.ascontiguousarray(Y, dtype=np.float32)
    N, F = Y.shape
    Z = np.empty_like(Y); H = np.empty_like(Y)
    lib().orc_bias_relu(N, F, Y, F, _ptr(bias), _ptr(Z), F, _ptr(H), F)
    return Z, H


def relu_bwd(dH, Z):
    dH = np.ascontiguousarray(dH, dtype=np.float32); Z = np.ascontiguousarray(Z, dtype=np.float32)
    N, F = Z.shape
    out = np.empty_like(Z)
    lib().orc_relu_bwd(N, F, dH, F, Z, F, out, F)
    return out


def bias_grad(dZ, order=0):
    dZ = np.ascontiguousarray(dZ, dtype=np.float32)
    N, F = dZ.shape
    db = np.empty(F, dtype=np.float32)
    lib().orc_bias_grad(N, F, dZ, F, db, order)
    return db


def softmax_xent(Z, y, order=0, want_grad=True):
    Z = np.ascontiguousarray(Z, dtype=np.float32)
    N, Cn = Z.shape
    dZ = np.empty_like(Z) if want_grad else None
    loss = lib().orc_softmax_xent(N, Cn, Z, Cn, np.ascontiguousarray(y, dtype=np.int32), _ptr(dZ), Cn, order)
    return float(loss), dZ


def sgd_step(p, g, vel=None, lr=0.01, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, first=True):
    lib().orc_sgd_step(p.size, p.reshape(-1), np.ascontiguousarray(g, dtype=np.float32).reshape(-1), _ptr(vel), lr,
                       momentum, dampening, weight_decay, int(nesterov), int(first))


def adam_step(p, g, m, v, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, step=1):
    lib().orc_adam_step(p.size, p.reshape(-1), np.ascontiguousarray(g, dtype=np.float32).reshape(-1), m.reshape(-1),
                        v.reshape(-1), lr, beta1, beta2, eps, weight_decay, int(step))


def softmax_xent_masked(Z, y, mask, want_grad=True):
    Z = np.ascontiguousarray(Z, dtype=np.float32)
    N, Cn = Z.shape
    dZ = np.empty_like(Z) if want_grad else None
    nsel = C.c_int64(0)
    loss = lib().orc_softmax_xent_masked(N, Cn, Z, Cn, np.ascontiguousarray(y, dtype=np.int32),
                                         np.ascontiguousarray(mask, dtype=np.uint8), _ptr(dZ), Cn, C.byref(nsel))
    return float(loss), dZ, int(nsel.value)


def argmax_correct(Z, y, mask=None):
    Z = np.ascontiguousarray(Z, dtype=np.float32)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    return int(lib().orc_argmax_correct(Z.shape[0], Z.shape[1], Z, Z.shape[1], np.ascontiguousarray(y, dtype=np.int32), _ptr(m)))


def batchnorm_fwd(X, gamma, beta, eps=1e-5, relu=False, order=0):
    X = np.ascontiguousarray(X, dtype=np.float32)
    N, F = X.shape
    Y = np.empty_like(X); mean = np.empty(F, np.float32); var = np.empty(F, np.float32)
    lib().orc_batchnorm_fwd(N, F, X, F, np.ascontiguousarray(gamma, dtype=np.float32), _ptr(beta), eps, int(relu), Y, F, mean,
                            var, order)
    return Y, mean, var


def batchnorm_bwd(X, mean, var, gamma, dY, eps=1e-5, relu_out=None):
    X = np.ascontiguousarray(X, dtype=np.float32); dY = np.ascontiguousarray(dY, dtype=np.float32)
    N, F = X.shape
    dX = np.empty_like(X); dg = np.empty(F, np.float32); db = np.empty(F, np.float32)
    lib().orc_batchnorm_bwd(N, F, X, F, mean, var, np.ascontiguousarray(gamma, dtype=np.float32), eps,
                            int(relu_out is not None), _ptr(relu_out), F, dY, F, dX, F, dg, db)
    return dX, dg, db


def layernorm_fwd(X, gamma=None, beta=None, eps=1e-5, relu=False, order=0):
    X = np.ascontiguousarray(X, dtype=np.float32)
    N, F = X.shape
    Y = np.empty_like(X); mean = np.empty(N, np.float32); rstd = np.empty(N, np.float32)
    lib().orc_layernorm_fwd(N, F, X, F, _ptr(gamma), _ptr(beta), eps, int(relu), Y, F, mean, rstd, order)
    return Y, mean, rstd


def layernorm_bwd(X, mean, rstd, gamma, dY, relu_out=None):
    X = np.ascontiguousarray(X, dtype=np.float32); dY = np.ascontiguousarray(dY, dtype=np.float32)
    N, F = X.shape
    dX = np.empty_like(X); dg = np.empty(F, np.float32); db = np.empty(F, np.float32)
    lib().orc_layernorm_bwd(N, F, X, F, mean, rstd, _ptr(gamma), int(relu_out is not None), _ptr(relu_out), F, dY, F, dX, F, dg, db)
    return dX, dg, db


def tanh_fwd(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().orc_tanh_fwd(x.size, x.reshape(-1), y.reshape(-1))
    return y


def dropout_fwd(x, p, seed):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().orc_dropout_fwd(x.size, x.reshape(-1), p, int(seed), y.reshape(-1))
    return y


def mlp_fwd(X, Ws, bs, gammas, betas, order=0):
    """nn::MLP::forward (reference include/nn.h:193-214) with dropout p = 0: per layer Linear, then LayerNorm + ReLU
    unless the layer's width equals the last width.  gammas/betas: per layer, None where the layer has no LayerNorm."""
    H = np.ascontiguousarray(X, dtype=np.float32)
    outs = []
    for W, b, g, be in zip(Ws, bs, gammas, betas):
        H = (gemm_nt(H, W, order) + np.asarray(b, np.float32)[None, :]).astype(np.float32)
        if g is not None:
            H, _, _ = layernorm_fwd(H, g, be, relu=True, order=order)
        outs.append(H)
    return outs


def gcnconv_as_written(src, dst, N, X, W, bias, gamma, beta, eps=1e-5, order=0):
    """graph::GCNConv::forward exactly as written (reference src/graph.cpp:170-212): loops removed, Linear (no bias) ->
    BatchNorm (training statistics) -> ReLU, then (A0 h) * norm + bias with norm = (A0 dinv) * dinv, deg = rowsum(A0)+1.
    Returns a dict with every intermediate (lin, bn, h, norm, Z) and the loop-free CSR."""
    rowptr, colidx = csr_build(src, dst, N, 0)
    lin = gemm_nt(X, W, order)
    bn, mean, var = batchnorm_fwd(lin, gamma, beta, eps, relu=False, order=order)
    h = np.maximum(bn, np.float32(0)).astype(np.float32)
    dinv = np.empty(N, np.float32); norm = np.empty(N, np.float32)
    lib().orc_aswritten_norm(N, rowptr, colidx, dinv, norm)
    ones = np.ones(max(len(colidx), 1), np.float32)
    agg = spmm(N, rowptr, colidx, ones, h, order)
    Z = (agg * norm[:, None] + np.asarray(bias, np.float32)[None, :]).astype(np.float32)
    return {"lin": lin, "bn": bn, "h": h, "mean": mean, "var": var, "dinv": dinv, "norm": norm, "Z": Z, "rowptr": rowptr,
            "colidx": colidx}


def partition_ptr(N, P):
    out = np.empty(P + 1, dtype=np.int64)
    lib().orc_partition_ptr(N, P, out)
    return out


def partition_rows(rowptr, colidx, val, lo, hi):
    n = int(rowptr[hi] - rowptr[lo])
    lr = np.empty(hi - lo + 1, dtype=np.int64)
    lc = np.empty(max(n, 1), dtype=np.int32)
    lv = np.empty(max(n, 1), dtype=np.float32) if val is not None else None
    lib().orc_partition_rows(rowptr, colidx, _ptr(val), lo, hi, lr, lc, _ptr(lv))
    return lr, lc[:n], (lv[:n] if lv is not None else None)


def partition_interior(l_rowptr, l_colidx, lo, hi):
    nrows = len(l_rowptr) - 1
    flags = np.empty(max(nrows, 1), dtype=np.uint8)
    cnt = lib().orc_partition_interior(l_rowptr, np.ascontiguousarray(l_colidx), nrows, lo, hi, flags)
    return flags[:nrows], int(cnt)


def partition_halo(l_rowptr, l_colidx, n_cols, lo, hi):
    """(halo_ids sorted unique remote columns, locally renumbered colidx) of a row block; see orc_partition_halo"""
    nrows = len(l_rowptr) - 1
    halo = np.empty(max(n_cols, 1), dtype=np.int32)
    local = np.empty(max(len(l_colidx), 1), dtype=np.int32)
    n = lib().orc_partition_halo(l_rowptr, np.ascontiguousarray(l_colidx), nrows, n_cols, lo, hi, halo, local)
    return halo[:n].copy(), local[:len(l_colidx)].copy()


def train_step(g, dims, X, y, W, b, lr=0.0, order=0):
    """Full fwd+loss+bwd(+SGD when lr>0, in place on W/b). Returns dict of Z1..ZL, dW*, db*, dZ, loss."""
    L = len(dims) - 1
    N = g.N
    X = np.ascontiguousarray(X, dtype=np.float32)
    Zs = [np.empty((N, dims[l + 1]), dtype=np.float32) for l in range(L)]
    dWs = [np.empty((dims[l + 1], dims[l]), dtype=np.float32) for l in range(L)]
    dbs = [np.empty(dims[l + 1], dtype=np.float32) for l in range(L)]
    dZL = np.empty((N, dims[L]), dtype=np.float32)
    PP = C.c_void_p * L

    def arr(lst):
        return PP(*[a.ctypes.data_as(C.c_void_p) for a in lst])

    for a in list(W) + list(b):
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    fn = lib().orc_gcn_train_step
    fn.argtypes = [C.c_int32, C.c_int32, _i32p, _i64p, _i32p, _f32p, _i64p, _i32p, _f32p, _f32p, _i32p, PP, PP, PP, PP,
                   PP, _f32p, C.c_float, C.c_int]
    loss = fn(N, L, np.asarray(dims, dtype=np.int32), g.rowptr, g.colidx, g.val, g.colptr,
              np.ascontiguousarray(g.rowidx), g.valT, X, np.ascontiguousarray(y, dtype=np.int32), arr(W), arr(b),
              arr(Zs), arr(dWs), arr(dbs), dZL, lr, order)
    out = {"loss": float(loss), "dZ": dZL}
    for l in range(L):
        out["Z%d" % (l + 1)] = Zs[l]
        out["dW%d" % (l + 1)] = dWs[l]
        out["db%d" % (l + 1)] = dbs[l]
    return out


def forward_composed(g, dims, X, W, b, order=0):
    """Forward pass from the primitives in the reference's own (transform-first) order: returns (Hs, Zs) with
    Hs[0] = X, Hs[l] = ReLU(Z_l) for hidden layers, Zs[l-1] = Z_l (pre-activations; logits for the last layer)."""
    L = len(dims) - 1
    Hs, Zs = [np.ascontiguousarray(X, dtype=np.float32)], []
    for l in range(L):
        P = gemm_nt(Hs[l], W[l], order)
        if l < L - 1:
            Z, Hn = bias_relu(spmm(g.N, g.rowptr, g.colidx, g.val, P, order), b[l])
        else:  # logits: no activation
            Y = spmm(g.N, g.rowptr, g.colidx, g.val, P, order)
            Z = np.empty_like(Y)
            lib().orc_bias_relu(Y.shape[0], Y.shape[1], Y, Y.shape[1], _ptr(np.ascontiguousarray(b[l], dtype=np.float32)),
                                _ptr(Z), Y.shape[1], None, 0)
            Hn = None
        del P
        Zs.append(Z)
        Hs.append(Hn)
    return Hs, Zs


def backward_composed(g, dims, Hs, Zs, y, W, order=0, masks=None):
    """Loss + backward over the activations of forward_composed.  `masks` (list of L-1 boolean [N, F_l] arrays,
    entries may be None) overrides which side of the ReLU kink the BACKWARD takes for hidden layer l: a checker for
    pre-activations that sit within the forward tolerance of zero, where `Z > 0` (operation.h:560) is not a
    continuous function of the inputs."""
    L = len(dims) - 1
    loss, dZ = softmax_xent(Zs[-1], y, order)
    out = {"loss": loss, "dZ": dZ}
    for l in range(L - 1, -1, -1):
        out["Z%d" % (l + 1)] = Zs[l]
        out["db%d" % (l + 1)] = bias_grad(dZ, order)
        dP = spmm(g.N, g.colptr, g.rowidx, g.valT, dZ, order)
        out["dW%d" % (l + 1)] = gemm_tn(dP, Hs[l], order)
        if l > 0:
            dH = gemm_nn(dP, W[l], order)
            if masks is not None and masks[l - 1] is not None:
                dZ = np.where(masks[l - 1], dH, np.float32(0)).astype(np.float32)
            else:
                dZ = relu_bwd(dH, Zs[l - 1])
            del dH
        del dP
    return out


def train_step_composed(g, dims, X, y, W, b, order=0, masks=None):
    """The same fwd+loss+bwd as orc_gcn_train_step, composed from the primitives above in the reference's own
    (transform-first) order; see backward_composed for `masks`."""
    Hs, Zs = forward_composed(g, dims, X, W, b, order)
    return backward_composed(g, dims, Hs, Zs, y, W, order, masks)

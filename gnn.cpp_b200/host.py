"""Thin Python host layer over the C ABI for the test/bench harness.

PyTorch provides device buffers (torch.Tensor.data_ptr()), the CUDA stream and torch.distributed; all
computation is done by libgnn_b200.so.  The user-facing API of this project is the C++ surface in
gnn.cpp_b200/host/ (cyg::tensor, graph::GCNConv, nn::...) — this module only lets Python drive the same
C ABI the C++ classes call.
"""
import ctypes as C

import numpy as np
import torch

from . import capi


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    if isinstance(t, np.ndarray):
        return t.ctypes.data_as(C.c_void_p)
    raise TypeError(type(t))


class Context:
    def __init__(self, device=0, use_torch_stream=True):
        if not torch.cuda.is_available():
            raise capi.GnnError("no CUDA device: the GCN hot path has no CPU fallback")
        torch.cuda.set_device(device)
        self.device = torch.device("cuda", device)
        # One stream for everything: torch ops (copies, allocations, events) and the library's kernels must be
        # ordered on the SAME stream.  torch's legacy default stream has handle 0, which the C ABI reads as
        # "create your own", so make a real stream current for torch and hand that one to the context.
        stream = None
        if use_torch_stream:
            if torch.cuda.current_stream(device).cuda_stream == 0:
                self.stream = torch.cuda.Stream(device)
                torch.cuda.set_stream(self.stream)
            else:
                self.stream = torch.cuda.current_stream(device)
            stream = self.stream.cuda_stream
        h = C.c_void_p()
        capi.call("gnn_ctx_create", device, C.c_void_p(stream) if stream else None, C.byref(h))
        self.h = h
        self.sm_count = capi.load().gnn_ctx_sm_count(h)

    def sync(self):
        capi.call("gnn_ctx_sync", self.h)

    @property
    def launches(self):
        return capi.load().gnn_ctx_launch_count(self.h)

    def close(self):
        if self.h:
            capi.call("gnn_ctx_destroy", self.h)
            self.h = None

    # ---- multi-GPU -------------------------------------------------------------------------------
    def init_comm_from_torch(self):
        """Bootstrap the library's NCCL communicator through an initialised torch.distributed group."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        ident = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            capi.call("gnn_comm_unique_id_h", _ptr(ident))
        t = torch.from_numpy(ident).to(self.device) if dist.get_backend() == "nccl" else torch.from_numpy(ident)
        dist.broadcast(t, 0)
        ident = t.cpu().numpy().copy()
        capi.call("gnn_comm_init", self.h, _ptr(ident), rank, world)
        return rank, world


class Graph:
    """Device CSR/CSC of A_hat = D^-1/2 (A0 + I) D^-1/2."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    @classmethod
    def build(cls, ctx, src, dst, N, fill_mode=1, csc=True, normalize=True, weights=None):
        """src/dst: int32 numpy (host) or torch cuda int32 tensors; weights: optional float32 edge weights (edge_attr)."""
        h = C.c_void_p()
        E = len(src)
        if weights is not None:
            to_dev = lambda a, dt: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(ctx.device)  # noqa: E731
            sd, dd, wd = to_dev(src, np.int32), to_dev(dst, np.int32), to_dev(weights, np.float32)
            capi.call("gnn_graph_build_weighted", ctx.h, _ptr(sd), _ptr(dd), _ptr(wd), E, N, fill_mode, C.byref(h))
            torch.cuda.current_stream().synchronize()   # the temporaries above may be released now
        elif isinstance(src, torch.Tensor):
            assert src.is_cuda and src.dtype == torch.int32 and dst.dtype == torch.int32
            capi.call("gnn_graph_build", ctx.h, _ptr(src), _ptr(dst), E, N, fill_mode, C.byref(h))
        else:
            src = np.ascontiguousarray(src, dtype=np.int32)
            dst = np.ascontiguousarray(dst, dtype=np.int32)
            capi.call("gnn_graph_build_h", ctx.h, _ptr(src), _ptr(dst), E, N, fill_mode, C.byref(h))
        g = cls(ctx, h)
        if csc:
            capi.call("gnn_graph_build_csc", ctx.h, h)
        if normalize:
            capi.call("gnn_graph_normalize", ctx.h, h)
        return g

    @property
    def nnz(self):
        return capi.load().gnn_graph_nnz(self.h)

    @property
    def n_rows(self):
        return capi.load().gnn_graph_rows(self.h)

    @property
    def n_cols(self):
        return capi.load().gnn_graph_cols(self.h)

    @property
    def symmetric(self):
        return bool(capi.load().gnn_graph_is_symmetric(self.h))

    def slice_rows(self, lo, hi):
        h = C.c_void_p()
        capi.call("gnn_graph_slice_rows", self.ctx.h, self.h, lo, hi, C.byref(h))
        return Graph(self.ctx, h)

    def export(self, csc=True, values=True):
        n, nnz, nc = self.n_rows, self.nnz, self.n_cols
        out = {"rowptr": np.empty(n + 1, np.int32), "colidx": np.empty(nnz, np.int32)}
        if values:
            out.update(val=np.empty(nnz, np.float32), deg=np.empty(n, np.int32), dinv=np.empty(n, np.float32))
        if csc:
            out.update(colptr=np.empty(nc + 1, np.int32), rowidx=np.empty(nnz, np.int32), perm=np.empty(nnz, np.int32))
            if values:
                out["valT"] = np.empty(nnz, np.float32)
        g = out.get
        capi.call("gnn_graph_export_h", self.ctx.h, self.h, _ptr(out["rowptr"]), _ptr(out["colidx"]), _ptr(g("val")),
                  _ptr(g("colptr")), _ptr(g("rowidx")), _ptr(g("perm")), _ptr(g("valT")), _ptr(g("deg")), _ptr(g("dinv")))
        return out

    def export_weights(self):
        out = np.empty(self.nnz, dtype=np.float32)
        capi.call("gnn_graph_export_weights_h", self.ctx.h, self.h, _ptr(out))
        return out

    def to_dense(self, weighted=True):
        out = torch.empty((self.n_rows, self.n_cols), dtype=torch.float32, device=self.ctx.device)
        capi.call("gnn_graph_to_dense", self.ctx.h, self.h, int(weighted), _ptr(out), self.n_cols)
        return out

    def normalize_as_written(self):
        """graph::GCNConv::forward's factorised normalisation (graph built with fill_mode=0); returns norm[N]."""
        norm = torch.empty(self.n_rows, dtype=torch.float32, device=self.ctx.device)
        capi.call("gnn_graph_normalize_as_written", self.ctx.h, self.h, _ptr(norm))
        return norm

    def spmm_fwd(self, P, bias=None, relu=False, mask=None, use_values=True, out=None):
        n, F = self.n_rows, P.shape[1]
        Y = out if out is not None else torch.empty((n, F), dtype=torch.float32, device=P.device)
        capi.call("gnn_spmm_fwd", self.ctx.h, self.h, _ptr(P), P.stride(0), F, _ptr(Y), Y.stride(0), _ptr(bias),
                  int(relu), _ptr(mask), mask.stride(0) if mask is not None else 0, int(use_values))
        return Y

    def spmm_bwd(self, dZ, mask=None, use_values=True, out=None, n_out=None):
        F = dZ.shape[1]
        n = n_out if n_out is not None else self.n_cols
        dP = out if out is not None else torch.empty((n, F), dtype=torch.float32, device=dZ.device)
        capi.call("gnn_spmm_bwd", self.ctx.h, self.h, _ptr(dZ), dZ.stride(0), F, _ptr(dP), dP.stride(0), _ptr(mask),
                  mask.stride(0) if mask is not None else 0, int(use_values))
        return dP

    def partition_arrays(self, lo, hi, transpose=False):
        """halo_ids / local_colidx / interior flags and row lists of this row block [lo, hi) (gnn_partition_build)"""
        ph = C.c_void_p()
        capi.call("gnn_partition_build", self.ctx.h, self.h, int(lo), int(hi), int(transpose), C.byref(ph))
        L = capi.load()
        n_halo, n_int, nnz = L.gnn_partition_halo_count(ph), L.gnn_partition_interior_count(ph), L.gnn_partition_nnz(ph)
        n = hi - lo
        out = {"halo_ids": np.empty(n_halo, np.int32), "local_colidx": np.empty(nnz, np.int32), "interior": np.empty(n, np.uint8),
               "interior_rows": np.empty(n_int, np.int32), "boundary_rows": np.empty(n - n_int, np.int32)}
        capi.call("gnn_partition_export_h", self.ctx.h, ph, _ptr(out["halo_ids"]), _ptr(out["local_colidx"]), _ptr(out["interior"]),
                  _ptr(out["interior_rows"]), _ptr(out["boundary_rows"]))
        capi.call("gnn_partition_destroy", self.ctx.h, ph)
        return out

    def close(self):
        if self.h:
            capi.call("gnn_graph_destroy", self.ctx.h, self.h)
            self.h = None


def gemm_nt(ctx, A, B, bias=None, relu=False, precision=0, out=None):
    M, K = A.shape
    N = B.shape[0]
    Cm = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=A.device)
    capi.call("gnn_gemm_nt", ctx.h, M, N, K, _ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(Cm), Cm.stride(0),
              _ptr(bias), int(relu), precision)
    return Cm


def gemm_nn(ctx, A, B, mask=None, precision=0, out=None):
    M, K = A.shape
    N = B.shape[1]
    Cm = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=A.device)
    capi.call("gnn_gemm_nn", ctx.h, M, N, K, _ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(Cm), Cm.stride(0),
              _ptr(mask), mask.stride(0) if mask is not None else 0, precision)
    return Cm


def gemm_tn(ctx, A, B, precision=0, out=None):
    M, K1 = A.shape
    K2 = B.shape[1]
    Cm = out if out is not None else torch.empty((K1, K2), dtype=torch.float32, device=A.device)
    capi.call("gnn_gemm_tn", ctx.h, M, K1, K2, _ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(Cm), Cm.stride(0),
              precision)
    return Cm


def bias_relu(ctx, Y, bias=None, relu=True):
    out = torch.empty_like(Y)
    capi.call("gnn_bias_relu_fwd", ctx.h, Y.shape[0], Y.shape[1], _ptr(Y), Y.stride(0), _ptr(bias), int(relu), _ptr(out),
              out.stride(0))
    return out


def relu_bwd(ctx, dH, act):
    out = torch.empty_like(dH)
    capi.call("gnn_relu_bwd", ctx.h, dH.shape[0], dH.shape[1], _ptr(dH), dH.stride(0), _ptr(act), act.stride(0),
              _ptr(out), out.stride(0))
    return out


def bias_grad(ctx, dZ):
    db = torch.empty(dZ.shape[1], dtype=torch.float32, device=dZ.device)
    capi.call("gnn_bias_grad", ctx.h, dZ.shape[0], dZ.shape[1], _ptr(dZ), dZ.stride(0), _ptr(db))
    return db


def softmax_xent(ctx, Z, y, n_total=0, want_grad=True):
    loss = torch.empty(1, dtype=torch.float32, device=Z.device)
    dZ = torch.empty_like(Z) if want_grad else None
    capi.call("gnn_softmax_xent", ctx.h, Z.shape[0], Z.shape[1], _ptr(Z), Z.stride(0), _ptr(y), n_total, _ptr(loss),
              _ptr(dZ), dZ.stride(0) if want_grad else 0)
    return loss, dZ


def sgd_step(ctx, p, g, vel=None, lr=0.01, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, first=True):
    capi.call("gnn_sgd_step", ctx.h, p.numel(), _ptr(p), _ptr(g), _ptr(vel), lr, momentum, dampening, weight_decay,
              int(nesterov), int(first))


def batchnorm_fwd(ctx, X, gamma, beta=None, eps=1e-5, relu=False):
    N, F = X.shape
    Y = torch.empty((N, F), dtype=torch.float32, device=X.device)
    mean = torch.empty(F, dtype=torch.float32, device=X.device); var = torch.empty(F, dtype=torch.float32, device=X.device)
    capi.call("gnn_batchnorm_fwd", ctx.h, N, F, _ptr(X), X.stride(0), _ptr(gamma), _ptr(beta), eps, int(relu), _ptr(Y), F,
              _ptr(mean), _ptr(var))
    return Y, mean, var


def batchnorm_bwd(ctx, X, mean, var, gamma, dY, eps=1e-5, relu_out=None):
    N, F = X.shape
    dX = torch.empty((N, F), dtype=torch.float32, device=X.device)
    dg = torch.empty(F, dtype=torch.float32, device=X.device); db = torch.empty(F, dtype=torch.float32, device=X.device)
    capi.call("gnn_batchnorm_bwd", ctx.h, N, F, _ptr(X), X.stride(0), _ptr(mean), _ptr(var), _ptr(gamma), eps, _ptr(relu_out),
              relu_out.stride(0) if relu_out is not None else 0, _ptr(dY), dY.stride(0), _ptr(dX), F, _ptr(dg), _ptr(db))
    return dX, dg, db


def layernorm_fwd(ctx, X, gamma=None, beta=None, eps=1e-5, relu=False):
    N, F = X.shape
    Y = torch.empty((N, F), dtype=torch.float32, device=X.device)
    mean = torch.empty(N, dtype=torch.float32, device=X.device); rstd = torch.empty(N, dtype=torch.float32, device=X.device)
    capi.call("gnn_layernorm_fwd", ctx.h, N, F, _ptr(X), X.stride(0), _ptr(gamma), _ptr(beta), eps, int(relu), _ptr(Y), F,
              _ptr(mean), _ptr(rstd))
    return Y, mean, rstd


def layernorm_bwd(ctx, X, mean, rstd, gamma, dY, relu_out=None):
    N, F = X.shape
    dX = torch.empty((N, F), dtype=torch.float32, device=X.device)
    dg = torch.empty(F, dtype=torch.float32, device=X.device); db = torch.empty(F, dtype=torch.float32, device=X.device)
    capi.call("gnn_layernorm_bwd", ctx.h, N, F, _ptr(X), X.stride(0), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(relu_out),
              relu_out.stride(0) if relu_out is not None else 0, _ptr(dY), dY.stride(0), _ptr(dX), F, _ptr(dg), _ptr(db))
    return dX, dg, db


def tanh_fwd(ctx, x):
    y = torch.empty_like(x)
    capi.call("gnn_tanh_fwd", ctx.h, x.numel(), _ptr(x), _ptr(y))
    return y


def tanh_bwd(ctx, y, dy):
    dx = torch.empty_like(y)
    capi.call("gnn_tanh_bwd", ctx.h, y.numel(), _ptr(y), _ptr(dy), _ptr(dx))
    return dx


def dropout(ctx, x, p, seed):
    y = torch.empty_like(x)
    capi.call("gnn_dropout", ctx.h, x.numel(), _ptr(x), float(p), int(seed), _ptr(y))
    return y


def adam_step(ctx, p, g, m, v, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, step=1):
    capi.call("gnn_adam_step", ctx.h, p.numel(), _ptr(p), _ptr(g), _ptr(m), _ptr(v), lr, beta1, beta2, eps, weight_decay, int(step))


def softmax_xent_masked(ctx, Z, y, mask, n_selected=None):
    """mask: torch uint8/bool [N] on the device; returns (loss tensor[1], dZ)."""
    N, Cn = Z.shape
    mask = mask.to(torch.uint8)
    loss = torch.zeros(1, dtype=torch.float32, device=Z.device)
    dZ = torch.empty((N, Cn), dtype=torch.float32, device=Z.device)
    n_sel = int(mask.sum().item()) if n_selected is None else int(n_selected)
    capi.call("gnn_softmax_xent_masked", ctx.h, N, Cn, _ptr(Z), Z.stride(0), _ptr(y), _ptr(mask), n_sel, _ptr(loss), _ptr(dZ), Cn)
    return loss, dZ


def argmax_correct(ctx, Z, y, mask=None):
    cnt = torch.zeros(1, dtype=torch.int64, device=Z.device)
    m8 = None if mask is None else mask.to(torch.uint8)
    capi.call("gnn_argmax_correct", ctx.h, Z.shape[0], Z.shape[1], _ptr(Z), Z.stride(0), _ptr(y), _ptr(m8), _ptr(cnt))
    return int(cnt.item())


class GCN:
    """Fused trainer (gnn_gcn_* entry points)."""

    def set_train_mask(self, mask, n_selected_total=None):
        """mask: torch uint8/bool [local rows] on the device, or None for all nodes."""
        if mask is None:
            self._mask = None
            capi.call("gnn_gcn_set_train_mask", self.ctx.h, self.h, None, 0)
            return
        self._mask = mask.to(torch.uint8).contiguous()   # keep alive: the library stores the pointer
        n = int(self._mask.sum().item()) if n_selected_total is None else int(n_selected_total)
        capi.call("gnn_gcn_set_train_mask", self.ctx.h, self.h, _ptr(self._mask), n)

    def set_relu_overrides(self, layer, rows, cols, positive):
        """rows/cols: int32 numpy (LOCAL row ids, column ids), positive: bool/uint8 numpy; see gnn_gcn_set_relu_overrides."""
        if not hasattr(self, "_ov"):
            self._ov = {}
        n = len(rows)
        if n == 0:
            self._ov.pop(layer, None)
            capi.call("gnn_gcn_set_relu_overrides", self.ctx.h, self.h, layer, None, None, None, 0)
            return
        dev = self.ctx.device
        t = (torch.from_numpy(np.ascontiguousarray(rows, dtype=np.int32)).to(dev),
             torch.from_numpy(np.ascontiguousarray(cols, dtype=np.int32)).to(dev),
             torch.from_numpy(np.ascontiguousarray(positive, dtype=np.uint8)).to(dev))
        self._ov[layer] = t        # keep alive: the library stores the pointers
        capi.call("gnn_gcn_set_relu_overrides", self.ctx.h, self.h, layer, _ptr(t[0]), _ptr(t[1]), _ptr(t[2]), n)

    def accuracy(self, y, mask=None):
        cnt = torch.zeros(1, dtype=torch.int64, device=y.device)
        m8 = None if mask is None else mask.to(torch.uint8).contiguous()
        capi.call("gnn_gcn_accuracy", self.ctx.h, self.h, _ptr(y), _ptr(m8), _ptr(cnt))
        return int(cnt.item())

    def __init__(self, ctx, graph, dims, grid=None, n_loc=None):
        """grid = (Pr, Pc): 2-D partition (gnn_gcn_create_grid); `graph` is then the row group's structure slice and
        n_loc the number of activation rows this rank owns."""
        self.ctx, self.graph, self.dims = ctx, graph, list(dims)
        self.L = len(dims) - 1
        h = C.c_void_p()
        d = np.asarray(dims, dtype=np.int32)
        if grid is not None:
            capi.call("gnn_gcn_create_grid", ctx.h, graph.h, self.L, _ptr(d), int(grid[0]), int(grid[1]), C.byref(h))
            self.n_loc = int(n_loc)
        else:
            capi.call("gnn_gcn_create", ctx.h, graph.h, self.L, _ptr(d), C.byref(h))
            self.n_loc = graph.n_rows
        self.h = h

    def set_option(self, key, value):
        capi.call("gnn_gcn_set_option", self.h, key.encode(), float(value))

    def set_params(self, Ws, bs):
        for l, (W, b) in enumerate(zip(Ws, bs), start=1):
            W = np.ascontiguousarray(W, dtype=np.float32)
            b = np.ascontiguousarray(b, dtype=np.float32)
            capi.call("gnn_gcn_set_params_h", self.ctx.h, self.h, l, _ptr(W), _ptr(b))

    def params(self, l):
        W = np.empty((self.dims[l], self.dims[l - 1]), np.float32)
        b = np.empty(self.dims[l], np.float32)
        capi.call("gnn_gcn_get_params_h", self.ctx.h, self.h, l, _ptr(W), _ptr(b))
        return W, b

    def grads(self, l):
        W = np.empty((self.dims[l], self.dims[l - 1]), np.float32)
        b = np.empty(self.dims[l], np.float32)
        capi.call("gnn_gcn_get_grads_h", self.ctx.h, self.h, l, _ptr(W), _ptr(b))
        return W, b

    def activation(self, l):
        out = np.empty((self.n_loc, self.dims[l]), np.float32)
        capi.call("gnn_gcn_get_activation_h", self.ctx.h, self.h, l, _ptr(out))
        return out

    def dlogits(self):
        out = np.empty((self.n_loc, self.dims[-1]), np.float32)
        capi.call("gnn_gcn_get_dlogits_h", self.ctx.h, self.h, _ptr(out))
        return out

    def train_step(self, X, y, lr, loss_out=None):
        """X: cuda float32 [n_loc, F0] (row stride = ld), y: cuda int32 [n_loc]. Returns the device loss tensor."""
        if loss_out is None:
            loss_out = torch.empty(1, dtype=torch.float32, device=X.device)
        capi.call("gnn_gcn_train_step", self.ctx.h, self.h, _ptr(X), X.stride(0), _ptr(y), lr, _ptr(loss_out))
        return loss_out

    def forward(self, X):
        capi.call("gnn_gcn_forward", self.ctx.h, self.h, _ptr(X), X.stride(0))

    def train_step_host(self, X_h, y_h, lr):
        """End-to-end step from host buffers (numpy or pinned torch CPU tensors); returns the loss as float."""
        loss = np.zeros(1, np.float32)
        capi.call("gnn_gcn_train_step_h", self.ctx.h, self.h, _ptr(X_h) if X_h is not None else None,
                  _ptr(y_h) if y_h is not None else None, lr, _ptr(loss))
        return float(loss[0])

    def prefetch_host(self, X_h, y_h):
        """Start uploading the next step's host inputs (overlaps the running step); see gnn_gcn_prefetch_h."""
        capi.call("gnn_gcn_prefetch_h", self.ctx.h, self.h, _ptr(X_h), _ptr(y_h))

    def breakdown(self):
        ms = np.zeros(6, np.float64)
        capi.call("gnn_gcn_last_breakdown", self.h, _ptr(ms), 6)
        return dict(zip(["spmm", "gemm", "loss", "bias_grad", "sgd", "other"], ms.tolist()))

    def spmm_spans(self, cap=256):
        """[(ms, algorithmic bytes, F)] of every aggregation launch of the last profiled step."""
        ms = np.zeros(cap, np.float64); by = np.zeros(cap, np.float64); F = np.zeros(cap, np.int32); n = C.c_int(0)
        capi.call("gnn_gcn_last_spmm_spans", self.h, _ptr(ms), _ptr(by), _ptr(F), cap, C.byref(n))
        k = min(n.value, cap)
        return [(float(ms[i]), float(by[i]), int(F[i])) for i in range(k)]

    def stats(self):
        b, n, f = C.c_double(), C.c_int32(), C.c_double()
        capi.call("gnn_gcn_spmm_stats", self.h, C.byref(b), C.byref(n), C.byref(f))
        return {"spmm_alg_bytes": b.value, "n_spmm": n.value, "gemm_flops": f.value}

    def exchange_mode(self):
        return capi.load().gnn_gcn_exchange_mode(self.h)

    def exchange_stats(self):
        hf, idf = C.c_double(1.0), C.c_double(0.0)
        hl, sp = C.c_int(0), C.c_int(0)
        capi.call("gnn_gcn_exchange_stats", self.h, C.byref(hf), C.byref(hl), C.byref(sp), C.byref(idf))
        return {"halo_fraction": hf.value, "halo_only_exchange": bool(hl.value), "interior_boundary_split": bool(sp.value),
                "interior_fraction": idf.value}

    def exchange_desc(self):
        return {0: "no exchange (single GPU)",
                1: "ncclAllGather of every aggregation input",
                2: "all-gather of every aggregation input by SM store pushes into IPC-mapped peer arenas over NVLink, pipelined by column panels",
                3: "all-gather of every aggregation input by copy-engine pushes into IPC-mapped peer arenas, pipelined by column panels",
                4: "all-gather of every aggregation input by in-place ncclAllGather per column panel (side stream)",
                6: "2-D partition of the aggregation (row groups x feature-column groups): column-slice scatter into IPC-mapped peer arenas before, row exchange fused into the SpMM epilogue stores",
                }.get(self.exchange_mode(), "unknown")

    def close(self):
        if self.h:
            capi.call("gnn_gcn_destroy", self.ctx.h, self.h)
            self.h = None

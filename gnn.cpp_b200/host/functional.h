// functional.h — the stateless arithmetic layer (counterpart of reference include/functional.h:17-495).
//
// In the reference every function here is a std::valarray expression on the CPU; this is exactly where its
// author meant to call a device backend (`// device::add(out_data, lhs_data, rhs_data);`, functional.h:174,180).
// Each function below keeps the reference name and meaning and enqueues a kernel through the C ABI instead.
// All templates are instantiated for T = float (device compute type); shapes are viewed as [rows, cols].
#ifndef GNNB200_FUNCTIONAL_H
#define GNNB200_FUNCTIONAL_H

#include <memory>
#include <tuple>
#include <vector>

#include "device.h"
#include "utils.h"

namespace cyg {
template <class T> class tensor;
}

namespace functional {
using cyg::device::check;
using cyg::device::ctx;

struct detail {
    // fresh float tensor of a given shape on the device
    template <class T> static std::shared_ptr<cyg::tensor<T>> make(const cyg::dims_t &dims, bool rg) {
        return std::make_shared<cyg::tensor<T>>(dims, cyg::device::alloc(cyg::count_elements(dims) * 4), rg);
    }
    static void strides(const cyg::dims_t &in, size_t R, size_t C, int64_t &rs, int64_t &cs) {
        size_t r, c;
        cyg::as_2d(in, r, c);
        rs = (r == 1 && R > 1) ? 0 : (int64_t)c;
        cs = (c == 1 && C > 1) ? 0 : 1;
        if (r == 1 && R == 1) rs = (int64_t)c;
    }
};

template <class T>
std::shared_ptr<cyg::tensor<T>> binary(int op, const cyg::tensor<T> &lhs, const cyg::tensor<T> &rhs, bool rg) {
    const cyg::dims_t out_dims = cyg::broadcast_shape(lhs.shape(), rhs.shape());
    size_t R, C;
    cyg::as_2d(out_dims, R, C);
    int64_t ars, acs, brs, bcs;
    detail::strides(lhs.shape(), R, C, ars, acs);
    detail::strides(rhs.shape(), R, C, brs, bcs);
    auto out = detail::make<T>(out_dims, rg);
    check(gnn_binary_f32(ctx(), op, (int64_t)R, (int64_t)C, lhs.dptr(), ars, acs, rhs.dptr(), brs, bcs, out->dptr()));
    return out;
}
/** elementwise with numpy broadcasting — reference functional.h:162-264 */
template <class T> std::shared_ptr<cyg::tensor<T>> add(const cyg::tensor<T> &lhs, const cyg::tensor<T> &rhs) {
    return binary(GNN_OP_ADD, lhs, rhs, lhs.requires_grad() || rhs.requires_grad());
}
template <class T> std::shared_ptr<cyg::tensor<T>> mul(const cyg::tensor<T> &lhs, const cyg::tensor<T> &rhs) {
    return binary(GNN_OP_MUL, lhs, rhs, lhs.requires_grad() || rhs.requires_grad());
}
template <class T> std::shared_ptr<cyg::tensor<T>> div(const cyg::tensor<T> &num, const cyg::tensor<T> &den) {
    return binary(GNN_OP_DIV, num, den, num.requires_grad() || den.requires_grad());
}
template <class T> std::shared_ptr<cyg::tensor<T>> pow(const cyg::tensor<T> &base, const cyg::tensor<T> &exponent) {
    return binary(GNN_OP_POW, base, exponent, base.requires_grad() || exponent.requires_grad());
}
/** t1 > t2 as a bool tensor (stored as 0/1 floats on the device) — reference functional.h:143-160 */
template <class T> std::shared_ptr<cyg::tensor<bool>> gt(const cyg::tensor<T> &t1, const cyg::tensor<T> &t2) {
    const cyg::dims_t out_dims = cyg::broadcast_shape(t1.shape(), t2.shape());
    size_t R, C;
    cyg::as_2d(out_dims, R, C);
    int64_t ars, acs, brs, bcs;
    detail::strides(t1.shape(), R, C, ars, acs);
    detail::strides(t2.shape(), R, C, brs, bcs);
    auto out = std::make_shared<cyg::tensor<bool>>(out_dims, cyg::device::alloc(R * C * 4), false);
    check(gnn_binary_f32(ctx(), GNN_OP_GT, (int64_t)R, (int64_t)C, t1.dptr(), ars, acs, t2.dptr(), brs, bcs, out->dptr()));
    return out;
}

/** sum along a dim / over everything — reference functional.h:266-296 */
template <class T> std::shared_ptr<cyg::tensor<T>> sum(const cyg::tensor<T> &base, int dim = INT_MAX, const bool &keepdim = false) {
    const int rank = base.rank();
    if (dim != INT_MAX && dim < 0) dim += rank;
    size_t R, C;
    cyg::as_2d(base.shape(), R, C);
    cyg::dims_t out_dims;
    int kdim; // kernel dim on the [R,C] view: 0 = over rows, 1 = over cols, -1 = all
    if (dim == INT_MAX) { out_dims = {1}; kdim = -1; }
    else if (rank == 1) { out_dims = {1}; kdim = -1; }
    else if (dim == 0) { out_dims = keepdim ? cyg::dims_t{1, C} : cyg::dims_t{C}; kdim = 0; }
    else { out_dims = keepdim ? cyg::dims_t{R, 1} : cyg::dims_t{R}; kdim = 1; }
    auto out = detail::make<T>(out_dims, base.requires_grad());
    check(gnn_sum_f32(ctx(), (int64_t)R, (int64_t)C, base.dptr(), kdim, out->dptr()));
    return out;
}
template <class T> std::shared_ptr<cyg::tensor<T>> mean(const cyg::tensor<T> &base, int dim = INT_MAX, const bool &keepdim = false) {
    auto s = sum(base, dim, keepdim);
    const size_t n = dim == INT_MAX ? base.numel() : base.shape()[dim < 0 ? dim + base.rank() : dim];
    cyg::tensor<T> denom(cyg::dims_t{1}, static_cast<T>(n), false);
    return div(*s, denom);
}
/** reference functional.h:309-329 */
template <class T> std::shared_ptr<cyg::tensor<T>> exp(const cyg::tensor<T> &base) {
    auto out = detail::make<T>(base.shape(), base.requires_grad());
    check(gnn_unary_f32(ctx(), GNN_UOP_EXP, (int64_t)base.numel(), base.dptr(), out->dptr()));
    return out;
}
template <class T> std::shared_ptr<cyg::tensor<T>> log(const cyg::tensor<T> &base) {
    auto out = detail::make<T>(base.shape(), base.requires_grad());
    check(gnn_unary_f32(ctx(), GNN_UOP_LOG, (int64_t)base.numel(), base.dptr(), out->dptr()));
    return out;
}
/** 2-D transpose — reference functional.h:330-357 */
template <class T> std::shared_ptr<cyg::tensor<T>> transpose(const cyg::tensor<T> &t, int, int) {
    if (t.rank() != 2) throw std::runtime_error(cyg::err::transpose());
    auto out = detail::make<T>({t.shape()[1], t.shape()[0]}, t.requires_grad());
    check(gnn_transpose_f32(ctx(), (int64_t)t.shape()[0], (int64_t)t.shape()[1], t.dptr(), out->dptr()));
    return out;
}
/** dense product [M,K] x [K,N] — reference functional.h:398-441 (FP32 FMA GEMM kernel; precision 1 = 3xTF32) */
template <class T> std::shared_ptr<cyg::tensor<T>> matmul(const cyg::tensor<T> &lhs, const cyg::tensor<T> &rhs, int precision = 0) {
    if (lhs.rank() != 2 || rhs.rank() != 2 || lhs.shape()[1] != rhs.shape()[0]) throw std::runtime_error(cyg::err::mm_compatible());
    const int64_t M = lhs.shape()[0], K = lhs.shape()[1], N = rhs.shape()[1];
    auto out = detail::make<T>({(size_t)M, (size_t)N}, lhs.requires_grad() || rhs.requires_grad());
    check(gnn_gemm_nn(ctx(), M, (int32_t)N, (int32_t)K, lhs.dptr(), K, rhs.dptr(), N, out->dptr(), N, nullptr, 0, precision));
    return out;
}
/** out = cond > 0 ? true_value : false_value — reference functional.h:443-471 */
template <class T>
std::shared_ptr<cyg::tensor<T>> mask(const cyg::tensor<T> &cond, const cyg::tensor<T> &tv, const cyg::tensor<T> &fv) {
    auto out = detail::make<T>(tv.shape(), tv.requires_grad() || fv.requires_grad());
    check(gnn_where_f32(ctx(), (int64_t)tv.numel(), cond.dptr(), tv.dptr(), fv.dptr(), out->dptr()));
    return out;
}
/** out[i] = t[i, idx[i]] — reference functional.h:482-494 */
template <class T, class B> std::shared_ptr<cyg::tensor<T>> slice(const cyg::tensor<T> &t, const cyg::tensor<B> &idx, int) {
    auto out = detail::make<T>({t.shape()[0]}, t.requires_grad());
    check(gnn_gather_cols_f32(ctx(), (int64_t)t.shape()[0], (int64_t)t.shape()[1], t.dptr(), idx.dptr(), out->dptr()));
    return out;
}
/** max / argmax (reference functional.h:24-71): a read-out op — evaluated on the host mirror (explicit I/O),
 * never part of the training arithmetic. */
template <class T>
std::tuple<std::shared_ptr<cyg::tensor<T>>, std::shared_ptr<cyg::tensor<int>>> max(const cyg::tensor<T> &t, int dim = INT_MAX, const bool &keepdim = false) {
    auto *h = t.data();
    size_t R, C;
    cyg::as_2d(t.shape(), R, C);
    if (dim != INT_MAX && dim < 0) dim += t.rank();
    if (dim == INT_MAX || t.rank() == 1) {
        size_t best = 0;
        for (size_t i = 1; i < h->size(); i++) if ((*h)[i] > (*h)[best]) best = i;
        return {std::make_shared<cyg::tensor<T>>(cyg::dims_t{1}, (*h)[best], false),
                std::make_shared<cyg::tensor<int>>(cyg::dims_t{1}, (int)best, false)};
    }
    const size_t n_out = dim == 0 ? C : R, n_in = dim == 0 ? R : C;
    auto *vals = new std::valarray<T>(n_out);
    auto *idxs = new std::valarray<int>(n_out);
    for (size_t o = 0; o < n_out; o++) {
        size_t best = 0;
        for (size_t i = 1; i < n_in; i++) {
            const T a = dim == 0 ? (*h)[i * C + o] : (*h)[o * C + i];
            const T b = dim == 0 ? (*h)[best * C + o] : (*h)[o * C + best];
            if (a > b) best = i;
        }
        (*vals)[o] = dim == 0 ? (*h)[best * C + o] : (*h)[o * C + best];
        (*idxs)[o] = (int)best;
    }
    cyg::dims_t od = dim == 0 ? (keepdim ? cyg::dims_t{1, C} : cyg::dims_t{C}) : (keepdim ? cyg::dims_t{R, 1} : cyg::dims_t{R});
    return {std::make_shared<cyg::tensor<T>>(od, vals, false), std::make_shared<cyg::tensor<int>>(od, idxs, false)};
}
} // namespace functional
#endif

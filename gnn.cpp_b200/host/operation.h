// operation.h — autograd nodes (counterpart of reference include/operation.h:19-616).
//
// Same contract as the reference: an Operation owns a Context of saved inputs, forward() returns the output
// tensor, and the virtual _backward(incoming) hands a local gradient to each input with `input->backward(g)`.
// Differences, all deliberate:
//   * the arithmetic is device kernels (functional:: -> C ABI), never std::valarray;
//   * `input->backward(g)` called from inside a node is DEFERRED to a reverse-topological engine
//     (cyg::autograd), which sums the contributions of every consumer before running the producer's node —
//     the reference's depth-first recursion returns early on the second visit and loses them (bug B2);
//   * three fused nodes the hot path needs are added: SpMM (K4/K5 + bias/ReLU epilogue), LinearOp (NT GEMM +
//     bias, TN/NN GEMMs backward) and SoftmaxCrossEntropy (K8, analytic gradient — the reference's composed
//     loss backward throws, bug B3).
#ifndef GNNB200_OPERATION_H
#define GNNB200_OPERATION_H

#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "functional.h"
#include "utils.h"

namespace cyg {

/** saved inputs + integer attributes of a node — reference operation.h:19-61 */
template <class T> class Context {
    std::vector<std::shared_ptr<T>> cache;

  public:
    std::map<std::string, int> saved_data;
    void save_for_backward(std::vector<std::shared_ptr<T>> ts) {
        for (const auto &t : ts) cache.push_back(t);
    }
    std::vector<std::shared_ptr<T>> get_variables() { return cache; }
};

/** base node — reference operation.h:63-100 */
template <class T> class Operation {
  protected:
    std::unique_ptr<Context<T>> context;

  public:
    bool _done = false;
    std::string name;
    Operation() : context(std::make_unique<Context<T>>()) {}
    virtual ~Operation() = default;
    void reset() { context = std::make_unique<Context<T>>(); }
    std::vector<std::shared_ptr<T>> inputs() { return context->get_variables(); }
    void backward(std::shared_ptr<T> incoming_grad) {
        if (_done) { // same warning as the reference (operation.h:82-86)
            std::cout << "trying to backprop on this node again, pls be sure this is intended" << "\n";
            return;
        }
        _backward(incoming_grad);
    }
    virtual void _backward(std::shared_ptr<T>) {}
    friend std::ostream &operator<<(std::ostream &out, const Operation &op) { return out << op.name << "Op"; }
};

template <class T> void check_backward(const std::vector<std::shared_ptr<T>> &var, size_t expected) {
    if (var.size() != expected) throw std::runtime_error("cant backprop without executing a forward computation first");
}

// ---------------------------------------------------------------------------------------------------------
// reverse-topological engine
// ---------------------------------------------------------------------------------------------------------
namespace autograd {
using ftensor = tensor<float>;
struct Engine {
    std::unordered_map<ftensor *, std::shared_ptr<ftensor>> pending;
    void add(ftensor *t, const std::shared_ptr<ftensor> &g) {
        auto it = pending.find(t);
        if (it == pending.end()) pending.emplace(t, g);
        else it->second = functional::add(*it->second, *g);
    }
};
inline Engine *&current() {
    static thread_local Engine *e = nullptr;
    return e;
}
inline void topo(ftensor *t, std::unordered_set<ftensor *> &seen, std::vector<ftensor *> &order);
void run_or_defer(ftensor *root, std::shared_ptr<ftensor> seed);
// reduce a gradient of shape `g` to the shape of a broadcast input (the reference's sum_to_size, tensor.h:618-638)
std::shared_ptr<ftensor> sum_to_shape(const std::shared_ptr<ftensor> &g, const dims_t &target);
} // namespace autograd

// ---------------------------------------------------------------------------------------------------------
// elementwise nodes
// ---------------------------------------------------------------------------------------------------------
template <class T> class Add : public Operation<T> { // reference operation.h:102-129
  public:
    Add() { this->name = "Add"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &lhs, const std::shared_ptr<T> &rhs) {
        auto out = functional::add(*lhs, *rhs);
        if (out->requires_grad()) this->context->save_for_backward({lhs, rhs});
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        check_backward(var, 2);
        for (const auto &t : var)
            if (t->requires_grad()) t->backward(autograd::sum_to_shape(g, t->shape()));
        this->_done = true;
    }
};

template <class T> class Mul : public Operation<T> { // reference operation.h:131-168
  public:
    Mul() { this->name = "Mul"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &lhs, const std::shared_ptr<T> &rhs) {
        auto out = functional::mul(*lhs, *rhs);
        if (out->requires_grad()) this->context->save_for_backward({lhs, rhs});
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        check_backward(var, 2);
        auto lhs = var[0], rhs = var[1];
        if (rhs->requires_grad()) rhs->backward(autograd::sum_to_shape(functional::binary(GNN_OP_MUL, *g, *lhs, false), rhs->shape()));
        if (lhs->requires_grad()) lhs->backward(autograd::sum_to_shape(functional::binary(GNN_OP_MUL, *g, *rhs, false), lhs->shape()));
        this->_done = true;
    }
};

template <class T> class Div : public Operation<T> { // reference operation.h:169-208
  public:
    Div() { this->name = "Div"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &num, const std::shared_ptr<T> &den) {
        auto out = functional::div(*num, *den);
        if (out->requires_grad()) this->context->save_for_backward({num, den});
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        check_backward(var, 2);
        auto num = var[0], den = var[1];
        if (num->requires_grad()) num->backward(autograd::sum_to_shape(functional::binary(GNN_OP_DIV, *g, *den, false), num->shape()));
        if (den->requires_grad()) { // d(a/b)/db = -a / b^2
            auto b2 = functional::binary(GNN_OP_MUL, *den, *den, false);
            auto q = functional::binary(GNN_OP_DIV, *num, *b2, false);
            T minus_one(dims_t{1}, -1.0f, false);
            auto local = functional::binary(GNN_OP_MUL, *functional::binary(GNN_OP_MUL, *g, *q, false), minus_one, false);
            den->backward(autograd::sum_to_shape(local, den->shape()));
        }
        this->_done = true;
    }
};

template <class T> class Exp : public Operation<T> { // reference operation.h:338-367
  public:
    Exp() { this->name = "Exp"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &base) {
        auto out = functional::exp(*base);
        if (out->requires_grad()) this->context->save_for_backward({base});
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        check_backward(var, 1);
        if (var[0]->requires_grad()) {
            auto e = functional::exp(*var[0]);
            var[0]->backward(functional::binary(GNN_OP_MUL, *g, *e, false));
        }
        this->_done = true;
    }
};

template <class T> class Log : public Operation<T> { // reference operation.h:368-396
  public:
    Log() { this->name = "Log"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &base) {
        auto out = functional::log(*base);
        if (out->requires_grad()) this->context->save_for_backward({base});
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        check_backward(var, 1);
        if (var[0]->requires_grad()) var[0]->backward(functional::binary(GNN_OP_DIV, *g, *var[0], false));
        this->_done = true;
    }
};

template <class T> class Sum : public Operation<T> { // reference operation.h:255-292
  public:
    Sum() { this->name = "Sum"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &base, int dim = INT_MAX, const bool &keepdim = false) {
        auto out = functional::sum(*base, dim, keepdim);
        if (out->requires_grad()) {
            this->context->save_for_backward({base});
            this->context->saved_data["dim"] = dim == INT_MAX ? INT_MAX : (dim < 0 ? base->rank() + dim : dim);
            this->context->saved_data["keepdim"] = keepdim;
        }
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        check_backward(var, 1);
        auto base = var[0];
        if (base->requires_grad()) { // every element of the summed fibre receives the fibre's gradient
            size_t R, C;
            as_2d(base->shape(), R, C);
            const int dim = this->context->saved_data["dim"];
            // view g as [1,C] (dim 0), [R,1] (dim 1) or a scalar, and broadcast against ones
            auto gv = std::make_shared<T>(dim == INT_MAX || base->rank() == 1 ? dims_t{1} : (dim == 0 ? dims_t{1, C} : dims_t{R, 1}),
                                          g->buffer(), false);
            T ones(base->shape(), 1.0f, false);
            base->backward(functional::binary(GNN_OP_MUL, ones, *gv, false));
        }
        this->_done = true;
    }
};

template <class T> class Transpose : public Operation<T> { // reference operation.h:398-434
  public:
    Transpose() { this->name = "Transpose"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &lhs, int d1 = -1, int d2 = -2) {
        auto out = functional::transpose(*lhs, d1, d2);
        if (out->requires_grad()) this->context->save_for_backward({lhs});
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        check_backward(var, 1);
        if (var[0]->requires_grad()) {
            auto gt = functional::transpose(*g, -1, -2);
            var[0]->backward(gt);
        }
        this->_done = true;
    }
};

template <class T> class MatMul : public Operation<T> { // reference operation.h:489-535
  public:
    MatMul() { this->name = "MatMul"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &lhs, const std::shared_ptr<T> &rhs) {
        auto out = functional::matmul(*lhs, *rhs);
        if (out->requires_grad()) this->context->save_for_backward({lhs, rhs});
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        check_backward(var, 2);
        auto lhs = var[0], rhs = var[1]; // lhs [M,K], rhs [K,N], g [M,N]
        const int64_t M = lhs->shape()[0], K = lhs->shape()[1], N = rhs->shape()[1];
        if (lhs->requires_grad()) { // dL = g * rhs^T : NT product, no transposed copy (the reference clones + transposes)
            auto out = functional::detail::make<float>({(size_t)M, (size_t)K}, false);
            device::check(gnn_gemm_nt(device::ctx(), M, (int32_t)K, (int32_t)N, g->dptr(), N, rhs->dptr(), N, out->dptr(), K, nullptr, 0, 0));
            lhs->backward(out);
        }
        if (rhs->requires_grad()) { // dR = lhs^T * g : TN product with a fixed-order split reduction over M
            auto out = functional::detail::make<float>({(size_t)K, (size_t)N}, false);
            device::check(gnn_gemm_tn(device::ctx(), M, (int32_t)K, (int32_t)N, lhs->dptr(), K, g->dptr(), N, out->dptr(), N, 0));
            rhs->backward(out);
        }
        this->_done = true;
    }
};

template <class T> class Mask : public Operation<T> { // reference operation.h:537-573
  public:
    Mask() { this->name = "Mask"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &cond, const std::shared_ptr<T> &tv, const std::shared_ptr<T> &fv) {
        auto out = functional::mask(*cond, *tv, *fv);
        if (out->requires_grad()) this->context->save_for_backward({tv, fv, cond});
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        check_backward(var, 3);
        auto tv = var[0], fv = var[1], cond = var[2];
        if (tv->requires_grad()) { // gradient passes where cond > 0
            auto out = functional::detail::make<float>(g->shape(), false);
            device::check(gnn_relu_bwd(device::ctx(), (int64_t)g->rows(), (int32_t)g->cols(), g->dptr(), (int64_t)g->cols(), cond->dptr(),
                                       (int64_t)g->cols(), out->dptr(), (int64_t)g->cols()));
            tv->backward(out);
        }
        if (fv->requires_grad()) { // ... and to the other branch where cond <= 0
            T zfull(g->shape(), 0.0f, false);
            fv->backward(functional::mask(*cond, zfull, *g));
        }
        this->_done = true;
    }
};

template <class T> class Slice : public Operation<T> { // reference operation.h:575-616 (its backward throws; fixed)
    std::shared_ptr<tensor<int>> idx_;

  public:
    Slice() { this->name = "Slice"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &t, const std::shared_ptr<tensor<int>> &idx, int dim = -1) {
        auto out = functional::slice<float, int>(*t, *idx, dim);
        if (out->requires_grad()) {
            this->context->save_for_backward({t});
            idx_ = idx;
        }
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        check_backward(var, 1);
        auto t = var[0];
        if (t->requires_grad()) { // scatter g[i] into column idx[i] of a zero matrix (host-side index build is I/O-sized: N ints)
            auto *ih = idx_->data();
            auto *gh = g->data();
            auto *dense = new std::valarray<float>(0.0f, t->numel());
            const size_t C = t->cols();
            for (size_t i = 0; i < ih->size(); i++) (*dense)[i * C + (size_t)(*ih)[i]] = (*gh)[i];
            t->backward(std::make_shared<T>(t->shape(), dense, false));
        }
        this->_done = true;
    }
};

// ---------------------------------------------------------------------------------------------------------
// fused hot-path nodes (new)
// ---------------------------------------------------------------------------------------------------------
/** y = x W^T (+ b)(ReLU) — nn::Linear::forward (reference nn.cpp:205-211) as one NT GEMM with fused epilogue;
 *  backward = TN GEMM for dW, column sum for db, NN GEMM (+ReLU mask) for dx. */
template <class T> class LinearOp : public Operation<T> {
    bool relu_ = false;
    device::buffer_ptr out_; // output buffer kept for the ReLU mask (a buffer, not the tensor: no ownership cycle)

  public:
    int precision = device::default_gemm_precision(); // round 1 left this at 0 (FP32 FMA): the op-node path ran 83 ms per products-shaped step against the fused trainer's 46
    LinearOp() { this->name = "Linear"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &x, const std::shared_ptr<T> &W, const std::shared_ptr<T> &b, bool relu = false) {
        if (x->rank() != 2 || W->rank() != 2 || x->shape()[1] != W->shape()[1]) throw std::runtime_error(err::mm_compatible());
        const int64_t M = x->shape()[0], K = x->shape()[1], N = W->shape()[0];
        const bool rg = x->requires_grad() || W->requires_grad() || (b && b->requires_grad());
        auto out = functional::detail::make<float>({(size_t)M, (size_t)N}, rg);
        device::check(gnn_gemm_nt(device::ctx(), M, (int32_t)N, (int32_t)K, x->dptr(), K, W->dptr(), K, out->dptr(), N, b ? b->dptr() : nullptr, relu, precision));
        if (rg) {
            this->context->save_for_backward(b ? std::vector<std::shared_ptr<T>>{x, W, b} : std::vector<std::shared_ptr<T>>{x, W});
            relu_ = relu;
            if (relu) out_ = out->buffer();
        }
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        if (var.size() < 2) throw std::runtime_error("cant backprop without executing a forward computation first");
        auto x = var[0], W = var[1];
        const int64_t M = x->shape()[0], K = x->shape()[1], N = W->shape()[0];
        if (relu_) { // g <- g . [out > 0]
                        auto gm = functional::detail::make<float>(g->shape(), false);
            device::check(gnn_relu_bwd(device::ctx(), M, (int32_t)N, g->dptr(), N, static_cast<const float *>(out_->ptr), N, gm->dptr(), N));
            g = gm;
        }
        if (var.size() == 3 && var[2]->requires_grad()) {
            auto db = functional::detail::make<float>(var[2]->shape(), false);
            device::check(gnn_bias_grad(device::ctx(), M, (int32_t)N, g->dptr(), N, db->dptr()));
            var[2]->backward(db);
        }
        if (W->requires_grad()) {
            auto dW = functional::detail::make<float>(W->shape(), false);
            device::check(gnn_gemm_tn(device::ctx(), M, (int32_t)N, (int32_t)K, g->dptr(), N, x->dptr(), K, dW->dptr(), K, precision));
            W->backward(dW);
        }
        if (x->requires_grad()) {
            auto dx = functional::detail::make<float>(x->shape(), false);
            device::check(gnn_gemm_nn(device::ctx(), M, (int32_t)K, (int32_t)N, g->dptr(), N, W->dptr(), K, dx->dptr(), K, nullptr, 0, precision));
            x->backward(dx);
        }
        out_.reset();
        this->_done = true;
    }
};

/** Y = A_hat P (+ b)(ReLU) over the device CSR; backward dP = A_hat^T (g . mask) over the CSC (no atomics).
 *  Replaces adj_mat->mm(x) on a dense N x N matrix (reference graph.cpp:208, operation.h:524-531). */
template <class T> class SpMM : public Operation<T> {
    device::graph_ptr graph_;
    bool relu_ = false, use_values_ = true;
    device::buffer_ptr out_;

  public:
    SpMM() { this->name = "SpMM"; }
    std::shared_ptr<T> forward(const device::graph_ptr &graph, const std::shared_ptr<T> &P, const std::shared_ptr<T> &b = nullptr, bool relu = false,
                               bool use_values = true) {
        const int64_t n = gnn_graph_rows(graph->g), F = P->shape()[1];
        // a row block of a partitioned graph (Data::partitioned): P holds this rank's rows, the aggregation needs every rank's
        const bool part = device::dist().active() && n != gnn_graph_cols(graph->g);
        if (P->rank() != 2 || (int64_t)P->shape()[0] != (part ? n : (int64_t)gnn_graph_cols(graph->g))) throw std::runtime_error(err::mm_compatible());
        const bool rg = P->requires_grad() || (b && b->requires_grad());
        auto out = functional::detail::make<float>({(size_t)n, (size_t)F}, rg);
        device::buffer_ptr gathered;
        const float *src = P->dptr();
        if (part) src = gather_rows(P->dptr(), n, F, gathered);
        device::check(gnn_spmm_fwd(device::ctx(), graph->g, src, F, (int32_t)F, out->dptr(), F, b ? b->dptr() : nullptr, relu, nullptr, 0, use_values));
        if (rg) {
            this->context->save_for_backward(b ? std::vector<std::shared_ptr<T>>{P, b} : std::vector<std::shared_ptr<T>>{P});
            graph_ = graph;
            relu_ = relu;
            use_values_ = use_values;
            if (relu) out_ = out->buffer();
        }
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        if (var.empty()) throw std::runtime_error("cant backprop without executing a forward computation first");
        auto P = var[0];
        const int64_t n = g->shape()[0], F = g->shape()[1];
        if (relu_) {
                        auto gm = functional::detail::make<float>(g->shape(), false);
            device::check(gnn_relu_bwd(device::ctx(), n, (int32_t)F, g->dptr(), F, static_cast<const float *>(out_->ptr), F, gm->dptr(), F));
            g = gm;
        }
        if (var.size() == 2 && var[1]->requires_grad()) {
            auto db = functional::detail::make<float>(var[1]->shape(), false);
            device::check(gnn_bias_grad(device::ctx(), n, (int32_t)F, g->dptr(), F, db->dptr()));
            var[1]->backward(db);
        }
        if (P->requires_grad()) {
            auto dP = functional::detail::make<float>(P->shape(), false);
            device::buffer_ptr gathered;
            const float *src = g->dptr();
            if (device::dist().active() && gnn_graph_rows(graph_->g) != gnn_graph_cols(graph_->g)) src = gather_rows(g->dptr(), n, F, gathered);
            device::check(gnn_spmm_bwd(device::ctx(), graph_->g, src, F, (int32_t)F, dP->dptr(), F, nullptr, 0, use_values_));
            P->backward(dP);
        }
        out_.reset();
        this->_done = true;
    }

  private:
    /** all ranks' row blocks in global row order (the exchange of the row-partitioned aggregation, K10): the rank's n rows
     *  are staged into a chunk-row block (the last rank may own fewer) and all-gathered over NCCL */
    static const float *gather_rows(const float *local, int64_t n, int64_t F, device::buffer_ptr &keep) {
        const auto &d = device::dist();
        keep = device::alloc((size_t)(d.world + 1) * d.chunk * F * 4);
        float *all = static_cast<float *>(keep->ptr), *stage = all + (size_t)d.world * d.chunk * F;
        device::check(gnn_memset(device::ctx(), stage, 0, (size_t)d.chunk * F * 4));
        device::check(gnn_memcpy_d2d(device::ctx(), stage, local, (size_t)n * F * 4));
        device::check(gnn_allgather_rows(device::ctx(), stage, all, d.chunk, (int32_t)F));
        return all;
    }
};

/** loss = mean_i -log(exp(z_iy) / (sum_c exp(z_ic) + 1e-20)) — nn::cross_entropy_loss (reference nn.cpp:442-453);
 *  backward = incoming * (softmax(Z) - onehot(y)) / N, produced by the same fused kernel. */
template <class T> class SoftmaxCrossEntropy : public Operation<T> {
    std::shared_ptr<T> dZ_;

  public:
    SoftmaxCrossEntropy() { this->name = "CrossEntropy"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &logits, const std::shared_ptr<tensor<int>> &target) {
        if (logits->rank() != 2 || target->rank() != 1 || target->numel() != logits->shape()[0])
            throw std::runtime_error("invalid input, logits must be of rank 2 and targets must be 1D tensor");
        const int64_t N = logits->shape()[0], C = logits->shape()[1];
        auto out = functional::detail::make<float>({1}, logits->requires_grad());
        if (logits->requires_grad()) {
            dZ_ = functional::detail::make<float>(logits->shape(), false);
            this->context->save_for_backward({logits});
        }
        // under a row partition the mean runs over ALL nodes: the local value is this rank's share of the global loss
        const int64_t n_total = device::dist().active() && device::dist().n_global ? device::dist().n_global : N;
        device::check(gnn_softmax_xent(device::ctx(), N, (int32_t)C, logits->dptr(), C, target->dptr(), n_total, out->dptr(), dZ_ ? dZ_->dptr() : nullptr, C));
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        check_backward(var, 1);
        if (var[0]->requires_grad()) var[0]->backward(functional::binary(GNN_OP_MUL, *dZ_, *g, false)); // g is the scalar seed
        dZ_.reset();
        this->_done = true;
    }
};

/** y = relu?((x - mean) / sqrt(var + eps) * gamma + beta) with batch statistics over the node dimension — one fused node
 *  for nn::BatchNorm (+ the nn::ReLU GCNConv applies next); the backward is the standard batch-norm gradient
 *  (gnn_batchnorm_bwd), which the reference's own autograd cannot produce (fan-out, bug B2). */
template <class T> class BatchNormOp : public Operation<T> {
    device::buffer_ptr mean_, var_, out_;
    float eps_ = 1e-5f;
    bool relu_ = false;

  public:
    BatchNormOp() { this->name = "BatchNorm"; }
    const device::buffer_ptr &mean() const { return mean_; }
    const device::buffer_ptr &var() const { return var_; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &x, const std::shared_ptr<T> &gamma, const std::shared_ptr<T> &beta, float eps, bool relu) {
        if (x->rank() != 2 || gamma->numel() != x->shape()[1]) throw std::runtime_error(err::size_mismatch());
        const int64_t N = x->shape()[0], F = x->shape()[1];
        const bool rg = x->requires_grad() || gamma->requires_grad() || (beta && beta->requires_grad());
        auto out = functional::detail::make<float>(x->shape(), rg);
        mean_ = device::alloc(F * 4);
        var_ = device::alloc(F * 4);
        device::check(gnn_batchnorm_fwd(device::ctx(), N, (int32_t)F, x->dptr(), F, gamma->dptr(), beta ? beta->dptr() : nullptr, eps, relu,
                                        out->dptr(), F, static_cast<float *>(mean_->ptr), static_cast<float *>(var_->ptr)));
        if (rg) {
            this->context->save_for_backward(beta ? std::vector<std::shared_ptr<T>>{x, gamma, beta} : std::vector<std::shared_ptr<T>>{x, gamma});
            eps_ = eps;
            relu_ = relu;
            if (relu) out_ = out->buffer();
        }
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        if (var.size() < 2) throw std::runtime_error("cant backprop without executing a forward computation first");
        auto x = var[0], gamma = var[1];
        const int64_t N = x->shape()[0], F = x->shape()[1];
        auto dx = functional::detail::make<float>(x->shape(), false);
        auto dg = functional::detail::make<float>(gamma->shape(), false);
        auto db = functional::detail::make<float>(gamma->shape(), false);
        device::check(gnn_batchnorm_bwd(device::ctx(), N, (int32_t)F, x->dptr(), F, static_cast<const float *>(mean_->ptr),
                                        static_cast<const float *>(var_->ptr), gamma->dptr(), eps_,
                                        relu_ ? static_cast<const float *>(out_->ptr) : nullptr, F, g->dptr(), F, dx->dptr(), F, dg->dptr(), db->dptr()));
        if (x->requires_grad()) x->backward(dx);
        if (gamma->requires_grad()) gamma->backward(dg);
        if (var.size() == 3 && var[2]->requires_grad()) var[2]->backward(db);
        out_.reset();
        this->_done = true;
    }
};

/** nn::LayerNorm (+ the nn::ReLU nn::MLP applies next) as one node; standard layer-norm gradient (gnn_layernorm_bwd) */
template <class T> class LayerNormOp : public Operation<T> {
    device::buffer_ptr mean_, rstd_, out_;
    bool relu_ = false;

  public:
    LayerNormOp() { this->name = "LayerNorm"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &x, const std::shared_ptr<T> &gamma, const std::shared_ptr<T> &beta, float eps, bool relu) {
        if (x->rank() != 2 || (gamma && gamma->numel() != x->shape()[1])) throw std::runtime_error(err::size_mismatch());
        const int64_t N = x->shape()[0], F = x->shape()[1];
        const bool rg = x->requires_grad() || (gamma && gamma->requires_grad()) || (beta && beta->requires_grad());
        auto out = functional::detail::make<float>(x->shape(), rg);
        mean_ = device::alloc(N * 4);
        rstd_ = device::alloc(N * 4);
        device::check(gnn_layernorm_fwd(device::ctx(), N, (int32_t)F, x->dptr(), F, gamma ? gamma->dptr() : nullptr, beta ? beta->dptr() : nullptr,
                                        eps, relu, out->dptr(), F, static_cast<float *>(mean_->ptr), static_cast<float *>(rstd_->ptr)));
        if (rg) {
            std::vector<std::shared_ptr<T>> saved{x};
            if (gamma) saved.push_back(gamma);
            if (beta) saved.push_back(beta);
            this->context->save_for_backward(saved);
            this->context->saved_data["gamma"] = gamma ? 1 : 0;
            this->context->saved_data["beta"] = beta ? 1 : 0;
            relu_ = relu;
            if (relu) out_ = out->buffer();
        }
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        if (var.empty()) throw std::runtime_error("cant backprop without executing a forward computation first");
        auto x = var[0];
        const bool has_g = this->context->saved_data["gamma"], has_b = this->context->saved_data["beta"];
        auto gamma = has_g ? var[1] : nullptr;
        auto beta = has_b ? var[has_g ? 2 : 1] : nullptr;
        const int64_t N = x->shape()[0], F = x->shape()[1];
        auto dx = functional::detail::make<float>(x->shape(), false);
        auto dg = gamma ? functional::detail::make<float>(gamma->shape(), false) : nullptr;
        auto db = beta ? functional::detail::make<float>(beta->shape(), false) : nullptr;
        device::check(gnn_layernorm_bwd(device::ctx(), N, (int32_t)F, x->dptr(), F, static_cast<const float *>(mean_->ptr),
                                        static_cast<const float *>(rstd_->ptr), gamma ? gamma->dptr() : nullptr,
                                        relu_ ? static_cast<const float *>(out_->ptr) : nullptr, F, g->dptr(), F, dx->dptr(), F,
                                        dg ? dg->dptr() : nullptr, db ? db->dptr() : nullptr));
        if (x->requires_grad()) x->backward(dx);
        if (gamma && gamma->requires_grad()) gamma->backward(dg);
        if (beta && beta->requires_grad()) beta->backward(db);
        out_.reset();
        this->_done = true;
    }
};

/** nn::tanh (reference nn.cpp:355-364) as one node: y = tanh(x + 1e-12), dx = dy (1 - y^2) */
template <class T> class TanhOp : public Operation<T> {
    device::buffer_ptr out_;

  public:
    TanhOp() { this->name = "Tanh"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &x) {
        auto out = functional::detail::make<float>(x->shape(), x->requires_grad());
        device::check(gnn_tanh_fwd(device::ctx(), (int64_t)x->numel(), x->dptr(), out->dptr()));
        if (x->requires_grad()) {
            this->context->save_for_backward({x});
            out_ = out->buffer();
        }
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        check_backward(var, 1);
        auto dx = functional::detail::make<float>(var[0]->shape(), false);
        device::check(gnn_tanh_bwd(device::ctx(), (int64_t)var[0]->numel(), static_cast<const float *>(out_->ptr), g->dptr(), dx->dptr()));
        var[0]->backward(dx);
        out_.reset();
        this->_done = true;
    }
};

/** nn::Dropout (reference nn.cpp:246-266) as one node with a seeded, counter-based keep mask */
template <class T> class DropoutOp : public Operation<T> {
    float p_ = 0.f;
    uint64_t seed_ = 0;

  public:
    DropoutOp() { this->name = "Dropout"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &x, float p, uint64_t seed) {
        auto out = functional::detail::make<float>(x->shape(), x->requires_grad());
        device::check(gnn_dropout(device::ctx(), (int64_t)x->numel(), x->dptr(), p, seed, out->dptr()));
        if (x->requires_grad()) {
            this->context->save_for_backward({x});
            p_ = p;
            seed_ = seed;
        }
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        check_backward(var, 1);
        auto dx = functional::detail::make<float>(var[0]->shape(), false);
        device::check(gnn_dropout(device::ctx(), (int64_t)var[0]->numel(), g->dptr(), p_, seed_, dx->dptr()));
        var[0]->backward(dx);
        this->_done = true;
    }
};

/** cross-entropy over the rows selected by a node mask (graph::Data::set_mask): gnn_softmax_xent_masked */
template <class T> class MaskedSoftmaxCrossEntropy : public Operation<T> {
    std::shared_ptr<T> dZ_;

  public:
    MaskedSoftmaxCrossEntropy() { this->name = "MaskedCrossEntropy"; }
    std::shared_ptr<T> forward(const std::shared_ptr<T> &logits, const std::shared_ptr<tensor<int>> &target,
                               const device::buffer_ptr &mask_u8, int64_t n_selected) {
        if (logits->rank() != 2 || target->rank() != 1 || target->numel() != logits->shape()[0])
            throw std::runtime_error("invalid input, logits must be of rank 2 and targets must be 1D tensor");
        const int64_t N = logits->shape()[0], C = logits->shape()[1];
        auto out = functional::detail::make<float>({1}, logits->requires_grad());
        if (logits->requires_grad()) {
            dZ_ = functional::detail::make<float>(logits->shape(), false);
            this->context->save_for_backward({logits});
        }
        device::check(gnn_softmax_xent_masked(device::ctx(), N, (int32_t)C, logits->dptr(), C, target->dptr(),
                                              static_cast<const uint8_t *>(mask_u8->ptr), n_selected, out->dptr(),
                                              dZ_ ? dZ_->dptr() : nullptr, C));
        return out;
    }
    void _backward(std::shared_ptr<T> g) override {
        auto var = this->context->get_variables();
        check_backward(var, 1);
        if (var[0]->requires_grad()) var[0]->backward(functional::binary(GNN_OP_MUL, *dZ_, *g, false));
        dZ_.reset();
        this->_done = true;
    }
};

// ---------------------------------------------------------------------------------------------------------
// engine implementation (needs the complete tensor type)
// ---------------------------------------------------------------------------------------------------------
namespace autograd {
inline void topo(ftensor *t, std::unordered_set<ftensor *> &seen, std::vector<ftensor *> &order) {
    if (!seen.insert(t).second) return;
    if (t->grad_fn)
        for (auto &in : t->grad_fn->inputs()) topo(in.get(), seen, order);
    order.push_back(t);
}
inline std::shared_ptr<ftensor> sum_to_shape(const std::shared_ptr<ftensor> &g, const dims_t &target) {
    size_t gr, gc, tr, tc;
    as_2d(g->shape(), gr, gc);
    as_2d(target, tr, tc);
    if (gr == tr && gc == tc) {
        if (g->shape() == target) return g;
        return std::make_shared<ftensor>(target, g->buffer(), false); // same data, other rank
    }
    int dim;
    if (tr == 1 && tc == 1) dim = -1;
    else if (tr == 1 && gr > 1 && tc == gc) dim = 0; // bias [F] from [N,F]: ascending-row column sum (tensor.h:618-638)
    else if (tc == 1 && gc > 1 && tr == gr) dim = 1;
    else throw std::runtime_error("dims is not broacastable to this tensor's size");
    auto out = functional::detail::make<float>(target, false);
    device::check(gnn_sum_f32(device::ctx(), (int64_t)gr, (int64_t)gc, g->dptr(), dim, out->dptr()));
    return out;
}
inline void run_or_defer(ftensor *root, std::shared_ptr<ftensor> seed) {
    Engine *&cur = current();
    if (cur != nullptr) { // called from inside a node's _backward: queue, the engine will sum and dispatch
        cur->add(root, seed);
        return;
    }
    Engine engine;
    struct Guard { Engine *&slot; ~Guard() { slot = nullptr; } } guard{cur};
    cur = &engine;
    std::unordered_set<ftensor *> seen;
    std::vector<ftensor *> order;
    topo(root, seen, order);
    engine.add(root, seed);
    for (auto it = order.rbegin(); it != order.rend(); ++it) {
        ftensor *t = *it;
        auto f = engine.pending.find(t);
        if (f == engine.pending.end()) continue;
        std::shared_ptr<ftensor> g = f->second;
        engine.pending.erase(f);
        if (t->requires_grad() && !t->grad_fn) { // leaf: accumulate into .grad (reference tensor.h:268-271)
            device::check(gnn_binary_f32(device::ctx(), GNN_OP_ADD, 1, (int64_t)t->numel(), t->grad_dptr(), 0, 1, g->dptr(), 0, 1, t->grad_dptr()));
        }
        if (t->grad_fn) t->grad_fn->backward(g);
    }
}
} // namespace autograd

} // namespace cyg
#endif

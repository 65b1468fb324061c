// tensor.h — cyg::tensor<T> with DEVICE-RESIDENT storage (counterpart of reference include/tensor.h:21-913).
//
// Same surface as the reference's tensor (constructors, data()/set_data, shape/numel/rank, grad(),
// requires_grad_, zero_grad, backward, add/mul/div/mm/t/where/exp/log/sum/mean/at/item/max/argmax/gt/clone/
// uniform/fill_diagonal_, the tptr operators, randn/eye/ones_like/zeros_like/no_grad/enable_grad), but
//   * `_data` is a device buffer owned through a shared_ptr (the reference leaks a heap std::valarray,
//     tensor.h:159,825-828); `data()` returns a HOST MIRROR refreshed by a device->host copy — explicit I/O,
//     not a compute path.  Writes must go through set_data().
//   * every arithmetic method enqueues sm_100a kernels through the C ABI (include/gnn_c.h); there is no CPU
//     arithmetic anywhere.  Rank <= 2 (the GCN hot path); `bool` tensors are stored as 0/1 floats.
//   * backward() is a reverse-topological engine: gradients arriving at a tensor used several times are
//     summed before flowing on (the reference drops them, SURVEY.md bug B2), and `.grad` storage is only
//     allocated for leaves.
#ifndef GNNB200_TENSOR_H
#define GNNB200_TENSOR_H

#include <algorithm>
#include <climits>
#include <memory>
#include <sstream>
#include <tuple>
#include <type_traits>
#include <unordered_map>
#include <unordered_set>
#include <valarray>
#include <vector>

#include "device.h"
#include "utils.h"

namespace cyg {

template <class T> class tensor;
template <class A> using tptr = std::shared_ptr<tensor<A>>;
template <class T> class Operation;

// element type used on the device for a host element type
template <class T> struct dev_elem { using type = T; };
template <> struct dev_elem<bool> { using type = float; };

} // namespace cyg

#include "functional.h" // declarations only need the forward-declared tensor

namespace cyg {

template <class T> class tensor : public std::enable_shared_from_this<tensor<T>> {
    using D = typename dev_elem<T>::type;
    static constexpr bool is_float = std::is_same<T, float>::value;

  public:
    typedef T value_type;
    std::unique_ptr<Operation<tensor<T>>> grad_fn;

    // (dims, fill value, requires_grad) — reference tensor.h:106-115
    explicit tensor(std::vector<size_t> dims, T value = 0, bool requires_grad = false) : _dims(dims), _requires_grad(requires_grad) {
        check_valid_dims(dims);
        if (requires_grad && !is_float) throw std::runtime_error(err::grad_dtype());
        _buf = device::alloc(count_elements(dims) * sizeof(D));
        fill(value);
    }
    // (dims, host data — ownership taken, uploaded then released —, requires_grad) — reference tensor.h:122-139
    explicit tensor(std::vector<size_t> dims, std::valarray<T> *data, bool requires_grad = false) : _dims(dims), _requires_grad(requires_grad) {
        check_valid_dims(dims);
        if (requires_grad && !is_float) throw std::runtime_error(err::grad_dtype());
        const size_t n = count_elements(dims);
        if (data != nullptr && data->size() != n) {
            delete data;
            throw std::runtime_error(err::size_mismatch());
        }
        _buf = device::alloc(n * sizeof(D));
        if (data != nullptr) {
            upload(*data);
            delete data;
        } else {
            fill(T(0));
        }
    }
    // adopt an existing device buffer (used by the op kernels' outputs)
    tensor(std::vector<size_t> dims, device::buffer_ptr buf, bool requires_grad) : _dims(dims), _buf(buf), _requires_grad(requires_grad) {}
    ~tensor() { delete _host; delete _host_grad; }
    tensor(const tensor &) = delete;
    tensor &operator=(const tensor &) = delete;

    // ---- storage ---------------------------------------------------------------------------------------
    D *dptr() const { return static_cast<D *>(_buf->ptr); }
    device::buffer_ptr buffer() const { return _buf; }
    /** host mirror of the data (device -> host copy on every call). reference tensor.h:160 */
    std::valarray<T> *data() const {
        const size_t n = numel();
        std::vector<D> tmp(n);
        device::check(gnn_memcpy_d2h(device::ctx(), tmp.data(), _buf->ptr, n * sizeof(D)));
        if (_host == nullptr) _host = new std::valarray<T>(n);
        if (_host->size() != n) _host->resize(n);
        for (size_t i = 0; i < n; i++) (*_host)[i] = static_cast<T>(tmp[i]);
        return _host;
    }
    template <class A> void set_data(std::valarray<A> *data) { // reference tensor.h:161-172
        if (data->size() != numel()) throw std::runtime_error(err::size_mismatch());
        std::valarray<T> conv(numel());
        for (size_t i = 0; i < numel(); i++) conv[i] = static_cast<T>((*data)[i]);
        upload(conv);
        if (_requires_grad) zero_grad();
    }
    std::shared_ptr<tensor<float>> to_float() { return cast_to<float>(); }
    std::shared_ptr<tensor<int>> to_int() { return cast_to<int>(); }
    std::shared_ptr<tensor<bool>> to_bool() { return cast_to<bool>(); }
    /** zero-copy float view of a float/bool tensor (bool is stored as 0/1 floats on the device) */
    std::shared_ptr<tensor<float>> to_float_view() {
        if constexpr (std::is_same<D, float>::value) return std::make_shared<tensor<float>>(_dims, _buf, false);
        else return to_float();
    }

    std::vector<size_t> shape() const { return _dims; }
    size_t numel() const { return count_elements(_dims); }
    int rank() const { return (int)_dims.size(); }
    size_t rows() const { size_t r, c; as_2d(_dims, r, c); return r; }
    size_t cols() const { size_t r, c; as_2d(_dims, r, c); return c; }

    // ---- autograd state --------------------------------------------------------------------------------
    /** gradient of a leaf tensor as a host mirror. reference tensor.h:207-214 */
    std::valarray<float> *grad() {
        if (grad_fn != nullptr) throw std::runtime_error(err::grad_not_leaf());
        if (!_requires_grad) throw std::runtime_error("invalid op, pls enable grad on this tensor");
        ensure_grad();
        const size_t n = numel();
        if (_host_grad == nullptr) _host_grad = new std::valarray<float>(n);
        device::check(gnn_memcpy_d2h(device::ctx(), &(*_host_grad)[0], _grad->ptr, n * 4));
        return _host_grad;
    }
    float *grad_dptr() { ensure_grad(); return static_cast<float *>(_grad->ptr); }
    bool requires_grad() const { return _requires_grad; }
    void requires_grad_(bool rg) { // reference tensor.h:220-233
        if (rg && !is_float) throw std::runtime_error(err::grad_dtype());
        _requires_grad = rg;
        if (rg) zero_grad();
        else _grad.reset();
    }
    void zero_grad() { // reference tensor.h:234-238
        if (_grad) device::check(gnn_memset(device::ctx(), _grad->ptr, 0, numel() * 4));
    }
    void squeeze() { _dims.erase(std::remove(_dims.begin(), _dims.end(), (size_t)1), _dims.end()); if (_dims.empty()) _dims = {1}; }
    tptr<T> unsqueeze(int dim) {
        if (dim > rank() || dim < -rank() - 1) throw std::runtime_error(err::out_of_range());
        _dims.insert(_dims.begin() + (dim < 0 ? dim + rank() + 1 : dim), 1);
        if (_dims.size() > 2) throw std::runtime_error(err::rank_limit());
        return this->shared_from_this();
    }
    /** reverse-mode differentiation from this tensor. reference tensor.h:260-276 */
    void backward(std::shared_ptr<tensor<float>> incoming_gradient = nullptr);

    // ---- indexing (host mirror reads: I/O) -------------------------------------------------------------
    template <typename... A> tptr<T> operator()(const A &...d) const { // reference tensor.h:281-293
        std::vector<size_t> idx = {(size_t)d...};
        if (idx.size() > _dims.size()) throw std::runtime_error(err::bad_dim());
        size_t flat = 0, stride = 1;
        for (int i = (int)_dims.size() - 1; i >= 0; i--) {
            const size_t v = i < (int)idx.size() ? idx[i] : 0;
            if (v >= _dims[i]) throw std::runtime_error(err::out_of_range());
            flat += v * stride;
            stride *= _dims[i];
        }
        D v;
        device::check(gnn_memcpy_d2h(device::ctx(), &v, dptr() + flat, sizeof(D)));
        return std::make_shared<tensor<T>>(std::vector<size_t>{1}, static_cast<T>(v), false);
    }
    T operator[](size_t i) const { // 1-D element read (the reference returns a reference into host storage)
        if (rank() != 1) throw std::runtime_error("tensor must be 1D");
        if (i >= numel()) throw std::runtime_error("invalid index");
        D v;
        device::check(gnn_memcpy_d2h(device::ctx(), &v, dptr() + i, sizeof(D)));
        return static_cast<T>(v);
    }
    T item() { // reference tensor.h:653-658
        if (numel() != 1) throw std::runtime_error("invalid op, tensor must be scalar");
        return flat_item();
    }

    // ---- arithmetic: every method builds an Operation node like the reference (tensor.h:309-613) -------
    tptr<T> add(const tptr<T> &other);
    template <class A> tptr<T> add(const A &other) { return add(std::make_shared<tensor<T>>(std::vector<size_t>{1}, static_cast<T>(other), false)); }
    tptr<T> mul(const tptr<T> &other);
    template <class A> tptr<T> mul(const A &other) { return mul(std::make_shared<tensor<T>>(std::vector<size_t>{1}, static_cast<T>(other), false)); }
    tptr<T> div(const tptr<T> &other);
    template <class A> tptr<T> div(const A &other) { return div(std::make_shared<tensor<T>>(std::vector<size_t>{1}, static_cast<T>(other), false)); }
    tptr<T> mm(const tptr<T> &other);
    tptr<T> where(const tptr<bool> cond, const tptr<T> &other);
    template <class A> tptr<T> where(const tptr<bool> &cond, const A &other) {
        return where(cond, std::make_shared<tensor<T>>(_dims, static_cast<T>(other), false));
    }
    tptr<T> exp();
    tptr<T> log();
    tptr<float> mean(int dim = INT_MAX, const bool &keepdim = false);
    tptr<T> sum(int dim = INT_MAX, const bool &keepdim = false);
    tptr<T> t(int d1 = -1, int d2 = -2);
    tptr<T> at(const tptr<int> &idx, int dim = -1);
    std::shared_ptr<tensor<bool>> gt(const tptr<T> &other);
    std::shared_ptr<tensor<bool>> gt(const float &other) { return gt(std::make_shared<tensor<T>>(std::vector<size_t>{1}, static_cast<T>(other), false)); }
    std::tuple<tptr<T>, tptr<int>> max(int dim = INT_MAX, const bool &keepdim = false) const { return functional::max<T>(*this, dim, keepdim); }
    tptr<int> argmax(int dim = INT_MAX, const bool &keepdim = false) { return std::get<1>(max(dim, keepdim)); }

    /** deep copy; fillValue != INT_MAX fills instead. reference tensor.h:757-765 (fixed: B3's zero-length assign) */
    tptr<T> clone(const bool &require_grad = false, const T fillValue = static_cast<T>(INT_MAX)) const {
        auto out = std::make_shared<tensor<T>>(_dims, device::alloc(numel() * sizeof(D)), false);
        if (fillValue != static_cast<T>(INT_MAX)) out->fill(fillValue);
        else device::check(gnn_memcpy_d2d(device::ctx(), out->_buf->ptr, _buf->ptr, numel() * sizeof(D)));
        if (require_grad) out->requires_grad_(true);
        return out;
    }
    void uniform(const float &low, const float &high) { // reference tensor.h:695-699 (seeded generator, see utils.h)
        std::valarray<T> h(numel());
        for (size_t i = 0; i < numel(); i++) h[i] = static_cast<T>(generate_random(low, high));
        upload(h);
    }
    template <class A> void fill_diagonal_(const A &value = 0) { // reference tensor.h:806-817
        if (_requires_grad) throw std::runtime_error(err::in_place_leaf());
        if (rank() != 2 || _dims[0] != _dims[1]) throw std::runtime_error("all dimensions must be of same length and tensor must be 2D");
        // strided device write of N elements (one 2-D memset-like copy from a small host vector)
        std::vector<D> diag(_dims[0], static_cast<D>(value));
        device::check(gnn_memcpy_h2d_strided(diag.data(), dptr(), _dims[0], _dims[1] + 1));
    }
    void fill(T value) {
        if (std::is_same<D, float>::value) device::check(gnn_fill_f32(device::ctx(), reinterpret_cast<float *>(dptr()), static_cast<float>(value), (int64_t)numel()));
        else {
            std::valarray<T> h(value, numel());
            upload(h);
        }
    }

  protected:
    std::vector<size_t> _dims;
    device::buffer_ptr _buf;
    device::buffer_ptr _grad;
    bool _requires_grad = false;
    mutable std::valarray<T> *_host = nullptr;
    mutable std::valarray<float> *_host_grad = nullptr;

    T flat_item() const {
        D v;
        device::check(gnn_memcpy_d2h(device::ctx(), &v, dptr(), sizeof(D)));
        return static_cast<T>(v);
    }
    void ensure_grad() {
        if (!_grad) {
            _grad = device::alloc(numel() * 4);
            device::check(gnn_memset(device::ctx(), _grad->ptr, 0, numel() * 4));
        }
    }
    void upload(const std::valarray<T> &h) {
        std::vector<D> tmp(h.size());
        for (size_t i = 0; i < h.size(); i++) tmp[i] = static_cast<D>(h[i]);
        device::check(gnn_memcpy_h2d(device::ctx(), _buf->ptr, tmp.data(), tmp.size() * sizeof(D)));
        device::sync(); // tmp goes out of scope
    }
    static int gnn_memcpy_h2d_strided(const D *src, D *dst, size_t n, size_t stride) {
        for (size_t i = 0; i < n; i++) {
            int rc = gnn_memcpy_h2d(device::ctx(), dst + i * stride, src + i, sizeof(D));
            if (rc) return rc;
        }
        return gnn_ctx_sync(device::ctx());
    }
    template <class A> std::shared_ptr<tensor<A>> cast_to() {
        auto *h = data();
        auto *conv = new std::valarray<A>(numel());
        for (size_t i = 0; i < numel(); i++) (*conv)[i] = static_cast<A>((*h)[i]);
        return std::make_shared<tensor<A>>(_dims, conv, false);
    }
    template <class U> friend class tensor;
    friend struct functional::detail;
};

// ---- operators on tptr (reference tensor.h:31-95) ------------------------------------------------------
template <class T, class A> tptr<T> operator+(tptr<T> lhs, const A &rhs) { return lhs->add(rhs); }
template <class T> tptr<T> operator-(const tptr<T> &lhs) { return lhs->mul(-1); }
template <class T> tptr<T> operator-(tptr<T> lhs, const tptr<T> &rhs) { return lhs->add(rhs->mul(-1)); }
template <class T> tptr<T> operator-(tptr<T> lhs, const float &rhs) { return lhs->add(-rhs); }
template <class T> tptr<T> operator*(tptr<T> lhs, const tptr<T> &rhs) { return lhs->mul(rhs); }
template <class T> tptr<T> operator*(tptr<T> lhs, const float &rhs) { return lhs->mul(rhs); }
template <class T> tptr<T> operator/(tptr<T> lhs, const tptr<T> &rhs) { return lhs->div(rhs); }
template <class T> tptr<T> operator/(tptr<T> lhs, const float &rhs) { return lhs->div(rhs); }
template <class T> tptr<bool> operator>(const tptr<T> lhs, const tptr<T> &rhs) { return lhs->gt(rhs); }
template <class T> tptr<bool> operator>(const tptr<T> lhs, const float &rhs) { return lhs->gt(rhs); }

template <class T> std::ostream &operator<<(std::ostream &out, const tensor<T> &t) {
    auto *h = t.data();
    out << "(";
    const size_t c = t.cols();
    for (size_t i = 0; i < h->size() && i < 64; i++) out << (i && i % c == 0 ? "\n " : " ") << (*h)[i];
    if (h->size() > 64) out << " ...";
    out << ", size = " << t.shape() << ", requires_grad = " << std::boolalpha << t.requires_grad() << " )\n";
    return out;
}

/** uniform random tensor U[low, high). reference tensor.h:864-872 */
template <class T> tptr<T> randn(std::vector<size_t> dims, T low = -1, T high = 1, bool requires_grad = false) {
    if (low >= high) throw std::runtime_error("pls check input params, low must be lower than high");
    auto *vec = new std::valarray<T>(count_elements(dims));
    for (auto &v : *vec) v = static_cast<T>(generate_random((float)low, (float)high));
    return std::make_shared<tensor<T>>(dims, vec, requires_grad);
}
template <class T> tptr<int> ones_like(const tptr<T> &t, bool requires_grad = false) { return std::make_shared<tensor<int>>(t->shape(), 1, requires_grad); }
template <class T> tptr<int> zeros_like(const tptr<T> &t, bool requires_grad = false) { return std::make_shared<tensor<int>>(t->shape(), 0, requires_grad); }
void no_grad(std::vector<tptr<float>> ts);
void enable_grad(std::vector<tptr<float>> ts);
tptr<int> eye(size_t n, size_t m = INT_MAX);

} // namespace cyg

#include "operation.h"

// ---- out-of-class definitions that need the Operation nodes ---------------------------------------------
namespace cyg {

template <class T> tptr<T> tensor<T>::add(const tptr<T> &other) {
    auto op = std::make_unique<Add<tensor<T>>>();
    auto out = op->forward(this->shared_from_this(), other);
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}
template <class T> tptr<T> tensor<T>::mul(const tptr<T> &other) {
    auto op = std::make_unique<Mul<tensor<T>>>();
    auto out = op->forward(this->shared_from_this(), other);
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}
template <class T> tptr<T> tensor<T>::div(const tptr<T> &other) {
    auto op = std::make_unique<Div<tensor<T>>>();
    auto out = op->forward(this->shared_from_this(), other);
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}
template <class T> tptr<T> tensor<T>::mm(const tptr<T> &other) {
    if (rank() != 2 || other->rank() != 2 || _dims[1] != other->shape()[0]) throw std::runtime_error(err::mm_compatible());
    auto op = std::make_unique<MatMul<tensor<T>>>();
    auto out = op->forward(this->shared_from_this(), other);
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}
template <class T> tptr<T> tensor<T>::where(const tptr<bool> cond, const tptr<T> &other) {
    if (cond->numel() != numel() || other->numel() != numel()) throw std::runtime_error(err::size_mismatch());
    auto op = std::make_unique<Mask<tensor<T>>>();
    auto out = op->forward(cond->to_float_view(), this->shared_from_this(), other);
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}
template <class T> tptr<T> tensor<T>::exp() {
    auto op = std::make_unique<Exp<tensor<T>>>();
    auto out = op->forward(this->shared_from_this());
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}
template <class T> tptr<T> tensor<T>::log() {
    auto op = std::make_unique<Log<tensor<T>>>();
    auto out = op->forward(this->shared_from_this());
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}
template <class T> tptr<T> tensor<T>::sum(int dim, const bool &keepdim) {
    check_dim(dim, rank());
    auto op = std::make_unique<Sum<tensor<T>>>();
    auto out = op->forward(this->shared_from_this(), dim, keepdim);
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}
template <class T> tptr<float> tensor<T>::mean(int dim, const bool &keepdim) {
    check_dim(dim, rank());
    auto s = sum(dim, keepdim);
    const size_t n = dim == INT_MAX ? numel() : _dims[dim < 0 ? dim + rank() : dim];
    return s->div((float)n);
}
template <class T> tptr<T> tensor<T>::t(int d1, int d2) {
    if (rank() != 2 || std::abs(d1 - d2) != 1) throw std::runtime_error(err::transpose());
    auto op = std::make_unique<Transpose<tensor<T>>>();
    auto out = op->forward(this->shared_from_this(), d1, d2);
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}
template <class T> tptr<T> tensor<T>::at(const tptr<int> &idx, int dim) {
    check_dim(dim, rank());
    if (rank() != 2 || idx->numel() != _dims[0]) throw std::runtime_error("invalid op, input tensor's shape must be compatible with this tensor");
    auto op = std::make_unique<Slice<tensor<T>>>();
    auto out = op->forward(this->shared_from_this(), idx);
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}
template <class T> std::shared_ptr<tensor<bool>> tensor<T>::gt(const tptr<T> &other) { return functional::gt(*this, *other); }

template <class T> void tensor<T>::backward(std::shared_ptr<tensor<float>> incoming) {
    if constexpr (!std::is_same<T, float>::value) {
        throw std::runtime_error(err::grad_dtype());
    } else {
        if (incoming == nullptr && numel() != 1) throw std::runtime_error(err::non_scalar_backprop());
        if (incoming == nullptr) incoming = std::make_shared<tensor<float>>(_dims, 1.0f, false);
        if (incoming->numel() != numel()) throw std::runtime_error(err::grad_mismatch());
        autograd::run_or_defer(this, incoming);
    }
}

} // namespace cyg
#endif

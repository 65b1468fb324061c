// nn.cpp — see nn.h.  Registration/lookup semantics follow reference src/nn.cpp:12-151 (ordered child list,
// recursive name maps with `parent_child` prefixing on collisions); the arithmetic is device kernels.
#include "nn.h"

#include <cmath>

using namespace cyg;

namespace nn {

void Module::register_module(std::string n, Module *module) {
    module->name = n;
    _modules.push_back({n, std::shared_ptr<Module>(module)});
}
void Module::register_parameter(std::string n, tptr<float> p) {
    if (!p->requires_grad() || p->grad_fn)
        throw std::runtime_error("cannot add tensor as param, tensor requires_grad must be set to true and tensor must be non-leaf - explicitly created");
    _parameters[n] = p;
}
void Module::register_buffer(std::string n, tptr<float> b) {
    if (b->requires_grad() || b->grad_fn)
        throw std::runtime_error("cannot add tensor as buffer, tensor requires_grad must be set to false and tensor must be non-leaf - explicitly created");
    _buffers[n] = b;
}
void Module::zero_grad() {
    for (auto &kv : _parameters) kv.second->zero_grad();
    for (auto &kv : _modules) kv.second->zero_grad();
}
void Module::train(const bool &isTrain) {
    training = isTrain;
    for (auto &kv : _parameters) kv.second->requires_grad_(isTrain);
    for (auto &kv : _modules) kv.second->train(isTrain);
}
tptr<float> Module::get_parameter(std::string n) {
    auto params = named_parameters();
    auto it = params.find(n);
    if (it == params.end()) throw std::runtime_error("invalid input, no parameter with given name");
    return it->second;
}
tptr<float> Module::get_buffer(std::string n) {
    auto bufs = named_buffers();
    auto it = bufs.find(n);
    if (it == bufs.end()) throw std::runtime_error("invalid input, no buffer with given name");
    return it->second;
}
std::shared_ptr<Module> Module::get_module(std::string n) {
    for (auto &kv : _modules)
        if (kv.first == n) return kv.second; // direct children first (cheap, unambiguous)
    auto mods = named_modules();
    auto it = mods.find(n);
    if (it == mods.end()) throw std::runtime_error("invalid input, no Module with given name");
    return it->second;
}
tptr<float> Module::operator()(const tptr<float> &input_tensor, tensor<int> *y) {
    if (y == nullptr) return forward(input_tensor);
    return forward(input_tensor, y);
}
std::vector<std::shared_ptr<Module>> Module::modules(const bool &recurse) {
    std::vector<std::shared_ptr<Module>> out;
    for (auto &kv : named_modules(recurse)) out.push_back(kv.second);
    return out;
}
std::unordered_map<std::string, std::shared_ptr<Module>> Module::named_modules(const bool &recurse) {
    std::unordered_map<std::string, std::shared_ptr<Module>> res;
    if (!recurse || _modules.empty()) {
        // a module owned by raw pointer (stack object) has no shared owner: report children only
        auto self = this->weak_from_this().lock();
        if (self) res[name] = self;
        return res;
    }
    for (auto &kv : _modules) {
        for (auto &sub : kv.second->named_modules(recurse)) {
            if (res.count(sub.first)) res[kv.first + "_" + sub.first] = sub.second;
            else res[sub.first] = sub.second;
        }
    }
    return res;
}
std::vector<tptr<float>> Module::parameters(const bool &recurse) {
    // deterministic order (registration order, depth first) so optimiser state lines up run to run; the reference
    // iterates an unordered_map (nn.cpp:103-109)
    std::vector<tptr<float>> out;
    std::vector<std::string> keys;
    for (auto &kv : _parameters) keys.push_back(kv.first);
    std::sort(keys.begin(), keys.end());
    for (auto &k : keys) out.push_back(_parameters[k]);
    if (recurse)
        for (auto &kv : _modules)
            for (auto &p : kv.second->parameters(true)) out.push_back(p);
    return out;
}
std::unordered_map<std::string, tptr<float>> Module::named_parameters(const bool &recurse) {
    std::unordered_map<std::string, tptr<float>> res = _parameters;
    if (!recurse || _modules.empty()) return res;
    for (auto &kv : _modules) {
        for (auto &sub : kv.second->named_parameters(recurse)) {
            if (res.count(sub.first)) res[kv.first + "_" + sub.first] = sub.second;
            else res[sub.first] = sub.second;
        }
    }
    return res;
}
std::unordered_map<std::string, tptr<float>> Module::named_buffers(const bool &recurse) const {
    std::unordered_map<std::string, tptr<float>> res = _buffers;
    if (!recurse || _modules.empty()) return res;
    for (auto &kv : _modules)
        for (auto &sub : kv.second->named_buffers(recurse)) res[kv.first + "_" + sub.first] = sub.second;
    return res;
}
std::vector<tptr<float>> Module::buffers(const bool &recurse) const {
    std::vector<tptr<float>> out;
    for (auto &kv : named_buffers(recurse)) out.push_back(kv.second);
    return out;
}

Linear::Linear(const size_t &in_features, const size_t &out_features, const bool &bias, const std::string &n)
    : Module(n), _bias(bias), _in_features(in_features), _out_features(out_features) {
    register_parameter("weight", std::make_shared<tensor<float>>(std::vector<size_t>{out_features, in_features}, 1.0f, true));
    if (_bias) register_parameter("bias", std::make_shared<tensor<float>>(std::vector<size_t>{out_features}, 1.0f, true));
    reset_parameters();
}
void Linear::reset_parameters() { // U(-1/sqrt(in), 1/sqrt(in)) — reference nn.cpp:198-204
    const float bound = 1.0f / std::sqrt((float)_in_features);
    _parameters["weight"]->uniform(-bound, bound);
    if (_bias) _parameters["bias"]->uniform(-bound, bound);
}
tptr<float> Linear::forward(const tptr<float> &x) {
    auto op = std::make_unique<LinearOp<tensor<float>>>();
    auto out = op->forward(x, _parameters["weight"], _bias ? _parameters["bias"] : nullptr, false);
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}

Sequential::Sequential(std::vector<std::pair<std::string, Module *>> input, const std::string &n) : Module(n) {
    for (auto &kv : input) register_module(kv.first, kv.second);
}
tptr<float> Sequential::forward(const tptr<float> &x) {
    tptr<float> out = x;
    for (auto &kv : _modules) out = (*kv.second)(out);
    return out;
}

tptr<float> ReLU::forward(const tptr<float> &x) {
    auto cond = x > 0.0f;
    auto out = x->where(cond, 0.0f);
    if (out->grad_fn) out->grad_fn->name = name;
    return out;
}

tptr<float> softmax(const tptr<float> &x, int dim) {
    auto sum_ = x->exp()->sum(dim, true);
    auto out = (x - sum_->log())->exp();
    if (out->grad_fn) out->grad_fn->name = "SoftmaxOp";
    return out;
}

BatchNorm::BatchNorm(const size_t &num_features, const float &eps, const float &momentum, const bool &affine,
                     const bool &track_running_stats, const std::string &n)
    : Module(n), _num_features((int)num_features), _eps(eps), _momentum(momentum), _affine(affine), _tracking_running_stats(track_running_stats) {
    std::vector<size_t> dims = {1, num_features};
    register_parameter("gammas", std::make_shared<tensor<float>>(dims, 1.0f, true));
    if (affine) register_parameter("betas", std::make_shared<tensor<float>>(dims, 0.0f, true));
    if (_tracking_running_stats) {
        register_buffer("running_mean", std::make_shared<tensor<float>>(dims, 0.0f, false));
        register_buffer("running_var", std::make_shared<tensor<float>>(dims, 0.0f, false));
    }
    training = true;
}

tptr<float> BatchNorm::forward_relu(const tptr<float> &x, bool relu) {
    if (x->rank() != 2 || (int)x->shape()[1] != _num_features) throw std::runtime_error(err::size_mismatch());
    auto gamma = _parameters["gammas"];
    auto beta = _affine ? _parameters["betas"] : nullptr;
    if (!training && _tracking_running_stats) { // evaluation: running statistics, composed from tensor ops (nn.cpp:307-309)
        auto stdv = ((_buffers["running_var"] + _eps)->log() * 0.5f)->exp(); // (var + eps)^0.5 from the ops this surface has
        auto scaled = (x - _buffers["running_mean"]) / stdv;
        auto out = scaled * gamma;
        if (beta) out = out + beta;
        return relu ? out->where(out > 0.0f, 0.0f) : out;
    }
    auto op = std::make_unique<BatchNormOp<tensor<float>>>();
    auto out = op->forward(x, gamma, beta, _eps, relu);
    if (_tracking_running_stats) { // nn.cpp:321-327: running = running*momentum + stat*(1-momentum), unbiased variance
        const size_t F = (size_t)_num_features, N = x->shape()[0];
        auto mean = std::make_shared<tensor<float>>(std::vector<size_t>{1, F}, op->mean(), false);
        auto var = std::make_shared<tensor<float>>(std::vector<size_t>{1, F}, op->var(), false);
        const float unbias = N > 1 ? (float)N / (float)(N - 1) : 0.0f;
        _buffers["running_mean"] = (_buffers["running_mean"] * _momentum) + mean * (1.0f - _momentum);
        _buffers["running_var"] = (_buffers["running_var"] * _momentum) + var * (unbias * (1.0f - _momentum));
    }
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}

LayerNorm::LayerNorm(const size_t &normalized_shape, const float &eps, const bool &elementwise_affine, const bool &bias, const std::string &n)
    : Module(n), _normalized_shape(normalized_shape), _eps(eps), _elementwise_affine(elementwise_affine), _bias(bias) {
    std::vector<size_t> dims = {1, normalized_shape};
    if (elementwise_affine) {
        register_parameter("gammas", std::make_shared<tensor<float>>(dims, 1.0f, true));
        if (_bias) register_parameter("betas", std::make_shared<tensor<float>>(dims, 0.0f, true));
    }
    training = true;
}

tptr<float> LayerNorm::forward_relu(const tptr<float> &x, bool relu) {
    if (x->rank() != 2 || x->shape()[1] != _normalized_shape) throw std::runtime_error(err::size_mismatch());
    auto gamma = _elementwise_affine ? _parameters["gammas"] : nullptr;
    auto beta = (_elementwise_affine && _bias) ? _parameters["betas"] : nullptr;
    auto op = std::make_unique<LayerNormOp<tensor<float>>>();
    auto out = op->forward(x, gamma, beta, _eps, relu);
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}

Dropout::Dropout(const float &p, const std::string &n) : Module(n), p(p) {
    if (p > 1.0 || p < 0.0) throw std::runtime_error("invalid input, prob should be between 0 and 1 (inclusive)");
    training = true;
}

tptr<float> Dropout::forward(const tptr<float> &input_tensor) {
    if (!training || p == 0.0f) return input_tensor; // p = 0 keeps everything at scale 1
    if (p >= 1.0f) return input_tensor * 0.0f;
    auto op = std::make_unique<DropoutOp<tensor<float>>>();
    auto out = op->forward(input_tensor, p, seed + 0x9E3779B97F4A7C15ull * (++calls));
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}

tptr<float> tanh(const tptr<float> &x) {
    auto op = std::make_unique<TanhOp<tensor<float>>>();
    auto out = op->forward(x);
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}

MLP::MLP(size_t in_channel, std::vector<size_t> hid_dims, const bool &bias, const float &dropout) : Module("MLP") {
    auto seq = new Sequential();
    int i = 0;
    for (auto hid_dim : hid_dims) {
        auto i_s = std::to_string(i);
        seq->add_module("lin_" + i_s, new Linear(in_channel, hid_dim, bias));
        if (hid_dim != hid_dims[hid_dims.size() - 1]) {
            seq->add_module("lnorm_" + i_s, new LayerNorm(hid_dim));
            seq->add_module("relu_" + i_s, new ReLU());
        }
        seq->add_module("drop_" + i_s, new Dropout(dropout));
        in_channel = hid_dim;
        i++;
    }
    register_module("seq", seq);
}

tptr<float> MLP::forward(const tptr<float> &input) {
    // the chain of the registered children, with LayerNorm + ReLU fused into one kernel
    auto &children = _modules[0].second->_modules;
    auto out = input;
    for (size_t k = 0; k < children.size(); k++) {
        auto ln = std::dynamic_pointer_cast<LayerNorm>(children[k].second);
        if (ln && k + 1 < children.size() && std::dynamic_pointer_cast<ReLU>(children[k + 1].second)) {
            out = ln->forward_relu(out, true);
            k++;
            continue;
        }
        out = (*children[k].second)(out);
    }
    return out;
}

tptr<float> cross_entropy_loss(const tptr<float> logits, const tptr<int> target) {
    auto op = std::make_unique<SoftmaxCrossEntropy<tensor<float>>>();
    auto out = op->forward(logits, target);
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}

// node mask (0/1) -> device uint8 array + number of selected nodes
static device::buffer_ptr mask_to_u8(const tensor<bool> &mask, int64_t n_rows, int64_t *n_selected) {
    if ((int64_t)mask.numel() != n_rows)
        throw std::runtime_error("invalid input, mask must be 1D and of same size with num of nodes in graph");
    const std::valarray<bool> *h = mask.data();
    std::vector<uint8_t> u8(mask.numel());
    int64_t n = 0;
    for (size_t i = 0; i < u8.size(); i++) {
        u8[i] = (*h)[i] ? 1 : 0;
        n += u8[i];
    }
    auto buf = device::alloc(u8.size() ? u8.size() : 1);
    device::check(gnn_memcpy_h2d(device::ctx(), buf->ptr, u8.data(), u8.size()));
    device::check(gnn_ctx_sync(device::ctx())); // u8 is a stack-local staging buffer
    *n_selected = n;
    return buf;
}

tptr<float> cross_entropy_loss(const tptr<float> logits, const tptr<int> target, const tensor<bool> &mask) {
    int64_t n_sel = 0;
    auto m8 = mask_to_u8(mask, (int64_t)logits->shape()[0], &n_sel);
    if (n_sel == 0) throw std::runtime_error("invalid input, mask selects no node");
    auto op = std::make_unique<MaskedSoftmaxCrossEntropy<tensor<float>>>();
    auto out = op->forward(logits, target, m8, n_sel);
    if (out->requires_grad()) out->grad_fn = std::move(op);
    return out;
}

size_t count_correct(const tptr<float> logits, const tptr<int> target, const tensor<bool> *mask) {
    if (logits->rank() != 2 || target->rank() != 1 || target->numel() != logits->shape()[0])
        throw std::runtime_error("invalid input, logits must be of rank 2 and targets must be 1D tensor");
    device::buffer_ptr m8;
    int64_t n_sel = 0;
    if (mask) m8 = mask_to_u8(*mask, (int64_t)logits->shape()[0], &n_sel);
    auto cnt = device::alloc(8);
    device::check(gnn_argmax_correct(device::ctx(), (int64_t)logits->shape()[0], (int32_t)logits->shape()[1], logits->dptr(),
                                     (int64_t)logits->shape()[1], target->dptr(), m8 ? static_cast<const uint8_t *>(m8->ptr) : nullptr,
                                     static_cast<int64_t *>(cnt->ptr)));
    int64_t h = 0;
    device::check(gnn_memcpy_d2h(device::ctx(), &h, cnt->ptr, 8));
    return (size_t)h;
}

void Optimizer::zero_grad() {
    for (auto &p : _parameters) p->zero_grad();
}

void SGD::step() {
    if (_velocity.size() != _parameters.size()) _velocity.assign(_parameters.size(), nullptr);
    for (size_t i = 0; i < _parameters.size(); i++) {
        auto &p = _parameters[i];
        float *vel = nullptr;
        if (_momentum != 0.0f) {
            if (!_velocity[i]) _velocity[i] = device::alloc(p->numel() * 4);
            vel = static_cast<float *>(_velocity[i]->ptr);
        }
        device::check(gnn_sgd_step(device::ctx(), (int64_t)p->numel(), p->dptr(), p->grad_dptr(), vel, _lr, _momentum, _dampening,
                                   _weight_decay, _nestorov, _steps == 0));
    }
    _steps++;
}

Adam::Adam(std::vector<tptr<float>> parameters, float lr, float b1, float b2, float eps, float weight_decay)
    : Optimizer(std::move(parameters)), _lr(lr), _b1(b1), _b2(b2), _eps(eps), _weight_decay(weight_decay) {
    for (const auto &p : _parameters) { // zero-initialised moments like the reference constructor (nn.cpp:419-426)
        _velocity.push_back(device::alloc(p->numel() * 4));
        _momentum.push_back(device::alloc(p->numel() * 4));
        device::check(gnn_memset(device::ctx(), _velocity.back()->ptr, 0, p->numel() * 4));
        device::check(gnn_memset(device::ctx(), _momentum.back()->ptr, 0, p->numel() * 4));
    }
}

void Adam::step() {
    _steps++;
    for (size_t i = 0; i < _parameters.size(); i++) {
        auto &p = _parameters[i];
        device::check(gnn_adam_step(device::ctx(), (int64_t)p->numel(), p->dptr(), p->grad_dptr(),
                                    static_cast<float *>(_momentum[i]->ptr), static_cast<float *>(_velocity[i]->ptr), _lr, _b1,
                                    _b2, _eps, _weight_decay, (int64_t)_steps));
    }
}

} // namespace nn

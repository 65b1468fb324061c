// nn.h — Module tree, Linear, ReLU, Sequential, BatchNorm, softmax, cross-entropy loss and the optimisers of the GCN
// training loop, plus LayerNorm, Dropout, tanh and MLP of the reference's Model (counterpart of reference
// include/nn.h:28-214).  Out of scope here: Sigmoid, Softmax/LogSoftmax modules, Embedding.
#ifndef GNNB200_NN_H
#define GNNB200_NN_H

#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "tensor.h"

namespace nn {

/** parameter / submodule registry — reference nn.h:28-61, nn.cpp:12-151 */
class Module : public std::enable_shared_from_this<Module> {
  public:
    explicit Module(std::string n = "Module") : name(std::move(n)) {}
    virtual ~Module() = default; // the reference's destructor is non-virtual (bug B5)
    bool training = true;
    std::string name;
    void register_module(std::string name, Module *module);
    void register_parameter(std::string name, cyg::tptr<float> p);
    void register_buffer(std::string name, cyg::tptr<float> p);
    void zero_grad();
    void eval() { train(false); }
    void train(const bool &isTrain = true);
    cyg::tptr<float> get_parameter(std::string name);
    cyg::tptr<float> get_buffer(std::string name);
    std::shared_ptr<Module> get_module(std::string name);
    cyg::tptr<float> operator()(const cyg::tptr<float> &input_tensor, cyg::tensor<int> *y = nullptr);
    virtual cyg::tptr<float> forward(const cyg::tptr<float> &) { throw std::runtime_error("not implemented"); }
    virtual cyg::tptr<float> forward(const cyg::tptr<float> &, cyg::tensor<int> *) { throw std::runtime_error("not implemented"); }
    std::vector<std::shared_ptr<Module>> modules(const bool &recurse = true);
    std::unordered_map<std::string, std::shared_ptr<Module>> named_modules(const bool &recurse = true);
    std::vector<cyg::tptr<float>> parameters(const bool &recurse = true);
    std::unordered_map<std::string, cyg::tptr<float>> named_parameters(const bool &recurse = true);
    std::vector<cyg::tptr<float>> buffers(const bool &recurse = true) const;
    std::unordered_map<std::string, cyg::tptr<float>> named_buffers(const bool &recurse = true) const;

    std::vector<std::pair<std::string, std::shared_ptr<Module>>> _modules;
    std::unordered_map<std::string, cyg::tptr<float>> _parameters;
    std::unordered_map<std::string, cyg::tptr<float>> _buffers;
};

/** y = x W^T + b, W[out,in] — reference nn.h:63-73, nn.cpp:187-211.  One fused NT-GEMM node. */
class Linear : public Module {
  public:
    Linear(const size_t &in_features, const size_t &out_features, const bool &bias = true, const std::string &n = "Linear");
    void reset_parameters();
    cyg::tptr<float> forward(const cyg::tptr<float> &input_tensor) override;
    bool _bias;
    size_t _in_features, _out_features;
};

class Sequential : public Module { // reference nn.h:75-82
  public:
    explicit Sequential(const std::string &n = "seq") : Module(n) {}
    Sequential(std::vector<std::pair<std::string, Module *>> input, const std::string &n = "seq");
    void add_module(std::string n, Module *m) { register_module(n, m); }
    cyg::tptr<float> forward(const cyg::tptr<float> &input_tensor) override;
};

class ReLU : public Module { // reference nn.h:84-91, nn.cpp:229-237
  public:
    explicit ReLU(const std::string &n = "ReLU") : Module(n) {}
    cyg::tptr<float> forward(const cyg::tptr<float> &input_tensor) override;
};

/** batch normalisation over the node dimension — reference nn.h:114-123, nn.cpp:285-330: parameters "gammas" (1) and
 *  "betas" (0) of shape {1, F}, buffers "running_mean"/"running_var"; training mode normalises with the batch
 *  statistics (biased variance) and updates the running statistics as running = running*momentum + stat*(1-momentum)
 *  (unbiased variance for running_var), evaluation mode uses the running statistics.  forward_relu fuses the ReLU
 *  that graph::GCNConv applies right after (graph.cpp:174-175). */
class BatchNorm : public Module {
  public:
    BatchNorm(const size_t &num_features, const float &eps = 1e-05, const float &momentum = 0.1, const bool &affine = true,
              const bool &track_running_stats = true, const std::string &n = "BatchNorm");
    cyg::tptr<float> forward(const cyg::tptr<float> &x) override { return forward_relu(x, false); }
    cyg::tptr<float> forward_relu(const cyg::tptr<float> &x, bool relu);
    int _num_features;
    float _eps, _momentum;
    bool _affine, _tracking_running_stats;
};

/** reference nn.h:125-135, nn.cpp:332-353: parameters "gammas" (1) / "betas" (0) of shape {1, F}; forward_relu fuses the
 *  nn::ReLU that nn::MLP applies next */
class LayerNorm : public Module {
  public:
    LayerNorm(const size_t &normalized_shape, const float &eps = 1e-05, const bool &elementwise_affine = true, const bool &bias = true,
              const std::string &n = "LayerNorm");
    cyg::tptr<float> forward(const cyg::tptr<float> &x) override { return forward_relu(x, false); }
    cyg::tptr<float> forward_relu(const cyg::tptr<float> &x, bool relu);
    size_t _normalized_shape;
    float _eps;
    bool _elementwise_affine, _bias;
};

/** reference nn.h:93-103, nn.cpp:239-266: identity in evaluation mode; in training mode zeroes with probability p and
 *  scales the rest by 1/(1-p).  The mask is a function of (seed, call counter, element index) instead of the
 *  reference's time-seeded engine (bug B6): reproducible runs, fresh mask on every call. */
class Dropout : public Module {
  public:
    explicit Dropout(const float &p = 0.5, const std::string &n = "Dropout");
    cyg::tptr<float> forward(const cyg::tptr<float> &input_tensor) override;
    float p;
    uint64_t seed = 0x5eed, calls = 0;
};

/** (e^x - e^-x)/(e^x + e^-x) of x + 1e-12 — reference nn.cpp:355-364 — one fused node */
cyg::tptr<float> tanh(const cyg::tptr<float> &x);

/** reference nn.h:193-214: per hidden width Linear, then LayerNorm + ReLU unless the width equals the last width, then
 *  Dropout; children "lin_i", "lnorm_i", "relu_i", "drop_i" inside a Sequential registered as "seq".
 *  forward() runs that chain (the reference's forward cannot resolve "seq" through its own named_modules()). */
class MLP : public Module {
  public:
    MLP(size_t in_channel, std::vector<size_t> hid_dims, const bool &bias = true, const float &dropout = 0.0);
    cyg::tptr<float> forward(const cyg::tptr<float> &input) override;
};

/** softmax(x) = exp(x - log(sum(exp(x)))) composed from tensor ops like the reference (nn.cpp:270-278) */
cyg::tptr<float> softmax(const cyg::tptr<float> &input_tensor, int dim);

/** mean_i -log(exp(z_iy)/(sum_c exp(z_ic)+1e-20)) — reference nn.h:191, nn.cpp:442-453 — one fused kernel; its
 *  backward is the analytic (softmax - onehot)/N (the reference's own backward throws, bug B3). */
cyg::tptr<float> cross_entropy_loss(const cyg::tptr<float> logits, const cyg::tptr<int> target);

class Optimizer { // reference nn.h:155-162
  public:
    explicit Optimizer(std::vector<cyg::tptr<float>> parameters) : _parameters(std::move(parameters)) {}
    void zero_grad();
    std::vector<cyg::tptr<float>> _parameters;
};

/** torch.optim.SGD semantics (the documented intent, nn.h:165-167; the reference body segfaults, bug B4) */
class SGD : public Optimizer {
  public:
    SGD(std::vector<cyg::tptr<float>> parameters, float lr, float momentum = 0, float dampening = 0, float weight_decay = 0,
        bool nestorov = false)
        : Optimizer(std::move(parameters)), _lr(lr), _dampening(dampening), _momentum(momentum), _weight_decay(weight_decay), _nestorov(nestorov) {}
    void step();
    float _lr, _dampening, _momentum, _weight_decay;
    bool _nestorov;
    std::vector<cyg::device::buffer_ptr> _velocity;
    size_t _steps = 0;
};

/** torch.optim.Adam semantics — the intent of reference nn.h:180-188.  The reference body (nn.cpp:419-441) divides
 *  by sqrt(v)*eps, uses the parameter INDEX as the step count and its default `eps = 10 - 8` evaluates to 2; here
 *  eps defaults to 1e-8 and the step count is the number of step() calls. */
class Adam : public Optimizer {
  public:
    Adam(std::vector<cyg::tptr<float>> parameters, float lr, float b1 = 0.9f, float b2 = 0.999f, float eps = 1e-8f,
         float weight_decay = 0);
    void step();
    float _lr, _b1, _b2, _eps, _weight_decay;
    std::vector<cyg::device::buffer_ptr> _velocity, _momentum; // second / first moment (reference member names)
    size_t _steps = 0;
};

/** cross-entropy over the nodes selected by a mask (graph::Data::set_mask TRAIN/VAL/TEST, reference graph.cpp:130-151):
 *  mean over the selected rows; the gradient of unselected rows is zero. */
cyg::tptr<float> cross_entropy_loss(const cyg::tptr<float> logits, const cyg::tptr<int> target, const cyg::tensor<bool> &mask);
/** number of selected rows (all rows when mask == nullptr) whose arg-max logit equals the target
 *  (tensor::argmax semantics, reference tensor.h:645-648: first maximum) */
size_t count_correct(const cyg::tptr<float> logits, const cyg::tptr<int> target, const cyg::tensor<bool> *mask = nullptr);

} // namespace nn
#endif

// ref_driver.cpp — drives the UNMODIFIED arithmetic of walexi/gnn.cpp (patched only so it
// compiles, see oracle/build_ref.sh) to produce oracle outputs for the GCN hot path.
//
// TEST INFRASTRUCTURE ONLY: built into oracle/_ref/ref_gcn, used by tests/golden/make_golden.py
// to pin oracle/gcn_oracle.c, and by `bench.py --impl reference` as the timed CPU reference.
// Nothing here is on the product path.
//
// "Mode B" (SURVEY.md §0.1, §8c): the standard Kipf-Welling layer composed from the reference's
// own primitives —
//   A   = graph::edge_to_adj_mat(ei, nullptr, N)           src/graph.cpp:21-44
//   A->fill_diagonal_(1)                                   include/tensor.h:806-817
//   deg = A->sum(-1,true); dinv = deg->pow(-0.5)           src/graph.cpp:178,183
//   Ahat= (A*dinv)*dinv->t()                               functional.h:189-213 (broadcast mul)
//   Z_l = Ahat->mm(H->mm(W->t())) + b ; H = ReLU(Z)        nn.cpp:205-211, graph.cpp:208, nn.cpp:229-237
//   loss= nn::cross_entropy_loss(Z_L, y)                   nn.cpp:442-453
//   dZ_L= (nn::softmax(Z_L) - onehot(y))/N                 nn.cpp:270-278 (reference CE backward throws, bug B3)
//   Z_L->backward(dZ_L)                                    tensor.h:260-276, operation.h:504-534 ...
//
// usage:
//   ref_gcn structure <problem.gcnp> <out.gcno>
//   ref_gcn step      <problem.gcnp> <out.gcno>
//   ref_gcn structure_w <problem.gcnp> <out.gcno>    (weighted adjacency + normalisation, mode B with edge_attr)
//   ref_gcn mlp       <problem.gcnp> <out.gcno>      (nn::MLP forward + nn::tanh on the problem's X / W / b)
//   ref_gcn aswritten <problem.gcnp> <out.gcno>      (one graph::GCNConv::forward exactly as written)
//   ref_gcn time      <problem.gcnp> <steps>         (prints one JSON line with per-stage ms)
#include "graph.h"
#include "nn.h"
#include REF_NN_CPP

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

using cyg::tensor;
using cyg::tptr;

namespace {

struct Problem {
    int64_t N = 0, E = 0, L = 0;
    std::vector<int64_t> dims;
    std::vector<int> src, dst;
    std::vector<float> X;
    std::vector<int> y;
    std::vector<std::vector<float>> W, b;
};

void rd(std::ifstream &f, void *p, size_t n) {
    f.read(reinterpret_cast<char *>(p), n);
    if (!f) { fprintf(stderr, "ref_driver: short read\n"); exit(2); }
}

Problem load(const char *path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { fprintf(stderr, "ref_driver: cannot open %s\n", path); exit(2); }
    Problem p;
    int64_t magic;
    rd(f, &magic, 8);
    if (magic != 0x47434E50) { fprintf(stderr, "ref_driver: bad magic\n"); exit(2); }
    rd(f, &p.N, 8); rd(f, &p.E, 8); rd(f, &p.L, 8);
    p.dims.resize(p.L + 1);
    rd(f, p.dims.data(), 8 * (p.L + 1));
    p.src.resize(p.E); p.dst.resize(p.E);
    if (p.E) { rd(f, p.src.data(), 4 * p.E); rd(f, p.dst.data(), 4 * p.E); }
    p.X.resize(p.N * p.dims[0]); rd(f, p.X.data(), 4 * p.X.size());
    p.y.resize(p.N); rd(f, p.y.data(), 4 * p.N);
    p.W.resize(p.L); p.b.resize(p.L);
    for (int64_t l = 0; l < p.L; l++) {
        p.W[l].resize(p.dims[l + 1] * p.dims[l]); rd(f, p.W[l].data(), 4 * p.W[l].size());
        p.b[l].resize(p.dims[l + 1]);             rd(f, p.b[l].data(), 4 * p.b[l].size());
    }
    return p;
}

struct Writer {
    std::ofstream f;
    explicit Writer(const char *path) : f(path, std::ios::binary) {}
    void put(const std::string &name, int dtype, const std::vector<int64_t> &shape, const void *data, size_t bytes) {
        int32_t nl = name.size(), nd = shape.size(), dt = dtype;
        f.write(reinterpret_cast<char *>(&nl), 4); f.write(name.data(), nl);
        f.write(reinterpret_cast<char *>(&dt), 4); f.write(reinterpret_cast<char *>(&nd), 4);
        f.write(reinterpret_cast<const char *>(shape.data()), 8 * nd);
        f.write(reinterpret_cast<const char *>(data), bytes);
    }
    void f32(const std::string &name, const std::valarray<float> &v, std::vector<size_t> shape) {
        std::vector<int64_t> s(shape.begin(), shape.end());
        put(name, 0, s, &v[0], 4 * v.size());
    }
    void i32(const std::string &name, const std::valarray<int> &v, std::vector<size_t> shape) {
        std::vector<int64_t> s(shape.begin(), shape.end());
        put(name, 1, s, v.size() ? &v[0] : nullptr, 4 * v.size());
    }
};

tptr<float> make_f(const std::vector<float> &v, std::vector<size_t> dims, bool rg) {
    auto *d = new std::valarray<float>(v.data(), v.size());
    return std::make_shared<tensor<float>>(dims, d, rg);
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Ahat built exactly from reference primitives (mode B).
tptr<float> build_ahat(const Problem &p, tptr<float> *deg_out = nullptr, tptr<float> *dinv_out = nullptr) {
    auto ei = graph::vec_to_edge_list(p.src, p.dst);
    auto A = graph::edge_to_adj_mat(*ei, nullptr, p.N);
    A->fill_diagonal_(1);
    auto deg = A->sum(-1, true);
    auto dinv = deg->pow(-0.5);
    auto Ahat = (A * dinv) * dinv->t(-1, -2);
    if (deg_out) *deg_out = deg;
    if (dinv_out) *dinv_out = dinv;
    return Ahat;
}

int cmd_structure(const Problem &p, const char *out) {
    Writer w(out);
    auto ei = graph::vec_to_edge_list(p.src, p.dst);
    for (int fill = 0; fill <= 1; fill++) {
        // add_self_loops = edge_to_adj_mat + fill_diagonal_ + adj_to_edge_list  (src/graph.cpp:68-75)
        auto [el, ew] = graph::add_self_loops(*ei, nullptr, (float)fill, (int)p.N);
        w.i32(std::string("coo_fill") + char('0' + fill), *el->data(), el->shape());
    }
    tptr<float> deg, dinv;
    auto Ahat = build_ahat(p, &deg, &dinv);
    w.f32("deg", *deg->data(), {(size_t)p.N});
    w.f32("dinv", *dinv->data(), {(size_t)p.N});
    // nonzero values of Ahat in row-major order (positions == coo_fill1)
    auto &ad = *Ahat->data();
    std::vector<float> nz;
    for (size_t i = 0; i < ad.size(); i++) if (ad[i] != 0.0f) nz.push_back(ad[i]);
    std::valarray<float> nzv(nz.data(), nz.size());
    w.f32("ahat_val", nzv, {nz.size()});
    return 0;
}

struct StepTimes { double ahat = 0, fwd = 0, loss = 0, bwd = 0; };

int run_step(const Problem &p, Writer *w, StepTimes *tm) {
    double t0 = now_ms();
    auto Ahat = build_ahat(p);
    double t1 = now_ms();
    size_t N = p.N;
    std::vector<tptr<float>> W(p.L), B(p.L), Z(p.L);
    for (int64_t l = 0; l < p.L; l++) {
        W[l] = make_f(p.W[l], {(size_t)p.dims[l + 1], (size_t)p.dims[l]}, true);
        B[l] = make_f(p.b[l], {(size_t)p.dims[l + 1]}, true);
    }
    auto H = make_f(p.X, {N, (size_t)p.dims[0]}, false);
    nn::ReLU relu;
    for (int64_t l = 0; l < p.L; l++) {
        auto P = H->mm(W[l]->t(-1, -2));      // nn::Linear::forward without bias (nn.cpp:205-211)
        auto Y = Ahat->mm(P);                 // aggregation (graph.cpp:208)
        Z[l] = Y + B[l];                      // bias (graph.cpp:188)
        if (w) w->f32("Z" + std::to_string(l + 1), *Z[l]->data(), Z[l]->shape());
        if (l + 1 < p.L) H = relu.forward(Z[l]);
    }
    double t2 = now_ms();
    auto logits = Z[p.L - 1];
    auto *yd = new std::valarray<int>(p.y.data(), p.y.size());
    auto yt = std::make_shared<tensor<int>>(std::vector<size_t>{N}, yd, false);
    // loss forward on a detached copy: the reference's own CE backward is broken (bug B3) and
    // logits must keep fan-out 1 for the reference autograd to be a valid gradient oracle (bug B2).
    // (a requires_grad leaf clone: cross_entropy_loss dereferences out->grad_fn, nn.cpp:451)
    auto logits_detached = logits->clone(true);
    auto loss = nn::cross_entropy_loss(logits_detached, yt);
    auto S = nn::softmax(logits_detached, -1);
    size_t C = p.dims[p.L];
    auto *dz = new std::valarray<float>(*S->data());
    for (size_t i = 0; i < N; i++) (*dz)[i * C + p.y[i]] -= 1.0f;
    *dz /= (float)N;
    auto dZ = std::make_shared<tensor<float>>(std::vector<size_t>{N, C}, dz, false);
    double t3 = now_ms();
    logits->backward(dZ);
    double t4 = now_ms();
    if (w) {
        w->f32("loss", *loss->data(), {1});
        w->f32("dZ", *dZ->data(), dZ->shape());
        for (int64_t l = 0; l < p.L; l++) {
            w->f32("dW" + std::to_string(l + 1), *W[l]->grad(), W[l]->shape());
            w->f32("db" + std::to_string(l + 1), *B[l]->grad(), B[l]->shape());
        }
    }
    if (tm) { tm->ahat += t1 - t0; tm->fwd += t2 - t1; tm->loss += t3 - t2; tm->bwd += t4 - t3; }
    return 0;
}

// Weighted adjacency through the reference's own code: edge_to_adj_mat with edge_attr (src/graph.cpp:21-44, last write
// wins), fill_diagonal_(1), deg = rowsum, dinv = pow(deg,-0.5), Ahat = (A*dinv)*dinv^T (mode B with weights).
// Weights: w_i = 0.25 + ((7919 i) mod 1024)/512 (positive, exact in fp32; same formula as oracle.edge_weights).
int cmd_structure_w(const Problem &p, const char *out) {
    Writer w(out);
    auto ei = graph::vec_to_edge_list(p.src, p.dst);
    auto *wv = new std::valarray<float>(p.E);
    for (int64_t i = 0; i < p.E; i++) (*wv)[i] = 0.25f + (float)((i * 7919) % 1024) / 512.0f;
    tensor<float> attr(std::vector<size_t>{(size_t)p.E}, wv, false);
    auto A = graph::edge_to_adj_mat(*ei, &attr, p.N);
    A->fill_diagonal_(1);
    auto deg = A->sum(-1, true);
    auto dinv = deg->pow(-0.5);
    auto Ahat = (A * dinv) * dinv->t(-1, -2);
    w.f32("w_deg", *deg->data(), {(size_t)p.N});
    w.f32("w_dinv", *dinv->data(), {(size_t)p.N});
    auto &a0 = *A->data();
    auto &ad = *Ahat->data();
    std::vector<float> nz, raw;
    std::vector<int> rows, cols;
    for (size_t i = 0; i < ad.size(); i++)
        if (a0[i] != 0.0f) { nz.push_back(ad[i]); raw.push_back(a0[i]); rows.push_back((int)(i / p.N)); cols.push_back((int)(i % p.N)); }
    std::valarray<float> nzv(nz.data(), nz.size()), rawv(raw.data(), raw.size());
    std::valarray<int> rv(rows.data(), rows.size()), cv(cols.data(), cols.size());
    w.f32("w_ahat_val", nzv, {nz.size()});
    w.f32("w_raw_val", rawv, {raw.size()});
    w.i32("w_rows", rv, {rows.size()});
    w.i32("w_cols", cv, {cols.size()});
    return 0;
}

// nn::MLP (include/nn.h:193-214: Linear [+ LayerNorm + ReLU] + Dropout(p = 0) per layer) and nn::tanh (src/nn.cpp:355-364)
// through the reference's own modules.  The MLP has widths dims[0] -> dims[1] -> ... -> dims[L]; Linear weights/biases
// are the problem's W/b, LayerNorm gammas = 1 + b/2, betas = b/4 on the layers that have one.
int cmd_mlp(const Problem &p, const char *out) {
    Writer w(out);
    std::vector<size_t> hid;
    for (int64_t l = 1; l <= p.L; l++) hid.push_back((size_t)p.dims[l]);
    nn::MLP mlp((size_t)p.dims[0], hid, true, 0.0f);
    // nn::MLP::forward looks its Sequential up by name, which the reference's named_modules() cannot resolve (leaves are
    // keyed by their class name, src/nn.cpp:87-102), so the children are reached through the public _modules lists and
    // the Sequential's own forward (src/nn.cpp:219-227) is run — the same module chain MLP::forward was meant to run.
    auto seq = mlp._modules[0].second;
    auto child = [&](const std::string &n) -> std::shared_ptr<nn::Module> {
        for (auto &kv : seq->_modules) if (kv.first == n) return kv.second;
        fprintf(stderr, "ref_driver: no child %s\n", n.c_str()); exit(2);
    };
    for (int64_t l = 0; l < p.L; l++) {
        std::valarray<float> wv(p.W[l].data(), p.W[l].size()), bv(p.b[l].data(), p.b[l].size());
        auto lin = child("lin_" + std::to_string(l));
        lin->_parameters["weight"]->set_data(&wv);
        lin->_parameters["bias"]->set_data(&bv);
        if (p.dims[l + 1] != p.dims[p.L]) {
            std::valarray<float> gam = 1.0f + 0.5f * bv, bet = 0.25f * bv;
            auto ln = child("lnorm_" + std::to_string(l));
            ln->_parameters["gammas"]->set_data(&gam);
            ln->_parameters["betas"]->set_data(&bet);
        }
    }
    auto x = make_f(p.X, {(size_t)p.N, (size_t)p.dims[0]}, false);
    auto y = seq->forward(x);
    w.f32("mlp_out", *y->data(), y->shape());
    auto t = nn::tanh(y);
    w.f32("tanh_out", *t->data(), t->shape());
    return 0;
}

// The reference's graph::GCNConv layer EXACTLY AS WRITTEN (src/graph.cpp:160-212), run through its own forward():
//   add_self_loops(..., 0)  -> loops removed;  lin (no bias) -> BatchNorm (training statistics) -> ReLU;
//   deg = rowsum(A0) + 1; dinv = deg^-0.5; norm = (A0 dinv) * dinv;  out = (A0 h) * norm + bias
// Parameters injected: lin.weight = W1, conv bias = b1, BatchNorm gammas = 1 + b1/2, betas = b1/4 (so the affine part
// is exercised).  The registered Dropout module is not called by forward().  Only the forward is a valid oracle:
// BatchNorm uses its input several times and the reference autograd loses fan-out gradients (SURVEY.md bug B2).
int cmd_aswritten(const Problem &p, const char *out) {
    Writer w(out);
    const size_t N = p.N, Fi = p.dims[0], Fo = p.dims[1];
    auto ei = graph::vec_to_edge_list(p.src, p.dst);
    auto *ei_raw = new tensor<int>(ei->shape(), new std::valarray<int>(*ei->data()), false); // Data keeps the raw pointer
    auto x = make_f(p.X, {N, Fi}, false);
    graph::GCNConv conv(Fi, Fo, 0.0f);
    std::valarray<float> wv(p.W[0].data(), p.W[0].size()), bv(p.b[0].data(), p.b[0].size());
    std::valarray<float> gam = 1.0f + 0.5f * bv, bet = 0.25f * bv;
    conv.get_module("lin")->get_parameter("weight")->set_data(&wv);
    conv.get_parameter("bias")->set_data(&bv);
    conv.get_module("bnorm")->get_parameter("gammas")->set_data(&gam);
    conv.get_module("bnorm")->get_parameter("betas")->set_data(&bet);
    auto lin = (*conv.get_module("lin"))(x);
    w.f32("aw_lin", *lin->data(), lin->shape());
    auto bn = (*conv.get_module("bnorm"))(lin);
    w.f32("aw_bn", *bn->data(), bn->shape());
    auto Z = conv.forward(graph::Data(x, ei_raw));
    w.f32("aw_Z", *Z->data(), Z->shape());
    return 0;
}

} // namespace

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s structure|step|time <problem.gcnp> <out.gcno|steps>\n", argv[0]);
        return 2;
    }
    std::string cmd = argv[1];
    Problem p = load(argv[2]);
    if (cmd == "structure") return cmd_structure(p, argv[3]);
    if (cmd == "step") { Writer w(argv[3]); return run_step(p, &w, nullptr); }
    if (cmd == "aswritten") return cmd_aswritten(p, argv[3]);
    if (cmd == "structure_w") return cmd_structure_w(p, argv[3]);
    if (cmd == "mlp") return cmd_mlp(p, argv[3]);
    if (cmd == "time") {
        int steps = atoi(argv[3]);
        StepTimes tm;
        double t0 = now_ms();
        for (int s = 0; s < steps; s++) run_step(p, nullptr, &tm);
        double total = now_ms() - t0;
        printf("{\"steps\": %d, \"ms_per_step\": %.3f, \"ahat_ms\": %.3f, \"fwd_ms\": %.3f, \"loss_ms\": %.3f, \"bwd_ms\": %.3f, \"threads\": 1}\n",
               steps, total / steps, tm.ahat / steps, tm.fwd / steps, tm.loss / steps, tm.bwd / steps);
        return 0;
    }
    fprintf(stderr, "unknown command %s\n", cmd.c_str());
    return 2;
}

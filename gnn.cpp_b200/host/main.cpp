// main.cpp — the training driver the reference sketches but never wrote (reference src/main.cpp:10-36: a `Model`
// of GCNConv layers and an empty main()).  Builds the model from graph::GCNConv, runs the full-batch loop
//     opt.zero_grad(); logits = model.forward(data); loss = cross_entropy_loss(logits, y); loss->backward(); opt.step();
//  on the GPU through the reference-shaped C++ API, and optionally dumps activations / gradients
// for the parity tests.
//
//   gcn_main --config cora|pubmed|arxiv|reddit|products|tiny|tiny_pl [--epochs 5] [--lr 0.01]
//   gcn_main --problem file.gcnp [--epochs 1] [--lr 0] [--dump out.gcno]
//   gcn_main --config tiny_pl --model reference   (the reference's Model: pre/post MLP, GCNConv as written, tanh; Adam)
//   gcn_main --gpus N --config products            (one process per GPU: the launcher re-executes itself N times with
//                                                   RANK / WORLD_SIZE / LOCAL_RANK set; nodes are 1-D row-partitioned,
//                                                   graph::Data::partitioned holds the rank's rows, the SpMM node
//                                                   exchanges aggregation inputs, gradients are all-reduced)
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstring>
#include <iostream>

#include "graph.h"
#include "problem_io.h"
#include "synth.h"

using namespace cyg;
using namespace graph;
using namespace nn;

// Layer stack in the shape of the reference's Model (main.cpp:10-30): enc1..encL GCNConv, ReLU between layers
// (fused into the producing kernel's epilogue), logits out of the last one.
class Model : public MessagePassing {
  public:
    Model(size_t input_dim, std::vector<size_t> layer_dims) : MessagePassing() {
        for (size_t i = 0; i < layer_dims.size(); i++) {
            const bool last = i + 1 == layer_dims.size();
            register_module("enc" + std::to_string(i + 1), new GCNConv(input_dim, layer_dims[i], 0.0f, /*fused_relu=*/!last));
            input_dim = layer_dims[i];
        }
    }
    tptr<float> forward(const Data &data) {
        tptr<float> out = data.x();
        activations.clear();
        for (auto &kv : _modules) {
            auto *conv = static_cast<GCNConv *>(kv.second.get());
            out = conv->forward(data.with_x(out));
            activations.push_back(out);
        }
        return out;
    }
    std::vector<tptr<float>> activations;
};

// The reference's Model exactly in its own shape (src/main.cpp:10-30): "pre" MLP (F -> 2F -> F), one GCNConv as written per
// hidden width with nn::tanh after each, "post" MLP (H -> 2H -> H -> classes).  (The reference's forward hands the
// convolutions (tensor, edge_index) through Module::operator(), an overload GCNConv does not have; here they get the
// Data object their forward takes.)
class ReferenceModel : public MessagePassing {
  public:
    ReferenceModel(size_t n_classes, size_t input_dim, std::vector<size_t> hidden_dims, float p, bool bias) : MessagePassing() {
        register_module("pre", new MLP(input_dim, {input_dim * 2, input_dim}, bias, p));
        for (size_t i = 0; i < hidden_dims.size(); i++) {
            register_module("enc" + std::to_string(i + 1), new GCNConvAsWritten(input_dim, hidden_dims[i]));
            input_dim = hidden_dims[i];
        }
        register_module("post", new MLP(input_dim, {input_dim * 2, input_dim, n_classes}, bias, p));
    }
    tptr<float> forward(const Data &data) {
        auto out = (*_modules[0].second)(data.x());
        for (size_t i = 1; i + 1 < _modules.size(); i++) {
            out = static_cast<GCNConvAsWritten *>(_modules[i].second.get())->forward(data.with_x(out));
            out = nn::tanh(out);
        }
        return (*_modules.back().second)(out);
    }
};

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// `--gpus N` without RANK in the environment: start N copies of this binary, one per GPU, and wait for them
static int launch_ranks(int n, char **argv) {
    char rdv[] = "/tmp/gcn_main_rdv_XXXXXX";
    const int fd = mkstemp(rdv);
    if (fd >= 0) close(fd);
    unlink(rdv); // rank 0 creates it (atomically) once the NCCL id exists
    std::vector<pid_t> pids;
    for (int r = 0; r < n; r++) {
        const pid_t pid = fork();
        if (pid == 0) {
            setenv("RANK", std::to_string(r).c_str(), 1);
            setenv("LOCAL_RANK", std::to_string(r).c_str(), 1);
            setenv("WORLD_SIZE", std::to_string(n).c_str(), 1);
            setenv("GNN_RDV", rdv, 1);
            execv("/proc/self/exe", argv);
            std::perror("execv");
            _exit(127);
        }
        pids.push_back(pid);
    }
    int rc = 0;
    for (pid_t pid : pids) {
        int st = 0;
        waitpid(pid, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = 1;
    }
    unlink(rdv);
    return rc;
}

int main(int argc, char **argv) {
    std::string config, problem_path, dump_path, model_kind = "gcn";
    int epochs = 5, gpus = 1;
    float lr = 0.01f;
    for (int i = 1; i < argc; i++) {
        auto next = [&](const char *flag) -> const char * {
            if (i + 1 >= argc) { std::cerr << "missing value for " << flag << "\n"; std::exit(2); }
            return argv[++i];
        };
        if (!strcmp(argv[i], "--config")) config = next("--config");
        else if (!strcmp(argv[i], "--problem")) problem_path = next("--problem");
        else if (!strcmp(argv[i], "--dump")) dump_path = next("--dump");
        else if (!strcmp(argv[i], "--epochs")) epochs = std::atoi(next("--epochs"));
        else if (!strcmp(argv[i], "--lr")) lr = (float)std::atof(next("--lr"));
        else if (!strcmp(argv[i], "--model")) model_kind = next("--model"); // gcn (default) | reference
        else if (!strcmp(argv[i], "--gpus")) gpus = std::atoi(next("--gpus"));
        else { std::cerr << "unknown argument " << argv[i] << "\n"; return 2; }
    }
    if (gpus > 1 && !std::getenv("RANK")) return launch_ranks(gpus, argv);
    try {
        device::init_distributed(); // no-op for a single process
        const auto &ds = device::dist();
        const bool root = ds.rank == 0;
        problem_io::Problem p;
        if (!problem_path.empty()) {
            p = problem_io::load(problem_path);
        } else {
            synth::Config c;
            if (!synth::lookup(config.empty() ? "cora" : config, c)) { std::cerr << "unknown config\n"; return 2; }
            p.N = c.N; p.E = c.E; p.L = (int64_t)c.dims.size() - 1;
            p.dims.assign(c.dims.begin(), c.dims.end());
            synth::edges(c.seed(), c.E, c.N, c.powerlaw, p.src, p.dst);
            p.X.resize((size_t)c.N * c.dims[0]);
            synth::uniform(c.seed(), 3, p.X.size(), -1.0f, 1.0f, p.X.data());
            p.y.resize(c.N);
            for (int32_t i = 0; i < c.N; i++) p.y[i] = (int)(synth::hash3(c.seed(), 4, i) % (uint64_t)c.dims.back());
            p.W.resize(p.L); p.b.resize(p.L);
            for (int64_t l = 1; l <= p.L; l++) {
                const float bound = 1.0f / std::sqrt((float)c.dims[l - 1]);
                p.W[l - 1].resize((size_t)c.dims[l] * c.dims[l - 1]);
                p.b[l - 1].resize(c.dims[l]);
                synth::uniform(c.seed(), 16 + 2 * l, p.W[l - 1].size(), -bound, bound, p.W[l - 1].data());
                synth::uniform(c.seed(), 16 + 2 * l + 1, p.b[l - 1].size(), -bound, bound, p.b[l - 1].data());
            }
        }
        const size_t N = (size_t)p.N;
        // this rank's rows [lo, hi) of the node features and labels (all of them for a single process)
        const size_t chunk = (N + ds.world - 1) / ds.world;
        const size_t lo = std::min(N, (size_t)ds.rank * chunk), hi = std::min(N, (size_t)(ds.rank + 1) * chunk), n_loc = hi - lo;
        const size_t F0 = (size_t)p.dims[0];
        auto x = std::make_shared<tensor<float>>(std::vector<size_t>{n_loc, F0}, new std::valarray<float>(p.X.data() + lo * F0, n_loc * F0), false);
        auto y = std::make_shared<tensor<int>>(std::vector<size_t>{n_loc}, new std::valarray<int>(p.y.data() + lo, n_loc), false);
        auto edge_index = vec_to_edge_list(p.src, p.dst);
        double t0 = now_ms();
        Data data = ds.active() ? Data::partitioned(x, edge_index.get(), N, lo, hi) : Data(x, edge_index.get());
        device::sync();
        if (root)
            std::cout << "graph: N=" << N << " E=" << p.E << " ranks=" << ds.world << " nnz(A+I) of rank 0's rows=" << gnn_graph_nnz(data.structure()->g)
                      << " structure build " << now_ms() - t0 << " ms\n";

        std::vector<size_t> layer_dims(p.dims.begin() + 1, p.dims.end());
        if (model_kind == "reference") { // the reference's own Model shape, trained with Adam on seeded parameters
            seed_rng(1234);
            std::vector<size_t> hidden(layer_dims.begin(), layer_dims.end() - 1);
            ReferenceModel rm((size_t)p.dims.back(), (size_t)p.dims[0], hidden, 0.1f, true);
            Adam adam(rm.parameters(), lr);
            for (int e = 0; e < epochs; e++) {
                device::sync();
                t0 = now_ms();
                adam.zero_grad();
                auto out = rm.forward(data);
                auto l = cross_entropy_loss(out, y);
                l->backward();
                if (lr != 0.0f) adam.step();
                std::cout << "epoch " << e << " loss " << l->item() << " step " << now_ms() - t0 << " ms\n";
            }
            return 0;
        }
        Model model((size_t)p.dims[0], layer_dims);
        for (int64_t l = 1; l <= p.L; l++) { // inject the seeded parameters (reference init is time-seeded, bug B6)
            auto conv = model.get_module("enc" + std::to_string(l));
            std::valarray<float> W(p.W[l - 1].data(), p.W[l - 1].size()), b(p.b[l - 1].data(), p.b[l - 1].size());
            conv->get_module("lin")->_parameters["weight"]->set_data(&W);
            conv->_parameters["bias"]->set_data(&b);
        }
        SGD opt(model.parameters(), lr);
        tptr<float> logits, loss;
        for (int e = 0; e < epochs; e++) {
            device::sync();
            t0 = now_ms();
            opt.zero_grad();
            logits = model.forward(data);
            loss = cross_entropy_loss(logits, y);
            loss->backward();
            if (ds.active()) { // every rank holds its rows' share: sum the parameter gradients and the loss over the ranks
                for (auto &prm : model.parameters()) device::allreduce_sum(prm->grad_dptr(), (int64_t)prm->numel());
                device::allreduce_sum(loss->dptr(), 1);
            }
            if (lr != 0.0f) opt.step();
            const float l = loss->item(); // device -> host read of the scalar (synchronises)
            if (root) std::cout << "epoch " << e << " loss " << l << " step " << now_ms() - t0 << " ms\n";
        }
        if (!dump_path.empty() && root) { // (a partitioned run dumps rank 0's rows of every activation)
            problem_io::Writer w(dump_path);
            float l = loss->item();
            w.f32("loss", &l, {1});
            for (size_t i = 0; i < model.activations.size(); i++) {
                auto *h = model.activations[i]->data();
                w.f32("A" + std::to_string(i + 1), &(*h)[0], {(int64_t)n_loc, (int64_t)model.activations[i]->shape()[1]});
            }
            for (int64_t l2 = 1; l2 <= p.L; l2++) {
                auto conv = model.get_module("enc" + std::to_string(l2));
                auto Wt = conv->get_module("lin")->_parameters["weight"];
                auto bt = conv->_parameters["bias"];
                w.f32("dW" + std::to_string(l2), &(*Wt->grad())[0], {(int64_t)Wt->shape()[0], (int64_t)Wt->shape()[1]});
                w.f32("db" + std::to_string(l2), &(*bt->grad())[0], {(int64_t)bt->shape()[0]});
                w.f32("W" + std::to_string(l2), &(*Wt->data())[0], {(int64_t)Wt->shape()[0], (int64_t)Wt->shape()[1]});
            }
        }
    } catch (const std::exception &e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
    return 0;
}

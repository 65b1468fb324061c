// host_tests.cpp — unit tests of the reference-shaped C++ API, in the style of the reference's own
// tests/tensor.test.cpp, tests/operation.test.cpp, tests/nn.test.cpp and tests/graph.test.cpp (doctest is not
// available, so a few macros stand in).  Runs on the GPU; expectations are closed-form values computed inline.
// Exit code = number of failed checks.  Driven by tests/test_host_cpp.py (pytest -m gpu).
#include <cmath>
#include <functional>
#include <iostream>

#include "graph.h"

using namespace cyg;
using namespace graph;
using namespace nn;

static int g_failed = 0, g_checked = 0;
#define CHECK(cond)                                                                              \
    do {                                                                                         \
        g_checked++;                                                                             \
        if (!(cond)) { g_failed++; std::cerr << __FILE__ << ":" << __LINE__ << " CHECK failed: " #cond "\n"; } \
    } while (0)
#define CHECK_THROWS_WITH(expr, msg)                                                             \
    do {                                                                                         \
        g_checked++;                                                                             \
        bool thrown = false;                                                                     \
        try { expr; } catch (const std::exception &e) { thrown = std::string(e.what()).find(msg) != std::string::npos; } \
        if (!thrown) { g_failed++; std::cerr << __FILE__ << ":" << __LINE__ << " expected throw: " #expr "\n"; } \
    } while (0)

static bool close(float a, float b, float tol = 1e-5f) { return std::fabs(a - b) <= tol * std::max(1.0f, std::fabs(b)); }
static bool all_close(const std::valarray<float> &a, const std::valarray<float> &b, float tol = 1e-5f) {
    if (a.size() != b.size()) return false;
    float scale = std::max(1e-30f, std::abs(b).max());
    return (std::abs(a - b).max()) <= tol * scale;
}
static tptr<float> T(std::vector<size_t> dims, std::initializer_list<float> v, bool rg = false) {
    return std::make_shared<tensor<float>>(dims, new std::valarray<float>(v), rg);
}

static void test_tensor_basics() { // reference tests/tensor.test.cpp:17-70
    CHECK_THROWS_WITH(std::make_shared<tensor<int>>(std::vector<size_t>{2, 3}, 1, true), "Only Tensors of floating point dtype can require gradients");
    CHECK_THROWS_WITH(std::make_shared<tensor<float>>(std::vector<size_t>{2, 3}, new std::valarray<float>(1.0f, 5), false), "mismatch between number of elements");
    CHECK_THROWS_WITH(std::make_shared<tensor<float>>(std::vector<size_t>{}, 1.0f, false), "dims cannot be empty or zero");
    auto a = T({2, 3}, {1, 2, 3, 4, 5, 6});
    auto *d = a->data();
    CHECK((*d)[0] == 1 && (*d)[5] == 6 && a->numel() == 6 && a->rank() == 2);
    CHECK_THROWS_WITH(a->grad(), "pls enable grad");
    a->requires_grad_(true);
    CHECK(a->requires_grad() && a->grad()->sum() == 0.0f);
    CHECK((*a)(1, 2)->item() == 6.0f);
    CHECK_THROWS_WITH((*a)(2, 0), "out of bound range");
    auto b = T({1, 4}, {1, 2, 3, 4});
    b->squeeze();
    CHECK(b->rank() == 1 && b->shape()[0] == 4);
    b->unsqueeze(0);
    CHECK(b->rank() == 2 && b->shape()[0] == 1);
    auto c = a->clone();
    CHECK((*c->data() == *a->data()).min() == true && !c->requires_grad());
    auto e = eye(3);
    CHECK(e->data()->sum() == 3);
    auto r = randn<float>({4, 5}, -1.0f, 1.0f, true);
    CHECK(r->requires_grad() && r->data()->max() < 1.0f && r->data()->min() >= -1.0f);
}

static void test_elementwise_ops() { // reference tests/operation.test.cpp:32-161
    auto a = T({2, 3}, {1, 2, 3, 4, 5, 6}, true);
    auto b = T({3}, {10, 20, 30}, true);       // broadcast over rows
    auto c = T({2, 1}, {2, 4}, true);          // broadcast over columns
    auto out = (a + b) * c;                    // [2,3]
    std::valarray<float> expect = {22, 44, 66, 56, 100, 144};
    CHECK(all_close(*out->data(), expect));
    out->backward(std::make_shared<tensor<float>>(std::vector<size_t>{2, 3}, 1.0f, false));
    CHECK(all_close(*a->grad(), std::valarray<float>{2, 2, 2, 4, 4, 4}));
    CHECK(all_close(*b->grad(), std::valarray<float>{6, 6, 6}));            // sum over rows of c
    CHECK(all_close(*c->grad(), std::valarray<float>{11 + 22 + 33, 14 + 25 + 36}));
    // div, exp, log, sum, mean
    auto x = T({2, 2}, {1, 2, 3, 4}, true), y = T({2, 2}, {2, 4, 8, 16}, true);
    auto q = (x / y)->sum();
    q->backward();
    CHECK(close(q->item(), 0.5f + 0.5f + 0.375f + 0.25f));
    CHECK(all_close(*x->grad(), std::valarray<float>{0.5f, 0.25f, 0.125f, 0.0625f}));
    CHECK(all_close(*y->grad(), std::valarray<float>{-0.25f, -0.125f, -3.0f / 64, -4.0f / 256}));
    auto z = T({3}, {0.5f, 1.0f, 2.0f}, true);
    auto l = z->exp()->log()->mean();          // identity then mean
    l->backward();
    CHECK(close(l->item(), 3.5f / 3));
    CHECK(all_close(*z->grad(), std::valarray<float>{1.0f / 3, 1.0f / 3, 1.0f / 3}));
    auto m = T({2, 3}, {1, 2, 3, 4, 5, 6});
    CHECK(all_close(*m->sum(0)->data(), std::valarray<float>{5, 7, 9}));
    CHECK(all_close(*m->sum(-1, true)->data(), std::valarray<float>{6, 15}));
    CHECK(m->sum(-1, true)->shape() == (std::vector<size_t>{2, 1}));
    CHECK(all_close(*m->t()->data(), std::valarray<float>{1, 4, 2, 5, 3, 6}));
    CHECK_THROWS_WITH(m + T({2, 2}, {1, 1, 1, 1}), "tensors must be of same shape/size");
}

static void test_fan_out_accumulates() { // SURVEY.md bug B2: the reference returns dz/da = 4a here
    auto a = T({3}, {1, 2, 3}, true);
    auto y = a * 2.0f;
    auto z = (y * y)->sum();
    z->backward();
    CHECK(all_close(*a->grad(), std::valarray<float>{8, 16, 24}));
    z->backward(); // second pass over the same graph: nodes are done, like the reference it only warns
}

static void test_matmul_and_relu() { // reference tests/operation.test.cpp:219-234, tests/nn.test.cpp
    auto A = T({2, 3}, {1, 2, 3, 4, 5, 6}, true);
    auto B = T({3, 2}, {1, 0, 0, 1, 1, 1}, true);
    auto C = A->mm(B);
    CHECK(all_close(*C->data(), std::valarray<float>{4, 5, 10, 11}));
    C->backward(std::make_shared<tensor<float>>(std::vector<size_t>{2, 2}, 1.0f, false));
    CHECK(all_close(*A->grad(), std::valarray<float>{1, 1, 2, 1, 1, 2}));   // g * B^T
    CHECK(all_close(*B->grad(), std::valarray<float>{5, 5, 7, 7, 9, 9}));   // A^T * g (column sums of A)
    CHECK_THROWS_WITH(A->mm(A), "tensors are not compatible");
    auto x = T({2, 2}, {-1, 2, 0, 3}, true);
    ReLU relu;
    auto h = relu.forward(x);
    CHECK(all_close(*h->data(), std::valarray<float>{0, 2, 0, 3}));
    h->backward(std::make_shared<tensor<float>>(std::vector<size_t>{2, 2}, 5.0f, false));
    CHECK(all_close(*x->grad(), std::valarray<float>{0, 5, 0, 5}));         // strict > 0 (nn.cpp:231)
}

static void test_module_and_linear() { // reference tests/nn.test.cpp:19-31,84-115
    seed_rng(7);
    Linear lin(4, 3, true);
    CHECK(lin.parameters().size() == 2);
    CHECK(lin.get_parameter("weight")->shape() == (std::vector<size_t>{3, 4}));
    const float bound = 1.0f / std::sqrt(4.0f);
    CHECK(std::abs(*lin.get_parameter("weight")->data()).max() <= bound);
    CHECK_THROWS_WITH(lin.get_parameter("nope"), "no parameter with given name");
    lin.eval();
    CHECK(!lin.training && !lin.get_parameter("weight")->requires_grad());
    lin.train();
    CHECK(lin.get_parameter("weight")->requires_grad());
    std::valarray<float> W = {1, 0, 0, 0, 0, 1, 0, 0, 1, 1, 1, 1}, b = {0.5f, -0.5f, 0};
    lin.get_parameter("weight")->set_data(&W);
    lin.get_parameter("bias")->set_data(&b);
    auto x = T({2, 4}, {1, 2, 3, 4, 5, 6, 7, 8});
    auto y = lin.forward(x);
    CHECK(all_close(*y->data(), std::valarray<float>{1.5f, 1.5f, 10, 5.5f, 5.5f, 26}));
    y->backward(std::make_shared<tensor<float>>(std::vector<size_t>{2, 3}, 1.0f, false));
    CHECK(all_close(*lin.get_parameter("weight")->grad(), std::valarray<float>{6, 8, 10, 12, 6, 8, 10, 12, 6, 8, 10, 12})); // rows = column sums of x
    CHECK(all_close(*lin.get_parameter("bias")->grad(), std::valarray<float>{2, 2, 2}));
    auto *seq = new Sequential({{"l1", new Linear(4, 8)}, {"act", new ReLU()}, {"l2", new Linear(8, 2)}});
    std::shared_ptr<Module> holder(seq);
    CHECK(seq->parameters().size() == 4 && seq->forward(x)->shape() == (std::vector<size_t>{2, 2}));
    CHECK(seq->named_parameters().size() == 4); // l2's names collide with l1's and get the `l2_` prefix (nn.cpp:110-125)
}

static void test_loss_and_sgd() { // reference nn.cpp:442-453 (forward), nn.h:165-167 (SGD intent)
    auto Z = T({2, 3}, {1, 2, 3, 1, 1, 1}, true);
    auto y = std::make_shared<tensor<int>>(std::vector<size_t>{2}, new std::valarray<int>{2, 0}, false);
    auto loss = cross_entropy_loss(Z, y);
    const float e1 = std::exp(1.f), e2 = std::exp(2.f), e3 = std::exp(3.f);
    const float expect = 0.5f * (-std::log(e3 / (e1 + e2 + e3)) - std::log(1.0f / 3));
    CHECK(close(loss->item(), expect));
    loss->backward();
    const float s = e1 + e2 + e3;
    CHECK(all_close(*Z->grad(), std::valarray<float>{e1 / s / 2, e2 / s / 2, (e3 / s - 1) / 2, (1.f / 3 - 1) / 2, 1.f / 6, 1.f / 6}));
    // the composed softmax of the reference API gives the same probabilities
    auto sm = softmax(T({2, 3}, {1, 2, 3, 1, 1, 1}), -1);
    CHECK(all_close(*sm->data(), std::valarray<float>{e1 / s, e2 / s, e3 / s, 1.f / 3, 1.f / 3, 1.f / 3}));
    CHECK_THROWS_WITH(cross_entropy_loss(T({3}, {1, 2, 3}), y), "logits must be of rank 2");
    SGD opt({Z}, 0.1f);
    std::valarray<float> before = *Z->data(), g = *Z->grad();
    opt.step();
    CHECK(all_close(*Z->data(), std::valarray<float>(before - 0.1f * g)));
    opt.zero_grad();
    CHECK(Z->grad()->sum() == 0.0f);
    auto p = T({2}, {1, 1}, true);
    SGD mom({p}, 0.1f, 0.9f);
    for (int i = 0; i < 2; i++) { mom.zero_grad(); (p * 2.0f)->sum()->backward(); mom.step(); }
    // torch: v1 = g = 2 -> p = 0.8 ; v2 = 0.9*2 + 2 = 3.8 -> p = 0.42
    CHECK(all_close(*p->data(), std::valarray<float>{0.42f, 0.42f}));
}

static void test_adam_masks_accuracy() { // nn::Adam intent (nn.h:180-188), Data::set_mask use
    auto p = T({2}, {1, 1}, true);
    Adam opt({p}, 0.1f);
    for (int i = 0; i < 2; i++) { opt.zero_grad(); (p * 2.0f)->sum()->backward(); opt.step(); }
    // torch.optim.Adam, constant gradient 2: every step moves by lr * 1/(1 + eps/2) -> 0.9, 0.8
    CHECK(all_close(*p->data(), std::valarray<float>{0.8f, 0.8f}));
    auto Z = T({3, 3}, {1, 2, 3, 1, 1, 1, 5, 0, 5}, true);
    auto y = std::make_shared<tensor<int>>(std::vector<size_t>{3}, new std::valarray<int>{2, 0, 0}, false);
    tensor<bool> mask(std::vector<size_t>{3}, new std::valarray<bool>{true, false, true}, false);
    auto loss = cross_entropy_loss(Z, y, mask);
    const float e1 = std::exp(1.f), e2 = std::exp(2.f), e3 = std::exp(3.f), e5 = std::exp(5.f);
    const float expect = 0.5f * (-std::log(e3 / (e1 + e2 + e3)) - std::log(e5 / (2 * e5 + 1)));
    CHECK(close(loss->item(), expect));
    loss->backward();
    std::valarray<float> g = *Z->grad();
    CHECK(g[3] == 0.0f && g[4] == 0.0f && g[5] == 0.0f); // unselected row
    CHECK(close(g[2], (e3 / (e1 + e2 + e3) - 1) / 2));
    CHECK(count_correct(Z, y) == 3);        // rows: argmax 2 (==2), 0 (first max of a tie, ==0), 0 (first of the 5,5 tie, ==0)
    CHECK(count_correct(Z, y, &mask) == 2);
    tensor<bool> bad(std::vector<size_t>{2}, new std::valarray<bool>{true, false}, false);
    CHECK_THROWS_WITH(cross_entropy_loss(Z, y, bad), "mask must be 1D and of same size");
}

static void test_graph_structure() { // reference tests/graph.test.cpp:16-43 (its toy graph; it asserts nothing)
    auto edge_list = vec_to_edge_list({1, 2, 3, 0, 4, 1, 2, 3}, {1, 2, 0, 1, 2, 2, 1, 1});
    CHECK(edge_list->shape() == (std::vector<size_t>{2, 8}));
    CHECK_THROWS_WITH(vec_to_edge_list({1, 2}, {1}), "input vectors must be of same length");
    auto mat = edge_to_adj_mat(*edge_list); // n_nodes inferred as max+1 = 5 (bug B1 fixed)
    CHECK(mat->shape() == (std::vector<size_t>{5, 5}) && mat->data()->sum() == 8.0f);
    CHECK((*mat)(3, 0)->item() == 1.0f && (*mat)(0, 3)->item() == 0.0f); // row = source, col = destination
    auto [el, ew] = adj_to_edge_list(*mat);
    std::valarray<int> want = {0, 1, 1, 2, 2, 3, 3, 4, 1, 1, 2, 1, 2, 0, 1, 2};
    CHECK((*el->data() == want).min() == true && ew->data()->sum() == 8.0f);
    // edge weights land at A[src][dst] (graph.cpp:38-40); duplicates (last write wins) are covered by the directed golden
    auto attr = T({8}, {0.5f, 1.5f, 2.0f, 3.0f, 4.0f, 5.0f, 6.0f, 7.0f});
    auto wmat = edge_to_adj_mat(*edge_list, attr.get(), 5);
    CHECK((*wmat)(1, 1)->item() == 0.5f && (*wmat)(3, 0)->item() == 2.0f && (*wmat)(3, 1)->item() == 7.0f && (*wmat)(2, 1)->item() == 6.0f);
    CHECK(close(wmat->data()->sum(), 0.5f + 1.5f + 2.0f + 3.0f + 4.0f + 5.0f + 6.0f + 7.0f));
    auto [e0, w0] = add_self_loops(*edge_list, nullptr, 0, 5); // fillValue 0 REMOVES loops (graph.cpp:68-75)
    std::valarray<int> want0 = {0, 1, 2, 3, 3, 4, 1, 2, 1, 0, 1, 2};
    CHECK((*e0->data() == want0).min() == true);
    auto [e1, w1] = add_self_loops(*edge_list, nullptr, 1, 5);
    CHECK(e1->shape()[1] == 11);
    // with weights the reference's dense round trip keeps entries with int(w) != 0 only (graph.cpp:54): 0.5 (edge 1->1)
    // is the overwritten diagonal anyway; every row gets the diagonal fillValue
    auto [e2, w2] = add_self_loops(*edge_list, attr.get(), 1, 5);
    CHECK(e2->shape()[1] == 11 && w2->shape()[0] == 11);
    {   // rows 0..4 in order: (0,0)=1 (0,1)=3 | (1,1)=1 (1,2)=5 | (2,1)=6 (2,2)=1 | (3,0)=2 (3,1)=7 (3,3)=1 | (4,2)=4 (4,4)=1
        std::valarray<int> want2 = {0, 0, 1, 1, 2, 2, 3, 3, 3, 4, 4, 0, 1, 1, 2, 1, 2, 0, 1, 3, 2, 4};
        std::valarray<float> wv = {1, 3, 1, 5, 6, 1, 2, 7, 1, 4, 1};
        CHECK((*e2->data() == want2).min() == true && all_close(*w2->data(), wv));
    }
    auto small_w = T({8}, {0.5f, 0.25f, 2.0f, 3.0f, 0.75f, 5.0f, 6.0f, 7.0f});     // |w| < 1 entries vanish like in the reference
    auto [e3, w3] = add_self_loops(*edge_list, small_w.get(), 0, 5);
    CHECK(e3->shape()[1] == 5 && close(w3->data()->sum(), 2.0f + 3.0f + 5.0f + 6.0f + 7.0f));
    auto x = randn<float>({15, 10}, 0.0f, 2.0f, true);
    {   // a weighted Data feeds the WEIGHTED normalised adjacency to the layer: same as the dense composition
        auto xs = randn<float>({5, 3}, 0.0f, 1.0f, false);
        auto attr2 = T({8, 1}, {0.5f, 1.5f, 2.0f, 3.0f, 4.0f, 5.0f, 6.0f, 7.0f});
        Data wd(xs, edge_list.get(), attr2);
        auto A = edge_to_adj_mat(*edge_list, attr.get(), 5);
        A->fill_diagonal_(1);
        auto deg = A->sum(-1, true);
        cyg::tensor<float> mhalf(std::vector<size_t>{1}, -0.5f, false);
        auto dinv = functional::pow(*deg, mhalf);
        auto Ahat = (A * dinv) * dinv->t();
        auto agg = std::make_unique<SpMM<tensor<float>>>();
        auto got = agg->forward(wd.structure(), xs);
        auto want_y = Ahat->mm(xs);
        CHECK(all_close(*got->data(), *want_y->data()));
        Data ud(xs, edge_list.get());
        auto agg2 = std::make_unique<SpMM<tensor<float>>>();
        CHECK(!all_close(*agg2->forward(ud.structure(), xs)->data(), *want_y->data()));   // the unweighted structure differs
    }
    Data data(x, edge_list.get());
    CHECK(data.num_nodes() == 15 && data.num_node_features() == 10 && data.num_edges() == 8);
    CHECK(data.to_adj()->shape() == (std::vector<size_t>{15, 15}));
    GCNConv m(10, 20);
    auto out = m(data);
    CHECK(out->shape() == (std::vector<size_t>{15, 20}));
    auto small = randn<float>({3, 4}, 0.0f, 1.0f, false);
    CHECK_THROWS_WITH(Data(small, edge_list.get()), "max value in edge_index should be less than the number of nodes");
}

// GCNConv (sparse device structure, fused nodes) against the SAME layer composed from dense tensor primitives the
// way the reference composes it (mode B): A_hat = (A*dinv)*dinv^T ; Z = A_hat.mm(x.mm(W^T)) + b.
static void check_gcn_against_dense(size_t in, size_t out_ch) {
    seed_rng(11 + in);
    const size_t N = 40;
    std::vector<int> src, dst;
    for (size_t i = 0; i < N; i++)
        for (size_t j : {(i * 7 + 3) % N, (i * 13 + 5) % N, i}) { src.push_back((int)i); dst.push_back((int)j); src.push_back((int)j); dst.push_back((int)i); }
    auto ei = vec_to_edge_list(src, dst);
    auto x = randn<float>({N, in}, -1.0f, 1.0f, false);
    Data data(x, ei.get());
    GCNConv conv(in, out_ch);
    auto W = conv.get_module("lin")->_parameters["weight"];
    auto b = conv._parameters["bias"];
    b->uniform(-0.5f, 0.5f);
    auto Z = conv.forward(data.with_x(x));
    auto G = randn<float>({N, out_ch}, -1.0f, 1.0f, false);
    Z->backward(G);
    std::valarray<float> dW = *W->grad(), db = *b->grad();
    // dense composition
    auto A = edge_to_adj_mat(*ei, nullptr, N);
    A->fill_diagonal_(1);
    auto deg = A->sum(-1, true);
    cyg::tensor<float> mhalf(std::vector<size_t>{1}, -0.5f, false);
    auto dinv = functional::pow(*deg, mhalf);
    auto Ahat = (A * dinv) * dinv->t();
    auto W2 = W->clone(true), b2 = b->clone(true);
    auto Zd = Ahat->mm(x->mm(W2->t())) + b2;
    Zd->backward(G);
    CHECK(all_close(*Z->data(), *Zd->data(), 2e-5f));
    CHECK(all_close(dW, *W2->grad(), 2e-5f));
    CHECK(all_close(db, *b2->grad(), 2e-5f));
}

// graph::GCNConv exactly as written in the reference (graph.cpp:160-212) vs the same layer composed from dense tensor
// ops, gradients included (the composed graph exercises the fan-out accumulation the reference loses, bug B2)
static void check_gcn_as_written_against_dense() {
    seed_rng(23);
    const size_t N = 40, in = 9, out_ch = 6;
    std::vector<int> src, dst;
    for (size_t i = 0; i < N; i++)
        for (size_t j : {(i * 7 + 3) % N, (i * 13 + 5) % N, i}) { src.push_back((int)i); dst.push_back((int)j); src.push_back((int)j); dst.push_back((int)i); }
    auto ei = vec_to_edge_list(src, dst);
    auto x = randn<float>({N, in}, -1.0f, 1.0f, false);
    Data data(x, ei.get());
    GCNConvAsWritten conv(in, out_ch);
    auto W = conv.get_module("lin")->_parameters["weight"];
    auto b = conv._parameters["bias"];
    auto bn = conv.get_module("bnorm");
    auto gam = bn->_parameters["gammas"], bet = bn->_parameters["betas"];
    b->uniform(-0.5f, 0.5f);
    gam->uniform(0.5f, 1.5f);
    bet->uniform(-0.3f, 0.3f);
    auto Z = conv.forward(data.with_x(x));
    auto G = randn<float>({N, out_ch}, -1.0f, 1.0f, false);
    Z->backward(G);
    std::valarray<float> dW = *W->grad(), db = *b->grad(), dgam = *gam->grad(), dbet = *bet->grad();
    // dense composition with the reference's formulas
    auto A = edge_to_adj_mat(*ei, nullptr, N);
    A->fill_diagonal_(0); // add_self_loops(.., 0) removes the loops (graph.cpp:172)
    auto deg = A->sum(-1, true) + 1.0f;
    cyg::tensor<float> mhalf(std::vector<size_t>{1}, -0.5f, false);
    auto dinv = functional::pow(*deg, mhalf);
    auto norm = A->mm(dinv) * dinv;
    auto W2 = W->clone(true), b2 = b->clone(true), g2 = gam->clone(true), be2 = bet->clone(true);
    auto lin = x->mm(W2->t());
    auto mean = lin->mean(-2, true);
    auto cen = lin - mean;
    auto var = (cen * cen)->mean(-2, true);
    auto stdv = ((var + 1e-5f)->log() * 0.5f)->exp();
    auto y = (cen / stdv) * g2 + be2;
    auto h = y->where(y > 0.0f, 0.0f);
    auto Zd = A->mm(h) * norm + b2;
    Zd->backward(G);
    CHECK(all_close(*Z->data(), *Zd->data(), 5e-5f));
    CHECK(all_close(dW, *W2->grad(), 5e-5f));
    CHECK(all_close(db, *b2->grad(), 5e-5f));
    CHECK(all_close(dgam, *g2->grad(), 5e-5f));
    CHECK(all_close(dbet, *be2->grad(), 5e-5f));
    // running statistics moved towards the batch statistics (nn.cpp:321-327), evaluation mode uses them
    CHECK(std::abs(bn->_buffers["running_mean"]->data()->sum()) > 0.0f);
    conv.eval();
    auto Ze = conv.forward(data.with_x(x));
    CHECK(Ze->shape() == Z->shape());
}

// nn::MLP / LayerNorm / tanh / Dropout (the reference Model's pre/post processing, main.cpp:10-30) vs the same chain
// composed from tensor ops, gradients included
static void test_mlp_layernorm_tanh_dropout() {
    seed_rng(31);
    const size_t N = 33, in = 7;
    MLP mlp(in, {14, 7, 3}, true, 0.0f);
    auto x = randn<float>({N, in}, -1.0f, 1.0f, false);
    auto &ch = mlp._modules[0].second->_modules;
    CHECK(ch.size() == 3 * 2 + 2 * 2); // lin + drop for 3 layers, lnorm + relu for the two widths != 3
    auto ln0 = ch[1].second;
    ln0->_parameters["gammas"]->uniform(0.5f, 1.5f);
    ln0->_parameters["betas"]->uniform(-0.3f, 0.3f);
    auto y = nn::tanh(mlp.forward(x));
    auto G = randn<float>({N, 3}, -1.0f, 1.0f, false);
    y->backward(G);
    // composition from tensor ops
    auto h = x;
    std::vector<tptr<float>> Ws, bs, gs, bes;
    for (size_t k = 0; k < ch.size(); k++) {
        if (auto lin = std::dynamic_pointer_cast<Linear>(ch[k].second)) {
            auto W = lin->_parameters["weight"]->clone(true), b = lin->_parameters["bias"]->clone(true);
            Ws.push_back(W); bs.push_back(b);
            h = h->mm(W->t()) + b;
        } else if (auto ln = std::dynamic_pointer_cast<LayerNorm>(ch[k].second)) {
            auto g2 = ln->_parameters["gammas"]->clone(true), be2 = ln->_parameters["betas"]->clone(true);
            gs.push_back(g2); bes.push_back(be2);
            auto mean = h->mean(-1, true);
            auto cen = h - mean;
            auto var = (cen * cen)->mean(-1, true);
            auto stdv = ((var + 1e-5f)->log() * 0.5f)->exp();
            h = (cen / stdv) * g2 + be2;
            h = h->where(h > 0.0f, 0.0f);
        }
    }
    auto e2 = (h * 2.0f)->exp();
    auto yd = (e2 - 1.0f) / (e2 + 1.0f); // tanh
    yd->backward(G);
    CHECK(all_close(*y->data(), *yd->data(), 5e-5f));
    auto lin0 = std::dynamic_pointer_cast<Linear>(ch[0].second);
    CHECK(all_close(*lin0->_parameters["weight"]->grad(), *Ws[0]->grad(), 1e-4f));
    CHECK(all_close(*lin0->_parameters["bias"]->grad(), *bs[0]->grad(), 1e-4f));
    CHECK(all_close(*ln0->_parameters["gammas"]->grad(), *gs[0]->grad(), 1e-4f));
    CHECK(all_close(*ln0->_parameters["betas"]->grad(), *bes[0]->grad(), 1e-4f));
    // dropout: training mode zeroes ~p and scales by 1/(1-p); the gradient passes through the same mask; eval = identity
    Dropout drop(0.25f);
    auto big = std::make_shared<tensor<float>>(std::vector<size_t>{200, 100}, 1.0f, true);
    auto d = drop.forward(big);
    std::valarray<float> dv = *d->data();
    size_t nz = 0;
    for (float v : dv) nz += v != 0.0f;
    const float kept = (float)nz / 20000.0f;
    CHECK(std::fabs(kept - 0.75f) < 0.02f && close(dv.max(), 1.0f / 0.75f));
    d->sum()->backward();
    CHECK(all_close(*big->grad(), dv));
    drop.eval();
    CHECK(drop.forward(big).get() == big.get());
    CHECK_THROWS_WITH(Dropout(1.5f), "prob should be between 0 and 1");
}

int main() {
    const std::pair<const char *, std::function<void()>> tests[] = {
        {"tensor_basics", test_tensor_basics},       {"elementwise_ops", test_elementwise_ops},
        {"fan_out_accumulates", test_fan_out_accumulates}, {"matmul_and_relu", test_matmul_and_relu},
        {"module_and_linear", test_module_and_linear}, {"loss_and_sgd", test_loss_and_sgd},
        {"adam_masks_accuracy", test_adam_masks_accuracy},
        {"graph_structure", test_graph_structure},
        {"gcn_vs_dense_transform_first", [] { check_gcn_against_dense(12, 5); }},
        {"gcn_vs_dense_aggregate_first", [] { check_gcn_against_dense(6, 17); }},
        {"gcn_as_written_vs_dense", check_gcn_as_written_against_dense},
        {"mlp_layernorm_tanh_dropout", test_mlp_layernorm_tanh_dropout},
    };
    for (auto &t : tests) {
        const int before = g_failed;
        try {
            t.second();
        } catch (const std::exception &e) {
            g_failed++;
            std::cerr << "test " << t.first << " threw: " << e.what() << "\n";
        }
        std::cout << (g_failed == before ? "[ ok ] " : "[FAIL] ") << t.first << "\n";
    }
    std::cout << g_checked << " checks, " << g_failed << " failed\n";
    return g_failed;
}

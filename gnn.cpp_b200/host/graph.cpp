// graph.cpp — see graph.h.  Structure work is done by the device kernels K1-K3 behind gnn_graph_build /
// gnn_graph_build_csc / gnn_graph_normalize; nothing here touches an N x N matrix unless the caller explicitly
// asks for the dense form (to_adj / edge_to_adj_mat).
#include "graph.h"

using namespace cyg;

namespace graph {

tptr<int> vec_to_edge_list(std::vector<int> source, std::vector<int> destination) {
    if (source.size() != destination.size()) throw std::runtime_error("input vectors must be of same length");
    const size_t E = source.size();
    auto *h = new std::valarray<int>(2 * E);
    for (size_t i = 0; i < E; i++) {
        (*h)[i] = source[i];
        (*h)[E + i] = destination[i];
    }
    return std::make_shared<tensor<int>>(std::vector<size_t>{2, E}, h, false);
}

device::graph_ptr build_structure(const tensor<int> &edge_index, size_t num_nodes, int fill_mode, bool normalize) {
    if (edge_index.rank() != 2 || edge_index.shape()[0] != 2) throw std::runtime_error("invalid input for x, must be of 2D");
    const int64_t E = (int64_t)edge_index.shape()[1];
    gnn_graph_t *g = nullptr;
    // row 0 of the [2,E] device tensor is the source array, row 1 the destination array
    device::check(gnn_graph_build(device::ctx(), edge_index.dptr(), edge_index.dptr() + E, E, (int32_t)num_nodes, fill_mode, &g));
    auto h = std::make_shared<device::GraphHandle>(g);
    device::check(gnn_graph_build_csc(device::ctx(), g));
    if (normalize) device::check(gnn_graph_normalize(device::ctx(), g));
    return h;
}

static size_t infer_nodes(const tensor<int> &edge_index, size_t n_nodes) {
    if (n_nodes != 0) return n_nodes;
    return (size_t)std::get<0>(edge_index.max())->item() + 1;
}

tptr<float> edge_to_adj_mat(const tensor<int> &edge_index, tensor<float> *edge_attr, size_t n_nodes) {
    const size_t N = infer_nodes(edge_index, n_nodes);
    if (edge_attr != nullptr) { // A[src][dst] = w, last write wins (graph.cpp:38-40): weighted device build, diagonal as given
        if (edge_index.shape()[1] != edge_attr->shape()[0])
            throw std::runtime_error("invalid inputs, number of edges in edge_index must be equal to size of edge_attr");
        const int64_t E = (int64_t)edge_index.shape()[1];
        gnn_graph_t *g = nullptr;
        device::check(gnn_graph_build_weighted(device::ctx(), edge_index.dptr(), edge_index.dptr() + E, edge_attr->dptr(), E, (int32_t)N,
                                               /*fill_mode=*/2, &g));
        auto h = std::make_shared<device::GraphHandle>(g);
        auto out = std::make_shared<tensor<float>>(std::vector<size_t>{N, N}, 0.0f, false);
        device::check(gnn_graph_to_dense(device::ctx(), g, /*raw weights*/ 2, out->dptr(), (int64_t)N));
        return out;
    }
    auto s = build_structure(edge_index, N, /*fill_mode=*/2, /*normalize=*/false);
    auto out = std::make_shared<tensor<float>>(std::vector<size_t>{N, N}, 0.0f, false);
    device::check(gnn_graph_to_dense(device::ctx(), s->g, 0, out->dptr(), (int64_t)N));
    return out;
}

static std::tuple<tptr<int>, tptr<float>> coo_from_device(const int32_t *rows_d, const int32_t *cols_d, const float *vals_d, size_t n) {
    auto ei = std::make_shared<tensor<int>>(std::vector<size_t>{2, n ? n : 1}, 0, false);
    auto ew = std::make_shared<tensor<float>>(std::vector<size_t>{n ? n : 1}, 1.0f, false);
    if (n) {
        device::check(gnn_memcpy_d2d(device::ctx(), ei->dptr(), rows_d, n * 4));
        device::check(gnn_memcpy_d2d(device::ctx(), ei->dptr() + n, cols_d, n * 4));
        if (vals_d) device::check(gnn_memcpy_d2d(device::ctx(), ew->dptr(), vals_d, n * 4));
    }
    return {ei, ew};
}

std::tuple<tptr<int>, tptr<float>> adj_to_edge_list(tensor<float> &adj) {
    if (adj.rank() != 2) throw std::runtime_error("all dimensions must be of same length and tensor must be 2D");
    const int64_t R = adj.shape()[0], C = adj.shape()[1];
    int64_t count = 0;
    device::check(gnn_dense_to_coo(device::ctx(), adj.dptr(), R, C, C, nullptr, nullptr, nullptr, 0, &count));
    auto tmp = device::alloc((size_t)(count ? count : 1) * 12);
    int32_t *rows = static_cast<int32_t *>(tmp->ptr), *cols = rows + count;
    float *vals = reinterpret_cast<float *>(cols + count);
    if (count) device::check(gnn_dense_to_coo(device::ctx(), adj.dptr(), R, C, C, rows, cols, vals, count, &count));
    return coo_from_device(rows, cols, vals, (size_t)count);
}

std::tuple<tptr<int>, tptr<float>> add_self_loops(const tensor<int> &edge_index, tensor<float> *edge_attr, const float &fillValue, const int &num_nodes) {
    const size_t N = infer_nodes(edge_index, (size_t)num_nodes);
    if (edge_attr != nullptr) {
        // the reference's dense round trip with weights (graph.cpp:68-75 over :21-67): A[src][dst] = w (last write wins),
        // the whole diagonal := fillValue, then every entry with int(a) != 0 in row-major order — so weights with
        // |w| < 1 disappear, exactly as they do in the reference
        if (edge_index.shape()[1] != edge_attr->numel())
            throw std::runtime_error("invalid inputs, number of edges in edge_index must be equal to size of edge_attr");
        const int64_t E = (int64_t)edge_index.shape()[1];
        gnn_graph_t *g = nullptr;
        device::check(gnn_graph_build_weighted(device::ctx(), edge_index.dptr(), edge_index.dptr() + E, edge_attr->dptr(), E, (int32_t)N,
                                               /*fill_mode=*/2, &g));
        device::GraphHandle h(g);
        const size_t nnz = (size_t)gnn_graph_nnz(g);
        std::vector<int32_t> rowptr(N + 1), col(nnz ? nnz : 1);
        std::vector<float> w(nnz ? nnz : 1);
        device::check(gnn_graph_export_h(device::ctx(), g, rowptr.data(), col.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
        device::check(gnn_graph_export_weights_h(device::ctx(), g, w.data()));
        std::vector<int> rs, cs;
        std::vector<float> vs;
        const bool keep_diag = (int)fillValue != 0;
        for (size_t r = 0; r < N; r++) {
            bool diag_done = !keep_diag;
            for (int32_t k = rowptr[r]; k < rowptr[r + 1]; k++) {
                const size_t c = (size_t)col[k];
                if (!diag_done && c >= r) { rs.push_back((int)r); cs.push_back((int)r); vs.push_back(fillValue); diag_done = true; }
                if (c == r) continue;                       // the stored diagonal was overwritten by fill_diagonal_
                if ((int)w[k] != 0) { rs.push_back((int)r); cs.push_back((int)c); vs.push_back(w[k]); }
            }
            if (!diag_done) { rs.push_back((int)r); cs.push_back((int)r); vs.push_back(fillValue); }
        }
        const size_t n = rs.size();
        if (n == 0) return coo_from_device(nullptr, nullptr, nullptr, 0);
        auto *eh = new std::valarray<int>(2 * n);
        auto *wh = new std::valarray<float>(n);
        for (size_t i = 0; i < n; i++) { (*eh)[i] = rs[i]; (*eh)[n + i] = cs[i]; (*wh)[i] = vs[i]; }
        return {std::make_shared<tensor<int>>(std::vector<size_t>{2, n}, eh, false), std::make_shared<tensor<float>>(std::vector<size_t>{n}, wh, false)};
    }
    const int fill_mode = ((int)fillValue != 0) ? 1 : 0; // adj_to_edge_list keeps int(a) != 0 (graph.cpp:54)
    gnn_graph_t *g = nullptr;
    const int64_t E = (int64_t)edge_index.shape()[1];
    device::check(gnn_graph_build(device::ctx(), edge_index.dptr(), edge_index.dptr() + E, E, (int32_t)N, fill_mode, &g));
    device::GraphHandle h(g);
    const size_t nnz = (size_t)gnn_graph_nnz(g);
    std::vector<int32_t> rowptr(N + 1), col(nnz ? nnz : 1);
    device::check(gnn_graph_export_h(device::ctx(), g, rowptr.data(), col.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
    auto *eh = new std::valarray<int>(2 * nnz);
    auto *wh = new std::valarray<float>(1.0f, nnz);
    for (size_t r = 0; r < N; r++)
        for (int32_t k = rowptr[r]; k < rowptr[r + 1]; k++) {
            (*eh)[k] = (int)r;
            (*eh)[nnz + k] = col[k];
            if ((size_t)col[k] == r && fill_mode == 1) (*wh)[k] = fillValue;
        }
    if (nnz == 0) { delete eh; delete wh; return coo_from_device(nullptr, nullptr, nullptr, 0); }
    return {std::make_shared<tensor<int>>(std::vector<size_t>{2, nnz}, eh, false), std::make_shared<tensor<float>>(std::vector<size_t>{nnz}, wh, false)};
}

Data::Data(const tptr<float> &x, tensor<int> *edge_index, tptr<float> edge_attr, tensor<float> *y)
    : _num_nodes(x->shape()[0]), _num_node_features(x->rank() > 1 ? x->shape()[1] : 1), _edge_index(edge_index), _y(y), _x(x), _edge_attr(edge_attr) {
    if (x->rank() != 2) throw std::runtime_error("invalid input for x, must be 2D");
    if (edge_index != nullptr) {
        if (edge_index->rank() != 2 || edge_index->shape()[0] != 2) throw std::runtime_error("invalid input for x, must be of 2D");
        _num_edges = edge_index->shape()[1];
        if (edge_attr != nullptr) {
            if (edge_attr->rank() != 2) throw std::runtime_error("pls check input tensors, must of 2D for x, edge_index and edge_attr");
            if (edge_index->shape()[1] != edge_attr->shape()[0])
                throw std::runtime_error("invalid edge_index and/or edge_attr input, edge_index should of [2, num_edges] and edge_attr should be of [num_edges, num_edge_feature]");
            _num_edge_features = edge_attr->shape()[1];
        }
        // building the device structure validates max(edge_index) < num_nodes with the reference's message (graph.cpp:87-88)
        structure();
    }
}
Data Data::partitioned(const tptr<float> &x_local, tensor<int> *edge_index, size_t n_global, size_t lo, size_t hi) {
    if (x_local->rank() != 2 || x_local->shape()[0] != hi - lo || hi > n_global || lo >= hi)
        throw std::runtime_error("invalid input for x, must be 2D");
    if (edge_index == nullptr || edge_index->rank() != 2 || edge_index->shape()[0] != 2) throw std::runtime_error("invalid input for x, must be of 2D");
    Data d;
    d._x = x_local;
    d._num_nodes = hi - lo;
    d._num_node_features = x_local->shape()[1];
    d._edge_index = edge_index;
    d._num_edges = edge_index->shape()[1];
    d._n_global = n_global; d._part_lo = lo; d._part_hi = hi;
    auto &ds = device::dist();
    ds.n_global = (int64_t)n_global; ds.lo = (int64_t)lo; ds.hi = (int64_t)hi;
    ds.chunk = ((int64_t)n_global + ds.world - 1) / ds.world;
    d.structure();
    return d;
}
tensor<int> *Data::edge_index() {
    if (_edge_index == nullptr) throw std::runtime_error("pls provide adj matr or edge");
    return _edge_index;
}
void Data::set_edge_index(tensor<int> *edge_index, tptr<float> edge_attr) {
    _edge_index = edge_index; // ownership stays with the caller (the reference deletes the old pointer, graph.cpp:112-117)
    _edge_attr = edge_attr;
    _num_edges = edge_index ? edge_index->shape()[1] : 0;
    _structure.reset();
    _structure_as_written.reset();
}
tptr<float> Data::to_adj() {
    if (_edge_index == nullptr) throw std::runtime_error("pls provide adj matr or edge");
    return edge_to_adj_mat(*_edge_index, _edge_attr.get(), _num_nodes); // reference graph.cpp:118-129 passes _edge_attr too
}
void Data::set_mask(tensor<bool> &mask, DataType type) {
    if (mask.numel() != _num_nodes) throw std::runtime_error("invalid input, mask must be 1D and of same size with num of nodes in graph");
    switch (type) {
    case TRAIN: _train_mask = &mask; break;
    case VAL: _val_mask = &mask; break;
    case TEST: _test_mask = &mask; break;
    }
}
device::graph_ptr Data::structure() const {
    if (_edge_index == nullptr) throw std::runtime_error("pls provide adj matr or edge");
    if (!_structure) {
        if (_n_global) { // row block [lo, hi) of the global structure (forward CSR rows + backward CSC rows), global column ids
            auto full = build_structure(*_edge_index, _n_global, /*fill_mode=*/1, /*normalize=*/true);
            gnn_graph_t *l = nullptr;
            device::check(gnn_graph_slice_rows(device::ctx(), full->g, (int64_t)_part_lo, (int64_t)_part_hi, &l));
            _structure = std::make_shared<device::GraphHandle>(l);
        } else if (_edge_attr != nullptr) {
            // edge weights (edge_to_adj_mat with edge_attr, reference src/graph.cpp:21-44): weighted A0, last write wins,
            // unit diagonal, then the weighted-degree normalisation
            if (_edge_attr->numel() != _num_edges) throw std::runtime_error("invalid inputs, number of edges in edge_index must be equal to size of edge_attr");
            const int64_t E = (int64_t)_num_edges;
            gnn_graph_t *g = nullptr;
            device::check(gnn_graph_build_weighted(device::ctx(), _edge_index->dptr(), _edge_index->dptr() + E, _edge_attr->dptr(), E,
                                                   (int32_t)_num_nodes, /*fill_mode=*/1, &g));
            _structure = std::make_shared<device::GraphHandle>(g);
            device::check(gnn_graph_build_csc(device::ctx(), g));
            device::check(gnn_graph_normalize(device::ctx(), g));
        } else {
            _structure = build_structure(*_edge_index, _num_nodes, /*fill_mode=*/1, /*normalize=*/true);
        }
    }
    return _structure;
}

device::graph_ptr Data::structure_as_written() const {
    if (_edge_index == nullptr) throw std::runtime_error("pls provide adj matr or edge");
    if (!_structure_as_written) {
        _structure_as_written = build_structure(*_edge_index, _num_nodes, /*fill_mode=*/0, /*normalize=*/false);
        device::check(gnn_graph_normalize_as_written(device::ctx(), _structure_as_written->g, nullptr));
    }
    return _structure_as_written;
}

tptr<float> MessagePassing::propagate(const tensor<int> &edge_index, const tptr<float> &x, const tptr<float> *norm) {
    return aggregate_and_update(x, edge_index, norm);
}

GCNConv::GCNConv(size_t in_channels, size_t out_channels, float dropout, bool fused_relu)
    : MessagePassing(), _in_channels(in_channels), _out_channels(out_channels), _dropout(dropout), _fused_relu(fused_relu) {
    register_module("lin", new nn::Linear(in_channels, out_channels, false));
    register_parameter("bias", std::make_shared<tensor<float>>(std::vector<size_t>{out_channels}, 0.0f, true));
}

tptr<float> GCNConv::forward(Data &&input) {
    auto structure = input.structure();
    auto x = input.x();
    auto W = get_module("lin")->_parameters["weight"];
    auto b = _parameters["bias"];
    if (_in_channels < _out_channels) { // aggregate first (narrower): Z = (A_hat X) W^T + b
        auto agg = std::make_unique<SpMM<tensor<float>>>();
        auto M = agg->forward(structure, x);
        if (M->requires_grad()) M->grad_fn = std::move(agg);
        auto lin = std::make_unique<LinearOp<tensor<float>>>();
        auto out = lin->forward(M, W, b, _fused_relu);
        if (out->requires_grad()) out->grad_fn = std::move(lin);
        return out;
    }
    auto lin = std::make_unique<LinearOp<tensor<float>>>(); // transform first: Z = A_hat (X W^T) + b
    auto P = lin->forward(x, W, nullptr, false);
    if (P->requires_grad()) P->grad_fn = std::move(lin);
    auto agg = std::make_unique<SpMM<tensor<float>>>();
    auto out = agg->forward(structure, P, b, _fused_relu);
    if (out->requires_grad()) out->grad_fn = std::move(agg);
    return out;
}

GCNConvAsWritten::GCNConvAsWritten(size_t in_channels, size_t out_channels, float dropout)
    : MessagePassing(), _in_channels(in_channels), _out_channels(out_channels), _dropout(dropout) {
    register_module("lin", new nn::Linear(in_channels, out_channels, false));
    register_module("bnorm", new nn::BatchNorm(out_channels));
    register_module("relu", new nn::ReLU());
    register_parameter("bias", std::make_shared<tensor<float>>(std::vector<size_t>{out_channels}, 0.0f, true));
}

tptr<float> GCNConvAsWritten::forward(Data &&input) {
    auto structure = input.structure_as_written();
    auto out = (*get_module("lin"))(input.x());
    out = std::static_pointer_cast<nn::BatchNorm>(get_module("bnorm"))->forward_relu(out, true); // bnorm + relu, one kernel
    auto agg = std::make_unique<SpMM<tensor<float>>>();
    auto Z = agg->forward(structure, out, _parameters["bias"], false); // (A0 h) * norm + bias: values hold norm[row]
    if (Z->requires_grad()) Z->grad_fn = std::move(agg);
    return Z;
}

tptr<float> GCNConv::propagate(const tensor<int> &edge_index, const tptr<float> &x, const tptr<float> *others) {
    return aggregate_and_update(x, edge_index, others);
}

tptr<float> GCNConv::aggregate_and_update(const tptr<float> &x, const tensor<int> &edge_index, const tptr<float> *other) {
    auto s = build_structure(edge_index, x->shape()[0], /*fill_mode=*/2, /*normalize=*/false);
    auto agg = std::make_unique<SpMM<tensor<float>>>();
    auto out = agg->forward(s, x, nullptr, false, /*use_values=*/false);
    if (out->requires_grad()) out->grad_fn = std::move(agg);
    if (other != nullptr && *other != nullptr) out = out * *other;
    return out;
}

} // namespace graph

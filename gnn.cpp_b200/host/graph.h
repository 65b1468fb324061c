// graph.h — graph container, structure helpers, MessagePassing and the GCN layer (counterpart of reference
// include/graph.h:9-138, src/graph.cpp).  The reference has NO sparse format: every aggregation rebuilds a dense
// N x N adjacency from the edge list (graph.cpp:21-44,177,207).  Here graph::Data owns a device-resident
// structure — CSR + CSC of A_hat = D^-1/2 (A0 + I) D^-1/2 with its degree normalisation — built once on the GPU.
#ifndef GNNB200_GRAPH_H
#define GNNB200_GRAPH_H

#include <memory>
#include <tuple>
#include <vector>

#include "nn.h"
#include "tensor.h"

namespace graph {

typedef enum DataType { TRAIN, VAL, TEST } DataType;

/** two index vectors -> tensor<int>[2, E] (row 0 = source, row 1 = destination). reference graph.cpp:10-19 */
cyg::tptr<int> vec_to_edge_list(std::vector<int> source, std::vector<int> destination);
/** dense 0/1 adjacency from the edge list (row = edge_index[0], col = edge_index[1], duplicates collapse).
 *  reference graph.cpp:21-44; n_nodes == 0 uses max(edge_index) + 1 (the reference uses max and overflows, bug B1).
 *  API fidelity for small N: the layer never calls this. */
cyg::tptr<float> edge_to_adj_mat(const cyg::tensor<int> &edge_index, cyg::tensor<float> *edge_attr = nullptr, size_t n_nodes = 0);
/** dense -> row-major sorted (edge_index[2,nnz], values[nnz]) of entries with int(a) != 0. reference graph.cpp:46-67 */
std::tuple<cyg::tptr<int>, cyg::tptr<float>> adj_to_edge_list(cyg::tensor<float> &adj_mat);
/** sorted, de-duplicated edge list with the diagonal set to fillValue (0 removes loops, non-zero adds them).
 *  reference graph.cpp:68-75 — computed on the device from the edge list, no dense round trip. */
std::tuple<cyg::tptr<int>, cyg::tptr<float>> add_self_loops(const cyg::tensor<int> &edge_index, cyg::tensor<float> *edge_attr = nullptr,
                                                             const float &fillValue = 0, const int &num_nodes = 0);
/** device structure (CSR/CSC/normalisation) of A_hat for an edge list; fill_mode as in gnn_graph_build */
cyg::device::graph_ptr build_structure(const cyg::tensor<int> &edge_index, size_t num_nodes, int fill_mode, bool normalize);

/** graph data holder — reference graph.h:50-100, graph.cpp:77-151 */
class Data {
  public:
    Data() {}
    Data(const cyg::tptr<float> &x, cyg::tensor<int> *edge_index = nullptr, cyg::tptr<float> edge_attr = nullptr, cyg::tensor<float> *y = nullptr);
    /** one process per GPU: this rank's rows [lo, hi) of the node features of an N-node graph, the WHOLE edge list (every
     *  rank builds the structure on its GPU and keeps its row block of A_hat and of A_hat^T with global column ids).
     *  Layers then exchange aggregation inputs between the ranks (SpMM node, operation.h); reference layer being
     *  sharded: src/graph.cpp:170-212. */
    static Data partitioned(const cyg::tptr<float> &x_local, cyg::tensor<int> *edge_index, size_t n_global, size_t lo, size_t hi);
    cyg::tensor<int> *edge_index();
    void set_edge_index(cyg::tensor<int> *edge_index, cyg::tptr<float> edge_attr = nullptr);
    cyg::tptr<float> to_adj();
    size_t num_nodes() const { return _num_nodes; }
    size_t num_node_features() const { return _num_node_features; }
    size_t num_edges() const { return _num_edges; }
    size_t num_edge_features() const { return _num_edge_features; }
    cyg::tptr<float> x() const { return _x; }
    cyg::tensor<int> *edge_index() const { return _edge_index; }
    cyg::tptr<float> edge_attr() const { return _edge_attr; }
    void set_mask(cyg::tensor<bool> &mask, DataType type = DataType::TRAIN);
    /** cached device CSR/CSC/normalisation of A_hat (built on first use, then reused by every layer and step) */
    cyg::device::graph_ptr structure() const;
    /** same graph (shares the cached device structure), other node features: what a layer stack feeds layer l+1 */
    Data with_x(const cyg::tptr<float> &x) const {
        Data d(*this);
        if (x->rank() != 2 || x->shape()[0] != _num_nodes) throw std::runtime_error("invalid input for x, must be 2D");
        d._x = x;
        d._num_node_features = x->shape()[1];
        return d;
    }

  protected:
    cyg::tensor<bool> *_train_mask = nullptr, *_val_mask = nullptr, *_test_mask = nullptr;
    size_t _num_nodes = 0, _num_node_features = 0, _num_edges = 0, _num_edge_features = 0;
    cyg::tensor<int> *_edge_index = nullptr;
    cyg::tensor<float> *_y = nullptr;
    cyg::tptr<float> _x, _edge_attr;
    mutable cyg::device::graph_ptr _structure, _structure_as_written;
    size_t _n_global = 0, _part_lo = 0, _part_hi = 0; // row partition (Data::partitioned); 0 = whole graph
  public:
    /** loop-free structure with the factorised normalisation of the reference's GCNConv::forward as written
     *  (values norm[row], graph.cpp:172-185), cached like structure() */
    cyg::device::graph_ptr structure_as_written() const;
};

class MessagePassing : public nn::Module { // reference graph.h:110-120
  public:
    MessagePassing() {}
    virtual cyg::tptr<float> message(const cyg::tptr<float> *, const cyg::tptr<float> *x_j, const cyg::tptr<float> * = nullptr) { return *x_j; }
    virtual cyg::tptr<float> aggregate_and_update(const cyg::tptr<float> &, const cyg::tensor<int> &, const cyg::tptr<float> *) {
        throw std::runtime_error("not yet implemented");
    }
    template <typename... T> cyg::tptr<float> operator()(T &...input) { return forward(std::forward<T>(input)...); }
    using nn::Module::forward;
    virtual cyg::tptr<float> forward(Data &&) { throw std::runtime_error("not yet implemented"); }
    virtual cyg::tptr<float> forward(Data &&, Data &) { throw std::runtime_error("not yet implemented"); }
    virtual cyg::tptr<float> propagate(const cyg::tensor<int> &edge_index, const cyg::tptr<float> &x, const cyg::tptr<float> *norm = nullptr);
};

/** Kipf-Welling layer Z = A_hat (X W^T) + b  (the north-star formula; see DESIGN.md for the reference's as-written
 *  variant).  Aggregation runs at min(in, out) width: A_hat (X W^T) == (A_hat X) W^T.  reference graph.h:123-138 */
class GCNConv : public MessagePassing {
  public:
    GCNConv(size_t in_channels, size_t out_channels, float dropout = 0.0, bool fused_relu = false);
    cyg::tptr<float> forward(Data &&input) override;
    cyg::tptr<float> propagate(const cyg::tensor<int> &edge_index, const cyg::tptr<float> &x, const cyg::tptr<float> *others) override;
    /** plain sum over neighbours (times *other when given): reference graph.cpp:204-212, as a device SpMM */
    cyg::tptr<float> aggregate_and_update(const cyg::tptr<float> &x, const cyg::tensor<int> &edge_index, const cyg::tptr<float> *other) override;
    size_t _in_channels, _out_channels;
    float _dropout;
    bool _fused_relu;
};

/** graph::GCNConv EXACTLY AS WRITTEN in the reference (graph.cpp:160-212): self loops removed, Linear without bias ->
 *  BatchNorm (training statistics) -> ReLU, then (A0 h) * norm + bias with deg = rowsum(A0)+1, norm = (A0 dinv) * dinv.
 *  Same registered children as the reference ("lin", "bnorm", "drop", "relu", parameter "bias"); like the reference's
 *  forward, the Dropout child is registered but not applied.  The north-star layer is graph::GCNConv above. */
class GCNConvAsWritten : public MessagePassing {
  public:
    GCNConvAsWritten(size_t in_channels, size_t out_channels, float dropout = 0.0);
    cyg::tptr<float> forward(Data &&input) override;
    size_t _in_channels, _out_channels;
    float _dropout;
};

} // namespace graph
#endif

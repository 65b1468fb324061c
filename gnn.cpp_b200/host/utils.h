// utils.h — shape checks, error strings and seeded RNG for the device-backed mirror of the reference API.
//
// Counterpart of reference include/utils.h + src/utils.cpp.  The reference's stride/offset helpers
// (generate_idxs, repeat_nd, broadcast: utils.cpp:96-115, utils.h:160-228) exist only to drive std::valarray
// slices on the CPU; here broadcasting is expressed as strides handed to the kernels (gnn_binary_f32), so
// those helpers have no counterpart.  Error conditions throw std::runtime_error with the reference's
// messages (utils.h:19-30).
#ifndef GNNB200_UTILS_H
#define GNNB200_UTILS_H

#include <cstddef>
#include <cstdint>
#include <iostream>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

namespace cyg {
namespace err {
// same wording as the reference so that tests matching on messages keep working
inline const char *grad_dtype() { return "Only Tensors of floating point dtype can require gradients"; }
inline const char *grad_not_leaf() {
    return "UserWarning: The .grad attribute of a Tensor that is not a leaf Tensor is being accessed. Its .grad "
           "attribute won't be populated during autograd.backward()";
}
inline const char *in_place_leaf() { return "RuntimeError: a leaf Variable that requires grad is being used in an in-place operation."; }
inline const char *size_mismatch() { return "tensors must be of same shape/size - mismatch between number of elements and dimension of tensor"; }
inline const char *out_of_range() { return "out of bound range"; }
inline const char *invalid_dims() { return "dims cannot be empty or zero"; }
inline const char *non_scalar_backprop() { return "pass in tensor to backprop on non-scalar tensor"; }
inline const char *mm_compatible() { return "tensors are not compatible, tensors should of shape [...,A,B] and [...,B,A]"; }
inline const char *bad_dim() { return "dim is out of range"; }
inline const char *grad_mismatch() { return "size mismatch, incoming gradient must be same dimension with tensor"; }
inline const char *transpose() { return "invalid inp"; }
inline const char *rank_limit() { return "the device path supports rank-1 and rank-2 tensors (the GCN hot path); higher ranks are out of scope"; }
} // namespace err

using dims_t = std::vector<size_t>;

inline size_t count_elements(const dims_t &d) {
    return std::accumulate(d.begin(), d.end(), (size_t)1, std::multiplies<size_t>());
}
inline void check_valid_dims(const dims_t &d) {
    if (d.empty()) throw std::runtime_error(err::invalid_dims());
    for (auto v : d)
        if (v < 1) throw std::runtime_error(err::invalid_dims());
    if (d.size() > 2) throw std::runtime_error(err::rank_limit());
}
// a dim index valid for `rank` (negative counts from the end); INT32_MAX means "all"
constexpr int ALL_DIMS = INT32_MAX;
inline void check_dim(int dim, int rank) {
    if (dim != ALL_DIMS && (dim >= rank || dim < -rank)) throw std::runtime_error(err::bad_dim());
}
// view every tensor as [rows, cols]: rank-1 [n] is a single row
inline void as_2d(const dims_t &d, size_t &rows, size_t &cols) {
    if (d.size() == 1) { rows = 1; cols = d[0]; }
    else { rows = d[0]; cols = d[1]; }
}
// numpy-style broadcast of two (<= 2-D) shapes; throws the reference's size-mismatch error
inline dims_t broadcast_shape(const dims_t &a, const dims_t &b) {
    size_t ar, ac, br, bc;
    as_2d(a, ar, ac);
    as_2d(b, br, bc);
    if ((ar != br && ar != 1 && br != 1) || (ac != bc && ac != 1 && bc != 1)) throw std::runtime_error(err::size_mismatch());
    const size_t r = ar > br ? ar : br, c = ac > bc ? ac : bc;
    if (a.size() == 1 && b.size() == 1) return {c};
    return {r, c};
}

std::ostream &operator<<(std::ostream &out, const dims_t &d);

// seeded uniform generator (the reference seeds a global engine from time(nullptr), utils.cpp:6 — not reproducible)
void seed_rng(uint64_t seed);
float generate_random(float low, float high);
} // namespace cyg
#endif

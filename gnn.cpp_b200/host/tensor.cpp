// free functions of the tensor layer (counterpart of reference src/tensor.cpp:25-47)
#include "tensor.h"

namespace cyg {
void no_grad(std::vector<tptr<float>> ts) {
    for (const auto &t : ts) t->requires_grad_(false);
}
void enable_grad(std::vector<tptr<float>> ts) {
    for (const auto &t : ts) t->requires_grad_(true);
}
tptr<int> eye(size_t n, size_t m) {
    if (m == (size_t)INT_MAX) m = n;
    auto *h = new std::valarray<int>(0, n * m);
    for (size_t i = 0; i < n && i < m; i++) (*h)[i * m + i] = 1;
    return std::make_shared<tensor<int>>(std::vector<size_t>{n, m}, h, false);
}
} // namespace cyg

#include "utils.h"

#include <random>

namespace cyg {
static std::mt19937_64 &engine() {
    static std::mt19937_64 e(0x5EEDull);
    return e;
}
void seed_rng(uint64_t seed) { engine().seed(seed); }
float generate_random(float low, float high) {
    std::uniform_real_distribution<float> u(low, high);
    return u(engine());
}
std::ostream &operator<<(std::ostream &out, const dims_t &d) {
    out << "(";
    for (size_t i = 0; i < d.size(); i++) out << (i ? " , " : "") << d[i];
    return out << ")";
}
} // namespace cyg

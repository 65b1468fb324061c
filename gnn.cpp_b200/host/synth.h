// synth.h — the shared deterministic problem generator, C++ side (bit-identical to gnn.cpp_b200/synth.py and
// oracle/gcn_oracle.c: counter-based splitmix64).  Used by main.cpp to build BASELINE.json's configs.
#ifndef GNNB200_SYNTH_H
#define GNNB200_SYNTH_H
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

namespace synth {
inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
inline uint64_t hash3(uint64_t seed, uint64_t stream, uint64_t i) {
    return mix64(mix64(seed * 0x9E3779B97F4A7C15ULL + stream * 0xD1B54A32D192ED03ULL) + i);
}
inline void uniform(uint64_t seed, uint64_t stream, size_t n, float lo, float hi, float *out) {
    const float span = hi - lo;
    for (size_t i = 0; i < n; i++) {
        const float u = (float)(hash3(seed, stream, i) >> 40) * 0x1p-24f;
        const float t = span * u;
        out[i] = lo + t;
    }
}
inline uint64_t gcd64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }
inline int32_t endpoint(uint64_t h, int32_t N, bool powerlaw, uint64_t pa, uint64_t pb) {
    if (!powerlaw) return (int32_t)(h % (uint64_t)N);
    const double x = (double)(h >> 32) * 0x1p-32;
    const double s = std::sqrt(x);
    const double t = x * s;
    int64_t id = (int64_t)(t * (double)N);
    if (id >= N) id = N - 1;
    return (int32_t)((pa * (uint64_t)id + pb) % (uint64_t)N);
}
inline void edges(uint64_t seed, int64_t E, int32_t N, bool powerlaw, std::vector<int> &src, std::vector<int> &dst) {
    src.resize(E); dst.resize(E);
    const int64_t np = E / 2;
    uint64_t pa = 0x9E3779B1ULL % (uint64_t)N;
    if (pa == 0) pa = 1;
    while (gcd64(pa, (uint64_t)N) != 1) pa++;
    const uint64_t pb = 0x7F4A7C15ULL % (uint64_t)N;
    for (int64_t k = 0; k < np; k++) {
        const int32_t u = endpoint(hash3(seed, 1, k), N, powerlaw, pa, pb), v = endpoint(hash3(seed, 2, k), N, powerlaw, pa, pb);
        src[k] = u; dst[k] = v; src[np + k] = v; dst[np + k] = u;
    }
    if (E & 1) { src[E - 1] = endpoint(hash3(seed, 1, np), N, powerlaw, pa, pb); dst[E - 1] = endpoint(hash3(seed, 2, np), N, powerlaw, pa, pb); }
}
struct Config { std::string name; int32_t N; int64_t E; std::vector<int32_t> dims; bool powerlaw; int id; uint64_t seed() const { return 1234 + id; } };
inline bool lookup(const std::string &name, Config &c) {
    static const Config all[] = {
        {"cora", 2708, 10556, {1433, 16, 7}, false, 1},       {"pubmed", 19717, 88648, {500, 64, 3}, false, 2},
        {"arxiv", 169343, 1170000, {128, 256, 256, 40}, false, 3}, {"reddit", 232965, 114600000, {602, 128, 41}, true, 4},
        {"products", 2450000, 61900000, {100, 256, 256, 47}, true, 5}, {"tiny", 200, 1200, {24, 16, 5}, false, 91},
        {"tiny_pl", 3000, 60000, {32, 48, 7}, true, 92}};
    for (const auto &x : all) if (x.name == name) { c = x; return true; }
    return false;
}
} // namespace synth
#endif

// problem_io.h — .gcnp problem / .gcno result files shared with gnn.cpp_b200/problem_io.py and the oracle drivers.
#ifndef GNNB200_PROBLEM_IO_H
#define GNNB200_PROBLEM_IO_H
#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace problem_io {
struct Problem {
    int64_t N = 0, E = 0, L = 0;
    std::vector<int64_t> dims;
    std::vector<int> src, dst, y;
    std::vector<float> X;
    std::vector<std::vector<float>> W, b;
};
inline void rd(std::ifstream &f, void *p, size_t n) {
    f.read(reinterpret_cast<char *>(p), (std::streamsize)n);
    if (!f) throw std::runtime_error("problem_io: short read");
}
inline Problem load(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("problem_io: cannot open " + path);
    Problem p;
    int64_t magic;
    rd(f, &magic, 8);
    if (magic != 0x47434E50) throw std::runtime_error("problem_io: bad magic");
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
        p.b[l].resize(p.dims[l + 1]); rd(f, p.b[l].data(), 4 * p.b[l].size());
    }
    return p;
}
struct Writer {
    std::ofstream f;
    explicit Writer(const std::string &path) : f(path, std::ios::binary) {}
    void f32(const std::string &name, const float *data, const std::vector<int64_t> &shape) {
        int32_t nl = (int32_t)name.size(), dt = 0, nd = (int32_t)shape.size();
        size_t n = 1;
        for (auto s : shape) n *= (size_t)s;
        f.write(reinterpret_cast<char *>(&nl), 4); f.write(name.data(), nl);
        f.write(reinterpret_cast<char *>(&dt), 4); f.write(reinterpret_cast<char *>(&nd), 4);
        f.write(reinterpret_cast<const char *>(shape.data()), 8 * nd);
        f.write(reinterpret_cast<const char *>(data), (std::streamsize)(4 * n));
    }
};
} // namespace problem_io
#endif

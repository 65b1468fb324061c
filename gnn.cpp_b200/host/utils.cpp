#include "utils.h"

#include <chrono>
#include <cstdio>
#include <fstream>
#include <random>
#include <thread>

#include "device.h"

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

namespace cyg {
namespace device {
void init_distributed() {
    const char *r = std::getenv("RANK"), *w = std::getenv("WORLD_SIZE");
    const int rank = r ? std::atoi(r) : 0, world = w ? std::atoi(w) : 1;
    if (world <= 1) return;
    const char *rdv = std::getenv("GNN_RDV");
    if (!rdv) throw std::runtime_error("multi-process run needs GNN_RDV (rendezvous file for the NCCL id)");
    unsigned char id[128];
    if (rank == 0) {
        check(gnn_comm_unique_id_h(id));
        const std::string tmp = std::string(rdv) + ".tmp";
        {
            std::ofstream f(tmp, std::ios::binary);
            f.write(reinterpret_cast<const char *>(id), sizeof(id));
        }
        if (std::rename(tmp.c_str(), rdv) != 0) throw std::runtime_error("cannot publish the NCCL id");
    } else {
        bool got = false;
        for (int i = 0; i < 6000 && !got; i++) { // up to 60 s
            std::ifstream f(rdv, std::ios::binary);
            if (f && f.read(reinterpret_cast<char *>(id), sizeof(id)) && f.gcount() == (std::streamsize)sizeof(id)) got = true;
            else std::this_thread::sleep_for(std::chrono::milliseconds(10));
        }
        if (!got) throw std::runtime_error("rank 0 never published the NCCL id");
    }
    check(gnn_comm_init(ctx(), id, rank, world));
    dist().rank = rank;
    dist().world = world;
}
} // namespace device
} // namespace cyg

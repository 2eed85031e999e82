// Times the host-side partition of pack.cu on Netflix-shaped degrees:
//   g++ -O2 -std=c++17 -I mfrec_b200/csrc tools/partition_bench.cpp -o /tmp/partition_bench && /tmp/partition_bench
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "partition.h"

int main(int argc, char **argv)
{
    const int nu = 480000, ni = 17700, B = 148, W = 8;
    std::mt19937_64 rng(1);
    std::lognormal_distribution<double> ln(0.0, 1.0);
    std::vector<int32_t> du(nu), di(ni);
    for (auto &d : du) d = (int32_t)(ln(rng) * 126.0);
    for (int i = 0; i < ni; ++i) di[i] = (int32_t)(1.0e8 * 0.0025 * 71.0 / (i + 71.0));
    auto sorted = [](const std::vector<int32_t> &deg) {
        std::vector<int32_t> ids(deg.size());
        std::iota(ids.begin(), ids.end(), 0);
        std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return deg[a] > deg[b]; });
        return ids;
    };
    const auto su = sorted(du), si = sorted(di);
    std::vector<int32_t> g, p, st, g2, p2, st2;
    mfrec_part::Workspace wu, wi;
    const int threads = argc > 1 ? atoi(argv[1]) : 4;
    for (int rep = 0; rep < 5; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        mfrec_part::partition_ids(du, su, B, W, 1, g, p, st, wu, threads);
        auto t1 = std::chrono::steady_clock::now();
        mfrec_part::partition_ids(di, si, B, W, 1, g2, p2, st2, wi, 1);
        auto t2 = std::chrono::steady_clock::now();
        std::printf("users %.2f ms, items %.2f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(),
                    std::chrono::duration<double, std::milli>(t2 - t1).count());
    }
    return 0;
}

// Checks the host-side partition of pack.cu (mfrec_b200/csrc/partition.h) on power-law degrees:
// prints one line "ok <max group load / mean> <max block load / mean> <min count> <max count>"
// or "FAIL <why>".  Driven by tests/test_partition_cpu.py.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "partition.h"

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 100000, nblocks = argc > 2 ? atoi(argv[2]) : 37, W = argc > 3 ? atoi(argv[3]) : 8;
    const int slabs = argc > 4 ? atoi(argv[4]) : 1;
    std::mt19937_64 rng(7);
    std::lognormal_distribution<double> ln(0.0, 1.2);
    std::vector<int32_t> deg(n);
    for (auto &d : deg) d = (int32_t)(ln(rng) * 50.0);
    std::vector<int32_t> sorted(n);
    std::iota(sorted.begin(), sorted.end(), 0);
    std::stable_sort(sorted.begin(), sorted.end(), [&](int a, int b) { return deg[a] > deg[b]; });
    std::vector<int32_t> group, perm, start, group2, perm2, start2;
    mfrec_part::partition_ids(deg, sorted, nblocks, W, slabs, group, perm, start);
    mfrec_part::partition_ids(deg, sorted, nblocks, W, slabs, group2, perm2, start2);
    if (group != group2 || perm != perm2 || start != start2) { puts("FAIL not deterministic"); return 0; }
    {   // the threaded second level with a reused workspace must give the same partition
        mfrec_part::Workspace ws;
        for (int rep = 0; rep < 2; ++rep) {
            mfrec_part::partition_ids(deg, sorted, nblocks, W, slabs, group2, perm2, start2, ws, 4);
            if (group != group2 || perm != perm2 || start != start2) { puts("FAIL threads change the result"); return 0; }
        }
    }
    const int ng = nblocks * W;
    if ((int)start.size() != ng + 1 || start[0] != 0 || start[ng] != n) { puts("FAIL start"); return 0; }
    std::vector<char> seen(n, 0);
    std::vector<double> load(ng, 0.0);
    std::vector<int> count(ng, 0);
    for (int id = 0; id < n; ++id) {
        const int g = group[id], p = perm[id];
        if (g < 0 || g >= ng || p < start[g] || p >= start[g + 1] || seen[p]) { puts("FAIL perm is not a bijection into the group ranges"); return 0; }
        seen[p] = 1;
        load[g] += deg[id] + 1;
        count[g] += 1;
    }
    // ascending original id inside a group
    std::vector<int> last(ng, -1);
    for (int id = 0; id < n; ++id) {
        if (perm[id] <= last[group[id]]) { puts("FAIL order inside group"); return 0; }
        last[group[id]] = perm[id];
    }
    double tot = 0, gmax = 0, bmax = 0;
    for (double l : load) { tot += l; gmax = std::max(gmax, l); }
    for (int b = 0; b < nblocks; ++b) {
        double bl = 0;
        for (int w = 0; w < W; ++w) bl += load[b * W + w];
        bmax = std::max(bmax, bl);
    }
    printf("ok %.5f %.5f %d %d\n", gmax / (tot / ng), bmax / (tot / nblocks), *std::min_element(count.begin(), count.end()),
           *std::max_element(count.begin(), count.end()));
    return 0;
}

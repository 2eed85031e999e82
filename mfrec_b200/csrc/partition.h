// Host-side balanced partition of users / items into blocks and groups (used by pack.cu).
// Header-only and free of CUDA so that it can be unit-tested and timed on a CPU
// (tools/partition_bench.cpp, tests/test_partition_cpu.py).
#pragma once

#include <stdint.h>

#include <algorithm>
#include <functional>
#include <numeric>
#include <queue>
#include <thread>
#include <utility>
#include <vector>

namespace mfrec_part {

// Balanced partition of n elements, given HEAVIEST FIRST by their degrees, into `bins` bins by
// (degree + 1).  Everything works on positions in that order ("ranks"): sequential memory only.
// The head (8 elements per bin) is placed by exact longest-processing-time-first with a heap; the
// long tail of light elements is dealt in rounds -- the bins still below the target load,
// lightest first, each take the next heaviest element -- which is O(n) instead of O(n log bins)
// and within ~0.1 % of LPT on power-law degrees.  Deterministic.
inline void balanced_bins(const int32_t *deg_sorted, int64_t n, int bins, int32_t *bin_of_rank)
{
    typedef std::pair<int64_t, int> Load;  // (load, bin): min-heap, ties -> lowest bin
    std::vector<int64_t> load(bins, 0);
    int64_t total = 0;
    for (int64_t j = 0; j < n; ++j) total += (int64_t)deg_sorted[j] + 1;
    const int64_t head = std::min<int64_t>(n, (int64_t)8 * bins);
    {
        std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
        for (int b = 0; b < bins; ++b) heap.push(Load(0, b));
        for (int64_t j = 0; j < head; ++j) {
            Load top = heap.top();
            heap.pop();
            bin_of_rank[j] = top.second;
            top.first += (int64_t)deg_sorted[j] + 1;
            load[top.second] = top.first;
            heap.push(top);
        }
    }
    if (head == n) return;
    const int64_t target = (total + bins - 1) / bins;
    std::vector<int> order(bins);   // bins by (load, index), kept sorted with an adaptive sort
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        return load[a] != load[b] ? load[a] < load[b] : a < b;
    });
    int64_t pos = head;
    while (pos < n) {
        int m = 0;   // bins below the target take part (all of them if none is)
        while (m < bins && load[order[m]] < target) ++m;
        if (m == 0) m = bins;
        m = (int)std::min<int64_t>(m, n - pos);
        for (int j = 0; j < m; ++j) {
            bin_of_rank[pos + j] = order[j];
            load[order[j]] += (int64_t)deg_sorted[pos + j] + 1;
        }
        pos += m;
        for (int j = 1; j < bins; ++j) {   // insertion sort: the order changes little per round
            const int bj = order[j];
            int i = j - 1;
            while (i >= 0 && (load[order[i]] > load[bj] || (load[order[i]] == load[bj] && order[i] > bj))) {
                order[i + 1] = order[i];
                --i;
            }
            order[i + 1] = bj;
        }
    }
}

// Scratch of partition_ids.  Kept by the caller across calls: the vectors are several MB at
// Netflix size and a fresh allocation of that size is page-faulted in on every call, which costs
// more than the partition itself.
struct Workspace {
    std::vector<int32_t> deg_sorted, block_r, members, mdeg, sub, count, cursor;
    std::vector<int64_t> bstart, cur;
};

// Hierarchical partition: ids -> nblocks blocks -> W groups each, balanced by (degree + 1).
// `sorted` lists all ids heaviest first (ties in a seeded pseudo-random order; computed on the
// device).  Outputs, for every id, its group (block * W + group-in-block) and its packed id
// (groups are contiguous id ranges, ascending original id inside a group); start[g] = first
// packed id of group g.  `threads` > 1 splits the per-block second level over that many host
// threads (the result does not depend on it).
inline void partition_ids(const std::vector<int32_t> &deg, const std::vector<int32_t> &sorted, int nblocks,
                          int W, int n_slabs, std::vector<int32_t> &group_of, std::vector<int32_t> &perm,
                          std::vector<int32_t> &start, Workspace &ws, int threads = 1)
{
    const int64_t n = (int64_t)deg.size();
    ws.deg_sorted.resize(n);
    ws.block_r.resize(n);
    int32_t *deg_sorted = ws.deg_sorted.data(), *block_r = ws.block_r.data();
    for (int64_t j = 0; j < n; ++j) deg_sorted[j] = deg[sorted[j]];   // the one gather by id
    balanced_bins(deg_sorted, n, nblocks, block_r);
    if (n_slabs > 1) {
        // The heaviest ids land in the lowest-numbered bins.  A heavy item is a long dependent
        // chain for the SGD kernel, so deal the bins round-robin over the slabs (bin j -> slab
        // j mod G): every slab then gets its share of hot items and the slabs of a DSGD ring take
        // equal time, not just equal counts.
        const int per = nblocks / n_slabs;
        for (int64_t j = 0; j < n; ++j) block_r[j] = (block_r[j] % n_slabs) * per + block_r[j] / n_slabs;
    }
    // ranks of each block, still heaviest first (counting sort by block)
    ws.bstart.assign(nblocks + 1, 0);
    int64_t *bstart = ws.bstart.data();
    for (int64_t j = 0; j < n; ++j) bstart[block_r[j] + 1] += 1;
    for (int b = 0; b < nblocks; ++b) bstart[b + 1] += bstart[b];
    ws.members.resize(n);
    ws.mdeg.resize(n);
    ws.sub.resize(n);
    int32_t *members = ws.members.data(), *mdeg = ws.mdeg.data(), *sub = ws.sub.data();
    {
        ws.cur.assign(ws.bstart.begin(), ws.bstart.end() - 1);
        int64_t *cur = ws.cur.data();
        for (int64_t j = 0; j < n; ++j) {
            const int64_t at = cur[block_r[j]]++;
            members[at] = (int32_t)j;
            mdeg[at] = deg_sorted[j];
        }
    }
    group_of.resize(n);
    int32_t *gof = group_of.data();
    const int32_t *srt = sorted.data();
    auto second_level = [&](int b0, int b1) {
        for (int b = b0; b < b1; ++b) {
            const int64_t a = bstart[b], m = bstart[b + 1] - a;
            balanced_bins(mdeg + a, m, W, sub + a);
            for (int64_t t = a; t < a + m; ++t) gof[srt[members[t]]] = b * W + sub[t];   // the one scatter by id
        }
    };
    threads = std::max(1, std::min(threads, nblocks));
    if (threads == 1 || n < 65536) {
        second_level(0, nblocks);
    } else {
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; ++t)
            pool.emplace_back(second_level, (int)((int64_t)nblocks * t / threads),
                              (int)((int64_t)nblocks * (t + 1) / threads));
        second_level(0, nblocks / threads);
        for (auto &th : pool) th.join();
    }
    const int ng = nblocks * W;
    ws.count.assign(ng + 1, 0);
    for (int64_t id = 0; id < n; ++id) ws.count[gof[id] + 1] += 1;
    start.assign(ng + 1, 0);
    for (int g = 0; g < ng; ++g) start[g + 1] = start[g] + ws.count[g + 1];
    ws.cursor.assign(start.begin(), start.end() - 1);
    int32_t *cursor = ws.cursor.data();
    perm.resize(n);
    for (int64_t id = 0; id < n; ++id) perm[id] = cursor[gof[id]]++;
}

// Item side with hot-item copies (pack.cu): the ids are VIRTUAL items -- item i appears as
// vbase[i+1] - vbase[i] copies, each trained on a disjoint share of the item's ratings and merged
// after every epoch -- and the partition has three levels: real items -> n_slabs slabs (so that
// all copies of an item travel together and are merged by whichever rank holds the slab),
// a slab's virtual ids -> blocks_per_slab blocks, a block's ids -> W groups; every level balanced
// by (degree + 1).  `deg` / `sorted` are over virtual ids (heaviest first), `real_deg` over real
// items.  With n_slabs == 1 and no copies this is partition_ids.
inline void partition_items(const std::vector<int32_t> &deg, const std::vector<int32_t> &sorted,
                            const std::vector<int32_t> &real_deg, const std::vector<int32_t> &vbase,
                            int n_slabs, int blocks_per_slab, int W, std::vector<int32_t> &group_of,
                            std::vector<int32_t> &perm, std::vector<int32_t> &start, Workspace &ws)
{
    const int64_t n = (int64_t)deg.size();
    const int32_t ni = (int32_t)real_deg.size();
    const int nblocks = n_slabs * blocks_per_slab;
    // level 0: real items -> slabs (heaviest first, ties by id: every rank of a ring computes the same)
    std::vector<int32_t> slab_of_item(ni, 0);
    if (n_slabs > 1) {
        std::vector<int32_t> order(ni), dsorted(ni), bin(ni);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return real_deg[a] > real_deg[b]; });
        for (int32_t j = 0; j < ni; ++j) dsorted[j] = real_deg[order[j]];
        balanced_bins(dsorted.data(), ni, n_slabs, bin.data());
        for (int32_t j = 0; j < ni; ++j) slab_of_item[order[j]] = bin[j];
    }
    std::vector<int32_t> item_of(n);
    for (int32_t i = 0; i < ni; ++i)
        for (int32_t v = vbase[i]; v < vbase[i + 1]; ++v) item_of[v] = i;
    // level 1: per slab, its virtual ids (still heaviest first) -> blocks
    ws.deg_sorted.resize(n);
    ws.block_r.resize(n);
    int32_t *deg_sorted = ws.deg_sorted.data(), *block_r = ws.block_r.data();
    for (int64_t j = 0; j < n; ++j) deg_sorted[j] = deg[sorted[j]];
    {
        std::vector<std::vector<int64_t>> ranks_of_slab(n_slabs);
        for (int64_t j = 0; j < n; ++j) ranks_of_slab[slab_of_item[item_of[sorted[j]]]].push_back(j);
        std::vector<int32_t> d, b;
        for (int g = 0; g < n_slabs; ++g) {
            const std::vector<int64_t> &rk = ranks_of_slab[g];
            d.resize(rk.size());
            b.resize(rk.size());
            for (size_t t = 0; t < rk.size(); ++t) d[t] = deg_sorted[rk[t]];
            if (!rk.empty()) balanced_bins(d.data(), (int64_t)rk.size(), blocks_per_slab, b.data());
            for (size_t t = 0; t < rk.size(); ++t) block_r[rk[t]] = g * blocks_per_slab + b[t];
        }
    }
    // level 2 and the relabelling: as in partition_ids
    ws.bstart.assign(nblocks + 1, 0);
    int64_t *bstart = ws.bstart.data();
    for (int64_t j = 0; j < n; ++j) bstart[block_r[j] + 1] += 1;
    for (int b = 0; b < nblocks; ++b) bstart[b + 1] += bstart[b];
    ws.members.resize(n);
    ws.mdeg.resize(n);
    ws.sub.resize(n);
    int32_t *members = ws.members.data(), *mdeg = ws.mdeg.data(), *sub = ws.sub.data();
    {
        ws.cur.assign(ws.bstart.begin(), ws.bstart.end() - 1);
        int64_t *cur = ws.cur.data();
        for (int64_t j = 0; j < n; ++j) {
            const int64_t at = cur[block_r[j]]++;
            members[at] = (int32_t)j;
            mdeg[at] = deg_sorted[j];
        }
    }
    group_of.resize(n);
    int32_t *gof = group_of.data();
    for (int b = 0; b < nblocks; ++b) {
        const int64_t a = bstart[b], m = bstart[b + 1] - a;
        if (m > 0) balanced_bins(mdeg + a, m, W, sub + a);
        for (int64_t t = a; t < a + m; ++t) gof[sorted[members[t]]] = b * W + sub[t];
    }
    const int ng = nblocks * W;
    ws.count.assign(ng + 1, 0);
    for (int64_t id = 0; id < n; ++id) ws.count[gof[id] + 1] += 1;
    start.assign(ng + 1, 0);
    for (int g = 0; g < ng; ++g) start[g + 1] = start[g] + ws.count[g + 1];
    ws.cursor.assign(start.begin(), start.end() - 1);
    int32_t *cursor = ws.cursor.data();
    perm.resize(n);
    for (int64_t id = 0; id < n; ++id) perm[id] = cursor[gof[id]]++;
}

inline void partition_ids(const std::vector<int32_t> &deg, const std::vector<int32_t> &sorted, int nblocks,
                          int W, int n_slabs, std::vector<int32_t> &group_of, std::vector<int32_t> &perm,
                          std::vector<int32_t> &start)
{
    Workspace ws;
    partition_ids(deg, sorted, nblocks, W, n_slabs, group_of, perm, start, ws, 1);
}

}  // namespace mfrec_part

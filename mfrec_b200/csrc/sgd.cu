// SGD update kernels for the all-k "KMF" loops of the reference
// (mfrec/lib/kmf_train.pyx:195-277 train_linear_kernel, :103-189 train_logistic_kernel).
//
// Stratified schedule (MFREC_SCHED_STRATIFIED)
// --------------------------------------------
// Two updates commute iff they share neither the user row nor the item row.  pack.cu cuts the
// rating matrix into B x B blocks (x G item slabs) and every block into W x W buckets.
// One epoch = B kernel launches per slab ("sub-epochs").  In sub-epoch s, CTA rb owns
//   row block rb  x  column block (rb + s) mod B            (a Latin square: no two CTAs share
//                                                            a row block or a column block)
// and inside the CTA, in phase p, warp w owns
//   row group w   x  column group (w + p) mod W             (again a Latin square)
// with a __syncthreads() between phases.  So at any instant all B*W warps of the grid work on
// pairwise disjoint users and items: no atomics, no locks, deterministic for a given layout.
// Any serial replay that visits (sub-epoch, CTA, phase, warp, bucket order) nested in that
// order is an equivalent sequential SGD order; tests replay exactly that with the CPU oracle.
//
// Data movement per CTA
//   * the column block's Q rows (contiguous in packed-id order) are pulled into shared memory
//     once with cp.async.bulk (TMA, mbarrier complete_tx) and written back once;
//   * each warp's W buckets are stored back to back, so the warp streams ONE contiguous run of
//     12-byte ratings through a private 4-stage shared-memory ring filled by cp.async.bulk;
//   * P rows are fetched with cp.async (16 bytes per lane, coalesced 512 B per row at k = 128)
//     into an 8-deep shared-memory ring, 8 ratings ahead of the consumer and across phase
//     boundaries; they are tracked by cp.async groups, not by the register scoreboard, so a
//     wait never stalls on the newest request; updated rows go back with 128-bit stores;
//   * the dot product is a 5-step warp-shuffle butterfly; all arithmetic is fp32, the epoch's
//     sum of squared errors is accumulated in fp64.
//
// Sequential schedule (MFREC_SCHED_SEQUENTIAL): one thread, fp64, no FMA contraction, the
// reference's exact order on the reference's own [k][n] layout.  Bit-exact with the reference
// for the linear kernel; used for verification and for tiny fold-in calls.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int kChunk = 64;    // ratings per TMA chunk
constexpr int kStages = 4;    // chunks in flight per warp (ring of kStages * kChunk ratings)
constexpr int kDepth = 16;    // P rows in flight per warp (cp.async ring), in stream positions
constexpr int kQuadsAhead = kDepth / 4;
constexpr int kRing = kChunk * kStages;
static_assert(kChunk % 4 == 0, "chunks must be 16-byte multiples");
static_assert(4 * kDepth <= kChunk, "the prefetch cursor may run at most one chunk ahead");

struct SgdParams {
    const PackedRating *packed;
    const int64_t *bucket_off;
    const int32_t *bucket_cnt;
    const int32_t *col_start;
    float *Q, *ib, *P, *ub;
    double *se_part;  // [B] partial sums of this launch
    int B, W, slab, s;
    int tile_rows;    // shared-memory rows reserved for the Q tile
    float lr, Ku, Ki, Kb;
    int update_users, update_items;
    float fx_scale, fx_inv;       // fixed-point scale of the warp reduction (power of two)
    unsigned long long *timing;   // debug (MFREC_SGD_TIMING=1): [B][W][8] cycle counters, or null
};

// ---- PTX helpers: mbarrier + bulk async copy (TMA, non-tensor form) + cp.async ------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cp_async(void *dst_smem, const void *src_gmem)
{
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- per-lane row fragment: E floats, as NV vectors of V floats, interleaved over the warp --
template <int E>
struct Frag {
    static constexpr int V = E >= 4 ? 4 : E;
    static constexpr int NV = E / V;
    float x[E];
};

template <int E>
__device__ __forceinline__ void frag_load(Frag<E> &f, const float *row, int lane)
{
    constexpr int V = Frag<E>::V, NV = Frag<E>::NV;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        const float *p = row + (c * 32 + lane) * V;
        if constexpr (V == 4) {
            const float4 t = *reinterpret_cast<const float4 *>(p);
            f.x[c * 4 + 0] = t.x; f.x[c * 4 + 1] = t.y; f.x[c * 4 + 2] = t.z; f.x[c * 4 + 3] = t.w;
        } else if constexpr (V == 2) {
            const float2 t = *reinterpret_cast<const float2 *>(p);
            f.x[0] = t.x; f.x[1] = t.y;
        } else {
            f.x[0] = *p;
        }
    }
}

template <int E>
__device__ __forceinline__ void frag_store(const Frag<E> &f, float *row, int lane)
{
    constexpr int V = Frag<E>::V, NV = Frag<E>::NV;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        float *p = row + (c * 32 + lane) * V;
        if constexpr (V == 4)
            *reinterpret_cast<float4 *>(p) = make_float4(f.x[c * 4], f.x[c * 4 + 1], f.x[c * 4 + 2], f.x[c * 4 + 3]);
        else if constexpr (V == 2)
            *reinterpret_cast<float2 *>(p) = make_float2(f.x[0], f.x[1]);
        else
            *p = f.x[0];
    }
}

// asynchronous global -> shared copy of one row, same lane <-> column mapping as frag_load
template <int E>
__device__ __forceinline__ void row_cp_async(float *dst_row, const float *src_row, int lane)
{
    constexpr int V = Frag<E>::V, NV = Frag<E>::NV;
#pragma unroll
    for (int c = 0; c < NV; ++c)
        cp_async<V * 4>(dst_row + (c * 32 + lane) * V, src_row + (c * 32 + lane) * V);
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// The stratified kernel.  grid = B CTAs, block = W warps.  E = kpad / 32 floats per lane.
//
// Every warp walks ONE contiguous stream of packed ratings (its W buckets, stored
// back to back) with three cursors:
//   load cursor     : cp.async.bulk chunks of kChunk ratings into a kStages-deep ring
//   prefetch cursor : kDepth ratings ahead of the consumer, cp.async of the P row (and user
//                     bias) of each upcoming rating into a kDepth-deep shared-memory ring;
//                     P rows of a row group belong to this warp for the whole launch, so the
//                     cursor may run across phase boundaries
//   consume cursor  : the update itself; __syncthreads() at every bucket (= phase) end
// Stale-prefetch hazard: the row of rating x is fetched while ratings x-kDepth .. x-1 are still
// being applied.  Inside a bucket equal users are adjacent (sorted) and are served from
// registers; across a bucket boundary a lane-distributed history of the last kDepth users
// detects the rare repeat and re-reads the row from global memory.
// ------------------------------------------------------------------------------------------
template <int E, int KERNEL>
__global__ void __launch_bounds__(512)
sgd_block_kernel(const SgdParams prm)
{
    constexpr int KPAD = E * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = prm.W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rb = blockIdx.x;
    const int cbl = (rb + prm.s) % prm.B;
    const int cbg = prm.slab * prm.B + cbl;
    const int cs = prm.col_start[cbg * W];
    const int nq = prm.col_start[(cbg + 1) * W] - cs;

    // shared-memory carve-up (every section is a multiple of 16 bytes)
    float *Qs = reinterpret_cast<float *>(smem_raw);
    float *ibs = Qs + (size_t)prm.tile_rows * KPAD;
    float *prow_all = ibs + ((prm.tile_rows + 3) & ~3);
    float *pbias_all = prow_all + (size_t)W * kDepth * KPAD;
    PackedRating *ring_all = reinterpret_cast<PackedRating *>(pbias_all + (size_t)W * kDepth);
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring_all + (size_t)W * kRing);
    int64_t *boff = reinterpret_cast<int64_t *>(bars + 1 + kStages * W);
    int32_t *bcnt = reinterpret_cast<int32_t *>(boff + W * W + 1);
    double *se_s = reinterpret_cast<double *>(bcnt + ((W * W + 1) & ~1));

    float *prow = prow_all + (size_t)warp * kDepth * KPAD;
    float *pbias = pbias_all + warp * kDepth;
    PackedRating *ring = ring_all + (size_t)warp * kRing;
    uint64_t *tile_bar = bars;
    uint64_t *my_bar = bars + 1 + kStages * warp;

    if (threadIdx.x == 0) {
        mbar_init(tile_bar, 1);
        for (int i = 0; i < kStages * W; ++i) mbar_init(bars + 1 + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // this CTA's W*W bucket descriptors (worker-major: index w * W + phase) + end sentinel
    const int64_t bucket_base = (((int64_t)prm.slab * prm.B + rb) * prm.B + cbl) * W * W;
    for (int i = threadIdx.x; i <= W * W; i += blockDim.x) {
        boff[i] = prm.bucket_off[bucket_base + i];
        if (i < W * W) bcnt[i] = prm.bucket_cnt[bucket_base + i];
    }
    __syncthreads();

    // Q tile: one elected thread issues the bulk copies, everyone waits on the mbarrier
    if (threadIdx.x == 0 && nq > 0) {
        const uint32_t total = (uint32_t)nq * KPAD * 4u;
        mbar_expect_tx(tile_bar, total);
        const char *src = reinterpret_cast<const char *>(prm.Q + (size_t)cs * KPAD);
        char *dst = reinterpret_cast<char *>(Qs);
        for (uint32_t o = 0; o < total; o += 32768u)
            bulk_g2s(dst + o, src + o, min(32768u, total - o), tile_bar);
    }
    for (int i = threadIdx.x; i < nq; i += blockDim.x) ibs[i] = prm.ib[cs + i];

    // ---- this warp's stream ---------------------------------------------------------------
    const int64_t S0 = boff[warp * W];
    const uint32_t slen = (uint32_t)(boff[warp * W + W] - S0);   // multiple of 4 (padded buckets)
    const uint32_t nchunks = (slen + kChunk - 1) / kChunk;
    const uint32_t nquads = slen / 4;
    const PackedRating *stream = prm.packed + S0;
    const int32_t *cnt_w = bcnt + warp * W;
    const int64_t *off_w = boff + warp * W;
    uint32_t issued = 0;

    auto issue_chunk = [&](uint32_t c) {
        if (lane == 0) {
            const uint32_t cnt = min((uint32_t)kChunk, slen - c * kChunk);
            uint64_t *bar = my_bar + (c % kStages);
            mbar_expect_tx(bar, cnt * 12u);
            bulk_g2s(ring + (c % kStages) * kChunk, stream + (size_t)c * kChunk, cnt * 12u, bar);
        }
    };
    while (issued < nchunks && issued < (uint32_t)kStages) issue_chunk(issued++);

    // Prefetch cursor: walks the stream one QUAD (4 positions = 48 bytes, 16-byte aligned) at a
    // time, padding entries included (they name packed user 0, a valid row, and are never
    // consumed).  One cp.async group per quad, so group index == quad index and
    // cp.async.wait_group<kQuadsAhead-1> at quad x guarantees its four rows have landed once
    // the cursor stands at x + kQuadsAhead.  Quads past the end commit empty groups.
    uint32_t pfq = 0;
    auto prefetch_quads_to = [&](uint32_t target) {
#pragma unroll 1
        while (pfq < target) {
            if (pfq < nquads) {
                const uint32_t pos = pfq * 4;
                if ((pos % kChunk) == 0) {   // first touch of a chunk: wait for its bulk copy
                    const uint32_t c = pos / kChunk;
                    mbar_wait(my_bar + (c % kStages), (c / kStages) & 1u);
                }
                const int4 *qsrc = reinterpret_cast<const int4 *>(ring + (pos % kRing));
                const int4 a = qsrc[0], b = qsrc[1], c2 = qsrc[2];
                const int us[4] = {a.x, a.w, b.z, c2.y};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint32_t slot = (pos + t) % kDepth;
                    row_cp_async<E>(prow + slot * KPAD, prm.P + (size_t)us[t] * KPAD, lane);
                    if (lane == 0) cp_async<4>(pbias + slot, prm.ub + us[t]);
                }
            }
            cp_async_commit();
            ++pfq;
        }
    };
    prefetch_quads_to(kQuadsAhead);

    if (nq > 0) mbar_wait(tile_bar, 0);
    __syncthreads();

    const float lr = prm.lr;
    const bool upd_u = prm.update_users != 0, upd_i = prm.update_items != 0;
    // q' = q + lr (g p - Ki q) = (1 - lr Ki) q + (lr g) p, likewise for p (frozen side: a = 1, gl = 0)
    const float a_i = upd_i ? 1.f - prm.lr * prm.Ki : 1.f;
    const float a_u = upd_u ? 1.f - prm.lr * prm.Ku : 1.f;
    const float a_b = 1.f - prm.lr * prm.Kb;
    const bool upd_bu = (KERNEL == MFREC_KERNEL_LINEAR) || upd_u;
    const bool upd_bi = (KERNEL == MFREC_KERNEL_LINEAR) || upd_i;
    const float fx_scale = prm.fx_scale, fx_inv = prm.fx_inv;
    double se = 0.0;       // fp64 total of fp32 per-bucket partials
    // opt-in section timer (cycles per warp): 0 bookkeeping + prefetch issue, 1 cp.async wait,
    // 2 quad load, 3 updates, 6 phase barrier
    unsigned long long tsec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool timing = prm.timing != nullptr;
    long long tmark = timing ? clock64() : 0;
    auto lap = [&](int k) {
        if (timing) {
            const long long now = clock64();
            tsec[k] += (unsigned long long)(now - tmark);
            tmark = now;
        }
    };
    Frag<E> cp, cq;        // current user row / item row (post-update values)
#pragma unroll
    for (int e = 0; e < E; ++e) { cp.x[e] = 0.f; cq.x[e] = 0.f; }
    float cbu = 0.f, cbi = 0.f, se_f = 0.f;
    int prev_u = -1, prev_i = -1, hist = -1;
    uint32_t cons_chunk = 0;

    // one rating: pos = stream position, (u, i, r) the triple
    auto update_one = [&](uint32_t pos, int u, int it, float r) {
        const uint32_t slot = pos % kDepth;
        const bool same_u = (u == prev_u);
        Frag<E> pu;
        float bu;
        if (!same_u && __any_sync(0xffffffffu, hist == u)) {
            // the row was updated after its prefetch was issued (the user repeats across a bucket
            // boundary inside the prefetch window): re-read it from global memory
            frag_load<E>(pu, prm.P + (size_t)u * KPAD, lane);
            bu = prm.ub[u];
        } else {
            frag_load<E>(pu, prow + slot * KPAD, lane);
            bu = pbias[slot];
        }
#pragma unroll
        for (int e = 0; e < E; ++e) pu.x[e] = same_u ? cp.x[e] : pu.x[e];
        bu = same_u ? cbu : bu;
        float *qrow = Qs + (size_t)(it - cs) * KPAD;
        if (it != prev_i) {   // otherwise the item row is still in registers
            frag_load<E>(cq, qrow, lane);
            cbi = ibs[it - cs];
        }
        float part = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) part = fmaf(pu.x[e], cq.x[e], part);
        // deterministic warp reduction in 32-bit fixed point: one REDUX instead of a 5-level
        // shuffle butterfly; integer addition is associative, so the result does not depend on
        // lane order.  fx_scale is a power of two chosen from max |rating| (see sgd_epoch).
        const float dot = (float)__reduce_add_sync(0xffffffffu, __float2int_rn(part * fx_scale)) * fx_inv;
        const float pred = cbi + bu + dot;
        float err, grad;
        if constexpr (KERNEL == MFREC_KERNEL_LINEAR) {
            err = r - pred;
            grad = err;
        } else {
            const float sig = 1.f / (1.f + expf(-pred));
            err = r - (1.f + 4.f * sig);
            grad = err * sig * (1.f - sig) * 4.f;
        }
        se_f = fmaf(err, err, se_f);
        const float gl = lr * grad;
        cbu = upd_bu ? fmaf(a_b, bu, gl) : bu;
        cbi = upd_bi ? fmaf(a_b, cbi, gl) : cbi;
        const float gli = upd_i ? gl : 0.f, glu = upd_u ? gl : 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const float pe = pu.x[e], qe = cq.x[e];
            cp.x[e] = fmaf(glu, qe, a_u * pe);
            cq.x[e] = fmaf(gli, pe, a_i * qe);
        }
        prev_u = u;
        prev_i = it;
        if (lane == (int)slot) hist = u;
        frag_store<E>(cq, qrow, lane);
        frag_store<E>(cp, prm.P + (size_t)u * KPAD, lane);
        if (lane == 0) {
            ibs[it - cs] = cbi;
            prm.ub[u] = cbu;
        }
    };

    for (int p = 0; p < W; ++p) {
        const uint32_t n = (uint32_t)cnt_w[p];
        const uint32_t rel0 = (uint32_t)(off_w[p] - S0);
        se_f = 0.f;
#pragma unroll 1
        for (uint32_t i = 0; i < n; i += 4) {
            const uint32_t rel = rel0 + i;   // quad aligned
            if (rel / kChunk != cons_chunk) {
                cons_chunk = rel / kChunk;   // every earlier chunk's stage is free again
                __syncwarp();
                while (issued < nchunks && issued < cons_chunk + kStages) issue_chunk(issued++);
            }
            prefetch_quads_to(rel / 4 + kQuadsAhead);
            lap(0);
            cp_async_wait<kQuadsAhead - 1>();   // the four rows of this quad have landed
            __syncwarp();                       // ... and lane 0's bias copies / stores are visible
            lap(1);
            const int4 *qsrc = reinterpret_cast<const int4 *>(ring + (rel % kRing));
            const int4 a = qsrc[0], b = qsrc[1], c2 = qsrc[2];
            const uint32_t nv = n - i;   // valid ratings in this quad (>= 1; >= 4 means all)
            lap(2);
            update_one(rel, a.x, a.y, __int_as_float(a.z));
            if (nv > 1) update_one(rel + 1, a.w, b.x, __int_as_float(b.y));
            if (nv > 2) update_one(rel + 2, b.z, b.w, __int_as_float(c2.x));
            if (nv > 3) update_one(rel + 3, c2.y, c2.z, __int_as_float(c2.w));
            lap(3);
        }
        se += (double)se_f;
        prev_i = -1;       // the column group changes hands: never forward Q across a phase
        __syncthreads();   // phase boundary
        lap(6);
    }
    cp_async_wait<0>();
    if (timing && lane == 0)
        for (int k2 = 0; k2 < 8; ++k2) prm.timing[((size_t)rb * W + warp) * 8 + k2] = tsec[k2];
    // write the Q tile back
    {
        const int nvec = nq * KPAD / 4;
        float4 *dst = reinterpret_cast<float4 *>(prm.Q + (size_t)cs * KPAD);
        const float4 *srcv = reinterpret_cast<const float4 *>(Qs);
        for (int i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = srcv[i];
        for (int i = threadIdx.x; i < nq; i += blockDim.x) prm.ib[cs + i] = ibs[i];
    }
    if (lane == 0) se_s[warp] = se;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < W; ++w) tot += se_s[w];
        prm.se_part[rb] = tot;
    }
}

// fixed-order sum of the per-CTA partials of one epoch
__global__ void __launch_bounds__(1024) se_reduce_kernel(const double *__restrict__ part, int64_t n,
                                                         double *__restrict__ out)
{
    __shared__ double sh[1024];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) acc += part[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

// ------------------------------------------------------------------------------------------
// Sequential schedule: the reference's loop verbatim (kmf_train.pyx:241-273), one thread,
// fp64, round-to-nearest multiplies and adds kept separate (the reference build has no FMA).
// ------------------------------------------------------------------------------------------
__global__ void kmf_sequential_kernel(int kernel, int nbr_epochs, int dim, double lr, double K_users,
                                      double K_items, double K_bias, double *u, double *v,
                                      const int32_t *idx, const double *ratings, int64_t nnz,
                                      int64_t ni, int64_t nu, double *ib, double *ub,
                                      int update_users, int update_items, double *rmse_out)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int epoch = 0; epoch < nbr_epochs; ++epoch) {
        double se = 0.0;
        for (int64_t n = 0; n < nnz; ++n) {
            const int user = idx[2 * n], item = idx[2 * n + 1];
            const double rating = ratings[n];
            double s = __dadd_rn(__dadd_rn(0.0, ib[item]), ub[user]);
            for (int f = 0; f < dim; ++f)
                s = __dadd_rn(s, __dmul_rn(u[(int64_t)f * ni + item], v[(int64_t)f * nu + user]));
            double err, grad;
            if (kernel == MFREC_KERNEL_LINEAR) {
                err = __dadd_rn(rating, -s);
                grad = err;
            } else {
                const double sig = 1.0 / __dadd_rn(1.0, exp(-s));
                const double p = __dadd_rn(1.0, __dmul_rn(sig, 4.0));
                err = __dadd_rn(rating, -p);
                grad = __dmul_rn(__dmul_rn(__dmul_rn(err, sig), __dadd_rn(1.0, -sig)), 4.0);
            }
            se = __dadd_rn(se, __dmul_rn(err, err));
            if (kernel == MFREC_KERNEL_LINEAR || update_users)
                ub[user] = __dadd_rn(ub[user], __dmul_rn(lr, __dadd_rn(grad, -__dmul_rn(K_bias, ub[user]))));
            if (kernel == MFREC_KERNEL_LINEAR || update_items)
                ib[item] = __dadd_rn(ib[item], __dmul_rn(lr, __dadd_rn(grad, -__dmul_rn(K_bias, ib[item]))));
            for (int f = 0; f < dim; ++f) {
                double *pu = &u[(int64_t)f * ni + item], *pv = &v[(int64_t)f * nu + user];
                const double cf = *pv, mf = *pu;
                if (update_items)
                    *pu = __dadd_rn(mf, __dmul_rn(lr, __dadd_rn(__dmul_rn(grad, cf), -__dmul_rn(K_items, mf))));
                if (update_users)
                    *pv = __dadd_rn(cf, __dmul_rn(lr, __dadd_rn(__dmul_rn(grad, mf), -__dmul_rn(K_users, cf))));
            }
        }
        if (rmse_out) rmse_out[epoch] = sqrt(se / (double)nnz);
    }
}

}  // namespace

// shared memory one CTA of the stratified kernel needs (also used by pack.cu to size blocks)
size_t mfrec_sgd_smem_bytes(int tile_rows, int kpad, int W)
{
    size_t b = (size_t)tile_rows * kpad * 4;               // Q tile
    b += (size_t)((tile_rows + 3) & ~3) * 4;               // item biases
    b += (size_t)W * kDepth * kpad * 4;                    // P-row rings
    b += (size_t)W * kDepth * 4;                           // user-bias rings
    b += (size_t)W * kRing * sizeof(PackedRating);         // rating rings
    b += (size_t)(1 + kStages * W) * 8;                    // mbarriers
    b += (size_t)(W * W + 1) * 8;                          // bucket offsets
    b += (size_t)((W * W + 1) & ~1) * 4;                   // bucket counts
    b += (size_t)W * 8;                                    // per-warp squared error
    return b + 128;
}

namespace {

template <int E>
int launch_sgd(mfrec_ctx *ctx, int kernel, const SgdParams &prm, size_t smem)
{
    auto fn = kernel == MFREC_KERNEL_LINEAR ? sgd_block_kernel<E, MFREC_KERNEL_LINEAR>
                                            : sgd_block_kernel<E, MFREC_KERNEL_LOGISTIC>;
    static size_t configured[2] = {0, 0};  // per instantiation (E) and kernel
    if (configured[kernel] < smem) {
        MF_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[kernel] = smem;
    }
    fn<<<prm.B, prm.W * 32, smem, ctx->stream>>>(prm);
    MF_LAUNCH_CHECK(ctx);
    return MFREC_OK;
}

}  // namespace

extern "C" int mfrec_sgd_epoch(mfrec_ctx *ctx, const mfrec_ratings *r, mfrec_model *m, int kernel,
                               double learning_rate, double K_users, double K_items, double K_bias,
                               int update_users, int update_items, int32_t slab, double *sq_err_out)
{
    if (!ctx || !r || !m) return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_sgd_epoch: NULL argument");
    if (kernel != MFREC_KERNEL_LINEAR && kernel != MFREC_KERNEL_LOGISTIC)
        return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_sgd_epoch: kernel=%d", kernel);
    if (m->ni != r->ni || m->nu != r->nu || !m->user_perm)
        return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_sgd_epoch: model was not created with this layout");
    if (slab >= r->G) return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_sgd_epoch: slab=%d of %d", slab, r->G);
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    const int s_lo = slab < 0 ? 0 : slab, s_hi = slab < 0 ? r->G : slab + 1;
    const int64_t nparts = (int64_t)(s_hi - s_lo) * r->B * r->B;
    if (ctx->se_cap < (size_t)nparts) {
        if (ctx->se_scratch) cudaFree(ctx->se_scratch);
        ctx->se_scratch = nullptr;
        ctx->se_cap = 0;
        MF_CUDA(ctx, cudaMalloc((void **)&ctx->se_scratch, (size_t)nparts * 8));
        ctx->se_cap = (size_t)nparts;
    }
    SgdParams prm;
    prm.packed = r->packed;
    prm.bucket_off = r->bucket_off;
    prm.bucket_cnt = r->bucket_cnt;
    prm.col_start = r->col_start;
    prm.Q = m->Q; prm.ib = m->ib; prm.P = m->P; prm.ub = m->ub;
    prm.B = r->B; prm.W = r->W;
    prm.tile_rows = r->max_cb_items;
    prm.lr = (float)learning_rate; prm.Ku = (float)K_users; prm.Ki = (float)K_items; prm.Kb = (float)K_bias;
    prm.update_users = update_users; prm.update_items = update_items;
    {
        // The dot product's cross-lane sum runs in 32-bit fixed point (REDUX).  Predictions live
        // on the rating scale, so allow |dot| up to 16 x max|rating| (at least 16) before the
        // integer sum wraps; the resolution is then <= 2^-23 of that range (fp32-like).
        const float range = 16.f * fmaxf(r->max_abs_rating, 1.f);
        int ex = 0;
        frexpf(range, &ex);              // range <= 2^ex
        prm.fx_scale = ldexpf(1.f, 30 - ex);
        prm.fx_inv = ldexpf(1.f, ex - 30);
    }
    const size_t smem = mfrec_sgd_smem_bytes(r->max_cb_items, m->kpad, r->W);
    if (smem > ctx->smem_optin)
        return mfrec_set_error(ctx, MFREC_ERR_UNSUPPORTED,
                               "mfrec_sgd_epoch: Q tile of %d rows x %d needs %zu B shared memory (> %zu); pack with more row_blocks",
                               r->max_cb_items, m->kpad, smem, ctx->smem_optin);
    // debug: MFREC_SGD_TIMING=1 prints per-section cycle counts of the first launch to stderr
    static int timing_env = -1;
    if (timing_env < 0) timing_env = getenv("MFREC_SGD_TIMING") ? 1 : 0;
    DevBuf<unsigned long long> d_timing;
    prm.timing = nullptr;
    if (timing_env) {
        MF_CUDA(ctx, d_timing.alloc((size_t)r->B * r->W * 8));
        MF_CUDA(ctx, cudaMemsetAsync(d_timing.p, 0, (size_t)r->B * r->W * 64, ctx->stream));
        prm.timing = d_timing.p;
    }
    int64_t part = 0;
    for (int g = s_lo; g < s_hi; ++g) {
        for (int s = 0; s < r->B; ++s) {
            prm.slab = g;
            prm.s = s;
            prm.se_part = ctx->se_scratch + part;
            part += r->B;
            int rc;
            switch (m->kpad) {
            case 32: rc = launch_sgd<1>(ctx, kernel, prm, smem); break;
            case 64: rc = launch_sgd<2>(ctx, kernel, prm, smem); break;
            case 128: rc = launch_sgd<4>(ctx, kernel, prm, smem); break;
            case 256: rc = launch_sgd<8>(ctx, kernel, prm, smem); break;
            default: rc = mfrec_set_error(ctx, MFREC_ERR_UNSUPPORTED, "kpad=%d", m->kpad);
            }
            MF_TRY(rc);
            if (timing_env) {
                std::vector<unsigned long long> h((size_t)r->B * r->W * 8);
                MF_CUDA(ctx, cudaMemcpyAsync(h.data(), d_timing.p, h.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
                MF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                double avg[8] = {0}, mx[8] = {0};
                double worst_tot = 0; int worst = 0;
                for (int c = 0; c < r->B * r->W; ++c) {
                    double tot = 0;
                    for (int k2 = 0; k2 < 7; ++k2) { avg[k2] += (double)h[(size_t)c * 8 + k2]; tot += (double)h[(size_t)c * 8 + k2]; }
                    if (tot - (double)h[(size_t)c * 8 + 6] > worst_tot) { worst_tot = tot - (double)h[(size_t)c * 8 + 6]; worst = c; }
                }
                fprintf(stderr, "[sgd timing] slab %d sub-epoch %d: avg cycles/warp by section:", g, s);
                for (int k2 = 0; k2 < 7; ++k2) fprintf(stderr, " %.0f", avg[k2] / (r->B * r->W));
                fprintf(stderr, " | busiest warp (cta %d warp %d):", worst / r->W, worst % r->W);
                for (int k2 = 0; k2 < 7; ++k2) fprintf(stderr, " %llu", h[(size_t)worst * 8 + k2]);
                fprintf(stderr, "\n");
                (void)mx;
                if (s >= 2) timing_env = 0;   // three launches are enough
            }
        }
    }
    if (sq_err_out) {
        se_reduce_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->se_scratch, nparts, sq_err_out);
        MF_LAUNCH_CHECK(ctx);
    }
    return MFREC_OK;
}

static int train_kmf_sequential(mfrec_ctx *ctx, int kernel, int nbr_epochs, int k, double lr,
                                double K_users, double K_items, double K_bias, double *u, double *v,
                                const int32_t *idx, const double *ratings, int64_t nnz, int32_t ni,
                                int32_t nu, double *ib, double *ub, int update_users,
                                int update_items, double *rmse_per_epoch)
{
    cudaStream_t st = ctx->stream;
    DevBuf<double> du, dv, dr, dib, dub, drm;
    DevBuf<int32_t> didx;
    MF_CUDA(ctx, du.alloc((size_t)k * ni));
    MF_CUDA(ctx, dv.alloc((size_t)k * nu));
    MF_CUDA(ctx, dr.alloc((size_t)nnz));
    MF_CUDA(ctx, didx.alloc((size_t)nnz * 2));
    MF_CUDA(ctx, dib.alloc(ni));
    MF_CUDA(ctx, dub.alloc(nu));
    MF_CUDA(ctx, drm.alloc(nbr_epochs > 0 ? nbr_epochs : 1));
    MF_CUDA(ctx, cudaMemcpyAsync(du.p, u, (size_t)k * ni * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(dv.p, v, (size_t)k * nu * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(dr.p, ratings, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(didx.p, idx, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(dib.p, ib, (size_t)ni * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(dub.p, ub, (size_t)nu * 8, cudaMemcpyHostToDevice, st));
    kmf_sequential_kernel<<<1, 1, 0, st>>>(kernel, nbr_epochs, k, lr, K_users, K_items, K_bias, du.p,
                                           dv.p, didx.p, dr.p, nnz, ni, nu, dib.p, dub.p,
                                           update_users, update_items, drm.p);
    MF_LAUNCH_CHECK(ctx);
    MF_CUDA(ctx, cudaMemcpyAsync(u, du.p, (size_t)k * ni * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaMemcpyAsync(v, dv.p, (size_t)k * nu * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaMemcpyAsync(ib, dib.p, (size_t)ni * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaMemcpyAsync(ub, dub.p, (size_t)nu * 8, cudaMemcpyDeviceToHost, st));
    if (rmse_per_epoch && nbr_epochs > 0)
        MF_CUDA(ctx, cudaMemcpyAsync(rmse_per_epoch, drm.p, (size_t)nbr_epochs * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaStreamSynchronize(st));
    return MFREC_OK;
}

extern "C" int mfrec_train_kmf(mfrec_ctx *ctx, int kernel, int nbr_epochs, int k, double learning_rate,
                               double K_users, double K_items, double K_bias, double *u, double *v,
                               const int32_t *ratings_index, const double *ratings, int64_t nnz,
                               int32_t ni, int32_t nu, double *items_bias, double *users_bias,
                               int update_users, int update_items, const mfrec_opts *opts,
                               double *rmse_per_epoch)
{
    if (!ctx) return mfrec_set_error(nullptr, MFREC_ERR_BAD_ARG, "mfrec_train_kmf: NULL ctx");
    if (!u || !v || !items_bias || !users_bias || (nnz > 0 && (!ratings_index || !ratings)))
        return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_train_kmf: NULL array");
    if (kernel != MFREC_KERNEL_LINEAR && kernel != MFREC_KERNEL_LOGISTIC)
        return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_train_kmf: kernel=%d", kernel);
    if (k <= 0 || ni <= 0 || nu <= 0 || nnz < 0 || nbr_epochs < 0)
        return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_train_kmf: k=%d ni=%d nu=%d nnz=%lld epochs=%d",
                               k, ni, nu, (long long)nnz, nbr_epochs);
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    if (nbr_epochs == 0) return MFREC_OK;
    if (nnz == 0) {
        // the reference divides 0/0 -> rmse = NaN and leaves the model untouched
        if (rmse_per_epoch)
            for (int e = 0; e < nbr_epochs; ++e) rmse_per_epoch[e] = NAN;
        return MFREC_OK;
    }
    if (opts && opts->schedule == MFREC_SCHED_SEQUENTIAL) {
        for (int64_t n = 0; n < nnz; ++n) {
            const int32_t a = ratings_index[2 * n], b = ratings_index[2 * n + 1];
            if (a < 0 || a >= nu || b < 0 || b >= ni)
                return mfrec_set_error(ctx, MFREC_ERR_INDEX, "mfrec_train_kmf: rating %lld has (user,item)=(%d,%d)",
                                       (long long)n, a, b);
        }
        return train_kmf_sequential(ctx, kernel, nbr_epochs, k, learning_rate, K_users, K_items, K_bias,
                                    u, v, ratings_index, ratings, nnz, ni, nu, items_bias, users_bias,
                                    update_users, update_items, rmse_per_epoch);
    }
    if (mfrec_kpad(k) < 0)
        return mfrec_set_error(ctx, MFREC_ERR_UNSUPPORTED, "mfrec_train_kmf: k=%d > 256 is not instantiated", k);
    mfrec_opts o;
    if (opts) o = *opts; else memset(&o, 0, sizeof(o));
    o.n_slabs = 1;
    if (o.k_hint == 0) o.k_hint = k;
    mfrec_ratings *R = nullptr;
    mfrec_model *M = nullptr;
    MF_TRY(mfrec_ratings_pack(ctx, ratings_index, ratings, 0, 0, nnz, ni, nu, nullptr, &o, &R));
    int rc = mfrec_model_create(ctx, R, k, ni, nu, u, v, items_bias, users_bias, &M);
    DevBuf<double> d_se;
    if (rc == MFREC_OK && d_se.alloc(nbr_epochs) != cudaSuccess)
        rc = mfrec_set_error(ctx, MFREC_ERR_OOM, "mfrec_train_kmf: device OOM");
    for (int e = 0; rc == MFREC_OK && e < nbr_epochs; ++e)
        rc = mfrec_sgd_epoch(ctx, R, M, kernel, learning_rate, K_users, K_items, K_bias, update_users,
                             update_items, -1, d_se.p + e);
    if (rc == MFREC_OK) {
        std::vector<double> h_se(nbr_epochs);
        cudaError_t ce = cudaMemcpyAsync(h_se.data(), d_se.p, (size_t)nbr_epochs * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
        if (ce != cudaSuccess) rc = mfrec_set_error(ctx, MFREC_ERR_CUDA, "mfrec_train_kmf: %s", cudaGetErrorString(ce));
        else if (rmse_per_epoch)
            for (int e = 0; e < nbr_epochs; ++e) rmse_per_epoch[e] = sqrt(h_se[e] / (double)nnz);
    }
    if (rc == MFREC_OK) rc = mfrec_model_read(ctx, M, u, v, items_bias, users_bias);
    mfrec_model_destroy(M);
    mfrec_ratings_destroy(R);
    return rc;
}

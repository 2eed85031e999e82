// One-call drop-in over SEVERAL GPUs of one process: train_linear_kernel / train_logistic_kernel
// (mfrec/lib/kmf_train.pyx:195-277, 103-189) with the reference's host arrays in and out, the
// work spread over a DSGD ring (sgd.cu, mfrec_ring_*): device d keeps user slice d and in step t
// of an epoch holds item slab (d + t) mod D; finished column blocks go to the next device through
// directly addressed peer memory (cudaDeviceEnablePeerAccess, NVLink).  This is what
// KMFRecommender.train reaches when `mfrec.lib.kmf_train.options["devices"]` lists more than one
// device; one process per GPU (bench.py under torchrun) uses the same ring through cudaIpc.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

#include "common.cuh"

namespace {

struct Slice {
    int device = 0;
    int32_t u_lo = 0, u_hi = 0;            // users [u_lo, u_hi)
    std::vector<int32_t> idx;              // [nnz_d][2], user ids relative to u_lo
    std::vector<double> r, v;              // ratings; user factors [k][nu_d]
    mfrec_ctx *ctx = nullptr;
    mfrec_ratings *R = nullptr;
    mfrec_model *M = nullptr;
    mfrec_ring *ring = nullptr;
    double *d_se = nullptr;
    int rc = MFREC_OK;
    std::string err;
    std::vector<double> u_out, ib_out;     // this device's view of the item side after training
    std::vector<int32_t> item_rows;
};

void release(std::vector<Slice> &S)
{
    for (Slice &s : S) {
        if (s.ctx) cudaSetDevice(s.device);
        if (s.ring) mfrec_ring_destroy(s.ring);
        if (s.d_se) cudaFree(s.d_se);
        if (s.M) mfrec_model_destroy(s.M);
        if (s.R) mfrec_ratings_destroy(s.R);
        if (s.ctx) mfrec_ctx_destroy(s.ctx);
        s.ring = nullptr; s.d_se = nullptr; s.M = nullptr; s.R = nullptr; s.ctx = nullptr;
    }
}

}  // namespace

extern "C" int mfrec_train_kmf_multi(const int32_t *devices, int n_dev, int kernel, int nbr_epochs, int k,
                                     double learning_rate, double K_users, double K_items, double K_bias,
                                     double *u, double *v, const int32_t *ratings_index, const double *ratings,
                                     int64_t nnz, int32_t ni, int32_t nu, double *items_bias, double *users_bias,
                                     const mfrec_opts *opts, double *rmse_per_epoch)
{
    if (!devices || n_dev < 1)
        return mfrec_set_error(nullptr, MFREC_ERR_BAD_ARG, "mfrec_train_kmf_multi: no devices");
    if (!u || !v || !items_bias || !users_bias || (nnz > 0 && (!ratings_index || !ratings)))
        return mfrec_set_error(nullptr, MFREC_ERR_BAD_ARG, "mfrec_train_kmf_multi: NULL array");
    if (k <= 0 || ni <= 0 || nu <= 0 || nnz < 0 || nbr_epochs < 0)
        return mfrec_set_error(nullptr, MFREC_ERR_BAD_ARG, "mfrec_train_kmf_multi: k=%d ni=%d nu=%d nnz=%lld epochs=%d",
                               k, ni, nu, (long long)nnz, nbr_epochs);
    for (int a = 0; a < n_dev; ++a)
        for (int b = a + 1; b < n_dev; ++b)
            if (devices[a] == devices[b])
                return mfrec_set_error(nullptr, MFREC_ERR_BAD_ARG, "mfrec_train_kmf_multi: device %d listed twice", devices[a]);
    if (n_dev == 1 || nnz == 0 || nbr_epochs == 0 || nu < 2 * n_dev) {
        mfrec_ctx *ctx = nullptr;
        MF_TRY(mfrec_ctx_create(devices[0], &ctx));
        const int rc = mfrec_train_kmf(ctx, kernel, nbr_epochs, k, learning_rate, K_users, K_items, K_bias, u, v,
                                       ratings_index, ratings, nnz, ni, nu, items_bias, users_bias, 1, 1, opts,
                                       rmse_per_epoch);
        if (rc != MFREC_OK) mfrec_set_error(nullptr, rc, "%s", mfrec_last_error(ctx));
        mfrec_ctx_destroy(ctx);
        return rc;
    }
    const int D = n_dev;
    // ---- user slices of equal rating count, global item degrees (host, one pass each) ------------
    std::vector<int64_t> deg_u((size_t)nu + 1, 0), deg_i(ni, 0);
    for (int64_t n = 0; n < nnz; ++n) {
        const int32_t a = ratings_index[2 * n], b = ratings_index[2 * n + 1];
        if (a < 0 || a >= nu || b < 0 || b >= ni)
            return mfrec_set_error(nullptr, MFREC_ERR_INDEX, "mfrec_train_kmf_multi: rating %lld has (user,item)=(%d,%d)",
                                   (long long)n, a, b);
        deg_u[a + 1] += 1;
        deg_i[b] += 1;
    }
    for (int32_t j = 0; j < nu; ++j) deg_u[j + 1] += deg_u[j];   // cumulative
    std::vector<Slice> S(D);
    std::vector<int32_t> slice_of(nu);
    {
        int32_t lo = 0;
        for (int d = 0; d < D; ++d) {
            int32_t hi = nu;
            if (d + 1 < D) {
                const int64_t target = nnz * (int64_t)(d + 1) / D;
                hi = (int32_t)(std::lower_bound(deg_u.begin(), deg_u.end(), target) - deg_u.begin());
                hi = std::max(lo + 1, std::min(hi, nu - (D - 1 - d)));
            }
            S[d].device = devices[d];
            S[d].u_lo = lo;
            S[d].u_hi = hi;
            for (int32_t j = lo; j < hi; ++j) slice_of[j] = d;
            lo = hi;
        }
    }
    for (int d = 0; d < D; ++d) {
        const int64_t cnt = deg_u[S[d].u_hi] - deg_u[S[d].u_lo];
        S[d].idx.resize((size_t)cnt * 2);
        S[d].r.resize((size_t)cnt);
    }
    {
        std::vector<int64_t> at(D, 0);
        for (int64_t n = 0; n < nnz; ++n) {   // input order is kept inside every slice
            const int32_t a = ratings_index[2 * n];
            Slice &s = S[slice_of[a]];
            int64_t &p = at[slice_of[a]];
            s.idx[2 * p] = a - s.u_lo;
            s.idx[2 * p + 1] = ratings_index[2 * n + 1];
            s.r[p] = ratings[n];
            ++p;
        }
    }
    mfrec_opts o;
    if (opts) o = *opts; else memset(&o, 0, sizeof(o));
    o.schedule = MFREC_SCHED_STRATIFIED;
    o.n_slabs = D;
    if (o.k_hint == 0) o.k_hint = k;
    // ---- per device: context, layout, model, ring (one host thread each) -------------------------
    auto setup = [&](int d) {
        Slice &s = S[d];
        const int32_t nu_d = s.u_hi - s.u_lo;
        s.v.resize((size_t)k * nu_d);
        for (int f = 0; f < k; ++f) memcpy(&s.v[(size_t)f * nu_d], v + (size_t)f * nu + s.u_lo, (size_t)nu_d * 8);
        int rc = mfrec_ctx_create(s.device, &s.ctx);
        if (rc == MFREC_OK)
            rc = mfrec_ratings_pack(s.ctx, s.idx.data(), s.r.data(), 0, 0, (int64_t)s.r.size(), ni, nu_d, deg_i.data(), &o, &s.R);
        if (rc == MFREC_OK)
            rc = mfrec_model_create(s.ctx, s.R, k, ni, nu_d, u, s.v.data(), items_bias, users_bias + s.u_lo, &s.M);
        if (rc == MFREC_OK) rc = mfrec_ring_create(s.ctx, s.R, s.M, d, D, &s.ring);
        if (rc == MFREC_OK && cudaMalloc((void **)&s.d_se, (size_t)nbr_epochs * 8) != cudaSuccess) rc = MFREC_ERR_OOM;
        s.rc = rc;
        if (rc != MFREC_OK) s.err = mfrec_last_error(s.ctx);
        std::vector<int32_t>().swap(s.idx);   // the packed copy lives in HBM now
        std::vector<double>().swap(s.r);
    };
    {
        std::vector<std::thread> th;
        for (int d = 1; d < D; ++d) th.emplace_back(setup, d);
        setup(0);
        for (auto &t : th) t.join();
    }
    auto fail = [&](int rc, const std::string &msg) {
        release(S);
        return mfrec_set_error(nullptr, rc, "mfrec_train_kmf_multi: %s", msg.c_str());
    };
    for (int d = 0; d < D; ++d)
        if (S[d].rc != MFREC_OK) return fail(S[d].rc, "device " + std::to_string(S[d].device) + ": " + S[d].err);
    for (int d = 1; d < D; ++d)   // every device derived the item layout from the same global degrees
        if (S[d].R->B != S[0].R->B || S[d].R->W != S[0].R->W || S[d].R->max_cb_items != S[0].R->max_cb_items ||
            S[d].R->h_col_start != S[0].R->h_col_start)
            return fail(MFREC_ERR_UNSUPPORTED, "the devices disagree on the item layout (pass opts.row_blocks / opts.workers)");
    for (int d = 0; d < D; ++d) {
        const int rc = mfrec_ring_connect_local(S[d].ring, S[(d + D - 1) % D].ring);
        if (rc != MFREC_OK) return fail(rc, mfrec_last_error(S[d].ctx));
    }
    // ---- all epochs: every device's launches are enqueued before anything is waited for -----------
    for (int d = 0; d < D; ++d) {
        const int rc = mfrec_ring_epochs(S[d].ring, kernel, learning_rate, K_users, K_items, K_bias, nbr_epochs, S[d].d_se);
        if (rc != MFREC_OK) return fail(rc, mfrec_last_error(S[d].ctx));
    }
    std::vector<double> se((size_t)nbr_epochs, 0.0), se_d((size_t)nbr_epochs);
    for (int d = 0; d < D; ++d) {
        int rc = mfrec_ring_wait(S[d].ring);
        if (rc != MFREC_OK) return fail(rc, mfrec_last_error(S[d].ctx));
        cudaSetDevice(S[d].device);
        if (cudaMemcpy(se_d.data(), S[d].d_se, (size_t)nbr_epochs * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail(MFREC_ERR_CUDA, "reading the error sums failed");
        for (int e = 0; e < nbr_epochs; ++e) se[e] += se_d[e];
    }
    // ---- read back: device d holds slab d and its own users ------------------------------------------
    auto download = [&](int d) {
        Slice &s = S[d];
        const int32_t nu_d = s.u_hi - s.u_lo;
        s.u_out.resize((size_t)k * ni);
        s.ib_out.resize(ni);
        s.item_rows.resize(ni);
        int rc = mfrec_ring_sync_model(s.ring);
        if (rc == MFREC_OK) rc = mfrec_model_read(s.ctx, s.M, s.u_out.data(), s.v.data(), s.ib_out.data(), users_bias + s.u_lo);
        if (rc == MFREC_OK) rc = mfrec_ratings_perm(s.ctx, s.R, nullptr, s.item_rows.data());
        s.rc = rc;
        if (rc != MFREC_OK) { s.err = mfrec_last_error(s.ctx); return; }
        for (int f = 0; f < k; ++f) memcpy(v + (size_t)f * nu + s.u_lo, &s.v[(size_t)f * nu_d], (size_t)nu_d * 8);
        int32_t a = 0, b = 0;
        mfrec_ratings_slab_items(s.R, d, &a, &b);
        for (int32_t i = 0; i < ni; ++i) {
            if (s.item_rows[i] < a || s.item_rows[i] >= b) continue;   // another device holds this item's slab
            items_bias[i] = s.ib_out[i];
            for (int f = 0; f < k; ++f) u[(size_t)f * ni + i] = s.u_out[(size_t)f * ni + i];
        }
    };
    {
        std::vector<std::thread> th;
        for (int d = 1; d < D; ++d) th.emplace_back(download, d);
        download(0);
        for (auto &t : th) t.join();
    }
    for (int d = 0; d < D; ++d)
        if (S[d].rc != MFREC_OK) return fail(S[d].rc, "device " + std::to_string(S[d].device) + ": " + S[d].err);
    if (rmse_per_epoch)
        for (int e = 0; e < nbr_epochs; ++e) rmse_per_epoch[e] = sqrt(se[e] / (double)nnz);
    release(S);
    return MFREC_OK;
}

// The development variants of the Funk loop (priority A3):
//   estimator_loop                      mfrec/lib/gd_estimator.pyx:210-303
//   estimator_loop2                     :308-395
//   estimator_subloop / predictor_subloop  :903-962 / :967-995
//   estimator_loop_with_learned_bias    :401-483 (+ full_estimator :115-148)
// They keep a DENSE rating cache indexed `user + item * nbr_users` (toy sizes by construction)
// or, for the learned-bias loop, recompute a full clamped k-dot per rating; the reference itself
// only drives them from its *_dev / feature_training2 / feature_training_bias entry points.
// Here they run in the reference's own order: one thread, float64, unfused multiplies and adds,
// bit-identical to the reference.  Correctness and API coverage, not speed: the production Funk
// path is funk.cu.
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {

__device__ __forceinline__ double dclamp15(double x)
{
    if (x > 5.0) x = 5.0;
    if (x < 1.0) x = 1.0;
    return x;
}

// gd_estimator.pyx:38-73 with the defaults the A3 loops use (overall 1.0, no biases)
__device__ __forceinline__ double dev_estimate(double uf, double vf, double cache, double trail, int trailing)
{
    double s = cache > 0 ? cache : 1.0;
    s = __dadd_rn(s, __dmul_rn(uf, vf));
    s = dclamp15(s);
    if (trailing) {
        s = __dadd_rn(s, trail);
        s = dclamp15(s);
    }
    return s;
}

// one training pass of feature f over all ratings (the body shared by :266-282, :352-366, :944-958)
__device__ double dev_train_pass(double *uf, double *vf, const int32_t *idx, const double *ratings,
                                 int64_t nnz, int64_t nu, const double *cache, double trail, double lr, double K)
{
    double se = 0.0;
    for (int64_t n = 0; n < nnz; ++n) {
        const int user = idx[2 * n], item = idx[2 * n + 1];
        const double p = dev_estimate(uf[item], vf[user], cache[user + (int64_t)item * nu], trail, 1);
        const double err = __dadd_rn(__dmul_rn(1.0, ratings[n]), -p);
        se = __dadd_rn(se, __dmul_rn(err, err));
        const double cf = vf[user], mf = uf[item];
        vf[user] = __dadd_rn(cf, __dmul_rn(lr, __dadd_rn(__dmul_rn(err, mf), -__dmul_rn(K, cf))));
        uf[item] = __dadd_rn(mf, __dmul_rn(lr, __dadd_rn(__dmul_rn(err, cf), -__dmul_rn(K, mf))));
    }
    return se;
}

__device__ void dev_refresh(const double *uf, const double *vf, const int32_t *idx, int64_t nnz, int64_t nu,
                            double *cache)
{
    for (int64_t n = 0; n < nnz; ++n) {
        const int user = idx[2 * n], item = idx[2 * n + 1];
        double *c = &cache[user + (int64_t)item * nu];
        *c = dev_estimate(uf[item], vf[user], *c, 0.0, 0);
    }
}

__global__ void funk_loop_dev_kernel(int min_epochs, int max_epochs, double min_improvement, int dim,
                                     double f_init, double lr, double K, double *u, double *v,
                                     const int32_t *idx, const double *ratings, int64_t nnz, int64_t ni,
                                     int64_t nu, double *cache, double *hist /* [dim * max_epochs] or null */,
                                     int32_t *feature_epochs, double *feature_rmse)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double rmse = 2.0, rmse_last = 0.0, improvement = 0.0;
    const bool with_hist = max_epochs >= 0;
    for (int f = 0; f < dim; ++f) {
        double *uf = u + (int64_t)f * ni, *vf = v + (int64_t)f * nu;
        const double trail = __dmul_rn(__dmul_rn((double)(dim - f - 1), f_init), f_init);
        int epoch = 0;
        while (with_hist ? ((epoch < min_epochs || improvement >= min_improvement) && epoch < max_epochs)
                         : (epoch < min_epochs || rmse <= __dadd_rn(rmse_last, -min_improvement))) {
            rmse_last = rmse;
            const double se = dev_train_pass(uf, vf, idx, ratings, nnz, nu, cache, trail, lr, K);
            rmse = sqrt(se / (double)nnz);
            if (with_hist) {
                hist[epoch + (int64_t)f * max_epochs] = rmse;
                improvement = __dadd_rn(rmse_last, -rmse);
            }
            ++epoch;
        }
        feature_epochs[f] = epoch;
        feature_rmse[f] = rmse;
        dev_refresh(uf, vf, idx, nnz, nu, cache);
    }
}

__global__ void funk_subloop_kernel(double trail, double lr, double K, double *uf, double *vf,
                                    const int32_t *idx, const double *ratings, int64_t nnz, int64_t nu,
                                    const double *cache, double *rmse_out)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const double se = dev_train_pass(uf, vf, idx, ratings, nnz, nu, cache, trail, lr, K);
    *rmse_out = sqrt(se / (double)nnz);
}

__global__ void funk_predictor_subloop_kernel(const double *uf, const double *vf, const int32_t *idx,
                                              int64_t nnz, int64_t nu, double *cache)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    dev_refresh(uf, vf, idx, nnz, nu, cache);
}

__global__ void funk_learned_bias_kernel(int min_epochs, double min_improvement, int dim, double f_init,
                                         double lr, double lr_users, double lr_items, double K_feature,
                                         double K_bias, double overall, double *u, double *v,
                                         const int32_t *idx, const double *ratings, int64_t nnz, int64_t ni,
                                         int64_t nu, double *ib, double *ub, int32_t *feature_epochs,
                                         double *feature_rmse)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double rmse = 2.0, rmse_last = 0.0;
    for (int f = 0; f < dim; ++f) {
        double *uf = u + (int64_t)f * ni, *vf = v + (int64_t)f * nu;
        const double trail = __dmul_rn(__dmul_rn((double)(dim - f - 1), f_init), f_init);
        int epoch = 0;
        while (epoch < min_epochs || rmse <= __dadd_rn(rmse_last, -min_improvement)) {
            double se = 0.0;
            rmse_last = rmse;
            for (int64_t n = 0; n < nnz; ++n) {
                const int user = idx[2 * n], item = idx[2 * n + 1];
                double s = __dadd_rn(__dadd_rn(overall, ib[item]), ub[user]);      // :138
                for (int g = 0; g < dim; ++g)
                    s = __dadd_rn(s, __dmul_rn(u[(int64_t)g * ni + item], v[(int64_t)g * nu + user]));
                s = dclamp15(s);
                s = __dadd_rn(s, trail);
                s = dclamp15(s);
                const double err = __dadd_rn(ratings[n], -s);
                se = __dadd_rn(se, __dmul_rn(err, err));
                const double cf = vf[user], mf = uf[item];
                ub[user] = __dadd_rn(ub[user], __dmul_rn(lr_users, __dadd_rn(err, -__dmul_rn(K_bias, ub[user]))));
                ib[item] = __dadd_rn(ib[item], __dmul_rn(lr_items, __dadd_rn(err, -__dmul_rn(K_bias, ib[item]))));
                uf[item] = __dadd_rn(mf, __dmul_rn(lr, __dadd_rn(__dmul_rn(err, cf), -__dmul_rn(K_feature, mf))));
                vf[user] = __dadd_rn(cf, __dmul_rn(lr, __dadd_rn(__dmul_rn(err, mf), -__dmul_rn(K_feature, cf))));
            }
            rmse = sqrt(se / (double)nnz);
            ++epoch;
        }
        feature_epochs[f] = epoch;
        feature_rmse[f] = rmse;
    }
}

int check_pairs(mfrec_ctx *ctx, const char *who, const int32_t *idx, int64_t nnz, int32_t ni, int32_t nu)
{
    for (int64_t n = 0; n < nnz; ++n) {
        const int32_t a = idx[2 * n], b = idx[2 * n + 1];
        if (a < 0 || a >= nu || b < 0 || b >= ni)
            return mfrec_set_error(ctx, MFREC_ERR_INDEX, "%s: rating %lld has (user,item)=(%d,%d)", who, (long long)n, a, b);
    }
    return MFREC_OK;
}

int check_dense(mfrec_ctx *ctx, const char *who, int32_t ni, int32_t nu)
{
    if ((int64_t)ni * nu > ((int64_t)1 << 29))
        return mfrec_set_error(ctx, MFREC_ERR_UNSUPPORTED,
                               "%s: the dense rating cache of this loop needs %lld x 8 bytes; use estimator_loop_without_bias",
                               who, (long long)ni * nu);
    return MFREC_OK;
}

}  // namespace

extern "C" int mfrec_funk_loop_dev(mfrec_ctx *ctx, int min_epochs, int max_epochs, double min_improvement,
                                   int k, double f_init, double learning_rate, double K, double *u, double *v,
                                   const int32_t *ratings_index, const double *ratings, int64_t nnz,
                                   int32_t ni, int32_t nu, int batch, double *rmse_hist,
                                   int32_t *feature_epochs, double *feature_rmse)
{
    const char *who = "mfrec_funk_loop_dev";
    if (!ctx) return mfrec_set_error(nullptr, MFREC_ERR_BAD_ARG, "%s: NULL ctx", who);
    if (!u || !v || !ratings_index || !ratings || k <= 0 || ni <= 0 || nu <= 0 || nnz <= 0 || batch < 0 ||
        (max_epochs >= 0 && !rmse_hist))
        return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "%s: bad argument", who);
    MF_TRY(check_pairs(ctx, who, ratings_index, nnz, ni, nu));
    MF_TRY(check_dense(ctx, who, ni, nu));
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevBuf<double> du, dv, dr, dcache, dhist, dfr;
    DevBuf<int32_t> didx, dfe;
    const size_t nh = max_epochs > 0 ? (size_t)max_epochs * k : 1;
    MF_CUDA(ctx, du.alloc((size_t)k * ni, st));
    MF_CUDA(ctx, dv.alloc((size_t)k * nu, st));
    MF_CUDA(ctx, dr.alloc(nnz, st));
    MF_CUDA(ctx, didx.alloc((size_t)nnz * 2, st));
    MF_CUDA(ctx, dcache.alloc((size_t)ni * nu, st));
    MF_CUDA(ctx, dhist.alloc(nh, st));
    MF_CUDA(ctx, dfe.alloc(k, st));
    MF_CUDA(ctx, dfr.alloc(k, st));
    MF_CUDA(ctx, cudaMemcpyAsync(du.p, u, (size_t)k * ni * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(dv.p, v, (size_t)k * nu * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(dr.p, ratings, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(didx.p, ratings_index, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemsetAsync(dcache.p, 0, (size_t)ni * nu * 8, st));
    funk_loop_dev_kernel<<<1, 1, 0, st>>>(min_epochs, max_epochs, min_improvement, k, f_init, learning_rate, K,
                                          du.p, dv.p, didx.p, dr.p, nnz, ni, nu, dcache.p, dhist.p, dfe.p, dfr.p);
    MF_LAUNCH_CHECK(ctx);
    std::vector<int32_t> h_fe(k);
    std::vector<double> h_fr(k), h_hist(nh);
    MF_CUDA(ctx, cudaMemcpyAsync(h_fe.data(), dfe.p, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaMemcpyAsync(h_fr.data(), dfr.p, (size_t)k * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaMemcpyAsync(h_hist.data(), dhist.p, nh * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaMemcpyAsync(u, du.p, (size_t)k * ni * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaMemcpyAsync(v, dv.p, (size_t)k * nu * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += 1;
    if (max_epochs >= 0)   // only the entries the reference writes: rmse_hist[epoch + f*max + batch*max*dim]
        for (int f = 0; f < k; ++f)
            for (int e = 0; e < h_fe[f]; ++e)
                rmse_hist[e + (int64_t)f * max_epochs + (int64_t)batch * max_epochs * k] = h_hist[e + (size_t)f * max_epochs];
    if (feature_epochs) memcpy(feature_epochs, h_fe.data(), (size_t)k * 4);
    if (feature_rmse) memcpy(feature_rmse, h_fr.data(), (size_t)k * 8);
    return MFREC_OK;
}

extern "C" int mfrec_funk_subloop(mfrec_ctx *ctx, int f, int k, double f_init, double learning_rate, double K,
                                  double *u, double *v, const int32_t *ratings_index, const double *ratings,
                                  int64_t nnz, int32_t ni, int32_t nu, const double *rating_cache,
                                  double *rmse_out)
{
    const char *who = "mfrec_funk_subloop";
    if (!ctx) return mfrec_set_error(nullptr, MFREC_ERR_BAD_ARG, "%s: NULL ctx", who);
    if (!u || !v || !ratings_index || !ratings || !rating_cache || !rmse_out || k <= 0 || f < 0 || f >= k ||
        ni <= 0 || nu <= 0 || nnz <= 0)
        return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "%s: bad argument", who);
    MF_TRY(check_pairs(ctx, who, ratings_index, nnz, ni, nu));
    MF_TRY(check_dense(ctx, who, ni, nu));
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevBuf<double> duf, dvf, dr, dcache, drm;
    DevBuf<int32_t> didx;
    MF_CUDA(ctx, duf.alloc(ni, st));
    MF_CUDA(ctx, dvf.alloc(nu, st));
    MF_CUDA(ctx, dr.alloc(nnz, st));
    MF_CUDA(ctx, didx.alloc((size_t)nnz * 2, st));
    MF_CUDA(ctx, dcache.alloc((size_t)ni * nu, st));
    MF_CUDA(ctx, drm.alloc(1, st));
    double *uf = u + (size_t)f * ni, *vf = v + (size_t)f * nu;   // only feature f is touched
    MF_CUDA(ctx, cudaMemcpyAsync(duf.p, uf, (size_t)ni * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(dvf.p, vf, (size_t)nu * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(dr.p, ratings, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(didx.p, ratings_index, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(dcache.p, rating_cache, (size_t)ni * nu * 8, cudaMemcpyHostToDevice, st));
    const double trail = (double)(k - f - 1) * f_init * f_init;
    funk_subloop_kernel<<<1, 1, 0, st>>>(trail, learning_rate, K, duf.p, dvf.p, didx.p, dr.p, nnz, nu, dcache.p, drm.p);
    MF_LAUNCH_CHECK(ctx);
    MF_CUDA(ctx, cudaMemcpyAsync(uf, duf.p, (size_t)ni * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaMemcpyAsync(vf, dvf.p, (size_t)nu * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaMemcpyAsync(rmse_out, drm.p, 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += 1;
    return MFREC_OK;
}

extern "C" int mfrec_funk_predictor_subloop(mfrec_ctx *ctx, int f, int k, double f_init, const double *u,
                                            const double *v, const int32_t *ratings_index, int64_t nnz,
                                            int32_t ni, int32_t nu, double *rating_cache)
{
    (void)f_init;   // trailing = 0 in the refresh (:993-994)
    const char *who = "mfrec_funk_predictor_subloop";
    if (!ctx) return mfrec_set_error(nullptr, MFREC_ERR_BAD_ARG, "%s: NULL ctx", who);
    if (!u || !v || !ratings_index || !rating_cache || k <= 0 || f < 0 || f >= k || ni <= 0 || nu <= 0 || nnz < 0)
        return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "%s: bad argument", who);
    if (nnz == 0) return MFREC_OK;
    MF_TRY(check_pairs(ctx, who, ratings_index, nnz, ni, nu));
    MF_TRY(check_dense(ctx, who, ni, nu));
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevBuf<double> duf, dvf, dcache;
    DevBuf<int32_t> didx;
    MF_CUDA(ctx, duf.alloc(ni, st));
    MF_CUDA(ctx, dvf.alloc(nu, st));
    MF_CUDA(ctx, didx.alloc((size_t)nnz * 2, st));
    MF_CUDA(ctx, dcache.alloc((size_t)ni * nu, st));
    MF_CUDA(ctx, cudaMemcpyAsync(duf.p, u + (size_t)f * ni, (size_t)ni * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(dvf.p, v + (size_t)f * nu, (size_t)nu * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(didx.p, ratings_index, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(dcache.p, rating_cache, (size_t)ni * nu * 8, cudaMemcpyHostToDevice, st));
    funk_predictor_subloop_kernel<<<1, 1, 0, st>>>(duf.p, dvf.p, didx.p, nnz, nu, dcache.p);
    MF_LAUNCH_CHECK(ctx);
    MF_CUDA(ctx, cudaMemcpyAsync(rating_cache, dcache.p, (size_t)ni * nu * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += 1;
    return MFREC_OK;
}

extern "C" int mfrec_train_funk_learned_bias(mfrec_ctx *ctx, int min_epochs, double min_improvement, int k,
                                             double f_init, double learning_rate, double learning_rate_users,
                                             double learning_rate_items, double K_feature, double K_bias,
                                             double overall_avg, double *u, double *v,
                                             const int32_t *ratings_index, const double *ratings, int64_t nnz,
                                             int32_t ni, int32_t nu, double *items_bias, double *users_bias,
                                             int32_t *feature_epochs, double *feature_rmse)
{
    const char *who = "mfrec_train_funk_learned_bias";
    if (!ctx) return mfrec_set_error(nullptr, MFREC_ERR_BAD_ARG, "%s: NULL ctx", who);
    if (!u || !v || !ratings_index || !ratings || !items_bias || !users_bias || k <= 0 || ni <= 0 || nu <= 0 ||
        nnz <= 0)
        return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "%s: bad argument", who);
    MF_TRY(check_pairs(ctx, who, ratings_index, nnz, ni, nu));
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevBuf<double> du, dv, dr, dib, dub, dfr;
    DevBuf<int32_t> didx, dfe;
    MF_CUDA(ctx, du.alloc((size_t)k * ni, st));
    MF_CUDA(ctx, dv.alloc((size_t)k * nu, st));
    MF_CUDA(ctx, dr.alloc(nnz, st));
    MF_CUDA(ctx, didx.alloc((size_t)nnz * 2, st));
    MF_CUDA(ctx, dib.alloc(ni, st));
    MF_CUDA(ctx, dub.alloc(nu, st));
    MF_CUDA(ctx, dfe.alloc(k, st));
    MF_CUDA(ctx, dfr.alloc(k, st));
    MF_CUDA(ctx, cudaMemcpyAsync(du.p, u, (size_t)k * ni * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(dv.p, v, (size_t)k * nu * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(dr.p, ratings, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(didx.p, ratings_index, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(dib.p, items_bias, (size_t)ni * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(dub.p, users_bias, (size_t)nu * 8, cudaMemcpyHostToDevice, st));
    funk_learned_bias_kernel<<<1, 1, 0, st>>>(min_epochs, min_improvement, k, f_init, learning_rate,
                                              learning_rate_users, learning_rate_items, K_feature, K_bias,
                                              overall_avg, du.p, dv.p, didx.p, dr.p, nnz, ni, nu, dib.p, dub.p,
                                              dfe.p, dfr.p);
    MF_LAUNCH_CHECK(ctx);
    std::vector<int32_t> h_fe(k);
    std::vector<double> h_fr(k);
    MF_CUDA(ctx, cudaMemcpyAsync(h_fe.data(), dfe.p, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaMemcpyAsync(h_fr.data(), dfr.p, (size_t)k * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaMemcpyAsync(u, du.p, (size_t)k * ni * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaMemcpyAsync(v, dv.p, (size_t)k * nu * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaMemcpyAsync(items_bias, dib.p, (size_t)ni * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaMemcpyAsync(users_bias, dub.p, (size_t)nu * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += 1;
    if (feature_epochs) memcpy(feature_epochs, h_fe.data(), (size_t)k * 4);
    if (feature_rmse) memcpy(feature_rmse, h_fr.data(), (size_t)k * 8);
    return MFREC_OK;
}

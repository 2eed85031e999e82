// Evaluation kernels: batched predict / RMSE over (user, item) pairs and the bias statistics
// that precede training.  All HBM-bound: one warp per pair, both factor rows read with
// coalesced 128-bit loads, shuffle-reduced dot product.
//
// Reference being replaced (all interpreted Python loops, one k-dot per call):
//   metrics.test_predict_rating            mfrec/recommendation/metrics.py:51-82
//   GDRecommender.predict_rating[_with_bias]  gradient_descent.py:621-648
//   KMFRecommender.predict_{logistic,linear,linear_neg}  kmf.py:79-103
//   compute_overall_avg base.py:504-508, compute_{items,users}_bias_bk mf.py:78-121
#include <cmath>

#include "common.cuh"

namespace {

struct PredParams {
    const float *Q, *ib, *P, *ub;
    const int32_t *user_perm, *item_perm;
    const int32_t *pairs;
    const void *real;
    int real_is_f32;
    int64_t n;
    int predictor;
    float mu, min_rating, max_rating;
    double *out;        // nullable
    double *part;       // [gridDim.x][3]  sum e^2, sum |e|, n_valid   (nullable)
    int32_t ni, nu;
    int32_t *bad;
};

template <int E>
__global__ void __launch_bounds__(256) predict_kernel(const PredParams p)
{
    constexpr int KPAD = E * 32;
    constexpr int V = E >= 4 ? 4 : E, NV = E / V;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    double s2 = 0.0, s1 = 0.0, cnt = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib; j < p.n; j += warps) {
        const int2 ui = reinterpret_cast<const int2 *>(p.pairs)[j];
        if (ui.x < 0 || ui.x >= p.nu || ui.y < 0 || ui.y >= p.ni) {
            if (lane == 0) atomicOr(p.bad, 1);
            continue;
        }
        const int64_t ur = p.user_perm ? p.user_perm[ui.x] : ui.x;
        const int64_t ir = p.item_perm ? p.item_perm[ui.y] : ui.y;
        const float *pu = p.P + ur * KPAD, *qi = p.Q + ir * KPAD;
        float part = 0.f;
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            const int off = (c * 32 + lane) * V;
            if constexpr (V == 4) {
                const float4 a = *reinterpret_cast<const float4 *>(pu + off);
                const float4 b = *reinterpret_cast<const float4 *>(qi + off);
                part = fmaf(a.x, b.x, part); part = fmaf(a.y, b.y, part);
                part = fmaf(a.z, b.z, part); part = fmaf(a.w, b.w, part);
            } else if constexpr (V == 2) {
                const float2 a = *reinterpret_cast<const float2 *>(pu + off);
                const float2 b = *reinterpret_cast<const float2 *>(qi + off);
                part = fmaf(a.x, b.x, part); part = fmaf(a.y, b.y, part);
            } else {
                part = fmaf(pu[off], qi[off], part);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) {
            const float dot = part;
            const float bsum = p.ib[ir] + p.ub[ur];
            float pred;
            switch (p.predictor) {
            case MFREC_PRED_GD_RATING: pred = dot + 1.0f; break;
            case MFREC_PRED_GD_RATING_BIAS: pred = dot + (p.mu + bsum); break;
            case MFREC_PRED_KMF_LINEAR: pred = dot + bsum; break;
            case MFREC_PRED_KMF_LOGISTIC:
                pred = p.min_rating + (1.f / (1.f + expf(-(dot + bsum)))) * (p.max_rating - p.min_rating);
                break;
            case MFREC_PRED_KMF_LINEAR_NEG:
                pred = p.min_rating + (dot + bsum) * (p.max_rating - p.min_rating);
                break;
            default: pred = dot; break;
            }
            if (p.out) p.out[j] = (double)pred;
            if (p.real) {
                const double real = p.real_is_f32 ? (double)((const float *)p.real)[j]
                                                  : ((const double *)p.real)[j];
                const double e = real - (double)pred;
                if (e == e) { s2 += e * e; s1 += fabs(e); cnt += 1.0; }
            }
        }
    }
    if (p.part) {
        __shared__ double sh[8][3];
        if (lane == 0) { sh[wib][0] = s2; sh[wib][1] = s1; sh[wib][2] = cnt; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0, b = 0, c = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sh[w][0]; b += sh[w][1]; c += sh[w][2]; }
            p.part[blockIdx.x * 3 + 0] = a;
            p.part[blockIdx.x * 3 + 1] = b;
            p.part[blockIdx.x * 3 + 2] = c;
        }
    }
}

__global__ void __launch_bounds__(256) stats_reduce_kernel(const double *__restrict__ part, int nblk,
                                                           double *__restrict__ out)
{
    __shared__ double sh[256][3];
    double a = 0, b = 0, c = 0;
    for (int i = threadIdx.x; i < nblk; i += 256) { a += part[i * 3]; b += part[i * 3 + 1]; c += part[i * 3 + 2]; }
    sh[threadIdx.x][0] = a; sh[threadIdx.x][1] = b; sh[threadIdx.x][2] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o)
            for (int q = 0; q < 3; ++q) sh[threadIdx.x][q] += sh[threadIdx.x + o][q];
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = sh[0][0]; out[1] = sh[0][1]; out[2] = sh[0][2]; out[3] = 0.0; }
}

// ---- bias statistics ----------------------------------------------------------------------
__global__ void sum_ratings_kernel(const double *__restrict__ r, int64_t nnz, double *__restrict__ part)
{
    __shared__ double sh[256];
    double acc = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nnz; n += stride) acc += r[n];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

__global__ void item_bias_acc_kernel(const int32_t *__restrict__ idx, const double *__restrict__ r,
                                     int64_t nnz, double mu, int32_t ni, int32_t nu,
                                     double *__restrict__ acc_i, int32_t *__restrict__ cnt_i,
                                     int32_t *__restrict__ bad)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nnz; n += stride) {
        const int2 ui = reinterpret_cast<const int2 *>(idx)[n];
        if (ui.x < 0 || ui.x >= nu || ui.y < 0 || ui.y >= ni) { atomicOr(bad, 1); continue; }
        atomicAdd(&acc_i[ui.y], r[n] - mu);
        atomicAdd(&cnt_i[ui.y], 1);
    }
}

__global__ void finish_bias_kernel(double *__restrict__ acc, const int32_t *__restrict__ cnt, int32_t n, double K)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] = cnt[i] > 0 ? acc[i] / (K + (double)cnt[i]) : 0.0;
}

__global__ void user_bias_acc_kernel(const int32_t *__restrict__ idx, const double *__restrict__ r,
                                     int64_t nnz, double mu, const double *__restrict__ bias_i,
                                     double *__restrict__ acc_u, int32_t *__restrict__ cnt_u)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nnz; n += stride) {
        const int2 ui = reinterpret_cast<const int2 *>(idx)[n];
        atomicAdd(&acc_u[ui.x], r[n] - mu - bias_i[ui.y]);
        atomicAdd(&cnt_u[ui.x], 1);
    }
}

}  // namespace

extern "C" int mfrec_model_predict(mfrec_ctx *ctx, const mfrec_model *m, int predictor,
                                   const int32_t *pairs, const void *real, int real_is_f32, int64_t n,
                                   int is_device, double mu, double min_rating, double max_rating,
                                   double *out, double stats_out[4])
{
    if (!ctx || !m) return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_model_predict: NULL argument");
    if (predictor < 0 || predictor > MFREC_PRED_DOT)
        return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_model_predict: predictor=%d", predictor);
    if (n < 0 || (n > 0 && !pairs)) return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_model_predict: n=%lld", (long long)n);
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (n == 0) {
        if (stats_out) stats_out[0] = stats_out[1] = stats_out[2] = stats_out[3] = 0.0;
        return MFREC_OK;
    }
    DevBuf<int32_t> d_pairs, d_bad;
    DevBuf<char> d_real;
    DevBuf<double> d_out, d_part, d_stats;
    const size_t rsz = real_is_f32 ? 4 : 8;
    PredParams p;
    p.pairs = pairs;
    p.real = real;
    p.out = out;
    if (!is_device) {
        MF_CUDA(ctx, d_pairs.alloc((size_t)n * 2, ctx->stream));
        MF_CUDA(ctx, cudaMemcpyAsync(d_pairs.p, pairs, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        p.pairs = d_pairs.p;
        if (real) {
            MF_CUDA(ctx, d_real.alloc((size_t)n * rsz, ctx->stream));
            MF_CUDA(ctx, cudaMemcpyAsync(d_real.p, real, (size_t)n * rsz, cudaMemcpyHostToDevice, st));
            p.real = d_real.p;
        }
        if (out) {
            MF_CUDA(ctx, d_out.alloc((size_t)n, ctx->stream));
            p.out = d_out.p;
        }
    }
    MF_CUDA(ctx, d_bad.alloc(1, ctx->stream));
    MF_CUDA(ctx, cudaMemsetAsync(d_bad.p, 0, 4, st));
    const int warps_per_block = 8;
    int grid = (int)std::min<int64_t>(ceil_div64(n, warps_per_block), (int64_t)ctx->sm_count * 16);
    if (stats_out) {
        MF_CUDA(ctx, d_part.alloc((size_t)grid * 3, ctx->stream));
        MF_CUDA(ctx, d_stats.alloc(4, ctx->stream));
    }
    p.Q = m->Q; p.ib = m->ib; p.P = m->P; p.ub = m->ub;
    p.user_perm = m->user_perm; p.item_perm = m->item_perm;
    p.real_is_f32 = real_is_f32;
    p.n = n;
    p.predictor = predictor;
    p.mu = (float)mu; p.min_rating = (float)min_rating; p.max_rating = (float)max_rating;
    p.part = stats_out ? d_part.p : nullptr;
    p.ni = m->ni; p.nu = m->nu;
    p.bad = d_bad.p;
    switch (m->kpad) {
    case 32: predict_kernel<1><<<grid, 256, 0, st>>>(p); break;
    case 64: predict_kernel<2><<<grid, 256, 0, st>>>(p); break;
    case 128: predict_kernel<4><<<grid, 256, 0, st>>>(p); break;
    case 256: predict_kernel<8><<<grid, 256, 0, st>>>(p); break;
    default: return mfrec_set_error(ctx, MFREC_ERR_UNSUPPORTED, "kpad=%d", m->kpad);
    }
    MF_LAUNCH_CHECK(ctx);
    if (stats_out) {
        stats_reduce_kernel<<<1, 256, 0, st>>>(d_part.p, grid, d_stats.p);
        MF_LAUNCH_CHECK(ctx);
        MF_CUDA(ctx, cudaMemcpyAsync(stats_out, d_stats.p, 32, cudaMemcpyDeviceToHost, st));
    }
    if (!is_device && out)
        MF_CUDA(ctx, cudaMemcpyAsync(out, d_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    int32_t h_bad = 0;
    MF_CUDA(ctx, cudaMemcpyAsync(&h_bad, d_bad.p, 4, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaStreamSynchronize(st));
    if (h_bad)
        return mfrec_set_error(ctx, MFREC_ERR_INDEX, "mfrec_model_predict: a pair has a user outside [0,%d) or an item outside [0,%d)",
                               m->nu, m->ni);
    return MFREC_OK;
}

extern "C" int mfrec_predict_pairs(mfrec_ctx *ctx, int predictor, int k, const double *u, const double *v,
                                   int32_t ni, int32_t nu, const int32_t *pairs, int64_t n, double mu,
                                   const double *items_bias, const double *users_bias,
                                   double min_rating, double max_rating, double *out)
{
    if (!ctx || !u || !v || !out) return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_predict_pairs: NULL argument");
    mfrec_model *M = nullptr;
    MF_TRY(mfrec_model_create(ctx, nullptr, k, ni, nu, u, v, items_bias, users_bias, &M));
    const int rc = mfrec_model_predict(ctx, M, predictor, pairs, nullptr, 0, n, 0, mu, min_rating, max_rating, out, nullptr);
    mfrec_model_destroy(M);
    return rc;
}

extern "C" int mfrec_rmse_pairs(mfrec_ctx *ctx, int predictor, int k, const double *u, const double *v,
                                int32_t ni, int32_t nu, const int32_t *pairs, const double *real,
                                int64_t n, double mu, const double *items_bias, const double *users_bias,
                                double min_rating, double max_rating, double *errors_out, double stats[4])
{
    if (!ctx || !u || !v || !stats || (n > 0 && !real))
        return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_rmse_pairs: NULL argument");
    mfrec_model *M = nullptr;
    MF_TRY(mfrec_model_create(ctx, nullptr, k, ni, nu, u, v, items_bias, users_bias, &M));
    double raw[4] = {0, 0, 0, 0};
    const int rc = mfrec_model_predict(ctx, M, predictor, pairs, real, 0, n, 0, mu, min_rating, max_rating,
                                       errors_out, raw);
    mfrec_model_destroy(M);
    if (rc != MFREC_OK) return rc;
    if (errors_out)
        for (int64_t j = 0; j < n; ++j) errors_out[j] = real[j] - errors_out[j];
    const double cnt = raw[2];
    if (cnt > 0) {
        const double mae = raw[1] / cnt, ms = raw[0] / cnt;
        stats[0] = sqrt(ms);
        stats[1] = mae;
        stats[2] = fmax(ms - mae * mae, 0.0);
        stats[3] = cnt;
    } else {
        stats[0] = stats[1] = stats[2] = NAN;
        stats[3] = 0.0;
    }
    return MFREC_OK;
}

extern "C" int mfrec_bias_stats(mfrec_ctx *ctx, const int32_t *ratings_index, const double *ratings,
                                int64_t nnz, int32_t ni, int32_t nu, double K2, double K3,
                                double *mu_out, double *items_bias, double *users_bias)
{
    if (!ctx || !mu_out || !items_bias || !users_bias || nnz < 0 || ni <= 0 || nu <= 0 ||
        (nnz > 0 && (!ratings_index || !ratings)))
        return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_bias_stats: bad argument");
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (nnz == 0) {
        *mu_out = NAN;
        for (int32_t i = 0; i < ni; ++i) items_bias[i] = 0.0;
        for (int32_t j = 0; j < nu; ++j) users_bias[j] = 0.0;
        return MFREC_OK;
    }
    DevBuf<int32_t> d_idx, cnt_i, cnt_u, bad;
    DevBuf<double> d_r, acc_i, acc_u, part;
    const int grid = ctx->sm_count * 8;
    MF_CUDA(ctx, d_idx.alloc((size_t)nnz * 2, ctx->stream));
    MF_CUDA(ctx, d_r.alloc((size_t)nnz, ctx->stream));
    MF_CUDA(ctx, cnt_i.alloc(ni)); MF_CUDA(ctx, cnt_u.alloc(nu, ctx->stream));
    MF_CUDA(ctx, acc_i.alloc(ni)); MF_CUDA(ctx, acc_u.alloc(nu, ctx->stream));
    MF_CUDA(ctx, part.alloc(grid)); MF_CUDA(ctx, bad.alloc(1, ctx->stream));
    MF_CUDA(ctx, cudaMemcpyAsync(d_idx.p, ratings_index, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(d_r.p, ratings, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemsetAsync(cnt_i.p, 0, (size_t)ni * 4, st));
    MF_CUDA(ctx, cudaMemsetAsync(cnt_u.p, 0, (size_t)nu * 4, st));
    MF_CUDA(ctx, cudaMemsetAsync(acc_i.p, 0, (size_t)ni * 8, st));
    MF_CUDA(ctx, cudaMemsetAsync(acc_u.p, 0, (size_t)nu * 8, st));
    MF_CUDA(ctx, cudaMemsetAsync(bad.p, 0, 4, st));
    sum_ratings_kernel<<<grid, 256, 0, st>>>(d_r.p, nnz, part.p);
    MF_LAUNCH_CHECK(ctx);
    std::vector<double> h_part(grid);
    MF_CUDA(ctx, cudaMemcpyAsync(h_part.data(), part.p, (size_t)grid * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaStreamSynchronize(st));
    double tot = 0.0;
    for (double x : h_part) tot += x;
    const double mu = tot / (double)nnz;
    item_bias_acc_kernel<<<grid, 256, 0, st>>>(d_idx.p, d_r.p, nnz, mu, ni, nu, acc_i.p, cnt_i.p, bad.p);
    MF_LAUNCH_CHECK(ctx);
    int32_t h_bad = 0;
    MF_CUDA(ctx, cudaMemcpyAsync(&h_bad, bad.p, 4, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaStreamSynchronize(st));
    if (h_bad) return mfrec_set_error(ctx, MFREC_ERR_INDEX, "mfrec_bias_stats: index out of range");
    finish_bias_kernel<<<(ni + 255) / 256, 256, 0, st>>>(acc_i.p, cnt_i.p, ni, K3);
    MF_LAUNCH_CHECK(ctx);
    user_bias_acc_kernel<<<grid, 256, 0, st>>>(d_idx.p, d_r.p, nnz, mu, acc_i.p, acc_u.p, cnt_u.p);
    MF_LAUNCH_CHECK(ctx);
    finish_bias_kernel<<<(nu + 255) / 256, 256, 0, st>>>(acc_u.p, cnt_u.p, nu, K2);
    MF_LAUNCH_CHECK(ctx);
    MF_CUDA(ctx, cudaMemcpyAsync(items_bias, acc_i.p, (size_t)ni * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaMemcpyAsync(users_bias, acc_u.p, (size_t)nu * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaStreamSynchronize(st));
    *mu_out = mu;
    return MFREC_OK;
}

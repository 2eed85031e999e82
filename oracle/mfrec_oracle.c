/*
 * mfrec_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded, float64 restatement of the reference's latent-factor
 * hot path, used as the parity oracle for the CUDA implementation.  Nothing in the
 * product (mfrec_b200/, include/) may link, import or call this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors
 * (SURVEY.md section 4), so this restatement is pinned by executing the reference's own
 * kernels -- built unmodified from mfrec/lib/{kmf_train,gd_estimator}.pyx by
 * oracle/build_ref.sh into oracle/_ref/ -- on seeded inputs; tests/test_oracle.py
 * demands bit-equality (linear / Funk) or <=2 ulp (logistic: libm exp) against it, and
 * tests/golden/ holds committed outputs of oracle/_ref made by tests/golden/make_golden.py.
 *
 * Layout follows the reference exactly: factors are feature-major [k][n] float64,
 * `u` = ITEM factors, `v` = USER factors, ratings_index is [nnz][2] = (user, item).
 *
 * Compile with -O2 -ffp-contract=off (the reference build has no FMA contraction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- mfrec/lib/gd_estimator.pyx:26-35 (same body kmf_train.pyx:22-31): the clamp
 *      ignores its min/max arguments, the range [1,5] is hard-coded. */
static double clamp15(double x)
{
    if (x > 5.0) x = 5.0;
    if (x < 1.0) x = 1.0;
    return x;
}

/* ---- mfrec/lib/gd_estimator.pyx:38-73 `estimator`: cached partial prediction + one
 *      feature product, clamp, optional trailing term for the untrained features. */
static double funk_estimate(double uf, double vf, int f, int dim, double f_init,
                            double cache, int trailing, double overall_avg,
                            double item_bias, double user_bias)
{
    double s;
    if (cache > 0) s = cache;
    else           s = overall_avg + item_bias + user_bias;
    s += uf * vf;
    s = clamp15(s);
    if (trailing == 1) {
        s += (dim - f - 1) * f_init * f_init;
        s = clamp15(s);
    }
    return s;
}

/* ---- mfrec/lib/kmf_train.pyx:72-97 `full_estimator`: overall + b_i + b_u + sum_f u*v,
 *      sequential in f, no clamp. */
static double kmf_full_estimate(const double *u, const double *v, int64_t ni, int64_t nu,
                                int dim, int item, int user, double overall_avg,
                                double item_bias, double user_bias)
{
    double s = overall_avg + item_bias + user_bias;
    for (int f = 0; f < dim; ++f)
        s += u[(int64_t)f * ni + item] * v[(int64_t)f * nu + user];
    return s;
}

/*
 * mfrec/lib/kmf_train.pyx:195-277 (kernel==0, train_linear_kernel) and
 * kmf_train.pyx:103-189 (kernel==1, train_logistic_kernel).
 *
 * Quirks kept on purpose:
 *  - overall_avg is passed to full_estimator as the literal 0.0 (:159, :250);
 *  - the linear kernel updates both biases unconditionally (:259-260), the logistic one
 *    gates them on update_users / update_items (:168-171);
 *  - only `learning_rate` is used; learning_rate_users/items and f_init are dead;
 *  - features are updated from the OLD cf / mf pair (:263-270).
 * rmse_out (nullable) receives sqrt(se/nnz) per epoch (the reference only prints it).
 */
void oracle_kmf_train(int kernel, int nbr_epochs, int dim, double lr, double K_users,
                      double K_items, double K_bias, double *u, double *v,
                      const int32_t *ratings_index, const double *ratings, int64_t nnz,
                      int64_t ni, int64_t nu, double *items_bias, double *users_bias,
                      int update_users, int update_items, double *rmse_out)
{
    const double rating_range_size = 4.0, min_rating = 1.0;
    for (int epoch = 0; epoch < nbr_epochs; ++epoch) {
        double se = 0.0;
        for (int64_t n = 0; n < nnz; ++n) {
            const int user = ratings_index[2 * n], item = ratings_index[2 * n + 1];
            const double rating = ratings[n];
            const double dot = kmf_full_estimate(u, v, ni, nu, dim, item, user, 0.0,
                                                 items_bias[item], users_bias[user]);
            double err, grad;
            if (kernel == 0) {
                err = rating - dot;
                se += err * err;
                grad = err;
                users_bias[user] += lr * (grad - K_bias * users_bias[user]);
                items_bias[item] += lr * (grad - K_bias * items_bias[item]);
            } else {
                const double sig = 1.0 / (1.0 + exp(-dot));
                const double p = min_rating + sig * rating_range_size;
                err = rating - p;
                se += err * err;
                grad = err * sig * (1.0 - sig) * rating_range_size;
                if (update_users)
                    users_bias[user] += lr * (grad - K_bias * users_bias[user]);
                if (update_items)
                    items_bias[item] += lr * (grad - K_bias * items_bias[item]);
            }
            for (int f = 0; f < dim; ++f) {
                double *pu = &u[(int64_t)f * ni + item], *pv = &v[(int64_t)f * nu + user];
                const double cf = *pv, mf = *pu;
                if (update_items) *pu += lr * (grad * cf - K_items * mf);
                if (update_users) *pv += lr * (grad * mf - K_users * cf);
            }
        }
        if (rmse_out) rmse_out[epoch] = sqrt(se / (double)nnz);
    }
}

/*
 * Funk-SVD per-feature loops:
 *   variant 0: estimator_loop_without_bias   mfrec/lib/gd_estimator.pyx:691-779
 *   variant 1: estimator_loop_with_bias      gd_estimator.pyx:489-582
 *   variant 2: estimator_loop_with_bias_dev  gd_estimator.pyx:588-685 (update gates)
 *
 * Quirks kept: rmse (2.0) and rmse_last (0.0) carry across features; max_epochs is
 * ignored; the per-rating cache is refreshed once per feature with trailing=0; biases
 * are read-only; variant 0 uses the estimator's defaults overall_avg=1.0, biases 0.
 * Returns the number of training passes done (sum over features); feature_epochs
 * (nullable, [dim]) gets the per-feature pass count, feature_rmse (nullable, [dim])
 * the rmse of the last pass of each feature.
 */
int64_t oracle_funk_train(int variant, int min_epochs, double min_improvement, int dim,
                          double f_init, double lr, double K, double overall_avg,
                          double *u, double *v, const int32_t *ratings_index,
                          const double *ratings, int64_t nnz, int64_t ni, int64_t nu,
                          const double *items_bias, const double *users_bias,
                          int update_users, int update_items, int32_t *feature_epochs,
                          double *feature_rmse)
{
    double rmse = 2.0, rmse_last = 0.0;
    int64_t passes = 0;
    double *cache = (double *)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(double));
    for (int64_t n = 0; n < nnz; ++n) cache[n] = 0.0;
    if (variant != 2) { update_users = 1; update_items = 1; }
    for (int f = 0; f < dim; ++f) {
        double *uf = u + (int64_t)f * ni, *vf = v + (int64_t)f * nu;
        int epoch = 0;
        while (epoch < min_epochs || rmse <= rmse_last - min_improvement) {
            double se = 0.0;
            rmse_last = rmse;
            for (int64_t n = 0; n < nnz; ++n) {
                const int user = ratings_index[2 * n], item = ratings_index[2 * n + 1];
                const double rating = ratings[n];
                double p;
                if (variant == 0)
                    p = funk_estimate(uf[item], vf[user], f, dim, f_init, cache[n], 1,
                                      1.0, 0.0, 0.0);
                else
                    p = funk_estimate(uf[item], vf[user], f, dim, f_init, cache[n], 1,
                                      overall_avg, items_bias[item], users_bias[user]);
                const double err = rating - p;
                se += err * err;
                const double cf = vf[user], mf = uf[item];
                if (update_items) uf[item] += lr * (err * cf - K * mf);
                if (update_users) vf[user] += lr * (err * mf - K * cf);
            }
            rmse = sqrt(se / (double)nnz);
            ++epoch;
            ++passes;
        }
        if (feature_epochs) feature_epochs[f] = epoch;
        if (feature_rmse) feature_rmse[f] = rmse;
        for (int64_t n = 0; n < nnz; ++n) {
            const int user = ratings_index[2 * n], item = ratings_index[2 * n + 1];
            if (variant == 0)
                cache[n] = funk_estimate(uf[item], vf[user], f, dim, f_init, cache[n], 0,
                                         1.0, 0.0, 0.0);
            else
                cache[n] = funk_estimate(uf[item], vf[user], f, dim, f_init, cache[n], 0,
                                         overall_avg, items_bias[item], users_bias[user]);
        }
    }
    free(cache);
    return passes;
}

/*
 * "Fair layout" CPU figure: the arithmetic of train_linear_kernel
 * (kmf_train.pyx:241-273, both sides updated) on ROW-major float32 factors -- the layout the GPU
 * path uses -- so that bench.py can report how much of the GPU / reference ratio is the
 * reference's feature-major float64 layout and how much is the device.  One thread, the input
 * order; a measurement aid, not a parity oracle.  Returns the sum of squared errors.
 */
double oracle_kmf_epoch_rowmajor_f32(int dim, float lr, float K_users, float K_items, float K_bias,
                                     float *P /* [nu][dim] */, float *Q /* [ni][dim] */,
                                     float *users_bias, float *items_bias,
                                     const int32_t *ratings_index, const float *ratings, int64_t nnz)
{
    double se = 0.0;
    for (int64_t n = 0; n < nnz; ++n) {
        const int user = ratings_index[2 * n], item = ratings_index[2 * n + 1];
        float *p = P + (int64_t)user * dim, *q = Q + (int64_t)item * dim;
        float s = items_bias[item] + users_bias[user];
        for (int f = 0; f < dim; ++f) s += p[f] * q[f];
        const float err = ratings[n] - s;
        se += (double)err * err;
        users_bias[user] += lr * (err - K_bias * users_bias[user]);
        items_bias[item] += lr * (err - K_bias * items_bias[item]);
        for (int f = 0; f < dim; ++f) {
            const float cf = p[f], mf = q[f];
            q[f] = mf + lr * (err * cf - K_items * mf);
            p[f] = cf + lr * (err * mf - K_users * cf);
        }
    }
    return se;
}

/* The same loop on T host threads, thread t taking the ratings of the users in
 * [t * nu / T, (t + 1) * nu / T) in array order: user rows are private to a thread, item rows and
 * item biases are updated without locks (races between threads: a THROUGHPUT figure for bench.py's
 * fair-layout CPU line, not a result anything is compared with). */
#include <pthread.h>
typedef struct {
    int dim; float lr, Ku, Ki, Kb; float *P, *Q, *ub, *ib; const int32_t *idx; const float *r; int64_t nnz;
    int32_t u_lo, u_hi; double se;
} rowmajor_job;

static void *rowmajor_worker(void *arg)
{
    rowmajor_job *j = (rowmajor_job *)arg;
    const int dim = j->dim;
    double se = 0.0;
    for (int64_t n = 0; n < j->nnz; ++n) {
        const int user = j->idx[2 * n];
        if (user < j->u_lo || user >= j->u_hi) continue;
        const int item = j->idx[2 * n + 1];
        float *p = j->P + (int64_t)user * dim, *q = j->Q + (int64_t)item * dim;
        float s = j->ib[item] + j->ub[user];
        for (int f = 0; f < dim; ++f) s += p[f] * q[f];
        const float err = j->r[n] - s;
        se += (double)err * err;
        j->ub[user] += j->lr * (err - j->Kb * j->ub[user]);
        j->ib[item] += j->lr * (err - j->Kb * j->ib[item]);
        for (int f = 0; f < dim; ++f) {
            const float cf = p[f], mf = q[f];
            q[f] = mf + j->lr * (err * cf - j->Ki * mf);
            p[f] = cf + j->lr * (err * mf - j->Ku * cf);
        }
    }
    j->se = se;
    return 0;
}

double oracle_kmf_epoch_rowmajor_f32_mt(int dim, float lr, float K_users, float K_items, float K_bias,
                                        float *P, float *Q, float *users_bias, float *items_bias,
                                        const int32_t *ratings_index, const float *ratings, int64_t nnz,
                                        int32_t nu, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    rowmajor_job jobs[64];
    pthread_t tid[64];
    for (int t = 0; t < threads; ++t) {
        rowmajor_job j = {dim, lr, K_users, K_items, K_bias, P, Q, users_bias, items_bias, ratings_index, ratings, nnz,
                          (int32_t)((int64_t)nu * t / threads), (int32_t)((int64_t)nu * (t + 1) / threads), 0.0};
        jobs[t] = j;
        pthread_create(&tid[t], 0, rowmajor_worker, &jobs[t]);
    }
    double se = 0.0;
    for (int t = 0; t < threads; ++t) {
        pthread_join(tid[t], 0);
        se += jobs[t].se;
    }
    return se;
}

/*
 * A3 development variants of the Funk loop; all share the training pass
 * of estimator_loop_without_bias and the `estimator` above.  Their rating cache is a DENSE
 * array indexed `user + item * nbr_users` (gd_estimator.pyx:250-255): toy sizes only.
 *
 * ---- mfrec/lib/gd_estimator.pyx:210-303 `estimator_loop` (feature_training_dev,
 *      gradient_descent.py:596): the only loop that honours max_epochs (:264) and records
 *      rmse_hist[epoch + f*max_epochs + batch*max_epochs*dim] (:285); `improvement` (0.0) carries
 *      across features like rmse / rmse_last; user factor written before the item factor
 *      (:281-282), both from the old values.
 * ---- :308-395 `estimator_loop2`: same with the control of estimator_loop_without_bias (no
 *      history, no max_epochs) -- pass max_epochs < 0 and rmse_hist = NULL.
 */
void oracle_funk_loop_dev(int min_epochs, int max_epochs, double min_improvement, int dim,
                          double f_init, double lr, double K, double *u, double *v,
                          const int32_t *ratings_index, const double *ratings, int64_t nnz,
                          int64_t ni, int64_t nu, int batch, double *rmse_hist,
                          int32_t *feature_epochs, double *feature_rmse)
{
    double rmse = 2.0, rmse_last = 0.0, improvement = 0.0;
    const int hist = max_epochs >= 0;
    double *cache = (double *)calloc((size_t)(ni * nu > 0 ? ni * nu : 1), sizeof(double));
    for (int f = 0; f < dim; ++f) {
        double *uf = u + (int64_t)f * ni, *vf = v + (int64_t)f * nu;
        int epoch = 0;
        while (hist ? ((epoch < min_epochs || improvement >= min_improvement) && epoch < max_epochs)
                    : (epoch < min_epochs || rmse <= rmse_last - min_improvement)) {
            double se = 0.0;
            rmse_last = rmse;
            for (int64_t n = 0; n < nnz; ++n) {
                const int user = ratings_index[2 * n], item = ratings_index[2 * n + 1];
                const double p = funk_estimate(uf[item], vf[user], f, dim, f_init,
                                               cache[user + (int64_t)item * nu], 1, 1.0, 0.0, 0.0);
                const double err = 1.0 * ratings[n] - p;
                se += err * err;
                const double cf = vf[user], mf = uf[item];
                vf[user] += lr * (err * mf - K * cf);
                uf[item] += lr * (err * cf - K * mf);
            }
            rmse = sqrt(se / (double)nnz);
            if (hist) {
                rmse_hist[epoch + (int64_t)f * max_epochs + (int64_t)batch * max_epochs * dim] = rmse;
                improvement = rmse_last - rmse;
            }
            ++epoch;
        }
        if (feature_epochs) feature_epochs[f] = epoch;
        if (feature_rmse) feature_rmse[f] = rmse;
        for (int64_t n = 0; n < nnz; ++n) {
            const int user = ratings_index[2 * n], item = ratings_index[2 * n + 1];
            double *c = &cache[user + (int64_t)item * nu];
            *c = funk_estimate(uf[item], vf[user], f, dim, f_init, *c, 0, 1.0, 0.0, 0.0);
        }
    }
    free(cache);
}

/* ---- mfrec/lib/gd_estimator.pyx:903-962 `estimator_subloop` (feature_training2,
 *      gradient_descent.py:322): exactly one pass of feature f with the caller's dense cache
 *      (read only); returns the rmse -- the only native function with a return value. */
double oracle_funk_subloop(int f, int dim, double f_init, double lr, double K, double *u, double *v,
                           const int32_t *ratings_index, const double *ratings, int64_t nnz,
                           int64_t ni, int64_t nu, const double *rating_cache)
{
    double *uf = u + (int64_t)f * ni, *vf = v + (int64_t)f * nu;
    double se = 0.0;
    for (int64_t n = 0; n < nnz; ++n) {
        const int user = ratings_index[2 * n], item = ratings_index[2 * n + 1];
        const double p = funk_estimate(uf[item], vf[user], f, dim, f_init,
                                       rating_cache[user + (int64_t)item * nu], 1, 1.0, 0.0, 0.0);
        const double err = 1.0 * ratings[n] - p;
        se += err * err;
        const double cf = vf[user], mf = uf[item];
        vf[user] += lr * (err * mf - K * cf);
        uf[item] += lr * (err * cf - K * mf);
    }
    return sqrt(se / (double)nnz);
}

/* ---- mfrec/lib/gd_estimator.pyx:967-995 `predictor_subloop` (gradient_descent.py:327): the
 *      cache refresh of feature f on the caller's dense cache (written). */
void oracle_funk_predictor_subloop(int f, int dim, double f_init, const double *u, const double *v,
                                   const int32_t *ratings_index, int64_t nnz, int64_t ni, int64_t nu,
                                   double *rating_cache)
{
    const double *uf = u + (int64_t)f * ni, *vf = v + (int64_t)f * nu;
    for (int64_t n = 0; n < nnz; ++n) {
        const int user = ratings_index[2 * n], item = ratings_index[2 * n + 1];
        double *c = &rating_cache[user + (int64_t)item * nu];
        *c = funk_estimate(uf[item], vf[user], f, dim, f_init, *c, 0, 1.0, 0.0, 0.0);
    }
}

/* ---- mfrec/lib/gd_estimator.pyx:401-483 `estimator_loop_with_learned_bias`
 *      (feature_training_bias, gradient_descent.py:501) with its `full_estimator` :115-148:
 *      a full clamped k-dot per rating (overall + b_i + b_u + sum_f, clamp, trailing term, clamp;
 *      no cache), but only feature f and the two biases are updated -- biases first, with their
 *      own learning rates and the freshly updated value inside the regulariser (`+=` re-reads the
 *      element), then the item factor, then the user factor, both from the old values. */
void oracle_funk_learned_bias(int min_epochs, double min_improvement, int dim, double f_init,
                              double lr, double lr_users, double lr_items, double K_feature,
                              double K_bias, double overall_avg, double *u, double *v,
                              const int32_t *ratings_index, const double *ratings, int64_t nnz,
                              int64_t ni, int64_t nu, double *items_bias, double *users_bias,
                              int32_t *feature_epochs, double *feature_rmse)
{
    double rmse = 2.0, rmse_last = 0.0;
    for (int f = 0; f < dim; ++f) {
        double *uf = u + (int64_t)f * ni, *vf = v + (int64_t)f * nu;
        int epoch = 0;
        while (epoch < min_epochs || rmse <= rmse_last - min_improvement) {
            double se = 0.0;
            rmse_last = rmse;
            for (int64_t n = 0; n < nnz; ++n) {
                const int user = ratings_index[2 * n], item = ratings_index[2 * n + 1];
                double s = overall_avg + items_bias[item] + users_bias[user];
                for (int g = 0; g < dim; ++g) s += u[(int64_t)g * ni + item] * v[(int64_t)g * nu + user];
                s = clamp15(s);
                s += (dim - f - 1) * f_init * f_init;
                s = clamp15(s);
                const double err = ratings[n] - s;
                se += err * err;
                const double cf = vf[user], mf = uf[item];
                users_bias[user] += lr_users * (err - K_bias * users_bias[user]);
                items_bias[item] += lr_items * (err - K_bias * items_bias[item]);
                uf[item] += lr * (err * cf - K_feature * mf);
                vf[user] += lr * (err * mf - K_feature * cf);
            }
            rmse = sqrt(se / (double)nnz);
            ++epoch;
        }
        if (feature_epochs) feature_epochs[f] = epoch;
        if (feature_rmse) feature_rmse[f] = rmse;
    }
}

/*
 * Predictors (one k-dot + affine / logistic map), by id:
 *   0  GDRecommender.predict_rating            gradient_descent.py:621-631   dot + 1.0
 *   1  GDRecommender.predict_rating_with_bias  gradient_descent.py:637-648   dot + (mu + (b_i + b_u))
 *   2  KMFRecommender.predict_linear           kmf.py:88-94                  dot + (b_i + b_u)
 *   3  KMFRecommender.predict_logistic         kmf.py:79-85                  1 + 4*sigmoid(dot + (b_i+b_u))
 *   4  KMFRecommender.predict_linear_neg       kmf.py:97-103                 1 + 4*(dot + (b_i+b_u))
 *   5  WRMFRecommender.predict                 wrmf.py:67-69                 dot
 * min_rating / max_rating are the BaseRecommender attributes (base.py:92-93), 1 and 5.
 */
static double predict_one(int predictor, const double *u, const double *v, int64_t ni,
                          int64_t nu, int dim, int item, int user, double mu,
                          const double *ib, const double *ub, double min_rating,
                          double max_rating)
{
    double dot = 0.0;
    for (int f = 0; f < dim; ++f)
        dot += u[(int64_t)f * ni + item] * v[(int64_t)f * nu + user];
    switch (predictor) {
    case 0: return dot + 1.0;
    case 1: return dot + (mu + (ib[item] + ub[user]));
    case 2: return dot + (ib[item] + ub[user]);
    case 3: {
        const double s = dot + (ib[item] + ub[user]);
        return min_rating + (1.0 / (1.0 + exp(-s))) * (max_rating - min_rating);
    }
    case 4: return min_rating + (dot + (ib[item] + ub[user])) * (max_rating - min_rating);
    default: return dot;
    }
}

/* pairs are (user, item) rows like u_test / ratings_index. */
void oracle_predict_pairs(int predictor, const double *u, const double *v, int64_t ni,
                          int64_t nu, int dim, const int32_t *pairs, int64_t n, double mu,
                          const double *ib, const double *ub, double min_rating,
                          double max_rating, double *out)
{
    for (int64_t j = 0; j < n; ++j)
        out[j] = predict_one(predictor, u, v, ni, nu, dim, pairs[2 * j + 1], pairs[2 * j],
                             mu, ib, ub, min_rating, max_rating);
}

/*
 * metrics.test_predict_rating  mfrec/recommendation/metrics.py:51-82:
 * errors = real - predicted over the pairs, NaN errors dropped, then
 * out = { rmse = sqrt(mean(|e|^2)), mae = mean|e|, var(|e|) (population), n_valid }.
 * errors_out (nullable, [n]) receives every error in order (NaN kept in place).
 */
void oracle_rmse_pairs(int predictor, const double *u, const double *v, int64_t ni,
                       int64_t nu, int dim, const int32_t *pairs, const double *real,
                       int64_t n, double mu, const double *ib, const double *ub,
                       double min_rating, double max_rating, double *errors_out,
                       double out[4])
{
    double s2 = 0.0, s1 = 0.0;
    int64_t cnt = 0;
    double *tmp = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    for (int64_t j = 0; j < n; ++j) {
        const double p = predict_one(predictor, u, v, ni, nu, dim, pairs[2 * j + 1],
                                     pairs[2 * j], mu, ib, ub, min_rating, max_rating);
        const double e = real[j] - p;
        tmp[j] = e;
        if (errors_out) errors_out[j] = e;
        if (e == e) { s2 += e * e; s1 += fabs(e); ++cnt; }
    }
    const double mae = cnt ? s1 / (double)cnt : NAN;
    double var = 0.0;
    for (int64_t j = 0; j < n; ++j)
        if (tmp[j] == tmp[j]) { const double d = fabs(tmp[j]) - mae; var += d * d; }
    out[0] = cnt ? sqrt(s2 / (double)cnt) : NAN;
    out[1] = mae;
    out[2] = cnt ? var / (double)cnt : NAN;
    out[3] = (double)cnt;
    free(tmp);
}

/*
 * Top-N for one user:
 *   GDRecommender.find_user_top_match  gradient_descent.py:769-802 (all items, predictor 0)
 *   MFRecommender.find_recommended_items  mf.py:144-193 (the first `n_candidates` item ids
 *     -- the loop scores the enumeration index, not the sampled id -- any predictor).
 * Score every candidate item i unless i is already rated by the user or i == user_index
 * (quirk: the user id is appended to the rated-item list, mf.py:162 / gradient_descent.py:779);
 * NaN -> 0 (mf.py:176); drop exact zeros; stable sort by score descending (ties keep
 * ascending item id, Python's sorted(reverse=True) is stable); keep the first N.
 * rated_items: the user's rated item ids (any order), n_rated of them.
 * Returns how many results were written (<= N).
 */
typedef struct { double score; int32_t item; } scored_t;

static int scored_cmp(const void *a, const void *b)
{
    const scored_t *x = (const scored_t *)a, *y = (const scored_t *)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return (x->item > y->item) - (x->item < y->item);
}

int oracle_topn_user(int predictor, const double *u, const double *v, int64_t ni, int64_t nu,
                     int dim, int user, int n_candidates, const int32_t *rated_items,
                     int64_t n_rated, double mu, const double *ib, const double *ub,
                     double min_rating, double max_rating, int N, int32_t *out_items,
                     double *out_scores)
{
    uint8_t *mask = (uint8_t *)calloc((size_t)(n_candidates > 0 ? n_candidates : 1), 1);
    scored_t *sc = (scored_t *)malloc((size_t)(n_candidates > 0 ? n_candidates : 1) * sizeof(scored_t));
    for (int64_t j = 0; j < n_rated; ++j)
        if (rated_items[j] >= 0 && rated_items[j] < n_candidates) mask[rated_items[j]] = 1;
    if (user >= 0 && user < n_candidates) mask[user] = 1;
    int m = 0;
    for (int i = 0; i < n_candidates; ++i) {
        if (mask[i]) continue;
        double s = predict_one(predictor, u, v, ni, nu, dim, i, user, mu, ib, ub,
                               min_rating, max_rating);
        if (s != s) s = 0.0;
        if (s == 0.0) continue;
        sc[m].score = s; sc[m].item = i; ++m;
    }
    qsort(sc, (size_t)m, sizeof(scored_t), scored_cmp);
    const int outn = m < N ? m : N;
    for (int j = 0; j < outn; ++j) { out_items[j] = sc[j].item; out_scores[j] = sc[j].score; }
    free(mask); free(sc);
    return outn;
}

/*
 * Bias statistics (the step just before the hot path):
 *   compute_overall_avg    base.py:504-508   mu = mean of stored ratings
 *   compute_items_bias_bk  mf.py:78-97       b_i = sum_u (r - mu) / (K3 + n_i)
 *   compute_users_bias_bk  mf.py:100-121     b_u = sum_i (r - mu - b_i) / (K2 + n_u)
 * Inputs are the COO triples; empty rows/cols keep bias 0.  Sums run in ascending
 * (user, item) order for items' columns / users' rows as scipy's csc / csr slicing does;
 * callers pass triples sorted by (user, item) (lil_matrix.tocoo order).
 */
double oracle_bias_stats(const int32_t *ratings_index, const double *ratings, int64_t nnz,
                         int64_t ni, int64_t nu, double K2, double K3, double *items_bias,
                         double *users_bias)
{
    double tot = 0.0;
    for (int64_t n = 0; n < nnz; ++n) tot += ratings[n];
    const double mu = nnz ? tot / (double)nnz : NAN;
    double *cnt_i = (double *)calloc((size_t)(ni > 0 ? ni : 1), sizeof(double));
    double *cnt_u = (double *)calloc((size_t)(nu > 0 ? nu : 1), sizeof(double));
    for (int64_t i = 0; i < ni; ++i) items_bias[i] = 0.0;
    for (int64_t j = 0; j < nu; ++j) users_bias[j] = 0.0;
    for (int64_t n = 0; n < nnz; ++n) {
        const int item = ratings_index[2 * n + 1];
        items_bias[item] += ratings[n] - mu;
        cnt_i[item] += 1.0;
    }
    for (int64_t i = 0; i < ni; ++i)
        if (cnt_i[i] > 0) items_bias[i] = items_bias[i] / (K3 + cnt_i[i]);
    for (int64_t n = 0; n < nnz; ++n) {
        const int user = ratings_index[2 * n], item = ratings_index[2 * n + 1];
        users_bias[user] += ratings[n] - mu - items_bias[item];
        cnt_u[user] += 1.0;
    }
    for (int64_t j = 0; j < nu; ++j)
        if (cnt_u[j] > 0) users_bias[j] = users_bias[j] / (K2 + cnt_u[j]);
    free(cnt_i); free(cnt_u);
    return mu;
}

/*
 * ALS for implicit-feedback WRMF: mfrec/lib/als_implicit.pyx:208-352 (als_wrmf), restated.
 * u = item factors [dim][ni], v = user factors [dim][nu]; *_row = [0, count_0, count_1, ...]
 * (lib/datasets.py:13-32), *_col = neighbour ids.  The reference inverts the dim x dim system
 * matrix with numpy.linalg.inv (LAPACK) and multiplies; here the system is solved by Gauss-Jordan
 * elimination with partial pivoting -- equal to float64 round-off for these SPD systems; the
 * pin against oracle/_ref/als_implicit is at 1e-9 (tests/test_oracle.py).
 */
static void solve_dense(double *m, double *b, int n)
{
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r)
            if (fabs(m[r * n + c]) > fabs(m[piv * n + c])) piv = r;
        if (piv != c) {
            for (int j = 0; j < n; ++j) { double t = m[c * n + j]; m[c * n + j] = m[piv * n + j]; m[piv * n + j] = t; }
            double t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
        const double d = m[c * n + c];
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = m[r * n + c] / d;
            if (f == 0.0) continue;
            for (int j = c; j < n; ++j) m[r * n + j] -= f * m[c * n + j];
            b[r] -= f * b[c];
        }
    }
    for (int c = 0; c < n; ++c) b[c] /= m[c * n + c];
}

static void als_pass(int dim, const double *x, int64_t nx, double *y, int64_t ny, const int32_t *row,
                     int64_t n_active, const int32_t *col, int c_pos, double reg, double *HH, double *M,
                     double *b)
{
    for (int f1 = 0; f1 < dim; ++f1)                     /* :257-262 */
        for (int f2 = 0; f2 < dim; ++f2) {
            double d = 0.0;
            for (int64_t i = 0; i < nx; ++i) d += x[f1 * nx + i] * x[f2 * nx + i];
            HH[f1 + dim * f2] = d;
        }
    int64_t start = 0;
    for (int64_t j = 0; j < n_active; ++j) {
        start += row[j];                                  /* :267 */
        const int64_t span = row[j + 1];
        for (int f1 = 0; f1 < dim; ++f1)
            for (int f2 = 0; f2 < dim; ++f2) {
                double d = 0.0;
                for (int64_t i = 0; i < span; ++i) {
                    const int64_t id = col[start + i];
                    d += x[f1 * nx + id] * x[f2 * nx + id] * c_pos;
                }
                M[f1 * dim + f2] = HH[f1 + dim * f2] + d + (f1 == f2 ? reg : 0.0);
            }
        for (int f = 0; f < dim; ++f) {
            double d = 0.0;
            for (int64_t i = 0; i < span; ++i) d += x[f * nx + col[start + i]] * (1 + c_pos);
            b[f] = d;
        }
        solve_dense(M, b, dim);
        for (int f = 0; f < dim; ++f) y[f * ny + j] = b[f];
    }
}

void oracle_als_wrmf(int nbr_epochs, int dim, double *u, double *v, const int32_t *users_row,
                     int64_t n_users_row, const int32_t *users_col, const int32_t *items_row,
                     int64_t n_items_row, const int32_t *items_col, int64_t nu, int64_t ni, int c_pos,
                     double reg)
{
    double *HH = (double *)malloc((size_t)dim * dim * sizeof(double));
    double *M = (double *)malloc((size_t)dim * dim * sizeof(double));
    double *b = (double *)malloc((size_t)dim * sizeof(double));
    for (int e = 0; e < nbr_epochs; ++e) {
        als_pass(dim, u, ni, v, nu, users_row, n_users_row - 1, users_col, c_pos, reg, HH, M, b);
        als_pass(dim, v, nu, u, ni, items_row, n_items_row - 1, items_col, c_pos, reg, HH, M, b);
    }
    free(HH); free(M); free(b);
}

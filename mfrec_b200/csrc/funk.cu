// Funk-SVD per-feature SGD (gd_estimator.pyx) -- placeholder until the kernel lands.
#include "common.cuh"

extern "C" int mfrec_train_funk(mfrec_ctx *ctx, int, int, int, double, int, double, double, double, double,
                                double *, double *, const int32_t *, const double *, int64_t, int32_t,
                                int32_t, const double *, const double *, int, int, const mfrec_opts *,
                                int32_t *, double *)
{
    return mfrec_set_error(ctx, MFREC_ERR_UNSUPPORTED, "mfrec_train_funk: not implemented yet");
}

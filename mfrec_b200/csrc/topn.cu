// Top-N scoring -- placeholder until the kernel lands.
#include "common.cuh"

extern "C" int mfrec_topn(mfrec_ctx *ctx, int, int, const double *, const double *, int32_t, int32_t,
                          const int32_t *, int32_t, int32_t, const int64_t *, const int32_t *, double,
                          const double *, const double *, double, double, int32_t, int32_t *, double *,
                          int32_t *)
{
    return mfrec_set_error(ctx, MFREC_ERR_UNSUPPORTED, "mfrec_topn: not implemented yet");
}

// Alternating least squares for implicit-feedback WRMF (Hu, Koren, Volinsky, ICDM 2008):
// als_wrmf of the reference (mfrec/lib/als_implicit.pyx:208-352), driven by
// WRMFRecommender.train (mfrec/recommendation/wrmf.py:83-110).  SURVEY.md 8(f) #3 -- the step
// next to the SGD hot path, not part of it.
//
// Per epoch, with U the item factors and V the user factors (both feature-major [k][n] float64 at
// the boundary, row-major [n][k] float64 on the device):
//   user pass:  HH = sum_i u_i u_i^T ;  for every active user j with rated items S_j
//               M = HH + c_pos * sum_{i in S_j} u_i u_i^T + reg * I ,  b = (1 + c_pos) * sum_{i in S_j} u_i
//               v_j = M^-1 b                        (als_implicit.pyx:257-306)
//   item pass:  the same with the roles swapped, on the UPDATED V   (:308-352)
// Inside a pass the rows are independent, so one CTA solves one row: M is assembled in shared
// memory (k^2 doubles: k <= 160), factorised by Cholesky (M is symmetric positive definite:
// Gram matrices plus reg > 0) and solved by two triangular sweeps.  The reference inverts M with
// numpy.linalg.inv and multiplies; the two agree to float64 round-off for these well-conditioned
// systems (tests compare at 1e-9).  Everything is float64: this is a dense small-matrix problem
// at the reference's own sizes (k = 20 in its example), not a bandwidth problem.
//
// The CSR-like inputs are the reference's own (mfrec/lib/datasets.py:13-32): row = [0, count_0,
// count_1, ...] (counts, NOT offsets; the kernel loop accumulates them, als_implicit.pyx:266-268),
// col = neighbour ids in row order.
#include <vector>

#include "common.cuh"

namespace {

constexpr int kTileMin = 8, kTileMax = 32;   // neighbour rows staged per step (as many as keep >= 4 CTAs per SM)

__global__ void __launch_bounds__(256)
to_rows_f64_kernel(const double *__restrict__ src_kn, int k, int64_t n, double *__restrict__ dst_nk)
{
    __shared__ double tile[32][33];
    const int64_t j0 = (int64_t)blockIdx.x * 32;
    const int f0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int f = f0 + r;
        const int64_t j = j0 + tx;
        tile[r][tx] = (f < k && j < n) ? src_kn[(int64_t)f * n + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t j = j0 + r;
        const int f = f0 + tx;
        if (j < n && f < k) dst_nk[j * k + f] = tile[tx][r];
    }
}

__global__ void __launch_bounds__(256)
from_rows_f64_kernel(const double *__restrict__ src_nk, int k, int64_t n, double *__restrict__ dst_kn)
{
    __shared__ double tile[32][33];
    const int64_t j0 = (int64_t)blockIdx.x * 32;
    const int f0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int64_t j = j0 + r;
        const int f = f0 + tx;
        tile[r][tx] = (j < n && f < k) ? src_nk[j * k + f] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int f = f0 + r;
        const int64_t j = j0 + tx;
        if (f < k && j < n) dst_kn[(int64_t)f * n + j] = tile[tx][r];
    }
}

// partial Gram matrices: block b sums x x^T over its slice of rows, in row order
__global__ void __launch_bounds__(256)
gram_partial_kernel(const double *__restrict__ X, int64_t n, int k, int64_t rows_per_block, double *__restrict__ part)
{
    const int64_t a = (int64_t)blockIdx.x * rows_per_block, b = min(n, a + rows_per_block);
    for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
        const int f1 = e / k, f2 = e % k;
        double acc = 0.0;
        for (int64_t i = a; i < b; ++i) acc += X[i * k + f1] * X[i * k + f2];
        part[(size_t)blockIdx.x * k * k + e] = acc;
    }
}

__global__ void gram_reduce_kernel(const double *__restrict__ part, int nblocks, int kk, double *__restrict__ HH)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= kk) return;
    double acc = 0.0;
    for (int b = 0; b < nblocks; ++b) acc += part[(size_t)b * kk + e];   // fixed order
    HH[e] = acc;
}

// offsets from the reference's count array: off[j] = sum_{t <= j} row[t]  (row[0] = 0)
__global__ void validate_cols_kernel(const int32_t *__restrict__ col, int64_t n, int32_t limit, int32_t *__restrict__ bad)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        if (col[j] < 0 || col[j] >= limit) atomicOr(bad, 1);
}

// one CTA per active row j: assemble M and b, Cholesky with the forward solve folded in, back
// substitution, write Y[j].  kp = the pitch of a staged neighbour row: k rounded up so that the
// register tiles below never leave the row (the padding is zero).
__host__ __device__ inline int als_pitch(int k) { return ((k + 7) / 8) * 8; }

__global__ void __launch_bounds__(128)
als_solve_kernel(const double *__restrict__ X, const double *__restrict__ HH, const int64_t *__restrict__ off,
                 const int32_t *__restrict__ col, int k, double c_pos, double reg, double *__restrict__ Y, int kTile)
{
    extern __shared__ double sm[];
    const int kp = als_pitch(k);
    double *M = sm;                 // [k][k]
    double *b = M + k * k;          // [k]
    double *colv = b + k;           // [k]  column c of L during the factorisation
    double *xs = colv + k;          // [kTile][kp]
    const int j = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < k * k; e += nt) M[e] = HH[e];
    for (int f = tid; f < k; f += nt) b[f] = 0.0;
    for (int e = tid; e < kTile * (kp - k); e += nt) xs[(e / (kp - k)) * kp + k + e % (kp - k)] = 0.0;   // (once per row solve)
    __syncthreads();
    for (int f = tid; f < k; f += nt) M[f * k + f] += reg;
    // register tile of the Gram update: 4 rows x 8 columns of M per thread (rows ti * 4 + i, columns
    // tj + ntc * jj: neighbouring lanes read neighbouring doubles, lanes of one ti share their row
    // loads): 12 shared-memory loads per 32 multiply-adds and neighbour instead of 64 (round 1: one
    // entry per thread and step; the loop was bound by its LDS traffic and by e / k, e % k).
    const int ntr = (k + 3) / 4, ntc = kp / 8;
    const int64_t a = off[j], z = off[j + 1];
    for (int64_t t0 = a; t0 < z; t0 += kTile) {
        const int nrow = (int)min((int64_t)kTile, z - t0);
        __syncthreads();
        for (int r = warp; r < nrow; r += 4) {   // a warp copies a neighbour's row (coalesced)
            const double *src = X + (int64_t)col[t0 + r] * k;
            for (int f = lane; f < k; f += 32) xs[r * kp + f] = src[f];
        }
        __syncthreads();
        for (int tile = tid; tile < ntr * ntc; tile += nt) {
            const int ti = tile / ntc, tj = tile - ti * ntc;
            // rows past k (k not a multiple of 4) read a valid, ignored entry
            const int r0 = min(ti * 4, k - 1), r1 = min(ti * 4 + 1, k - 1), r2 = min(ti * 4 + 2, k - 1), r3 = min(ti * 4 + 3, k - 1);
            double acc[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) acc[i][jj] = 0.0;
            const double *xr = xs;
            for (int r = 0; r < nrow; ++r, xr += kp) {
                const double rv[4] = {xr[r0], xr[r1], xr[r2], xr[r3]};
                double cv[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) cv[jj] = xr[tj + ntc * jj];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) acc[i][jj] += rv[i] * cv[jj];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int f1 = ti * 4 + i, f2 = tj + ntc * jj;
                    if (f1 < k && f2 < k) M[f1 * k + f2] += c_pos * acc[i][jj];
                }
        }
        for (int f = tid; f < k; f += nt) {
            double acc = 0.0;
            for (int r = 0; r < nrow; ++r) acc += xs[r * kp + f];
            b[f] += (1.0 + c_pos) * acc;
        }
    }
    __syncthreads();
    // Cholesky M = L L^T in place (lower triangle), right-looking, two barriers per column, with the
    // forward solve L y = b riding along as one more row of the trailing update (b[r] -= L[r][c] y[c]).
    // The trailing update runs on an 8 x 16 thread grid (rows strided by 8, columns by 16: no division to
    // find an entry -- round 1 decoded (row, column) from a flat index with e / m and e % m and walked
    // the full square, 123 k of the kernel's 216 k warp instructions per row) and reads column c from
    // a contiguous copy (M[c2][c] for neighbouring c2 is one bank).
    {
        const int tr = tid >> 4, tc = tid & 15;
        for (int c = 0; c < k; ++c) {
            const double d = sqrt(M[c * k + c]);   // (every thread: the entry is final since the last barrier)
            const double yc = b[c] / d;            // y[c] (likewise)
            for (int r = c + 1 + tid; r < k; r += nt) {
                const double v = M[r * k + c] / d;
                M[r * k + c] = v;
                colv[r] = v;
            }
            __syncthreads();
            if (tid == 0) { M[c * k + c] = d; b[c] = yc; }
            for (int r = c + 1 + tr; r < k; r += 8) {
                const double lr = colv[r];
                for (int c2 = c + 1 + tc; c2 <= r; c2 += 16) M[r * k + c2] -= lr * colv[c2];
                if (tc == 15) b[r] -= lr * yc;
            }
            __syncthreads();
        }
    }
    // back substitution L^T x = y, column-oriented on one warp (row r of L is contiguous):
    // x[r] = y[r] / L[r][r], then y[c] -= L[r][c] x[r] for every c < r
    if (tid < 32) {
        for (int r = k - 1; r >= 0; --r) {
            const double xr_ = b[r] / M[r * k + r];
            for (int c = lane; c < r; c += 32) b[c] -= M[r * k + c] * xr_;
            __syncwarp();
            if (lane == 0) b[r] = xr_;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int f = tid; f < k; f += nt) Y[(int64_t)j * k + f] = b[f];
}

int gram(mfrec_ctx *ctx, const double *X, int64_t n, int k, double *part, int nblocks, double *HH)
{
    const int64_t rpb = (n + nblocks - 1) / nblocks;
    gram_partial_kernel<<<nblocks, 256, 0, ctx->stream>>>(X, n, k, rpb, part);
    MF_LAUNCH_CHECK(ctx);
    gram_reduce_kernel<<<(k * k + 255) / 256, 256, 0, ctx->stream>>>(part, nblocks, k * k, HH);
    MF_LAUNCH_CHECK(ctx);
    return MFREC_OK;
}

// host: offsets of the active rows from the reference's [0, counts...] array
std::vector<int64_t> offsets_from_counts(const int32_t *row, int64_t n_row)
{
    std::vector<int64_t> off(n_row > 0 ? n_row : 1, 0);
    int64_t start = 0;
    for (int64_t j = 0; j + 1 < n_row; ++j) {
        start += row[j];                 // als_implicit.pyx:267
        off[j] = start;
        off[j + 1] = start + row[j + 1];
    }
    return off;
}

}  // namespace

extern "C" int mfrec_train_als_wrmf(mfrec_ctx *ctx, int nbr_epochs, int k, double *u, double *v,
                                    const int32_t *users_row, int64_t n_users_row, const int32_t *users_col,
                                    const int32_t *items_row, int64_t n_items_row, const int32_t *items_col,
                                    int32_t nbr_users, int32_t nbr_items, int c_pos, double reg)
{
    if (!ctx || !u || !v || !users_row || !items_row)
        return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_train_als_wrmf: NULL argument");
    if (k <= 0 || nbr_users <= 0 || nbr_items <= 0 || nbr_epochs < 0 || n_users_row < 1 || n_items_row < 1)
        return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_train_als_wrmf: k=%d users=%d items=%d epochs=%d", k,
                               nbr_users, nbr_items, nbr_epochs);
    const int64_t nau = n_users_row - 1, nai = n_items_row - 1;   // active rows (als_implicit.pyx:248-249)
    if (nau > nbr_users || nai > nbr_items)
        return mfrec_set_error(ctx, MFREC_ERR_INDEX, "mfrec_train_als_wrmf: more rows in the sparse structure than users / items");
    MF_CUDA(ctx, cudaSetDevice(ctx->device));
    // neighbour rows per staging step: the loads of a step are one L2 round trip, so more rows per step
    // hide it better, as long as four CTAs still fit an SM
    int kTile = kTileMax;
    while (kTile > kTileMin && ((size_t)k * k + 2 * k + (size_t)kTile * als_pitch(k)) * sizeof(double) * 4 > ctx->smem_per_sm) kTile /= 2;
    const size_t smem = ((size_t)k * k + 2 * k + (size_t)kTile * als_pitch(k)) * sizeof(double);
    if (smem > ctx->smem_optin)
        return mfrec_set_error(ctx, MFREC_ERR_UNSUPPORTED, "mfrec_train_als_wrmf: k=%d needs %zu B of shared memory", k, smem);
    if (nbr_epochs == 0) return MFREC_OK;
    const std::vector<int64_t> uoff = offsets_from_counts(users_row, n_users_row);
    const std::vector<int64_t> ioff = offsets_from_counts(items_row, n_items_row);
    const int64_t nuc = uoff[nau], nic = ioff[nai];
    if ((nuc > 0 && !users_col) || (nic > 0 && !items_col))
        return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_train_als_wrmf: NULL column array");
    for (int64_t j = 0; j < nau; ++j)
        if (uoff[j + 1] < uoff[j]) return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_train_als_wrmf: negative count");
    for (int64_t j = 0; j < nai; ++j)
        if (ioff[j + 1] < ioff[j]) return mfrec_set_error(ctx, MFREC_ERR_BAD_ARG, "mfrec_train_als_wrmf: negative count");
    cudaStream_t st = ctx->stream;
    const int nblocks = ctx->sm_count;
    DevBuf<double> d_stage, d_U, d_V, d_part, d_HH;
    DevBuf<int64_t> d_uoff, d_ioff;
    DevBuf<int32_t> d_ucol, d_icol, d_bad;
    const int64_t nmax = std::max<int64_t>(nbr_users, nbr_items);
    MF_CUDA(ctx, d_stage.alloc((size_t)k * nmax, ctx->stream));
    MF_CUDA(ctx, d_U.alloc((size_t)k * nbr_items, ctx->stream));
    MF_CUDA(ctx, d_V.alloc((size_t)k * nbr_users, ctx->stream));
    MF_CUDA(ctx, d_part.alloc((size_t)nblocks * k * k, ctx->stream));
    MF_CUDA(ctx, d_HH.alloc((size_t)k * k, ctx->stream));
    MF_CUDA(ctx, d_uoff.alloc(uoff.size(), ctx->stream));
    MF_CUDA(ctx, d_ioff.alloc(ioff.size(), ctx->stream));
    MF_CUDA(ctx, d_ucol.alloc(nuc, ctx->stream));
    MF_CUDA(ctx, d_icol.alloc(nic, ctx->stream));
    MF_CUDA(ctx, d_bad.alloc(1, ctx->stream));
    MF_CUDA(ctx, cudaMemsetAsync(d_bad.p, 0, 4, st));
    MF_CUDA(ctx, cudaMemcpyAsync(d_uoff.p, uoff.data(), uoff.size() * 8, cudaMemcpyHostToDevice, st));
    MF_CUDA(ctx, cudaMemcpyAsync(d_ioff.p, ioff.data(), ioff.size() * 8, cudaMemcpyHostToDevice, st));
    if (nuc) MF_CUDA(ctx, cudaMemcpyAsync(d_ucol.p, users_col, (size_t)nuc * 4, cudaMemcpyHostToDevice, st));
    if (nic) MF_CUDA(ctx, cudaMemcpyAsync(d_icol.p, items_col, (size_t)nic * 4, cudaMemcpyHostToDevice, st));
    if (nuc) { validate_cols_kernel<<<ctx->sm_count, 256, 0, st>>>(d_ucol.p, nuc, nbr_items, d_bad.p); MF_LAUNCH_CHECK(ctx); }
    if (nic) { validate_cols_kernel<<<ctx->sm_count, 256, 0, st>>>(d_icol.p, nic, nbr_users, d_bad.p); MF_LAUNCH_CHECK(ctx); }
    int32_t h_bad = 0;
    MF_CUDA(ctx, cudaMemcpyAsync(&h_bad, d_bad.p, 4, cudaMemcpyDeviceToHost, st));
    // factors: [k][n] host -> [n][k] device
    MF_CUDA(ctx, cudaMemcpyAsync(d_stage.p, u, (size_t)k * nbr_items * 8, cudaMemcpyHostToDevice, st));
    to_rows_f64_kernel<<<dim3((unsigned)ceil_div64(nbr_items, 32), (k + 31) / 32), 256, 0, st>>>(d_stage.p, k, nbr_items, d_U.p);
    MF_LAUNCH_CHECK(ctx);
    MF_CUDA(ctx, cudaStreamSynchronize(st));   // `u` is borrowed; the stage buffer is reused for v
    if (h_bad) return mfrec_set_error(ctx, MFREC_ERR_INDEX, "mfrec_train_als_wrmf: neighbour id out of range");
    MF_CUDA(ctx, cudaMemcpyAsync(d_stage.p, v, (size_t)k * nbr_users * 8, cudaMemcpyHostToDevice, st));
    to_rows_f64_kernel<<<dim3((unsigned)ceil_div64(nbr_users, 32), (k + 31) / 32), 256, 0, st>>>(d_stage.p, k, nbr_users, d_V.p);
    MF_LAUNCH_CHECK(ctx);
    MF_CUDA(ctx, cudaFuncSetAttribute(als_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int e = 0; e < nbr_epochs; ++e) {
        MF_TRY(gram(ctx, d_U.p, nbr_items, k, d_part.p, nblocks, d_HH.p));
        if (nau > 0) {
            als_solve_kernel<<<(unsigned)nau, 128, smem, st>>>(d_U.p, d_HH.p, d_uoff.p, d_ucol.p, k, (double)c_pos, reg, d_V.p, kTile);
            MF_LAUNCH_CHECK(ctx);
        }
        MF_TRY(gram(ctx, d_V.p, nbr_users, k, d_part.p, nblocks, d_HH.p));
        if (nai > 0) {
            als_solve_kernel<<<(unsigned)nai, 128, smem, st>>>(d_V.p, d_HH.p, d_ioff.p, d_icol.p, k, (double)c_pos, reg, d_U.p, kTile);
            MF_LAUNCH_CHECK(ctx);
        }
    }
    from_rows_f64_kernel<<<dim3((unsigned)ceil_div64(nbr_users, 32), (k + 31) / 32), 256, 0, st>>>(d_V.p, k, nbr_users, d_stage.p);
    MF_LAUNCH_CHECK(ctx);
    MF_CUDA(ctx, cudaMemcpyAsync(v, d_stage.p, (size_t)k * nbr_users * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaStreamSynchronize(st));
    from_rows_f64_kernel<<<dim3((unsigned)ceil_div64(nbr_items, 32), (k + 31) / 32), 256, 0, st>>>(d_U.p, k, nbr_items, d_stage.p);
    MF_LAUNCH_CHECK(ctx);
    MF_CUDA(ctx, cudaMemcpyAsync(u, d_stage.p, (size_t)k * nbr_items * 8, cudaMemcpyDeviceToHost, st));
    MF_CUDA(ctx, cudaStreamSynchronize(st));
    return MFREC_OK;
}

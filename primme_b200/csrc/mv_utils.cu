// mv_utils.cu -- K6: multivector utilities (fp64): column permute/copy, per-column axpy/scale,
// batched column dots, in-place residual, Jacobi preconditioner.  All are single-pass,
// coalesced, grid-stride kernels; the per-column scalars come from the host through the
// context's pinned staging buffer (replaces cublasAxpyEx/DotcEx/ScalEx + cudaMemcpy2D calls of
// reference src/linalg/cublas_wrapper.c:616-705,739-783 and the column-by-column permute of
// src/linalg/auxiliary.c:763-779).
#include "pb200_internal.cuh"
#include <math.h>
#include <string.h>
#include <vector>

namespace {

constexpr int UT = 256;

inline int grid_for(pb200_ctx *ctx, int64_t n) {
   int64_t g = (n + UT - 1) / UT;
   int64_t cap = (int64_t)ctx->num_sms * 8;
   if (g > cap) g = cap;
   if (g < 1) g = 1;
   return (int)g;
}

// Y(:, yin[j]) = X(:, xin[j])  (index arrays in device memory, -1 list = identity)
__global__ void copy_cols_kernel(int64_t n, const double *__restrict__ X, int64_t ldx,
      const int *__restrict__ xin, double *__restrict__ Y, int64_t ldy,
      const int *__restrict__ yin, int ncols) {
   for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
         r += (int64_t)gridDim.x * blockDim.x)
      for (int j = 0; j < ncols; j++) {
         int xs = xin ? xin[j] : j, ys = yin ? yin[j] : j;
         Y[r + (size_t)ys * ldy] = X[r + (size_t)xs * ldx];
      }
}

// in-place row-wise gather: each thread owns a row, reads all its entries, then writes
template <int CH>
__global__ void permute_rows_kernel(
      int64_t n, double *__restrict__ X, int64_t ldx, const int *__restrict__ perm, int c0, int nc) {
   // columns [c0, c0+nc) of the cycle-closed set are handled by this launch (nc <= CH)
   for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
         r += (int64_t)gridDim.x * blockDim.x) {
      double v[CH];
#pragma unroll
      for (int j = 0; j < CH; j++)
         if (j < nc) v[j] = X[r + (size_t)perm[c0 + j] * ldx];
#pragma unroll
      for (int j = 0; j < CH; j++)
         if (j < nc) X[r + (size_t)(c0 + j) * ldx] = v[j];
   }
}

__global__ void axpy_cols_kernel(int64_t n, const double *__restrict__ alpha,
      const double *__restrict__ X, int64_t ldx, double *__restrict__ Y, int64_t ldy, int ncols) {
   for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
         r += (int64_t)gridDim.x * blockDim.x)
      for (int j = 0; j < ncols; j++) Y[r + (size_t)j * ldy] += alpha[j] * X[r + (size_t)j * ldx];
}

__global__ void scale_cols_kernel(
      int64_t n, const double *__restrict__ alpha, double *__restrict__ X, int64_t ldx, int ncols) {
   for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
         r += (int64_t)gridDim.x * blockDim.x)
      for (int j = 0; j < ncols; j++) X[r + (size_t)j * ldx] *= alpha[j];
}

// mode 0: out[j] = sum X_j .* Y_j ;  mode 1: W_j -= theta_j V_j (X=V, Y=W), out[j] = |W_j|^2
template <int NC>
__global__ void __launch_bounds__(UT) dots_kernel(int64_t n, const double *__restrict__ X,
      int64_t ldx, double *__restrict__ Y, int64_t ldy, int ncols, int mode,
      const double *__restrict__ theta, double *__restrict__ partials) {
   double acc[NC];
#pragma unroll
   for (int j = 0; j < NC; j++) acc[j] = 0.0;
   for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
         r += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
      for (int j = 0; j < NC; j++)
         if (j < ncols) {
            double x = X[r + (size_t)j * ldx], y = Y[r + (size_t)j * ldy];
            if (mode == 1) {
               y -= theta[j] * x;
               Y[r + (size_t)j * ldy] = y;
               acc[j] += y * y;
            } else
               acc[j] += x * y;
         }
   }
   __shared__ double red[UT / 32][NC];
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
   for (int j = 0; j < NC; j++) {
      double v = acc[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][j] = v;
   }
   __syncthreads();
   if (threadIdx.x < ncols) {
      double s = 0.0;
      for (int w = 0; w < UT / 32; w++) s += red[w][threadIdx.x];
      partials[(size_t)blockIdx.x * ncols + threadIdx.x] = s;
   }
}

__global__ void jacobi_kernel(int64_t n, const double *__restrict__ diag,
      const double *__restrict__ shifts, int has_shifts, double minabs,
      const double *__restrict__ X, int64_t ldx, double *__restrict__ Y, int64_t ldy, int ncols) {
   for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
         r += (int64_t)gridDim.x * blockDim.x) {
      double d0 = diag[r];
      for (int j = 0; j < ncols; j++) {
         double d = d0 - (has_shifts ? shifts[j] : 0.0);
         if (fabs(d) < minabs) d = d < 0 ? -minabs : minabs;
         Y[r + (size_t)j * ldy] = X[r + (size_t)j * ldx] / d;
      }
   }
}

int stage_doubles(pb200_ctx *ctx, const double *h, int cnt, size_t off) {
   PB_CHK(pb_ensure_small(ctx, off + cnt));
   memcpy(ctx->h_pinned + off, h, sizeof(double) * cnt);
   PB_CUDA(cudaMemcpyAsync(ctx->d_small + off, ctx->h_pinned + off, sizeof(double) * cnt,
         cudaMemcpyHostToDevice, ctx->stream));
   return 0;
}
int stage_ints(pb200_ctx *ctx, const int *h, int cnt, size_t off_doubles, const int **dev) {
   size_t dbl = ((size_t)cnt + 1) / 2;
   PB_CHK(pb_ensure_small(ctx, off_doubles + dbl));
   memcpy(ctx->h_pinned + off_doubles, h, sizeof(int) * cnt);
   PB_CUDA(cudaMemcpyAsync(ctx->d_small + off_doubles, ctx->h_pinned + off_doubles,
         sizeof(int) * cnt, cudaMemcpyHostToDevice, ctx->stream));
   *dev = (const int *)(ctx->d_small + off_doubles);
   return 0;
}

int dots_impl(pb200_ctx *ctx, int64_t n, const double *X, int64_t ldx, double *Y, int64_t ldy,
      int ncols, int mode, const double *theta_host, double *out_host) {
   for (int c0 = 0; c0 < ncols; c0 += 8) {
      int nc = ncols - c0 < 8 ? ncols - c0 : 8;
      // the pinned buffer is reused by pb_finish_panel: wait for earlier users
      if (mode == 1) PB_CHK(stage_doubles(ctx, theta_host + c0, nc, 0));
      int grid = grid_for(ctx, n);
      PB_CHK(pb_ensure_partials(ctx, (size_t)grid * nc));
      dots_kernel<8><<<grid, UT, 0, ctx->stream>>>(n, X + (size_t)c0 * ldx, ldx,
            Y + (size_t)c0 * ldy, ldy, nc, mode, ctx->d_small, ctx->d_partials);
      ctx->launches++;
      PB_CUDA(cudaGetLastError());
      PB_CHK(pb_finish_panel(ctx, grid, nc));
      for (int j = 0; j < nc; j++) out_host[c0 + j] = ctx->h_pinned[j];
   }
   return 0;
}

}  // namespace

extern "C" int pb200_dcopy_columns(pb200_ctx *ctx, int64_t n, const double *X, int64_t ldx,
      const int *xin_host, double *Y, int64_t ldy, const int *yin_host, int ncols) {
   if (ncols <= 0 || n <= 0) return 0;
   const int *dx = NULL, *dy = NULL;
   // earlier async uses of the pinned buffer must be complete before it is overwritten
   PB_CUDA(cudaStreamSynchronize(ctx->stream));
   if (xin_host) PB_CHK(stage_ints(ctx, xin_host, ncols, 0, &dx));
   if (yin_host) PB_CHK(stage_ints(ctx, yin_host, ncols, (size_t)(ncols + 1) / 2 + 1, &dy));
   copy_cols_kernel<<<grid_for(ctx, n), UT, 0, ctx->stream>>>(n, X, ldx, dx, Y, ldy, dy, ncols);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return 0;
}

extern "C" int pb200_dpermute_columns(
      pb200_ctx *ctx, int64_t n, double *X, int64_t ldx, const int *perm_host, int ncols) {
   if (ncols <= 0 || n <= 0) return 0;
   bool ident = true;
   for (int i = 0; i < ncols; i++) ident = ident && perm_host[i] == i;
   if (ident) return 0;
   // Row-wise in place is only safe if a thread reads every source it overwrites first; with at
   // most 32 columns per launch we instead go through scratch: tmp(:,i) = X(:,perm[i]) for the
   // moved columns only, then copy back.
   std::vector<int> moved;
   for (int i = 0; i < ncols; i++)
      if (perm_host[i] != i) moved.push_back(i);
   int nm = (int)moved.size();
   std::vector<int> src(nm), ident_idx(nm);
   for (int i = 0; i < nm; i++) src[i] = perm_host[moved[i]];
   PB_CHK(pb_ensure_scratch(ctx, sizeof(double) * (size_t)n * nm));
   double *tmp = (double *)ctx->d_scratch;
   PB_CHK(pb200_dcopy_columns(ctx, n, X, ldx, src.data(), tmp, n, NULL, nm));
   PB_CHK(pb200_dcopy_columns(ctx, n, tmp, n, NULL, X, ldx, moved.data(), nm));
   return 0;
}

extern "C" int pb200_daxpy_columns(pb200_ctx *ctx, int64_t n, const double *alpha_host,
      const double *X, int64_t ldx, double *Y, int64_t ldy, int ncols) {
   if (ncols <= 0 || n <= 0) return 0;
   PB_CUDA(cudaStreamSynchronize(ctx->stream));
   PB_CHK(stage_doubles(ctx, alpha_host, ncols, 0));
   axpy_cols_kernel<<<grid_for(ctx, n), UT, 0, ctx->stream>>>(n, ctx->d_small, X, ldx, Y, ldy, ncols);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return 0;
}

extern "C" int pb200_dscale_columns(
      pb200_ctx *ctx, int64_t n, const double *alpha_host, double *X, int64_t ldx, int ncols) {
   if (ncols <= 0 || n <= 0) return 0;
   PB_CUDA(cudaStreamSynchronize(ctx->stream));
   PB_CHK(stage_doubles(ctx, alpha_host, ncols, 0));
   scale_cols_kernel<<<grid_for(ctx, n), UT, 0, ctx->stream>>>(n, ctx->d_small, X, ldx, ncols);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return 0;
}

extern "C" int pb200_dcolumn_dots(pb200_ctx *ctx, int64_t n, const double *X, int64_t ldx,
      const double *Y, int64_t ldy, int ncols, double *out_host) {
   if (ncols <= 0) return 0;
   return dots_impl(ctx, n, X, ldx, (double *)Y, ldy, ncols, 0, NULL, out_host);
}

extern "C" int pb200_dresidual_inplace(pb200_ctx *ctx, int64_t n, const double *theta_host,
      const double *V, int64_t ldv, double *W, int64_t ldw, int ncols, double *out_host) {
   if (ncols <= 0) return 0;
   PB_CUDA(cudaStreamSynchronize(ctx->stream));
   return dots_impl(ctx, n, V, ldv, W, ldw, ncols, 1, theta_host, out_host);
}

extern "C" int pb200_djacobi(pb200_ctx *ctx, int64_t n, const double *diag,
      const double *shifts_host, double minabs, const double *X, int64_t ldx, double *Y,
      int64_t ldy, int ncols) {
   if (ncols <= 0 || n <= 0) return 0;
   PB_CUDA(cudaStreamSynchronize(ctx->stream));
   if (shifts_host) PB_CHK(stage_doubles(ctx, shifts_host, ncols, 0));
   jacobi_kernel<<<grid_for(ctx, n), UT, 0, ctx->stream>>>(
         n, diag, ctx->d_small, shifts_host != NULL, minabs, X, ldx, Y, ldy, ncols);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return 0;
}

// mv_utils.cu -- K6: multivector utilities (fp64): column permute/copy, per-column axpy/scale,
// batched column dots, in-place residual, Jacobi preconditioner, the fused QMR solution update.  All are
// single-pass, coalesced, grid-stride kernels; the per-column scalars and index lists travel as kernel
// parameters (no staging copy, no stream synchronisation per call) and the reductions finish inside the
// kernel (PbFin) -- one launch per operation (replaces cublasAxpyEx/DotcEx/ScalEx + cudaMemcpy2D calls of
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

// Y(:, idx.y[j]) = X(:, idx.x[j]); the index lists travel as kernel parameters
struct ColIdx {
   int x[32], y[32];
};
__global__ void copy_cols_kernel(int64_t n, const double *__restrict__ X, int64_t ldx,
      double *__restrict__ Y, int64_t ldy, int ncols, const ColIdx idx) {
   for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
         r += (int64_t)gridDim.x * blockDim.x)
      for (int j = 0; j < ncols; j++) Y[r + (size_t)idx.y[j] * ldy] = X[r + (size_t)idx.x[j] * ldx];
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

// per-column scalars travel in the kernel parameter space: no staging copy, no stream synchronisation
struct DScal {
   double v[8];
};

__global__ void axpy_cols_kernel(int64_t n, const DScal alpha,
      const double *__restrict__ X, int64_t ldx, double *__restrict__ Y, int64_t ldy, int ncols) {
   for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
         r += (int64_t)gridDim.x * blockDim.x)
      for (int j = 0; j < ncols; j++) Y[r + (size_t)j * ldy] += alpha.v[j] * X[r + (size_t)j * ldx];
}

__global__ void scale_cols_kernel(
      int64_t n, const DScal alpha, double *__restrict__ X, int64_t ldx, int ncols) {
   for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
         r += (int64_t)gridDim.x * blockDim.x)
      for (int j = 0; j < ncols; j++) X[r + (size_t)j * ldx] *= alpha.v[j];
}

// one step of the QMR solution update (reference inner_solve.c:395-413), all columns in one pass:
//   delta_j = gamma_j delta_j + eta_j d_j ;  sol_j += delta_j ;  partial |sol_j|^2 when wanted
// (the same roundings as scale + axpy + axpy + dot, one read and one write of every vector)
__global__ void __launch_bounds__(UT) qmr_update_kernel(int64_t n, const DScal gam, const DScal eta,
      const double *__restrict__ D, int64_t ldd, double *__restrict__ Dl, int64_t ldl, double *__restrict__ S,
      int64_t lds, int ncols, int want_dots, double *__restrict__ partials, const PbFin fin) {
   __shared__ double red[UT / 32][8];
   __shared__ int flag;
   double acc[8];
#pragma unroll
   for (int j = 0; j < 8; j++) acc[j] = 0.0;
   for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
      for (int j = 0; j < 8; j++)
         if (j < ncols) {
            double t = Dl[r + (size_t)j * ldl] * gam.v[j];
            t += eta.v[j] * D[r + (size_t)j * ldd];
            Dl[r + (size_t)j * ldl] = t;
            const double s = S[r + (size_t)j * lds] + t;
            S[r + (size_t)j * lds] = s;
            acc[j] += s * s;
         }
   }
   if (!want_dots) return;
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
   for (int j = 0; j < 8; j++) {
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
   pb_finish_device(fin, threadIdx.x, UT, 15, &flag);
}

// mode 0: out[j] = sum X_j .* Y_j ;  mode 1: W_j -= theta_j V_j (X=V, Y=W), out[j] = |W_j|^2
template <int NC>
__global__ void __launch_bounds__(UT) dots_kernel(int64_t n, const double *__restrict__ X,
      int64_t ldx, double *__restrict__ Y, int64_t ldy, int ncols, int mode,
      const DScal theta, double *__restrict__ partials, const PbFin fin) {
   __shared__ int flag;
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
               y -= theta.v[j] * x;
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
   pb_finish_device(fin, threadIdx.x, UT, 15, &flag);
}

__global__ void jacobi_kernel(int64_t n, const double *__restrict__ diag,
      const DScal shifts, int has_shifts, double minabs,
      const double *__restrict__ X, int64_t ldx, double *__restrict__ Y, int64_t ldy, int ncols) {
   for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
         r += (int64_t)gridDim.x * blockDim.x) {
      double d0 = diag[r];
      for (int j = 0; j < ncols; j++) {
         double d = d0 - (has_shifts ? shifts.v[j] : 0.0);
         if (fabs(d) < minabs) d = d < 0 ? -minabs : minabs;
         Y[r + (size_t)j * ldy] = X[r + (size_t)j * ldx] / d;
      }
   }
}

// partial-panel plumbing of the reducing utilities: in-kernel finish (PbFin: also the multi-rank peer
// exchange) when the launch shape allows it, else one partial per CTA for pb_finish_panel
int util_panel_setup(pb200_ctx *ctx, int grid, int cnt, PbFin *fin, double **partials) {
   const int r = pb_fin_prepare(ctx, grid, 1, cnt, fin);
   if (r < 0) return r;
   if (r == 1) {
      memset(fin, 0, sizeof(*fin));
      PB_CHK(pb_ensure_partials(ctx, (size_t)grid * cnt));
      *partials = ctx->d_partials;
   } else
      *partials = fin->partials;
   return 0;
}
int util_panel_collect(pb200_ctx *ctx, const PbFin *fin, int grid, int cnt) {
   if (fin->cnt > 0) return pb_collect_panel(ctx, fin);
   return pb_finish_panel(ctx, grid, cnt);
}
// a rank without local rows still takes part in the reduction; result (zeros summed with the others') in h_pinned
int util_empty_panel(pb200_ctx *ctx, int cnt) {
   if (ctx->nranks > 1) {
      PB_CHK(pb_ensure_small(ctx, (size_t)cnt));
      const int zr = pb_fin_contribute_zeros(ctx, cnt);
      if (zr < 0) return zr;
      if (zr == 1) {
         PB_CUDA(cudaMemsetAsync(ctx->d_panel, 0, sizeof(double) * cnt, ctx->stream));
         PB_CHK(pb_nccl_allreduce_dev(ctx, ctx->d_panel, cnt));
         PB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_panel, sizeof(double) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
         PB_CUDA(cudaStreamSynchronize(ctx->stream));
      }
   } else
      for (int j = 0; j < cnt; j++) ctx->h_pinned[j] = 0.0;
   return 0;
}

int dots_impl(pb200_ctx *ctx, int64_t n, const double *X, int64_t ldx, double *Y, int64_t ldy,
      int ncols, int mode, const double *theta_host, double *out_host) {
   for (int c0 = 0; c0 < ncols; c0 += 8) {
      int nc = ncols - c0 < 8 ? ncols - c0 : 8;
      if (n <= 0) {
         PB_CHK(util_empty_panel(ctx, nc));
      } else {
         DScal th;
         memset(&th, 0, sizeof(th));
         if (mode == 1)
            for (int j = 0; j < nc; j++) th.v[j] = theta_host[c0 + j];
         int grid = grid_for(ctx, n);
         PbFin fin;
         double *partials = NULL;
         PB_CHK(util_panel_setup(ctx, grid, nc, &fin, &partials));
         int ps = pb_prof_begin(ctx, PB_K_UTIL);
         dots_kernel<8><<<grid, UT, 0, ctx->stream>>>(n, X + (size_t)c0 * ldx, ldx,
               Y + (size_t)c0 * ldy, ldy, nc, mode, th, partials, fin);
         pb_prof_end(ctx, ps, 8.0 * (double)n * nc * (mode == 1 ? 3 : 2));
         ctx->launches++;
         PB_CUDA(cudaGetLastError());
         PB_CHK(util_panel_collect(ctx, &fin, grid, nc));
      }
      for (int j = 0; j < nc; j++) out_host[c0 + j] = ctx->h_pinned[j];
   }
   return 0;
}

}  // namespace

extern "C" int pb200_dcopy_columns(pb200_ctx *ctx, int64_t n, const double *X, int64_t ldx,
      const int *xin_host, double *Y, int64_t ldy, const int *yin_host, int ncols) {
   if (ncols <= 0 || n <= 0) return 0;
   for (int c0 = 0; c0 < ncols; c0 += 32) {
      const int nc = ncols - c0 < 32 ? ncols - c0 : 32;
      ColIdx idx;
      for (int j = 0; j < nc; j++)
         idx.x[j] = xin_host ? xin_host[c0 + j] : c0 + j, idx.y[j] = yin_host ? yin_host[c0 + j] : c0 + j;
      copy_cols_kernel<<<grid_for(ctx, n), UT, 0, ctx->stream>>>(n, X, ldx, Y, ldy, nc, idx);
      ctx->launches++;
      PB_CUDA(cudaGetLastError());
   }
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
   for (int c0 = 0; c0 < ncols; c0 += 8) {
      const int nc = ncols - c0 < 8 ? ncols - c0 : 8;
      DScal al;
      memset(&al, 0, sizeof(al));
      memcpy(al.v, alpha_host + c0, sizeof(double) * nc);
      int ps = pb_prof_begin(ctx, PB_K_UTIL);
      axpy_cols_kernel<<<grid_for(ctx, n), UT, 0, ctx->stream>>>(n, al, X + (size_t)c0 * ldx, ldx, Y + (size_t)c0 * ldy, ldy, nc);
      pb_prof_end(ctx, ps, 24.0 * (double)n * nc);
      ctx->launches++;
      PB_CUDA(cudaGetLastError());
   }
   return 0;
}

extern "C" int pb200_dscale_columns(
      pb200_ctx *ctx, int64_t n, const double *alpha_host, double *X, int64_t ldx, int ncols) {
   if (ncols <= 0 || n <= 0) return 0;
   for (int c0 = 0; c0 < ncols; c0 += 8) {
      const int nc = ncols - c0 < 8 ? ncols - c0 : 8;
      DScal al;
      memset(&al, 0, sizeof(al));
      memcpy(al.v, alpha_host + c0, sizeof(double) * nc);
      int ps = pb_prof_begin(ctx, PB_K_UTIL);
      scale_cols_kernel<<<grid_for(ctx, n), UT, 0, ctx->stream>>>(n, al, X + (size_t)c0 * ldx, ldx, nc);
      pb_prof_end(ctx, ps, 16.0 * (double)n * nc);
      ctx->launches++;
      PB_CUDA(cudaGetLastError());
   }
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
   return dots_impl(ctx, n, V, ldv, W, ldw, ncols, 1, theta_host, out_host);
}

extern "C" int pb200_djacobi(pb200_ctx *ctx, int64_t n, const double *diag,
      const double *shifts_host, double minabs, const double *X, int64_t ldx, double *Y,
      int64_t ldy, int ncols) {
   if (ncols <= 0 || n <= 0) return 0;
   for (int c0 = 0; c0 < ncols; c0 += 8) {
      const int nc = ncols - c0 < 8 ? ncols - c0 : 8;
      DScal sh;
      memset(&sh, 0, sizeof(sh));
      if (shifts_host) memcpy(sh.v, shifts_host + c0, sizeof(double) * nc);
      int ps = pb_prof_begin(ctx, PB_K_UTIL);
      jacobi_kernel<<<grid_for(ctx, n), UT, 0, ctx->stream>>>(n, diag, sh, shifts_host != NULL, minabs,
            X + (size_t)c0 * ldx, ldx, Y + (size_t)c0 * ldy, ldy, nc);
      pb_prof_end(ctx, ps, 16.0 * (double)n * nc + 8.0 * (double)n);
      ctx->launches++;
      PB_CUDA(cudaGetLastError());
   }
   return 0;
}

// delta_j = gamma_j delta_j + eta_j d_j;  sol_j += delta_j;  dots_host[j] = |sol_j|^2 when dots_host != NULL
// (the solution update of one QMR step, reference src/eigs/inner_solve.c:384-413, in one pass)
extern "C" int pb200_dqmr_update(pb200_ctx *ctx, int64_t n, const double *gamma_host, const double *eta_host,
      const double *D, int64_t ldd, double *Delta, int64_t ldl, double *Sol, int64_t lds, int ncols,
      double *dots_host) {
   for (int c0 = 0; c0 < ncols; c0 += 8) {
      const int nc = ncols - c0 < 8 ? ncols - c0 : 8;
      if (n <= 0) {
         if (dots_host) {
            PB_CHK(util_empty_panel(ctx, nc));
            for (int j = 0; j < nc; j++) dots_host[c0 + j] = ctx->h_pinned[j];
         }
         continue;
      }
      DScal g, e;
      memset(&g, 0, sizeof(g)), memset(&e, 0, sizeof(e));
      memcpy(g.v, gamma_host + c0, sizeof(double) * nc), memcpy(e.v, eta_host + c0, sizeof(double) * nc);
      const int grid = grid_for(ctx, n);
      PbFin fin;
      memset(&fin, 0, sizeof(fin));
      double *partials = NULL;
      if (dots_host) PB_CHK(util_panel_setup(ctx, grid, nc, &fin, &partials));
      int ps = pb_prof_begin(ctx, PB_K_UTIL);
      qmr_update_kernel<<<grid, UT, 0, ctx->stream>>>(n, g, e, D + (size_t)c0 * ldd, ldd, Delta + (size_t)c0 * ldl, ldl,
            Sol + (size_t)c0 * lds, lds, nc, dots_host != NULL, partials, fin);
      pb_prof_end(ctx, ps, 40.0 * (double)n * nc);
      ctx->launches++;
      PB_CUDA(cudaGetLastError());
      if (dots_host) {
         PB_CHK(util_panel_collect(ctx, &fin, grid, nc));
         for (int j = 0; j < nc; j++) dots_host[c0 + j] = ctx->h_pinned[j];
      }
   }
   return 0;
}

// ------------------------------------------------------------------------------------------
// LAPACK's dlarnv(idist = 2) on the device, bit for bit.  The reference fills its random vectors on the
// host (Num_larnv, src/linalg/blaslapack.c:953-977 -> dlarnv -> dlaruv): at n = 10^6 and a block of 4 that
// is 40 ms of a 0.6 s solve, and the vector still has to cross PCIe.  dlaruv is a 48-bit multiplicative
// congruential generator, x_k = a^k x_0 mod 2^48 with a = 33952834046453 (its table MM holds a^1 .. a^128
// in base 4096), value k = x_k / 2^48 assembled from the four 12-bit digits (exact in fp64), dlarnv maps
// it to 2 u - 1.  A power of a is reached by square-and-multiply, so every thread starts its own run of
// the SAME sequence; the seed returned is x_n in dlaruv's digit form, exactly what dlarnv leaves behind.
namespace {
constexpr unsigned long long LARNV_A = 33952834046453ull, LARNV_MASK = (1ull << 48) - 1;
__host__ __device__ inline unsigned long long larnv_pow(unsigned long long k) {
   unsigned long long r = 1, b = LARNV_A;
   while (k) {
      if (k & 1) r = (r * b) & LARNV_MASK;
      b = (b * b) & LARNV_MASK;
      k >>= 1;
   }
   return r;
}
constexpr int LARNV_RUN = 64;  // consecutive values per thread
__global__ void __launch_bounds__(256) larnv_kernel(unsigned long long x0, int64_t count, int64_t col_len, int64_t ld,
      double *__restrict__ X) {
   // value k of the stream (k = 0 .. count-1) is element (k % col_len) of column k / col_len
   const int64_t nruns = (count + LARNV_RUN - 1) / LARNV_RUN;
   for (int64_t run = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; run < nruns; run += (int64_t)gridDim.x * blockDim.x) {
      const int64_t k0 = run * LARNV_RUN;
      unsigned long long x = (x0 * larnv_pow((unsigned long long)k0 + 1)) & LARNV_MASK;
      const int64_t k1 = k0 + LARNV_RUN < count ? k0 + LARNV_RUN : count;
      int64_t col = k0 / col_len, row = k0 % col_len;
      for (int64_t k = k0; k < k1; k++) {
         const double r = 1.0 / 4096.0;
         const double u = r * ((double)(x >> 36) + r * ((double)((x >> 24) & 4095) + r * ((double)((x >> 12) & 4095) + r * (double)(x & 4095))));
         X[row + col * ld] = 2.0 * u - 1.0;
         x = (x * LARNV_A) & LARNV_MASK;
         if (++row == col_len) row = 0, col++;
      }
   }
}
}  // namespace

// X(0:col_len, 0:ncols) (leading dimension ld, in doubles) = the next col_len * ncols values of dlarnv(2, iseed):
// column after column, like ncols successive calls; iseed is advanced as dlarnv would.  For complex blocks pass
// col_len = 2 n and ld = 2 ldx (interleaved parts, the reference draws 2 n values per column).
extern "C" int pb200_dlarnv(pb200_ctx *ctx, long long iseed[4], int64_t col_len, int ncols, double *X, int64_t ld) {
   if (col_len <= 0 || ncols <= 0) return 0;
   for (int i = 0; i < 4; i++)
      if (iseed[i] < 0 || iseed[i] > 4095) return PB200_ERR_ARG;
   const unsigned long long x0 = ((unsigned long long)iseed[0] << 36) | ((unsigned long long)iseed[1] << 24) |
                                 ((unsigned long long)iseed[2] << 12) | (unsigned long long)iseed[3];
   const int64_t count = col_len * ncols;
   const int64_t nruns = (count + LARNV_RUN - 1) / LARNV_RUN;
   int64_t blocks = (nruns + 255) / 256;
   const int64_t cap = (int64_t)ctx->num_sms * 8;
   if (blocks > cap) blocks = cap;
   int ps = pb_prof_begin(ctx, PB_K_UTIL);
   larnv_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(x0, count, col_len, ld, X);
   pb_prof_end(ctx, ps, 8.0 * (double)count);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   const unsigned long long xn = (x0 * larnv_pow((unsigned long long)count)) & LARNV_MASK;
   iseed[0] = (long long)(xn >> 36), iseed[1] = (long long)((xn >> 24) & 4095), iseed[2] = (long long)((xn >> 12) & 4095),
   iseed[3] = (long long)(xn & 4095);
   return 0;
}

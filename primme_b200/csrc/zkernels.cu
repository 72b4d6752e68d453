// zkernels.cu -- complex fp64 twins of K2-K6 behind zprimme / cublas_zprimme (config C3).
//
// The reference gets its complex solver by re-including every source with SCALAR = complex
// (reference src/include/template_types.h:51-204) over zgemm/zgemv/zdotc calls of a vendor BLAS
// (src/linalg/cublas_wrapper.c).  Here the complex instantiation of the hot path is three kernel
// families on interleaved (re,im) data, HBM-bound like their real counterparts (a complex FMA is 4 real
// FMAs on 16 bytes: 0.25-2 flop/B at these panel widths, far below the fp64 ridge, so the plain DFMA
// pipe is enough and no DMMA staging is needed):
//   zsweep_kernel     X <- (X - [Q V] C) Y and P = [Q V X]^H X in one row sweep   (ortho.c:963-1072;
//                     update_projection.c:99-102; the CGS gemv pair of Bortho_gen ortho.c:237-291, which
//                     is what C3's orth = implicit_I path runs)
//   ztall_kernel      P = V h, Q = W h for up to 16 columns of h per pass         (auxiliary_eigs_normal.c:254,271)
//   utilities         column copy / permute / axpy / scale / dots / residual / Jacobi: the per-column
//                     scalars travel as kernel parameters (no staging copy, no stream synchronisation
//                     per call) and reductions finish inside the kernel (PbFin): one launch per operation
//                     of the inner QMR solver.
// Panels are reduced in a fixed order (bitwise reproducible) and all-reduced over peer memory / NCCL
// exactly like the real ones (a complex panel is a real panel of twice the length).
#include "pb200_internal.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

typedef double2 zc;
__device__ __forceinline__ zc zmake(double a, double b) { return make_double2(a, b); }
// acc += a * b
__device__ __forceinline__ void zfma(zc &acc, const zc a, const zc b) {
   acc.x = fma(a.x, b.x, acc.x);
   acc.x = fma(-a.y, b.y, acc.x);
   acc.y = fma(a.x, b.y, acc.y);
   acc.y = fma(a.y, b.x, acc.y);
}
// acc -= a * b
__device__ __forceinline__ void zfms(zc &acc, const zc a, const zc b) {
   acc.x = fma(-a.x, b.x, acc.x);
   acc.x = fma(a.y, b.y, acc.x);
   acc.y = fma(-a.x, b.y, acc.y);
   acc.y = fma(-a.y, b.x, acc.y);
}
// acc += conj(a) * b
__device__ __forceinline__ void zfmac(zc &acc, const zc a, const zc b) {
   acc.x = fma(a.x, b.x, acc.x);
   acc.x = fma(a.y, b.y, acc.x);
   acc.y = fma(a.x, b.y, acc.y);
   acc.y = fma(-a.y, b.x, acc.y);
}

constexpr int ZT = 256;  // threads per CTA == rows per tile
constexpr int ZW = ZT / 32;

inline int grid_for(pb200_ctx *ctx, int64_t n, int per_sm) {
   int64_t g = (n + ZT - 1) / ZT;
   const int64_t cap = (int64_t)ctx->num_sms * per_sm;
   if (g > cap) g = cap;
   if (g < 1) g = 1;
   return (int)g;
}

// partial-panel plumbing shared by every reducing kernel: in-kernel finish when the launch shape allows
// it (also the multi-rank peer exchange), else one partial per CTA for pb_finish_panel
struct ZPanel {
   PbFin fin;
   double *partials;
   int grid, cnt;
};
int zpanel_setup(pb200_ctx *ctx, int grid, int cnt, ZPanel *p) {
   p->grid = grid, p->cnt = cnt;
   const int r = pb_fin_prepare(ctx, grid, 1, cnt, &p->fin);
   if (r < 0) return r;
   if (r == 1) {
      memset(&p->fin, 0, sizeof(p->fin));
      PB_CHK(pb_ensure_partials(ctx, (size_t)grid * cnt));
      p->partials = ctx->d_partials;
   } else
      p->partials = p->fin.partials;
   return 0;
}
// result in ctx->h_pinned[0..cnt)
int zpanel_collect(pb200_ctx *ctx, const ZPanel *p) {
   if (p->fin.cnt > 0) return pb_collect_panel(ctx, &p->fin);
   return pb_finish_panel(ctx, p->grid, p->cnt);
}

// --------------------------------------------------------------------------- ortho sweep ----
struct ZSweepArgs {
   const zc *Q, *V;
   zc *X;
   int64_t n, ldq, ldv, ldx;
   int q, mv, b;
   int do_update, has_Y, do_gram, xx;
   const zc *coef;  // device: C (k x BT, row stride BT) then Y (BT x BT)
   int coef_inline; // 1: the block travels in the kernel parameter space instead (ZCoef)
   ZPanel pan;
};
// coefficient block of the update by value: no staging copy and no stream synchronisation in front of the
// launch (the CGS passes of C3 and the projectors of the inner solver update with <= 60 x 1 blocks)
#define ZCOEF_MAX 448
struct ZCoef {
   zc v[ZCOEF_MAX];
};
__device__ __forceinline__ const zc *zcol(const ZSweepArgs &a, int j) {
   return j < a.q ? a.Q + (size_t)j * a.ldq : a.V + (size_t)(j - a.q) * a.ldv;
}

// BT: block columns (padded), CPW: basis columns per warp in the Gram phase (k <= ZW * CPW)
template <int BT, int CPW>
__global__ void __launch_bounds__(ZT) zsweep_kernel(const ZSweepArgs a, const __grid_constant__ ZCoef cf) {
   extern __shared__ __align__(16) unsigned char zsm[];
   const int k = a.q + a.mv;
   zc *Cs = reinterpret_cast<zc *>(zsm);   // k * BT
   zc *Ys = Cs + (size_t)k * BT;           // BT * BT   ([r][c])
   zc *xs = Ys + BT * BT;                  // BT * ZT
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   if (a.do_update) {
      const zc *src = a.coef_inline ? cf.v : a.coef;
      for (int i = tid; i < k * BT + BT * BT; i += ZT) Cs[i] = src[i];
   }
   __syncthreads();

   zc acc[CPW][BT], accx[BT];
#pragma unroll
   for (int j = 0; j < CPW; j++)
#pragma unroll
      for (int c = 0; c < BT; c++) acc[j][c] = zmake(0.0, 0.0);
#pragma unroll
   for (int c = 0; c < BT; c++) accx[c] = zmake(0.0, 0.0);
   const zc *cols[CPW];
#pragma unroll
   for (int j = 0; j < CPW; j++) {
      const int jj = warp * CPW + j;
      cols[j] = (a.do_gram && jj < k) ? zcol(a, jj) : nullptr;
   }

   const int64_t ntiles = (a.n + ZT - 1) / ZT;
   for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t r0 = tile * ZT;
      {
         const int64_t r = r0 + tid;
         zc x[BT];
#pragma unroll
         for (int c = 0; c < BT; c++) x[c] = (r < a.n && c < a.b) ? a.X[r + (size_t)c * a.ldx] : zmake(0.0, 0.0);
         if (a.do_update && r < a.n) {
            int j = 0;
            for (; j + 2 <= k; j += 2) {
               const zc v0 = zcol(a, j)[r], v1 = zcol(a, j + 1)[r];
#pragma unroll
               for (int c = 0; c < BT; c++) {
                  zfms(x[c], v0, Cs[(j + 0) * BT + c]);
                  zfms(x[c], v1, Cs[(j + 1) * BT + c]);
               }
            }
            for (; j < k; j++) {
               const zc v0 = zcol(a, j)[r];
#pragma unroll
               for (int c = 0; c < BT; c++) zfms(x[c], v0, Cs[j * BT + c]);
            }
            if (a.has_Y) {
               zc y[BT];
#pragma unroll
               for (int c = 0; c < BT; c++) {
                  zc s = zmake(0.0, 0.0);
#pragma unroll
                  for (int cc = 0; cc < BT; cc++) zfma(s, x[cc], Ys[cc * BT + c]);
                  y[c] = s;
               }
#pragma unroll
               for (int c = 0; c < BT; c++) x[c] = y[c];
            }
#pragma unroll
            for (int c = 0; c < BT; c++)
               if (c < a.b) a.X[r + (size_t)c * a.ldx] = x[c];
         }
         if (a.do_gram) {
#pragma unroll
            for (int c = 0; c < BT; c++) xs[c * ZT + tid] = x[c];
         }
      }
      if (a.do_gram) {
         __syncthreads();
#pragma unroll 2
         for (int i = 0; i < ZT / 32; i++) {
            const int row = i * 32 + lane;
            const int64_t r = r0 + row;
            zc xv[BT];
#pragma unroll
            for (int c = 0; c < BT; c++) xv[c] = xs[c * ZT + row];
            zc av[CPW];
#pragma unroll
            for (int j = 0; j < CPW; j++) av[j] = (cols[j] && r < a.n) ? cols[j][r] : zmake(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < CPW; j++)
#pragma unroll
               for (int c = 0; c < BT; c++) zfmac(acc[j][c], av[j], xv[c]);
            if (a.xx && warp < a.b) {
               const zc xw = xs[warp * ZT + row];
#pragma unroll
               for (int c = 0; c < BT; c++) zfmac(accx[c], xw, xv[c]);
            }
         }
         __syncthreads();
      }
   }
   if (!a.do_gram) return;
   // one partial panel per CTA: rows x b complex = 2 * rows * b doubles, column-major
   const int rows = k + (a.xx ? a.b : 0);
   zc *out = reinterpret_cast<zc *>(a.pan.partials) + (size_t)blockIdx.x * rows * a.b;
#pragma unroll
   for (int j = 0; j < CPW; j++)
#pragma unroll
      for (int c = 0; c < BT; c++) {
         zc v = acc[j][c];
#pragma unroll
         for (int o = 16; o > 0; o >>= 1) {
            v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
            v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
         }
         const int jj = warp * CPW + j;
         if (lane == 0 && jj < k && c < a.b) out[jj + (size_t)c * rows] = v;
      }
   if (a.xx) {
#pragma unroll
      for (int c = 0; c < BT; c++) {
         zc v = accx[c];
#pragma unroll
         for (int o = 16; o > 0; o >>= 1) {
            v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
            v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
         }
         if (lane == 0 && warp < a.b && c < a.b) out[k + warp + (size_t)c * rows] = v;
      }
   }
   pb_finish_device(a.pan.fin, tid, ZT, 15, reinterpret_cast<int *>(zsm));
}

template <int BT, int CPW>
int launch_zsweep(pb200_ctx *ctx, const ZSweepArgs &a, const ZCoef &cf, int grid, size_t shmem) {
   auto kern = zsweep_kernel<BT, CPW>;
   static size_t attr = 0;
   if (shmem > 48 * 1024 && shmem > attr) {
      PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
      attr = shmem;
   }
   kern<<<grid, ZT, shmem, ctx->stream>>>(a, cf);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return 0;
}
// basis columns one launch can take in its Gram phase
inline int zsweep_kmax(int BT) { return ZW * (BT == 1 ? 12 : BT == 2 ? 8 : BT == 4 ? 4 : 2); }

int dispatch_zsweep(pb200_ctx *ctx, const ZSweepArgs &a, const ZCoef &cf, int BT, int grid, size_t shmem) {
   const int k = a.q + a.mv;
   const int cpw = a.do_gram ? (k + ZW - 1) / ZW : 1;
#define ZS(B, C) return launch_zsweep<B, C>(ctx, a, cf, grid, shmem)
   switch (BT) {
   case 1:
      if (cpw <= 1) ZS(1, 1);
      if (cpw <= 2) ZS(1, 2);
      if (cpw <= 4) ZS(1, 4);
      if (cpw <= 6) ZS(1, 6);
      if (cpw <= 8) ZS(1, 8);
      ZS(1, 12);
   case 2:
      if (cpw <= 1) ZS(2, 1);
      if (cpw <= 2) ZS(2, 2);
      if (cpw <= 4) ZS(2, 4);
      ZS(2, 8);
   case 4:
      if (cpw <= 1) ZS(4, 1);
      if (cpw <= 2) ZS(4, 2);
      ZS(4, 4);
   default:
      if (cpw <= 1) ZS(8, 1);
      ZS(8, 2);
   }
#undef ZS
}

// one launch: update with ALL the given columns and/or Gram against them (k <= kmax when do_gram)
int zsweep_once(pb200_ctx *ctx, int64_t n, const zc *Q, int q, int64_t ldq, const zc *V, int mv, int64_t ldv, zc *X,
      int b, int64_t ldx, const zc *C_host, int ldc, const zc *Y_host, int ldy, int xx, zc *P_host, int ldp) {
   const int k = q + mv;
   const int BT = b <= 1 ? 1 : b <= 2 ? 2 : b <= 4 ? 4 : 8;
   ZSweepArgs a;
   memset(&a, 0, sizeof(a));
   a.Q = Q, a.V = V, a.X = X, a.n = n, a.ldq = ldq, a.ldv = ldv, a.ldx = ldx, a.q = q, a.mv = mv, a.b = b;
   a.do_update = C_host != NULL || Y_host != NULL;
   a.has_Y = Y_host != NULL;
   a.do_gram = P_host != NULL;
   a.xx = xx ? 1 : 0;
   if (!a.do_update && !a.do_gram) return 0;
   ZCoef cfv;  // host staging of the by-value block
   if (a.do_update) {
      const size_t ne = (size_t)k * BT + BT * BT;
      zc *hp;
      if (ne <= ZCOEF_MAX) {
         a.coef_inline = 1;
         hp = cfv.v;
      } else {
         // large blocks go through the pinned staging buffer (complex: 2 doubles per entry)
         PB_CHK(pb_ensure_small(ctx, 2 * ne));
         PB_CUDA(cudaStreamSynchronize(ctx->stream));  // the staging buffer may still feed an earlier copy
         hp = reinterpret_cast<zc *>(ctx->h_pinned);
      }
      for (int j = 0; j < k; j++)
         for (int c = 0; c < BT; c++)
            hp[(size_t)j * BT + c] = (c < b && C_host) ? C_host[j + (size_t)c * ldc] : zc{0.0, 0.0};
      zc *hy = hp + (size_t)k * BT;
      for (int r = 0; r < BT; r++)
         for (int c = 0; c < BT; c++)
            hy[r * BT + c] = Y_host ? ((r < b && c < b) ? Y_host[r + (size_t)c * ldy] : zc{0.0, 0.0})
                                    : zc{r == c ? 1.0 : 0.0, 0.0};
      if (!a.coef_inline) {
         PB_CUDA(cudaMemcpyAsync(ctx->d_small, hp, 2 * ne * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
         a.coef = reinterpret_cast<const zc *>(ctx->d_small);
      }
   }
   const int rows = k + (a.xx ? b : 0);
   const int cnt = a.do_gram ? 2 * rows * b : 0;
   const int grid = grid_for(ctx, n, 3);
   if (a.do_gram) PB_CHK(zpanel_setup(ctx, grid, cnt, &a.pan));
   const size_t shmem = ((size_t)k * BT + BT * BT + (size_t)BT * ZT) * sizeof(zc) + 16;
   // algorithmic bytes: basis read once, X read (+ written when updated), 16 bytes per element
   const double abytes = 16.0 * (double)n * (k + b * (a.do_update ? 2 : 1));
   int ps = pb_prof_begin(ctx, PB_K_ORTHO);
   int rc = dispatch_zsweep(ctx, a, cfv, BT, grid, shmem);
   pb_prof_end(ctx, ps, abytes);
   PB_CHK(rc);
   if (a.do_gram) {
      PB_CHK(zpanel_collect(ctx, &a.pan));
      const zc *hp = reinterpret_cast<const zc *>(ctx->h_pinned);
      for (int c = 0; c < b; c++)
         for (int j = 0; j < rows; j++) P_host[j + (size_t)c * ldp] = hp[j + (size_t)c * rows];
   }
   return 0;
}

// ------------------------------------------------------------------ tall-skinny products ----
struct ZTallArgs {
   const zc *V, *W;
   int64_t n, ld, ldo;
   int m, nh, need_y;
   const zc *hdev;  // m x NT, row stride NT (zero padded)
   zc *P, *Qo;      // n x nh outputs, leading dimension ldo
};
template <int NT>
__global__ void __launch_bounds__(ZT) ztall_kernel(const ZTallArgs a) {
   extern __shared__ __align__(16) unsigned char zsm[];
   zc *hs = reinterpret_cast<zc *>(zsm);
   for (int i = threadIdx.x; i < a.m * NT; i += ZT) hs[i] = a.hdev[i];
   __syncthreads();
   for (int64_t r = (int64_t)blockIdx.x * ZT + threadIdx.x; r < a.n; r += (int64_t)gridDim.x * ZT) {
      zc x[NT], y[NT];
#pragma unroll
      for (int c = 0; c < NT; c++) x[c] = y[c] = zmake(0.0, 0.0);
      const zc *vp = a.V + r, *wp = a.W + r;
      int k = 0;
      for (; k + 2 <= a.m; k += 2) {
         const zc v0 = vp[(size_t)k * a.ld], v1 = vp[(size_t)(k + 1) * a.ld];
         zc w0 = zmake(0.0, 0.0), w1 = w0;
         if (a.need_y) w0 = wp[(size_t)k * a.ld], w1 = wp[(size_t)(k + 1) * a.ld];
         const zc *h0 = hs + (size_t)k * NT;
#pragma unroll
         for (int c = 0; c < NT; c++) {
            zfma(x[c], v0, h0[c]);
            zfma(x[c], v1, h0[NT + c]);
            zfma(y[c], w0, h0[c]);
            zfma(y[c], w1, h0[NT + c]);
         }
      }
      for (; k < a.m; k++) {
         const zc v0 = vp[(size_t)k * a.ld];
         const zc w0 = a.need_y ? wp[(size_t)k * a.ld] : zmake(0.0, 0.0);
         const zc *h0 = hs + (size_t)k * NT;
#pragma unroll
         for (int c = 0; c < NT; c++) {
            zfma(x[c], v0, h0[c]);
            zfma(y[c], w0, h0[c]);
         }
      }
#pragma unroll
      for (int c = 0; c < NT; c++)
         if (c < a.nh) {
            a.P[r + (size_t)c * a.ldo] = x[c];
            if (a.need_y) a.Qo[r + (size_t)c * a.ldo] = y[c];
         }
   }
}

// ---------------------------------------------------------------------------- utilities ----
struct ZScal {
   zc v[8];
};
struct ZIdx {
   int x[32], y[32];
};
__global__ void __launch_bounds__(ZT) zcopy_cols_kernel(int64_t n, const zc *__restrict__ X, int64_t ldx, zc *__restrict__ Y,
      int64_t ldy, int ncols, const ZIdx idx) {
   for (int64_t r = (int64_t)blockIdx.x * ZT + threadIdx.x; r < n; r += (int64_t)gridDim.x * ZT)
      for (int j = 0; j < ncols; j++) Y[r + (size_t)idx.y[j] * ldy] = X[r + (size_t)idx.x[j] * ldx];
}
__global__ void __launch_bounds__(ZT) zaxpy_kernel(int64_t n, const ZScal al, const zc *__restrict__ X, int64_t ldx,
      zc *__restrict__ Y, int64_t ldy, int ncols) {
   for (int64_t r = (int64_t)blockIdx.x * ZT + threadIdx.x; r < n; r += (int64_t)gridDim.x * ZT)
      for (int j = 0; j < ncols; j++) {
         zc y = Y[r + (size_t)j * ldy];
         zfma(y, al.v[j], X[r + (size_t)j * ldx]);
         Y[r + (size_t)j * ldy] = y;
      }
}
__global__ void __launch_bounds__(ZT) zscale_kernel(int64_t n, const ZScal al, zc *__restrict__ X, int64_t ldx, int ncols) {
   for (int64_t r = (int64_t)blockIdx.x * ZT + threadIdx.x; r < n; r += (int64_t)gridDim.x * ZT)
      for (int j = 0; j < ncols; j++) {
         zc s = zmake(0.0, 0.0);
         zfma(s, al.v[j], X[r + (size_t)j * ldx]);
         X[r + (size_t)j * ldx] = s;
      }
}
// mode 0: out[j] = X_j^H Y_j (complex, 2 doubles per column); mode 1: Y_j -= theta_j X_j, out[j] = |Y_j|^2
// (the same 2-double slot, imaginary part 0)
__global__ void __launch_bounds__(ZT) zdots_kernel(int64_t n, const zc *__restrict__ X, int64_t ldx, zc *__restrict__ Y,
      int64_t ldy, int ncols, int mode, const ZScal th, const ZPanel pan) {
   __shared__ double red[ZW][16];
   __shared__ int flag;
   zc acc[8];
#pragma unroll
   for (int j = 0; j < 8; j++) acc[j] = zmake(0.0, 0.0);
   for (int64_t r = (int64_t)blockIdx.x * ZT + threadIdx.x; r < n; r += (int64_t)gridDim.x * ZT) {
#pragma unroll
      for (int j = 0; j < 8; j++)
         if (j < ncols) {
            const zc x = X[r + (size_t)j * ldx];
            zc y = Y[r + (size_t)j * ldy];
            if (mode == 1) {
               y.x = fma(-th.v[j].x, x.x, y.x);
               y.y = fma(-th.v[j].x, x.y, y.y);
               Y[r + (size_t)j * ldy] = y;
               acc[j].x = fma(y.x, y.x, acc[j].x);
               acc[j].x = fma(y.y, y.y, acc[j].x);
            } else
               zfmac(acc[j], x, y);
         }
   }
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
   for (int j = 0; j < 8; j++) {
      zc v = acc[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
         v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
         v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
      }
      if (lane == 0) red[warp][2 * j] = v.x, red[warp][2 * j + 1] = v.y;
   }
   __syncthreads();
   if (threadIdx.x < 2 * ncols) {
      double s = 0.0;
      for (int w = 0; w < ZW; w++) s += red[w][threadIdx.x];
      pan.partials[(size_t)blockIdx.x * 2 * ncols + threadIdx.x] = s;
   }
   pb_finish_device(pan.fin, threadIdx.x, ZT, 15, &flag);
}
// complex vectors, REAL recurrence scalars (reference inner_solve.c:155-182): delta = gamma delta + eta d;
// sol += delta; partial |sol|^2
__global__ void __launch_bounds__(ZT) zqmr_update_kernel(int64_t n, const ZScal gam, const ZScal eta, const zc *__restrict__ D,
      int64_t ldd, zc *__restrict__ Dl, int64_t ldl, zc *__restrict__ S, int64_t lds, int ncols, int want_dots,
      const ZPanel pan) {
   __shared__ double red[ZW][8];
   __shared__ int flag;
   double acc[8];
#pragma unroll
   for (int j = 0; j < 8; j++) acc[j] = 0.0;
   for (int64_t r = (int64_t)blockIdx.x * ZT + threadIdx.x; r < n; r += (int64_t)gridDim.x * ZT) {
#pragma unroll
      for (int j = 0; j < 8; j++)
         if (j < ncols) {
            // the same roundings as zscale (complex product with (gamma, 0)) + zaxpy + zaxpy + dot
            zc t = zmake(0.0, 0.0);
            zfma(t, gam.v[j], Dl[r + (size_t)j * ldl]);
            zfma(t, eta.v[j], D[r + (size_t)j * ldd]);
            Dl[r + (size_t)j * ldl] = t;
            zc s = S[r + (size_t)j * lds];
            zfma(s, zmake(1.0, 0.0), t);
            S[r + (size_t)j * lds] = s;
            acc[j] = fma(s.x, s.x, acc[j]);
            acc[j] = fma(s.y, s.y, acc[j]);
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
      for (int w = 0; w < ZW; w++) s += red[w][threadIdx.x];
      pan.partials[(size_t)blockIdx.x * ncols + threadIdx.x] = s;
   }
   pb_finish_device(pan.fin, threadIdx.x, ZT, 15, &flag);
}

__global__ void __launch_bounds__(ZT) zjacobi_kernel(int64_t n, const double *__restrict__ diag, const ZScal sh, double minabs,
      const zc *__restrict__ X, int64_t ldx, zc *__restrict__ Y, int64_t ldy, int ncols) {
   for (int64_t r = (int64_t)blockIdx.x * ZT + threadIdx.x; r < n; r += (int64_t)gridDim.x * ZT) {
      const double d0 = diag[r];
      for (int j = 0; j < ncols; j++) {
         double d = d0 - sh.v[j].x;
         if (fabs(d) < minabs) d = d < 0 ? -minabs : minabs;
         const zc x = X[r + (size_t)j * ldx];
         Y[r + (size_t)j * ldy] = zmake(x.x / d, x.y / d);
      }
   }
}

int zdots_impl(pb200_ctx *ctx, int64_t n, const zc *X, int64_t ldx, zc *Y, int64_t ldy, int ncols, int mode,
      const double *theta_host, double *out_host /* 2 per column (mode 0) or 1 per column (mode 1) */) {
   for (int c0 = 0; c0 < ncols; c0 += 8) {
      const int nc = ncols - c0 < 8 ? ncols - c0 : 8;
      if (n <= 0) {
         // a rank without rows still takes part in the reduction of the others
         if (ctx->nranks > 1) {
            PB_CHK(pb_ensure_small(ctx, (size_t)2 * nc));
            const int zr = pb_fin_contribute_zeros(ctx, 2 * nc);
            if (zr < 0) return zr;
            if (zr == 1) {
               PB_CUDA(cudaMemsetAsync(ctx->d_panel, 0, sizeof(double) * 2 * nc, ctx->stream));
               PB_CHK(pb_nccl_allreduce_dev(ctx, ctx->d_panel, 2 * nc));
               PB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_panel, sizeof(double) * 2 * nc, cudaMemcpyDeviceToHost, ctx->stream));
               PB_CUDA(cudaStreamSynchronize(ctx->stream));
            }
         } else
            for (int j = 0; j < 2 * nc; j++) ctx->h_pinned[j] = 0.0;
      } else {
         ZScal th;
         memset(&th, 0, sizeof(th));
         if (mode == 1)
            for (int j = 0; j < nc; j++) th.v[j].x = theta_host[c0 + j];
         const int grid = grid_for(ctx, n, 4);
         ZPanel pan;
         PB_CHK(zpanel_setup(ctx, grid, 2 * nc, &pan));
         int ps = pb_prof_begin(ctx, PB_K_UTIL);
         zdots_kernel<<<grid, ZT, 0, ctx->stream>>>(n, X + (size_t)c0 * ldx, ldx, Y + (size_t)c0 * ldy, ldy, nc, mode, th, pan);
         pb_prof_end(ctx, ps, 16.0 * (double)n * nc * (mode == 1 ? 3 : 2));
         ctx->launches++;
         PB_CUDA(cudaGetLastError());
         PB_CHK(zpanel_collect(ctx, &pan));
      }
      for (int j = 0; j < nc; j++) {
         if (mode == 1) out_host[c0 + j] = ctx->h_pinned[2 * j];
         else out_host[2 * (c0 + j)] = ctx->h_pinned[2 * j], out_host[2 * (c0 + j) + 1] = ctx->h_pinned[2 * j + 1];
      }
   }
   return 0;
}

}  // namespace

// ================================================================================ C-ABI ====
extern "C" int pb200_zortho_sweep(pb200_ctx *ctx, int64_t n, const void *Q_, int q, int64_t ldq, const void *V_, int mv,
      int64_t ldv, void *X_, int b, int64_t ldx, const void *C_host_, int ldc, const void *Y_host_, int ldy, int xx,
      void *P_host_, int ldp) {
   if (b <= 0) return 0;
   if (b > 8 || q < 0 || mv < 0) return PB200_ERR_ARG;
   const zc *Q = (const zc *)Q_, *V = (const zc *)V_, *C_host = (const zc *)C_host_, *Y_host = (const zc *)Y_host_;
   zc *X = (zc *)X_, *P_host = (zc *)P_host_;
   const int k = q + mv;
   if (n <= 0) {
      if (P_host) {
         const int rows = k + (xx ? b : 0);
         for (int c = 0; c < b; c++)
            for (int j = 0; j < rows; j++) P_host[j + (size_t)c * ldp] = zc{0.0, 0.0};
         if (ctx->nranks > 1) {
            const int cnt = 2 * rows * b;
            PB_CHK(pb_ensure_small(ctx, (size_t)cnt));
            const int zr = pb_fin_contribute_zeros(ctx, cnt);
            if (zr < 0) return zr;
            if (zr == 1) {
               PB_CUDA(cudaMemsetAsync(ctx->d_panel, 0, sizeof(double) * cnt, ctx->stream));
               PB_CHK(pb_nccl_allreduce_dev(ctx, ctx->d_panel, cnt));
               PB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_panel, sizeof(double) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
               PB_CUDA(cudaStreamSynchronize(ctx->stream));
            }
            const zc *hp = reinterpret_cast<const zc *>(ctx->h_pinned);
            for (int c = 0; c < b; c++)
               for (int j = 0; j < rows; j++) P_host[j + (size_t)c * ldp] = hp[j + (size_t)c * rows];
         }
      }
      return 0;
   }
   const int BT = b <= 1 ? 1 : b <= 2 ? 2 : b <= 4 ? 4 : 8;
   const int kmax = zsweep_kmax(BT);
   const int kupd = 96;  // columns per update launch (coefficient block in shared memory)
   const bool upd = C_host != NULL || Y_host != NULL;
   if (k <= kmax && k <= kupd)
      return zsweep_once(ctx, n, Q, q, ldq, V, mv, ldv, X, b, ldx, C_host, ldc, Y_host, ldy, xx, P_host, ldp);
   // more columns than one launch covers: the update chunk by chunk (Y with the last chunk), then the
   // Gram panel chunk by chunk (the X block only with the last chunk)
   auto chunk = [&](int j0, int j1, const zc **Qc, int *qa, const zc **Vc, int *va) {
      *qa = j0 < q ? (j1 < q ? j1 : q) - j0 : 0;
      const int va0 = j0 > q ? j0 - q : 0;
      *va = j1 > q ? (j1 - q) - va0 : 0;
      *Qc = *qa > 0 ? Q + (size_t)j0 * ldq : Q;
      *Vc = V ? V + (size_t)va0 * ldv : V;
   };
   if (upd) {
      if (k == 0)
         PB_CHK(zsweep_once(ctx, n, NULL, 0, 0, NULL, 0, 0, X, b, ldx, NULL, 0, Y_host, ldy, 0, NULL, 0));
      for (int j0 = 0; j0 < k; j0 += kupd) {
         const int j1 = j0 + kupd < k ? j0 + kupd : k;
         const zc *Qc, *Vc;
         int qa, va;
         chunk(j0, j1, &Qc, &qa, &Vc, &va);
         const bool last = j1 == k;
         if (!C_host && !last) continue;
         PB_CHK(zsweep_once(ctx, n, Qc, qa, ldq, Vc, va, ldv, X, b, ldx, C_host ? C_host + j0 : NULL, ldc,
               last ? Y_host : NULL, ldy, 0, NULL, 0));
      }
   }
   if (P_host) {
      if (k == 0) PB_CHK(zsweep_once(ctx, n, NULL, 0, 0, NULL, 0, 0, X, b, ldx, NULL, 0, NULL, 0, xx, P_host, ldp));
      for (int j0 = 0; j0 < k; j0 += kmax) {
         const int j1 = j0 + kmax < k ? j0 + kmax : k;
         const zc *Qc, *Vc;
         int qa, va;
         chunk(j0, j1, &Qc, &qa, &Vc, &va);
         PB_CHK(zsweep_once(ctx, n, Qc, qa, ldq, Vc, va, ldv, X, b, ldx, NULL, 0, NULL, 0, j1 == k ? xx : 0,
               P_host + j0, ldp));
      }
   }
   return 0;
}

extern "C" int pb200_zcopy_columns(pb200_ctx *ctx, int64_t n, const void *X, int64_t ldx, const int *xin_host, void *Y,
      int64_t ldy, const int *yin_host, int ncols) {
   if (ncols <= 0 || n <= 0) return 0;
   for (int c0 = 0; c0 < ncols; c0 += 32) {
      const int nc = ncols - c0 < 32 ? ncols - c0 : 32;
      ZIdx idx;
      for (int j = 0; j < nc; j++) idx.x[j] = xin_host ? xin_host[c0 + j] : c0 + j, idx.y[j] = yin_host ? yin_host[c0 + j] : c0 + j;
      zcopy_cols_kernel<<<grid_for(ctx, n, 8), ZT, 0, ctx->stream>>>(n, (const zc *)X, ldx, (zc *)Y, ldy, nc, idx);
      ctx->launches++;
      PB_CUDA(cudaGetLastError());
   }
   return 0;
}

extern "C" int pb200_zpermute_columns(pb200_ctx *ctx, int64_t n, void *X, int64_t ldx, const int *perm_host, int ncols) {
   if (ncols <= 0 || n <= 0) return 0;
   std::vector<int> moved, src;
   for (int i = 0; i < ncols; i++)
      if (perm_host[i] != i) moved.push_back(i), src.push_back(perm_host[i]);
   const int nm = (int)moved.size();
   if (nm == 0) return 0;
   // through scratch: tmp(:,i) = X(:,perm[moved i]) for the moved columns, then back
   PB_CHK(pb_ensure_scratch(ctx, sizeof(zc) * (size_t)n * nm));
   PB_CHK(pb200_zcopy_columns(ctx, n, X, ldx, src.data(), ctx->d_scratch, n, NULL, nm));
   PB_CHK(pb200_zcopy_columns(ctx, n, ctx->d_scratch, n, NULL, X, ldx, moved.data(), nm));
   return 0;
}

extern "C" int pb200_zaxpy_columns(pb200_ctx *ctx, int64_t n, const void *alpha_host, const void *X, int64_t ldx, void *Y,
      int64_t ldy, int ncols) {
   if (ncols <= 0 || n <= 0) return 0;
   for (int c0 = 0; c0 < ncols; c0 += 8) {
      const int nc = ncols - c0 < 8 ? ncols - c0 : 8;
      ZScal al;
      memset(&al, 0, sizeof(al));
      memcpy(al.v, (const zc *)alpha_host + c0, sizeof(zc) * nc);
      int ps = pb_prof_begin(ctx, PB_K_UTIL);
      zaxpy_kernel<<<grid_for(ctx, n, 8), ZT, 0, ctx->stream>>>(n, al, (const zc *)X + (size_t)c0 * ldx, ldx,
            (zc *)Y + (size_t)c0 * ldy, ldy, nc);
      pb_prof_end(ctx, ps, 48.0 * (double)n * nc);
      ctx->launches++;
      PB_CUDA(cudaGetLastError());
   }
   return 0;
}

extern "C" int pb200_zscale_columns(pb200_ctx *ctx, int64_t n, const void *alpha_host, void *X, int64_t ldx, int ncols) {
   if (ncols <= 0 || n <= 0) return 0;
   for (int c0 = 0; c0 < ncols; c0 += 8) {
      const int nc = ncols - c0 < 8 ? ncols - c0 : 8;
      ZScal al;
      memset(&al, 0, sizeof(al));
      memcpy(al.v, (const zc *)alpha_host + c0, sizeof(zc) * nc);
      int ps = pb_prof_begin(ctx, PB_K_UTIL);
      zscale_kernel<<<grid_for(ctx, n, 8), ZT, 0, ctx->stream>>>(n, al, (zc *)X + (size_t)c0 * ldx, ldx, nc);
      pb_prof_end(ctx, ps, 32.0 * (double)n * nc);
      ctx->launches++;
      PB_CUDA(cudaGetLastError());
   }
   return 0;
}

extern "C" int pb200_zcolumn_dots(pb200_ctx *ctx, int64_t n, const void *X, int64_t ldx, const void *Y, int64_t ldy,
      int ncols, void *out_host) {
   if (ncols <= 0) return 0;
   return zdots_impl(ctx, n, (const zc *)X, ldx, (zc *)Y, ldy, ncols, 0, NULL, (double *)out_host);
}

extern "C" int pb200_zresidual_inplace(pb200_ctx *ctx, int64_t n, const double *theta_host, const void *V, int64_t ldv,
      void *W, int64_t ldw, int ncols, double *out_host) {
   if (ncols <= 0) return 0;
   return zdots_impl(ctx, n, (const zc *)V, ldv, (zc *)W, ldw, ncols, 1, theta_host, out_host);
}

extern "C" int pb200_zjacobi(pb200_ctx *ctx, int64_t n, const double *diag, const double *shifts_host, double minabs,
      const void *X, int64_t ldx, void *Y, int64_t ldy, int ncols) {
   if (ncols <= 0 || n <= 0) return 0;
   for (int c0 = 0; c0 < ncols; c0 += 8) {
      const int nc = ncols - c0 < 8 ? ncols - c0 : 8;
      ZScal sh;
      memset(&sh, 0, sizeof(sh));
      if (shifts_host)
         for (int j = 0; j < nc; j++) sh.v[j].x = shifts_host[c0 + j];
      int ps = pb_prof_begin(ctx, PB_K_UTIL);
      zjacobi_kernel<<<grid_for(ctx, n, 8), ZT, 0, ctx->stream>>>(n, diag, sh, minabs, (const zc *)X + (size_t)c0 * ldx, ldx,
            (zc *)Y + (size_t)c0 * ldy, ldy, nc);
      pb_prof_end(ctx, ps, 32.0 * (double)n * nc + 8.0 * (double)n);
      ctx->launches++;
      PB_CUDA(cudaGetLastError());
   }
   return 0;
}

extern "C" int pb200_zqmr_update(pb200_ctx *ctx, int64_t n, const double *gamma_host, const double *eta_host, const void *D,
      int64_t ldd, void *Delta, int64_t ldl, void *Sol, int64_t lds, int ncols, double *dots_host) {
   for (int c0 = 0; c0 < ncols; c0 += 8) {
      const int nc = ncols - c0 < 8 ? ncols - c0 : 8;
      if (n <= 0) {
         if (dots_host) {
            if (ctx->nranks > 1) {
               PB_CHK(pb_ensure_small(ctx, (size_t)nc));
               const int zr = pb_fin_contribute_zeros(ctx, nc);
               if (zr < 0) return zr;
               if (zr == 1) {
                  PB_CUDA(cudaMemsetAsync(ctx->d_panel, 0, sizeof(double) * nc, ctx->stream));
                  PB_CHK(pb_nccl_allreduce_dev(ctx, ctx->d_panel, nc));
                  PB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_panel, sizeof(double) * nc, cudaMemcpyDeviceToHost, ctx->stream));
                  PB_CUDA(cudaStreamSynchronize(ctx->stream));
               }
            } else
               for (int j = 0; j < nc; j++) ctx->h_pinned[j] = 0.0;
            for (int j = 0; j < nc; j++) dots_host[c0 + j] = ctx->h_pinned[j];
         }
         continue;
      }
      ZScal g, e;
      memset(&g, 0, sizeof(g)), memset(&e, 0, sizeof(e));
      for (int j = 0; j < nc; j++) g.v[j].x = gamma_host[c0 + j], e.v[j].x = eta_host[c0 + j];
      const int grid = grid_for(ctx, n, 4);
      ZPanel pan;
      memset(&pan, 0, sizeof(pan));
      if (dots_host) PB_CHK(zpanel_setup(ctx, grid, nc, &pan));
      int ps = pb_prof_begin(ctx, PB_K_UTIL);
      zqmr_update_kernel<<<grid, ZT, 0, ctx->stream>>>(n, g, e, (const zc *)D + (size_t)c0 * ldd, ldd, (zc *)Delta + (size_t)c0 * ldl,
            ldl, (zc *)Sol + (size_t)c0 * lds, lds, nc, dots_host != NULL, pan);
      pb_prof_end(ctx, ps, 80.0 * (double)n * nc);
      ctx->launches++;
      PB_CUDA(cudaGetLastError());
      if (dots_host) {
         PB_CHK(zpanel_collect(ctx, &pan));
         for (int j = 0; j < nc; j++) dots_host[c0 + j] = ctx->h_pinned[j];
      }
   }
   return 0;
}

extern "C" int pb200_zvwxr_can_fuse_gram(pb200_ctx *ctx, int64_t n, const void *V, const void *W, int m, int64_t ld, int nh,
      const pb200_vwxr_out *o) {
   (void)ctx, (void)n, (void)V, (void)W, (void)m, (void)ld, (void)nh, (void)o;
   return 0;
}

// K5 for complex data: the products P = V h and Q = W h go to scratch in passes of up to 16 columns of
// h (every row of V and W is read once per pass; candidates sweeps have nh <= 8: one pass), then every
// output is derived from P and Q with the other kernels -- Gram blocks by the sweep kernel, residuals in
// place, column ranges by copies.  Outputs may alias V / W: nothing is written to them before all
// products exist.
extern "C" int pb200_zvwxr(pb200_ctx *ctx, int64_t n, const void *V_, const void *W_, int m, int64_t ld, const void *h_host_,
      int ldh, int nh, const double *theta_host, const pb200_vwxr_out *o) {
   if (nh <= 0 || m < 0) return 0;
   if (o->P_host) return PB200_ERR_ARG;
   const zc *V = (const zc *)V_, *W = (const zc *)W_, *h_host = (const zc *)h_host_;
   const bool need_y = (o->Wo.ptr && o->Wo.ce > o->Wo.cb) || (o->R.ptr && o->R.ce > o->R.cb) ||
                       (o->rnorms_host && o->re > o->rb) || (o->H_host && o->nH > 0);
   const int64_t lds = n > 0 ? (n + 15) / 16 * 16 : 16;
   PB_CHK(pb_ensure_scratch(ctx, sizeof(zc) * (size_t)lds * nh * (need_y ? 2 : 1)));
   zc *P = (zc *)ctx->d_scratch, *Q = need_y ? P + (size_t)lds * nh : NULL;
   if (n > 0) {
      // passes of NT columns; the coefficient block of a pass must fit the default 48 KB of shared memory
      const int NT = (nh > 8 && (size_t)m * 16 * sizeof(zc) <= 48 * 1024) ? 16 : nh > 4 ? 8 : nh > 2 ? 4 : nh > 1 ? 2 : 1;
      for (int c0 = 0; c0 < nh; c0 += NT) {
         const int nc = nh - c0 < NT ? nh - c0 : NT;
         const size_t need = 2 * (size_t)m * NT;
         PB_CHK(pb_ensure_small(ctx, need));
         PB_CUDA(cudaStreamSynchronize(ctx->stream));
         zc *hp = reinterpret_cast<zc *>(ctx->h_pinned);
         for (int k = 0; k < m; k++)
            for (int c = 0; c < NT; c++) hp[(size_t)k * NT + c] = c < nc ? h_host[k + (size_t)(c0 + c) * ldh] : zc{0.0, 0.0};
         PB_CUDA(cudaMemcpyAsync(ctx->d_small, hp, need * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
         ZTallArgs a;
         a.V = V, a.W = W, a.n = n, a.ld = ld, a.ldo = lds, a.m = m, a.nh = nc, a.need_y = need_y ? 1 : 0;
         a.hdev = reinterpret_cast<const zc *>(ctx->d_small);
         a.P = P + (size_t)lds * c0, a.Qo = need_y ? Q + (size_t)lds * c0 : NULL;
         const size_t shmem = (size_t)m * NT * sizeof(zc) + 16;
         const int grid = grid_for(ctx, n, 4);
         int ps = pb_prof_begin(ctx, PB_K_VWXR);
         switch (NT) {
         case 1: ztall_kernel<1><<<grid, ZT, shmem, ctx->stream>>>(a); break;
         case 2: ztall_kernel<2><<<grid, ZT, shmem, ctx->stream>>>(a); break;
         case 4: ztall_kernel<4><<<grid, ZT, shmem, ctx->stream>>>(a); break;
         case 8: ztall_kernel<8><<<grid, ZT, shmem, ctx->stream>>>(a); break;
         default: ztall_kernel<16><<<grid, ZT, shmem, ctx->stream>>>(a); break;
         }
         pb_prof_end(ctx, ps, 16.0 * (double)n * ((need_y ? 2 : 1) * (m + nc)));
         ctx->launches++;
         PB_CUDA(cudaGetLastError());
      }
   }
   // Gram blocks G = P(:,0:nG)^H P(:,0:nG), H = P(:,0:nH)^H Q(:,0:nH): panels of <= 8 columns
   if (o->G_host && o->nG > 0)
      for (int c0 = 0; c0 < o->nG; c0 += 8) {
         const int nc = o->nG - c0 < 8 ? o->nG - c0 : 8;
         PB_CHK(pb200_zortho_sweep(ctx, n, NULL, 0, 0, P, o->nG, lds, P + (size_t)lds * c0, nc, lds, NULL, 0, NULL, 0, 0,
               (zc *)o->G_host + (size_t)o->ldG * c0, o->ldG));
      }
   if (o->H_host && o->nH > 0)
      for (int c0 = 0; c0 < o->nH; c0 += 8) {
         const int nc = o->nH - c0 < 8 ? o->nH - c0 : 8;
         PB_CHK(pb200_zortho_sweep(ctx, n, NULL, 0, 0, P, o->nH, lds, Q + (size_t)lds * c0, nc, lds, NULL, 0, NULL, 0, 0,
               (zc *)o->H_host + (size_t)o->ldH * c0, o->ldH));
      }
   for (int t = 0; t < 3; t++)
      if (o->X[t].ptr && o->X[t].ce > o->X[t].cb)
         PB_CHK(pb200_copy_d2d(ctx, P + (size_t)lds * o->X[t].cb, lds, o->X[t].ptr, o->X[t].ld, n, o->X[t].ce - o->X[t].cb, 16));
   if (o->Wo.ptr && o->Wo.ce > o->Wo.cb)
      PB_CHK(pb200_copy_d2d(ctx, Q + (size_t)lds * o->Wo.cb, lds, o->Wo.ptr, o->Wo.ld, n, o->Wo.ce - o->Wo.cb, 16));
   // residuals: Q_j <- Q_j - theta_j P_j in scratch with their squared norms
   std::vector<double> n2(nh, -1.0);
   auto residual_cols = [&](int cb, int ce) -> int {
      for (int c0 = cb; c0 < ce; c0 += 8) {
         int c1 = c0 + 8 < ce ? c0 + 8 : ce, lo = c0;
         while (lo < c1 && n2[lo] >= 0.0) lo++;
         if (lo >= c1) continue;
         PB_CHK(pb200_zresidual_inplace(ctx, n, theta_host + lo, P + (size_t)lds * lo, lds, Q + (size_t)lds * lo, lds, c1 - lo,
               &n2[lo]));
      }
      return 0;
   };
   if (o->R.ptr && o->R.ce > o->R.cb) {
      PB_CHK(residual_cols(o->R.cb, o->R.ce));
      PB_CHK(pb200_copy_d2d(ctx, Q + (size_t)lds * o->R.cb, lds, o->R.ptr, o->R.ld, n, o->R.ce - o->R.cb, 16));
      if (o->R2) PB_CHK(pb200_copy_d2d(ctx, Q + (size_t)lds * o->R.cb, lds, o->R2, o->ldR2, n, o->R.ce - o->R.cb, 16));
      if (o->Rnorms_host)
         for (int c = o->R.cb; c < o->R.ce; c++) o->Rnorms_host[c - o->R.cb] = sqrt(n2[c]);
   }
   if (o->rnorms_host && o->re > o->rb) {
      PB_CHK(residual_cols(o->rb, o->re));
      for (int c = o->rb; c < o->re; c++) o->rnorms_host[c - o->rb] = sqrt(n2[c]);
   }
   return 0;
}

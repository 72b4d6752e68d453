/* oracle/kernels_ref_z.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C99 complex restatement of the hot-path kernels behind zprimme: the complex twins declared in
 * include/primme_b200.h (pb200_zspmm, pb200_zortho_sweep, pb200_zvwxr, the multivector utilities).
 * Same role as kernels_ref.c: the checker of the CUDA kernels (tests/test_zkernels_gpu.py) and the
 * kernel layer of oracle/_build/libprimme_hostcheck.so.  Parity pin: the whole zprimme solver built on
 * these functions is compared with the reference's own zprimme (oracle/_ref/libprimme_ref.so) and its
 * stored golden vectors sol_10N_doublecomplex (tests/test_zprimme_cpu.py, tests/test_driver_cpu.py).
 *
 * "^H" is the conjugate transpose exactly where the reference's complex instantiation uses one
 * (template_types.h:51-204; Num_gemm 'C', Num_dot = zdotc).
 */
#include "kernels_ref.h"
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef double _Complex zc;

/* restates tests/COMMON/mat.c:68-100 for the complex driver (zamux per column) */
int pb200_zspmm(pb200_ctx *ctx, const pb200_csr *A, const void *X_, int64_t ldx, void *Y_, int64_t ldy, int ncols) {
   if (!A->is_complex) return PB200_ERR_ARG;
   ctx->launches++;
   const zc *X = (const zc *)X_, *vals = (const zc *)A->vals;
   zc *Y = (zc *)Y_;
   for (int c = 0; c < ncols; c++) {
      const zc *x = X + (size_t)c * ldx;
      zc *y = Y + (size_t)c * ldy;
      for (int64_t i = 0; i < A->nrows; i++) {
         zc t = 0.0;
         for (int64_t k = A->rowptr[i]; k < A->rowptr[i + 1]; k++) t += vals[k] * x[A->colind[k]];
         y[i] = t;
      }
   }
   return 0;
}

/* restates src/eigs/ortho.c:963-1072 (Num_ortho_kernel), complex instantiation: update :1017-1038
 * (no conjugation), Gram :1043-1059 ('C' = conjugate transpose of [Q V X]) */
int pb200_zortho_sweep(pb200_ctx *ctx, int64_t n, const void *Q_, int q, int64_t ldq, const void *V_, int mv,
      int64_t ldv, void *X_, int b, int64_t ldx, const void *C_, int ldc, const void *Y_, int ldy, int xx,
      void *P_, int ldp) {
   ctx->launches++;
   const zc *Q = (const zc *)Q_, *V = (const zc *)V_, *C = (const zc *)C_, *Y = (const zc *)Y_;
   zc *X = (zc *)X_, *P = (zc *)P_;
   const int k = q + mv;
   if (C || Y) {
      zc *t = (zc *)malloc(sizeof(zc) * (b > 0 ? b : 1));
      for (int64_t r = 0; r < n; r++) {
         for (int c = 0; c < b; c++) {
            zc s = X[r + (size_t)c * ldx];
            if (C) {
               for (int j = 0; j < q; j++) s -= Q[r + (size_t)j * ldq] * C[j + (size_t)c * ldc];
               for (int j = 0; j < mv; j++) s -= V[r + (size_t)j * ldv] * C[q + j + (size_t)c * ldc];
            }
            t[c] = s;
         }
         for (int c = 0; c < b; c++) {
            zc s;
            if (Y) {
               s = 0.0;
               for (int cc = 0; cc < b; cc++) s += t[cc] * Y[cc + (size_t)c * ldy];
            } else
               s = t[c];
            X[r + (size_t)c * ldx] = s;
         }
      }
      free(t);
   }
   if (P) {
      const int rows = k + (xx ? b : 0);
      for (int c = 0; c < b; c++) {
         const zc *x = X + (size_t)c * ldx;
         for (int j = 0; j < rows; j++) {
            const zc *a = j < q ? Q + (size_t)j * ldq : j < k ? V + (size_t)(j - q) * ldv : X + (size_t)(j - k) * ldx;
            zc s = 0.0;
            for (int64_t r = 0; r < n; r++) s += conj(a[r]) * x[r];
            P[j + (size_t)c * ldp] = s;
         }
      }
   }
   return 0;
}

/* restates src/eigs/auxiliary_eigs_normal.c:155-388 (Num_update_VWXR_Sprimme), complex instantiation,
 * B = I: X = V*h :254, Y = W*h :271, G = X^H X :299, H = X^H Y :306, R = Y - X*theta and ||R|| :313-339 */
int pb200_zvwxr(pb200_ctx *ctx, int64_t n, const void *V_, const void *W_, int m, int64_t ld, const void *h_,
      int ldh, int nh, const double *theta, const pb200_vwxr_out *o) {
   ctx->launches++;
   const zc *V = (const zc *)V_, *W = (const zc *)W_, *h = (const zc *)h_;
   const int nR = o->R.ptr ? o->R.ce - o->R.cb : 0;
   const int nr = o->rnorms_host ? o->re - o->rb : 0;
   double *Rn = (double *)calloc(nR + nr + 1, sizeof(double));
   double *rn = Rn + nR;
   zc *xrow = (zc *)malloc(sizeof(zc) * 2 * (nh > 0 ? nh : 1));
   zc *yrow = xrow + nh;
   zc *G = (zc *)o->G_host, *H = (zc *)o->H_host, *P = (zc *)o->P_host;
   if (G)
      for (int j = 0; j < o->nG; j++)
         for (int i = 0; i < o->nG; i++) G[i + (size_t)j * o->ldG] = 0.0;
   if (H)
      for (int j = 0; j < o->nH; j++)
         for (int i = 0; i < o->nH; i++) H[i + (size_t)j * o->ldH] = 0.0;
   if (P)
      for (int j = 0; j < nR; j++)
         for (int i = 0; i < m + nR; i++) P[i + (size_t)j * o->ldP] = 0.0;
   for (int64_t r = 0; r < n; r++) {
      for (int c = 0; c < nh; c++) {
         zc sx = 0.0, sy = 0.0;
         for (int k = 0; k < m; k++) {
            sx += V[r + (size_t)k * ld] * h[k + (size_t)c * ldh];
            sy += W[r + (size_t)k * ld] * h[k + (size_t)c * ldh];
         }
         xrow[c] = sx, yrow[c] = sy;
      }
      if (P)
         for (int j = 0; j < nR; j++) {
            const zc rj = yrow[o->R.cb + j] - xrow[o->R.cb + j] * theta[o->R.cb + j];
            for (int i = 0; i < m; i++) P[i + (size_t)j * o->ldP] += conj(V[r + (size_t)i * ld]) * rj;
            for (int i = 0; i < nR; i++)
               P[m + i + (size_t)j * o->ldP] += conj(yrow[o->R.cb + i] - xrow[o->R.cb + i] * theta[o->R.cb + i]) * rj;
         }
      for (int t = 0; t < 3; t++)
         if (o->X[t].ptr)
            for (int c = o->X[t].cb; c < o->X[t].ce; c++)
               ((zc *)o->X[t].ptr)[r + (size_t)(c - o->X[t].cb) * o->X[t].ld] = xrow[c];
      if (o->Wo.ptr)
         for (int c = o->Wo.cb; c < o->Wo.ce; c++) ((zc *)o->Wo.ptr)[r + (size_t)(c - o->Wo.cb) * o->Wo.ld] = yrow[c];
      for (int c = 0; c < nR; c++) {
         const int cc = o->R.cb + c;
         const zc v = yrow[cc] - xrow[cc] * theta[cc];
         ((zc *)o->R.ptr)[r + (size_t)c * o->R.ld] = v;
         if (o->R2) ((zc *)o->R2)[r + (size_t)c * o->ldR2] = v;
         Rn[c] += creal(v) * creal(v) + cimag(v) * cimag(v);
      }
      for (int c = 0; c < nr; c++) {
         const int cc = o->rb + c;
         const zc v = yrow[cc] - xrow[cc] * theta[cc];
         rn[c] += creal(v) * creal(v) + cimag(v) * cimag(v);
      }
      if (G)
         for (int j = 0; j < o->nG; j++)
            for (int i = 0; i < o->nG; i++) G[i + (size_t)j * o->ldG] += conj(xrow[i]) * xrow[j];
      if (H)
         for (int j = 0; j < o->nH; j++)
            for (int i = 0; i < o->nH; i++) H[i + (size_t)j * o->ldH] += conj(xrow[i]) * yrow[j];
   }
   if (o->Rnorms_host)
      for (int c = 0; c < nR; c++) o->Rnorms_host[c] = sqrt(Rn[c]);
   for (int c = 0; c < nr; c++) o->rnorms_host[c] = sqrt(rn[c]);
   free(Rn), free(xrow);
   return 0;
}

int pb200_zvwxr_can_fuse_gram(pb200_ctx *ctx, int64_t n, const void *V, const void *W, int m, int64_t ld, int nh,
      const pb200_vwxr_out *o) {
   (void)ctx, (void)n, (void)V, (void)W, (void)m, (void)ld, (void)nh, (void)o;
   return 0; /* the complex candidates sweep is not fused with the first Gram panel */
}

/* restates src/linalg/auxiliary.c:716-793 (permute_vecs) */
int pb200_zpermute_columns(pb200_ctx *ctx, int64_t n, void *X_, int64_t ldx, const int *perm, int ncols) {
   ctx->launches++;
   zc *X = (zc *)X_;
   zc *tmp = (zc *)malloc(sizeof(zc) * (size_t)(n > 0 ? n : 1) * (ncols > 0 ? ncols : 1));
   for (int i = 0; i < ncols; i++) memcpy(tmp + (size_t)i * n, X + (size_t)perm[i] * ldx, sizeof(zc) * n);
   for (int i = 0; i < ncols; i++) memcpy(X + (size_t)i * ldx, tmp + (size_t)i * n, sizeof(zc) * n);
   free(tmp);
   return 0;
}
int pb200_zcopy_columns(pb200_ctx *ctx, int64_t n, const void *X_, int64_t ldx, const int *xin, void *Y_, int64_t ldy,
      const int *yin, int ncols) {
   ctx->launches++;
   const zc *X = (const zc *)X_;
   zc *Y = (zc *)Y_;
   for (int i = 0; i < ncols; i++)
      memmove(Y + (size_t)(yin ? yin[i] : i) * ldy, X + (size_t)(xin ? xin[i] : i) * ldx, sizeof(zc) * n);
   return 0;
}
int pb200_zaxpy_columns(pb200_ctx *ctx, int64_t n, const void *alpha_, const void *X_, int64_t ldx, void *Y_,
      int64_t ldy, int ncols) {
   ctx->launches++;
   const zc *alpha = (const zc *)alpha_, *X = (const zc *)X_;
   zc *Y = (zc *)Y_;
   for (int j = 0; j < ncols; j++)
      for (int64_t r = 0; r < n; r++) Y[r + (size_t)j * ldy] += alpha[j] * X[r + (size_t)j * ldx];
   return 0;
}
int pb200_zscale_columns(pb200_ctx *ctx, int64_t n, const void *alpha_, void *X_, int64_t ldx, int ncols) {
   ctx->launches++;
   const zc *alpha = (const zc *)alpha_;
   zc *X = (zc *)X_;
   for (int j = 0; j < ncols; j++)
      for (int64_t r = 0; r < n; r++) X[r + (size_t)j * ldx] *= alpha[j];
   return 0;
}
/* restates Num_dist_dots (src/eigs/auxiliary_eigs.c:662-673): x^H y */
int pb200_zcolumn_dots(pb200_ctx *ctx, int64_t n, const void *X_, int64_t ldx, const void *Y_, int64_t ldy, int ncols,
      void *out_) {
   ctx->launches++;
   const zc *X = (const zc *)X_, *Y = (const zc *)Y_;
   zc *out = (zc *)out_;
   for (int j = 0; j < ncols; j++) {
      zc s = 0.0;
      for (int64_t r = 0; r < n; r++) s += conj(X[r + (size_t)j * ldx]) * Y[r + (size_t)j * ldy];
      out[j] = s;
   }
   return 0;
}
/* restates verify_norms' loop (src/eigs/main_iter.c:1872-1877) */
int pb200_zresidual_inplace(pb200_ctx *ctx, int64_t n, const double *theta, const void *V_, int64_t ldv, void *W_,
      int64_t ldw, int ncols, double *out) {
   ctx->launches++;
   const zc *V = (const zc *)V_;
   zc *W = (zc *)W_;
   for (int j = 0; j < ncols; j++) {
      double s = 0.0;
      for (int64_t r = 0; r < n; r++) {
         const zc v = W[r + (size_t)j * ldw] - theta[j] * V[r + (size_t)j * ldv];
         W[r + (size_t)j * ldw] = v;
         s += creal(v) * creal(v) + cimag(v) * cimag(v);
      }
      out[j] = s;
   }
   return 0;
}
/* restates tests/COMMON/mat.c:137-165 for complex blocks and the (real) diagonal of a Hermitian matrix */
int pb200_zjacobi(pb200_ctx *ctx, int64_t n, const double *diag, const double *shifts, double minabs, const void *X_,
      int64_t ldx, void *Y_, int64_t ldy, int ncols) {
   ctx->launches++;
   const zc *X = (const zc *)X_;
   zc *Y = (zc *)Y_;
   for (int j = 0; j < ncols; j++) {
      const double sh = shifts ? shifts[j] : 0.0;
      for (int64_t r = 0; r < n; r++) {
         double d = diag[r] - sh;
         if (fabs(d) < minabs) d = d < 0 ? -minabs : minabs;
         Y[r + (size_t)j * ldy] = X[r + (size_t)j * ldx] / d;
      }
   }
   return 0;
}

/* restates the solution update of one QMR step (src/eigs/inner_solve.c:384-413), complex vectors, real scalars */
int pb200_zqmr_update(pb200_ctx *ctx, int64_t n, const double *gamma, const double *eta, const void *D_, int64_t ldd,
      void *Delta_, int64_t ldl, void *Sol_, int64_t lds, int ncols, double *dots) {
   ctx->launches++;
   const zc *D = (const zc *)D_;
   zc *Delta = (zc *)Delta_, *Sol = (zc *)Sol_;
   for (int j = 0; j < ncols; j++) {
      double s2 = 0.0;
      for (int64_t r = 0; r < n; r++) {
         zc t = Delta[r + (size_t)j * ldl] * gamma[j];
         t += eta[j] * D[r + (size_t)j * ldd];
         Delta[r + (size_t)j * ldl] = t;
         const zc s = Sol[r + (size_t)j * lds] + t;
         Sol[r + (size_t)j * lds] = s;
         s2 += creal(s) * creal(s) + cimag(s) * cimag(s);
      }
      if (dots) dots[j] = s2;
   }
   return 0;
}

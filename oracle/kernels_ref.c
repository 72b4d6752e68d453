/* oracle/kernels_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the hot-path kernels, exporting the same C-ABI as
 * include/primme_b200.h so that (a) tests can compare the CUDA kernels against it entry point by
 * entry point and (b) the host solver sources can be linked against it into
 * oracle/_build/libprimme_hostcheck.so to check the host logic on a machine without a GPU.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the
 * product library (primme_b200/libprimme_b200.so) never links or calls it.
 *
 * Parity pin: tests/test_oracle_vs_reference.py checks these functions against the reference's
 * own Num_update_VWXR_dprimme / Num_gemm / CSR matvec built into oracle/_ref/libprimme_ref.so,
 * and the whole-solver results against the reference's dprimme and its golden sol_* files.
 *
 * Each function cites the reference lines it restates (relative to /root/reference).
 */
#include "../include/primme_b200.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kernels_ref.h"

int pb200_device_count(void) { return 1; /* the CPU itself; oracle only */ }

int pb200_ctx_create(pb200_ctx **ctx, int device) {
   (void)device;
   *ctx = (pb200_ctx *)calloc(1, sizeof(pb200_ctx));
   if (!*ctx) return PB200_ERR_ALLOC;
   (*ctx)->nranks = 1;
   return 0;
}
int pb200_ctx_destroy(pb200_ctx *ctx) {
   if (ctx)
      for (int s = 0; s < 4; s++) free(ctx->ws_ptr[s]);
   free(ctx);
   return 0;
}
int64_t pb200_ctx_l2_persist(pb200_ctx *ctx, const void *ptr, size_t bytes) {
   (void)ctx, (void)ptr, (void)bytes;
   return 0;
}
int pb200_ctx_begin_solve(pb200_ctx *ctx) {
   (void)ctx;
   return 0;
}
int pb200_ctx_sync(pb200_ctx *ctx) {
   (void)ctx;
   return 0;
}
void *pb200_ctx_stream(pb200_ctx *ctx) {
   (void)ctx;
   return NULL;
}
int64_t pb200_ctx_launches(pb200_ctx *ctx) { return ctx->launches; }
int pb200_ctx_set_comm(pb200_ctx *ctx, void *comm, int nranks, int rank) {
   (void)comm;
   ctx->nranks = nranks;
   ctx->rank = rank;
   return nranks == 1 ? 0 : PB200_ERR_ARG; /* the oracle is sequential */
}
int pb200_ctx_nranks(pb200_ctx *ctx) { return ctx->nranks; }
int pb200_ctx_set_profiling(pb200_ctx *ctx, int on) {
   (void)ctx, (void)on;
   return 0;
}
int pb200_ctx_get_profile(pb200_ctx *ctx, int kind, int64_t *count, double *ms, double *bytes) {
   (void)ctx, (void)kind;
   *count = 0, *ms = 0.0, *bytes = 0.0;
   return 0;
}
int pb200_allreduce_host(pb200_ctx *ctx, double *buf, int count) {
   (void)ctx, (void)buf, (void)count;
   return 0;
}
int pb200_bcast_host(pb200_ctx *ctx, double *buf, int count, int root) {
   (void)ctx, (void)buf, (void)count, (void)root;
   return 0;
}

/* multi-rank pieces: the oracle is sequential (N > 1 host logic is tested with gloo callbacks) */
int pb200_comm_unique_id(void *id128) { (void)id128; return PB200_ERR_ARG; }
int pb200_ctx_comm_init(pb200_ctx *ctx, int nranks, int rank, const void *id) { (void)ctx, (void)nranks, (void)rank, (void)id; return PB200_ERR_ARG; }
int pb200_ctx_comm_free(pb200_ctx *ctx) { (void)ctx; return 0; }
int pb200_dist_csr_create(pb200_ctx *ctx, pb200_csr *A, const int64_t *c, int nr, pb200_dist_csr **D) { (void)ctx, (void)A, (void)c, (void)nr, (void)D; return PB200_ERR_ARG; }
int pb200_dist_csr_destroy(pb200_ctx *ctx, pb200_dist_csr *D) { (void)ctx, (void)D; return 0; }
int pb200_ddist_spmm(pb200_ctx *ctx, pb200_dist_csr *D, const double *X, int64_t ldx, double *Y, int64_t ldy, int nc) { (void)ctx, (void)D, (void)X, (void)ldx, (void)Y, (void)ldy, (void)nc; return PB200_ERR_ARG; }
int pb200_zdist_spmm(pb200_ctx *ctx, pb200_dist_csr *D, const void *X, int64_t ldx, void *Y, int64_t ldy, int nc) { (void)ctx, (void)D, (void)X, (void)ldx, (void)Y, (void)ldy, (void)nc; return PB200_ERR_ARG; }
int pb200_dist_csr_info(const pb200_dist_csr *D, int64_t *a, int64_t *b, int64_t *c, int *d) { (void)D, (void)a, (void)b, (void)c, (void)d; return PB200_ERR_ARG; }
int pb200_ctx_peer_export(pb200_ctx *ctx, void *h) { (void)ctx, (void)h; return PB200_ERR_ARG; }
int pb200_ctx_peer_attach(pb200_ctx *ctx, int n, int r, const void *h) { (void)ctx, (void)n, (void)r, (void)h; return PB200_ERR_ARG; }
int pb200_ctx_peer_active(pb200_ctx *ctx) { (void)ctx; return 0; }
void primme_b200_svds_dist_csr_matvec(void *x, int64_t *ldx, void *y, int64_t *ldy, int *bs, int *tr, struct primme_svds_params *p, int *ierr) { (void)x, (void)ldx, (void)y, (void)ldy, (void)bs, (void)tr, (void)p; *ierr = -1; }
void primme_b200_dist_csr_matvec(void *x, int64_t *ldx, void *y, int64_t *ldy, int *bs, struct primme_params *p, int *ierr) { (void)x, (void)ldx, (void)y, (void)ldy, (void)bs, (void)p; *ierr = -1; }

int pb200_malloc(pb200_ctx *ctx, size_t bytes, void **dptr) {
   (void)ctx;
   *dptr = malloc(bytes ? bytes : 1);
   return *dptr ? 0 : PB200_ERR_ALLOC;
}
int pb200_free(pb200_ctx *ctx, void *dptr) {
   (void)ctx;
   free(dptr);
   return 0;
}
int pb200_ctx_workspace(pb200_ctx *ctx, int slot, size_t bytes, void **dptr) {
   if (slot < 0 || slot >= 4) return PB200_ERR_ARG;
   if (bytes > ctx->ws_bytes[slot]) {
      free(ctx->ws_ptr[slot]);
      ctx->ws_ptr[slot] = malloc(bytes ? bytes : 1);
      ctx->ws_bytes[slot] = ctx->ws_ptr[slot] ? bytes : 0;
   }
   *dptr = ctx->ws_ptr[slot];
   return *dptr ? 0 : PB200_ERR_ALLOC;
}
int pb200_memset0(pb200_ctx *ctx, void *dptr, size_t bytes) {
   (void)ctx;
   memset(dptr, 0, bytes);
   return 0;
}
static int copy2d(const void *src, int64_t lds, void *dst, int64_t ldd, int64_t rows, int cols,
      int es) {
   for (int j = 0; j < cols; j++)
      memmove((char *)dst + (size_t)j * ldd * es, (const char *)src + (size_t)j * lds * es,
            (size_t)rows * es);
   return 0;
}
int pb200_copy_h2d(pb200_ctx *ctx, const void *s, int64_t lds, void *d, int64_t ldd,
      int64_t rows, int cols, int es) {
   (void)ctx;
   return copy2d(s, lds, d, ldd, rows, cols, es);
}
int pb200_copy_d2h(pb200_ctx *ctx, const void *s, int64_t lds, void *d, int64_t ldd,
      int64_t rows, int cols, int es) {
   (void)ctx;
   return copy2d(s, lds, d, ldd, rows, cols, es);
}
int pb200_copy_d2d(pb200_ctx *ctx, const void *s, int64_t lds, void *d, int64_t ldd,
      int64_t rows, int cols, int es) {
   (void)ctx;
   return copy2d(s, lds, d, ldd, rows, cols, es);
}
int pb200_is_device_pointer(const void *p) {
   (void)p;
   return 1; /* in the oracle every pointer is "device" memory */
}

/* ---------------------------------------------------------------- K1: CSR SpMM ------------ */
/* restates tests/COMMON/mat.c:68-100 (CSRMatrixMatvec: SPARSKIT amux per column) and
 * examples/ex_eigs_dcublas.c:238-263 (cusparseSpMM, CSR x dense col-major block). */
int pb200_csr_create(pb200_ctx *ctx, int64_t nrows, int64_t ncols, int64_t nnz,
      const int64_t *rowptr, const int32_t *colind, const void *vals, int index_base,
      int is_complex, pb200_csr **out) {
   (void)ctx;
   pb200_csr *A = (pb200_csr *)calloc(1, sizeof(*A));
   if (!A) return PB200_ERR_ALLOC;
   A->nrows = nrows, A->ncols = ncols, A->nnz = nnz, A->is_complex = is_complex;
   A->rowptr = (int64_t *)malloc(sizeof(int64_t) * (nrows + 1));
   A->colind = (int32_t *)malloc(sizeof(int32_t) * (nnz ? nnz : 1));
   size_t vb = sizeof(double) * (is_complex ? 2 : 1) * (nnz ? nnz : 1);
   A->vals = (double *)malloc(vb);
   if (!A->rowptr || !A->colind || !A->vals) return PB200_ERR_ALLOC;
   for (int64_t i = 0; i <= nrows; i++) A->rowptr[i] = rowptr[i] - index_base;
   for (int64_t i = 0; i < nnz; i++) A->colind[i] = colind[i] - index_base;
   memcpy(A->vals, vals, sizeof(double) * (is_complex ? 2 : 1) * nnz);
   *out = A;
   return 0;
}
int pb200_csr_create_pooled(pb200_ctx *ctx, int64_t nrows, int64_t ncols, int64_t nnz, const int64_t *rowptr,
      const int32_t *colind, const void *vals, int index_base, int is_complex, pb200_csr **out) {
   return pb200_csr_create(ctx, nrows, ncols, nnz, rowptr, colind, vals, index_base, is_complex, out);
}
int pb200_csr_destroy(pb200_ctx *ctx, pb200_csr *A) {
   if (!A) return 0;
   if (A->T) pb200_csr_destroy(ctx, A->T);
   free(A->rowptr), free(A->colind), free(A->vals), free(A);
   return 0;
}
int64_t pb200_csr_nnz(const pb200_csr *A) { return A->nnz; }
int pb200_csr_layout(const pb200_csr *A, int ncols) { (void)A, (void)ncols; return 1; }
int pb200_csr_is_complex(const pb200_csr *A) { return A->is_complex; }

int pb200_dspmm(pb200_ctx *ctx, const pb200_csr *A, const double *X, int64_t ldx, double *Y,
      int64_t ldy, int ncols) {
   if (A->is_complex) return PB200_ERR_ARG;
   ctx->launches++;
   for (int c = 0; c < ncols; c++) {
      const double *x = X + (size_t)c * ldx;
      double *y = Y + (size_t)c * ldy;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < A->nrows; i++) {
         double t = 0.0;
         for (int64_t k = A->rowptr[i]; k < A->rowptr[i + 1]; k++)
            t += A->vals[k] * x[A->colind[k]];
         y[i] = t;
      }
   }
   return 0;
}

int pb200_csr_build_transpose(pb200_ctx *ctx, pb200_csr *A) {
   if (A->T) return 0;
   pb200_csr *T = (pb200_csr *)calloc(1, sizeof(*T));
   T->nrows = A->ncols, T->ncols = A->nrows, T->nnz = A->nnz;
   T->rowptr = (int64_t *)calloc(T->nrows + 1, sizeof(int64_t));
   T->colind = (int32_t *)malloc(sizeof(int32_t) * (A->nnz ? A->nnz : 1));
   T->vals = (double *)malloc(sizeof(double) * (A->nnz ? A->nnz : 1));
   for (int64_t k = 0; k < A->nnz; k++) T->rowptr[A->colind[k] + 1]++;
   for (int64_t i = 0; i < T->nrows; i++) T->rowptr[i + 1] += T->rowptr[i];
   int64_t *next = (int64_t *)malloc(sizeof(int64_t) * (T->nrows + 1));
   memcpy(next, T->rowptr, sizeof(int64_t) * (T->nrows + 1));
   for (int64_t i = 0; i < A->nrows; i++)
      for (int64_t k = A->rowptr[i]; k < A->rowptr[i + 1]; k++) {
         int64_t p = next[A->colind[k]]++;
         T->colind[p] = (int32_t)i;
         T->vals[p] = A->vals[k];
      }
   free(next);
   A->T = T;
   (void)ctx;
   return 0;
}
/* restates src/svds/primme_svds_c.c:1337-1351 "transpose" leg of the normal-equations operator
 * (tests/COMMON/mat.c CSRMatrixMatvecSVD -> atmuxr) */
int pb200_dspmm_t(pb200_ctx *ctx, const pb200_csr *A, const double *X, int64_t ldx, double *Y,
      int64_t ldy, int ncols) {
   if (!A->T) return PB200_ERR_ARG;
   return pb200_dspmm(ctx, A->T, X, ldx, Y, ldy, ncols);
}

/* ------------------------------------------------------- K2/K3/K4: fused ortho row sweep -- */
/* restates src/eigs/ortho.c:963-1072 (Num_ortho_kernel): update :1017-1038, Gram :1043-1059;
 * with xx=0 and X = W block it restates src/eigs/update_projection.c:99-102. */
int pb200_dortho_sweep(pb200_ctx *ctx, int64_t n, const double *Q, int q, int64_t ldq,
      const double *V, int mv, int64_t ldv, double *X, int b, int64_t ldx, const double *C,
      int ldc, const double *Y, int ldy, int xx, double *P, int ldp) {
   ctx->launches++;
   int k = q + mv;
   if (C) {
      double *t = (double *)malloc(sizeof(double) * (b > 0 ? b : 1));
      for (int64_t r = 0; r < n; r++) {
         for (int c = 0; c < b; c++) {
            double s = X[r + (size_t)c * ldx];
            for (int j = 0; j < q; j++) s -= Q[r + (size_t)j * ldq] * C[j + (size_t)c * ldc];
            for (int j = 0; j < mv; j++)
               s -= V[r + (size_t)j * ldv] * C[q + j + (size_t)c * ldc];
            t[c] = s;
         }
         for (int c = 0; c < b; c++) {
            double s;
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
      int rows = k + (xx ? b : 0);
      for (int c = 0; c < b; c++) {
         const double *x = X + (size_t)c * ldx;
         for (int j = 0; j < rows; j++) {
            const double *a = j < q        ? Q + (size_t)j * ldq
                              : j < k      ? V + (size_t)(j - q) * ldv
                                           : X + (size_t)(j - k) * ldx;
            double s = 0.0;
            for (int64_t r = 0; r < n; r++) s += a[r] * x[r];
            P[j + (size_t)c * ldp] = s;
         }
      }
   }
   return 0;
}

/* ------------------------------------------------------------------------ K5: VWXR -------- */
/* restates src/eigs/auxiliary_eigs_normal.c:155-388 (Num_update_VWXR_Sprimme) with B = I:
 * X=V*h :254, Y=W*h :271, scatter :258-276, G :299, H :306, R and norms :313-339 (+ :70-99). */
int pb200_dvwxr(pb200_ctx *ctx, int64_t n, const double *V, const double *W, int m, int64_t ld,
      const double *h, int ldh, int nh, const double *theta, const pb200_vwxr_out *o) {
   ctx->launches++;
   int nR = o->R.ptr ? o->R.ce - o->R.cb : 0;
   int nr = o->rnorms_host ? o->re - o->rb : 0;
   double *Rn = (double *)calloc(nR + nr + 1, sizeof(double));
   double *rn = Rn + nR;
   double *xrow = (double *)malloc(sizeof(double) * 2 * (nh > 0 ? nh : 1));
   double *yrow = xrow + nh;
   if (o->G_host)
      for (int j = 0; j < o->nG; j++)
         for (int i = 0; i < o->nG; i++) o->G_host[i + (size_t)j * o->ldG] = 0.0;
   if (o->H_host)
      for (int j = 0; j < o->nH; j++)
         for (int i = 0; i < o->nH; i++) o->H_host[i + (size_t)j * o->ldH] = 0.0;
   if (o->P_host)
      for (int j = 0; j < nR; j++)
         for (int i = 0; i < m + nR; i++) o->P_host[i + (size_t)j * o->ldP] = 0.0;
   for (int64_t r = 0; r < n; r++) {
      /* whole row of V*h and W*h first: outputs may alias V and W (restart.c:692-705) */
      for (int c = 0; c < nh; c++) {
         double sx = 0.0, sy = 0.0;
         for (int k = 0; k < m; k++) {
            sx += V[r + (size_t)k * ld] * h[k + (size_t)c * ldh];
            sy += W[r + (size_t)k * ld] * h[k + (size_t)c * ldh];
         }
         xrow[c] = sx, yrow[c] = sy;
      }
      /* optional first Gram panel of the block orthogonalisation, P = [V R]^H R
       * (src/eigs/ortho.c:1043-1059 with X = R), before any output overwrites the row */
      if (o->P_host)
         for (int j = 0; j < nR; j++) {
            const double rj = yrow[o->R.cb + j] - xrow[o->R.cb + j] * theta[o->R.cb + j];
            for (int i = 0; i < m; i++) o->P_host[i + (size_t)j * o->ldP] += V[r + (size_t)i * ld] * rj;
            for (int i = 0; i < nR; i++)
               o->P_host[m + i + (size_t)j * o->ldP] +=
                     (yrow[o->R.cb + i] - xrow[o->R.cb + i] * theta[o->R.cb + i]) * rj;
         }
      for (int t = 0; t < 3; t++)
         if (o->X[t].ptr)
            for (int c = o->X[t].cb; c < o->X[t].ce; c++)
               o->X[t].ptr[r + (size_t)(c - o->X[t].cb) * o->X[t].ld] = xrow[c];
      if (o->Wo.ptr)
         for (int c = o->Wo.cb; c < o->Wo.ce; c++)
            o->Wo.ptr[r + (size_t)(c - o->Wo.cb) * o->Wo.ld] = yrow[c];
      for (int c = 0; c < nR; c++) {
         int cc = o->R.cb + c;
         double v = yrow[cc] - xrow[cc] * theta[cc];
         o->R.ptr[r + (size_t)c * o->R.ld] = v;
         if (o->R2) o->R2[r + (size_t)c * o->ldR2] = v; /* second destination of the residual block */
         Rn[c] += v * v;
      }
      for (int c = 0; c < nr; c++) {
         int cc = o->rb + c;
         double v = yrow[cc] - xrow[cc] * theta[cc];
         rn[c] += v * v;
      }
      if (o->G_host)
         for (int j = 0; j < o->nG; j++)
            for (int i = 0; i < o->nG; i++) o->G_host[i + (size_t)j * o->ldG] += xrow[i] * xrow[j];
      if (o->H_host)
         for (int j = 0; j < o->nH; j++)
            for (int i = 0; i < o->nH; i++) o->H_host[i + (size_t)j * o->ldH] += xrow[i] * yrow[j];
   }
   if (o->Rnorms_host)
      for (int c = 0; c < nR; c++) o->Rnorms_host[c] = sqrt(Rn[c]);
   for (int c = 0; c < nr; c++) o->rnorms_host[c] = sqrt(rn[c]);
   free(Rn), free(xrow);
   return 0;
}

int pb200_dvwxr_can_fuse_gram(pb200_ctx *ctx, int64_t n, const double *V, const double *W, int m,
      int64_t ld, int nh, const pb200_vwxr_out *o) {
   (void)ctx, (void)V, (void)W, (void)ld, (void)n;
   if (getenv("PB200_NO_FUSE_GRAM")) return 0;
   return nh > 0 && nh <= 8 && !(o->G_host && o->nG > 0) && !(o->H_host && o->nH > 0) && o->R.ptr &&
          o->R.ce > o->R.cb && m > 0;
}

/* ------------------------------------------------------------------ K6: utilities --------- */
/* restates src/linalg/auxiliary.c:716-793 (permute_vecs): X(:,i) <- X(:,perm[i]) */
int pb200_dpermute_columns(pb200_ctx *ctx, int64_t n, double *X, int64_t ldx, const int *perm,
      int ncols) {
   ctx->launches++;
   double *tmp = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1) * (ncols > 0 ? ncols : 1));
   for (int i = 0; i < ncols; i++)
      memcpy(tmp + (size_t)i * n, X + (size_t)perm[i] * ldx, sizeof(double) * n);
   for (int i = 0; i < ncols; i++) memcpy(X + (size_t)i * ldx, tmp + (size_t)i * n, sizeof(double) * n);
   free(tmp);
   return 0;
}
/* restates src/linalg/auxiliary.c:649-662 (Num_copy_matrix_columns) */
int pb200_dcopy_columns(pb200_ctx *ctx, int64_t n, const double *X, int64_t ldx, const int *xin,
      double *Y, int64_t ldy, const int *yin, int ncols) {
   ctx->launches++;
   for (int i = 0; i < ncols; i++)
      memmove(Y + (size_t)(yin ? yin[i] : i) * ldy, X + (size_t)(xin ? xin[i] : i) * ldx,
            sizeof(double) * n);
   return 0;
}
/* restates Num_axpy per column (src/eigs/correction.c:366-369) */
int pb200_daxpy_columns(pb200_ctx *ctx, int64_t n, const double *alpha, const double *X,
      int64_t ldx, double *Y, int64_t ldy, int ncols) {
   ctx->launches++;
   for (int j = 0; j < ncols; j++)
      for (int64_t r = 0; r < n; r++) Y[r + (size_t)j * ldy] += alpha[j] * X[r + (size_t)j * ldx];
   return 0;
}
int pb200_dscale_columns(pb200_ctx *ctx, int64_t n, const double *alpha, double *X, int64_t ldx,
      int ncols) {
   ctx->launches++;
   for (int j = 0; j < ncols; j++)
      for (int64_t r = 0; r < n; r++) X[r + (size_t)j * ldx] *= alpha[j];
   return 0;
}
/* restates Num_dist_dots (src/eigs/auxiliary_eigs.c:662-673) */
int pb200_dcolumn_dots(pb200_ctx *ctx, int64_t n, const double *X, int64_t ldx, const double *Y,
      int64_t ldy, int ncols, double *out) {
   ctx->launches++;
   for (int j = 0; j < ncols; j++) {
      double s = 0.0;
      for (int64_t r = 0; r < n; r++) s += X[r + (size_t)j * ldx] * Y[r + (size_t)j * ldy];
      out[j] = s;
   }
   return 0;
}
/* restates verify_norms' loop (src/eigs/main_iter.c:1872-1877) */
int pb200_dresidual_inplace(pb200_ctx *ctx, int64_t n, const double *theta, const double *V,
      int64_t ldv, double *W, int64_t ldw, int ncols, double *out) {
   ctx->launches++;
   for (int j = 0; j < ncols; j++) {
      double s = 0.0;
      for (int64_t r = 0; r < n; r++) {
         double v = W[r + (size_t)j * ldw] - theta[j] * V[r + (size_t)j * ldv];
         W[r + (size_t)j * ldw] = v;
         s += v * v;
      }
      out[j] = s;
   }
   return 0;
}
/* restates tests/COMMON/mat.c:137-165 (ApplyInvDavidsonDiagPrecNative): y = x / (d - shift),
 * with |d - shift| floored at minabs keeping the sign */
int pb200_djacobi(pb200_ctx *ctx, int64_t n, const double *diag, const double *shifts,
      double minabs, const double *X, int64_t ldx, double *Y, int64_t ldy, int ncols) {
   ctx->launches++;
   for (int j = 0; j < ncols; j++) {
      double sh = shifts ? shifts[j] : 0.0;
      for (int64_t r = 0; r < n; r++) {
         double d = diag[r] - sh;
         if (fabs(d) < minabs) d = d < 0 ? -minabs : minabs;
         Y[r + (size_t)j * ldy] = X[r + (size_t)j * ldx] / d;
      }
   }
   return 0;
}

/* restates the solution update of one QMR step (src/eigs/inner_solve.c:384-413) */
int pb200_dqmr_update(pb200_ctx *ctx, int64_t n, const double *gamma, const double *eta, const double *D, int64_t ldd,
      double *Delta, int64_t ldl, double *Sol, int64_t lds, int ncols, double *dots) {
   ctx->launches++;
   for (int j = 0; j < ncols; j++) {
      double s2 = 0.0;
      for (int64_t r = 0; r < n; r++) {
         double t = Delta[r + (size_t)j * ldl] * gamma[j];
         t += eta[j] * D[r + (size_t)j * ldd];
         Delta[r + (size_t)j * ldl] = t;
         const double s = Sol[r + (size_t)j * lds] + t;
         Sol[r + (size_t)j * lds] = s;
         s2 += s * s;
      }
      if (dots) dots[j] = s2;
   }
   return 0;
}

/* the device generator's reference: LAPACK's dlarnv itself, column after column */
extern void dlarnv_(const int *idist, int *iseed, const int *n, double *x);
int pb200_dlarnv(pb200_ctx *ctx, long long iseed[4], int64_t col_len, int ncols, double *X, int64_t ld) {
   (void)ctx;
   int idist = 2, seed[4];
   for (int i = 0; i < 4; i++) seed[i] = (int)iseed[i];
   for (int j = 0; j < ncols; j++) {
      int64_t left = col_len;
      double *x = X + (size_t)j * ld;
      while (left > 0) {
         int chunk = left > 0x7ffffff0LL ? 0x7ffffff0 : (int)left;
         dlarnv_(&idist, seed, &chunk, x);
         x += chunk, left -= chunk;
      }
   }
   for (int i = 0; i < 4; i++) iseed[i] = seed[i];
   return 0;
}

/* operators.c -- ready-made device callbacks and the solver-context registry.
 *
 * The reference leaves the operator to the user (matrixMatvec callback); its GPU example wires
 * cusparseSpMM into that callback (examples/ex_eigs_dcublas.c:238-263) and its test driver a
 * host CSR loop plus a Jacobi preconditioner (tests/COMMON/mat.c:68-100,137-165).  These are the
 * B200-native counterparts with the same callback signature, so a caller only swaps the
 * function pointer:
 *     primme.matrix         = pb200_csr* (from pb200_csr_create)
 *     primme.matrixMatvec   = primme_b200_csr_matvec
 *     primme.preconditioner = primme_b200_jacobi* ; primme.applyPreconditioner = primme_b200_jacobi_apply
 * They launch on the running solver's stream, found through the registry below.
 */
#include "pb_host.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>

/* ---- registry: which kernel context serves which primme_params (a handful of entries) ---- */
#define PB_MAX_ACTIVE 64
static struct {
   const primme_params *primme;
   pb200_ctx *solver; /* set for the duration of a solve */
   pb200_ctx *user;   /* attached by the caller (e.g. with an NCCL communicator) */
} registry[PB_MAX_ACTIVE];

/* solves on different primme_params may run on different threads (the reference is re-entrant per
 * primme_params): slot allocation and release are serialised */
static pthread_mutex_t reg_lock = PTHREAD_MUTEX_INITIALIZER;

static int reg_find(const primme_params *p, int create) {
   int free_slot = -1;
   for (int i = 0; i < PB_MAX_ACTIVE; i++) {
      if (registry[i].primme == p) return i;
      if (!registry[i].primme && free_slot < 0) free_slot = i;
   }
   if (create && free_slot >= 0) {
      registry[free_slot].primme = p;
      registry[free_slot].solver = registry[free_slot].user = NULL;
   }
   return create ? free_slot : -1;
}

int primme_b200_attach_ctx(primme_params *primme, pb200_ctx *ctx) {
   pthread_mutex_lock(&reg_lock);
   int i = reg_find(primme, ctx != NULL), rc = 0;
   if (i < 0)
      rc = ctx ? PRIMME_MALLOC_FAILURE : 0;
   else {
      registry[i].user = ctx;
      if (!registry[i].user && !registry[i].solver) registry[i].primme = NULL;
   }
   pthread_mutex_unlock(&reg_lock);
   return rc;
}

pb200_ctx *primme_b200_attached_ctx(const primme_params *primme) {
   pthread_mutex_lock(&reg_lock);
   int i = reg_find(primme, 0);
   pb200_ctx *c = i < 0 ? NULL : registry[i].user;
   pthread_mutex_unlock(&reg_lock);
   return c;
}

void pb_registry_set_solver(const primme_params *primme, pb200_ctx *ctx) {
   pthread_mutex_lock(&reg_lock);
   int i = reg_find(primme, ctx != NULL);
   if (i >= 0) {
      registry[i].solver = ctx;
      if (!registry[i].user && !registry[i].solver) registry[i].primme = NULL;
   }
   pthread_mutex_unlock(&reg_lock);
}

pb200_ctx *primme_b200_solver_ctx(const primme_params *primme) {
   pthread_mutex_lock(&reg_lock);
   int i = reg_find(primme, 0);
   pb200_ctx *c = i < 0 ? NULL : (registry[i].solver ? registry[i].solver : registry[i].user);
   pthread_mutex_unlock(&reg_lock);
   return c;
}

/* ---- CSR block matvec: y = A x on device pointers ---- */
void primme_b200_csr_matvec(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy, int *blockSize,
      primme_params *primme, int *ierr) {
   pb200_ctx *ctx = primme_b200_solver_ctx(primme);
   const pb200_csr *A = (const pb200_csr *)primme->matrix;
   if (!ctx || !A) {
      *ierr = -1;
      return;
   }
   /* the matrix knows its scalar type: the same callback serves dprimme and zprimme */
   *ierr = pb200_csr_is_complex(A) ? pb200_zspmm(ctx, A, x, *ldx, y, *ldy, *blockSize)
                                   : pb200_dspmm(ctx, A, (const double *)x, *ldx, (double *)y, *ldy, *blockSize);
}

/* ---- Jacobi (diagonal) preconditioner with per-column shifts ---- */
/* struct primme_b200_jacobi is declared in include/primme_b200.h */

void primme_b200_jacobi_apply(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy, int *blockSize,
      primme_params *primme, int *ierr) {
   pb200_ctx *ctx = primme_b200_solver_ctx(primme);
   const primme_b200_jacobi *J = (const primme_b200_jacobi *)primme->preconditioner;
   if (!ctx || !J) {
      *ierr = -1;
      return;
   }
   const double *shifts = J->use_shifts ? primme->ShiftsForPreconditioner : NULL;
   *ierr = pb200_djacobi(ctx, primme->nLocal, J->diag_dev, shifts, J->minabs, (const double *)x,
         *ldx, (double *)y, *ldy, *blockSize);
}

/* complex blocks, real diagonal (the diagonal of a Hermitian matrix): zprimme's twin of the above */
void primme_b200_zjacobi_apply(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy, int *blockSize,
      primme_params *primme, int *ierr) {
   pb200_ctx *ctx = primme_b200_solver_ctx(primme);
   const primme_b200_jacobi *J = (const primme_b200_jacobi *)primme->preconditioner;
   if (!ctx || !J) {
      *ierr = -1;
      return;
   }
   const double *shifts = J->use_shifts ? primme->ShiftsForPreconditioner : NULL;
   *ierr = pb200_zjacobi(ctx, primme->nLocal, J->diag_dev, shifts, J->minabs, x, *ldx, y, *ldy, *blockSize);
}

/* ---- one-call convenience: host CSR + host result arrays, everything else on the device ----
 * What a CSR user of the reference would bind: upload the matrix, run the device solver with
 * the built-in SpMM, download the eigenvectors.  evecs_host is n x (numOrthoConst +
 * max(numEvals, initSize)) with leading dimension primme->ldevecs (or n). */
static int solve_csr(double *evals, void *evecs_host, double *resNorms, primme_params *primme,
      const int64_t *rowptr_host, const int32_t *colind_host, const void *vals_host, int index_base, int is_complex) {
   if (!primme || !rowptr_host) return -4;
   const int es = is_complex ? 16 : 8;
   pb200_ctx *ctx = NULL;
   int own = 0, rc;
   ctx = primme_b200_attached_ctx(primme);
   if (!ctx) {
      rc = pb200_ctx_create(&ctx, -1);
      if (rc) return PRIMME_FUNCTION_UNAVAILABLE;
      own = 1;
      primme_b200_attach_ctx(primme, ctx);
   }
   const int64_t n = primme->numProcs > 1 ? primme->nLocal : primme->n;
   const int64_t nnz = rowptr_host[n] - index_base;
   pb200_csr *A = NULL;
   void *devecs = NULL;
   const double tc0 = hl_wtime();
   /* a caller-attached (long-lived) context keeps the device storage of the matrix and of the eigenvectors
    * between calls: no cudaMalloc / cudaFree inside an end-to-end solve */
   rc = own ? pb200_csr_create(ctx, n, primme->n, nnz, rowptr_host, colind_host, vals_host, index_base, is_complex, &A)
            : pb200_csr_create_pooled(ctx, n, primme->n, nnz, rowptr_host, colind_host, vals_host, index_base, is_complex, &A);
   const double tc1 = hl_wtime();
   const int ncols = primme->numOrthoConst + PB_MAX(primme->numEvals, primme->initSize);
   const int64_t ldh = primme->ldevecs > 0 ? primme->ldevecs : n;
   if (!rc)
      rc = own ? pb200_malloc(ctx, (size_t)es * (size_t)PB_MAX(n, 1) * PB_MAX(ncols, 1), &devecs)
               : pb200_ctx_workspace(ctx, 2, (size_t)es * (size_t)PB_MAX(n, 1) * PB_MAX(ncols, 1), &devecs);
   if (!rc && primme->numOrthoConst + primme->initSize > 0)
      rc = pb200_copy_h2d(ctx, evecs_host, ldh, devecs, n, n, primme->numOrthoConst + primme->initSize, es);
   if (!rc) {
      void *old_matrix = primme->matrix;
      primme_block_op_fn old_mv = primme->matrixMatvec;
      PRIMME_INT old_ld = primme->ldevecs;
      primme->matrix = A;
      primme->matrixMatvec = primme_b200_csr_matvec;
      primme->ldevecs = n;
      if (primme->numProcs <= 1) primme->nLocal = n;
      rc = is_complex ? cublas_zprimme(evals, (PRIMME_COMPLEX_DOUBLE *)devecs, resNorms, primme)
                      : cublas_dprimme(evals, (double *)devecs, resNorms, primme);
      primme->matrix = old_matrix;
      primme->matrixMatvec = old_mv;
      primme->ldevecs = old_ld;
      int nret = primme->numOrthoConst + (primme->initSize > 0 ? primme->initSize : 0);
      if (nret > ncols) nret = ncols;
      if (rc == 0 || rc == PRIMME_MAIN_ITER_FAILURE) {
         int r2 = pb200_copy_d2h(ctx, devecs, n, evecs_host, ldh, n, PB_MAX(nret, primme->numOrthoConst + primme->numEvals < ncols ? primme->numOrthoConst + primme->numEvals : ncols), es);
         if (r2 && !rc) rc = PRIMME_UNEXPECTED_FAILURE;
      }
   }
   const double tc2 = hl_wtime();
   if (devecs && own) pb200_free(ctx, devecs);
   if (A) pb200_csr_destroy(ctx, A);
   if (getenv("PB200_DEBUG"))
      fprintf(stderr, "PRIMME-B200: dprimme_csr (s): csr upload+schedule %.4f, solve+download %.4f, release %.4f\n",
            tc1 - tc0, tc2 - tc1, hl_wtime() - tc2);
   if (own) {
      primme_b200_attach_ctx(primme, NULL);
      pb200_ctx_destroy(ctx);
   }
   return rc;
}

int primme_b200_dprimme_csr(double *evals, double *evecs_host, double *resNorms, primme_params *primme,
      const int64_t *rowptr_host, const int32_t *colind_host, const double *vals_host, int index_base) {
   return solve_csr(evals, evecs_host, resNorms, primme, rowptr_host, colind_host, vals_host, index_base, 0);
}
/* complex Hermitian twin: vals_host and evecs_host interleaved (re,im) */
int primme_b200_zprimme_csr(double *evals, void *evecs_host, double *resNorms, primme_params *primme,
      const int64_t *rowptr_host, const int32_t *colind_host, const void *vals_host, int index_base) {
   return solve_csr(evals, evecs_host, resNorms, primme, rowptr_host, colind_host, vals_host, index_base, 1);
}

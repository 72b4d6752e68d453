/* oracle/csr_host.c -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Host CSR block matvec and Jacobi preconditioner with PRIMME's callback signature, used to
 * drive the UNMODIFIED reference (oracle/_ref/libprimme_ref.so) on the CPU: the parity oracle of
 * the tests and the cpu_baseline / --impl reference arm of bench.py.  Restates the reference's
 * own test-driver callbacks (tests/COMMON/mat.c:68-100 CSRMatrixMatvec, :137-165
 * ApplyInvDavidsonDiagPrecNative); rows are split over pthreads so the baseline can use every
 * host core (OpenBLAS threads cover the dense part).
 */
#include "../include/primme.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>

typedef struct csr_host {
   int64_t n;
   const int64_t *rowptr; /* 0-based */
   const int32_t *colind; /* 0-based */
   const double *vals;
   int nthreads;
   const double *diag;    /* for the Jacobi preconditioner (optional) */
   double minabs;
   int use_shifts;
} csr_host;

typedef struct {
   const csr_host *A;
   const double *x;
   double *y;
   int64_t ldx, ldy, r0, r1;
   int bs;
} mv_job;

static void *mv_worker(void *arg) {
   mv_job *j = (mv_job *)arg;
   const csr_host *A = j->A;
   for (int c = 0; c < j->bs; c++) {
      const double *x = j->x + (size_t)c * j->ldx;
      double *y = j->y + (size_t)c * j->ldy;
      for (int64_t i = j->r0; i < j->r1; i++) {
         double t = 0.0;
         for (int64_t k = A->rowptr[i]; k < A->rowptr[i + 1]; k++) t += A->vals[k] * x[A->colind[k]];
         y[i] = t;
      }
   }
   return NULL;
}

void csr_host_matvec(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy, int *blockSize,
      primme_params *primme, int *ierr) {
   const csr_host *A = (const csr_host *)primme->matrix;
   int nt = A->nthreads > 0 ? A->nthreads : 1;
   if (nt > 256) nt = 256;
   if (A->n < 4096) nt = 1;
   mv_job jobs[256];
   pthread_t th[256];
   for (int t = 0; t < nt; t++) {
      jobs[t].A = A, jobs[t].x = (const double *)x, jobs[t].y = (double *)y;
      jobs[t].ldx = *ldx, jobs[t].ldy = *ldy, jobs[t].bs = *blockSize;
      jobs[t].r0 = A->n * t / nt, jobs[t].r1 = A->n * (t + 1) / nt;
   }
   for (int t = 1; t < nt; t++) pthread_create(&th[t], NULL, mv_worker, &jobs[t]);
   mv_worker(&jobs[0]);
   for (int t = 1; t < nt; t++) pthread_join(th[t], NULL);
   *ierr = 0;
}

void csr_host_jacobi(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy, int *blockSize,
      primme_params *primme, int *ierr) {
   const csr_host *A = (const csr_host *)primme->preconditioner;
   for (int c = 0; c < *blockSize; c++) {
      double sh = (A->use_shifts && primme->ShiftsForPreconditioner) ? primme->ShiftsForPreconditioner[c] : 0.0;
      const double *xc = (const double *)x + (size_t)c * *ldx;
      double *yc = (double *)y + (size_t)c * *ldy;
      for (int64_t i = 0; i < A->n; i++) {
         double d = A->diag[i] - sh;
         if (fabs(d) < A->minabs) d = d < 0 ? -A->minabs : A->minabs;
         yc[i] = xc[i] / d;
      }
   }
   *ierr = 0;
}

/* ---- complex Hermitian twins (zprimme reference runs): vals interleaved (re,im), diag real ---- */
#include <complex.h>
typedef double _Complex csr_zc;

static void *zmv_worker(void *arg) {
   mv_job *j = (mv_job *)arg;
   const csr_host *A = j->A;
   const csr_zc *vals = (const csr_zc *)A->vals;
   for (int c = 0; c < j->bs; c++) {
      const csr_zc *x = (const csr_zc *)j->x + (size_t)c * j->ldx;
      csr_zc *y = (csr_zc *)j->y + (size_t)c * j->ldy;
      for (int64_t i = j->r0; i < j->r1; i++) {
         csr_zc t = 0.0;
         for (int64_t k = A->rowptr[i]; k < A->rowptr[i + 1]; k++) t += vals[k] * x[A->colind[k]];
         y[i] = t;
      }
   }
   return NULL;
}

void csr_host_zmatvec(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy, int *blockSize,
      primme_params *primme, int *ierr) {
   const csr_host *A = (const csr_host *)primme->matrix;
   int nt = A->nthreads > 0 ? A->nthreads : 1;
   if (nt > 256) nt = 256;
   if (A->n < 4096) nt = 1;
   mv_job jobs[256];
   pthread_t th[256];
   for (int t = 0; t < nt; t++) {
      jobs[t].A = A, jobs[t].x = (const double *)x, jobs[t].y = (double *)y;
      jobs[t].ldx = *ldx, jobs[t].ldy = *ldy, jobs[t].bs = *blockSize;
      jobs[t].r0 = A->n * t / nt, jobs[t].r1 = A->n * (t + 1) / nt;
   }
   for (int t = 1; t < nt; t++) pthread_create(&th[t], NULL, zmv_worker, &jobs[t]);
   zmv_worker(&jobs[0]);
   for (int t = 1; t < nt; t++) pthread_join(th[t], NULL);
   *ierr = 0;
}

void csr_host_zjacobi(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy, int *blockSize,
      primme_params *primme, int *ierr) {
   const csr_host *A = (const csr_host *)primme->preconditioner;
   for (int c = 0; c < *blockSize; c++) {
      double sh = (A->use_shifts && primme->ShiftsForPreconditioner) ? primme->ShiftsForPreconditioner[c] : 0.0;
      const csr_zc *xc = (const csr_zc *)x + (size_t)c * *ldx;
      csr_zc *yc = (csr_zc *)y + (size_t)c * *ldy;
      for (int64_t i = 0; i < A->n; i++) {
         double d = A->diag[i] - sh;
         if (fabs(d) < A->minabs) d = d < 0 ? -A->minabs : A->minabs;
         yc[i] = xc[i] / d;
      }
   }
   *ierr = 0;
}

/* ---- rectangular operator with the SVD callback signature (reference examples/ex_svds_dseq.c:188-230,
 * tests/COMMON/mat.c CSRMatrixMatvecSVD): y = A x (transpose == 0) or y = A' x, host blocks ---- */
typedef struct csr_host_rect {
   int64_t m, n;
   const int64_t *rowptr; /* m + 1, 0-based */
   const int32_t *colind;
   const double *vals;
} csr_host_rect;

void csr_host_svds_matvec(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy, int *blockSize, int *transpose,
      primme_svds_params *primme_svds, int *ierr) {
   const csr_host_rect *A = (const csr_host_rect *)primme_svds->matrix;
   for (int c = 0; c < *blockSize; c++) {
      const double *xc = (const double *)x + (size_t)c * *ldx;
      double *yc = (double *)y + (size_t)c * *ldy;
      if (!*transpose) {
         for (int64_t i = 0; i < A->m; i++) {
            double t = 0.0;
            for (int64_t k = A->rowptr[i]; k < A->rowptr[i + 1]; k++) t += A->vals[k] * xc[A->colind[k]];
            yc[i] = t;
         }
      } else {
         for (int64_t j = 0; j < A->n; j++) yc[j] = 0.0;
         for (int64_t i = 0; i < A->m; i++)
            for (int64_t k = A->rowptr[i]; k < A->rowptr[i + 1]; k++) yc[A->colind[k]] += A->vals[k] * xc[i];
      }
   }
   *ierr = 0;
}

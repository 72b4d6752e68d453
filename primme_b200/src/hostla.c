/* hostla.c -- host-side small dense helpers (see hostla.h). */
#include "hostla.h"
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

/* Fortran BLAS/LAPACK (LP64) */
extern void dgemm_(const char *, const char *, const int *, const int *, const int *,
      const double *, const double *, const int *, const double *, const int *, const double *,
      double *, const int *);
extern void dsymm_(const char *, const char *, const int *, const int *, const double *,
      const double *, const int *, const double *, const int *, const double *, double *,
      const int *);
extern void dtrsm_(const char *, const char *, const char *, const char *, const int *,
      const int *, const double *, const double *, const int *, double *, const int *);
extern void dpotrf_(const char *, const int *, double *, const int *, int *);
extern void dgetrf_(const int *, const int *, double *, const int *, int *, int *);
extern void dgetrs_(const char *, const int *, const int *, const double *, const int *, const int *, double *,
      const int *, int *);
extern void dtrmm_(const char *, const char *, const char *, const char *, const int *,
      const int *, const double *, const double *, const int *, double *, const int *);
extern void dgesvd_(const char *, const char *, const int *, const int *, double *, const int *, double *,
      double *, const int *, double *, const int *, double *, const int *, int *);
extern void dsyevx_(const char *, const char *, const char *, const int *, double *, const int *,
      const double *, const double *, const int *, const int *, const double *, int *, double *,
      double *, const int *, double *, const int *, int *, int *, int *);
extern void dsygvx_(const int *, const char *, const char *, const char *, const int *, double *,
      const int *, double *, const int *, const double *, const double *, const int *,
      const int *, const double *, int *, double *, double *, const int *, double *,
      const int *, int *, int *, int *);
extern void dlarnv_(const int *, int *, const int *, double *);
extern double ddot_(const int *, const double *, const int *, const double *, const int *);

void hl_permute_cols(double *x, int m, int n, int ld, const int *perm) {
   if (n <= 0 || m <= 0) return;
   double *tmp = (double *)malloc(sizeof(double) * (size_t)m * n);
   for (int i = 0; i < n; i++) memcpy(tmp + (size_t)i * m, x + (size_t)perm[i] * ld, sizeof(double) * m);
   for (int i = 0; i < n; i++) memcpy(x + (size_t)i * ld, tmp + (size_t)i * m, sizeof(double) * m);
   free(tmp);
}

void hl_permute_ints(int *x, int n, const int *perm) {
   if (n <= 0) return;
   int *tmp = (int *)malloc(sizeof(int) * n);
   for (int i = 0; i < n; i++) tmp[i] = x[perm[i]];
   memcpy(x, tmp, sizeof(int) * n);
   free(tmp);
}

void hl_copy(const double *x, int m, int n, int ldx, double *y, int ldy) {
   if (m <= 0) return;
   for (int j = 0; j < n; j++) memmove(y + (size_t)j * ldy, x + (size_t)j * ldx, sizeof(double) * m);
}

void hl_zero(double *x, int m, int n, int ld) {
   if (m <= 0) return;
   for (int j = 0; j < n; j++) memset(x + (size_t)j * ld, 0, sizeof(double) * m);
}

void hl_gemm(char ta, char tb, int m, int n, int k, double alpha, const double *A, int lda,
      const double *B, int ldb, double beta, double *C, int ldc) {
   if (m == 0 || n == 0) return;
   if (k == 0) {
      for (int j = 0; j < n; j++)
         for (int i = 0; i < m; i++)
            C[i + (size_t)j * ldc] = beta == 0.0 ? 0.0 : beta * C[i + (size_t)j * ldc];
      return;
   }
   if (lda < 1) lda = 1;
   if (ldb < 1) ldb = 1;
   dgemm_(&ta, &tb, &m, &n, &k, &alpha, A, &lda, B, &ldb, &beta, C, &ldc);
}

void hl_symm_lu(int m, int n, double alpha, const double *A, int lda, const double *B, int ldb,
      double beta, double *C, int ldc) {
   if (m == 0 || n == 0) return;
   dsymm_("L", "U", &m, &n, &alpha, A, &lda, B, &ldb, &beta, C, &ldc);
}

void hl_trsm(char side, char uplo, char trans, char diag, int m, int n, double alpha,
      const double *A, int lda, double *B, int ldb) {
   if (m == 0 || n == 0) return;
   dtrsm_(&side, &uplo, &trans, &diag, &m, &n, &alpha, A, &lda, B, &ldb);
}

int hl_potrf_upper(int n, double *A, int lda) {
   int info = 0;
   if (n == 0) return 0;
   dpotrf_("U", &n, A, &lda, &info);
   return info;
}

int hl_sygv_upper(int n, double *A, int lda, const double *B, int ldb, double *w) {
   if (n == 0) return 0;
   int info = 0, lwork = -1, nfound = 0, idum = 0, one = 1;
   double abstol = 0.0, rdum = 0.0, wq = 0.0;
   double *z = (double *)malloc(sizeof(double) * (size_t)n * n);
   int *iwork = (int *)malloc(sizeof(int) * 5 * n);
   int *ifail = (int *)malloc(sizeof(int) * n);
   double *b = NULL;
   if (B) {
      /* only the upper triangle is meaningful in the caller's array */
      b = (double *)calloc((size_t)n * n, sizeof(double));
      for (int j = 0; j < n; j++)
         for (int i = 0; i <= j; i++) b[i + (size_t)j * n] = B[i + (size_t)j * ldb];
      dsygvx_(&one, "V", "A", "U", &n, A, &lda, b, &n, &rdum, &rdum, &idum, &idum, &abstol,
            &nfound, w, z, &n, &wq, &lwork, iwork, ifail, &info);
   } else {
      dsyevx_("V", "A", "U", &n, A, &lda, &rdum, &rdum, &idum, &idum, &abstol, &nfound, w, z, &n,
            &wq, &lwork, iwork, ifail, &info);
   }
   if (info == 0) {
      /* same workspace rule as the reference (blaslapack.c:1040-1043,1198): the optimal size
         from the query, so LAPACK takes the same blocked/unblocked code path */
      lwork = (int)wq;
      if (!B && lwork < 2 * n) lwork = 2 * n;
      if (lwork < 1) lwork = 1;
      double *work = (double *)malloc(sizeof(double) * lwork);
      if (B)
         dsygvx_(&one, "V", "A", "U", &n, A, &lda, b, &n, &rdum, &rdum, &idum, &idum, &abstol,
               &nfound, w, z, &n, work, &lwork, iwork, ifail, &info);
      else
         dsyevx_("V", "A", "U", &n, A, &lda, &rdum, &rdum, &idum, &idum, &abstol, &nfound, w, z,
               &n, work, &lwork, iwork, ifail, &info);
      free(work);
   }
   if (info == 0) hl_copy(z, n, n, n, A, lda);
   free(z), free(iwork), free(ifail), free(b);
   return info;
}

int hl_getrf(int m, int n, double *A, int lda, int *ipiv) {
   int info = 0;
   if (m == 0 || n == 0) return 0;
   dgetrf_(&m, &n, A, &lda, ipiv, &info);
   return info;
}

int hl_getrs(char trans, int n, int nrhs, const double *A, int lda, const int *ipiv, double *B, int ldb) {
   int info = 0;
   if (n == 0 || nrhs == 0) return 0;
   dgetrs_(&trans, &n, &nrhs, A, &lda, ipiv, B, &ldb, &info);
   return info;
}

void hl_trmm(char side, char uplo, char trans, char diag, int m, int n, double alpha, const double *A,
      int lda, double *B, int ldb) {
   if (m == 0 || n == 0) return;
   dtrmm_(&side, &uplo, &trans, &diag, &m, &n, &alpha, A, &lda, B, &ldb);
}

/* dgesvd jobu = 'S', jobvt = 'O' (reference Num_gesvd, blaslapack.c:1350-1403, as called by
 * solve_H_Ref): on return U holds the left singular vectors, A the transposed right ones, s the
 * singular values in descending order.  Workspace from the query, like the reference. */
int hl_gesvd_SO(int m, int n, double *A, int lda, double *s, double *U, int ldu) {
   if (m == 0 || n == 0) return 0;
   int info = 0, lwork = -1;
   double wq = 0.0;
   dgesvd_("S", "O", &m, &n, A, &lda, s, U, &ldu, A, &lda, &wq, &lwork, &info);
   if (info == 0) {
      lwork = (int)wq;
      if (lwork < 1) lwork = 1;
      double *work = (double *)malloc(sizeof(double) * lwork);
      dgesvd_("S", "O", &m, &n, A, &lda, s, U, &ldu, A, &lda, work, &lwork, &info);
      free(work);
   }
   return info;
}

void hl_larnv2(long long iseed[4], long long n, double *x) {
   int idist = 2, seed[4];
   for (int i = 0; i < 4; i++) seed[i] = (int)iseed[i];
   while (n > 0) {
      int chunk = n > 0x7ffffff0LL ? 0x7ffffff0 : (int)n;
      dlarnv_(&idist, seed, &chunk, x);
      x += chunk, n -= chunk;
   }
   for (int i = 0; i < 4; i++) iseed[i] = seed[i];
}

double hl_dot(int n, const double *x, const double *y) {
   /* BLAS ddot, as the reference's Num_dot (blaslapack.c:923) */
   int one = 1;
   if (n <= 0) return 0.0;
   return ddot_(&n, x, &one, y, &one);
}

double hl_wtime(void) {
   struct timeval tv;
   gettimeofday(&tv, NULL);
   return (double)tv.tv_sec + (double)tv.tv_usec / 1e6;
}

/* The projected problems are tiny (<= 64 x 64): a threaded BLAS spends far more time waking its
 * pool than computing.  If the linked BLAS is OpenBLAS, run it single-threaded during a solve. */
extern void openblas_set_num_threads(int) __attribute__((weak));
extern int openblas_get_num_threads(void) __attribute__((weak));
int hl_blas_threads(int nthreads) {
   int prev = openblas_get_num_threads ? openblas_get_num_threads() : 0;
   if (openblas_set_num_threads && nthreads > 0) openblas_set_num_threads(nthreads);
   return prev;
}

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
extern void dsytrf_(const char *, const int *, double *, const int *, int *, double *, const int *, int *);
extern void dsytrs_(const char *, const int *, const int *, const double *, const int *, const int *, double *, const int *, int *);
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
#ifdef PB_COMPLEX
#include <math.h>
extern void zgemm_(const char *, const char *, const int *, const int *, const int *, const SCALAR *, const SCALAR *,
      const int *, const SCALAR *, const int *, const SCALAR *, SCALAR *, const int *);
extern void zhemm_(const char *, const char *, const int *, const int *, const SCALAR *, const SCALAR *, const int *,
      const SCALAR *, const int *, const SCALAR *, SCALAR *, const int *);
extern void ztrsm_(const char *, const char *, const char *, const char *, const int *, const int *, const SCALAR *,
      const SCALAR *, const int *, SCALAR *, const int *);
extern void ztrmm_(const char *, const char *, const char *, const char *, const int *, const int *, const SCALAR *,
      const SCALAR *, const int *, SCALAR *, const int *);
extern void zpotrf_(const char *, const int *, SCALAR *, const int *, int *);
extern void zgetrf_(const int *, const int *, SCALAR *, const int *, int *, int *);
extern void zhetrf_(const char *, const int *, SCALAR *, const int *, int *, SCALAR *, const int *, int *);
extern void zhetrs_(const char *, const int *, const int *, const SCALAR *, const int *, const int *, SCALAR *, const int *, int *);
extern void zgetrs_(const char *, const int *, const int *, const SCALAR *, const int *, const int *, SCALAR *,
      const int *, int *);
extern void zgesvd_(const char *, const char *, const int *, const int *, SCALAR *, const int *, double *, SCALAR *,
      const int *, SCALAR *, const int *, SCALAR *, const int *, double *, int *);
extern void zheevx_(const char *, const char *, const char *, const int *, SCALAR *, const int *, const double *,
      const double *, const int *, const int *, const double *, int *, double *, SCALAR *, const int *, SCALAR *,
      const int *, double *, int *, int *, int *);
extern void zhegvx_(const int *, const char *, const char *, const char *, const int *, SCALAR *, const int *, SCALAR *,
      const int *, const double *, const double *, const int *, const int *, const double *, int *, double *,
      SCALAR *, const int *, SCALAR *, const int *, double *, int *, int *, int *);
#define XGEMM zgemm_
#define XSYMM zhemm_
#define XTRSM ztrsm_
#define XTRMM ztrmm_
#define XPOTRF zpotrf_
#define XGETRF zgetrf_
#define XHETRF zhetrf_
#define XHETRS zhetrs_
#define XGETRS zgetrs_
#else
#define XGEMM dgemm_
#define XSYMM dsymm_
#define XTRSM dtrsm_
#define XTRMM dtrmm_
#define XPOTRF dpotrf_
#define XGETRF dgetrf_
#define XHETRF dsytrf_
#define XHETRS dsytrs_
#define XGETRS dgetrs_
#endif

void hl_permute_cols(SCALAR *x, int m, int n, int ld, const int *perm) {
   if (n <= 0 || m <= 0) return;
   SCALAR *tmp = (SCALAR *)malloc(sizeof(SCALAR) * (size_t)m * n);
   for (int i = 0; i < n; i++) memcpy(tmp + (size_t)i * m, x + (size_t)perm[i] * ld, sizeof(SCALAR) * m);
   for (int i = 0; i < n; i++) memcpy(x + (size_t)i * ld, tmp + (size_t)i * m, sizeof(SCALAR) * m);
   free(tmp);
}

#ifndef PB_COMPLEX
void hl_permute_reals(double *x, int n, const int *perm) {
   if (n <= 0) return;
   double *tmp = (double *)malloc(sizeof(double) * n);
   for (int i = 0; i < n; i++) tmp[i] = x[perm[i]];
   memcpy(x, tmp, sizeof(double) * n);
   free(tmp);
}

void hl_permute_ints(int *x, int n, const int *perm) {
   if (n <= 0) return;
   int *tmp = (int *)malloc(sizeof(int) * n);
   for (int i = 0; i < n; i++) tmp[i] = x[perm[i]];
   memcpy(x, tmp, sizeof(int) * n);
   free(tmp);
}
#endif

void hl_copy(const SCALAR *x, int m, int n, int ldx, SCALAR *y, int ldy) {
   if (m <= 0) return;
   for (int j = 0; j < n; j++) memmove(y + (size_t)j * ldy, x + (size_t)j * ldx, sizeof(SCALAR) * m);
}

void hl_zero(SCALAR *x, int m, int n, int ld) {
   if (m <= 0) return;
   for (int j = 0; j < n; j++) memset(x + (size_t)j * ld, 0, sizeof(SCALAR) * m);
}

void hl_gemm(char ta, char tb, int m, int n, int k, SCALAR alpha, const SCALAR *A, int lda,
      const SCALAR *B, int ldb, SCALAR beta, SCALAR *C, int ldc) {
   if (m == 0 || n == 0) return;
   if (k == 0) {
      for (int j = 0; j < n; j++)
         for (int i = 0; i < m; i++)
            C[i + (size_t)j * ldc] = beta == 0.0 ? 0.0 : beta * C[i + (size_t)j * ldc];
      return;
   }
   if (lda < 1) lda = 1;
   if (ldb < 1) ldb = 1;
   XGEMM(&ta, &tb, &m, &n, &k, &alpha, A, &lda, B, &ldb, &beta, C, &ldc);
}

void hl_symm_lu(int m, int n, SCALAR alpha, const SCALAR *A, int lda, const SCALAR *B, int ldb,
      SCALAR beta, SCALAR *C, int ldc) {
   if (m == 0 || n == 0) return;
   XSYMM("L", "U", &m, &n, &alpha, A, &lda, B, &ldb, &beta, C, &ldc);
}

void hl_trsm(char side, char uplo, char trans, char diag, int m, int n, SCALAR alpha,
      const SCALAR *A, int lda, SCALAR *B, int ldb) {
   if (m == 0 || n == 0) return;
   XTRSM(&side, &uplo, &trans, &diag, &m, &n, &alpha, A, &lda, B, &ldb);
}

int hl_potrf_upper(int n, SCALAR *A, int lda) {
   int info = 0;
   if (n == 0) return 0;
   XPOTRF("U", &n, A, &lda, &info);
   return info;
}

#ifndef PB_COMPLEX
double hl_prof_eig_s = 0.0; /* PB200_HOST_PROFILE: time inside the projected eigen-solves */
long hl_prof_eig_n = 0;
#else
extern double hl_prof_eig_s;
extern long hl_prof_eig_n;
#endif
static int hl_sygv_upper_(int n, SCALAR *A, int lda, const SCALAR *B, int ldb, double *w);
int hl_sygv_upper(int n, SCALAR *A, int lda, const SCALAR *B, int ldb, double *w) {
   const double t0 = hl_wtime();
   const int rc = hl_sygv_upper_(n, A, lda, B, ldb, w);
   hl_prof_eig_s += hl_wtime() - t0, hl_prof_eig_n++;
   return rc;
}
static int hl_sygv_upper_(int n, SCALAR *A, int lda, const SCALAR *B, int ldb, double *w) {
   if (n == 0) return 0;
   int info = 0, lwork = -1, nfound = 0, idum = 0, one = 1;
   double abstol = 0.0, rdum = 0.0;
   SCALAR wq = 0.0;
   SCALAR *z = (SCALAR *)malloc(sizeof(SCALAR) * (size_t)n * n);
   int *iwork = (int *)malloc(sizeof(int) * 5 * n);
   int *ifail = (int *)malloc(sizeof(int) * n);
   SCALAR *b = NULL;
#ifdef PB_COMPLEX
   double *rwork = (double *)malloc(sizeof(double) * 7 * n); /* blaslapack.c:1045,1187 */
#define HL_RWORK rwork,
#define XHEEVX zheevx_
#define XHEGVX zhegvx_
#else
#define HL_RWORK
#define XHEEVX dsyevx_
#define XHEGVX dsygvx_
#endif
   if (B) {
      /* only the upper triangle is meaningful in the caller's array */
      b = (SCALAR *)calloc((size_t)n * n, sizeof(SCALAR));
      for (int j = 0; j < n; j++)
         for (int i = 0; i <= j; i++) b[i + (size_t)j * n] = B[i + (size_t)j * ldb];
      XHEGVX(&one, "V", "A", "U", &n, A, &lda, b, &n, &rdum, &rdum, &idum, &idum, &abstol,
            &nfound, w, z, &n, &wq, &lwork, HL_RWORK iwork, ifail, &info);
   } else {
      XHEEVX("V", "A", "U", &n, A, &lda, &rdum, &rdum, &idum, &idum, &abstol, &nfound, w, z, &n,
            &wq, &lwork, HL_RWORK iwork, ifail, &info);
   }
   if (info == 0) {
      /* same workspace rule as the reference (blaslapack.c:1040-1043,1198): the optimal size
         from the query, so LAPACK takes the same blocked/unblocked code path */
      lwork = (int)PB_REAL(wq);
      if (!B && lwork < 2 * n) lwork = 2 * n;
      if (lwork < 1) lwork = 1;
      SCALAR *work = (SCALAR *)malloc(sizeof(SCALAR) * lwork);
      if (B)
         XHEGVX(&one, "V", "A", "U", &n, A, &lda, b, &n, &rdum, &rdum, &idum, &idum, &abstol,
               &nfound, w, z, &n, work, &lwork, HL_RWORK iwork, ifail, &info);
      else
         XHEEVX("V", "A", "U", &n, A, &lda, &rdum, &rdum, &idum, &idum, &abstol, &nfound, w, z,
               &n, work, &lwork, HL_RWORK iwork, ifail, &info);
      free(work);
   }
   if (info == 0) hl_copy(z, n, n, n, A, lda);
   free(z), free(iwork), free(ifail), free(b);
#ifdef PB_COMPLEX
   free(rwork);
#endif
   return info;
}

int hl_getrf(int m, int n, SCALAR *A, int lda, int *ipiv) {
   int info = 0;
   if (m == 0 || n == 0) return 0;
   XGETRF(&m, &n, A, &lda, ipiv, &info);
   return info;
}

int hl_getrs(char trans, int n, int nrhs, const SCALAR *A, int lda, const int *ipiv, SCALAR *B, int ldb) {
   int info = 0;
   if (n == 0 || nrhs == 0) return 0;
   XGETRS(&trans, &n, &nrhs, A, &lda, ipiv, B, &ldb, &info);
   return info;
}

int hl_hetrf_upper(int n, SCALAR *A, int lda, int *ipiv) {
   int info = 0, lwork = -1;
   if (n == 0) return 0;
   SCALAR wq = 0.0;
   XHETRF("U", &n, A, &lda, ipiv, &wq, &lwork, &info); /* optimal workspace, like the reference */
   if (info == 0) {
      lwork = (int)PB_REAL(wq);
      if (lwork < 1) lwork = 1;
      SCALAR *work = (SCALAR *)malloc(sizeof(SCALAR) * lwork);
      XHETRF("U", &n, A, &lda, ipiv, work, &lwork, &info);
      free(work);
   }
   return info;
}

int hl_hetrs_upper(int n, int nrhs, const SCALAR *A, int lda, const int *ipiv, SCALAR *B, int ldb) {
   int info = 0;
   if (n == 0 || nrhs == 0) return 0;
   XHETRS("U", &n, &nrhs, A, &lda, ipiv, B, &ldb, &info);
   return info;
}

void hl_trmm(char side, char uplo, char trans, char diag, int m, int n, SCALAR alpha, const SCALAR *A,
      int lda, SCALAR *B, int ldb) {
   if (m == 0 || n == 0) return;
   XTRMM(&side, &uplo, &trans, &diag, &m, &n, &alpha, A, &lda, B, &ldb);
}

/* dgesvd jobu = 'S', jobvt = 'O' (reference Num_gesvd, blaslapack.c:1350-1403, as called by
 * solve_H_Ref): on return U holds the left singular vectors, A the transposed right ones, s the
 * singular values in descending order.  Workspace from the query, like the reference. */
int hl_gesvd_SO(int m, int n, SCALAR *A, int lda, double *s, SCALAR *U, int ldu) {
   if (m == 0 || n == 0) return 0;
   int info = 0, lwork = -1;
   SCALAR wq = 0.0;
#ifdef PB_COMPLEX
   double *rwork = (double *)malloc(sizeof(double) * 5 * (m < n ? m : n));
   zgesvd_("S", "O", &m, &n, A, &lda, s, U, &ldu, A, &lda, &wq, &lwork, rwork, &info);
#else
   dgesvd_("S", "O", &m, &n, A, &lda, s, U, &ldu, A, &lda, &wq, &lwork, &info);
#endif
   if (info == 0) {
      lwork = (int)PB_REAL(wq);
      if (lwork < 1) lwork = 1;
      SCALAR *work = (SCALAR *)malloc(sizeof(SCALAR) * lwork);
#ifdef PB_COMPLEX
      zgesvd_("S", "O", &m, &n, A, &lda, s, U, &ldu, A, &lda, work, &lwork, rwork, &info);
#else
      dgesvd_("S", "O", &m, &n, A, &lda, s, U, &ldu, A, &lda, work, &lwork, &info);
#endif
      free(work);
   }
#ifdef PB_COMPLEX
   free(rwork);
#endif
   return info;
}

void hl_larnv2(long long iseed[4], long long n, SCALAR *x_) {
   int idist = 2, seed[4];
   double *x = (double *)x_;
   n *= (long long)(sizeof(SCALAR) / sizeof(double));
   for (int i = 0; i < 4; i++) seed[i] = (int)iseed[i];
   while (n > 0) {
      int chunk = n > 0x7ffffff0LL ? 0x7ffffff0 : (int)n;
      dlarnv_(&idist, seed, &chunk, x);
      x += chunk, n -= chunk;
   }
   for (int i = 0; i < 4; i++) iseed[i] = seed[i];
}

SCALAR hl_dot(int n, const SCALAR *x, const SCALAR *y) {
   if (n <= 0) return 0.0;
#ifdef PB_COMPLEX
   SCALAR s = 0.0; /* the reference's explicit zdotc loop (blaslapack.c:899-913) */
   for (int i = 0; i < n; i++) s += conj(x[i]) * y[i];
   return s;
#else
   /* BLAS ddot, as the reference's Num_dot (blaslapack.c:923) */
   int one = 1;
   return ddot_(&n, x, &one, y, &one);
#endif
}

#ifndef PB_COMPLEX
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
#endif

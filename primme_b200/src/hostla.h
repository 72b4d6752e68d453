/* hostla.h -- small dense host linear algebra used by the Davidson control code.
 *
 * The projected problem (<= maxBasisSize ~ 40-64) is solved on the host with LAPACK exactly as
 * the reference does (reference src/linalg/blaslapack.c:1058-1235: ?syevx / ?sygvx, range 'A',
 * abstol 0; potrf, trsm, gemm, larnv).  Only Fortran-ABI BLAS/LAPACK symbols are used
 * (dgemm_, dsygvx_, ...), any LP64 implementation works; the build links the OpenBLAS found on
 * the box (see primme_b200/build.py).
 */
#ifndef PB_HOSTLA_H
#define PB_HOSTLA_H

#include <stddef.h>

#define PB_EPS 2.220446049250313e-16 /* DBL_EPSILON, reference MACHINE_EPSILON (common.h:157) */

#ifndef PB_MIN
#define PB_MIN(a, b) ((a) < (b) ? (a) : (b))
#define PB_MAX(a, b) ((a) > (b) ? (a) : (b))
#endif

/* y(:,i) <- x(:,perm[i]) in place, for i < n (reference permute_vecs, auxiliary.c:716-793) */
void hl_permute_cols(double *x, int m, int n, int ld, const int *perm);
void hl_permute_ints(int *x, int n, const int *perm);
void hl_copy(const double *x, int m, int n, int ldx, double *y, int ldy);
void hl_zero(double *x, int m, int n, int ld);

/* C = alpha*op(A)*op(B) + beta*C */
void hl_gemm(char ta, char tb, int m, int n, int k, double alpha, const double *A, int lda,
      const double *B, int ldb, double beta, double *C, int ldc);
/* C = alpha*A*B + beta*C with A symmetric (upper stored), side L */
void hl_symm_lu(int m, int n, double alpha, const double *A, int lda, const double *B, int ldb,
      double beta, double *C, int ldc);
void hl_trsm(char side, char uplo, char trans, char diag, int m, int n, double alpha,
      const double *A, int lda, double *B, int ldb);
void hl_trmm(char side, char uplo, char trans, char diag, int m, int n, double alpha,
      const double *A, int lda, double *B, int ldb);
int hl_potrf_upper(int n, double *A, int lda); /* returns LAPACK info */
/* LU with partial pivoting and its solves (reference Num_getrf / Num_getrs, blaslapack.c) */
int hl_getrf(int m, int n, double *A, int lda, int *ipiv);
int hl_getrs(char trans, int n, int nrhs, const double *A, int lda, const int *ipiv, double *B, int ldb);
/* singular value decomposition, left vectors in U, transposed right vectors overwrite A */
int hl_gesvd_SO(int m, int n, double *A, int lda, double *s, double *U, int ldu);
/* eigen-decomposition of the symmetric matrix stored in the upper triangle of A (n x n, lda);
 * on return A holds the eigenvectors, w ascending eigenvalues.  B != NULL: generalized problem
 * A x = w B x (upper triangle of B referenced, B not modified).  Returns LAPACK info. */
int hl_sygv_upper(int n, double *A, int lda, const double *B, int ldb, double *w);
/* uniform(-1,1) numbers from LAPACK's dlarnv(idist=2) with the evolving 4-integer seed */
void hl_larnv2(long long iseed[4], long long n, double *x);
double hl_dot(int n, const double *x, const double *y);
double hl_wtime(void);
/* set the BLAS thread count (OpenBLAS only); returns the previous value or 0 */
int hl_blas_threads(int nthreads);

#endif

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

/* Scalar type of this compilation unit.  Every typed host source is compiled twice -- plain for
 * dprimme, with -DPB_COMPLEX for zprimme -- the way the reference instantiates its templates by
 * self-inclusion (reference src/include/template.h, template_types.h:51-204): SCALAR is the type of
 * the basis and of the projected matrices, `double` stays the type of Ritz values, norms and
 * tolerances; typed symbols get the suffix _d or _z (PB_SUF). */
#ifdef PB_COMPLEX
#include <complex.h>
typedef double _Complex SCALAR;
#define PB_SUF(x) x##_z
#define PB_CONJ(x) conj(x)
#define PB_REAL(x) creal(x)
#define PB_ABS(x) cabs(x)
#define PB_ES 16
#else
typedef double SCALAR;
#define PB_SUF(x) x##_d
#define PB_CONJ(x) (x)
#define PB_REAL(x) (x)
#define PB_ABS(x) fabs(x)
#define PB_ES 8
#endif
#define hl_permute_cols PB_SUF(hl_permute_cols)
#define hl_copy PB_SUF(hl_copy)
#define hl_zero PB_SUF(hl_zero)
#define hl_gemm PB_SUF(hl_gemm)
#define hl_symm_lu PB_SUF(hl_symm_lu)
#define hl_trsm PB_SUF(hl_trsm)
#define hl_trmm PB_SUF(hl_trmm)
#define hl_potrf_upper PB_SUF(hl_potrf_upper)
#define hl_getrf PB_SUF(hl_getrf)
#define hl_getrs PB_SUF(hl_getrs)
#define hl_hetrf_upper PB_SUF(hl_hetrf_upper)
#define hl_hetrs_upper PB_SUF(hl_hetrs_upper)
#define hl_gesvd_SO PB_SUF(hl_gesvd_SO)
#define hl_sygv_upper PB_SUF(hl_sygv_upper)
#define hl_larnv2 PB_SUF(hl_larnv2)
#define hl_dot PB_SUF(hl_dot)

#define PB_EPS 2.220446049250313e-16 /* DBL_EPSILON, reference MACHINE_EPSILON (common.h:157) */

#ifndef PB_MIN
#define PB_MIN(a, b) ((a) < (b) ? (a) : (b))
#define PB_MAX(a, b) ((a) > (b) ? (a) : (b))
#endif

/* y(:,i) <- x(:,perm[i]) in place, for i < n (reference permute_vecs, auxiliary.c:716-793) */
void hl_permute_cols(SCALAR *x, int m, int n, int ld, const int *perm);
void hl_permute_ints(int *x, int n, const int *perm);
void hl_permute_reals(double *x, int n, const int *perm);
void hl_copy(const SCALAR *x, int m, int n, int ldx, SCALAR *y, int ldy);
void hl_zero(SCALAR *x, int m, int n, int ld);

/* C = alpha*op(A)*op(B) + beta*C ('C' = conjugate transpose) */
void hl_gemm(char ta, char tb, int m, int n, int k, SCALAR alpha, const SCALAR *A, int lda,
      const SCALAR *B, int ldb, SCALAR beta, SCALAR *C, int ldc);
/* C = alpha*A*B + beta*C with A symmetric / Hermitian (upper stored), side L */
void hl_symm_lu(int m, int n, SCALAR alpha, const SCALAR *A, int lda, const SCALAR *B, int ldb,
      SCALAR beta, SCALAR *C, int ldc);
void hl_trsm(char side, char uplo, char trans, char diag, int m, int n, SCALAR alpha,
      const SCALAR *A, int lda, SCALAR *B, int ldb);
void hl_trmm(char side, char uplo, char trans, char diag, int m, int n, SCALAR alpha,
      const SCALAR *A, int lda, SCALAR *B, int ldb);
int hl_potrf_upper(int n, SCALAR *A, int lda); /* returns LAPACK info */
/* LU with partial pivoting and its solves (reference Num_getrf / Num_getrs, blaslapack.c) */
int hl_getrf(int m, int n, SCALAR *A, int lda, int *ipiv);
int hl_getrs(char trans, int n, int nrhs, const SCALAR *A, int lda, const int *ipiv, SCALAR *B, int ldb);
/* Bunch-Kaufman factorisation / solve of a Hermitian matrix given by its upper triangle: dsytrf / zhetrf with the
 * workspace from the query, dsytrs / zhetrs (reference Num_hetrf / Num_hetrs, blaslapack.c:1412-1560) */
int hl_hetrf_upper(int n, SCALAR *A, int lda, int *ipiv);
int hl_hetrs_upper(int n, int nrhs, const SCALAR *A, int lda, const int *ipiv, SCALAR *B, int ldb);
/* singular value decomposition, left vectors in U, transposed right vectors overwrite A */
int hl_gesvd_SO(int m, int n, SCALAR *A, int lda, double *s, SCALAR *U, int ldu);
/* eigen-decomposition of the symmetric / Hermitian matrix stored in the upper triangle of A (n x n, lda);
 * on return A holds the eigenvectors, w ascending eigenvalues.  B != NULL: generalized problem
 * A x = w B x (upper triangle of B referenced, B not modified).  Returns LAPACK info. */
int hl_sygv_upper(int n, SCALAR *A, int lda, const SCALAR *B, int ldb, double *w);
/* uniform(-1,1) numbers from LAPACK's dlarnv(idist=2) with the evolving 4-integer seed; complex: real
 * and imaginary parts from the real generator on 2n numbers (reference blaslapack.c:938-949) */
void hl_larnv2(long long iseed[4], long long n, SCALAR *x);
/* x^H y (complex: the reference's explicit loop, blaslapack.c:899-913) */
SCALAR hl_dot(int n, const SCALAR *x, const SCALAR *y);
double hl_wtime(void);
/* set the BLAS thread count (OpenBLAS only); returns the previous value or 0 */
int hl_blas_threads(int nthreads);

#endif

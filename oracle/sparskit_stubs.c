/* sparskit_stubs.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 * C stand-ins for the Fortran SPARSKIT entry points the reference's test driver declares in
 * tests/COMMON/mat.c:43-55 (objects COMMON/matvec.f, ilut.f of tests/Makefile:26-30); there is no
 * Fortran compiler in this image.  CSR arrays are 1-based, as the driver's reader produces them
 * (tests/COMMON/csr.c:157-163). */
#include <stdio.h>
#include <stdlib.h>

/* y = A x */
void amux_(int *n, double *x, double *y, double *a, int *ja, int *ia) {
   for (int i = 0; i < *n; i++) {
      double s = 0.0;
      for (int k = ia[i]; k < ia[i + 1]; k++) s += a[k - 1] * x[ja[k - 1] - 1];
      y[i] = s;
   }
}
/* y = A' x for an m x n matrix */
void atmuxr_(int *m, int *n, double *x, double *y, double *a, int *ja, int *ia) {
   for (int j = 0; j < *m; j++) y[j] = 0.0;
   for (int i = 0; i < *n; i++)
      for (int k = ia[i]; k < ia[i + 1]; k++) y[ja[k - 1] - 1] += a[k - 1] * x[i];
}
/* ILUT preconditioner: no shipped test configuration selects it (driver.PrecChoice = ilut) */
void ilut_(void) {
   fprintf(stderr, "ilut_: not available in this build (no Fortran SPARSKIT)\n");
   abort();
}
void lusol0_(void) {
   fprintf(stderr, "lusol0_: not available in this build (no Fortran SPARSKIT)\n");
   abort();
}

/* complex twins (driver built with -DUSE_DOUBLECOMPLEX): y = A x; y = A^T x (zatmuxr does NOT conjugate,
 * like SPARSKIT's) */
#include <complex.h>
typedef double _Complex sk_z;
void zamux_(int *n, sk_z *x, sk_z *y, sk_z *a, int *ja, int *ia) {
   for (int i = 0; i < *n; i++) {
      sk_z s = 0.0;
      for (int k = ia[i]; k < ia[i + 1]; k++) s += a[k - 1] * x[ja[k - 1] - 1];
      y[i] = s;
   }
}
void zatmuxr_(int *m, int *n, sk_z *x, sk_z *y, sk_z *a, int *ja, int *ia) {
   for (int j = 0; j < *m; j++) y[j] = 0.0;
   for (int i = 0; i < *n; i++)
      for (int k = ia[i]; k < ia[i + 1]; k++) y[ja[k - 1] - 1] += a[k - 1] * x[i];
}
void zilut_(void) {
   fprintf(stderr, "zilut_: not available in this build (no Fortran SPARSKIT)\n");
   abort();
}
void zlusol_(void) {
   fprintf(stderr, "zlusol_: not available in this build (no Fortran SPARSKIT)\n");
   abort();
}

/* primme_ref_internal.h -- reference-INTERNAL entry points exported for the reference's own test
 * driver (tests/driver.c, tests/COMMON/ioandtest.c), which includes the reference's private
 * headers and links these names (SURVEY 8b "extra symbols").  Applications never need this file:
 * the driver sees the prototypes through the reference's src/include/blaslapack.h:1513,1884,1976,
 * src/eigs/ortho.h:617 and src/eigs/auxiliary_eigs.h:126,217; here they are restated with the
 * mirrored context type so that the boundary is documented and testable.  Host arrays only. */
#ifndef PRIMME_REF_INTERNAL_H
#define PRIMME_REF_INTERNAL_H
#include <stdio.h>
#include "primme.h"
#ifdef __cplusplus
extern "C" {
#endif

struct primme_context_str;
typedef struct primme_frame_mirror {
   void *prev_alloc;
   int keep_frame;
   struct primme_frame_mirror *prev;
} primme_frame_mirror;

/* field-by-field mirror of primme_context, src/include/common.h:610-641 (no PRIMME_PROFILE) */
typedef struct primme_context_str {
   primme_params *primme;
   primme_svds_params *primme_svds;
   int printLevel;
   FILE *outputFile;
   int (*report)(const char *fun, double time, struct primme_context_str ctx);
   primme_frame_mirror *mm;
   int numProcs;
   int procID;
   void *mpicomm;
   int (*bcast)(void *buffer, primme_op_datatype buffer_type, int count, struct primme_context_str ctx);
   int (*globalSum)(void *buffer, primme_op_datatype buffer_type, int count, struct primme_context_str ctx);
   void *queue;
} primme_context_mirror;

primme_context_mirror primme_get_context(primme_params *primme);
void primme_free_context(primme_context_mirror ctx);
int Mem_pop_frame(primme_context_mirror *ctx);
int Mem_pop_clean_frame(primme_context_mirror ctx);
int Mem_debug_frame(const char *where, primme_context_mirror ctx);
double Num_dot_dprimme(PRIMME_INT n, double *x, PRIMME_INT incx, double *y, PRIMME_INT incy,
      primme_context_mirror ctx);
int Num_gemv_dprimme(const char *transa, PRIMME_INT m, PRIMME_INT n, double alpha, double *a,
      PRIMME_INT lda, double *x, PRIMME_INT incx, double beta, double *y, PRIMME_INT incy,
      primme_context_mirror ctx);
int Num_larnv_dprimme(int idist, PRIMME_INT *iseed, PRIMME_INT length, double *x,
      primme_context_mirror ctx);
int ortho_single_iteration_dprimme(double *Q, int nQ, PRIMME_INT ldQ, double *BQ, PRIMME_INT ldBQ,
      double *QtBQ, int ldQtBQ, double *X, int *inX, int nX, PRIMME_INT ldX, double *norms,
      primme_context_mirror ctx);
/* complex double flavours (the driver built with -DUSE_DOUBLECOMPLEX links these) */
PRIMME_COMPLEX_DOUBLE Num_dot_zprimme(PRIMME_INT n, PRIMME_COMPLEX_DOUBLE *x, PRIMME_INT incx,
      PRIMME_COMPLEX_DOUBLE *y, PRIMME_INT incy, primme_context_mirror ctx);
int Num_gemv_zprimme(const char *transa, PRIMME_INT m, PRIMME_INT n, PRIMME_COMPLEX_DOUBLE alpha,
      PRIMME_COMPLEX_DOUBLE *a, PRIMME_INT lda, PRIMME_COMPLEX_DOUBLE *x, PRIMME_INT incx,
      PRIMME_COMPLEX_DOUBLE beta, PRIMME_COMPLEX_DOUBLE *y, PRIMME_INT incy, primme_context_mirror ctx);
int Num_larnv_zprimme(int idist, PRIMME_INT *iseed, PRIMME_INT length, PRIMME_COMPLEX_DOUBLE *x,
      primme_context_mirror ctx);
int ortho_single_iteration_zprimme(PRIMME_COMPLEX_DOUBLE *Q, int nQ, PRIMME_INT ldQ, PRIMME_COMPLEX_DOUBLE *BQ,
      PRIMME_INT ldBQ, PRIMME_COMPLEX_DOUBLE *QtBQ, int ldQtBQ, PRIMME_COMPLEX_DOUBLE *X, int *inX, int nX,
      PRIMME_INT ldX, double *norms, primme_context_mirror ctx);
/* layout of the mirrored context as this library was compiled (tests/test_abi.py) */
int primme_b200_ref_context_size(void);
int primme_b200_ref_context_offset(int field);

#ifdef __cplusplus
}
#endif
#endif

/* ref_internal.c -- the INTERNAL reference symbols that the reference's own test driver links
 * against, so that tests/driver.c and the tests/COMMON sources compile and link UNCHANGED against this
 * library (last row).  The driver includes the reference's private headers
 * (tests/COMMON/num.h:32-34 -> src/include/common.h, template.h, blaslapack.h;
 * tests/COMMON/ioandtest.c:37-38 -> src/eigs/auxiliary_eigs.h, ortho.h) and calls, in the double
 * precision host flavour:
 *
 *    primme_get_context / primme_free_context   src/eigs/auxiliary_eigs.c:94-160
 *    Mem_pop_frame / Mem_pop_clean_frame        src/linalg/memman.c:100-200 (through CHKERR,
 *                                               src/include/common.h:484-494)
 *    Num_dot_dprimme / Num_gemv_dprimme / Num_larnv_dprimme
 *                                               src/linalg/blaslapack.c:923,700-760,938-977
 *    ortho_single_iteration_dprimme             src/eigs/ortho.c:826-934
 *
 * All of them work on HOST arrays here (the driver allocates evecs with malloc and checks the
 * returned solution on the host, tests/COMMON/ioandtest.c:86-150); nothing in this file is on the
 * solver's hot path.  `pb_ref_context` mirrors `primme_context` (src/include/common.h:610-641,
 * built without PRIMME_PROFILE) field by field: it is passed BY VALUE across this boundary
 * (tests/test_abi.py checks size and offsets against the reference header).
 */
#include "pb_host.h"
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct pb_ref_context;
typedef struct pb_ref_alloc {
   void *p;
   int (*free_fn)(void *, struct pb_ref_context);
   struct pb_ref_alloc *prev;
} pb_ref_alloc;

typedef struct pb_ref_frame {
   pb_ref_alloc *prev_alloc;
   int keep_frame;
   struct pb_ref_frame *prev;
} pb_ref_frame;

typedef struct pb_ref_context {
   primme_params *primme;
   void *primme_svds;
   int printLevel;
   FILE *outputFile;
   int (*report)(const char *fun, double time, struct pb_ref_context ctx);
   pb_ref_frame *mm;
   int numProcs;
   int procID;
   void *mpicomm;
   int (*bcast)(void *buffer, primme_op_datatype type, int count, struct pb_ref_context ctx);
   int (*globalSum)(void *buffer, primme_op_datatype type, int count, struct pb_ref_context ctx);
   void *queue;
} pb_ref_context;

/* sizes for tests/abi_probe (compared with sizeof(primme_context) of the reference header) */
int primme_b200_ref_context_size(void) { return (int)sizeof(pb_ref_context); }
int primme_b200_ref_context_offset(int field) {
   switch (field) {
   case 0: return (int)offsetof(pb_ref_context, primme);
   case 1: return (int)offsetof(pb_ref_context, printLevel);
   case 2: return (int)offsetof(pb_ref_context, report);
   case 3: return (int)offsetof(pb_ref_context, mm);
   case 4: return (int)offsetof(pb_ref_context, numProcs);
   case 5: return (int)offsetof(pb_ref_context, bcast);
   case 6: return (int)offsetof(pb_ref_context, globalSum);
   case 7: return (int)offsetof(pb_ref_context, queue);
   default: return -1;
   }
}

static int ref_report(const char *fun, double time, pb_ref_context ctx) {
   (void)time;
   if (ctx.outputFile && fun) fprintf(ctx.outputFile, "%s\n", fun);
   return 0;
}

static int ref_global_sum(void *buffer, primme_op_datatype type, int count, pb_ref_context ctx) {
   primme_params *primme = ctx.primme;
   if (!primme || primme->numProcs <= 1 || count <= 0) return 0;
   if (!primme->globalSumReal || (type != primme_op_double && type != primme_op_default))
      return PRIMME_FUNCTION_UNAVAILABLE;
   int ierr = 0;
   primme->globalSumReal(buffer, buffer, &count, primme, &ierr);
   return ierr ? PRIMME_USER_FAILURE : 0;
}

static int ref_bcast(void *buffer, primme_op_datatype type, int count, pb_ref_context ctx) {
   primme_params *primme = ctx.primme;
   if (!primme || primme->numProcs <= 1 || count <= 0) return 0;
   if (!primme->broadcastReal || (type != primme_op_double && type != primme_op_default))
      return PRIMME_FUNCTION_UNAVAILABLE;
   int ierr = 0;
   primme->broadcastReal(buffer, &count, primme, &ierr);
   return ierr ? PRIMME_USER_FAILURE : 0;
}

/* every error macro of the reference assumes one frame below it (auxiliary_eigs.c:140-144) */
pb_ref_context primme_get_context(primme_params *primme) {
   pb_ref_context ctx;
   memset(&ctx, 0, sizeof(ctx));
   if (primme) {
      ctx.primme = primme;
      ctx.printLevel = primme->printLevel;
      ctx.outputFile = primme->outputFile;
      ctx.numProcs = primme->numProcs;
      ctx.procID = primme->procID;
      ctx.mpicomm = primme->commInfo;
      ctx.globalSum = ref_global_sum;
      ctx.bcast = ref_bcast;
      ctx.queue = primme->queue;
      ctx.report = ref_report;
   }
   pb_ref_frame *f = (pb_ref_frame *)calloc(1, sizeof(*f));
   if (f) f->prev = NULL;
   ctx.mm = f;
   return ctx;
}

/* releases what was registered in the top frame (nothing registers allocations on this side of
 * the boundary; the loop is there for callers that do) */
static void release_frame_allocations(pb_ref_frame *f, pb_ref_context ctx) {
   pb_ref_alloc *a = f->prev_alloc;
   while (a) {
      pb_ref_alloc *prev = a->prev;
      if (a->free_fn && a->p != (void *)f) a->free_fn(a->p, ctx);
      free(a);
      a = prev;
   }
   f->prev_alloc = NULL;
}

int Mem_pop_frame(pb_ref_context *ctx) {
   if (!ctx || !ctx->mm) return 0;
   pb_ref_frame *f = ctx->mm;
   if (f->keep_frame && f->prev) {
      /* hand the registrations over to the enclosing frame (memman.c:119-133) */
      pb_ref_alloc *a = f->prev_alloc;
      if (a) {
         while (a->prev) a = a->prev;
         a->prev = f->prev->prev_alloc;
         f->prev->prev_alloc = f->prev_alloc;
         f->prev_alloc = NULL;
      }
   } else {
      release_frame_allocations(f, *ctx);
   }
   ctx->mm = f->prev;
   return 0;
}

int Mem_pop_clean_frame(pb_ref_context ctx) {
   if (ctx.mm) release_frame_allocations(ctx.mm, ctx);
   return 0;
}

int Mem_debug_frame(const char *where, pb_ref_context ctx) {
   (void)where, (void)ctx;
   return 0;
}

void primme_free_context(pb_ref_context ctx) {
   if (!ctx.mm) return;
   release_frame_allocations(ctx.mm, ctx);
   free(ctx.mm);
}

/* ---- dense helpers on host arrays (blaslapack.c:923, :700-760, :938-977) ---- */
extern double ddot_(int *, const double *, int *, const double *, int *);
extern void dgemv_(const char *, int *, int *, double *, const double *, int *, const double *,
      int *, double *, double *, int *);
extern void dlarnv_(int *, int *, int *, double *);

double Num_dot_dprimme(PRIMME_INT n, double *x, PRIMME_INT incx, double *y, PRIMME_INT incy,
      pb_ref_context ctx) {
   (void)ctx;
   double s = 0.0;
   /* BLAS takes 32-bit sizes: long vectors go piecewise */
   while (n > 0) {
      int ln = n > 0x7ffffff0LL ? 0x7ffffff0 : (int)n, ix = (int)incx, iy = (int)incy;
      s += ddot_(&ln, x, &ix, y, &iy);
      x += (size_t)ln * incx, y += (size_t)ln * incy, n -= ln;
   }
   return s;
}

int Num_gemv_dprimme(const char *transa, PRIMME_INT m, PRIMME_INT n, double alpha, double *a,
      PRIMME_INT lda, double *x, PRIMME_INT incx, double beta, double *y, PRIMME_INT incy,
      pb_ref_context ctx) {
   (void)ctx;
   const int tr = (*transa == 'n' || *transa == 'N') ? 0 : 1;
   const PRIMME_INT leny = tr ? n : m;
   if (leny <= 0) return 0;
   if ((tr ? m : n) <= 0) {
      /* empty product: y <- beta*y (blaslapack.c:712-722) */
      for (PRIMME_INT i = 0; i < leny; i++) y[i * incy] = beta == 0.0 ? 0.0 : beta * y[i * incy];
      return 0;
   }
   if (m > 0x7fffffffLL || n > 0x7fffffffLL || lda > 0x7fffffffLL) return PRIMME_FUNCTION_UNAVAILABLE;
   int lm = (int)m, ln = (int)n, llda = (int)lda, ix = (int)incx, iy = (int)incy;
   dgemv_(transa, &lm, &ln, &alpha, a, &llda, x, &ix, &beta, y, &iy);
   return 0;
}

int Num_larnv_dprimme(int idist, PRIMME_INT *iseed, PRIMME_INT length, double *x, pb_ref_context ctx) {
   (void)ctx;
   int seed[4];
   for (int i = 0; i < 4; i++) seed[i] = (int)iseed[i];
   while (length > 0) {
      int chunk = length > 0x7ffffff0LL ? 0x7ffffff0 : (int)length;
      dlarnv_(&idist, seed, &chunk, x);
      x += chunk, length -= chunk;
   }
   for (int i = 0; i < 4; i++) iseed[i] = seed[i];
   return 0;
}

/* One projection X <- X - BQ (Q' X) and the norms of the result, on host arrays
 * (ortho.c:826-934; the coefficients are NOT solved against QtBQ before the update there
 * either: z is computed and dropped, :884-905). */
int ortho_single_iteration_dprimme(double *Q, int nQ, PRIMME_INT ldQ, double *BQ, PRIMME_INT ldBQ,
      double *QtBQ, int ldQtBQ, double *X, int *inX, int nX, PRIMME_INT ldX, double *norms,
      pb_ref_context ctx) {
   (void)QtBQ, (void)ldQtBQ;
   primme_params *primme = ctx.primme;
   if (!primme) return PRIMME_UNEXPECTED_FAILURE;
   const PRIMME_INT n = primme->nLocal;
   if (nX <= 0) return 0;
   double *y = (double *)calloc((size_t)(nQ > 0 ? nQ : 1) * nX, sizeof(double));
   if (!y) return PRIMME_MALLOC_FAILURE;
   for (int j = 0; j < nX; j++) {
      const double *x = X + (size_t)ldX * (inX ? inX[j] : j);
      for (int i = 0; i < nQ; i++) {
         const double *q = Q + (size_t)ldQ * i;
         double s = 0.0;
         for (PRIMME_INT r = 0; r < n; r++) s += q[r] * x[r];
         y[i + (size_t)nQ * j] = s;
      }
   }
   primme->stats.numOrthoInnerProds += (double)nQ * nX;
   int rc = ctx.globalSum ? ctx.globalSum(y, primme_op_double, nQ * nX, ctx) : 0;
   if (!rc) {
      for (int j = 0; j < nX; j++) {
         double *x = X + (size_t)ldX * (inX ? inX[j] : j);
         for (int i = 0; i < nQ; i++) {
            const double *bq = (BQ ? BQ : Q) + (size_t)(BQ ? ldBQ : ldQ) * i;
            const double c = y[i + (size_t)nQ * j];
            for (PRIMME_INT r = 0; r < n; r++) x[r] -= bq[r] * c;
         }
         if (norms) {
            double s = 0.0;
            for (PRIMME_INT r = 0; r < n; r++) s += x[r] * x[r];
            norms[j] = s;
         }
      }
      if (norms) {
         rc = ctx.globalSum ? ctx.globalSum(norms, primme_op_double, nX, ctx) : 0;
         for (int j = 0; j < nX; j++) norms[j] = sqrt(norms[j]);
         primme->stats.numOrthoInnerProds += nX;
      }
   }
   free(y);
   return rc;
}

/* ---- complex double flavours (tests/driver.c built with -DUSE_DOUBLECOMPLEX) ---- */
#include <complex.h>
typedef double _Complex pb_z;
extern void zgemv_(const char *, int *, int *, pb_z *, const pb_z *, int *, const pb_z *, int *, pb_z *, pb_z *, int *);

pb_z Num_dot_zprimme(PRIMME_INT n, pb_z *x, PRIMME_INT incx, pb_z *y, PRIMME_INT incy, pb_ref_context ctx) {
   (void)ctx;
   pb_z s = 0.0; /* the reference's explicit zdotc (blaslapack.c:899-913) */
   for (PRIMME_INT i = 0; i < n; i++) s += conj(x[i * incx]) * y[i * incy];
   return s;
}

int Num_gemv_zprimme(const char *transa, PRIMME_INT m, PRIMME_INT n, pb_z alpha, pb_z *a, PRIMME_INT lda, pb_z *x,
      PRIMME_INT incx, pb_z beta, pb_z *y, PRIMME_INT incy, pb_ref_context ctx) {
   (void)ctx;
   const int tr = (*transa == 'n' || *transa == 'N') ? 0 : 1;
   const PRIMME_INT leny = tr ? n : m;
   if (leny <= 0) return 0;
   if ((tr ? m : n) <= 0) {
      for (PRIMME_INT i = 0; i < leny; i++) y[i * incy] = beta == 0.0 ? 0.0 : beta * y[i * incy];
      return 0;
   }
   if (m > 0x7fffffffLL || n > 0x7fffffffLL || lda > 0x7fffffffLL) return PRIMME_FUNCTION_UNAVAILABLE;
   int lm = (int)m, ln = (int)n, llda = (int)lda, ix = (int)incx, iy = (int)incy;
   zgemv_(transa, &lm, &ln, &alpha, a, &llda, x, &ix, &beta, y, &iy);
   return 0;
}

int Num_larnv_zprimme(int idist, PRIMME_INT *iseed, PRIMME_INT length, pb_z *x, pb_ref_context ctx) {
   /* the real generator on twice the length (blaslapack.c:938-949) */
   return Num_larnv_dprimme(idist, iseed, 2 * length, (double *)x, ctx);
}

int ortho_single_iteration_zprimme(pb_z *Q, int nQ, PRIMME_INT ldQ, pb_z *BQ, PRIMME_INT ldBQ, pb_z *QtBQ, int ldQtBQ,
      pb_z *X, int *inX, int nX, PRIMME_INT ldX, double *norms, pb_ref_context ctx) {
   (void)QtBQ, (void)ldQtBQ;
   primme_params *primme = ctx.primme;
   if (!primme) return PRIMME_UNEXPECTED_FAILURE;
   const PRIMME_INT n = primme->nLocal;
   if (nX <= 0) return 0;
   pb_z *y = (pb_z *)calloc((size_t)(nQ > 0 ? nQ : 1) * nX, sizeof(pb_z));
   if (!y) return PRIMME_MALLOC_FAILURE;
   for (int j = 0; j < nX; j++) {
      const pb_z *x = X + (size_t)ldX * (inX ? inX[j] : j);
      for (int i = 0; i < nQ; i++) {
         const pb_z *q = Q + (size_t)ldQ * i;
         pb_z s = 0.0;
         for (PRIMME_INT r = 0; r < n; r++) s += conj(q[r]) * x[r];
         y[i + (size_t)nQ * j] = s;
      }
   }
   primme->stats.numOrthoInnerProds += (double)nQ * nX;
   int rc = ctx.globalSum ? ctx.globalSum(y, primme_op_double, 2 * nQ * nX, ctx) : 0;
   if (!rc) {
      for (int j = 0; j < nX; j++) {
         pb_z *x = X + (size_t)ldX * (inX ? inX[j] : j);
         for (int i = 0; i < nQ; i++) {
            const pb_z *bq = (BQ ? BQ : Q) + (size_t)(BQ ? ldBQ : ldQ) * i;
            const pb_z c = y[i + (size_t)nQ * j];
            for (PRIMME_INT r = 0; r < n; r++) x[r] -= bq[r] * c;
         }
         if (norms) {
            double s = 0.0;
            for (PRIMME_INT r = 0; r < n; r++) s += creal(conj(x[r]) * x[r]);
            norms[j] = s;
         }
      }
      if (norms) {
         rc = ctx.globalSum ? ctx.globalSum(norms, primme_op_double, nX, ctx) : 0;
         for (int j = 0; j < nX; j++) norms[j] = sqrt(norms[j]);
         primme->stats.numOrthoInnerProds += nX;
      }
   }
   free(y);
   return rc;
}

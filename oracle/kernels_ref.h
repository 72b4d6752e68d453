/* oracle/kernels_ref.h -- TEST INFRASTRUCTURE: structures shared by the CPU restatement of the kernels
 * (kernels_ref.c: fp64, kernels_ref_z.c: complex fp64). */
#ifndef KERNELS_REF_H
#define KERNELS_REF_H
#include "../include/primme_b200.h"

struct pb200_ctx {
   int64_t launches;
   int nranks, rank;
   void *ws_ptr[4];
   size_t ws_bytes[4];
};

struct pb200_csr {
   int64_t nrows, ncols, nnz;
   int64_t *rowptr; /* 0-based */
   int32_t *colind; /* 0-based */
   double *vals;    /* nnz (or 2*nnz if complex) */
   int is_complex;
   struct pb200_csr *T;
};

#endif

// dist.cu -- row-sharded operator for one-process-per-GPU runs.
//
// PRIMME's SPMD model (reference include/primme_eigs.h:187-198, examples/ex_eigs_mpi.c:106-112):
// rank r owns a contiguous block of rows of A and of every multivector.  The only data-path
// collective of the Davidson loop besides the small panel reductions is the SpMV halo: every
// rank needs the entries of the block x that its off-diagonal columns reference.  v1 gathers
// the whole block with one grouped NCCL broadcast per (column, rank) over NVLink (all-gather
// with unequal counts) into a persistent device buffer and runs the local block-CSR SpMM with
// global column indices on it.
#include "pb200_internal.cuh"
#include "../../include/primme.h"
#include <stdlib.h>
#include <string.h>

struct pb200_dist_csr {
   pb200_csr *A;       // nLocal rows, global column indices
   int nranks;
   int64_t nglobal;
   int64_t *counts, *displs;  // host
   double *Xg;         // nglobal x bcap gathered block
   int bcap;
};

extern "C" int pb200_dist_csr_create(pb200_ctx *ctx, pb200_csr *A_local, const int64_t *counts_host,
      int nranks, pb200_dist_csr **out) {
   if (nranks != ctx->nranks) return PB200_ERR_ARG;
   pb200_dist_csr *D = (pb200_dist_csr *)calloc(1, sizeof(*D));
   if (!D) return PB200_ERR_ALLOC;
   D->A = A_local, D->nranks = nranks;
   D->counts = (int64_t *)malloc(sizeof(int64_t) * nranks);
   D->displs = (int64_t *)malloc(sizeof(int64_t) * nranks);
   int64_t off = 0;
   for (int r = 0; r < nranks; r++) D->counts[r] = counts_host[r], D->displs[r] = off, off += counts_host[r];
   D->nglobal = off;
   D->bcap = 8;
   PB_CUDA(cudaMalloc((void **)&D->Xg, sizeof(double) * (size_t)(off > 0 ? off : 1) * D->bcap));
   *out = D;
   return 0;
}

extern "C" int pb200_dist_csr_destroy(pb200_ctx *ctx, pb200_dist_csr *D) {
   if (!D) return 0;
   if (ctx) cudaStreamSynchronize(ctx->stream);
   cudaFree(D->Xg);
   free(D->counts), free(D->displs), free(D);
   return 0;
}

// Y(local rows, 0:b) = A_local * allgather(X)
extern "C" int pb200_ddist_spmm(pb200_ctx *ctx, pb200_dist_csr *D, const double *X, int64_t ldx,
      double *Y, int64_t ldy, int ncols) {
   for (int c0 = 0; c0 < ncols; c0 += D->bcap) {
      int b = ncols - c0 < D->bcap ? ncols - c0 : D->bcap;
      const double *Xc = X + (size_t)c0 * ldx;
      if (ctx->nranks > 1) {
         PB_CHK(pb_nccl_allgatherv_cols(ctx, Xc, ldx, D->Xg, D->nglobal, D->counts, D->displs, b));
         PB_CHK(pb200_dspmm(ctx, D->A, D->Xg, D->nglobal, Y + (size_t)c0 * ldy, ldy, b));
      } else {
         PB_CHK(pb200_dspmm(ctx, D->A, Xc, ldx, Y + (size_t)c0 * ldy, ldy, b));
      }
   }
   return 0;
}

// primme.matrix = pb200_dist_csr*; primme.matrixMatvec = primme_b200_dist_csr_matvec
extern "C" void primme_b200_dist_csr_matvec(void *x, int64_t *ldx, void *y, int64_t *ldy,
      int *blockSize, struct primme_params *primme, int *ierr) {
   pb200_ctx *ctx = primme_b200_solver_ctx(primme);
   pb200_dist_csr *D = (pb200_dist_csr *)primme->matrix;
   if (!ctx || !D) {
      *ierr = -1;
      return;
   }
   *ierr = pb200_ddist_spmm(ctx, D, (const double *)x, *ldx, (double *)y, *ldy, *blockSize);
}

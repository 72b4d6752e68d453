/* primme_b200.h -- C-ABI of the sm_100a kernel layer under the Davidson inner loop.
 *
 * This is the drop-in boundary below the host solver: plain pointers and sizes, no C++/torch
 * types.  It replaces, one level higher (fused L2 operations instead of per-BLAS calls), the 22
 * device functions of the reference's cuBLAS back end (reference src/linalg/cublas_wrapper.c:
 * 162-987) and the MAGMA twin (src/linalg/magma_wrapper.c:125-1041), plus the user-side
 * cusparseSpMM callback of the reference GPU example (examples/ex_eigs_dcublas.c:238-263).
 *
 * Conventions: column-major, leading dimensions in elements, "d" = fp64, "z" = complex fp64
 * (interleaved re,im).  Pointers named *_host are host memory (any, pinned not required);
 * every other data pointer is device memory on the context's GPU.  Small operands (coefficient
 * blocks, panels) are host-side exactly like the reference's HSCALAR operands.  Every function
 * returns 0 or a negative PRIMME-style error code and is synchronous with respect to its
 * *_host outputs (the panel is valid on return).
 *
 * The library has NO CPU fallback: without a CUDA device pb200_ctx_create fails with
 * PB200_ERR_NO_DEVICE and the solvers return PRIMME_FUNCTION_UNAVAILABLE.
 */
#ifndef PRIMME_B200_H
#define PRIMME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_ERR_NO_DEVICE (-44) /* == PRIMME_FUNCTION_UNAVAILABLE */
#define PB200_ERR_CUDA      (-1)
#define PB200_ERR_ALLOC     (-2)
#define PB200_ERR_ARG       (-5)

typedef struct pb200_ctx pb200_ctx; /* stream, scratch pool, pinned staging, optional NCCL comm */
typedef struct pb200_csr pb200_csr; /* device-resident CSR matrix + row-block schedule */

/* ------------------------------------------------------------------ context / memory ---- */
/* replaces: cublasCreate + per-call cudaMalloc/cudaFree (cublas_wrapper.c:187-232) */
int pb200_device_count(void);
int pb200_ctx_create(pb200_ctx **ctx, int device /* -1: current */);
int pb200_ctx_destroy(pb200_ctx *ctx);
/* persisting L2 window over the head of a basis array of `bytes` bytes that every sweep re-reads (0 bytes:
 * remove it); dprimme / zprimme set it for V during a solve when the basis exceeds the L2 */
int64_t pb200_ctx_l2_persist(pb200_ctx *ctx, const void *ptr, size_t bytes);
int pb200_ctx_sync(pb200_ctx *ctx);
/* marks the start of a solve on a long-lived context (resets the alternating row-sweep direction so
 * that repeated solves are bitwise reproducible); dprimme / zprimme call it themselves */
int pb200_ctx_begin_solve(pb200_ctx *ctx);
void *pb200_ctx_stream(pb200_ctx *ctx); /* cudaStream_t */
/* kernel-launch counter (bench.py reports it as gpu_launches) */
int64_t pb200_ctx_launches(pb200_ctx *ctx);
/* Per-kernel-kind CUDA-event timing on the context's stream (kind: 0 spmm, 1 ortho sweep,
 * 2 vwxr, 3 utilities, 4 panel reduction): number of launches, total milliseconds and total
 * algorithmic bytes since profiling was switched on.  Used by bench.py for the roofline. */
int pb200_ctx_set_profiling(pb200_ctx *ctx, int on);
int pb200_ctx_get_profile(pb200_ctx *ctx, int kind, int64_t *count, double *ms, double *bytes);
/* Attach an NCCL communicator (ncclComm_t as void*): every *_host panel is then all-reduced
 * (sum) across ranks on the device before it is copied back; nranks==1 detaches. */
int pb200_ctx_set_comm(pb200_ctx *ctx, void *nccl_comm, int nranks, int rank);
int pb200_ctx_nranks(pb200_ctx *ctx);
/* host-buffer allreduce/bcast over the attached communicator (globalSumReal equivalents) */
int pb200_allreduce_host(pb200_ctx *ctx, double *buf_host, int count);
int pb200_bcast_host(pb200_ctx *ctx, double *buf_host, int count, int root);

int pb200_malloc(pb200_ctx *ctx, size_t bytes, void **dptr);
/* Workspace owned by the context (slot 0..3), grown on demand and kept until pb200_ctx_destroy:
 * the solver's basis arrays when the caller attached a long-lived context. */
int pb200_ctx_workspace(pb200_ctx *ctx, int slot, size_t bytes, void **dptr);
int pb200_free(pb200_ctx *ctx, void *dptr);
int pb200_memset0(pb200_ctx *ctx, void *dptr, size_t bytes);
/* 2-D copies, element size es bytes; replaces Num_set_matrix/get_matrix/copy_matrix
 * (cublas_wrapper.c:335,370,739) */
int pb200_copy_h2d(pb200_ctx *ctx, const void *src_host, int64_t lds, void *dst, int64_t ldd,
      int64_t rows, int cols, int es);
int pb200_copy_d2h(pb200_ctx *ctx, const void *src, int64_t lds, void *dst_host, int64_t ldd,
      int64_t rows, int cols, int es);
int pb200_copy_d2d(pb200_ctx *ctx, const void *src, int64_t lds, void *dst, int64_t ldd,
      int64_t rows, int cols, int es);
/* replaces Num_check_pointer (cublas_wrapper.c:162): 1 device, 0 host, <0 error */
int pb200_is_device_pointer(const void *p);

/* ------------------------------------------------------------------ K1: block-CSR SpMM -- */
/* Y(:,0:ncols) = A * X(:,0:ncols); replaces the user-side cusparseSpMM of
 * examples/ex_eigs_dcublas.c:238-263 and tests/COMMON/mat.c:68 (CSRMatrixMatvec).
 * index_base 0 or 1 (tests/COMMON/csr.c uses 1).  is_complex selects z values. */
int pb200_csr_create(pb200_ctx *ctx, int64_t nrows, int64_t ncols, int64_t nnz,
      const int64_t *rowptr_host, const int32_t *colind_host, const void *vals_host,
      int index_base, int is_complex, pb200_csr **A);
/* same with the device arrays taken from the context's matrix pool (no cudaMalloc / cudaFree per
 * matrix): for callers that upload a matrix per solve; valid until the next pooled create on ctx */
int pb200_csr_create_pooled(pb200_ctx *ctx, int64_t nrows, int64_t ncols, int64_t nnz,
      const int64_t *rowptr_host, const int32_t *colind_host, const void *vals_host,
      int index_base, int is_complex, pb200_csr **A);
int pb200_csr_destroy(pb200_ctx *ctx, pb200_csr *A);
int64_t pb200_csr_nnz(const pb200_csr *A);
/* gather layout the product of width ncols runs with on this matrix, decided by timing at the first
 * product: 0 not decided yet, 1 column-major gathers, 2 row-major gather copy, 3 windowed right-hand
 * sides staged in shared memory (matrices with column locality; see csrc/spmm.cu) */
int pb200_csr_layout(const pb200_csr *A, int ncols);
int pb200_dspmm(pb200_ctx *ctx, const pb200_csr *A, const double *X, int64_t ldx, double *Y,
      int64_t ldy, int ncols);
/* complex twin (zprimme): A created with is_complex = 1; X, Y interleaved (re,im), column-major,
 * leading dimensions in complex elements */
int pb200_zspmm(pb200_ctx *ctx, const pb200_csr *A, const void *X, int64_t ldx, void *Y,
      int64_t ldy, int ncols);
/* Y = A^T X (used by the SVD normal-equations operator, src/svds/primme_svds_c.c:1337-1351);
 * requires the transposed copy built by pb200_csr_build_transpose. */
int pb200_csr_build_transpose(pb200_ctx *ctx, pb200_csr *A);
int pb200_dspmm_t(pb200_ctx *ctx, const pb200_csr *A, const double *X, int64_t ldx, double *Y,
      int64_t ldy, int ncols);

/* ---- row-sharded operator (one process per GPU) --------------------------------
 * A_local holds this rank's rows with GLOBAL column indices; counts_host[r] = rows of rank r.
 * pb200_dist_csr_create (collective over the context's communicator) finds the distinct remote
 * columns each rank references, exchanges the lists, REMAPS A_local's column indices in place to
 * [own rows | compacted halo] (A_local then belongs to the operator: do not use it with pb200_dspmm)
 * and maps every rank's gather buffer into its peers (CUDA IPC over NVLink).  Per block:
 * Y_local = A_local * [X_local; halo]: one kernel packs the local block and pushes the rows each peer
 * needs into the peer's buffer, the SpMM kernel waits for the peers' flags -- no NCCL call on the data
 * path (compacted ncclSend/ncclRecv only if peer mapping is unavailable).  Replaces the MPI matvec of
 * reference examples/ex_eigs_mpi.c:106-112,209-218. */
typedef struct pb200_dist_csr pb200_dist_csr;
int pb200_comm_unique_id(void *id128);
int pb200_ctx_comm_init(pb200_ctx *ctx, int nranks, int rank, const void *id128);
int pb200_ctx_comm_free(pb200_ctx *ctx);
/* Peer-memory panel exchange (one node, one process per GPU, after pb200_ctx_comm_init): every
 * panel kernel then all-reduces its own panel over NVLink peer memory instead of a separate NCCL
 * call (see PbFin in csrc/pb200_internal.cuh).  export: allocates this rank's exchange buffer and
 * returns its cudaIpcMemHandle_t (64 bytes); the launcher all-gathers the handles; attach: opens the
 * peers' buffers.  attach is COLLECTIVE over the context's communicator: the ranks agree on the
 * outcome, so it returns 0 on every rank (exchange on) or 1 on every rank (some rank could not open a
 * peer buffer, or PB200_NO_PEER_EXCHANGE is set somewhere: all panels stay on NCCL).
 * Platform assumption of the exchange (csrc/pb200_internal.cuh, PbFin): a 16-byte aligned
 * st.volatile.v2.f64 to peer or mapped host memory is observed as one unit (true on x86-64 hosts
 * and NVLink/PCIe peers of this platform; the tag sits in the upper 8 bytes and the host reads it
 * with an acquire fence before the value). */
int pb200_ctx_peer_export(pb200_ctx *ctx, void *handle64);
int pb200_ctx_peer_attach(pb200_ctx *ctx, int nranks, int rank, const void *handles);
int pb200_ctx_peer_active(pb200_ctx *ctx);
int pb200_dist_csr_create(pb200_ctx *ctx, pb200_csr *A_local, const int64_t *counts_host,
      int nranks, pb200_dist_csr **D);
int pb200_dist_csr_destroy(pb200_ctx *ctx, pb200_dist_csr *D);
int pb200_ddist_spmm(pb200_ctx *ctx, pb200_dist_csr *D, const double *X, int64_t ldx, double *Y,
      int64_t ldy, int ncols);
int pb200_zdist_spmm(pb200_ctx *ctx, pb200_dist_csr *D, const void *X, int64_t ldx, void *Y,
      int64_t ldy, int ncols);
/* local rows, compacted halo rows received per block, rows pushed to peers per block, 1 if the halo
 * travels over peer memory (0: NCCL point-to-point) */
int pb200_dist_csr_info(const pb200_dist_csr *D, int64_t *nloc, int64_t *nhalo, int64_t *rows_sent_per_block,
      int *peer_halo);

/* ------------------------------------------------- K2/K3/K4: fused block-ortho row sweep -- */
/* One pass over the rows of [Q V X]:
 *    if C_host:  X <- (X - [Q V(:,0:mv)] * C) * Y      (Y = identity if Y_host == NULL)
 *    if P_host:  P <- [Q V(:,0:mv)  (X if xx)]^H * X    ((q+mv(+b)) x b, written to host)
 * X is n x b.  Covers Num_ortho_kernel (reference src/eigs/ortho.c:963-1072: update :1017-1038,
 * Gram :1043-1059), update_projection's V^H W panel (update_projection.c:99-102; call with
 * X = W block, xx = 0), the CGS gemv pair of Bortho_gen (ortho.c:237-291; b = 1) and
 * ortho_single_iteration (ortho.c:826-934).  Replaces Num_gemm_ddh/gemm_dhd/trsm_hd/
 * compute_gramm_ddh (cublas_wrapper.c:452-505,785,898-987).
 * C is (q+mv) x b with leading dimension ldc; Y is b x b (general matrix, applied from the
 * right).  The panel is reduced in a fixed order (bitwise reproducible run to run). */
int pb200_dortho_sweep(pb200_ctx *ctx, int64_t n, const double *Q, int q, int64_t ldq,
      const double *V, int mv, int64_t ldv, double *X, int b, int64_t ldx,
      const double *C_host, int ldc, const double *Y_host, int ldy, int xx, double *P_host,
      int ldp);

/* ------------------------------------------------------------------------- K5: VWXR ----- */
/* Column range [cb,ce) of the small matrix h, written to ptr (n x (ce-cb), leading dim ld). */
typedef struct pb200_cols {
   double *ptr;
   int64_t ld;
   int cb, ce;
} pb200_cols;

/* Outputs of one VWXR sweep; unused members have ptr == NULL / counts 0. */
typedef struct pb200_vwxr_out {
   pb200_cols X[3];     /* X_k = V*h(:,cb:ce)            (may alias V: restart V <- V*h) */
   pb200_cols Wo;       /* Wo  = W*h(:,cb:ce)            (may alias W) */
   pb200_cols R;        /* R   = W*h(:,cb:ce) - V*h(:,cb:ce)*diag(theta(cb:ce)) */
   double *Rnorms_host; /* ||R(:,j)||_2 (sqrt applied), length R.ce-R.cb, optional */
   int rb, re;          /* extra residual norms of columns [rb,re) without storing them */
   double *rnorms_host;
   int nG;              /* G = (V*h(:,0:nG))^H (V*h(:,0:nG)), full nG x nG on host */
   double *G_host;
   int ldG;
   int nH;              /* H = (V*h(:,0:nH))^H (W*h(:,0:nH)) */
   double *H_host;
   int ldH;
   double *P_host;      /* optional: P = [V R]^H R, (m + nR) x nR with nR = R.ce - R.cb: the first Gram panel
                           of the block orthogonalisation (ortho.c:1043-1059) when the new block is the
                           residual block itself; only when pb200_dvwxr_can_fuse_gram() says so */
   int ldP;
   double *R2;          /* optional, with P_host only: the residual columns are ALSO written here (n x nR,
                           leading dimension ldR2, 16-byte aligned) -- the next basis block when the
                           correction is the residual itself, so no copy kernel follows */
   int64_t ldR2;
} pb200_vwxr_out;

/* One sweep over the rows of V and W (both n x m, leading dimension ld):
 * reference Num_update_VWXR_Sprimme (src/eigs/auxiliary_eigs_normal.c:155-388) restricted to
 * B = I.  h_host is m x nh (ldh); theta_host has nh entries (indexed like h's columns). */
int pb200_dvwxr(pb200_ctx *ctx, int64_t n, const double *V, const double *W, int m,
      int64_t ld, const double *h_host, int ldh, int nh, const double *theta_host,
      const pb200_vwxr_out *out);
/* 1 when a sweep of this shape can also deliver out->P_host (nh <= 8, no G/H, aligned operands) */
int pb200_dvwxr_can_fuse_gram(pb200_ctx *ctx, int64_t n, const double *V, const double *W, int m,
      int64_t ld, int nh, const pb200_vwxr_out *out);

/* ------------------------------------------------------------ K6: multivector utilities -- */
/* X(:,i) <- X(:,perm[i]) in place (reference permute_vecs, src/linalg/auxiliary.c:716-793) */
int pb200_dpermute_columns(pb200_ctx *ctx, int64_t n, double *X, int64_t ldx,
      const int *perm_host, int ncols);
/* Y(:,yin[i]) <- X(:,xin[i]) (NULL index list = identity; auxiliary.c:649-662) */
int pb200_dcopy_columns(pb200_ctx *ctx, int64_t n, const double *X, int64_t ldx,
      const int *xin_host, double *Y, int64_t ldy, const int *yin_host, int ncols);
/* Y(:,j) += alpha[j] * X(:,j)   (Num_axpy per column: correction.c:368, main_iter.c:1873) */
int pb200_daxpy_columns(pb200_ctx *ctx, int64_t n, const double *alpha_host, const double *X,
      int64_t ldx, double *Y, int64_t ldy, int ncols);
/* X(:,j) *= alpha[j] */
int pb200_dscale_columns(pb200_ctx *ctx, int64_t n, const double *alpha_host, double *X,
      int64_t ldx, int ncols);
/* out[j] = X(:,j)^H Y(:,j)  (Num_dist_dots, auxiliary_eigs.c:662) */
int pb200_dcolumn_dots(pb200_ctx *ctx, int64_t n, const double *X, int64_t ldx,
      const double *Y, int64_t ldy, int ncols, double *out_host);
/* W(:,j) -= theta[j] V(:,j); out[j] = ||W(:,j)||^2   (verify_norms, main_iter.c:1864-1881) */
int pb200_dresidual_inplace(pb200_ctx *ctx, int64_t n, const double *theta_host,
      const double *V, int64_t ldv, double *W, int64_t ldw, int ncols, double *out_host);
/* solution update of one QMR step in one pass (reference src/eigs/inner_solve.c:384-413):
 * Delta(:,j) = gamma[j] Delta(:,j) + eta[j] D(:,j);  Sol(:,j) += Delta(:,j);  dots_host[j] = |Sol(:,j)|^2
 * when dots_host != NULL.  Same roundings as scale + axpy + axpy + dot. */
int pb200_dqmr_update(pb200_ctx *ctx, int64_t n, const double *gamma_host, const double *eta_host,
      const double *D, int64_t ldd, double *Delta, int64_t ldl, double *Sol, int64_t lds, int ncols,
      double *dots_host);
/* y = x ./ (d - shift_j) style Jacobi preconditioner on a block (tests/COMMON/mat.c:137-165):
 * Y(:,j) = X(:,j) ./ safeguard(diag - shifts[j]) */
int pb200_djacobi(pb200_ctx *ctx, int64_t n, const double *diag, const double *shifts_host,
      double minabs, const double *X, int64_t ldx, double *Y, int64_t ldy, int ncols);

/* X(0:col_len, 0:ncols) = the next col_len * ncols values of LAPACK's dlarnv(idist = 2, iseed), bit for bit,
 * generated on the device (reference Num_larnv, src/linalg/blaslapack.c:953-977, fills its random vectors on the
 * host); iseed (entries 0..4095) is advanced as dlarnv would.  Complex blocks: col_len = 2 n, ld = 2 ldx. */
int pb200_dlarnv(pb200_ctx *ctx, long long iseed[4], int64_t col_len, int ncols, double *X, int64_t ld);

/* ------------------------------------------------ complex twins of K2-K6 (zprimme / C3) ----
 * Same operations and argument lists as the d functions above on interleaved (re,im) fp64 data:
 * device arrays column-major with leading dimensions in COMPLEX elements; host coefficient blocks
 * and panels complex (C, Y, P, h, G, H, alpha, dots); Ritz values, norms, shifts and the Jacobi
 * diagonal real.  "^H" is the conjugate transpose: P = [Q V X]^H X, G = X^H X, H = X^H Y,
 * dots[j] = X(:,j)^H Y(:,j).  pb200_vwxr_out is shared: its column pointers and G_host / H_host /
 * P_host point to complex data.  They are what zprimme / cublas_zprimme (reference
 * include/primme_eigs.h:392,416; instantiation src/include/template_types.h:51-204) run on. */
int pb200_zortho_sweep(pb200_ctx *ctx, int64_t n, const void *Q, int q, int64_t ldq, const void *V, int mv,
      int64_t ldv, void *X, int b, int64_t ldx, const void *C_host, int ldc, const void *Y_host, int ldy, int xx,
      void *P_host, int ldp);
int pb200_zvwxr(pb200_ctx *ctx, int64_t n, const void *V, const void *W, int m, int64_t ld, const void *h_host,
      int ldh, int nh, const double *theta_host, const pb200_vwxr_out *out);
int pb200_zvwxr_can_fuse_gram(pb200_ctx *ctx, int64_t n, const void *V, const void *W, int m, int64_t ld, int nh,
      const pb200_vwxr_out *out);
int pb200_zpermute_columns(pb200_ctx *ctx, int64_t n, void *X, int64_t ldx, const int *perm_host, int ncols);
int pb200_zcopy_columns(pb200_ctx *ctx, int64_t n, const void *X, int64_t ldx, const int *xin_host, void *Y,
      int64_t ldy, const int *yin_host, int ncols);
int pb200_zaxpy_columns(pb200_ctx *ctx, int64_t n, const void *alpha_host, const void *X, int64_t ldx, void *Y,
      int64_t ldy, int ncols);
int pb200_zscale_columns(pb200_ctx *ctx, int64_t n, const void *alpha_host, void *X, int64_t ldx, int ncols);
int pb200_zcolumn_dots(pb200_ctx *ctx, int64_t n, const void *X, int64_t ldx, const void *Y, int64_t ldy,
      int ncols, void *out_host);
int pb200_zresidual_inplace(pb200_ctx *ctx, int64_t n, const double *theta_host, const void *V, int64_t ldv,
      void *W, int64_t ldw, int ncols, double *out_host);
int pb200_zjacobi(pb200_ctx *ctx, int64_t n, const double *diag, const double *shifts_host, double minabs,
      const void *X, int64_t ldx, void *Y, int64_t ldy, int ncols);
/* gamma, eta real (the recurrences of the inner solver are real in every precision) */
int pb200_zqmr_update(pb200_ctx *ctx, int64_t n, const double *gamma_host, const double *eta_host, const void *D,
      int64_t ldd, void *Delta, int64_t ldl, void *Sol, int64_t lds, int ncols, double *dots_host);
/* 1 when the matrix holds complex values */
int pb200_csr_is_complex(const pb200_csr *A);

/* ------------------------------------------------- ready-made PRIMME callbacks (operators.c) --
 * Same signature as primme_params.matrixMatvec / applyPreconditioner (reference
 * include/primme_eigs.h:170-180); x and y are DEVICE pointers (cublas_dprimme contract). */
struct primme_params;
/* primme.matrix = pb200_csr*;  primme.matrixMatvec = primme_b200_csr_matvec */
void primme_b200_csr_matvec(void *x, int64_t *ldx, void *y, int64_t *ldy, int *blockSize,
      struct primme_params *primme, int *ierr);
/* primme.matrix = pb200_dist_csr*; primme.matrixMatvec = primme_b200_dist_csr_matvec */
void primme_b200_dist_csr_matvec(void *x, int64_t *ldx, void *y, int64_t *ldy, int *blockSize,
      struct primme_params *primme, int *ierr);
/* primme.preconditioner = primme_b200_jacobi*;  primme.applyPreconditioner = ..._jacobi_apply
 * (Davidson diagonal preconditioner of the reference test driver, tests/COMMON/mat.c:137-165) */
typedef struct primme_b200_jacobi {
   const double *diag_dev;
   double minabs;
   int use_shifts;
} primme_b200_jacobi;
void primme_b200_jacobi_apply(void *x, int64_t *ldx, void *y, int64_t *ldy, int *blockSize,
      struct primme_params *primme, int *ierr);
/* Make dprimme/cublas_dprimme run on this context (stream, NCCL communicator) instead of a
 * private one; pass NULL to detach. */
int primme_b200_attach_ctx(struct primme_params *primme, pb200_ctx *ctx);
pb200_ctx *primme_b200_attached_ctx(const struct primme_params *primme);
/* The context of the solve currently running on `primme` (for user callbacks). */
pb200_ctx *primme_b200_solver_ctx(const struct primme_params *primme);
/* Host CSR in, host eigenpairs out; matrix upload, device solve with the built-in SpMM and the
 * eigenvector download happen inside. */
int primme_b200_dprimme_csr(double *evals, double *evecs_host, double *resNorms,
      struct primme_params *primme, const int64_t *rowptr_host, const int32_t *colind_host,
      const double *vals_host, int index_base);
int primme_b200_zprimme_csr(double *evals, void *evecs_host, double *resNorms, struct primme_params *primme,
      const int64_t *rowptr_host, const int32_t *colind_host, const void *vals_host, int index_base);
void primme_b200_zjacobi_apply(void *x, int64_t *ldx, void *y, int64_t *ldy, int *blockSize,
      struct primme_params *primme, int *ierr);
/* Built-in operator of the SVD front end (cublas_dprimme_svds): primme_svds.matrix = pb200_csr*
 * of the m x n matrix with its transposed copy (pb200_csr_build_transpose), primme_svds.matrixMatvec
 * = this function; y = A x or A' x on device blocks (reference user callback contract
 * include/primme_svds.h:113-116, examples/ex_svds_dseq.c:188-230). */
struct primme_svds_params;
void primme_b200_svds_csr_matvec(void *x, int64_t *ldx, void *y, int64_t *ldy, int *blockSize,
      int *transpose, struct primme_svds_params *primme_svds, int *ierr);

/* Row-partitioned built-in operator of the SVD front end (config C4; primme_svds.numProcs > 1, mLocal /
 * nLocal set): primme_svds.matrix = primme_b200_svds_dist*, primme_svds.matrixMatvec = this function.
 *   A  : pb200_dist_csr over this rank's rows of A   (mLocal x n, global column ids; counts = the nLocal's)
 *   At : pb200_dist_csr over this rank's rows of A^T (nLocal x m, global column ids; counts = the mLocal's)
 * y = A x gathers the n-side block, y = A^T x gathers the m-side block, both over the compacted
 * peer-memory halo of the row-sharded operator (no reduce-scatter).  Replaces the user's MPI operator
 * around reference src/svds/primme_svds_c.c:1323-1383. */
typedef struct primme_b200_svds_dist {
   pb200_dist_csr *A;
   pb200_dist_csr *At;
} primme_b200_svds_dist;
void primme_b200_svds_dist_csr_matvec(void *x, int64_t *ldx, void *y, int64_t *ldy, int *blockSize,
      int *transpose, struct primme_svds_params *primme_svds, int *ierr);

#ifdef __cplusplus
}
#endif

#endif /* PRIMME_B200_H */

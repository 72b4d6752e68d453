/* pb_host.h -- internal declarations of the host control code (C99).
 *
 * The host side keeps PRIMME's structure: a front end that validates primme_params
 * (front.c), the outer block-Davidson iteration (davidson.c) and the small projected problem on
 * the host with LAPACK (hostla.c).  Every n-long operation goes through the C-ABI of
 * include/primme_b200.h, i.e. the sm_100a kernels in the product library.
 */
#ifndef PB_HOST_H
#define PB_HOST_H

#include "../../include/primme.h"
#include "../../include/primme_b200.h"
#include "hostla.h"

/* ---- typed names: every function below that touches SCALAR data exists as name_d and name_z ---- */
#define pb_solver PB_SUF(pb_solver)
#define pb_main_iter PB_SUF(pb_main_iter)
#define pb_apply_matvec PB_SUF(pb_apply_matvec)
#define pb_apply_precond PB_SUF(pb_apply_precond)
#define pb_global_sum PB_SUF(pb_global_sum)
#define pb_bcast PB_SUF(pb_bcast)
#define pb_bcast_int PB_SUF(pb_bcast_int)
#define pb_conv_test PB_SUF(pb_conv_test)
#define pb_monitor PB_SUF(pb_monitor)
#define pb_fill_random PB_SUF(pb_fill_random)
#define pb_update_projection PB_SUF(pb_update_projection)
#define pb_update_XKinvBX PB_SUF(pb_update_XKinvBX)
#define pb_skew_evecs_after_restart PB_SUF(pb_skew_evecs_after_restart)
#define pb_solve_H PB_SUF(pb_solve_H)
#define pb_map_vecs PB_SUF(pb_map_vecs)
#define pb_reduce_panel PB_SUF(pb_reduce_panel)
#define pb_ortho_block PB_SUF(pb_ortho_block)
#define pb_ortho_block_p0 PB_SUF(pb_ortho_block_p0)
#define pb_update_cholesky PB_SUF(pb_update_cholesky)
#define pb_update_cholesky_gram PB_SUF(pb_update_cholesky_gram)
#define pb_ortho_block_R PB_SUF(pb_ortho_block_R)
#define pb_ortho_single_iteration PB_SUF(pb_ortho_single_iteration)
#define pb_ortho_local PB_SUF(pb_ortho_local)
#define pb_update_Q PB_SUF(pb_update_Q)
#define pb_solve_H_ref PB_SUF(pb_solve_H_ref)
#define pb_prepare_vecs PB_SUF(pb_prepare_vecs)
#define pb_restart_refined PB_SUF(pb_restart_refined)
#define pb_update_QtV PB_SUF(pb_update_QtV)
#define pb_solve_H_harm PB_SUF(pb_solve_H_harm)
#define pb_restart_harmonic PB_SUF(pb_restart_harmonic)
#define pb_compute_submatrix PB_SUF(pb_compute_submatrix)
#define pb_ortho_local_R PB_SUF(pb_ortho_local_R)
#define pb_dyn_switch_from_jdqmr PB_SUF(pb_dyn_switch_from_jdqmr)
#define pb_dyn_switch_from_gdpk PB_SUF(pb_dyn_switch_from_gdpk)
#define pb_inner_solve PB_SUF(pb_inner_solve)
#define pb_restart PB_SUF(pb_restart)
#define pb_check_convergence PB_SUF(pb_check_convergence)
#ifdef PB_COMPLEX
/* the kernel layer's complex entry points (include/primme_b200.h): same argument lists, interleaved
 * (re,im) data, leading dimensions in complex elements */
#define pb200_dortho_sweep pb200_zortho_sweep
#define pb200_dvwxr pb200_zvwxr
#define pb200_dvwxr_can_fuse_gram pb200_zvwxr_can_fuse_gram
#define pb200_dpermute_columns pb200_zpermute_columns
#define pb200_dcopy_columns pb200_zcopy_columns
#define pb200_daxpy_columns pb200_zaxpy_columns
#define pb200_dscale_columns pb200_zscale_columns
#define pb200_dcolumn_dots pb200_zcolumn_dots
#define pb200_dresidual_inplace pb200_zresidual_inplace
#define pb200_djacobi pb200_zjacobi
#define pb200_dqmr_update pb200_zqmr_update
#endif
#define PB_DP(p) ((double *)(p)) /* pb200_cols / pb200_vwxr_out carry untyped column pointers */

/* convergence flags (reference src/eigs/common_eigs.h:41-46) */
enum { UNCONVERGED = 0, SKIP_UNTIL_RESTART = 1, CONVERGED = 2, PRACTICALLY_CONVERGED = 3 };

void primme_set_defaults(primme_params *primme);
void primme_display_params_prefix(const char *prefix, primme_params primme);

/* run-time cost model of PRIMME_DYNAMIC (reference main_iter_private.h:60-110) */
typedef struct pb_cost_model {
   double MV_PR, MV, PR, qmr_only, qmr_plus_MV_PR, gdk_plus_MV_PR, gdk_plus_MV, project_locked, reortho_locked;
   double gdk_conv_rate, jdq_conv_rate, JDQMR_slowdown, ratio_MV_outer;
   int nextReset;
   double gdk_sum_logResReductions, jdq_sum_logResReductions, gdk_sum_MV, jdq_sum_MV;
   int nevals_by_gdk, nevals_by_jdq;
   PRIMME_INT numIt_0, numMV_0;
   double timer_0, time_in_inner, resid_0;
   double accum_jdq_gdk, accum_jdq, accum_gdk;
} pb_cost_model;

/* Solver state shared by the pieces of the outer iteration. */
typedef struct pb_solver {
   primme_params *primme;
   pb200_ctx *dev;
   int device_callbacks; /* 1: user callbacks take device pointers (cublas_dprimme contract) */
   int64_t n;            /* nLocal */
   int64_t ld;           /* leading dimension of V and W (ldOPs) */
   SCALAR *V, *W;        /* device, ld x maxBasisSize */
   SCALAR *evecs;        /* device, ldevecs x (numOrthoConst + max(numEvals, initSize)) */
   int64_t ldevecs;
   SCALAR *hstage;       /* host staging for host callbacks / random vectors: n x maxBlockSize */
   SCALAR *hstage2;
   int hstage_cols;
   /* replicated small matrices (host) */
   int maxBasis, maxRank;
   SCALAR *H;         /* maxBasis x maxBasis, upper triangle of V'AV */
   SCALAR *hVecs;     /* maxBasis x maxBasis (leading dimension maxBasis) */
   SCALAR *prevhVecs; /* maxBasis x maxBasis */
   SCALAR *VtBV;      /* maxRank x maxRank or NULL (orth implicit) */
   SCALAR *fVtBV;     /* Cholesky factor of VtBV */
   double *hVals, *prevRitzVals, *blockNorms, *basisNorms;
   int *flags, *map, *iev, *perm, *lockedFlags;
   double t0;
   double tProj, tSolveH, tRestart; /* wall-clock of the phases the reference does not time (PB200_DEBUG report) */
   int numPrevRitzVals;
   /* first Gram panel of the next block orthogonalisation, delivered by the candidates sweep when
    * the new block is the residual block itself (no preconditioner, no locked vectors):
    * fusedP = [V(:,0:m) R]' R, (m + nb) x nb, leading dimension maxBasis + 8 */
   SCALAR *fusedP;
   int fusedP_m, fusedP_nb; /* fusedP_nb > 0: valid for basis size m and a block of nb columns */
   int fuse_allowed, fuse_enabled, fuse_sweeps;
   /* inner QMR solver (JDQMR family): g, d, delta, w, sol -- 5 x ld x maxBlockSize, device */
   SCALAR *jd_work;
   int touch; /* stopping-criterion state of the inner solver (main_iter.c:206,597-599) */
   /* skew-Q projector with a preconditioner (PRIMME_JDQR + applyPreconditioner; main_iter.c:324-333,
    * correction.c:948-954): evecsHat = K^{-1} evecs column by column, M = evecs' evecsHat (upper triangle) and its
    * Bunch-Kaufman factors, maintained by pb_update_XKinvBX (factorize.c:183-235) */
   SCALAR *evecsHat;      /* device, ld x maxEvecsSize, or NULL */
   SCALAR *Mskew, *Mfact; /* host, maxEvecsSize x maxEvecsSize (Mfact packed with leading dimension nM) */
   int *ipivot;
   int maxEvecsSize, numConvergedStored; /* numConvergedStored: Ritz vectors kept in evecs without locking (:197) */
   /* refined extraction (dav_refined.c): (A - tau I) V = Q R next to V and W */
   int refined;        /* primme_proj_refined */
   int numQR;          /* refined or harmonic: Q and R are carried (main_iter.c:268-273) */
   SCALAR *QtV;        /* harmonic: Q'V, maxBasis x maxBasis */
   SCALAR *Q;          /* device, ld x maxBasisSize */
   SCALAR *R, *hU, *hVecsRot; /* maxBasis x maxBasis */
   SCALAR *QtQ, *fQtQ; /* Q'Q and its Cholesky factor when orth is explicit, else NULL */
   double *hSVals;     /* singular values of R */
   int numArbitraryVecs;
   pb_cost_model cost; /* PRIMME_DYNAMIC */
   double tstart;      /* start of the last correction / restart evaluation (main_iter.c:260,655,1182) */
} pb_solver;

/* error propagation in the style of the reference's CHKERR (common.h:484-494) */
#define CHK(call)                                                                        \
   do {                                                                                  \
      int chk_err_ = (call);                                                             \
      if (chk_err_ != 0) {                                                               \
         pb_report(S ? S->primme : NULL, __FILE__, __LINE__, chk_err_, #call);           \
         return chk_err_;                                                                \
      }                                                                                  \
   } while (0)
/* same, releasing the function's temporaries first */
#define CHKX(call, cleanup)                                                              \
   do {                                                                                  \
      int chk_err_ = (call);                                                             \
      if (chk_err_ != 0) {                                                               \
         cleanup;                                                                        \
         pb_report(S ? S->primme : NULL, __FILE__, __LINE__, chk_err_, #call);           \
         return chk_err_;                                                                \
      }                                                                                  \
   } while (0)
void pb_report(primme_params *primme, const char *file, int line, int err, const char *what);

/* davidson.c */
int pb_main_iter(pb_solver *S, double *evals, double *resNorms, int *ret, int *numRet);
/* front.c helpers used by davidson.c */
int pb_apply_matvec(pb_solver *S, SCALAR *Vblk, int64_t ldv, SCALAR *Wblk, int64_t ldw, int bs);
int pb_apply_precond(pb_solver *S, SCALAR *X, int64_t ldx, SCALAR *Y, int64_t ldy, int bs);
int pb_global_sum(pb_solver *S, double *buf, int count);
int pb_bcast(pb_solver *S, double *buf, int count);
int pb_bcast_int(pb_solver *S, int *buf, int count);
int pb_conv_test(pb_solver *S, double eval, const SCALAR *evec, double rnorm, int *isconv);
int pb_monitor(pb_solver *S, double *basisEvals, int basisSize, int *basisFlags, int *iblock,
      int blockSize, double *basisNorms, int numConverged, double *lockedEvals, int numLocked,
      int *lockedFlags, double *lockedNorms, int inner_its, double LSRes, const char *msg,
      double time, primme_event event);
double pb_problem_norm(int overrideUserEstimations, primme_params *primme);
int pb_fill_random(pb_solver *S, SCALAR *X, int64_t ldx, int ncols);


/* operators.c */
void pb_registry_set_solver(const primme_params *primme, pb200_ctx *ctx);

/* dav_project.c */
int pb_update_projection(pb_solver *S, int numCols, int blockSize);
int pb_update_XKinvBX(pb_solver *S, int numCols, int blockSize);
int pb_skew_evecs_after_restart(pb_solver *S, int numConverged);
int pb_solve_H(pb_solver *S, const SCALAR *H, int ldH, int n, const SCALAR *VtBVblk, int ldVtBV,
      SCALAR *hVecs, int ldhVecs, double *hVals, int numConverged, int updateStats);
int pb_map_vecs(const SCALAR *V, int m, int nV, int ldV, const SCALAR *W, int n0, int n, int ldW,
      int *p);
int pb_reduce_panel(pb_solver *S, SCALAR *P, int rows, int cols, int ldp);

/* dav_ortho.c */
int pb_ortho_block(pb_solver *S, SCALAR *V, int64_t ldV, int b1, int b2, const SCALAR *locked,
      int64_t ldLocked, int numLocked, SCALAR *RLocked, int ldRLocked, int *b2_out);
/* same, with the panel of the first sweep already known (P0 = [V(:,0:b1) X]' X, or NULL) */
int pb_ortho_block_p0(pb_solver *S, SCALAR *V, int64_t ldV, int b1, int b2, const SCALAR *locked,
      int64_t ldLocked, int numLocked, SCALAR *RLocked, int ldRLocked, int *b2_out, const SCALAR *P0,
      int ldP0);
int pb_update_cholesky(pb_solver *S, int n0, int n);
int pb_update_cholesky_gram(const SCALAR *G, SCALAR *fG, int ld, int n0, int n);
int pb_ortho_block_R(pb_solver *S, SCALAR *Q, int64_t ldQ, SCALAR *QtQ, SCALAR *fQtQ, int ldQtQ, int maxRank,
      SCALAR *R, int ldR, int b1, int b2, int *b2_out);
int pb_ortho_single_iteration(pb_solver *S, const SCALAR *Q, int nQ, int64_t ldQ,
      const SCALAR *QtQ, int ldQtQ, SCALAR *X, const int *inX, int nX, int64_t ldX,
      double *norms);
int pb_ortho_local(SCALAR *V, int ldV, SCALAR *R, int b1, int b2, SCALAR *locked, int ldLocked,
      int numLocked, int n, const SCALAR *B, int ldB, long long *iseed);

/* dav_refined.c */
int pb_update_Q(pb_solver *S, double shift, int basisSize, int blockSize, int *nQ);
int pb_solve_H_ref(pb_solver *S, int n, const SCALAR *VtBVblk, int ldVtBV, int numConverged);
int pb_prepare_vecs(pb_solver *S, int basisSize, int i0, int blockSize, int targetShiftIndex, int *arbitraryVecs,
      double smallestResNorm, const int *flags, int RRForAll);
int pb_restart_refined(pb_solver *S, int restartSize, int basisSize, int numConverged, int numPrevRetained,
      int indexOfPreviousVecs, int indexOfPreviousVecsBeforeRestart, const int *restartPerm, const int *hVecsPerm,
      int *targetShiftIndex);
int pb_update_QtV(pb_solver *S, int numCols, int blockSize);
int pb_solve_H_harm(pb_solver *S, int n, const SCALAR *VtBVblk, int ldVtBV, int numConverged);
int pb_restart_harmonic(pb_solver *S, int restartSize, int basisSize, int numConverged, int *targetShiftIndex);
int pb_compute_submatrix(const SCALAR *X, int nX, int ldX, const SCALAR *H, int nH, int ldH, SCALAR *R, int ldR);
int pb_ortho_local_R(SCALAR *V, int ldV, SCALAR *R, int ldR, int b1, int b2, int n, const SCALAR *B, int ldB,
      long long *iseed);

/* dav_dynamic.c */
void pb_dyn_init(pb_cost_model *m, primme_params *primme);
int pb_dyn_update_statistics(pb_cost_model *m, primme_params *primme, double current_time, int recentConv,
      int calledAtRestart, int numConverged, double currentResNorm);
int pb_dyn_switch_from_jdqmr(pb_solver *S, pb_cost_model *m);
int pb_dyn_switch_from_gdpk(pb_solver *S, pb_cost_model *m);
void pb_dyn_recommend(primme_params *primme, const pb_cost_model *m);

/* dav_jdqmr.c */
int pb_inner_solve(pb_solver *S, int blockSize, SCALAR *x, int64_t ldx, SCALAR *r, int64_t ldr, const double *rnorm,
      const SCALAR *Q, int64_t ldQ, int nQ, int useX, SCALAR *sol, int64_t ldsol, const double *eval, double *shift,
      int *touch, SCALAR *work, const SCALAR *RQ, int64_t ldRQ, int nRQ, SCALAR *RX, int64_t ldRX, SCALAR *xKinvBx,
      const SCALAR *skewQ, int64_t ldskewQ, const SCALAR *Mfact, const int *ipivot);

/* dav_restart.c */
int pb_restart(pb_solver *S, int basisSize, int *ievSize, double *evals, double *resNorms,
      int *numConverged, int *numLocked, int nprevhVecs, int numGuesses, int *restartSizeOut,
      int *targetShiftIndex, int *restartsSinceReset);
int pb_check_convergence(pb_solver *S, SCALAR *X, int64_t ldX, int givenX, SCALAR *R,
      int64_t ldR, int givenR, int numLocked, int left, int right, int *flags,
      double *blockNorms, double *hVals, int *reset, int practConvCheck);
int pb_insertion_sort(double newVal, double *evals, double newNorm, double *resNorms,
      int newFlag, int *flags, int *perm, int n, int initialShift, primme_params *primme);

#endif

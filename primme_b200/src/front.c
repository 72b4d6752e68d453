/* front.c -- public solver entry points, input validation, callback plumbing.
 *
 * Restates reference src/eigs/primme_c.c (Xprimme_aux :159-244, wrapper_Sprimme :278-422,
 * check_input :438-538, convTestFunAbsolute :555-570, default_monitor :602-721) and the callback
 * wrappers of src/eigs/auxiliary_eigs.c (matrixMatvec_ :183-230, applyPreconditioner_ :317-364,
 * globalSum_ :369-427, broadcast_ :430-480, problemNorm :567-591).
 *
 * Two flavours, same solver:
 *   dprimme         evecs and the callback blocks are HOST arrays (reference dprimme contract).
 *                   The basis lives in HBM; each user matvec sees a host copy of the block
 *                   (D2H, callback, H2D).  Correct drop-in, not the fast path.
 *   cublas_dprimme  the reference's device contract (include/primme_eigs.h:414-417,
 *                   examples/ex_eigs_dcublas.c): evecs and callback blocks are DEVICE pointers.
 * There is no CPU fallback: without a GPU both return PRIMME_FUNCTION_UNAVAILABLE.
 */
#include "pb_host.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef PB_COMPLEX /* untyped helpers: defined once */
void pb_report(primme_params *primme, const char *file, int line, int err, const char *what) {
   if (getenv("PB200_DEBUG")) fprintf(stderr, "PRIMME-B200: error %d at %s:%d in '%s'\n", err, file, line, what);
   if (primme && primme->procID == 0 && primme->outputFile && primme->printLevel >= 1) {
      fprintf(primme->outputFile, "PRIMME-B200: error %d at %s:%d in '%s'\n", err, file, line, what);
      fflush(primme->outputFile);
   } else if (!primme) {
      fprintf(stderr, "PRIMME-B200: error %d at %s:%d in '%s'\n", err, file, line, what);
   }
}

double pb_problem_norm(int overrideUserEstimations, primme_params *primme) {
   if (!overrideUserEstimations)
      return primme->aNorm > 0.0 ? primme->aNorm : primme->stats.estimateLargestSVal;
   return PB_MAX(primme->aNorm > 0.0 ? primme->aNorm : 0.0, primme->stats.estimateLargestSVal);
}
#endif

/* ------------------------------------------------------------------------- callbacks ---- */
static int host_block(pb_solver *S, int cols) {
   if (cols <= S->hstage_cols) return 0;
   free(S->hstage), free(S->hstage2);
   size_t bytes = sizeof(SCALAR) * (size_t)(S->n > 0 ? S->n : 1) * cols;
   S->hstage = (SCALAR *)malloc(bytes);
   S->hstage2 = (SCALAR *)malloc(bytes);
   S->hstage_cols = cols;
   return (S->hstage && S->hstage2) ? 0 : PRIMME_MALLOC_FAILURE;
}

static int call_block_op(pb_solver *S, primme_block_op_fn fn, SCALAR *X, int64_t ldx, SCALAR *Y,
      int64_t ldy, int bs, const char *what) {
   primme_params *primme = S->primme;
   int ierr = 0;
   if (S->device_callbacks) {
      PRIMME_INT lx = ldx, ly = ldy;
      fn(X, &lx, Y, &ly, &bs, primme, &ierr);
   } else {
      /* host contract: stage the block through host memory */
      CHK(host_block(S, bs));
      PRIMME_INT ln = S->n;
      CHK(pb200_copy_d2h(S->dev, X, ldx, S->hstage, S->n, S->n, bs, PB_ES));
      fn(S->hstage, &ln, S->hstage2, &ln, &bs, primme, &ierr);
      if (!ierr) CHK(pb200_copy_h2d(S->dev, S->hstage2, S->n, Y, ldy, S->n, bs, PB_ES));
   }
   if (ierr != 0) {
      pb_report(primme, __FILE__, __LINE__, ierr, what);
      return PRIMME_USER_FAILURE;
   }
   return 0;
}

int pb_apply_matvec(pb_solver *S, SCALAR *Vblk, int64_t ldv, SCALAR *Wblk, int64_t ldw, int bs) {
   primme_params *primme = S->primme;
   if (bs <= 0) return 0;
   const double t0 = hl_wtime();
   /* the stream is a blocking stream: default-stream work inside the callback is ordered
    * with ours; callbacks using their own non-blocking streams must synchronise themselves */
   CHK(call_block_op(S, primme->matrixMatvec, Vblk, ldv, Wblk, ldw, bs, "matrixMatvec"));
   primme->stats.timeMatvec += hl_wtime() - t0;
   primme->stats.numMatvecs += bs;
   return 0;
}

int pb_apply_precond(pb_solver *S, SCALAR *X, int64_t ldx, SCALAR *Y, int64_t ldy, int bs) {
   primme_params *primme = S->primme;
   if (bs <= 0) return 0;
   const double t0 = hl_wtime();
   if (primme->correctionParams.precondition) {
      CHK(call_block_op(S, primme->applyPreconditioner, X, ldx, Y, ldy, bs, "applyPreconditioner"));
      primme->stats.numPreconds += bs;
   } else {
      CHK(pb200_copy_d2d(S->dev, X, ldx, Y, ldy, S->n, bs, PB_ES));
   }
   primme->stats.timePrecond += hl_wtime() - t0;
   return 0;
}

int pb_global_sum(pb_solver *S, double *buf, int count) {
   primme_params *primme = S->primme;
   if (primme->numProcs <= 1 || count <= 0) return 0;
   const double t0 = hl_wtime();
   int ierr = 0;
   if (primme->globalSumReal) {
      primme->globalSumReal(buf, buf, &count, primme, &ierr); /* in place is allowed (:413) */
      if (ierr) return PRIMME_USER_FAILURE;
   } else if (pb200_ctx_nranks(S->dev) > 1) {
      CHK(pb200_allreduce_host(S->dev, buf, count));
   } else {
      return PRIMME_PARALLEL_FAILURE;
   }
   primme->stats.numGlobalSum++;
   primme->stats.volumeGlobalSum += count;
   primme->stats.timeGlobalSum += hl_wtime() - t0;
   return 0;
}

int pb_bcast(pb_solver *S, double *buf, int count) {
   primme_params *primme = S->primme;
   if (primme->numProcs <= 1 || count <= 0) return 0;
   int ierr = 0;
   if (primme->broadcastReal) {
      primme->broadcastReal(buf, &count, primme, &ierr);
      return ierr ? PRIMME_USER_FAILURE : 0;
   }
   /* zero on the others + global sum (auxiliary_eigs.c:444-466) */
   if (primme->procID != 0) memset(buf, 0, sizeof(double) * count);
   return pb_global_sum(S, buf, count);
}

int pb_bcast_int(pb_solver *S, int *buf, int count) {
   if (S->primme->numProcs <= 1 || count <= 0) return 0;
   double *tmp = (double *)malloc(sizeof(double) * count);
   for (int i = 0; i < count; i++) tmp[i] = buf[i];
   int rc = pb_bcast(S, tmp, count);
   for (int i = 0; i < count; i++) buf[i] = (int)tmp[i];
   free(tmp);
   return rc;
}

static void conv_test_absolute(double *eval, void *evec, double *rNorm, int *isConv,
      primme_params *primme, int *ierr);

/* convTestFun_ (auxiliary_eigs_normal.c:408-443): the vector, when the caller has it, is handed to
 * the callback in the memory space of its contract -- as is for cublas_dprimme, through a host copy
 * for dprimme (skipped for the built-in test, which ignores it) */
int pb_conv_test(pb_solver *S, double eval, const SCALAR *evec, double rnorm, int *isconv) {
   primme_params *primme = S->primme;
   int ierr = 0;
   void *v = (void *)evec;
   if (evec && !S->device_callbacks) {
      v = NULL;
      if (primme->convTestFun != conv_test_absolute) {
         CHK(host_block(S, 1));
         CHK(pb200_copy_d2h(S->dev, evec, S->n, S->hstage, S->n, S->n, 1, PB_ES));
         v = S->hstage;
      }
   }
   primme->convTestFun(&eval, v, &rnorm, isconv, primme, &ierr);
   if (ierr) {
      pb_report(primme, __FILE__, __LINE__, ierr, "convTestFun");
      return PRIMME_UNEXPECTED_FAILURE;
   }
   return 0;
}

int pb_monitor(pb_solver *S, double *basisEvals, int basisSize, int *basisFlags, int *iblock,
      int blockSize, double *basisNorms, int numConverged, double *lockedEvals, int numLocked,
      int *lockedFlags, double *lockedNorms, int inner_its, double LSRes, const char *msg,
      double time, primme_event event) {
   primme_params *primme = S->primme;
   if (!primme->monitorFun) return 0;
   primme->stats.elapsedTime = hl_wtime() - S->t0;
   int err = 0;
   primme->monitorFun(basisEvals, &basisSize, basisFlags, iblock, &blockSize, basisNorms,
         &numConverged, lockedEvals, &numLocked, lockedFlags, lockedNorms,
         inner_its >= 0 ? &inner_its : NULL, LSRes >= 0 ? &LSRes : NULL, msg, &time, &event,
         primme, &err);
   if (err) {
      pb_report(primme, __FILE__, __LINE__, err, "monitorFun");
      return PRIMME_UNEXPECTED_FAILURE;
   }
   return 0;
}

/* Random columns exactly as the reference generates them, even in its GPU builds
 * (cublas_wrapper.c:707-720): LAPACK dlarnv(2) on the host with the evolving primme.iseed,
 * then one upload. */
int pb_fill_random(pb_solver *S, SCALAR *X, int64_t ldx, int ncols) {
   primme_params *primme = S->primme;
   if (ncols <= 0) return 0;
   /* the same numbers as the reference's Num_larnv (dlarnv, idist 2), drawn on the device */
   long long seed[4];
   for (int i = 0; i < 4; i++) seed[i] = primme->iseed[i];
   const int w = (int)(sizeof(SCALAR) / sizeof(double));
   int in_range = 1;
   for (int i = 0; i < 4; i++) in_range = in_range && seed[i] >= 0 && seed[i] <= 4095;
   if (in_range) {
      CHK(pb200_dlarnv(S->dev, seed, (int64_t)w * S->n, ncols, (double *)X, (int64_t)w * ldx));
      for (int i = 0; i < 4; i++) primme->iseed[i] = seed[i];
      return 0;
   }
   /* seeds outside dlarnv's documented range: whatever LAPACK makes of them, on the host */
   CHK(host_block(S, ncols));
   for (int j = 0; j < ncols; j++) hl_larnv2(seed, S->n, S->hstage + (size_t)S->n * j);
   for (int i = 0; i < 4; i++) primme->iseed[i] = seed[i];
   return pb200_copy_h2d(S->dev, S->hstage, S->n, X, ldx, S->n, ncols, PB_ES);
}

/* ------------------------------------------------------------------- default callbacks -- */
static void conv_test_absolute(double *eval, void *evec, double *rNorm, int *isConv,
      primme_params *primme, int *ierr) {
   (void)eval, (void)evec;
   *isConv = *rNorm < PB_MAX(primme->eps, PB_EPS * 2) * pb_problem_norm(0, primme);
   *ierr = 0;
}

static void default_monitor(void *basisEvals_, int *basisSize, int *basisFlags, int *iblock,
      int *blockSize, void *basisNorms_, int *numConverged, void *lockedEvals_, int *numLocked,
      int *lockedFlags, void *lockedNorms_, int *inner_its, void *LSRes_, const char *msg,
      double *time, primme_event *event, primme_params *primme, int *err) {
   (void)basisSize, (void)basisFlags, (void)inner_its;
   double *basisEvals = (double *)basisEvals_, *basisNorms = (double *)basisNorms_;
   double *lockedEvals = (double *)lockedEvals_, *lockedNorms = (double *)lockedNorms_;
   FILE *f = primme->outputFile;
   *err = 0;
   if (!f || !(primme->procID == 0 || *event == primme_event_profile)) return;
   switch (*event) {
   case primme_event_outer_iteration:
      if (primme->printLevel >= 3) {
         int found = primme->locking ? *numLocked : *numConverged;
         for (int i = 0; i < *blockSize; i++)
            fprintf(f, "OUT %" PRIMME_INT_P " conv %d blk %d MV %" PRIMME_INT_P
                       " Sec %E EV %13E |r| %.3E\n",
                  primme->stats.numOuterIterations, found, i, primme->stats.numMatvecs,
                  primme->stats.elapsedTime, basisEvals[iblock[i]], basisNorms[iblock[i]]);
      }
      break;
   case primme_event_inner_iteration:
      if (primme->printLevel >= 4)
         fprintf(f, "INN MV %" PRIMME_INT_P " Sec %e Eval %13E Lin|r| %.3e EV|r| %.3e\n",
               primme->stats.numMatvecs, primme->stats.elapsedTime, basisEvals[iblock[0]],
               *(double *)LSRes_, basisNorms[iblock[0]]);
      break;
   case primme_event_converged:
      if ((!primme->locking && primme->printLevel >= 2) || (primme->locking && primme->printLevel >= 5))
         fprintf(f, "#Converged %d eval[ %d ]= %13E norm %e Mvecs %" PRIMME_INT_P " Time %g\n",
               *numConverged, iblock[0], basisEvals[iblock[0]], basisNorms[iblock[0]],
               primme->stats.numMatvecs, primme->stats.elapsedTime);
      break;
   case primme_event_locked:
      if (primme->printLevel >= 2)
         fprintf(f, "Lock epair[ %d ]= %13E norm %.4e Mvecs %" PRIMME_INT_P
                    " Time %.4e Flag %d\n",
               *numLocked - 1, lockedEvals[*numLocked - 1], lockedNorms[*numLocked - 1],
               primme->stats.numMatvecs, primme->stats.elapsedTime, lockedFlags[*numLocked - 1]);
      break;
   case primme_event_message:
      if (primme->printLevel >= 2 && msg) fprintf(f, "%s\n", msg);
      break;
   case primme_event_profile:
      if (msg && time) {
         if (primme->printLevel >= 3 && *time < 0.0)
            fprintf(f, "entering in %s proc %d\n", msg, primme->procID);
         if (primme->printLevel >= 2 && *time >= 0.0)
            fprintf(f, "time %g for %s proc %d\n", *time, msg, primme->procID);
      }
      break;
   default: break;
   }
   fflush(f);
}

/* ------------------------------------------------------------------ input validation ---- */
/* same codes as the reference (primme_c.c:438-538) */
static int check_input(void *evals, void *evecs, void *resNorms, primme_params *p) {
   if (p == NULL) return -4;
   if (p->n < 0 || p->nLocal < 0 || p->nLocal > p->n) return -5;
   if (p->numProcs < 1) return -6;
   if (p->matrixMatvec == NULL) return -7;
   if (p->applyPreconditioner == NULL && p->correctionParams.precondition > 0) return -8;
   if (p->numEvals > p->n) return -10;
   if (p->numEvals < 0) return -11;
   if (p->convTestFun != NULL && fabs(p->eps) != 0.0 && p->eps < PB_EPS) return -12;
   if (p->target != primme_smallest && p->target != primme_largest &&
         p->target != primme_largest_abs && p->target != primme_closest_geq &&
         p->target != primme_closest_leq && p->target != primme_closest_abs)
      return -13;
   if (p->numOrthoConst < 0 || p->numOrthoConst > p->n) return -16;
   if (p->maxBasisSize < 2 && p->n > 2) return -17;
   if (p->minRestartSize < 0 || (p->minRestartSize == 0 && p->n > 2 && p->numEvals > 0)) return -18;
   if (p->maxBlockSize < 0 || (p->maxBlockSize == 0 && p->numEvals > 0)) return -19;
   if (p->restartingParams.maxPrevRetain < 0) return -20;
   if (p->initSize < 0) return -22;
   if (p->locking == 0 && p->initSize > p->maxBasisSize) return -23;
   if (p->locking > 0 && p->initSize > p->numEvals) return -24;
   if (p->minRestartSize + p->restartingParams.maxPrevRetain >= p->maxBasisSize &&
         p->n > p->maxBasisSize)
      return -25;
   if (p->minRestartSize > p->n && p->n > 2) return -26;
   if (p->printLevel < 0 || p->printLevel > 5) return -27;
   if (p->correctionParams.convTest != primme_full_LTolerance &&
         p->correctionParams.convTest != primme_decreasing_LTolerance &&
         p->correctionParams.convTest != primme_adaptive_ETolerance &&
         p->correctionParams.convTest != primme_adaptive)
      return -28;
   if (p->correctionParams.convTest == primme_decreasing_LTolerance &&
         p->correctionParams.relTolBase <= 1.0)
      return -29;
   if (evals == NULL) return -30;
   if (evecs == NULL) return -31;
   if (resNorms == NULL) return -32;
   if (p->locking == 0 && p->minRestartSize < p->numEvals && p->n > 2) return -33;
   if (p->ldevecs < p->nLocal) return -34;
   if (p->ldOPs != 0 && p->ldOPs < p->nLocal) return -35;
   if (p->locking == 0 && (p->target == primme_closest_leq || p->target == primme_closest_geq))
      return -38;
   if (p->massMatrixMatvec && p->projectionParams.projection != primme_proj_RR) return -39;
   if (p->target == primme_largest_abs || p->target == primme_closest_geq ||
         p->target == primme_closest_leq || p->target == primme_closest_abs) {
      if (p->numTargetShifts <= 0) return -14;
      if (p->targetShifts == NULL) return -15;
   }
   return 0;
}

/* Features of the reference outside the hot-path scope of this library. */
static int check_scope(primme_params *p) {
   const char *why = NULL;
   if (p->massMatrixMatvec) why = "generalized problems (massMatrixMatvec)";
   else if (p->projectionParams.projection == primme_proj_harmonic && p->target != primme_closest_geq &&
            p->target != primme_closest_leq && p->target != primme_closest_abs)
      why = "harmonic extraction with a target other than closest_geq / closest_leq / closest_abs";
   else if (p->internalPrecision != primme_op_default && p->internalPrecision != primme_op_double)
      why = "internalPrecision other than double";
   else if ((p->matrixMatvec_type != primme_op_default && p->matrixMatvec_type != primme_op_double) ||
            (p->applyPreconditioner && p->applyPreconditioner_type != primme_op_default &&
                  p->applyPreconditioner_type != primme_op_double) ||
            (p->globalSumReal && p->globalSumReal_type != primme_op_default &&
                  p->globalSumReal_type != primme_op_double))
      why = "callback datatypes other than double";
   /* blocks wider than the kernels' 8-column panels are processed in chunks of 8 (dav_ortho.c, dav_project.c,
    * the launchers); the inner QMR solver keeps its 8 systems per block */
#ifdef PB_COMPLEX
   else if (p->projectionParams.projection != primme_proj_default && p->projectionParams.projection != primme_proj_RR)
      why = "refined / harmonic extraction in complex arithmetic (Rayleigh-Ritz only)";
#endif
   if (!why) return 0;
   if (p->outputFile && p->printLevel >= 1 && p->procID == 0)
      fprintf(p->outputFile, "PRIMME-B200: %s is outside the scope of this build\n", why);
   return PRIMME_FUNCTION_UNAVAILABLE;
}

static void free_solver(pb_solver *S, int own_evecs) {
   if (S->dev) {
      pb200_free(S->dev, S->V), pb200_free(S->dev, S->W);
      if (own_evecs) pb200_free(S->dev, S->evecs);
   }
   free(S->hstage), free(S->hstage2);
   free(S->H), free(S->hVecs), free(S->prevhVecs), free(S->VtBV), free(S->fVtBV), free(S->fusedP);
   free(S->R), free(S->hU), free(S->hVecsRot), free(S->QtQ), free(S->fQtQ), free(S->hSVals), free(S->QtV);
   if (S->dev && S->Q) pb200_free(S->dev, S->Q);
   if (S->dev && S->jd_work) pb200_free(S->dev, S->jd_work);
   if (S->dev && S->evecsHat) pb200_free(S->dev, S->evecsHat);
   free(S->Mskew), free(S->Mfact), free(S->ipivot);
   free(S->hVals), free(S->prevRitzVals), free(S->blockNorms), free(S->basisNorms);
   free(S->flags), free(S->map), free(S->iev), free(S->perm), free(S->lockedFlags);
   if (S->dev) pb200_ctx_destroy(S->dev);
}

/* The body shared by dprimme and cublas_dprimme (wrapper_Sprimme, primme_c.c:278-422). */
static int solve_typed(double *evals, SCALAR *evecs, double *resNorms, primme_params *primme,
      int device_mode) {
   pb_solver Sv, *S = &Sv;
   memset(S, 0, sizeof(*S));
   if (!primme) return -4;
   S->primme = primme;
   S->device_callbacks = device_mode;
   S->t0 = hl_wtime();
   const int outInitSize0 = 0;
   (void)outInitSize0;

   /* defaults for sequential programs and for members left at their sentinels */
   if (primme->numProcs <= 1 && evals && evecs && resNorms) {
      primme->nLocal = primme->n;
      primme->procID = 0;
   }
   primme_set_defaults(primme);
   if (primme->orth == primme_orth_default)
      primme->orth = primme->maxBlockSize > 1 ? primme_orth_explicit_I : primme_orth_implicit_I;
   /* primme.ldOPs as the reference leaves it (primme_c.c:325-332: nLocal).  The basis itself is
    * free to use a padded leading dimension (16 elements: every column starts on a 128-byte
    * boundary for bulk copies / vector loads); that choice lives in S->ld only, so the caller's
    * struct can be reused for a later, larger problem */
   const int ld_default = primme->ldOPs == -1 || primme->ldOPs == 0;
   if (ld_default) primme->ldOPs = primme->nLocal;
   if (evals == NULL && evecs == NULL && resNorms == NULL) return 0;

   if (primme->iseed[0] < 0 || primme->iseed[0] > 4095) primme->iseed[0] = primme->procID % 4096;
   if (primme->iseed[1] < 0 || primme->iseed[1] > 4095) primme->iseed[1] = (int)(primme->procID / 4096 + 1) % 4096;
   if (primme->iseed[2] < 0 || primme->iseed[2] > 4095) primme->iseed[2] = (int)((primme->procID / 4096) / 4096 + 2) % 4096;
   if (primme->iseed[3] < 0 || primme->iseed[3] > 4095) primme->iseed[3] = (2 * (int)(((primme->procID / 4096) / 4096) / 4096) + 1) % 4096;

   if (!primme->convTestFun) {
      primme->convTestFun = conv_test_absolute;
      primme->convTestFun_type = primme_op_double;
      if (primme->eps == 0.0) primme->eps = PB_EPS * 1e4;
   }
   if (!primme->monitorFun) {
      primme->monitorFun = default_monitor;
      primme->monitorFun_type = primme_op_double;
   }
   if (primme->matrixMatvec && primme->matrixMatvec_type == primme_op_default) primme->matrixMatvec_type = primme_op_double;
   if (primme->applyPreconditioner && primme->applyPreconditioner_type == primme_op_default) primme->applyPreconditioner_type = primme_op_double;
   if (primme->globalSumReal && primme->globalSumReal_type == primme_op_default) primme->globalSumReal_type = primme_op_double;
   if (primme->broadcastReal && primme->broadcastReal_type == primme_op_default) primme->broadcastReal_type = primme_op_double;

   int rc = check_input(evals, evecs, resNorms, primme);
   if (rc) return rc;
   rc = check_scope(primme);
   if (rc) {
      primme->initSize = 0;
      return rc;
   }

   /* the device context: the caller's (primme_b200_attach_ctx, e.g. carrying an NCCL
    * communicator) or a private one for this solve */
   int own_ctx = 0;
   S->dev = primme_b200_attached_ctx(primme);
   if (!S->dev) {
      rc = pb200_ctx_create(&S->dev, -1);
      if (rc) {
         pb_report(primme, __FILE__, __LINE__, rc, "no CUDA device: this library has no CPU path");
         primme->initSize = 0;
         return PRIMME_FUNCTION_UNAVAILABLE;
      }
      own_ctx = 1;
   }
   if (device_mode && pb200_is_device_pointer(evecs) != 1) {
      if (own_ctx) pb200_ctx_destroy(S->dev);
      primme->initSize = 0;
      return -31;
   }
   pb_registry_set_solver(primme, S->dev);
   pb200_ctx_begin_solve(S->dev);

   S->n = primme->nLocal;
   S->ld = ld_default ? (primme->nLocal + 15) / 16 * 16 : primme->ldOPs;
   S->maxBasis = primme->maxBasisSize;
   S->maxRank = primme->numOrthoConst + primme->maxBasisSize + (primme->locking ? primme->numEvals : 0);
   const int mb = S->maxBasis, mr = S->maxRank;
   const int nevecs = primme->numOrthoConst + PB_MAX(primme->numEvals, primme->initSize);

   /* a caller-attached (long-lived) context keeps the basis arrays between solves */
   const size_t basis_bytes = sizeof(SCALAR) * (size_t)PB_MAX(S->ld, 1) * mb;
   if (own_ctx) {
      rc = pb200_malloc(S->dev, basis_bytes, (void **)&S->V);
      if (!rc) rc = pb200_malloc(S->dev, basis_bytes, (void **)&S->W);
   } else {
      rc = pb200_ctx_workspace(S->dev, 0, basis_bytes, (void **)&S->V);
      if (!rc) rc = pb200_ctx_workspace(S->dev, 1, basis_bytes, (void **)&S->W);
   }
   /* persisting L2 window over the head of V: the panel sweeps of the GD-type methods re-read V from its first
    * column every launch; with inner QMR iterations the hot set is the locked vectors instead, left to the L2 */
   if (!rc && primme->correctionParams.maxInnerIterations == 0 && primme->dynamicMethodSwitch <= 0)
      pb200_ctx_l2_persist(S->dev, S->V, basis_bytes);
   int own_evecs = 0;
   if (!rc) {
      if (device_mode) {
         S->evecs = evecs;
         S->ldevecs = primme->ldevecs;
      } else {
         S->ldevecs = PB_MAX(S->n, 1);
         rc = pb200_malloc(S->dev, sizeof(SCALAR) * (size_t)S->ldevecs * nevecs, (void **)&S->evecs);
         own_evecs = 1;
         if (!rc && primme->numOrthoConst + primme->initSize > 0)
            rc = pb200_copy_h2d(S->dev, evecs, primme->ldevecs, S->evecs, S->ldevecs, S->n,
                  primme->numOrthoConst + primme->initSize, PB_ES);
      }
   }
   S->H = (SCALAR *)calloc((size_t)mb * mb, sizeof(SCALAR));
   S->hVecs = (SCALAR *)calloc((size_t)mb * mb, sizeof(SCALAR));
   S->prevhVecs = (SCALAR *)calloc((size_t)mb * mb, sizeof(SCALAR));
   S->fusedP = (SCALAR *)calloc((size_t)(mb + 8) * 8, sizeof(SCALAR));
   /* the fused candidates sweep leaves residuals, not Ritz vectors, in the block: only the
    * built-in test (which ignores evec) may run on it */
   S->fuse_allowed = (getenv("PB200_NO_FUSE_GRAM") || primme->convTestFun != conv_test_absolute) ? 0 : 1;
   if (primme->orth == primme_orth_explicit_I) {
      S->VtBV = (SCALAR *)calloc((size_t)mr * mr, sizeof(SCALAR));
      S->fVtBV = (SCALAR *)calloc((size_t)mr * mr, sizeof(SCALAR));
   }
   S->refined = primme->projectionParams.projection == primme_proj_refined;
   S->numQR = primme->projectionParams.projection != primme_proj_RR;
   if (primme->projectionParams.projection == primme_proj_harmonic) {
      S->QtV = (SCALAR *)calloc((size_t)mb * mb, sizeof(SCALAR));
      if (!S->QtV) rc = PRIMME_MALLOC_FAILURE;
   }
   if (S->numQR) {
      /* Q next to V and W, its small factors on the host (main_iter.c:284-320) */
      if (!rc) rc = pb200_malloc(S->dev, basis_bytes, (void **)&S->Q);
      S->R = (SCALAR *)calloc((size_t)mb * mb, sizeof(SCALAR));
      S->hU = (SCALAR *)calloc((size_t)mb * mb, sizeof(SCALAR));
      S->hVecsRot = (SCALAR *)calloc((size_t)mb * mb, sizeof(SCALAR));
      S->hSVals = (double *)calloc(mb, sizeof(double));
      if (primme->orth == primme_orth_explicit_I) {
         S->QtQ = (SCALAR *)calloc((size_t)mb * mb, sizeof(SCALAR));
         S->fQtQ = (SCALAR *)calloc((size_t)mb * mb, sizeof(SCALAR));
      }
      if (!S->R || !S->hU || !S->hVecsRot || !S->hSVals) rc = PRIMME_MALLOC_FAILURE;
   }
   if (primme->correctionParams.precondition && primme->correctionParams.maxInnerIterations != 0 &&
         primme->correctionParams.projectors.RightQ && primme->correctionParams.projectors.SkewQ) {
      /* main_iter.c:324-333 */
      S->maxEvecsSize = primme->numOrthoConst + primme->numEvals;
      const size_t me = (size_t)PB_MAX(S->maxEvecsSize, 1);
      if (!rc) rc = pb200_malloc(S->dev, sizeof(SCALAR) * (size_t)PB_MAX(S->ld, 1) * me, (void **)&S->evecsHat);
      S->Mskew = (SCALAR *)calloc(me * me, sizeof(SCALAR));
      S->Mfact = (SCALAR *)calloc(me * me, sizeof(SCALAR));
      S->ipivot = (int *)calloc(me, sizeof(int));
      if (!S->Mskew || !S->Mfact || !S->ipivot) rc = PRIMME_MALLOC_FAILURE;
   }
   S->hVals = (double *)calloc(mb, sizeof(double));
   S->prevRitzVals = (double *)calloc(mb + primme->numEvals, sizeof(double));
   S->blockNorms = (double *)calloc(PB_MAX(primme->maxBlockSize, 1), sizeof(double));
   S->basisNorms = (double *)calloc(mb, sizeof(double));
   S->flags = (int *)calloc(mb, sizeof(int));
   S->map = (int *)calloc(mb, sizeof(int));
   S->iev = (int *)calloc(PB_MAX(primme->maxBlockSize, 1), sizeof(int));
   S->perm = (int *)calloc(PB_MAX(primme->numEvals, 1), sizeof(int));
   S->lockedFlags = (int *)calloc(PB_MAX(primme->numEvals, 1), sizeof(int));
   for (int i = 0; i < mb; i++) S->map[i] = i;
   if (rc || !S->H || !S->hVecs || !S->prevhVecs || !S->hVals) {
      pb_registry_set_solver(primme, NULL);
      if (!own_ctx) {
         /* V and W belong to the attached context (pb200_ctx_workspace) */
         if (own_evecs) pb200_free(S->dev, S->evecs);
         if (S->jd_work) pb200_free(S->dev, S->jd_work);
         if (S->Q) pb200_free(S->dev, S->Q);
         if (S->evecsHat) pb200_free(S->dev, S->evecsHat);
         S->V = S->W = S->evecs = S->jd_work = S->Q = S->evecsHat = NULL;
         S->dev = NULL;
      }
      free_solver(S, own_evecs);
      return PRIMME_MALLOC_FAILURE;
   }

   int ret = 0, numRet = 0;
   const int blas_threads = hl_blas_threads(1);
   const double tprof0 = hl_wtime();
   rc = pb_main_iter(S, evals, resNorms, &ret, &numRet);
   if (blas_threads > 0) hl_blas_threads(blas_threads);
   if (getenv("PB200_HOST_PROFILE")) {
      /* where the wall time of the solve went: waiting for panels (device + latency), the projected
       * eigen-solves, everything else on the host (launch calls, small dense algebra, bookkeeping) */
      extern double pb200_debug_wait_seconds(long *calls) __attribute__((weak));
      extern double hl_prof_eig_s;
      extern long hl_prof_eig_n;
      static double wait_prev = 0.0, eig_prev = 0.0;
      static long waitn_prev = 0, eign_prev = 0;
      long wn = 0;
      const double w = pb200_debug_wait_seconds ? pb200_debug_wait_seconds(&wn) : 0.0;
      const double el = hl_wtime() - tprof0;
      fprintf(stderr, "primme_b200 host profile: solve %.1f ms, %lld outer its; waiting for panels %.1f ms (%ld waits); "
                      "projected eigen-solves %.1f ms (%ld calls); other host work %.1f ms\n",
            1e3 * el, (long long)primme->stats.numOuterIterations, 1e3 * (w - wait_prev), wn - waitn_prev,
            1e3 * (hl_prof_eig_s - eig_prev), hl_prof_eig_n - eign_prev, 1e3 * (el - (w - wait_prev) - (hl_prof_eig_s - eig_prev)));
      wait_prev = w, waitn_prev = wn, eig_prev = hl_prof_eig_s, eign_prev = hl_prof_eig_n;
   }
   if (rc == 0) {
      rc = ret;
      if (!device_mode && numRet > 0)
         rc = pb200_copy_d2h(S->dev, S->evecs + (size_t)S->ldevecs * primme->numOrthoConst,
                    S->ldevecs, evecs + (size_t)primme->ldevecs * primme->numOrthoConst,
                    primme->ldevecs, S->n, numRet, PB_ES)
                    ? PRIMME_UNEXPECTED_FAILURE
                    : ret;
   } else {
      primme->initSize = 0;
   }
   pb200_ctx_sync(S->dev);
   pb200_ctx_l2_persist(S->dev, NULL, 0);
   pb_registry_set_solver(primme, NULL);
   primme->stats.elapsedTime = hl_wtime() - S->t0;
   if (getenv("PB200_DEBUG") && primme->procID == 0)
      fprintf(stderr,
            "PRIMME-B200: phases (s): elapsed %.4f matvec %.4f ortho %.4f vwxr %.4f projection %.4f solveH %.4f "
            "restart %.4f | outer %lld restarts %lld matvecs %lld\n",
            primme->stats.elapsedTime, primme->stats.timeMatvec, primme->stats.timeOrtho,
            primme->stats.timeDense, S->tProj, S->tSolveH, S->tRestart,
            (long long)primme->stats.numOuterIterations, (long long)primme->stats.numRestarts,
            (long long)primme->stats.numMatvecs);
   if (!own_ctx) {
      /* V and W belong to the attached context (pb200_ctx_workspace) */
      if (own_evecs) pb200_free(S->dev, S->evecs);
      if (S->jd_work) pb200_free(S->dev, S->jd_work);
      if (S->Q) pb200_free(S->dev, S->Q);
      if (S->evecsHat) pb200_free(S->dev, S->evecsHat);
      S->V = S->W = S->evecs = S->jd_work = S->Q = S->evecsHat = NULL;
      S->dev = NULL;
   }
   free_solver(S, own_evecs);
   return rc;
}

#ifdef PB_COMPLEX
/* complex Hermitian twins (reference include/primme_eigs.h:392,416): evecs and the callback blocks are
 * interleaved (re,im) fp64; evals and resNorms real */
int zprimme(double *evals, PRIMME_COMPLEX_DOUBLE *evecs, double *resNorms, primme_params *primme) {
   return solve_typed(evals, (SCALAR *)evecs, resNorms, primme, 0);
}
int cublas_zprimme(double *evals, PRIMME_COMPLEX_DOUBLE *evecs, double *resNorms, primme_params *primme) {
   return solve_typed(evals, (SCALAR *)evecs, resNorms, primme, 1);
}
#else
int dprimme(double *evals, double *evecs, double *resNorms, primme_params *primme) {
   return solve_typed(evals, evecs, resNorms, primme, 0);
}
int cublas_dprimme(double *evals, double *evecs, double *resNorms, primme_params *primme) {
   return solve_typed(evals, evecs, resNorms, primme, 1);
}

/* Entry points of precisions / back ends outside the scope: same behaviour as a reference build
 * without that type (primme_c.c:233-242). */
#define PB_UNAVAILABLE(name, EV, VEC, RN)                                       \
   int name(EV *evals, VEC *evecs, RN *resNorms, primme_params *primme) {       \
      (void)evals, (void)evecs, (void)resNorms;                                 \
      if (primme) primme->initSize = 0;                                         \
      return PRIMME_FUNCTION_UNAVAILABLE;                                       \
   }
#define PB_UNAVAILABLE3(name, EV, VEC, RN) \
   PB_UNAVAILABLE(name, EV, VEC, RN) PB_UNAVAILABLE(magma_##name, EV, VEC, RN) PB_UNAVAILABLE(cublas_##name, EV, VEC, RN)

PB_UNAVAILABLE3(hprimme, PRIMME_HALF, PRIMME_HALF, PRIMME_HALF)
PB_UNAVAILABLE3(kprimme, PRIMME_HALF, PRIMME_COMPLEX_HALF, PRIMME_HALF)
PB_UNAVAILABLE3(sprimme, float, float, float)
PB_UNAVAILABLE3(cprimme, float, PRIMME_COMPLEX_FLOAT, float)
PB_UNAVAILABLE(magma_dprimme, double, double, double)
PB_UNAVAILABLE(magma_zprimme, double, PRIMME_COMPLEX_DOUBLE, double)
PB_UNAVAILABLE3(hsprimme, float, PRIMME_HALF, float)
PB_UNAVAILABLE3(ksprimme, float, PRIMME_COMPLEX_HALF, float)
PB_UNAVAILABLE3(kprimme_normal, PRIMME_COMPLEX_HALF, PRIMME_COMPLEX_HALF, PRIMME_HALF)
PB_UNAVAILABLE3(cprimme_normal, PRIMME_COMPLEX_FLOAT, PRIMME_COMPLEX_FLOAT, float)
PB_UNAVAILABLE3(zprimme_normal, PRIMME_COMPLEX_DOUBLE, PRIMME_COMPLEX_DOUBLE, double)
PB_UNAVAILABLE3(kcprimme_normal, PRIMME_COMPLEX_FLOAT, PRIMME_COMPLEX_HALF, float)
#endif /* !PB_COMPLEX */

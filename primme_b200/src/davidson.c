/* davidson.c -- outer block (Generalized/Jacobi-)Davidson iteration: the caller of the hot path.
 *
 * Restates the control flow of reference src/eigs/main_iter.c (main_iter_Sprimme :176-1444,
 * prepare_candidates :1470-1709, copy_back_candidates :1745-1841, verify_norms :1864-1894),
 * src/eigs/init.c (init_basis :125-238, init_block_krylov :258-323) and the GD/Olsen branch of
 * src/eigs/correction.c (:134-381, computeRobustShift :524-601, mergeSort :637-693) for the
 * configuration this library covers: Hermitian standard problem (B = I); Rayleigh-Ritz, refined and
 * harmonic extraction (dav_refined.c); GD-family corrections and the inner QMR solver (dav_jdqmr.c).
 * Decisions (flags, block selection,
 * restart sizes, tolerances) follow the reference line by line so that iteration counts match;
 * every n-long operation is a call into the sm_100a kernel layer (include/primme_b200.h):
 *     candidates X,R,|R|   -> pb200_dvwxr            (Num_update_VWXR)
 *     block ortho          -> pb200_dortho_sweep     (Num_ortho_kernel), see dav_ortho.c
 *     W = A V              -> user matrixMatvec (pb200_dspmm for the built-in CSR operator)
 *     H(:,new) = V' W      -> pb200_dortho_sweep     (update_projection), see dav_project.c
 *     restart              -> pb200_dvwxr            (Num_aux_update_VWXR), see dav_restart.c
 */
#include "pb_host.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* sums over the processes are still to be done on the host (no multi-rank kernel context) */
#define primme_numProcs_gt1(S) ((S)->primme->numProcs > 1 && pb200_ctx_nranks((S)->dev) <= 1)
#define VtBV_BLK(S, off) ((S)->VtBV ? &(S)->VtBV[(size_t)(S)->maxRank * (off) + (off)] : NULL)

/* solve_H (solve_projection.c:95-154): Rayleigh-Ritz pairs of (H, V'V) or the refined vectors from R */
/* PB200_HOST_PROFILE: host-only time per section of the outer iteration (wall - panel waits - eigen-solves) */
#ifndef PB_COMPLEX
double pb_sect_host[8];
int pb_sect_on = -1;
#else
extern double pb_sect_host[8];
extern int pb_sect_on;
#endif
extern double pb200_debug_wait_seconds(long *calls) __attribute__((weak));
extern double hl_prof_eig_s;
static double sect_now(void) {
   return hl_wtime() - (pb200_debug_wait_seconds ? pb200_debug_wait_seconds(NULL) : 0.0) - hl_prof_eig_s;
}
#define SECT(i, call)                                         \
   do {                                                       \
      if (pb_sect_on > 0) {                                   \
         const double t_ = sect_now();                        \
         call;                                                \
         pb_sect_host[i] += sect_now() - t_;                  \
      } else {                                                \
         call;                                                \
      }                                                       \
   } while (0)

static int solve_projected(pb_solver *S, int basisSize, int nLocked, int numConverged) {
   if (S->QtV) return pb_solve_H_harm(S, basisSize, VtBV_BLK(S, nLocked), S->maxRank, numConverged);
   if (S->refined) return pb_solve_H_ref(S, basisSize, VtBV_BLK(S, nLocked), S->maxRank, numConverged);
   return pb_solve_H(S, S->H, S->maxBasis, basisSize, VtBV_BLK(S, nLocked), S->maxRank, S->hVecs, S->maxBasis, S->hVals,
         numConverged, 1);
}

/* ------------------------------------------------------------------------------------------
 * Initial basis (init.c:125-323)
 * ---------------------------------------------------------------------------------------- */
static int init_block_krylov(pb_solver *S, int dv1, int dv2, int numLocked) {
   primme_params *primme = S->primme;
   const int numNew = dv2 - dv1 + 1;
   if (numNew <= 0) return 0;
   const int bs = numNew <= primme->maxBlockSize ? 1 : primme->maxBlockSize;
   int nV = 0;

   CHK(pb_fill_random(S, S->V + (size_t)S->ld * dv1, S->ld, bs));
   CHK(pb_ortho_block(S, S->V, S->ld, dv1, dv1 + bs - 1, S->evecs, S->ldevecs, numLocked, NULL, 0, &nV));
   if (nV != dv1 + bs) return PRIMME_UNEXPECTED_FAILURE;

   int m = bs;
   for (int i = dv1 + bs, mm = PB_MIN(m, dv2 - i + 1); i <= dv2; i += mm, mm = PB_MIN(mm, dv2 - i + 1)) {
      /* next Krylov block: A*V(:,i-bs:...) lands in V(:,i:...), and is also W(:,i-bs:...) */
      CHK(pb_apply_matvec(S, S->V + (size_t)S->ld * (i - bs), S->ld, S->V + (size_t)S->ld * i, S->ld, mm));
      CHK(pb200_copy_d2d(S->dev, S->V + (size_t)S->ld * i, S->ld, S->W + (size_t)S->ld * (i - bs),
            S->ld, S->n, mm, PB_ES));
      CHK(pb_ortho_block(S, S->V, S->ld, i, i + mm - 1, S->evecs, S->ldevecs, numLocked, NULL, 0, &nV));
      if (nV < i + mm) {
         CHK(pb_fill_random(S, S->V + (size_t)S->ld * nV, S->ld, i + mm - nV));
         CHK(pb_ortho_block(S, S->V, S->ld, nV, i + mm - 1, S->evecs, S->ldevecs, numLocked, NULL, 0, &nV));
      } else {
         /* the reference re-enters Bortho_block with an empty range here: no-op */
      }
      if (nV != i + mm) return PRIMME_UNEXPECTED_FAILURE;
   }
   CHK(pb_apply_matvec(S, S->V + (size_t)S->ld * (dv2 - bs + 1), S->ld,
         S->W + (size_t)S->ld * (dv2 - bs + 1), S->ld, bs));
   return 0;
}

static int init_basis(pb_solver *S, int *basisSize, int *nextGuess, int *numGuesses) {
   primme_params *primme = S->primme;
   int initSize, random = 0;

   if (primme->numOrthoConst > 0) {
      int nV = 0;
      CHK(pb_ortho_block(S, S->evecs, S->ldevecs, 0, primme->numOrthoConst - 1, NULL, 0, 0, NULL, 0, &nV));
      if (nV != primme->numOrthoConst) return PRIMME_ORTHO_CONST_FAILURE;
      /* the constraints join the skew projector (I - K^{-1}Q (Q'K^{-1}Q)^{-1} Q') (init.c:150-169) */
      if (S->evecsHat) {
         primme->ShiftsForPreconditioner = NULL;
         CHK(pb_apply_precond(S, S->evecs, S->ldevecs, S->evecsHat, S->ld, primme->numOrthoConst));
         CHK(pb_update_XKinvBX(S, 0, primme->numOrthoConst));
      }
   }
   initSize = PB_MIN(primme->locking ? primme->minRestartSize : primme->maxBasisSize, primme->initSize);
   {
      PRIMME_INT room = primme->n - primme->numOrthoConst;
      if (room < initSize) initSize = (int)room;
      if (initSize < 0) initSize = 0;
   }
   *numGuesses = primme->initSize - initSize;
   *nextGuess = primme->numOrthoConst + initSize;
   CHK(pb200_copy_d2d(S->dev, S->evecs + (size_t)S->ldevecs * primme->numOrthoConst, S->ldevecs,
         S->V, S->ld, S->n, initSize, PB_ES));

   switch (primme->initBasisMode) {
   case primme_init_krylov: random = 0; break;
   case primme_init_random: random = PB_MAX(0, primme->minRestartSize - initSize); break;
   case primme_init_user: random = PB_MAX(primme->maxBlockSize - initSize, 0); break;
   default: return PRIMME_UNEXPECTED_FAILURE;
   }
   {
      PRIMME_INT room = primme->n - primme->numOrthoConst - initSize;
      if (room < random) random = (int)room;
      if (random < 0) random = 0;
   }
   if (random > 0) CHK(pb_fill_random(S, S->V + (size_t)S->ld * initSize, S->ld, random));
   *basisSize = initSize + random;

   CHK(pb_ortho_block(S, S->V, S->ld, 0, *basisSize - 1, S->evecs, S->ldevecs,
         primme->numOrthoConst, NULL, 0, basisSize));
   CHK(pb_apply_matvec(S, S->V, S->ld, S->W, S->ld, *basisSize));

   if (primme->initBasisMode == primme_init_krylov) {
      int minRestartSize = primme->minRestartSize;
      if (primme->n - primme->numOrthoConst < minRestartSize)
         minRestartSize = (int)(primme->n - primme->numOrthoConst);
      CHK(init_block_krylov(S, *basisSize, minRestartSize - 1, primme->numOrthoConst));
      *basisSize = minRestartSize;
   }
   return 0;
}

/* ------------------------------------------------------------------------------------------
 * One candidates sweep: X = V*h, R = W*h - X*diag(theta), norms (main_iter.c:1656-1688)
 * ---------------------------------------------------------------------------------------- */
static int candidates_sweep(pb_solver *S, int basisSize, const SCALAR *hblk, const double *theta,
      int nb, SCALAR *X, SCALAR *R, int computeXR, double *norms) {
   primme_params *primme = S->primme;
   const double t0 = hl_wtime();
   pb200_vwxr_out o;
   memset(&o, 0, sizeof(o));
   int fused = 0;
   if (computeXR) {
      o.X[0].ptr = PB_DP(X), o.X[0].ld = S->ld, o.X[0].cb = 0, o.X[0].ce = nb;
      o.R.ptr = PB_DP(R), o.R.ld = S->ld, o.R.cb = 0, o.R.ce = nb;
      o.Rnorms_host = norms;
      /* the residual block becomes the new basis block unchanged (no preconditioner, no locked
       * vectors): the same sweep delivers the first Gram panel of its orthogonalisation */
      S->fusedP_nb = 0;
      S->fuse_sweeps++;
      if (S->fuse_enabled && X == S->V + (size_t)S->ld * basisSize &&
            pb200_dvwxr_can_fuse_gram(S->dev, S->n, S->V, S->W, basisSize, S->ld, nb, &o)) {
         o.P_host = PB_DP(S->fusedP), o.ldP = S->maxBasis + 8;
         /* the residuals go straight into the basis tail as well (the Ritz vectors are not needed
          * by this correction): no copy kernel in solve_correction */
         o.X[0].ptr = NULL;
         o.R2 = PB_DP(X), o.ldR2 = S->ld;
         fused = 1;
      }
   } else {
      o.rb = 0, o.re = nb, o.rnorms_host = norms;
   }
   CHK(pb200_dvwxr(S->dev, S->n, S->V, S->W, basisSize, S->ld, hblk, S->maxBasis, nb, theta, &o));
   if (fused) {
      CHK(pb_reduce_panel(S, S->fusedP, basisSize + nb, nb, S->maxBasis + 8));
      S->fusedP_m = basisSize, S->fusedP_nb = nb;
   }
   if (primme->numProcs > 1 && pb200_ctx_nranks(S->dev) <= 1) {
      for (int i = 0; i < nb; i++) norms[i] *= norms[i];
      CHK(pb_global_sum(S, norms, nb));
      for (int i = 0; i < nb; i++) norms[i] = sqrt(norms[i]);
   }
   primme->stats.timeDense += hl_wtime() - t0;
   primme->stats.flopsDense += 2.0 * (double)S->n * basisSize * nb + 2.0 * (double)S->n * nb;
   return 0;
}

/* ------------------------------------------------------------------------------------------
 * prepare_candidates (main_iter.c:1470-1709): fill the block with the first unconverged Ritz
 * pairs, computing residuals as needed.
 * ---------------------------------------------------------------------------------------- */
static int prepare_candidates(pb_solver *S, int basisSize, SCALAR *X, SCALAR *R, int computeXR,
      int remainedEvals, int blockNormsSize, int maxBlockSize, int numLocked, double *evals,
      double *resNorms, int targetShiftIndex, int *blockSize, int *recentlyConverged,
      double *smallestResNorm, int numConverged, int *reset, int nprevhVecs,
      int practConvChecking) {
   primme_params *primme = S->primme;
   int *flags = S->flags, *iev = S->iev, *map = S->map;
   double *hVals = S->hVals, *blockNorms = S->blockNorms, *basisNorms = S->basisNorms;
   const int ldh = S->maxBasis;
   int i, blki, lasti = -1, rc = 0;
   (void)targetShiftIndex;

   *blockSize = 0;
   double *hValsBlock = (double *)malloc(sizeof(double) * (maxBlockSize > 0 ? maxBlockSize : 1));
   SCALAR *hVecsBlock = (SCALAR *)malloc(sizeof(SCALAR) * (size_t)ldh * (maxBlockSize > 0 ? maxBlockSize : 1));
   int *flagsBlock = (int *)malloc(sizeof(int) * (maxBlockSize > 0 ? maxBlockSize : 1));

   for (i = 0; i < blockNormsSize; i++) hValsBlock[i] = hVals[iev[*blockSize + i]];
   if (blockNormsSize > 0) {
      for (*smallestResNorm = HUGE_VAL, i = 0; i < blockNormsSize; i++)
         *smallestResNorm = PB_MIN(*smallestResNorm, blockNorms[i]);
   }

   /* carry the flags of the previous iteration to the pairs closest in angle (:1524-1533) */
   pb_map_vecs(S->prevhVecs, basisSize, nprevhVecs, ldh, S->hVecs, 0, basisSize, ldh, map);
   hl_permute_ints(flags, basisSize, map);

   *recentlyConverged = 0;
   for (;;) {
      for (i = *blockSize; i < *blockSize + blockNormsSize; i++) flagsBlock[i - *blockSize] = flags[iev[i]];
      rc = pb_check_convergence(S, X ? X + (size_t)S->ld * *blockSize : NULL, S->ld, computeXR,
            R ? R + (size_t)S->ld * *blockSize : NULL, S->ld, computeXR, numLocked, 0,
            blockNormsSize, flagsBlock, &blockNorms[*blockSize], hValsBlock, reset,
            practConvChecking);
      if (rc) goto done;

      for (blki = *blockSize, i = 0; i < blockNormsSize && *blockSize < maxBlockSize; i++, blki++) {
         flags[iev[blki]] = flagsBlock[i];
         basisNorms[iev[blki]] = blockNorms[blki];
         double shift = primme->targetShifts ? primme->targetShifts[targetShiftIndex] : 0.0;
         if ((primme->target == primme_closest_leq && hVals[iev[blki]] - blockNorms[blki] > shift) ||
               (primme->target == primme_closest_geq && hVals[iev[blki]] + blockNorms[blki] < shift)) {
            /* wrong side of the shift: ignore */
         } else if (flagsBlock[i] != UNCONVERGED && *recentlyConverged < remainedEvals &&
                    (iev[blki] < primme->numEvals - numLocked ||
                          primme->target == primme_closest_geq ||
                          primme->target == primme_closest_leq)) {
            if (!primme->locking) {
               evals[iev[blki]] = hVals[iev[blki]];
               resNorms[iev[blki]] = blockNorms[blki];
               if (flagsBlock[i] == CONVERGED)
                  primme->stats.maxConvTol = PB_MAX(primme->stats.maxConvTol, blockNorms[blki]);
            }
            (*recentlyConverged)++;
            if (*blockSize == 0) *smallestResNorm = HUGE_VAL;
            maxBlockSize = PB_MIN(maxBlockSize,
                  primme->numEvals + 1 - *recentlyConverged - numConverged);
            rc = pb_monitor(S, hVals, basisSize, flags, &iev[blki], 1, basisNorms,
                  numConverged + *recentlyConverged, NULL, 0, NULL, NULL, -1, -1.0, NULL, 0.0,
                  primme_event_converged);
            if (rc) goto done;
         } else if (flagsBlock[i] == UNCONVERGED) {
            if (*blockSize == 0) *smallestResNorm = HUGE_VAL;
            *smallestResNorm = PB_MIN(*smallestResNorm, blockNorms[blki]);
            blockNorms[*blockSize] = blockNorms[blki];
            iev[*blockSize] = iev[blki];
            if (computeXR && blki != *blockSize) {
               rc = pb200_copy_d2d(S->dev, X + (size_t)S->ld * blki, S->ld,
                     X + (size_t)S->ld * *blockSize, S->ld, S->n, 1, PB_ES);
               if (!rc)
                  rc = pb200_copy_d2d(S->dev, R + (size_t)S->ld * blki, S->ld,
                        R + (size_t)S->ld * *blockSize, S->ld, S->n, 1, PB_ES);
               if (rc) goto done;
            }
            (*blockSize)++;
         }
         lasti = iev[blki];
      }

      /* well conditioned coefficient vectors for the next candidates (prepare_vecs: a no-op for
       * Rayleigh-Ritz), then the candidates after the last pair visited (:1629-1643) */
      blki = *blockSize;
      rc = pb_prepare_vecs(S, basisSize, lasti + 1, maxBlockSize - blki, targetShiftIndex, &S->numArbitraryVecs,
            *smallestResNorm, flags, 1);
      if (rc) goto done;
      for (i = lasti + 1; i < basisSize && blki < maxBlockSize; i++)
         if (flags[i] == UNCONVERGED) iev[blki++] = i;
      if (blki == *blockSize || *recentlyConverged >= remainedEvals) break;
      blockNormsSize = blki - *blockSize;

      for (i = 0; i < blockNormsSize; i++) {
         hValsBlock[i] = hVals[iev[*blockSize + i]];
         memcpy(&hVecsBlock[(size_t)ldh * i], &S->hVecs[(size_t)ldh * iev[*blockSize + i]],
               sizeof(SCALAR) * basisSize);
      }
      rc = candidates_sweep(S, basisSize, hVecsBlock, hValsBlock, blockNormsSize,
            X ? X + (size_t)S->ld * *blockSize : NULL, R ? R + (size_t)S->ld * *blockSize : NULL,
            computeXR, &blockNorms[*blockSize]);
      if (rc) goto done;

      /* the reference's clamp loop (:1686-1688) runs over i in [*blockSize, blockNormsSize) */
      for (i = *blockSize; i < blockNormsSize; i++)
         blockNorms[i] = PB_MAX(blockNorms[i], primme->stats.estimateResidualError);
   }
done:
   free(hValsBlock), free(hVecsBlock), free(flagsBlock);
   return rc;
}

/* ------------------------------------------------------------------------------------------
 * GD / Olsen correction (correction.c:134-381 with maxInnerIterations == 0)
 * ---------------------------------------------------------------------------------------- */
static void merge_sort(const double *lockedEvals, int numLocked, const double *ritzVals,
      const int *flags, int basisSize, double *sorted, int *ilev, int blockSize,
      primme_params *primme) {
   int count = 0, eval = 0, ritzVal = 0, blockIndex = 0;
   while (count < numLocked + basisSize) {
      if (eval >= numLocked ||
            (ritzVal < basisSize &&
                  ((primme->target == primme_largest && ritzVals[ritzVal] >= lockedEvals[eval]) ||
                        (primme->target == primme_smallest && ritzVals[ritzVal] <= lockedEvals[eval])))) {
         sorted[count] = ritzVals[ritzVal];
         if (blockIndex < blockSize && flags[ritzVal] == UNCONVERGED) ilev[blockIndex++] = count;
         ritzVal++;
      } else if (ritzVal >= basisSize ||
                 (primme->target == primme_largest && lockedEvals[eval] >= ritzVals[ritzVal]) ||
                 (primme->target == primme_smallest && lockedEvals[eval] <= ritzVals[ritzVal])) {
         sorted[count] = lockedEvals[eval];
         eval++;
      }
      count++;
   }
}

static double robust_shift(int blockIndex, double resNorm, const double *prevRitzVals,
      int numPrevRitzVals, const double *sorted, double *approxOlsenShift, int numSorted,
      const int *ilev, primme_params *primme) {
   const double invB = primme->stats.estimateInvBNorm;
   if (primme->stats.numOuterIterations <= 1) {
      *approxOlsenShift = resNorm * sqrt(invB);
      return resNorm * sqrt(invB);
   }
   int si = ilev[blockIndex];
   double gap, lowerGap, upperGap, delta, epsilon;
   if (si == 0 && numSorted >= 2) {
      lowerGap = DBL_MAX;
      gap = fabs(sorted[1] - sorted[0]);
   } else if (si > 0 && numSorted >= 2 && si + 1 < numSorted) {
      lowerGap = fabs(sorted[si] - sorted[si - 1]);
      upperGap = fabs(sorted[si + 1] - sorted[si]);
      gap = PB_MIN(lowerGap, upperGap);
   } else {
      lowerGap = fabs(sorted[si] - sorted[si - 1]);
      gap = lowerGap;
   }
   delta = si < numPrevRitzVals ? fabs(prevRitzVals[si] - sorted[si]) : DBL_MAX;
   if (gap > resNorm)
      epsilon = PB_MIN(delta, PB_MIN(resNorm * resNorm * invB / gap, lowerGap));
   else
      epsilon = PB_MIN(resNorm * sqrt(invB), lowerGap);
   *approxOlsenShift = PB_MIN(delta, epsilon);
   return epsilon;
}

static int solve_correction(pb_solver *S, double *evals, int numLocked, int basisSize,
      int blockSize) {
   primme_params *primme = S->primme;
   const correction_params *cp = &primme->correctionParams;
   double *ritzVals = S->hVals, *prevRitzVals = S->prevRitzVals, *blockNorms = S->blockNorms;
   int *iev = S->iev;
   const double sqrtInvB = sqrt(primme->stats.estimateInvBNorm);
   const int extremal = primme->target == primme_smallest || primme->target == primme_largest;
   double *shifts = (double *)malloc(sizeof(double) * blockSize);
   double *olsenEps = (double *)malloc(sizeof(double) * blockSize);
   double *sorted = ritzVals;
   int *ilev = iev, owns = 0, rc = 0;

   if (primme->locking && extremal) {
      sorted = (double *)malloc(sizeof(double) * (numLocked + basisSize));
      ilev = (int *)malloc(sizeof(int) * blockSize);
      owns = 1;
      merge_sort(evals, numLocked, ritzVals, S->flags, basisSize, sorted, ilev, blockSize, primme);
   }

   if (!extremal) {
      for (int b = 0; b < blockSize; b++) {
         double t = primme->numTargetShifts > 0
                          ? primme->targetShifts[PB_MIN(primme->numTargetShifts - 1, numLocked)]
                          : 0.0;
         int si = ilev[b];
         if (fabs(sorted[si] - t) < blockNorms[b] * sqrtInvB)
            shifts[b] = t;
         else if (S->refined)
            shifts[b] = sorted[si]; /* |Ritz value - target| <= singular value: trusted earlier (:226-228) */
         else
            shifts[b] = sorted[si] + (blockNorms[b] * sqrtInvB) * (t - sorted[si]) / fabs(t - sorted[si]);
         olsenEps[b] = si < S->numPrevRitzVals ? fabs(prevRitzVals[si] - sorted[si])
                                               : blockNorms[b] * sqrtInvB;
      }
      S->numPrevRitzVals = basisSize;
      memcpy(prevRitzVals, sorted, sizeof(double) * basisSize);
   } else {
      if (cp->robustShifts) {
         for (int b = 0; b < blockSize; b++) {
            int si = ilev[b];
            double rs = robust_shift(b, blockNorms[b], prevRitzVals, S->numPrevRitzVals, sorted,
                  &olsenEps[b], numLocked + basisSize, ilev, primme);
            if (primme->target == primme_smallest) {
               shifts[b] = sorted[si] - rs;
               if (si > 0) shifts[b] = PB_MAX(shifts[b], sorted[si - 1]);
            } else {
               shifts[b] = sorted[si] + rs;
               if (si > 0) shifts[b] = PB_MIN(shifts[b], sorted[si - 1]);
            }
         }
      } else {
         for (int b = 0; b < blockSize; b++) {
            int si = ilev[b];
            shifts[b] = ritzVals[iev[b]];
            olsenEps[b] = si < S->numPrevRitzVals ? fabs(prevRitzVals[si] - sorted[si])
                                                  : blockNorms[b] * sqrtInvB;
         }
      }
      S->numPrevRitzVals = numLocked + basisSize;
      memcpy(prevRitzVals, sorted, sizeof(double) * S->numPrevRitzVals);
   }

   primme->ShiftsForPreconditioner = shifts;

   SCALAR *r = S->W + (size_t)S->ld * basisSize; /* block residuals */
   SCALAR *x = S->V + (size_t)S->ld * basisSize; /* block Ritz vectors, receives the corrections */
   if (cp->maxInnerIterations != 0) {
      /* inner-outer JDQMR (correction.c:385-447, setup_JD_projectors :862-999 for the presets
       * without right projectors): Q = [constraints locked] when LeftQ; the Ritz vector joins Q
       * when the block is a single vector, else every system is projected against its own x_i */
      const int sizeEvecs = primme->numOrthoConst + (primme->locking ? numLocked : S->numConvergedStored);
      const SCALAR *Q = NULL;
      int nQ = 0, useX = 0;
      if (cp->projectors.LeftQ) {
         nQ = sizeEvecs, Q = S->evecs;
         if (cp->projectors.LeftX) {
            if (blockSize <= 1) {
               rc = pb200_copy_d2d(S->dev, x, S->ld, S->evecs + (size_t)S->ldevecs * sizeEvecs, S->ldevecs, S->n,
                     blockSize, PB_ES);
               nQ += blockSize;
            } else
               useX = 1;
         }
      } else if (cp->projectors.LeftX)
         useX = 1;
      if (!rc && !S->jd_work)
         rc = pb200_malloc(S->dev, sizeof(SCALAR) * (size_t)S->ld * 5 * PB_MAX(primme->maxBlockSize, 1),
               (void **)&S->jd_work);
      /* right projectors (:942-980): the locked vectors, and x or K^{-1}x with x'K^{-1}x */
      const SCALAR *RQ = NULL;
      SCALAR *RX = NULL, *KinvX = NULL;
      SCALAR *xKinvBx = (SCALAR *)malloc(sizeof(SCALAR) * PB_MAX(blockSize, 1));
      int nRQ = 0;
      /* right projector on Q: orthogonal (RQ = evecs) or, with a preconditioner and SkewQ, the skew one
       * (RQ = evecsHat = K^{-1} evecs, overlaps with evecs solved through the factors of M; :948-954) */
      int64_t ldRQ = S->ldevecs;
      const SCALAR *skewQ = NULL;
      if (cp->projectors.RightQ) {
         RQ = S->evecs, nRQ = sizeEvecs;
         if (cp->precondition && cp->projectors.SkewQ && S->evecsHat) RQ = S->evecsHat, ldRQ = S->ld, skewQ = S->evecs;
      }
      if (!rc && cp->projectors.RightX) {
         if (cp->precondition && cp->projectors.SkewX) {
            rc = pb200_malloc(S->dev, sizeof(SCALAR) * (size_t)S->ld * blockSize, (void **)&KinvX);
            if (!rc) rc = pb_apply_precond(S, x, S->ld, KinvX, S->ld, blockSize);
            if (!rc) rc = pb200_dcolumn_dots(S->dev, S->n, x, S->ld, KinvX, S->ld, blockSize, xKinvBx);
            if (!rc) rc = pb_reduce_panel(S, xKinvBx, blockSize, 1, blockSize);
            RX = KinvX;
         } else {
            RX = x;
            for (int b = 0; b < blockSize; b++) xKinvBx[b] = 1.0;
         }
      }
      if (!rc) {
         /* The systems of a block are independent (every scalar of the recurrences is per system, inner_solve.c:
          * 155-182): blocks wider than the 8 systems the inner solver carries are solved in chunks of 8, every
          * chunk starting from the same stopping-criterion state `touch`, whose increments add up as in one call
          * (once per call for the decreasing tolerance, once per system stopped by the convergence test else). */
         double *blockRitzVals = (double *)malloc(sizeof(double) * PB_MAX(blockSize, 1));
         for (int b = 0; b < blockSize; b++) blockRitzVals[b] = ritzVals[iev[b]];
         SCALAR *sol = S->jd_work + (size_t)S->ld * 4 * PB_MAX(primme->maxBlockSize, 1);
         const int touch0 = S->touch;
         int inc = 0;
         for (int c0 = 0; c0 < blockSize && !rc; c0 += 8) {
            const int nc = PB_MIN(8, blockSize - c0);
            int touch1 = touch0;
            rc = pb_inner_solve(S, nc, x + (size_t)S->ld * c0, S->ld, r + (size_t)S->ld * c0, S->ld, blockNorms + c0, Q,
                  S->ldevecs, nQ, useX, sol + (size_t)S->ld * c0, S->ld, blockRitzVals + c0, shifts + c0, &touch1,
                  S->jd_work, RQ, ldRQ, nRQ, RX ? RX + (size_t)S->ld * c0 : NULL, S->ld, xKinvBx + c0, skewQ, S->ldevecs,
                  S->Mfact, S->ipivot);
            inc = cp->convTest == primme_decreasing_LTolerance ? PB_MAX(inc, touch1 - touch0) : inc + (touch1 - touch0);
         }
         S->touch = PB_MAX(S->touch, touch0 + inc);
         if (!rc) rc = pb200_copy_d2d(S->dev, sol, S->ld, x, S->ld, S->n, blockSize, PB_ES);
         free(blockRitzVals);
      }
      if (KinvX) pb200_free(S->dev, KinvX);
      free(xKinvBx);
   } else if (cp->projectors.RightX && cp->projectors.SkewX) {
      /* exact Olsen projector (correction.c:695-774): x <- K^{-1}r - (x'K^{-1}r / x'K^{-1}x) K^{-1}x */
      SCALAR *tmp = NULL;
      rc = pb200_malloc(S->dev, sizeof(SCALAR) * (size_t)S->ld * blockSize * 2, (void **)&tmp);
      if (!rc) {
         SCALAR *Kx = tmp, *Kr = tmp + (size_t)S->ld * blockSize;
         SCALAR *xKx = (SCALAR *)malloc(sizeof(SCALAR) * 2 * blockSize), *xKr = xKx + blockSize;
         rc = pb_apply_precond(S, x, S->ld, Kx, S->ld, blockSize);
         if (!rc) rc = pb_apply_precond(S, r, S->ld, Kr, S->ld, blockSize);
         if (!rc) rc = pb200_dcolumn_dots(S->dev, S->n, x, S->ld, Kx, S->ld, blockSize, xKx);
         if (!rc) rc = pb_reduce_panel(S, xKx, blockSize, 1, blockSize);
         if (!rc) rc = pb200_dcolumn_dots(S->dev, S->n, x, S->ld, Kr, S->ld, blockSize, xKr);
         if (!rc) rc = pb_reduce_panel(S, xKr, blockSize, 1, blockSize);
         if (!rc) rc = pb200_copy_d2d(S->dev, Kr, S->ld, x, S->ld, S->n, blockSize, PB_ES);
         if (!rc) {
            for (int b = 0; b < blockSize; b++) xKr[b] = PB_ABS(xKx[b]) > 0.0 ? -xKr[b] / xKx[b] : 0.0;
            rc = pb200_daxpy_columns(S->dev, S->n, xKr, Kx, S->ld, x, S->ld, blockSize);
         }
         free(xKx);
         pb200_free(S->dev, tmp);
      }
   } else {
      /* approximate Olsen: r <- r - eps*x before preconditioning (:356-372) */
      if (cp->projectors.RightX &&
            ((cp->precondition && primme->applyPreconditioner) ||
                  (primme->locking && primme->orth == primme_orth_implicit_I))) {
         SCALAR *negEps = (SCALAR *)malloc(sizeof(SCALAR) * blockSize);
         for (int b = 0; b < blockSize; b++) negEps[b] = olsenEps[b] = -olsenEps[b];
         rc = pb200_daxpy_columns(S->dev, S->n, negEps, x, S->ld, r, S->ld, blockSize);
         free(negEps);
      }
      if (!rc && !(S->fusedP_nb == blockSize && blockSize > 0 && !cp->precondition))
         rc = pb_apply_precond(S, r, S->ld, x, S->ld, blockSize);
      /* else: the candidates sweep already wrote the residual block into x (fused path) */
   }
   primme->ShiftsForPreconditioner = NULL;
   if (owns) free(sorted), free(ilev);
   free(shifts), free(olsenEps);
   return rc;
}

/* ------------------------------------------------------------------------------------------
 * verify_norms (main_iter.c:1864-1894) -- destroys W(:,0:nv) (it holds residuals afterwards)
 * ---------------------------------------------------------------------------------------- */
static int verify_norms(pb_solver *S, int nv, double *resNorms, int *numConverged) {
   CHK(pb200_dresidual_inplace(S->dev, S->n, S->hVals, S->V, S->ld, S->W, S->ld, nv, resNorms));
   if (primme_numProcs_gt1(S)) CHK(pb_global_sum(S, resNorms, nv));
   for (int i = 0; i < nv; i++) resNorms[i] = sqrt(resNorms[i]);
   CHK(pb_check_convergence(S, S->V, S->ld, 1, S->W, S->ld, 1, 0, 0, nv, S->flags, resNorms,
         S->hVals, NULL, 0));
   int i;
   for (i = 0; i < nv && S->flags[i] != UNCONVERGED; i++)
      ;
   *numConverged = i;
   return 0;
}

/* copy_back_candidates (main_iter.c:1745-1841): with locking, hand back unconverged Ritz pairs
 * when the iteration budget ran out */
static int copy_back_candidates(pb_solver *S, int basisSize, double *evals, double *resNorms,
      int targetShiftIndex, int numConverged, int *numRet) {
   primme_params *primme = S->primme;
   *numRet = numConverged; /* the locked vectors already sit in evecs: the host contract copies them back */
   if (numConverged >= primme->numEvals || basisSize <= 0) return 0;
   SCALAR *ev = S->evecs + (size_t)S->ldevecs * primme->numOrthoConst;
   int i = 0;
   while (i < basisSize && numConverged < primme->numEvals) {
      int bs = PB_MAX(0, PB_MIN(primme->numEvals - numConverged, basisSize - i));
      bs = PB_MIN(bs, PB_ES);
      pb200_vwxr_out o;
      memset(&o, 0, sizeof(o));
      o.X[0].ptr = PB_DP(ev + (size_t)S->ldevecs * numConverged), o.X[0].ld = S->ldevecs, o.X[0].cb = 0, o.X[0].ce = bs;
      o.rb = 0, o.re = bs, o.rnorms_host = &resNorms[numConverged];
      CHK(pb200_dvwxr(S->dev, S->n, S->V, S->W, basisSize, S->ld, &S->hVecs[(size_t)S->maxBasis * i],
            S->maxBasis, bs, &S->hVals[i], &o));
      if (primme->numProcs > 1 && pb200_ctx_nranks(S->dev) <= 1) {
         double *rn = &resNorms[numConverged];
         for (int t = 0; t < bs; t++) rn[t] *= rn[t];
         CHK(pb_global_sum(S, rn, bs));
         for (int t = 0; t < bs; t++) rn[t] = sqrt(rn[t]);
      }
      int numConverged0 = numConverged;
      for (int blki = 0; blki < bs; blki++, i++) {
         double shift = primme->targetShifts ? primme->targetShifts[targetShiftIndex] : 0.0;
         double rn = resNorms[numConverged0 + blki];
         if ((primme->target == primme_closest_leq && S->hVals[i] - rn > shift) ||
               (primme->target == primme_closest_geq && S->hVals[i] + rn < shift))
            continue;
         evals[numConverged] = S->hVals[i];
         resNorms[numConverged] = rn;
         if (numConverged != numConverged0 + blki)
            CHK(pb200_copy_d2d(S->dev, ev + (size_t)S->ldevecs * (numConverged0 + blki), S->ldevecs,
                  ev + (size_t)S->ldevecs * numConverged, S->ldevecs, S->n, 1, PB_ES));
         numConverged++;
      }
   }
   for (i = numConverged; i < primme->numEvals; i++) resNorms[i] = -1;
   *numRet = numConverged;
   return 0;
}

/* ==========================================================================================
 * The outer iteration (main_iter.c:176-1444)
 * ======================================================================================== */
int pb_main_iter(pb_solver *S, double *evals, double *resNorms, int *ret, int *numRet) {
   primme_params *primme = S->primme;
   int i, rc = 0;
   int blockSize = 0, availableBlockSize = 0, basisSize = 0, numLocked = 0, numGuesses = 0,
       nextGuess = 0, numConverged = 0, targetShiftIndex = 0, recentlyConverged = 0,
       maxRecentlyConverged = 0, LockingProblem = 0, restartLimitReached, nprevhVecs = 0,
       reset = 0, restartsSinceReset = 0, wholeSpace = 0;
   const int maxNumRandoms = 10;
   const int maxBasis = primme->maxBasisSize, ldh = maxBasis;
   double smallestResNorm;
   int *flags = S->flags, *iev = S->iev, *map = S->map;
   /* evaluated at every use: PRIMME_DYNAMIC changes maxInnerIterations while iterating */
#define gdNoPrecLocking \
   (primme->locking && !primme->correctionParams.precondition && primme->correctionParams.maxInnerIterations == 0)

   *ret = PRIMME_MAIN_ITER_FAILURE;
   *numRet = 0;
   if (pb_sect_on < 0) pb_sect_on = getenv("PB200_HOST_PROFILE") != NULL;
   if (pb_sect_on > 0) memset(pb_sect_host, 0, sizeof(pb_sect_host));

   /* counters (main_iter.c:371-396) */
   memset(&primme->stats, 0, sizeof(primme->stats));
   primme->stats.estimateMinEVal = HUGE_VAL;
   primme->stats.estimateMaxEVal = -HUGE_VAL;
   primme->stats.estimateLargestSVal = -HUGE_VAL;
   primme->stats.estimateBNorm = 1.0;
   primme->stats.estimateInvBNorm = 1.0;
   for (i = 0; i < primme->numEvals; i++) S->perm[i] = i;
   for (i = 0; i < maxBasis; i++) S->basisNorms[i] = 0.0;

   if (primme->numEvals == 0) {
      primme->initSize = 0;
      *ret = 0;
      goto clean;
   }

   SECT(7, rc = init_basis(S, &basisSize, &nextGuess, &numGuesses));
   if (rc) return rc;
   primme->initSize = 0;

   /* dynamic method switching: inner iterations are allowed or not from run-time timings
    * (main_iter.c:423-437) */
   if (primme->dynamicMethodSwitch > 0) {
      pb_dyn_init(&S->cost, primme);
      S->cost.MV = primme->stats.timeMatvec / (double)PB_MAX(primme->stats.numMatvecs, 1);
      if (primme->numEvals < 5 || primme->maxBasisSize + (primme->locking ? primme->numEvals : 0) >= primme->n)
         primme->dynamicMethodSwitch = 1; /* tentatively GD+k */
      else
         primme->dynamicMethodSwitch = 3; /* GD+k for the first pair */
      primme->correctionParams.maxInnerIterations = 0;
   }

   while (primme->stats.numMatvecs < primme->maxMatvecs &&
          (primme->maxOuterIterations == 0 ||
                primme->stats.numOuterIterations < primme->maxOuterIterations)) {

      primme->initSize = numConverged = S->numConvergedStored = numLocked;
      reset = 0;
      for (i = 0; i < maxBasis; i++) flags[i] = UNCONVERGED;
      targetShiftIndex = 0;
      if (S->numQR) {
         int nQ = 0;
         CHK(pb_update_Q(S, primme->targetShifts[targetShiftIndex], 0, basisSize, &nQ));
         if (basisSize != nQ) return PRIMME_UNEXPECTED_FAILURE; /* "Not supported deficient QR" (:466) */
      }

      CHK(pb_update_projection(S, 0, basisSize));
      CHK(pb_update_QtV(S, 0, basisSize));
      CHK(solve_projected(S, basisSize, primme->numOrthoConst + numLocked, numConverged));
      S->numArbitraryVecs = 0;

      maxRecentlyConverged = availableBlockSize = blockSize = 0;
      smallestResNorm = HUGE_VAL;
      primme->stats.estimateResidualError = 0.0;
      if (!primme->locking) primme->stats.maxConvTol = 0.0;
      restartsSinceReset = 0;

      /* ---- restart loop ---- */
      while (numConverged < primme->numEvals && primme->stats.numMatvecs < primme->maxMatvecs &&
             (primme->maxOuterIterations == 0 ||
                   primme->stats.numOuterIterations < primme->maxOuterIterations) &&
             !wholeSpace) {

         nprevhVecs = 0;
         int candidates_prepared = 0;

         /* ---- expansion loop: grow the basis block by block ---- */
         while (basisSize < maxBasis && primme->stats.numMatvecs < primme->maxMatvecs &&
                (primme->maxOuterIterations == 0 ||
                      primme->stats.numOuterIterations < primme->maxOuterIterations)) {

            primme->stats.numOuterIterations++;

            if (primme->numTargetShifts > numConverged + 1 && S->numQR) {
               /* one pair at a time while the QR factorisation depends on the shift (:525-528) */
               availableBlockSize = 1;
               maxRecentlyConverged = numConverged - numLocked + 1;
            } else {
               availableBlockSize = primme->maxBlockSize;
               maxRecentlyConverged = PB_MAX(0, primme->numEvals - numConverged);
            }
            availableBlockSize = PB_MIN(availableBlockSize, maxBasis - basisSize);
            availableBlockSize = PB_MIN(availableBlockSize, maxRecentlyConverged + 1);

            if (availableBlockSize > 0) {
               int practConvCheck = 0;
               if (primme->n <= basisSize + numLocked + primme->numOrthoConst)
                  practConvCheck = 1;
               else if (gdNoPrecLocking)
                  practConvCheck = -1;
               /* fused first Gram panel: only when the correction below is a plain copy of the
                * residuals and the block is orthogonalised against the basis alone */
               S->fuse_enabled = S->fuse_allowed && S->VtBV != NULL && !primme->correctionParams.precondition &&
                                 primme->correctionParams.maxInnerIterations == 0 &&
                                 !(primme->correctionParams.projectors.RightX && primme->correctionParams.projectors.SkewX) &&
                                 !(primme->correctionParams.projectors.RightX && primme->locking &&
                                       primme->orth == primme_orth_implicit_I) &&
                                 primme->numOrthoConst + numLocked == 0;
               S->fuse_sweeps = 0, S->fusedP_nb = 0;
               SECT(0, CHK(prepare_candidates(S, basisSize, S->V + (size_t)S->ld * basisSize,
                     S->W + (size_t)S->ld * basisSize, 1, maxRecentlyConverged, blockSize,
                     availableBlockSize, numLocked, evals, resNorms, targetShiftIndex, &blockSize,
                     &recentlyConverged, &smallestResNorm, numConverged, &reset, nprevhVecs,
                     practConvCheck)));
               /* valid only if ONE sweep produced exactly the final block, in place and in order */
               if (!(S->fuse_sweeps == 1 && S->fusedP_nb == blockSize && S->fusedP_m == basisSize && blockSize > 0))
                  S->fusedP_nb = 0;
               candidates_prepared = 1;
            } else {
               blockSize = recentlyConverged = 0;
               S->fusedP_nb = 0;
            }

            numConverged += recentlyConverged;
            if (recentlyConverged > 0) S->touch = 0; /* main_iter.c:597-599 */
            if (primme->dynamicMethodSwitch > 0) {
               /* main_iter.c:601-624 */
               if (S->cost.resid_0 == -1.0) S->cost.resid_0 = S->blockNorms[0];
               if (recentlyConverged > 0 || primme->dynamicMethodSwitch == 2) {
                  S->cost.MV = primme->stats.timeMatvec / (double)PB_MAX(primme->stats.numMatvecs, 1);
                  if (pb_dyn_update_statistics(&S->cost, primme, S->tstart, recentlyConverged, 0, numConverged,
                            S->blockNorms[0])) {
                     if (primme->dynamicMethodSwitch == 3)
                        CHK(pb_dyn_switch_from_gdpk(S, &S->cost));
                     else if (primme->dynamicMethodSwitch == 2 || primme->dynamicMethodSwitch == 4)
                        CHK(pb_dyn_switch_from_jdqmr(S, &S->cost));
                  }
               }
            }

            CHK(pb_monitor(S, S->hVals, basisSize, flags, iev, blockSize, S->basisNorms,
                  numConverged, evals, numLocked, S->lockedFlags, resNorms, -1, -1.0, NULL, 0.0,
                  primme_event_outer_iteration));

            if (numConverged >= primme->numEvals ||
                  (primme->locking && numConverged > numLocked &&
                        primme->target != primme_smallest && primme->target != primme_largest &&
                        (!S->numQR || primme->target == primme_closest_geq || primme->target == primme_closest_leq)) ||
                  targetShiftIndex < 0 || (blockSize == 0 && recentlyConverged > 0) ||
                  (S->numQR && fabs(primme->targetShifts[targetShiftIndex] -
                                       primme->targetShifts[PB_MIN(primme->numTargetShifts - 1, numConverged)]) >=
                                       PB_MAX(primme->aNorm, primme->stats.estimateLargestSVal)) ||
                  (numConverged >= nextGuess - primme->numOrthoConst && numGuesses > 0)) {
               break;
            }

            if (blockSize > 0) {
               S->tstart = hl_wtime(); /* the model accumulates the time spent in the correction */
               SECT(1, CHK(solve_correction(S, evals, numLocked, basisSize, blockSize)));
               if (primme->dynamicMethodSwitch > 0) S->cost.time_in_inner += hl_wtime() - S->tstart;
            }

            /* with locking, GD and no preconditioner the practical convergence of the block is
             * judged after orthogonalisation from V_locked' r (main_iter.c:674-797) */
            SCALAR *Rlocked = NULL;
            const int ldRlocked = primme->numOrthoConst + numLocked;
            const int blockSize0 = blockSize;
            if (gdNoPrecLocking)
               Rlocked = (SCALAR *)calloc((size_t)(ldRlocked > 0 ? ldRlocked : 1) * (blockSize > 0 ? blockSize : 1), sizeof(SCALAR));

            for (i = 0; i < maxNumRandoms; i++) {
               int basisSizeOut;
               const int useP0 = i == 0 && S->fusedP_nb == blockSize && S->fusedP_nb > 0;
               SECT(2, CHKX(pb_ortho_block_p0(S, S->V, S->ld, basisSize, basisSize + blockSize - 1, S->evecs,
                     S->ldevecs, primme->numOrthoConst + numLocked, i == 0 ? Rlocked : NULL,
                     ldRlocked, &basisSizeOut, useP0 ? S->fusedP : NULL, S->maxBasis + 8), free(Rlocked)));
               S->fusedP_nb = 0;
               blockSize = basisSizeOut - basisSize;
               if (blockSize > 0 || availableBlockSize <= 0) break;
               CHKX(pb_fill_random(S, S->V + (size_t)S->ld * basisSize, S->ld, 1), free(Rlocked));
               blockSize = 1;
            }
            if (i >= maxNumRandoms) {
               if (availableBlockSize > 0 && blockSize0 <= 0 && reset == 0)
                  wholeSpace = 1;
               else
                  reset = 2;
               blockSize = 0;
               free(Rlocked);
               break;
            }

            if (gdNoPrecLocking) {
               if (numLocked > 0) {
                  for (i = 0; i < blockSize0 && numConverged < primme->numEvals; i++) {
                     double normXx = 0.0;
                     if (primme->orth == primme_orth_explicit_I) {
                        /* Xx = VtBV(0:numLocked, numLocked: ) * hVecs(:,iev[i]) */
                        SCALAR *Xx = (SCALAR *)calloc(numLocked, sizeof(SCALAR));
                        hl_gemm('N', 'N', numLocked, 1, basisSize, 1.0,
                              &S->VtBV[(size_t)S->maxRank * numLocked], S->maxRank,
                              &S->hVecs[(size_t)ldh * iev[i]], ldh, 0.0, Xx, numLocked);
                        normXx = PB_ABS(hl_dot(numLocked, Xx, Xx));
                        free(Xx);
                     }
                     double normR = PB_ABS(hl_dot(ldRlocked, &Rlocked[(size_t)ldRlocked * i],
                           &Rlocked[(size_t)ldRlocked * i]));
                     double bn = S->blockNorms[i];
                     double newBlockNorm = sqrt(PB_MAX(bn * bn - normR * (1. + normXx), 0.0));
                     CHKX(pb_check_convergence(S, S->V + (size_t)S->ld * (basisSize + i), S->ld, 1,
                           NULL, 0, 0, numLocked, 0, 1, &flags[iev[i]], &newBlockNorm,
                           &S->hVals[iev[i]], &reset, -1), free(Rlocked));
                     S->basisNorms[iev[i]] = newBlockNorm;
                     if (flags[iev[i]] == CONVERGED) {
                        flags[iev[i]] = PRACTICALLY_CONVERGED;
                        numConverged++;
                        CHKX(pb_monitor(S, S->hVals, basisSize, flags, &iev[i], 1, S->basisNorms,
                              numConverged, NULL, 0, NULL, NULL, -1, -1.0, NULL, 0.0,
                              primme_event_converged), free(Rlocked));
                     }
                  }
               }
               free(Rlocked);
               Rlocked = NULL;
               if (numConverged > numLocked && primme->target != primme_smallest &&
                     primme->target != primme_largest &&
                     (!S->numQR || primme->target == primme_closest_geq || primme->target == primme_closest_leq))
                  break;
            }

            /* W(:,new) = A V(:,new);  H(:,new) = V' W(:,new) */
            SECT(3, CHK(pb_apply_matvec(S, S->V + (size_t)S->ld * basisSize, S->ld,
                  S->W + (size_t)S->ld * basisSize, S->ld, blockSize)));
            if (S->numQR) {
               int nQ = basisSize;
               CHK(pb_update_Q(S, primme->targetShifts[targetShiftIndex], basisSize, blockSize, &nQ));
               if (basisSize + blockSize != nQ) {
                  blockSize = 0;
                  reset = 1;
                  break;
               }
            }
            SECT(4, CHK(pb_update_projection(S, basisSize, blockSize)));
            CHK(pb_update_QtV(S, basisSize, blockSize));

            hl_copy(S->hVecs, basisSize, basisSize, ldh, S->prevhVecs, ldh);
            hl_zero(&S->prevhVecs[basisSize], maxBasis - basisSize, basisSize, ldh);
            nprevhVecs = basisSize;

            basisSize += blockSize;
            blockSize = 0;

            SECT(5, CHK(solve_projected(S, basisSize, primme->numOrthoConst + numLocked, numConverged)));
            S->numArbitraryVecs = 0;
            candidates_prepared = 0;

            /* |l_0 - tau| <= s_0 must hold for the smallest triplet of R: otherwise the factorisation has
             * accumulated too much error and is rebuilt (:852-884) */
            if (S->refined && basisSize > 0 && restartsSinceReset > 1 && targetShiftIndex >= 0 &&
                  fabs(primme->targetShifts[targetShiftIndex] - S->hVals[0]) -
                              PB_MAX(primme->aNorm, primme->stats.estimateLargestSVal) * PB_EPS >
                        S->hSVals[0]) {
               reset = 2;
               break;
            }
         } /* expansion loop */

         if (basisSize >= primme->n - primme->numOrthoConst - numLocked) {
            if (primme->stats.maxConvTol < primme->stats.estimateResidualError) reset = 1;
         }
         if (reset > 0) break;

         /* ---- make hVecs/iev ready for the restart (main_iter.c:908-1073) ---- */
         if (!candidates_prepared) {
            if (blockSize > 0) {
               availableBlockSize = blockSize;
               maxRecentlyConverged = 0;
            } else if (primme->numTargetShifts > numConverged + 1) {
               if (primme->locking)
                  maxRecentlyConverged = PB_MAX(PB_MIN(primme->numEvals, numLocked + 1) - numConverged, 0);
               else
                  maxRecentlyConverged = PB_MAX(PB_MIN(primme->numEvals, numConverged + 1) - numConverged, 0);
               availableBlockSize = maxRecentlyConverged;
            } else {
               maxRecentlyConverged = PB_MAX(0, primme->numEvals - numConverged);
               availableBlockSize = PB_MAX(0, PB_MIN(primme->maxBlockSize, maxBasis - (numConverged - numLocked)));
               availableBlockSize = PB_MIN(availableBlockSize, maxRecentlyConverged + 1);
            }
            {
               PRIMME_INT room = primme->n - numLocked - primme->numOrthoConst;
               if (room < availableBlockSize) availableBlockSize = (int)room;
               if (availableBlockSize < 0) availableBlockSize = 0;
            }

            if (availableBlockSize <= 0 ||
                  primme->minRestartSize + primme->restartingParams.maxPrevRetain + availableBlockSize < maxBasis ||
                  primme->numOrthoConst + numLocked + basisSize >= primme->n || S->numQR) {
               double dummyZero = 0.0;
               double *srn = (primme->target == primme_closest_abs || primme->target == primme_largest_abs)
                                   ? &dummyZero
                                   : &smallestResNorm;
               CHK(prepare_candidates(S, basisSize, NULL, NULL, 0, maxRecentlyConverged, blockSize,
                     availableBlockSize, numLocked, evals, resNorms, targetShiftIndex, &blockSize,
                     &recentlyConverged, srn, numConverged, &reset, nprevhVecs, 0));

               /* with several shifts, a pair that converged may be the closest to another target: no
                * candidates for the next iteration (:1003-1006) */
               if (S->numQR && numConverged + recentlyConverged > numLocked && primme->numTargetShifts > numLocked + 1)
                  blockSize = 0;

               for (i = 0, numConverged = numLocked; i < basisSize; i++)
                  if (flags[i] != UNCONVERGED && numConverged < primme->numEvals &&
                        (i < primme->numEvals - numLocked || primme->target == primme_closest_geq ||
                              primme->target == primme_closest_leq))
                     numConverged++;

               /* converged pairs and the block first */
               int *iwork = (int *)malloc(sizeof(int) * basisSize);
               int j, k, l, m;
               for (i = k = l = m = 0; i < basisSize; i++) {
                  int inIev = 0;
                  for (j = 0; j < blockSize; j++)
                     if (iev[j] == i) inIev = 1;
                  if ((flags[i] != UNCONVERGED && m++ < numConverged - numLocked) || inIev)
                     iwork[k++] = i;
                  else
                     iwork[numConverged - numLocked + blockSize + l++] = i;
               }
               hl_permute_reals(S->hVals, basisSize, iwork);
               hl_permute_cols(S->hVecs, basisSize, basisSize, ldh, iwork);
               hl_permute_ints(flags, basisSize, iwork);
               if (S->hVecsRot) {
                  hl_zero(&S->hVecsRot[(size_t)ldh * S->numArbitraryVecs], maxBasis, basisSize - S->numArbitraryVecs, ldh);
                  for (i = S->numArbitraryVecs; i < basisSize; i++) S->hVecsRot[(size_t)ldh * i + i] = 1.0;
                  hl_permute_cols(S->hVecsRot, basisSize, basisSize, ldh, iwork);
                  for (i = j = 0; i < basisSize; i++)
                     if (iwork[i] != i) j = i + 1;
                  S->numArbitraryVecs = PB_MAX(S->numArbitraryVecs, j);
               }
               free(iwork);
            } else {
               blockSize = availableBlockSize;
               for (i = 0; i < blockSize; i++) iev[i] = i;
               pb_map_vecs(S->prevhVecs, basisSize, nprevhVecs, ldh, S->hVecs, 0, basisSize, ldh, map);
            }
         }
         if (reset > 0) break;

         hl_permute_cols(S->prevhVecs, basisSize, nprevhVecs, ldh, map);

         {
            const double tr0 = hl_wtime();
            SECT(6, CHK(pb_restart(S, basisSize, &blockSize, evals, resNorms, &numConverged, &numLocked,
                  nprevhVecs, numGuesses, &basisSize, &targetShiftIndex, &restartsSinceReset)));
            S->tRestart += hl_wtime() - tr0;
         }
         restartsSinceReset++;

         /* feed remaining initial guesses into the basis (main_iter.c:1098-1168) */
         if (numGuesses > 0) {
            int numNew = PB_MAX(0, PB_MIN(primme->minRestartSize + numConverged - (nextGuess - primme->numOrthoConst), numGuesses));
            numNew = PB_MAX(0, PB_MIN(basisSize + numNew, maxBasis) - basisSize);
            {
               PRIMME_INT cap = basisSize + numNew + primme->numOrthoConst + numLocked;
               if (cap > primme->n) cap = primme->n;
               cap -= primme->numOrthoConst + numLocked + basisSize;
               numNew = cap > 0 ? (int)cap : 0;
            }
            CHK(pb200_copy_d2d(S->dev, S->evecs + (size_t)S->ldevecs * nextGuess, S->ldevecs,
                  S->V + (size_t)S->ld * basisSize, S->ld, S->n, numNew, PB_ES));
            nextGuess += numNew;
            numGuesses -= numNew;
            int basisSizeOut;
            CHK(pb_ortho_block(S, S->V, S->ld, basisSize, basisSize + numNew - 1, S->evecs,
                  S->ldevecs, numLocked + primme->numOrthoConst, NULL, 0, &basisSizeOut));
            numNew = basisSizeOut - basisSize;
            CHK(pb_apply_matvec(S, S->V + (size_t)S->ld * basisSize, S->ld,
                  S->W + (size_t)S->ld * basisSize, S->ld, numNew));
            if (S->numQR) {
               int nQ = basisSize;
               CHK(pb_update_Q(S, primme->targetShifts[targetShiftIndex], basisSize, numNew, &nQ));
               if (basisSize + numNew != nQ) return PRIMME_UNEXPECTED_FAILURE;
            }
            CHK(pb_update_projection(S, basisSize, numNew));
            CHK(pb_update_QtV(S, basisSize, numNew));
            basisSize += numNew;
            CHK(solve_projected(S, basisSize, primme->numOrthoConst + numLocked, numConverged));
            if (numNew > 0) S->numArbitraryVecs = 0;
         }

         primme->stats.numRestarts++;
         primme->initSize = numConverged;
         /* with few eigenvalues GD+k is evaluated against JDQMR after every restart (main_iter.c:1181-1187) */
         if (primme->dynamicMethodSwitch == 1) {
            S->tstart = hl_wtime();
            S->cost.MV = primme->stats.timeMatvec / (double)PB_MAX(primme->stats.numMatvecs, 1);
            pb_dyn_update_statistics(&S->cost, primme, S->tstart, 0, 1, numConverged, S->blockNorms[0]);
            CHK(pb_dyn_switch_from_gdpk(S, &S->cost));
         }
         for (i = 0; i < maxBasis; i++) map[i] = i;
      } /* restart loop */

      if (reset > 0) continue;

      if (primme->locking) {
         CHK(copy_back_candidates(S, basisSize, evals, resNorms, targetShiftIndex, numConverged, numRet));
         pb_dyn_recommend(primme, &S->cost);
         primme->stats.lockingIssue = LockingProblem;
         *ret = (numConverged == primme->numEvals || wholeSpace) ? 0 : PRIMME_MAIN_ITER_FAILURE;
         goto clean;
      }

      /* no locking: all returned pairs must still pass the test after the last restart */
      restartLimitReached = !(primme->stats.numMatvecs < primme->maxMatvecs &&
                              (primme->maxOuterIterations == 0 ||
                                    primme->stats.numOuterIterations < primme->maxOuterIterations));
      CHK(verify_norms(S, restartLimitReached ? primme->numEvals : numConverged, resNorms, &numConverged));

      if (restartLimitReached || numConverged >= primme->numEvals || wholeSpace) {
         for (i = 0; i < primme->numEvals; i++) {
            evals[i] = S->hVals[i];
            S->perm[i] = i;
         }
         CHK(pb200_copy_d2d(S->dev, S->V, S->ld,
               S->evecs + (size_t)S->ldevecs * primme->numOrthoConst, S->ldevecs, S->n,
               primme->numEvals, PB_ES));
         *numRet = primme->numEvals;
         primme->initSize = numConverged;
         pb_dyn_recommend(primme, &S->cost);
         *ret = numConverged >= primme->numEvals ? 0 : PRIMME_MAIN_ITER_FAILURE;
         goto clean;
      }

      /* some pairs slipped: re-orthogonalise the basis, recompute W = A V, keep iterating */
      CHK(pb_ortho_block(S, S->V, S->ld, 0, basisSize - 1, S->evecs, S->ldevecs,
            primme->numOrthoConst, NULL, 0, &basisSize));
      CHK(pb_apply_matvec(S, S->V, S->ld, S->W, S->ld, basisSize));
      restartsSinceReset = 0;
      reset = 0;
      primme->stats.estimateResidualError = 0.0;
      numConverged = 0;
   } /* verification loop */

clean:
   if (pb_sect_on > 0)
      fprintf(stderr, "primme_b200 host profile, host-only ms per section: candidates %.1f, correction %.1f, ortho %.1f, "
                      "matvec %.1f, projection %.1f, solve_H (without the eigen-solve) %.1f, restart %.1f, initial basis %.1f\n",
            1e3 * pb_sect_host[0], 1e3 * pb_sect_host[1], 1e3 * pb_sect_host[2], 1e3 * pb_sect_host[3],
            1e3 * pb_sect_host[4], 1e3 * pb_sect_host[5], 1e3 * pb_sect_host[6], 1e3 * pb_sect_host[7]);
   if (primme->aNorm <= 0.0)
      primme->aNorm = primme->stats.estimateLargestSVal / primme->stats.estimateInvBNorm;
   /* locked vectors are stored in order of convergence: sort them like evals (:1355-1357) */
   rc = pb200_dpermute_columns(S->dev, S->n,
         S->evecs + (size_t)S->ldevecs * primme->numOrthoConst, S->ldevecs, S->perm,
         primme->initSize);
   return rc;
}
#undef gdNoPrecLocking

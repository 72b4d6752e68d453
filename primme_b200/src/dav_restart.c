/* dav_restart.c -- thick restart with +k retained directions, soft and hard locking, and the
 * convergence test.
 *
 * Restates reference src/eigs/restart.c: restart_Sprimme (:200-494), restart_soft_locking
 * (:598-722), restart_locking (:832-1187), Num_aux_update_VWXR (:1233-1294), restart_RR
 * (:1614-1735), ortho_coefficient_vectors (:2347-2408), compute_residual_columns (:2464-2536);
 * src/eigs/convergence.c: check_convergence (:86-204), check_practical_convergence (:238-281);
 * src/eigs/auxiliary_eigs_normal.c: insertionSort (:525-634).
 * The n-long work of a restart -- V <- V*h, W <- W*h, next block's Ritz vectors and residuals,
 * their norms and the Gram blocks G = V'V, H = V'W of the restarted basis -- is one fused device
 * sweep (pb200_dvwxr).
 */
#include "pb_host.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
#ifndef PB_COMPLEX
int pb_insertion_sort(double newVal, double *evals, double newNorm, double *resNorms,
      int newFlag, int *flags, int *perm, int n, int initialShift, primme_params *primme) {
   int i;
   if (primme->target == primme_smallest) {
      for (i = n; i > 0; i--)
         if (newVal >= evals[i - 1]) break;
   } else if (primme->target == primme_largest) {
      for (i = n; i > 0; i--)
         if (newVal <= evals[i - 1]) break;
   } else {
      /* interior: keep the order of convergence except among pairs of the same shift */
      const int last = primme->numTargetShifts - 1;
      double cur = primme->targetShifts[PB_MIN(last, initialShift + n)];
      for (i = n; i > 0; i--) {
         double ith = primme->targetShifts[PB_MIN(last, initialShift + i - 1)];
         if (ith != cur) break;
         double dn, de;
         switch (primme->target) {
         case primme_closest_geq: dn = newVal - cur, de = evals[i - 1] - cur; break;
         case primme_closest_leq: dn = cur - newVal, de = cur - evals[i - 1]; break;
         case primme_closest_abs: dn = fabs(newVal - cur), de = fabs(evals[i - 1] - cur); break;
         case primme_largest_abs: dn = -fabs(newVal - cur), de = -fabs(evals[i - 1] - cur); break;
         default: return PRIMME_FUNCTION_UNAVAILABLE;
         }
         if (dn >= de) break;
      }
   }
   for (int c = n - 1; c >= i; c--) {
      evals[c + 1] = evals[c];
      if (resNorms) resNorms[c + 1] = resNorms[c];
      if (perm) perm[c + 1] = perm[c];
      if (flags) flags[c + 1] = flags[c];
   }
   evals[i] = newVal;
   if (resNorms) resNorms[i] = newNorm;
   if (perm) perm[i] = n;
   if (flags) flags[i] = newFlag;
   return 0;
}
#endif

/* ------------------------------------------------------------------------------------------
 * Flag pairs [left,right) as converged / unconverged from their residual norms
 * (convergence.c:86-204).  R (device, optional) may lose its components along the locked
 * vectors when the practical-convergence test runs.
 * ---------------------------------------------------------------------------------------- */
int pb_check_convergence(pb_solver *S, SCALAR *X, int64_t ldX, int givenX, SCALAR *R,
      int64_t ldR, int givenR, int numLocked, int left, int right, int *flags,
      double *blockNorms, double *hVals, int *reset, int practConvCheck) {
   primme_params *primme = S->primme;
   if (right <= left) return 0;
   int *toProject = (int *)malloc(sizeof(int) * (right - left));
   int numToProject = 0;
   double tol = PB_MAX(PB_EPS * pb_problem_norm(1, primme), primme->stats.maxConvTol);
   double attainableTol = 0.0;
   if (primme->locking) attainableTol = sqrt((double)(primme->numOrthoConst + numLocked)) * tol;

   for (int i = left; i < right; i++) {
      double shift = primme->numTargetShifts > 0
                           ? primme->targetShifts[PB_MIN(primme->initSize, primme->numTargetShifts - 1)]
                           : 0.0;
      double rn = blockNorms[i - left];
      if ((primme->target == primme_closest_leq && hVals[i] - rn > shift) ||
            (primme->target == primme_closest_geq && hVals[i] + rn < shift)) {
         flags[i] = UNCONVERGED;
         continue;
      }
      if (rn <= primme->stats.maxConvTol) {
         flags[i] = CONVERGED;
         continue;
      }
      int isConv = 0;
      int rc = pb_conv_test(S, hVals[i], (X && givenX) ? X + (size_t)ldX * (i - left) : NULL, rn, &isConv);
      if (rc) {
         free(toProject);
         return rc;
      }
      if (isConv) {
         flags[i] = CONVERGED;
      } else if (rn <= primme->stats.estimateResidualError && reset) {
         /* the residual is at the level of the accumulated error: reset V, W at next restart */
         flags[i] = SKIP_UNTIL_RESTART;
         *reset = 1;
      } else if (primme->locking && numLocked > 0 && practConvCheck >= 0) {
         if (givenR && rn < attainableTol)
            toProject[numToProject++] = i - left;
         else if (flags[i] != PRACTICALLY_CONVERGED)
            flags[i] = UNCONVERGED;
      } else {
         flags[i] = UNCONVERGED;
      }
   }

   if (numToProject > 0) {
      /* || (I - QQ') r || <= tol  => practically converged (convergence.c:238-281) */
      double *norms = (double *)malloc(sizeof(double) * numToProject);
      int rc = pb_ortho_single_iteration(S, S->evecs, primme->numOrthoConst + numLocked,
            S->ldevecs, NULL, 0, R, toProject, numToProject, ldR, norms);
      if (rc) {
         free(norms), free(toProject);
         return rc;
      }
      for (int i = 0; i < numToProject; i++) {
         blockNorms[toProject[i]] = norms[i];
         flags[left + toProject[i]] = norms[i] <= tol ? PRACTICALLY_CONVERGED : UNCONVERGED;
      }
      free(norms);
   }
   free(toProject);
   return 0;
}

/* ------------------------------------------------------------------------------------------
 * Orthonormalise (in the V'V inner product) up to *numPrevRetained coefficient vectors of the
 * previous iteration against the first indexOfPreviousVecs current ones and append them
 * (restart.c:2347-2408).
 * ---------------------------------------------------------------------------------------- */
static int ortho_coefficient_vectors(pb_solver *S, int basisSize, int indexOfPreviousVecs,
      const SCALAR *VtBVblk, int nprevhVecs, int *numPrevRetained) {
   primme_params *primme = S->primme;
   const int ld = S->maxBasis;
   int retained = 0;
   for (int i = 0; i < nprevhVecs && retained < *numPrevRetained &&
                   indexOfPreviousVecs + retained < basisSize;
         i++) {
      if (primme->locking == 0 && S->flags[i] != UNCONVERGED) continue;
      SCALAR R = 0.0;
      long long seed[4];
      for (int t = 0; t < 4; t++) seed[t] = primme->iseed[t];
      int rc = pb_ortho_local(&S->prevhVecs[(size_t)ld * i], ld, &R, 0, 0, S->hVecs, ld,
            indexOfPreviousVecs + retained, basisSize, VtBVblk, S->maxRank, seed);
      for (int t = 0; t < 4; t++) primme->iseed[t] = seed[t];
      if (rc) return PRIMME_UNEXPECTED_FAILURE;
      if (PB_ABS(R) < PB_EPS * sqrt(retained + 1.0)) continue;
      hl_copy(&S->prevhVecs[(size_t)ld * i], basisSize, 1, ld,
            &S->hVecs[(size_t)ld * (indexOfPreviousVecs + retained)], ld);
      retained++;
   }
   *numPrevRetained = retained;
   return 0;
}

/* ------------------------------------------------------------------------------------------
 * The fused restart sweep + host bookkeeping of VtBV (restart.c:1233-1294).
 * h is S->hVecs (basisSize x nh).  evecsSize = numOrthoConst + numLocked (columns already in
 * evecs); columns [x2b,x2e) of V*h are appended to evecs.
 * ---------------------------------------------------------------------------------------- */
static int aux_update_VWXR(pb_solver *S, int basisSize, int nh, int restartSize, SCALAR *X1,
      int x1b, int x1e, int evecsSize, int x2b, int x2e, SCALAR *Rout, double *Rnorms, double *rnorms,
      int rb, int re, int Hfull) {
   primme_params *primme = S->primme;
   pb200_vwxr_out o;
   memset(&o, 0, sizeof(o));
   const double t0 = hl_wtime();
   o.X[0].ptr = PB_DP(S->V), o.X[0].ld = S->ld, o.X[0].cb = 0, o.X[0].ce = restartSize;
   if (X1 && x1e > x1b) o.X[1].ptr = PB_DP(X1), o.X[1].ld = S->ld, o.X[1].cb = x1b, o.X[1].ce = x1e;
   if (x2e > x2b)
      o.X[2].ptr = PB_DP(S->evecs + (size_t)S->ldevecs * evecsSize), o.X[2].ld = S->ldevecs,
      o.X[2].cb = x2b, o.X[2].ce = x2e;
   o.Wo.ptr = PB_DP(S->W), o.Wo.ld = S->ld, o.Wo.cb = 0, o.Wo.ce = restartSize;
   if (Rout && x1e > x1b) o.R.ptr = PB_DP(Rout), o.R.ld = S->ld, o.R.cb = x1b, o.R.ce = x1e;
   o.Rnorms_host = Rnorms;
   if (rnorms && re > rb) o.rb = rb, o.re = re, o.rnorms_host = rnorms;
   SCALAR *Gtmp = NULL;
   if (S->VtBV) {
      o.nG = restartSize;
      o.G_host = PB_DP(&S->VtBV[(size_t)S->maxRank * evecsSize + evecsSize]);
      o.ldG = S->maxRank;
   }
   if (Hfull) o.nH = restartSize, o.H_host = PB_DP(S->H), o.ldH = S->maxBasis;

   /* cross block VtBV(0:evecsSize, new) = VtBV(0:evecsSize, old) * h: needs the old columns,
    * which the G output overwrites only below row evecsSize -- compute it first */
   SCALAR *cross = NULL;
   if (S->VtBV && evecsSize > 0) {
      cross = (SCALAR *)calloc((size_t)evecsSize * restartSize, sizeof(SCALAR));
      hl_gemm('N', 'N', evecsSize, restartSize, basisSize, 1.0,
            &S->VtBV[(size_t)S->maxRank * evecsSize], S->maxRank, S->hVecs, S->maxBasis, 0.0,
            cross, evecsSize);
   }

   int rc = pb200_dvwxr(S->dev, S->n, S->V, S->W, basisSize, S->ld, S->hVecs, S->maxBasis, nh,
         S->hVals, &o);
   if (rc) {
      free(cross);
      return rc;
   }
   if (primme->numProcs > 1 && pb200_ctx_nranks(S->dev) <= 1) {
      /* host-callback reduction of everything the sweep produced */
      int nR = o.R.ptr ? o.R.ce - o.R.cb : 0, nr = o.rnorms_host ? re - rb : 0;
      if (o.G_host) rc = pb_reduce_panel(S, (SCALAR *)o.G_host, o.nG, o.nG, o.ldG);
      if (!rc && o.H_host) rc = pb_reduce_panel(S, (SCALAR *)o.H_host, o.nH, o.nH, o.ldH);
      if (!rc && Rnorms && nR) {
         for (int i = 0; i < nR; i++) Rnorms[i] *= Rnorms[i];
         rc = pb_global_sum(S, Rnorms, nR);
         for (int i = 0; i < nR; i++) Rnorms[i] = sqrt(Rnorms[i]);
      }
      if (!rc && nr) {
         for (int i = 0; i < nr; i++) rnorms[i] *= rnorms[i];
         rc = pb_global_sum(S, rnorms, nr);
         for (int i = 0; i < nr; i++) rnorms[i] = sqrt(rnorms[i]);
      }
   }
   /* don't trust residual norms below the error level (restart.c:1270-1272) */
   if (rnorms)
      for (int i = 0; i < re - rb; i++)
         rnorms[i] = PB_MAX(rnorms[i], primme->stats.estimateResidualError);
   if (cross) {
      hl_copy(cross, evecsSize, restartSize, evecsSize, &S->VtBV[(size_t)S->maxRank * evecsSize],
            S->maxRank);
      free(cross);
   }
   (void)Gtmp;
   primme->stats.timeDense += hl_wtime() - t0;
   primme->stats.flopsDense += 2.0 * (double)S->n * basisSize * nh;
   return rc;
}

/* ------------------------------------------------------------------------------------------
 * restart without locking (restart.c:598-722)
 * ---------------------------------------------------------------------------------------- */
static int restart_soft_locking(pb_solver *S, int *restartSize, int basisSize, int *restartPerm,
      int *ievSize, double *evals, double *resNorms, int *numConverged, int numPrevRetained,
      int *indexOfPreviousVecs, int *hVecsPerm) {
   primme_params *primme = S->primme;
   int *flags = S->flags, *iev = S->iev;
   double *hVals = S->hVals;
   int i, j, k;

   /* pairs flagged converged whose Ritz value drifted by more than their residual norm are
    * targeted again (:621-633) */
   *numConverged = 0;
   for (i = 0; i < primme->numEvals; i++) {
      if (flags[i] != UNCONVERGED && fabs(hVals[i] - evals[i]) > resNorms[i]) {
         flags[i] = UNCONVERGED;
      } else if (flags[i] != UNCONVERGED) {
         if (flags[i] == CONVERGED) {
            if (*numConverged == 0) primme->stats.maxConvTol = 0.0;
            primme->stats.maxConvTol = PB_MAX(primme->stats.maxConvTol, resNorms[i]);
         }
         (*numConverged)++;
      }
   }

   *indexOfPreviousVecs = *restartSize;
   *restartSize += numPrevRetained;
   {
      int v = PB_MIN(*ievSize, primme->maxBlockSize);
      v = PB_MIN(v, primme->numEvals - *numConverged + 1);
      v = PB_MIN(v, primme->maxBasisSize - *restartSize);
      v = PB_MIN(v, basisSize - *numConverged);
      v = PB_MIN(v, primme->minRestartSize - *numConverged);
      *ievSize = PB_MAX(0, v);
   }

   /* converged pairs first, the rest after them, original order kept within each group */
   for (i = j = k = 0; i < basisSize; i++) {
      if (k >= *numConverged || flags[i] == UNCONVERGED)
         restartPerm[*numConverged + j++] = i;
      else
         restartPerm[k++] = i;
   }
   hl_permute_reals(hVals, basisSize, restartPerm);
   hl_permute_cols(S->hVecs, basisSize, basisSize, S->maxBasis, restartPerm);

   SCALAR *X = S->V + (size_t)S->ld * *restartSize;
   SCALAR *R = S->W + (size_t)S->ld * *restartSize;
   int rc = aux_update_VWXR(S, basisSize, *restartSize, *restartSize, X, *numConverged,
         *numConverged + *ievSize, primme->numOrthoConst, 0, S->evecsHat ? *numConverged : 0 /* see below */,
         R, S->blockNorms, NULL, 0, 0, primme->orth == primme_orth_explicit_I);
   if (rc) return rc;
   /* NOTE: the reference also copies the converged Ritz vectors into evecs here
    * (restart.c:694: X2 = columns [0,numConverged)); without locking evecs is overwritten by
    * V(:,0:numEvals) before returning (main_iter.c:1285), and nothing reads it in between for
    * B = I unless the skew-Q projector is on (evecsHat: the copy feeds K^{-1}Q and M at the end of the
    * restart, pb_skew_evecs_after_restart), so the copy is made only then. */

   for (i = 0; i < basisSize; i++) hVecsPerm[restartPerm[i]] = i;
   for (i = 0; i < *ievSize; i++)
      for (j = 0; j < *restartSize; j++)
         if (hVecsPerm[j] == *numConverged + i) iev[i] = j;
   return 0;
}

/* Device helper for hard locking: see reference compute_residual_columns (restart.c:2464-2536).
 *   x(:,0:n) <- x(:,p), Ax likewise; (xd, rd)(:,id) for id < nd = next block column taken either
 *   from (xo, ro) (when pd[id] == next xo index) or from a failed pair (x(:,p[i]), its residual). */
static int residual_columns(pb_solver *S, const double *evals, SCALAR *x, int n, const int *p,
      SCALAR *Ax, const SCALAR *xo, int no, const SCALAR *ro, SCALAR *xd, int nd, const int *pd,
      SCALAR *rd) {
   const int64_t ld = S->ld, N = S->n;
   SCALAR *tmp = NULL;
   if (nd > 0) CHK(pb200_malloc(S->dev, sizeof(SCALAR) * (size_t)(N > 0 ? N : 1) * nd * 2, (void **)&tmp));
   SCALAR *XD = tmp, *RD = tmp ? tmp + (size_t)N * nd : NULL;
   int i = 0, id = 0, io = 0, rc = 0;
   if (n == 0) {
      int c = PB_MIN(no, nd);
      if (c > 0) {
         rc = pb200_copy_d2d(S->dev, xo, ld, XD, N, N, c, PB_ES);
         if (!rc) rc = pb200_copy_d2d(S->dev, ro, ld, RD, N, N, c, PB_ES);
         if (!rc) rc = pb200_copy_d2d(S->dev, XD, N, xd, ld, N, c, PB_ES);
         if (!rc) rc = pb200_copy_d2d(S->dev, RD, N, rd, ld, N, c, PB_ES);
      }
      goto done;
   }
   for (i = id = io = 0; (i < n || id < nd) && !rc; id++) {
      if (id < nd && io < no && pd[id] == io) {
         rc = pb200_copy_d2d(S->dev, xo + (size_t)ld * io, ld, XD + (size_t)N * id, N, N, 1, PB_ES);
         if (!rc) rc = pb200_copy_d2d(S->dev, ro + (size_t)ld * io, ld, RD + (size_t)N * id, N, N, 1, PB_ES);
         io++;
      } else {
         if (id < nd && i < n) {
            SCALAR alpha = -evals[p[i]];
            rc = pb200_copy_d2d(S->dev, x + (size_t)ld * p[i], ld, XD + (size_t)N * id, N, N, 1, PB_ES);
            if (!rc) rc = pb200_copy_d2d(S->dev, Ax + (size_t)ld * p[i], ld, RD + (size_t)N * id, N, N, 1, PB_ES);
            if (!rc) rc = pb200_daxpy_columns(S->dev, N, &alpha, x + (size_t)ld * p[i], ld, RD + (size_t)N * id, N, 1);
         }
         i++;
      }
   }
   /* compaction of the failed pairs to the front (p is increasing: forward copies are safe) */
   for (i = 0; i < n && !rc; i++) {
      if (p[i] == i) continue;
      rc = pb200_copy_d2d(S->dev, x + (size_t)ld * p[i], ld, x + (size_t)ld * i, ld, N, 1, PB_ES);
      if (!rc) rc = pb200_copy_d2d(S->dev, Ax + (size_t)ld * p[i], ld, Ax + (size_t)ld * i, ld, N, 1, PB_ES);
   }
   if (!rc && nd > 0) {
      rc = pb200_copy_d2d(S->dev, XD, N, xd, ld, N, nd, PB_ES);
      if (!rc) rc = pb200_copy_d2d(S->dev, RD, N, rd, ld, N, nd, PB_ES);
   }
done:
   if (tmp) pb200_free(S->dev, tmp);
   return rc;
}

/* ------------------------------------------------------------------------------------------
 * restart with (hard) locking (restart.c:832-1187)
 * ---------------------------------------------------------------------------------------- */
static int restart_locking(pb_solver *S, int *restartSize, int basisSize, int *restartPerm,
      int *ievSize, double *evals, double *resNorms, int *numConverged, int *numLocked,
      int numPrevRetained, int *indexOfPreviousVecs, int *hVecsPerm) {
   primme_params *primme = S->primme;
   int *flags = S->flags, *iev = S->iev;
   double *hVals = S->hVals, *blockNorms = S->blockNorms;
   const int ldh = S->maxBasis;
   int i, j, k, numPacked, failed;
   const int numLocked0 = *numLocked;

   int maxBlockSize = PB_MAX(0, PB_MIN(PB_MIN(*restartSize, primme->maxBlockSize),
                                      primme->numEvals - *numConverged + 1));
   int sizeBlockNorms = PB_MAX(0, PB_MIN(maxBlockSize,
         primme->maxBasisSize - *restartSize - numPrevRetained - *numConverged + *numLocked));
   *indexOfPreviousVecs = *restartSize;
   const int left = *restartSize + numPrevRetained;

   for (i = k = numPacked = 0; i < basisSize; i++) {
      if (flags[i] != UNCONVERGED && numPacked < primme->numEvals - *numLocked &&
            (i < primme->numEvals - *numLocked || primme->target == primme_closest_geq ||
                  primme->target == primme_closest_leq)) {
         restartPerm[left + numPacked++] = i;
      } else if (k < left) {
         restartPerm[k++] = i;
      } else {
         restartPerm[PB_MIN(*numConverged, primme->numEvals) - *numLocked + k++] = i;
      }
   }
   *restartSize = left + numPacked;
   hl_permute_reals(hVals, basisSize, restartPerm);
   hl_permute_cols(S->hVecs, basisSize, basisSize, ldh, restartPerm);

   double *lockedResNorms = &resNorms[*numLocked];
   SCALAR *X = S->V + (size_t)S->ld * *restartSize;
   SCALAR *R = S->W + (size_t)S->ld * *restartSize;
   int rc = aux_update_VWXR(S, basisSize, *restartSize, *restartSize, X, 0, sizeBlockNorms,
         *numLocked + primme->numOrthoConst, left, left + numPacked, R, blockNorms,
         lockedResNorms, left, *restartSize, primme->orth == primme_orth_explicit_I);
   if (rc) return rc;

   /* re-test the pairs about to be locked with their fresh residual norms (:947-950) */
   hl_permute_ints(flags, basisSize, restartPerm);
   rc = pb_check_convergence(S, S->V + (size_t)S->ld * left, S->ld, 1, NULL, 0, 0, *numLocked,
         left, left + numPacked, flags, lockedResNorms, hVals, NULL, 0);
   if (rc) return rc;

   for (i = left, j = 0; i < left + numPacked; i++) {
      if (flags[i] != UNCONVERGED && *numLocked + j < primme->numEvals)
         evals[*numLocked + j++] = hVals[i];
      else
         flags[i] = UNCONVERGED;
   }

   int *ifailed = (int *)malloc(sizeof(int) * (numPacked > 0 ? numPacked : 1));
   for (i = left, failed = 0; i < left + numPacked; i++)
      if (flags[i] == UNCONVERGED) ifailed[failed++] = i - left;
   for (i = left, j = 0; i < left + numPacked; i++)
      if (flags[i] != UNCONVERGED) ifailed[failed + j++] = i - left;

   /* pairs that failed to lock rejoin the basis and compete for the next block (:986-1064) */
   maxBlockSize = PB_MAX(0, PB_MIN(maxBlockSize,
         primme->maxBasisSize - *restartSize - numPrevRetained - *numConverged + *numLocked));
   double *blockNorms0 = (double *)malloc(sizeof(double) * (sizeBlockNorms > 0 ? sizeBlockNorms : 1));
   for (i = 0; i < sizeBlockNorms; i++) blockNorms0[i] = blockNorms[i];
   for (i = j = k = 0; i < *indexOfPreviousVecs || j < failed; k++) {
      if (i < *indexOfPreviousVecs &&
            (j >= failed || restartPerm[i] < restartPerm[left + ifailed[j]])) {
         if (k < maxBlockSize && i < sizeBlockNorms) blockNorms[k] = blockNorms0[i];
         hVecsPerm[k] = i++;
      } else {
         if (k < maxBlockSize) blockNorms[k] = resNorms[numLocked0 + ifailed[j]];
         hVecsPerm[k] = left + j++;
      }
   }
   free(blockNorms0);
   for (i = 0; i < numPrevRetained; i++) hVecsPerm[k++] = i + *indexOfPreviousVecs;
   for (; k < basisSize; k++) hVecsPerm[k] = -1;

   rc = residual_columns(S, &hVals[left], S->V + (size_t)S->ld * left, failed, ifailed,
         S->W + (size_t)S->ld * left, X, sizeBlockNorms, R,
         S->V + (size_t)S->ld * (left + failed), maxBlockSize, hVecsPerm,
         S->W + (size_t)S->ld * (left + failed));
   if (rc) {
      free(ifailed);
      return rc;
   }

   /* same rearrangement on the small matrices */
   {
      /* compact hVecs(:,left+ifailed) and hVals to the front of the packed range */
      SCALAR *tmp = (SCALAR *)malloc(sizeof(SCALAR) * (size_t)basisSize * (failed > 0 ? failed : 1));
      for (i = 0; i < failed; i++)
         memcpy(tmp + (size_t)basisSize * i, &S->hVecs[(size_t)ldh * (left + ifailed[i])], sizeof(SCALAR) * basisSize);
      for (i = 0; i < failed; i++)
         memcpy(&S->hVecs[(size_t)ldh * (left + i)], tmp + (size_t)basisSize * i, sizeof(SCALAR) * basisSize);
      for (i = 0; i < failed; i++) tmp[i] = hVals[left + ifailed[i]];
      for (i = 0; i < failed; i++) hVals[left + i] = PB_REAL(tmp[i]);
      free(tmp);
      hl_permute_ints(&restartPerm[left], numPacked, ifailed);
   }

   if (S->VtBV) {
      /* reorder rows/columns of the Gram matrix: [locked-now | kept | failed] (:1088-1110) */
      const int ldG = S->maxRank, nLocked = primme->numOrthoConst + *numLocked;
      const int nc = left + numPacked, nG = nLocked + nc;
      int *iV = (int *)malloc(sizeof(int) * (nc > 0 ? nc : 1));
      for (i = 0; i < numPacked - failed; i++) iV[i] = ifailed[failed + i] + left;
      for (i = 0; i < left; i++) iV[i + numPacked - failed] = i;
      for (i = 0; i < failed; i++) iV[i + left + numPacked - failed] = ifailed[i] + left;
      /* the sweep wrote only the upper triangle's worth reliably? G is written full: use it */
      SCALAR *rw = (SCALAR *)malloc(sizeof(SCALAR) * (size_t)nG * (nc > 0 ? nc : 1));
      for (i = 0; i < nc; i++)
         memcpy(rw + (size_t)nG * i, &S->VtBV[(size_t)ldG * (nLocked + iV[i])], sizeof(SCALAR) * nG);
      hl_copy(rw, nLocked, nc, nG, &S->VtBV[(size_t)ldG * nLocked], ldG);
      for (j = 0; j < nc; j++)
         for (i = 0; i < nc; i++)
            S->VtBV[(size_t)ldG * (nLocked + j) + nLocked + i] = rw[(size_t)nG * j + nLocked + iV[i]];
      free(rw), free(iV);
   }
   if (primme->orth == primme_orth_explicit_I) {
      /* H: failed pairs move right after the kept ones (:1114-1119) */
      const int ldH = S->maxBasis;
      SCALAR *rw = (SCALAR *)malloc(sizeof(SCALAR) * (size_t)(left + numPacked) * (failed > 0 ? failed : 1));
      for (i = 0; i < failed; i++)
         memcpy(rw + (size_t)(left + numPacked) * i, &S->H[(size_t)ldH * (left + ifailed[i])], sizeof(SCALAR) * (left + numPacked));
      for (i = 0; i < failed; i++)
         memcpy(&S->H[(size_t)ldH * (left + i)], rw + (size_t)(left + numPacked) * i, sizeof(SCALAR) * (left + numPacked));
      /* rows */
      for (j = 0; j < left + failed; j++) {
         for (i = 0; i < failed; i++) rw[i] = S->H[(size_t)ldH * j + left + ifailed[i]];
         for (i = 0; i < failed; i++) S->H[(size_t)ldH * j + left + i] = rw[i];
      }
      free(rw);
   }

   /* pack the vectors that really converged inside evecs and sort their values (:1127-1166) */
   for (i = left; i < left + numPacked; i++) {
      if (flags[i] != UNCONVERGED && *numLocked < primme->numEvals) {
         double resNorm = resNorms[*numLocked] = lockedResNorms[i - left];
         double eval = evals[*numLocked];
         int src = numLocked0 + i - left + primme->numOrthoConst;
         int dst = *numLocked + primme->numOrthoConst;
         if (src != dst)
            CHKX(pb200_copy_d2d(S->dev, S->evecs + (size_t)S->ldevecs * src, S->ldevecs,
                  S->evecs + (size_t)S->ldevecs * dst, S->ldevecs, S->n, 1, PB_ES), free(ifailed));
         (*numLocked)++;
         if (S->lockedFlags) S->lockedFlags[*numLocked - 1] = flags[i];
         CHKX(pb_monitor(S, NULL, 0, NULL, NULL, 0, NULL, 0, evals, *numLocked, S->lockedFlags,
               resNorms, -1, 0.0, NULL, 0.0, primme_event_locked), free(ifailed));
         CHKX(pb_insertion_sort(eval, evals, resNorm, resNorms, flags[i], S->lockedFlags, S->perm,
               *numLocked - 1, 0, primme), free(ifailed));
         if (flags[i] == CONVERGED)
            primme->stats.maxConvTol = PB_MAX(primme->stats.maxConvTol, resNorm);
      }
   }
   free(ifailed);

   *restartSize = left + failed;
   *ievSize = PB_MIN(maxBlockSize, sizeBlockNorms + failed);
   *numConverged = *numLocked;
   for (i = 0; i < *ievSize; i++) iev[i] = i;
   for (i = 0; i < basisSize; i++) flags[i] = UNCONVERGED;
   return 0;
}

/* ------------------------------------------------------------------------------------------
 * Projected matrices after the restart (restart.c:1614-1735, RR only)
 * ---------------------------------------------------------------------------------------- */
int pb_compute_submatrix(const SCALAR *X, int nX, int ldX, const SCALAR *H, int nH, int ldH,
      SCALAR *R, int ldR) {
   if (nH == 0 || nX == 0) return 0;
   SCALAR *rw = (SCALAR *)calloc((size_t)nH * nX, sizeof(SCALAR));
   hl_symm_lu(nH, nX, 1.0, H, ldH, X, ldX, 0.0, rw, nH);
   hl_gemm('C', 'N', nX, nX, nH, 1.0, X, ldX, rw, nH, 0.0, R, ldR);
   free(rw);
   return 0;
}

static int restart_RR(pb_solver *S, int restartSize, int basisSize, int numConverged,
      int numPrevRetained, int indexOfPreviousVecs, const int *hVecsPerm, int *targetShiftIndex) {
   primme_params *primme = S->primme;
   const int ldH = S->maxBasis, ldh = S->maxBasis, ldG = S->maxRank;
   SCALAR *H = S->H, *hVecs = S->hVecs;
   double *hVals = S->hVals;
   double aNorm = PB_MAX(primme->aNorm, primme->stats.estimateLargestSVal);
   int i, j;

   if (primme->orth == primme_orth_implicit_I) {
      /* H = diag(hVals) except the block of the retained directions (:1636-1664) */
      SCALAR *sub = (SCALAR *)calloc((size_t)(numPrevRetained > 0 ? numPrevRetained : 1) * (numPrevRetained > 0 ? numPrevRetained : 1), sizeof(SCALAR));
      pb_compute_submatrix(&hVecs[(size_t)ldh * indexOfPreviousVecs], numPrevRetained, ldh, H,
            basisSize, ldH, sub, numPrevRetained > 0 ? numPrevRetained : 1);
      hl_zero(H, restartSize, restartSize, ldH);
      for (j = 0; j < numPrevRetained; j++)
         for (i = 0; i < numPrevRetained; i++)
            H[(size_t)ldH * (indexOfPreviousVecs + j) + indexOfPreviousVecs + i] =
                  sub[(size_t)numPrevRetained * j + i];
      free(sub);
      for (j = 0; j < indexOfPreviousVecs; j++) H[(size_t)ldH * j + j] = hVals[j];
      for (j = indexOfPreviousVecs + numPrevRetained; j < restartSize; j++)
         H[(size_t)ldH * j + j] = hVals[j];
   }

   const int nLocked = primme->numOrthoConst + (primme->locking ? numConverged : 0);
   if (targetShiftIndex && primme->targetShifts &&
         (*targetShiftIndex < 0 ||
               fabs(primme->targetShifts[*targetShiftIndex] -
                     primme->targetShifts[PB_MIN(primme->numTargetShifts - 1, numConverged)]) >
                     PB_EPS * aNorm)) {
      /* the target shift moves: order everything again (:1680-1692) */
      *targetShiftIndex = PB_MIN(primme->numTargetShifts - 1, numConverged);
      return pb_solve_H(S, H, ldH, restartSize,
            S->VtBV ? &S->VtBV[(size_t)ldG * nLocked + nLocked] : NULL, ldG, hVecs, ldh, hVals,
            numConverged, 1);
   }

   int ordered = restartSize;
   for (i = 0; i < restartSize; i++)
      if (hVecsPerm[i] == indexOfPreviousVecs) {
         ordered = i;
         break;
      }

   /* Ritz vectors are canonical vectors except inside the retained block (:1716-1722) */
   for (j = 0; j < restartSize; j++) {
      for (i = 0; i < restartSize; i++) hVecs[(size_t)ldh * j + i] = 0.0;
      hVecs[(size_t)ldh * j + hVecsPerm[j]] = 1.0;
   }
   hl_permute_reals(hVals, restartSize, hVecsPerm);

   if (numPrevRetained > 0) {
      /* small eigenproblem of the retained block; note the output goes to a shifted window
       * of hVecs, so solve into a scratch matrix first */
      SCALAR *sub = (SCALAR *)malloc(sizeof(SCALAR) * numPrevRetained * numPrevRetained);
      double *w = (double *)malloc(sizeof(double) * numPrevRetained);
      int rc = pb_solve_H(S, &H[(size_t)ldH * indexOfPreviousVecs + indexOfPreviousVecs], ldH,
            numPrevRetained,
            S->VtBV ? &S->VtBV[(size_t)ldG * (nLocked + indexOfPreviousVecs) + nLocked + indexOfPreviousVecs] : NULL,
            ldG, sub, numPrevRetained, w, numConverged, 1);
      if (!rc) {
         for (j = 0; j < numPrevRetained; j++) {
            for (i = 0; i < numPrevRetained; i++)
               hVecs[(size_t)ldh * (ordered + j) + indexOfPreviousVecs + i] = sub[(size_t)numPrevRetained * j + i];
            hVals[ordered + j] = w[j];
         }
      }
      free(sub), free(w);
      if (rc) return rc;
   }
   return 0;
}

/* ------------------------------------------------------------------------------------------
 * restart_Sprimme (restart.c:200-446)
 * ---------------------------------------------------------------------------------------- */
int pb_restart(pb_solver *S, int basisSize, int *ievSize, double *evals, double *resNorms,
      int *numConverged, int *numLocked, int nprevhVecs, int numGuesses, int *restartSizeOut,
      int *targetShiftIndex, int *restartsSinceReset) {
   primme_params *primme = S->primme;
   int *flags = S->flags;
   int i, restartSize;

   for (i = 0, *numConverged = *numLocked; i < basisSize; i++) {
      if (flags[i] == SKIP_UNTIL_RESTART)
         flags[i] = UNCONVERGED;
      else if (flags[i] != UNCONVERGED && *numConverged < primme->numEvals &&
               (i < primme->numEvals - *numLocked || primme->target == primme_closest_geq ||
                     primme->target == primme_closest_leq))
         (*numConverged)++;
   }

   int numPrevRetained = primme->restartingParams.maxPrevRetain;
   if (!primme->locking && basisSize + *numLocked + primme->numOrthoConst >= primme->n) {
      restartSize = basisSize, numPrevRetained = 0;
   } else if (basisSize <= primme->maxBasisSize - primme->maxBlockSize) {
      restartSize = basisSize, numPrevRetained = 0; /* basis not full: keep everything */
   } else {
      restartSize = PB_MIN(basisSize, primme->minRestartSize);
   }
   restartSize -= PB_MIN(PB_MIN(numGuesses, *numConverged - *numLocked), restartSize);
   if (primme->locking) restartSize = PB_MIN(restartSize, basisSize - (*numConverged - *numLocked));
   {
      PRIMME_INT lim = primme->n - restartSize - *numConverged - primme->numOrthoConst;
      int v = PB_MIN(numPrevRetained, primme->maxBasisSize - restartSize - 1);
      if (lim < v) v = (int)lim;
      numPrevRetained = PB_MAX(0, v);
   }

   int indexOfPreviousVecs =
         primme->locking ? restartSize + *numConverged - *numLocked : restartSize;
   const int indexOfPreviousVecsBeforeRestart = indexOfPreviousVecs;
   const int nLocked = primme->numOrthoConst + *numLocked;
   CHK(ortho_coefficient_vectors(S, basisSize, indexOfPreviousVecs,
         S->VtBV ? &S->VtBV[(size_t)S->maxRank * nLocked + nLocked] : NULL, nprevhVecs,
         &numPrevRetained));

   int *restartPerm = (int *)malloc(sizeof(int) * basisSize);
   int *hVecsPerm = (int *)malloc(sizeof(int) * basisSize);
   int rc;
   if (!primme->locking)
      rc = restart_soft_locking(S, &restartSize, basisSize, restartPerm, ievSize, evals, resNorms,
            numConverged, numPrevRetained, &indexOfPreviousVecs, hVecsPerm);
   else
      rc = restart_locking(S, &restartSize, basisSize, restartPerm, ievSize, evals, resNorms,
            numConverged, numLocked, numPrevRetained, &indexOfPreviousVecs, hVecsPerm);
   if (rc) {
      free(restartPerm), free(hVecsPerm);
      return rc;
   }

   if (S->fVtBV) {
      int newnLocked = primme->numOrthoConst + *numLocked;
      rc = pb_update_cholesky(S, nLocked, newnLocked + restartSize);
   }

   /* previous Ritz values follow the permutation (interior targets only, :355-369) */
   if (!rc && primme->target != primme_smallest && primme->target != primme_largest) {
      double *prv = S->prevRitzVals;
      if (S->numPrevRitzVals > 0) {
         for (i = S->numPrevRitzVals; i < basisSize; i++) prv[i] = prv[S->numPrevRitzVals - 1];
         hl_permute_reals(prv, basisSize, restartPerm);
      }
      for (i = 0; i < restartSize; i++)
         if (restartPerm[i] >= S->numPrevRitzVals) prv[i] = S->hVals[i];
      hl_permute_reals(prv, restartSize, hVecsPerm);
      S->numPrevRitzVals = restartSize;
   }

   if (!rc && S->QtV)
      rc = pb_restart_harmonic(S, restartSize, basisSize, *numConverged, targetShiftIndex);
   else if (!rc && S->refined)
      rc = pb_restart_refined(S, restartSize, basisSize, *numConverged, numPrevRetained, indexOfPreviousVecs,
            indexOfPreviousVecsBeforeRestart, restartPerm, hVecsPerm, targetShiftIndex);
   else if (!rc)
      rc = restart_RR(S, restartSize, basisSize, *numConverged, numPrevRetained,
            indexOfPreviousVecs, hVecsPerm, targetShiftIndex);
   free(restartPerm);
   if (!rc) rc = pb_skew_evecs_after_restart(S, *numConverged);

   /* all wanted pairs converged: bring them to the front of V (:384-392) */
   if (!rc && *numConverged >= primme->numEvals && !primme->locking) {
      rc = pb200_dpermute_columns(S->dev, S->n, S->V, S->ld, hVecsPerm, restartSize);
      if (!rc) rc = pb200_dpermute_columns(S->dev, S->n, S->W, S->ld, hVecsPerm, restartSize);
   }
   free(hVecsPerm);
   if (rc) return rc;
   *restartSizeOut = restartSize;

   /* loss of orthogonality -> bound on the attainable residual norm (:402-446) */
   double fn = 0.0;
   if (S->VtBV) {
      double acc = 0.0;
      const int ldG = S->maxRank, nG = primme->numOrthoConst + *numLocked + restartSize;
      for (i = 0; i < nG; i++)
         for (int j = 0; j < i; j++) {
            SCALAR g = S->VtBV[(size_t)i * ldG + j];
            acc += 2 * PB_REAL(PB_CONJ(g) * g) / PB_ABS(S->VtBV[(size_t)i * ldG + i]) / PB_ABS(S->VtBV[(size_t)j * ldG + j]);
         }
      fn = sqrt(acc);
   }
   if (fn > 0.0) {
      if (*restartsSinceReset <= 1)
         primme->stats.maxConvTol =
               PB_MAX(primme->stats.maxConvTol, fn * primme->stats.estimateLargestSVal);
      primme->stats.estimateResidualError =
            sqrt((double)*restartsSinceReset) * fn * pb_problem_norm(1, primme);
   } else {
      primme->stats.estimateResidualError =
            2 * sqrt((double)*restartsSinceReset) * PB_EPS * pb_problem_norm(1, primme);
   }
   return 0;
}

/* dav_project.c -- Rayleigh-Ritz projection: H = V'AV panel update and the small eigenproblem.
 *
 * Restates reference src/eigs/update_projection.c:81-165 (new block columns of H through one
 * fused row sweep on the device) and src/eigs/solve_projection.c:95-331 (solve_H, solve_H_RR:
 * ?sygvx on (H, V'V), ordering by target), :1009-1064 (map_vecs).
 */
#include "pb_host.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Sum a host panel over the ranks unless the kernel layer already did it on the device. */
int pb_reduce_panel(pb_solver *S, SCALAR *P, int rows, int cols, int ldp) {
   primme_params *primme = S->primme;
   if (primme->numProcs <= 1 || pb200_ctx_nranks(S->dev) > 1) return 0;
   int cnt = rows * cols;
   if (cnt <= 0) return 0;
   const int rpe = PB_ES / 8; /* reals per element (globalSumReal counts reals, auxiliary_eigs.c:371-373) */
   if (ldp == rows) return pb_global_sum(S, (double *)P, cnt * rpe);
   SCALAR *tmp = (SCALAR *)malloc(sizeof(SCALAR) * cnt);
   if (!tmp) return PRIMME_MALLOC_FAILURE;
   hl_copy(P, rows, cols, ldp, tmp, rows);
   int r = pb_global_sum(S, (double *)tmp, cnt * rpe);
   hl_copy(tmp, rows, cols, rows, P, ldp);
   free(tmp);
   return r;
}

/* H(0:m, numCols:m) = V(:,0:m)' * W(:,numCols:m), m = numCols + blockSize
 * (reference update_projection.c:99-102; only the upper triangle is used afterwards). */
static int update_projection_impl(pb_solver *S, int numCols, int blockSize) {
   if (blockSize <= 0) return 0;
   const int m = numCols + blockSize;
   SCALAR *Hcol = &S->H[(size_t)S->maxBasis * numCols];
   for (int c0 = 0; c0 < blockSize; c0 += 8) {
      int bc = PB_MIN(8, blockSize - c0);
      SCALAR *P = Hcol + (size_t)S->maxBasis * c0;
      CHK(pb200_dortho_sweep(S->dev, S->n, NULL, 0, 0, S->V, m, S->ld,
            S->W + (size_t)S->ld * (numCols + c0), bc, S->ld, NULL, 0, NULL, 0, 0, P,
            S->maxBasis));
      CHK(pb_reduce_panel(S, P, m, bc, S->maxBasis));
   }
   return 0;
}

/* Solve the projected (generalized) eigenproblem and order the pairs by primme.target
 * (reference solve_projection.c:188-331). */
static int solve_H_impl(pb_solver *S, const SCALAR *H, int ldH, int n, const SCALAR *VtBVblk, int ldVtBV,
      SCALAR *hVecs, int ldhVecs, double *hVals, int numConverged, int updateStats) {
   primme_params *primme = S->primme;
   if (n == 0) return 0;
   const double sign = primme->target == primme_largest ? -1.0 : 1.0;
   for (int j = 0; j < n; j++)
      for (int i = 0; i <= j; i++) hVecs[i + (size_t)j * ldhVecs] = sign * H[i + (size_t)j * ldH];

   int info = hl_sygv_upper(n, hVecs, ldhVecs, VtBVblk, ldVtBV, hVals);
   if (info != 0) {
      pb_report(primme, __FILE__, __LINE__, PRIMME_LAPACK_FAILURE, "sygvx/syevx on the projected problem");
      return PRIMME_LAPACK_FAILURE;
   }

   if (primme->target == primme_largest) {
      for (int i = 0; i < n; i++) hVals[i] = -hVals[i];
   } else if (primme->target != primme_smallest) {
      /* interior: order by distance to the first shift not yet satisfied (:232-329) */
      int *permu = (int *)malloc(sizeof(int) * n);
      int i, j, index = 0;
      double shift = primme->targetShifts[PB_MIN(primme->numTargetShifts - 1, numConverged)];
      if (primme->target == primme_closest_geq) {
         for (j = 0; j < n; j++)
            if (hVals[j] >= shift) break;
         for (i = j; i < n; i++) permu[index++] = i;
         for (i = 0; i < j; i++) permu[index++] = i;
      } else if (primme->target == primme_closest_leq) {
         for (j = n - 1; j >= 0; j--)
            if (hVals[j] <= shift) break;
         for (i = j; i >= 0; i--) permu[index++] = i;
         for (i = n - 1; i > j; i--) permu[index++] = i;
      } else if (primme->target == primme_closest_abs) {
         for (j = 0; j < n; j++)
            if (hVals[j] >= shift) break;
         i = j - 1;
         while (i >= 0 && j < n) {
            if (fabs(hVals[i] - shift) < fabs(hVals[j] - shift))
               permu[index++] = i--;
            else
               permu[index++] = j++;
         }
         if (i < 0)
            for (i = j; i < n; i++) permu[index++] = i;
         else if (j >= n)
            for (j = i; j >= 0; j--) permu[index++] = j;
      } else { /* primme_largest_abs */
         j = 0, i = n - 1;
         while (i >= j) {
            if (fabs(hVals[i] - shift) > fabs(hVals[j] - shift))
               permu[index++] = i--;
            else
               permu[index++] = j++;
         }
      }
      hl_permute_reals(hVals, n, permu);
      hl_permute_cols(hVecs, n, n, ldhVecs, permu);
      free(permu);
   }

   if (updateStats) {
      /* spectrum estimates feed the default convergence tolerance (:144-151) */
      for (int i = 0; i < n; i++) {
         primme->stats.estimateMinEVal = PB_MIN(primme->stats.estimateMinEVal, hVals[i]);
         primme->stats.estimateMaxEVal = PB_MAX(primme->stats.estimateMaxEVal, hVals[i]);
         primme->stats.estimateLargestSVal =
               PB_MAX(primme->stats.estimateLargestSVal, fabs(hVals[i]));
      }
   }
   return 0;
}

/* For every column i in [n0,n) of W pick the not-yet-taken column of V with the largest
 * |cos| (reference solve_projection.c:1009-1064). */
int pb_map_vecs(const SCALAR *V, int m, int nV, int ldV, const SCALAR *W, int n0, int n, int ldW,
      int *p) {
   double *Vnorms = (double *)malloc(sizeof(double) * (nV > 0 ? nV : 1));
   SCALAR *ip = (SCALAR *)calloc((size_t)(nV > 0 ? nV : 1) * (n - n0 > 0 ? n - n0 : 1), sizeof(SCALAR));
   for (int i = 0; i < nV; i++)
      Vnorms[i] = sqrt(PB_REAL(hl_dot(m, &V[(size_t)ldV * i], &V[(size_t)ldV * i])));
   hl_gemm('C', 'N', nV, n - n0, m, 1.0, V, ldV, &W[(size_t)ldW * n0], ldW, 0.0, ip, nV > 0 ? nV : 1);
   for (int i = n0; i < n; i++) {
      int jmax = -1;
      double ipmax = -1;
      for (int j = 0; j < nV; j++) {
         double ipij = PB_ABS(ip[(size_t)nV * (i - n0) + j]);
         if (ipij > ipmax * Vnorms[j]) {
            int k;
            for (k = 0; k < i && p[k] != j; k++)
               ;
            if (k < i) continue;
            ipmax = fabs(ipij / Vnorms[j]);
            jmax = j;
         }
      }
      if (jmax < 0) jmax = i;
      p[i] = jmax;
   }
   free(Vnorms), free(ip);
   return 0;
}

/* timed entry points (the reference times neither phase; the PB200_DEBUG report does) */
int pb_update_projection(pb_solver *S, int numCols, int blockSize) {
   const double t0 = hl_wtime();
   int rc = update_projection_impl(S, numCols, blockSize);
   S->tProj += hl_wtime() - t0;
   return rc;
}

int pb_solve_H(pb_solver *S, const SCALAR *H, int ldH, int n, const SCALAR *VtBVblk, int ldVtBV,
      SCALAR *hVecs, int ldhVecs, double *hVals, int numConverged, int updateStats) {
   const double t0 = hl_wtime();
   int rc = solve_H_impl(S, H, ldH, n, VtBVblk, ldVtBV, hVecs, ldhVecs, hVals, numConverged, updateStats);
   S->tSolveH += hl_wtime() - t0;
   return rc;
}


/* ------------------------------------------------------------------------------------------
 * Skew-Q projector with a preconditioner (PRIMME_JDQR + applyPreconditioner).
 * update_XKinvBX (factorize.c:183-235 with B = I): grow M = evecs' evecsHat by blockSize columns
 * (update_projection.c:61-102: the new columns, rows 0 .. numCols + blockSize) and refactorise it
 * (Bunch-Kaufman on the upper triangle, packed with leading dimension nM the way MSolve reads it,
 * factorize.c:268-297).  numCols counts the orthogonality constraints too.
 * NOTE (parity): the reference reaches this code with ldMfact == 0 (main_iter.c:1089) and crashes before it
 * (restart.c:1511); the oracle of tests/test_jdqmr_cpu.py is the reference with those two lines fixed
 * (oracle/Makefile: refskewq).
 * ---------------------------------------------------------------------------------------- */
int pb_update_XKinvBX(pb_solver *S, int numCols, int blockSize) {
   const int ldM = S->maxEvecsSize, nM = numCols + blockSize;
   for (int c0 = 0; c0 < blockSize; c0 += 8) {
      const int nc = PB_MIN(8, blockSize - c0);
      SCALAR *P = &S->Mskew[(size_t)ldM * (numCols + c0)];
      for (int j = 0; j < nc; j++)
         for (int i = 0; i < nM; i++) P[i + (size_t)ldM * j] = 0.0;
      CHK(pb200_dortho_sweep(S->dev, S->n, S->evecs, nM, S->ldevecs, NULL, 0, 0,
            S->evecsHat + (size_t)S->ld * (numCols + c0), nc, S->ld, NULL, 0, NULL, 0, 0, P, ldM));
      CHK(pb_reduce_panel(S, P, nM, nc, ldM));
   }
   if (nM == 0) return 0;
   if (nM == 1) {
      S->Mfact[0] = S->Mskew[0];
      return 0;
   }
   for (int j = 0; j < nM; j++)
      for (int i = 0; i <= j; i++) S->Mfact[i + (size_t)nM * j] = S->Mskew[i + (size_t)ldM * j];
   return hl_hetrf_upper(nM, S->Mfact, nM, S->ipivot) ? PRIMME_LAPACK_FAILURE : 0;
}

/* end of restart_projection (restart.c:1471-1531): K^{-1} of the vectors that joined evecs at this restart
 * (all converged Ritz vectors when there is no locking: they change from one restart to the next), then M */
int pb_skew_evecs_after_restart(pb_solver *S, int numConverged) {
   primme_params *primme = S->primme;
   if (!S->evecsHat) return 0;
   if (!primme->locking) S->numConvergedStored = 0;
   const int evecsSize = S->numConvergedStored, nOC = primme->numOrthoConst;
   const int numRecentlyConverged = numConverged - evecsSize;
   double *shifts = NULL;
   int owned = 0;
   if (numConverged <= primme->numTargetShifts)
      shifts = &primme->targetShifts[evecsSize];
   else if (primme->numTargetShifts > 0) {
      shifts = (double *)malloc(sizeof(double) * PB_MAX(numConverged, 1));
      for (int i = 0; i < numRecentlyConverged; i++)
         shifts[i] = primme->targetShifts[PB_MIN(i + evecsSize, primme->numTargetShifts - 1)];
      owned = 1;
   }
   primme->ShiftsForPreconditioner = shifts;
   int rc = 0;
   if (numRecentlyConverged > 0)
      rc = pb_apply_precond(S, S->evecs + (size_t)S->ldevecs * (evecsSize + nOC), S->ldevecs,
            S->evecsHat + (size_t)S->ld * (evecsSize + nOC), S->ld, numRecentlyConverged);
   primme->ShiftsForPreconditioner = NULL;
   if (owned) free(shifts);
   if (!rc) rc = pb_update_XKinvBX(S, nOC + evecsSize, numRecentlyConverged);
   S->numConvergedStored = numConverged;
   return rc;
}

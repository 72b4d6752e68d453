/* dav_refined.c -- refined and harmonic extraction: the QR factorisation (A - tau I) V = Q R carried next
 * to V and W.
 *
 * Restates, for B = I and fp64:
 *   update_Q              src/eigs/update_W.c:69-113        new columns of Q and R
 *   solve_H_Ref           src/eigs/solve_projection.c:541-628   SVD of R, Rayleigh quotients
 *   prepare_vecs          src/eigs/solve_projection.c:842-985   Rayleigh-Ritz inside clusters of
 *                                                            close singular values
 *   restart_refined       src/eigs/restart.c:1837-2160      Q, R after V <- V*hVecs
 *   solve_H_Harm          src/eigs/solve_projection.c:395-470    harmonic Ritz pairs from (Q'V inv(R), Q'Q)
 *   restart_harmonic      src/eigs/restart.c:2256-2326      QR, Q'V and the projected problem again
 *   update_projection (unsymmetric)  src/eigs/update_projection.c:81-165   new columns and rows of Q'V
 * The n-long work is the same three kernel families as the rest of the solver: the residual
 * utility (Q = W - tau V), the block-ortho sweep (on Q, with its own Gram matrix when orth is
 * explicit) and the VWXR sweep (Q <- Q*hU with Q'Q).  Everything else is maxBasis x maxBasis host
 * algebra with the reference's LAPACK calls (dgesvd 'S','O', dpotrf, dtrmm, dtrsm).
 */
#include "pb_host.h"

#ifdef PB_COMPLEX
/* zprimme covers Rayleigh-Ritz extraction (what the reference's complex test configurations and the
 * benchmark configuration C3 use); refined / harmonic extraction in complex arithmetic is refused by
 * front.c:check_scope, so these entry points are never reached with work to do. */
int pb_update_Q(pb_solver *S, double shift, int basisSize, int blockSize, int *nQ) {
   (void)S, (void)shift, (void)basisSize, (void)blockSize, (void)nQ;
   return PRIMME_FUNCTION_UNAVAILABLE;
}
int pb_solve_H_ref(pb_solver *S, int n, const SCALAR *VtBVblk, int ldVtBV, int numConverged) {
   (void)S, (void)n, (void)VtBVblk, (void)ldVtBV, (void)numConverged;
   return PRIMME_FUNCTION_UNAVAILABLE;
}
int pb_prepare_vecs(pb_solver *S, int basisSize, int i0, int blockSize, int targetShiftIndex, int *arbitraryVecs,
      double smallestResNorm, const int *flags, int RRForAll) {
   (void)S, (void)basisSize, (void)i0, (void)blockSize, (void)targetShiftIndex, (void)arbitraryVecs;
   (void)smallestResNorm, (void)flags, (void)RRForAll;
   return 0; /* Rayleigh-Ritz coefficient vectors need no preparation (solve_projection.c:842-850) */
}
int pb_restart_refined(pb_solver *S, int restartSize, int basisSize, int numConverged, int numPrevRetained,
      int indexOfPreviousVecs, int indexOfPreviousVecsBeforeRestart, const int *restartPerm, const int *hVecsPerm,
      int *targetShiftIndex) {
   (void)S, (void)restartSize, (void)basisSize, (void)numConverged, (void)numPrevRetained, (void)indexOfPreviousVecs;
   (void)indexOfPreviousVecsBeforeRestart, (void)restartPerm, (void)hVecsPerm, (void)targetShiftIndex;
   return PRIMME_FUNCTION_UNAVAILABLE;
}
int pb_update_QtV(pb_solver *S, int numCols, int blockSize) {
   (void)S, (void)numCols, (void)blockSize;
   return 0;
}
int pb_solve_H_harm(pb_solver *S, int n, const SCALAR *VtBVblk, int ldVtBV, int numConverged) {
   (void)S, (void)n, (void)VtBVblk, (void)ldVtBV, (void)numConverged;
   return PRIMME_FUNCTION_UNAVAILABLE;
}
int pb_restart_harmonic(pb_solver *S, int restartSize, int basisSize, int numConverged, int *targetShiftIndex) {
   (void)S, (void)restartSize, (void)basisSize, (void)numConverged, (void)targetShiftIndex;
   return PRIMME_FUNCTION_UNAVAILABLE;
}
#else
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Q(:,basisSize:+blockSize) = W(:,...) - shift*V(:,...), orthonormalised against Q(:,0:*nQ) with the
 * rotations appended to R */
int pb_update_Q(pb_solver *S, double shift, int basisSize, int blockSize, int *nQ) {
   if (blockSize <= 0 || !S->R) return 0;
   const int ldR = S->maxBasis;
   double t[8], nrm[8];
   for (int i = 0; i < 8; i++) t[i] = shift;
   for (int c0 = 0; c0 < blockSize; c0 += 8) {
      const int nc = PB_MIN(8, blockSize - c0);
      double *q = S->Q + (size_t)S->ld * (basisSize + c0);
      CHK(pb200_copy_d2d(S->dev, S->W + (size_t)S->ld * (basisSize + c0), S->ld, q, S->ld, S->n, nc, 8));
      CHK(pb200_dresidual_inplace(S->dev, S->n, t, S->V + (size_t)S->ld * (basisSize + c0), S->ld, q, S->ld, nc, nrm));
   }
   CHK(pb_ortho_block_R(S, S->Q, S->ld, S->QtQ, S->fQtQ, S->maxBasis, S->primme->maxBasisSize, S->R, ldR, *nQ,
         *nQ + blockSize - 1, nQ));
   hl_zero(&S->R[basisSize], blockSize, basisSize, ldR);
   return 0;
}

/* hVecs, hU, hSVals from the SVD of R (in the V'V and Q'Q inner products when they are carried),
 * ordered by the target, and the Rayleigh quotients of the refined vectors */
int pb_solve_H_ref(pb_solver *S, int n, const double *VtBVblk, int ldVtBV, int numConverged) {
   primme_params *primme = S->primme;
   (void)numConverged;
   if (n == 0) return 0;
   const double t0 = hl_wtime();
   const int ld = S->maxBasis;
   double *hVecs = S->hVecs, *hU = S->hU, *hSVals = S->hSVals, *hVals = S->hVals;
   int rc = 0;

   hl_copy(S->R, n, n, ld, hVecs, ld);
   if (S->QtQ) {
      hl_copy(S->QtQ, n, n, ld, hU, ld);
      if (hl_potrf_upper(n, hU, ld) != 0) return PRIMME_LAPACK_FAILURE;
      hl_trmm('L', 'U', 'N', 'N', n, n, 1.0, hU, ld, hVecs, ld);
   }
   double *U = NULL;
   if (VtBVblk) {
      U = (double *)malloc(sizeof(double) * (size_t)n * n);
      if (!U) return PRIMME_MALLOC_FAILURE;
      hl_copy(VtBVblk, n, n, ldVtBV, U, n);
      if (hl_potrf_upper(n, U, n) != 0) {
         free(U);
         return PRIMME_LAPACK_FAILURE;
      }
      hl_trsm('R', 'U', 'N', 'N', n, n, 1.0, U, n, hVecs, ld);
   }
   /* gesvd returns V' and descending singular values */
   if (hl_gesvd_SO(n, n, hVecs, ld, hSVals, hU, ld) != 0) {
      free(U);
      return PRIMME_LAPACK_FAILURE;
   }
   double *rwork = (double *)malloc(sizeof(double) * (size_t)n * n);
   for (int j = 0; j < n; j++)
      for (int i = 0; i < n; i++) rwork[(size_t)n * j + i] = hVecs[(size_t)ld * i + j];
   hl_copy(rwork, n, n, n, hVecs, ld);
   if (U) {
      hl_trsm('L', 'U', 'N', 'N', n, n, 1.0, U, n, hVecs, ld);
      free(U);
   }
   if (primme->target == primme_closest_abs || primme->target == primme_closest_leq ||
         primme->target == primme_closest_geq) {
      int *perm = (int *)malloc(sizeof(int) * n);
      for (int i = 0; i < n; i++) perm[i] = n - 1 - i;
      hl_permute_cols(hSVals, 1, n, 1, perm);
      hl_permute_cols(hVecs, n, n, ld, perm);
      hl_permute_cols(hU, n, n, ld, perm);
      free(perm);
   }
   hl_zero(rwork, n, n, n);
   hl_symm_lu(n, n, 1.0, S->H, ld, hVecs, ld, 0.0, rwork, n);
   for (int i = 0; i < n; i++) hVals[i] = hl_dot(n, &hVecs[(size_t)ld * i], &rwork[(size_t)n * i]);
   free(rwork);
   for (int i = 0; i < n; i++) {
      primme->stats.estimateMinEVal = PB_MIN(primme->stats.estimateMinEVal, hVals[i]);
      primme->stats.estimateMaxEVal = PB_MAX(primme->stats.estimateMaxEVal, hVals[i]);
      primme->stats.estimateLargestSVal = PB_MAX(primme->stats.estimateLargestSVal, fabs(hVals[i]));
   }
   S->tSolveH += hl_wtime() - t0;
   return rc;
}

/* Refined vectors whose singular values are too close to tell apart are not good coefficient
 * vectors: replace every such cluster by its Rayleigh-Ritz vectors, until blockSize candidates
 * after position i0 are well defined.  *arbitraryVecs = number of leading columns of hVecs that are
 * no longer singular vectors of R; hVecsRot holds the rotations (hVecs = hV * hVecsRot). */
int pb_prepare_vecs(pb_solver *S, int basisSize, int i0, int blockSize, int targetShiftIndex, int *arbitraryVecs,
      double smallestResNorm, const int *flags, int RRForAll) {
   primme_params *primme = S->primme;
   if (!S->refined || basisSize == 0 || blockSize == 0) return 0;
   const int ld = S->maxBasis;
   double *hVecs = S->hVecs, *hVals = S->hVals, *hSVals = S->hSVals, *hVecsRot = S->hVecsRot;
   const double aNorm = primme->aNorm <= 0.0 ? primme->stats.estimateLargestSVal : primme->aNorm;
   double eps = primme->stats.maxConvTol > 0.0 ? primme->stats.maxConvTol
                                                : (smallestResNorm < HUGE_VAL ? smallestResNorm / 10.0 : 0.0);
   eps = PB_MAX(6.28 * PB_EPS, eps);
   int i, j, k, candidates, someCandidate;

   for (candidates = 0, i = PB_MIN(*arbitraryVecs, basisSize), j = i0; j < basisSize && candidates < blockSize;) {
      double ip;
      for (; j < i; j++)
         if (!flags || flags[j] == UNCONVERGED) candidates++;
      if (candidates >= blockSize) break;

      /* first i > j whose singular value is separated enough from its predecessor's */
      for (i = j + 1, someCandidate = 0, ip = 0.0; i < basisSize; i++) {
         double minDiff = sqrt(2.0) * hSVals[basisSize - 1] * PB_EPS / (aNorm * eps / fabs(hVals[i] - hVals[i - 1]));
         double ip0 = fabs(hVecs[(size_t)(i - 1) * ld + basisSize - 1]);
         double ip1 = ((ip += ip0 * ip0) != 0.0) ? ip : HUGE_VAL;
         someCandidate = 1;
         if (fabs(hSVals[i] - hSVals[i - 1]) >= minDiff &&
               (smallestResNorm >= HUGE_VAL || sqrt(ip1) >= smallestResNorm / aNorm / 3.16))
            break;
      }
      i = PB_MIN(i, basisSize);

      if (i - j > 1 && (someCandidate || RRForAll)) {
         const int an = i - j;
         double *aH = (double *)calloc((size_t)basisSize * an, sizeof(double));
         double *ahVecs = &hVecsRot[(size_t)ld * j + j];
         if (!aH) return PRIMME_MALLOC_FAILURE;
         hl_zero(&hVecsRot[(size_t)ld * *arbitraryVecs], primme->maxBasisSize, i - *arbitraryVecs, ld);
         for (k = *arbitraryVecs; k < i; k++) hVecsRot[(size_t)ld * k + k] = 1.0;
         /* aH = hVecs(:,j:i)' H hVecs(:,j:i), its eigenpairs ordered by the target */
         pb_compute_submatrix(&hVecs[(size_t)ld * j], an, ld, S->H, basisSize, ld, aH, an);
         int rc = pb_solve_H(S, aH, an, an, NULL, 0, ahVecs, ld, &hVals[j], targetShiftIndex, 0);
         if (rc) {
            free(aH);
            return rc;
         }
         hl_zero(aH, basisSize, an, basisSize);
         hl_gemm('N', 'N', basisSize, an, an, 1.0, &hVecs[(size_t)ld * j], ld, ahVecs, ld, 0.0, aH, basisSize);
         hl_copy(aH, basisSize, an, basisSize, &hVecs[(size_t)ld * j], ld);
         free(aH);
         *arbitraryVecs = i;
      }
   }
   return 0;
}

/* Q <- Q*hU(:,0:restartSize) in place with Q'Q refreshed (Num_update_VWXR on Q, restart.c:2062-2079).
 * The sweep is the restart instance of the VWXR kernel with V = W = Q. */
static int restart_Q(pb_solver *S, int basisSize, int restartSize) {
   pb200_vwxr_out o;
   memset(&o, 0, sizeof(o));
   const int ld = S->maxBasis;
   double *scratch = NULL;
   o.X[0].ptr = S->Q, o.X[0].ld = S->ld, o.X[0].cb = 0, o.X[0].ce = restartSize;
   o.Wo.ptr = S->Q, o.Wo.ld = S->ld, o.Wo.cb = 0, o.Wo.ce = restartSize; /* same values, same place */
   if (S->QtQ) {
      scratch = (double *)malloc(sizeof(double) * (size_t)ld * ld);
      if (!scratch) return PRIMME_MALLOC_FAILURE;
      o.nG = restartSize, o.G_host = S->QtQ, o.ldG = ld;
      o.nH = restartSize, o.H_host = scratch, o.ldH = ld;
   }
   int rc = pb200_dvwxr(S->dev, S->n, S->Q, S->Q, basisSize, S->ld, S->hU, ld, restartSize, S->hVals, &o);
   if (!rc && S->QtQ && S->primme->numProcs > 1 && pb200_ctx_nranks(S->dev) <= 1)
      rc = pb_reduce_panel(S, S->QtQ, restartSize, restartSize, ld);
   free(scratch);
   if (!rc && S->QtQ) rc = pb_update_cholesky_gram(S->QtQ, S->fQtQ, ld, 0, restartSize);
   return rc;
}

int pb_restart_refined(pb_solver *S, int restartSize, int basisSize, int numConverged, int numPrevRetained,
      int indexOfPreviousVecs, int indexOfPreviousVecsBeforeRestart, const int *restartPerm, const int *hVecsPerm,
      int *targetShiftIndex) {
   primme_params *primme = S->primme;
   const int ld = S->maxBasis, ldG = S->maxRank;
   double *H = S->H, *hVecs = S->hVecs, *hVals = S->hVals, *hSVals = S->hSVals, *hU = S->hU, *R = S->R;
   double *hVecsRot = S->hVecsRot;
   int *numArbitraryVecs = &S->numArbitraryVecs;
   const double aNorm = PB_MAX(primme->aNorm, primme->stats.estimateLargestSVal);
   int i, j;

   if (primme->orth == primme_orth_implicit_I) pb_compute_submatrix(hVecs, restartSize, ld, H, basisSize, ld, H, ld);

   const int nLocked = primme->numOrthoConst + (primme->locking ? numConverged : 0);
   const double *VtBVblk = S->VtBV ? &S->VtBV[(size_t)ldG * nLocked + nLocked] : NULL;

   /* the target moved: factorise (A - tau I) V again from scratch (:1878-1899) */
   if (*targetShiftIndex < 0 ||
         fabs(primme->targetShifts[*targetShiftIndex] -
               primme->targetShifts[PB_MIN(primme->numTargetShifts - 1, numConverged)]) > PB_EPS * aNorm) {
      *targetShiftIndex = PB_MIN(primme->numTargetShifts - 1, numConverged);
      int nQ = 0;
      CHK(pb_update_Q(S, primme->targetShifts[*targetShiftIndex], 0, restartSize, &nQ));
      if (restartSize != nQ) return PRIMME_UNEXPECTED_FAILURE;
      CHK(pb_solve_H_ref(S, restartSize, VtBVblk, ldG, numConverged));
      *numArbitraryVecs = 0;
      return 0;
   }

   int *restartPerm0 = (int *)malloc(sizeof(int) * PB_MAX(restartSize, 1));
   for (i = 0; i < restartSize; i++) restartPerm0[i] = restartPerm[hVecsPerm[i]];
   int newNumArbitraryVecs = 0;
   for (i = 0; i < restartSize - numPrevRetained; i++)
      if (restartPerm0[i] < *numArbitraryVecs) newNumArbitraryVecs++;

   /* R*Y = [hU diag(hSVals) hVecsRot, R prevhVecs]: the first part without forming R*Y (:1929-1955) */
   double *RPrev = (double *)calloc((size_t)PB_MAX(numPrevRetained, 1) * basisSize, sizeof(double));
   hl_gemm('N', 'N', basisSize, numPrevRetained, basisSize, 1.0, R, ld, &hVecs[(size_t)ld * indexOfPreviousVecs], ld, 0.0,
         RPrev, basisSize);

   const int nRegular = restartSize - numPrevRetained;
   int mh = *numArbitraryVecs;
   for (i = 0; i < nRegular; i++) mh = PB_MAX(mh, restartPerm0[i] + 1);
   double *rot0 = (double *)calloc((size_t)PB_MAX(mh, 1) * PB_MAX(nRegular, 1), sizeof(double));
   for (i = 0; i < newNumArbitraryVecs; i++)
      memcpy(&rot0[(size_t)mh * i], &hVecsRot[(size_t)ld * restartPerm0[i]], sizeof(double) * *numArbitraryVecs);
   for (i = 0; i < newNumArbitraryVecs; i++)
      for (j = 0; j < *numArbitraryVecs; j++) rot0[(size_t)mh * i + j] *= hSVals[j];
   for (i = newNumArbitraryVecs; i < nRegular; i++) rot0[(size_t)mh * i + restartPerm0[i]] = hSVals[restartPerm0[i]];

   hl_zero(R, primme->maxBasisSize, primme->maxBasisSize, ld);
   /* [rot0, R] = ortho(rot0) in the Q'Q inner product, column by column */
   long long seed[4];
   int rc = 0;
   for (i = 0; i < 4; i++) seed[i] = primme->iseed[i];
   rc = pb_ortho_local_R(rot0, mh, R, ld, 0, nRegular - 1, mh, S->QtQ, ld, seed);
   if (!rc) {
      /* hU = hU * rot0 (then in the coordinates of the Cholesky factor of Q'Q), next to R*prevhVecs */
      double *rw = (double *)calloc((size_t)basisSize * PB_MAX(nRegular, 1), sizeof(double));
      hl_gemm('N', 'N', basisSize, nRegular, mh, 1.0, hU, ld, rot0, PB_MAX(mh, 1), 0.0, rw, basisSize);
      hl_copy(rw, basisSize, nRegular, basisSize, hU, ld);
      free(rw);
      if (S->QtQ) hl_trsm('R', 'U', 'N', 'N', basisSize, nRegular, 1.0, S->fQtQ, ld, hU, ld);
      hl_copy(RPrev, basisSize, numPrevRetained, basisSize, &hU[(size_t)ld * nRegular], ld);
      rc = pb_ortho_local_R(hU, ld, R, ld, nRegular, nRegular + numPrevRetained - 1, basisSize, S->QtQ, ld, seed);
   }
   for (i = 0; i < 4; i++) primme->iseed[i] = seed[i];
   free(rot0), free(RPrev);
   if (rc) {
      free(restartPerm0);
      return PRIMME_UNEXPECTED_FAILURE;
   }

   /* columns that were plain singular vectors keep a diagonal R (:2018-2030) */
   for (i = newNumArbitraryVecs; i < nRegular; i++)
      if (restartPerm0[i] >= *numArbitraryVecs) {
         for (j = 0; j <= i; j++) R[(size_t)ld * i + j] = 0.0;
         R[(size_t)ld * i + i] = hSVals[restartPerm0[i]];
      }
   free(restartPerm0);
   if (*numArbitraryVecs <= indexOfPreviousVecsBeforeRestart) hl_zero(&R[(size_t)ld * nRegular], nRegular, numPrevRetained, ld);

   CHK(restart_Q(S, basisSize, restartSize));

   /* R lost its triangular shape: recompute hVecs, hU, hSVals; hVals only follow the permutation */
   double *keep = (double *)malloc(sizeof(double) * PB_MAX(restartSize, 1));
   memcpy(keep, hVals, sizeof(double) * restartSize);
   rc = pb_solve_H_ref(S, restartSize, VtBVblk, ldG, numConverged);
   memcpy(hVals, keep, sizeof(double) * restartSize);
   free(keep);
   if (rc) return rc;
   hl_permute_cols(hVals, 1, restartSize, 1, hVecsPerm);

   int *inv = (int *)malloc(sizeof(int) * PB_MAX(restartSize, 1));
   for (i = 0; i < restartSize; i++) inv[hVecsPerm[i]] = i;
   hl_permute_cols(R, restartSize, restartSize, ld, inv);
   free(inv);

   if (*numArbitraryVecs <= indexOfPreviousVecsBeforeRestart) {
      for (i = *numArbitraryVecs = newNumArbitraryVecs; i < restartSize; i++)
         if (hVecsPerm[i] != i) *numArbitraryVecs = i + 1;
   } else
      *numArbitraryVecs = restartSize;

   /* hVecsRot = hVecs' for the arbitrary vectors, whose coefficient vectors are canonical (:2132-2156) */
   hl_zero(hVecsRot, primme->maxBasisSize, primme->maxBasisSize, ld);
   for (j = 0; j < *numArbitraryVecs; j++)
      for (i = 0; i < restartSize; i++) hVecsRot[(size_t)ld * j + i] = hVecs[(size_t)ld * i + j];
   hl_zero(hVecs, restartSize, *numArbitraryVecs, ld);
   for (j = 0; j < *numArbitraryVecs; j++) hVecs[(size_t)ld * j + hVecsPerm[j]] = 1.0;
   (void)indexOfPreviousVecs;
   return 0;
}

/* ------------------------------------------------------------------------------------------
 * Harmonic extraction
 * ---------------------------------------------------------------------------------------- */
/* QtV(0:m, numCols:m) = Q(:,0:m)' V(:,numCols:m) and QtV(numCols:m, 0:numCols) = Q(:,numCols:m)' V(:,0:numCols),
 * m = numCols + blockSize: two row sweeps of the projection kernel */
int pb_update_QtV(pb_solver *S, int numCols, int blockSize) {
   if (blockSize <= 0 || !S->QtV) return 0;
   const int ld = S->maxBasis, m = numCols + blockSize;
   for (int c0 = 0; c0 < blockSize; c0 += 8) {
      const int bc = PB_MIN(8, blockSize - c0);
      double *P = &S->QtV[(size_t)ld * (numCols + c0)];
      CHK(pb200_dortho_sweep(S->dev, S->n, NULL, 0, 0, S->Q, m, S->ld, S->V + (size_t)S->ld * (numCols + c0), bc, S->ld,
            NULL, 0, NULL, 0, 0, P, ld));
      CHK(pb_reduce_panel(S, P, m, bc, ld));
   }
   if (numCols > 0) {
      /* the new rows are the transposed panel V(:,0:numCols)' Q(:,numCols:m) */
      double *P = (double *)malloc(sizeof(double) * (size_t)numCols * 8);
      if (!P) return PRIMME_MALLOC_FAILURE;
      for (int c0 = 0; c0 < blockSize; c0 += 8) {
         const int bc = PB_MIN(8, blockSize - c0);
         CHK(pb200_dortho_sweep(S->dev, S->n, NULL, 0, 0, S->V, numCols, S->ld, S->Q + (size_t)S->ld * (numCols + c0), bc,
               S->ld, NULL, 0, NULL, 0, 0, P, numCols));
         CHK(pb_reduce_panel(S, P, numCols, bc, numCols));
         for (int j = 0; j < bc; j++)
            for (int i = 0; i < numCols; i++) S->QtV[(size_t)ld * i + numCols + c0 + j] = P[(size_t)numCols * j + i];
      }
      free(P);
   }
   return 0;
}

/* eigenpairs of (Q'V inv(R), Q'Q) ordered as the harmonic values 1/(theta - tau), mapped back with
 * hVecs = inv(R) g, orthonormalised in the V'V inner product; Rayleigh quotients as values */
int pb_solve_H_harm(pb_solver *S, int n, const double *VtBVblk, int ldVtBV, int numConverged) {
   primme_params *primme = S->primme;
   (void)numConverged;
   if (n == 0) return 0;
   const double t0 = hl_wtime();
   const int ld = S->maxBasis;
   double *hVecs = S->hVecs, *hVals = S->hVals;
   double *fR = (double *)malloc(sizeof(double) * (size_t)n * n);
   int *piv = (int *)malloc(sizeof(int) * n);
   if (!fR || !piv) return PRIMME_MALLOC_FAILURE;
   int rc = 0;
   hl_copy(S->R, n, n, ld, fR, n);
   if (hl_getrf(n, n, fR, n, piv) != 0) rc = PRIMME_LAPACK_FAILURE;
   if (!rc) {
      for (int j = 0; j < n; j++)
         for (int i = 0; i < n; i++) hVecs[(size_t)ld * j + i] = S->QtV[(size_t)ld * i + j];
      if (hl_getrs('C', n, n, fR, n, piv, hVecs, ld) != 0) rc = PRIMME_LAPACK_FAILURE;
   }
   if (!rc) {
      double zero = 0.0, *oldShifts = primme->targetShifts;
      const primme_target oldTarget = primme->target;
      primme->targetShifts = &zero;
      primme->target = oldTarget == primme_closest_geq   ? primme_largest
                       : oldTarget == primme_closest_leq ? primme_smallest
                                                         : primme_largest_abs;
      rc = pb_solve_H(S, hVecs, ld, n, S->QtQ, ld, hVecs, ld, hVals, 0, 0);
      primme->targetShifts = oldShifts;
      primme->target = oldTarget;
   }
   if (!rc) {
      hl_copy(hVecs, n, n, ld, S->hU, ld);
      if (hl_getrs('N', n, n, fR, n, piv, hVecs, ld) != 0) rc = PRIMME_LAPACK_FAILURE;
   }
   if (!rc) {
      long long seed[4];
      for (int i = 0; i < 4; i++) seed[i] = primme->iseed[i];
      rc = pb_ortho_local_R(hVecs, ld, NULL, 0, 0, n - 1, n, VtBVblk, ldVtBV, seed) ? PRIMME_UNEXPECTED_FAILURE : 0;
      for (int i = 0; i < 4; i++) primme->iseed[i] = seed[i];
   }
   if (!rc) {
      double *rw = (double *)calloc((size_t)n * n, sizeof(double));
      hl_symm_lu(n, n, 1.0, S->H, ld, hVecs, ld, 0.0, rw, n);
      for (int i = 0; i < n; i++) hVals[i] = hl_dot(n, &hVecs[(size_t)ld * i], &rw[(size_t)n * i]);
      free(rw);
      for (int i = 0; i < n; i++) {
         primme->stats.estimateMinEVal = PB_MIN(primme->stats.estimateMinEVal, hVals[i]);
         primme->stats.estimateMaxEVal = PB_MAX(primme->stats.estimateMaxEVal, hVals[i]);
         primme->stats.estimateLargestSVal = PB_MAX(primme->stats.estimateLargestSVal, fabs(hVals[i]));
      }
   }
   free(fR), free(piv);
   S->tSolveH += hl_wtime() - t0;
   return rc;
}

/* after V <- V*hVecs, W <- W*hVecs: everything that depends on Q from scratch */
int pb_restart_harmonic(pb_solver *S, int restartSize, int basisSize, int numConverged, int *targetShiftIndex) {
   primme_params *primme = S->primme;
   const int ld = S->maxBasis, ldG = S->maxRank;
   if (primme->orth == primme_orth_implicit_I) pb_compute_submatrix(S->hVecs, restartSize, ld, S->H, basisSize, ld, S->H, ld);
   *targetShiftIndex = PB_MIN(primme->numTargetShifts - 1, numConverged);
   int nQ = 0;
   CHK(pb_update_Q(S, primme->targetShifts[*targetShiftIndex], 0, restartSize, &nQ));
   if (restartSize != nQ) return PRIMME_UNEXPECTED_FAILURE;
   CHK(pb_update_QtV(S, 0, restartSize));
   const int nLocked = primme->numOrthoConst + (primme->locking ? numConverged : 0);
   CHK(pb_solve_H_harm(S, restartSize, S->VtBV ? &S->VtBV[(size_t)ldG * nLocked + nLocked] : NULL, ldG, numConverged));
   S->numArbitraryVecs = 0;
   return 0;
}
#endif /* !PB_COMPLEX */

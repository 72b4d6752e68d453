/* dav_ortho.c -- block orthogonalisation of new basis columns.
 *
 * Restates reference src/eigs/ortho.c:
 *   Bortho_block_gen (:497-803)  iterated CholQR / SVQB driven by the maintained Gram matrix
 *                                VtBV and its Cholesky factor; the n-long work of every sweep
 *                                (Num_ortho_kernel :963-1072) is ONE fused device launch
 *                                pb200_dortho_sweep, the b x b algebra stays on the host.
 *   Bortho_gen (:124-371)        classical Gram-Schmidt with reorthogonalisation (Daniel test)
 *                                used when no Gram matrix is carried (orth implicit); each pass
 *                                is one fused sweep: update with the previous overlaps + the
 *                                overlaps/norm of the updated vector.
 *   ortho_single_iteration (:826-934), update_cholesky (:1200-1220), rank_estimation (:1165),
 *   decomposition (:1097), Bortho_local (:395-413).
 * B = I only (no mass matrix).
 */
#include "pb_host.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Number of leading columns n0..i-1 judged linearly independent from the Gram matrix G:
 * positive diagonal and |cos| to every earlier column below 0.8/maxRank (:1165-1181). */
static int rank_estimation(const SCALAR *G, int n0, int n1, int maxRank, int ldG) {
   int i, j;
   for (i = n0; i < n1; i++) {
      double gii = PB_REAL(G[(size_t)i * ldG + i]);
      if (!isfinite(gii) || gii <= 0.0) break;
      for (j = 0; j < i; j++)
         if (PB_ABS(G[(size_t)i * ldG + j]) > .8 / maxRank * sqrt(gii * PB_REAL(G[(size_t)j * ldG + j]))) break;
      if (j < i) break;
   }
   return i;
}

/* Cholesky if possible, else eigen-decomposition with descending eigenvalues (:1097-1137). */
static int decomposition(const SCALAR *C, int n, int ldC, SCALAR *Y, int ldY, double *evals,
      int *Yortho) {
   hl_copy(C, n, n, ldC, Y, ldY);
   if (hl_potrf_upper(n, Y, ldY) == 0) {
      *Yortho = 0;
      for (int i = 0; i < n; i++) evals[i] = 1.0;
      return 0;
   }
   for (int i = 0; i < n; i++)
      for (int j = 0; j <= i; j++) Y[(size_t)ldY * i + j] = -C[(size_t)ldC * i + j];
   if (hl_sygv_upper(n, Y, ldY, NULL, 0, evals) != 0) return PRIMME_LAPACK_FAILURE;
   for (int i = 0; i < n; i++) evals[i] = -evals[i];
   *Yortho = 1;
   return 0;
}

/* Append columns n0..n-1 to the Cholesky factor of VtBV (:1200-1220). */
int pb_update_cholesky_gram(const SCALAR *G, SCALAR *fG, int ld, int n0, int n) {
   if (!fG || n <= n0) return 0;
   SCALAR *A = (SCALAR *)malloc(sizeof(SCALAR) * (size_t)n * (n - n0));
   if (!A) return PRIMME_MALLOC_FAILURE;
   hl_copy(&G[(size_t)ld * n0], n, n - n0, ld, A, n);
   hl_trsm('L', 'U', 'C', 'N', n0, n - n0, 1.0, fG, ld, A, n);
   hl_gemm('C', 'N', n - n0, n - n0, n0, -1.0, A, n, A, n, 1.0, &A[n0], n);
   hl_potrf_upper(n - n0, &A[n0], n); /* failure is tolerated exactly as in the reference */
   hl_copy(A, n, n - n0, n, &fG[(size_t)ld * n0], ld);
   free(A);
   return 0;
}

int pb_update_cholesky(pb_solver *S, int n0, int n) {
   return pb_update_cholesky_gram(S->VtBV, S->fVtBV, S->maxRank, n0, n);
}

/* ------------------------------------------------------------------------------------------
 * CGS with reorthogonalisation, device version (no Gram matrix carried).
 * ---------------------------------------------------------------------------------------- */
static int ortho_cgs(pb_solver *S, SCALAR *V, int64_t ldV, SCALAR *R, int ldR, int b1, int b2, const SCALAR *locked,
      int64_t ldLocked, int numLocked, SCALAR *RLocked, int ldRLocked, int *b2_out) {
   primme_params *primme = S->primme;
   const int maxNumOrthos = 3, maxNumRandoms = 10;
   const double tol = sqrt(2.0) / 2.0; /* Daniel et al. test */
   const double eps_orth = PB_EPS;
   const double t0 = hl_wtime();
   int rc = 0;

   if (R) hl_zero(&R[(size_t)ldR * b1], b2 + 1, b2 - b1 + 1, ldR); /* (:146-149) */
   if (RLocked) hl_zero(RLocked, numLocked, b2 - b1 + 1, ldRLocked);
   SCALAR *panel = (SCALAR *)calloc((size_t)(numLocked + b2 + 2), sizeof(SCALAR));
   SCALAR *coef = (SCALAR *)calloc((size_t)(numLocked + b2 + 2), sizeof(SCALAR));
   if (!panel || !coef) return PRIMME_MALLOC_FAILURE;
   *b2_out = b1;

   for (int i = b1; i <= b2; i++) {
      SCALAR *v = V + (size_t)ldV * i;
      const int k = numLocked + i;
      int nOrth = 0, randomizations = 0, updateR = (R || RLocked) ? 1 : 0, have_panel = 0;
      double s0 = 0.0, s02 = 0.0, s1 = 0.0, s12 = 0.0;
      for (;;) {
         if (nOrth >= maxNumOrthos) {
            /* the column is replaced: it no longer factors the input (:183-188) */
            if (updateR && R) R[(size_t)ldR * i + i] = 0.0;
            updateR = 0;
            if (randomizations >= maxNumRandoms) goto done;
            rc = pb_fill_random(S, v, ldV, 1);
            if (rc) goto done;
            randomizations++;
            nOrth = 0;
            have_panel = 0;
         }
         nOrth++;
         if (!have_panel) {
            /* overlaps [locked V(:,0:i)]' v and v'v in one sweep (:230-250) */
            rc = pb200_dortho_sweep(S->dev, S->n, locked, numLocked, ldLocked, V, i, ldV, v, 1,
                  ldV, NULL, 0, NULL, 0, 1, panel, k + 1);
            if (rc) goto done;
            rc = pb_reduce_panel(S, panel, k + 1, 1, k + 1);
            if (rc) goto done;
            primme->stats.numOrthoInnerProds += k + 1;
         }
         if (nOrth == 1) s02 = PB_REAL(panel[k]);
         if (updateR && R)
            for (int j = 0; j < i; j++) R[(size_t)ldR * i + j] += panel[numLocked + j];
         if (updateR && RLocked)
            for (int j = 0; j < numLocked; j++) RLocked[(size_t)ldRLocked * (i - b1) + j] += panel[j];
         memcpy(coef, panel, sizeof(SCALAR) * k);
         /* v -= [locked V] * overlaps, then the overlaps and norm of the new v (:262-291) */
         rc = pb200_dortho_sweep(S->dev, S->n, locked, numLocked, ldLocked, V, i, ldV, v, 1, ldV,
               coef, k, NULL, 0, 1, panel, k + 1);
         if (rc) goto done;
         rc = pb_reduce_panel(S, panel, k + 1, 1, k + 1);
         if (rc) goto done;
         have_panel = 1;
         primme->stats.numOrthoInnerProds += 2 * k + 1;
         if (nOrth == 1) s0 = sqrt(s02);
         s12 = PB_REAL(panel[k]);
         s1 = sqrt(s12);

         if (!isfinite(s0) || !isfinite(s1) || s1 <= eps_orth * s0) {
            nOrth = maxNumOrthos; /* lost all significant digits: randomise */
         } else if (s1 <= tol * s0) {
            s0 = s1, s02 = s12; /* reorthogonalise */
         } else {
            if (updateR && R) R[(size_t)ldR * i + i] = s1;
            double inv = 1.0 / s1;
            if (isfinite(inv)) {
               SCALAR invs = inv;
               rc = pb200_dscale_columns(S->dev, S->n, &invs, v, ldV, 1);
               if (rc) goto done;
               break;
            }
            nOrth = maxNumOrthos;
         }
      }
      *b2_out = i + 1;
   }
done:
   primme->stats.timeOrtho += hl_wtime() - t0;
   free(panel), free(coef);
   return rc;
}

/* ------------------------------------------------------------------------------------------
 * Block orthogonalisation of V(:,b1:b2) (b2 inclusive) against [locked V(:,0:b1)] and itself.
 * With S->VtBV: iterated CholQR/SVQB (:497-803); columns b1..b2 of the Gram matrix
 * [locked V]'[locked V] and of its Cholesky factor are updated.  Without: CGS.
 * ---------------------------------------------------------------------------------------- */
static int ortho_block_gram(pb_solver *S, SCALAR *G, SCALAR *fG, int ldG, int maxRank, SCALAR *V, int64_t ldV,
      SCALAR *R, int ldR, int b1, int b2, const SCALAR *locked, int64_t ldLocked, int numLocked, SCALAR *RLocked,
      int ldRLocked, int *b2_out, const SCALAR *P0, int ldP0);

int pb_ortho_block(pb_solver *S, SCALAR *V, int64_t ldV, int b1, int b2, const SCALAR *locked,
      int64_t ldLocked, int numLocked, SCALAR *RLocked, int ldRLocked, int *b2_out) {
   return pb_ortho_block_p0(S, V, ldV, b1, b2, locked, ldLocked, numLocked, RLocked, ldRLocked, b2_out, NULL, 0);
}

int pb_ortho_block_p0(pb_solver *S, SCALAR *V, int64_t ldV, int b1, int b2, const SCALAR *locked,
      int64_t ldLocked, int numLocked, SCALAR *RLocked, int ldRLocked, int *b2_out, const SCALAR *P0,
      int ldP0) {
   return ortho_block_gram(S, S->VtBV, S->fVtBV, S->maxRank, S->maxRank, V, ldV, NULL, 0, b1, b2, locked, ldLocked,
         numLocked, RLocked, ldRLocked, b2_out, P0, ldP0);
}

/* ortho_block (ortho.c:477-493) of a second set of vectors with its own Gram matrix (or none: CGS), returning
 * the factor R with input = output * R: the QR factorisation of (A - tau I) V kept by the refined extraction */
int pb_ortho_block_R(pb_solver *S, SCALAR *Q, int64_t ldQ, SCALAR *QtQ, SCALAR *fQtQ, int ldQtQ, int maxRank,
      SCALAR *R, int ldR, int b1, int b2, int *b2_out) {
   return ortho_block_gram(S, QtQ, fQtQ, ldQtQ, maxRank, Q, ldQ, R, ldR, b1, b2, NULL, 0, 0, NULL, 0, b2_out, NULL, 0);
}

static int ortho_block_gram(pb_solver *S, SCALAR *G, SCALAR *fG, int ldG, int maxRank, SCALAR *V, int64_t ldV,
      SCALAR *R, int ldR, int b1, int b2, const SCALAR *locked, int64_t ldLocked, int numLocked, SCALAR *RLocked,
      int ldRLocked, int *b2_out, const SCALAR *P0, int ldP0) {
   primme_params *primme = S->primme;
   b2++; /* C range convention from here on */
   if (b2 <= b1) {
      *b2_out = b2;
      return 0;
   }
   if (!G)
      return ortho_cgs(S, V, ldV, R, ldR, b1, b2 - 1, locked, ldLocked, numLocked, RLocked, ldRLocked, b2_out);

   /* the device sweep handles at most 8 columns at a time: larger blocks (initial guesses,
    * re-orthogonalisation of the whole basis) go chunk by chunk */
   if (b2 - b1 > 8) {
      int cur = b1;
      while (cur < b2) {
         int hi = PB_MIN(cur + 8, b2), out = 0;
         CHK(ortho_block_gram(S, G, fG, ldG, maxRank, V, ldV, R, ldR, cur, hi - 1, locked, ldLocked, numLocked,
               RLocked ? RLocked + (size_t)ldRLocked * (cur - b1) : NULL, ldRLocked, &out, NULL, 0));
         /* chunk by chunk the factor is block upper triangular with 8-column blocks */
         if (R) hl_zero(&R[(size_t)ldR * cur + hi], b2 - hi, hi - cur, ldR);
         if (out < hi) {
            *b2_out = out;
            return 0;
         }
         cur = hi;
      }
      *b2_out = b2;
      return 0;
   }

   const double eps_orth = PB_EPS;
   const double t0 = hl_wtime();
   const int nb = b2 - b1, nVL = b1 + numLocked;
   SCALAR *A = &G[(size_t)ldG * nVL]; /* new columns of the Gram matrix */
   int rc = 0;

   double *D = (double *)malloc(sizeof(double) * nb), *N = (double *)malloc(sizeof(double) * nb);
   SCALAR *GdA = (SCALAR *)malloc(sizeof(SCALAR) * (size_t)(nVL > 0 ? nVL : 1) * nb);
   SCALAR *Y = (SCALAR *)malloc(sizeof(SCALAR) * nb * nb), *Yapply = (SCALAR *)malloc(sizeof(SCALAR) * nb * nb);
   SCALAR *C = (SCALAR *)malloc(sizeof(SCALAR) * nb * nb), *r = NULL;
   if (!D || !N || !GdA || !Y || !C || !Yapply) return PRIMME_MALLOC_FAILURE;
   SCALAR *r_own = NULL;
   int ldr = nb;
   if (R) {
      /* the diagonal block of R accumulates the rotations of the block itself (:556-580) */
      hl_zero(&R[(size_t)ldR * b1], b1, nb, ldR);
      r = &R[(size_t)ldR * b1 + b1], ldr = ldR;
   }
   if (RLocked) {
      hl_zero(RLocked, numLocked, nb, ldRLocked);
      if (!r) r = r_own = (SCALAR *)malloc(sizeof(SCALAR) * (size_t)nb * nb);
   }
   if (r) {
      hl_zero(r, nb, nb, ldr);
      for (int i = 0; i < nb; i++) r[(size_t)ldr * i + i] = 1.0;
   }

   *b2_out = b2;
   const int maxits = 5;
   int plus1 = 5, Yortho = 1;
   for (int its = 0; its < maxits; its++) {
      /* one row sweep: X <- (X - [locked V] GdA) Yapply (not in the first pass), then the
       * Gram panel A = [locked V(:,0:b2)]' X  (:631-636, kernel :963-1072) */
      const SCALAR *Cc = NULL, *Yc = NULL;
      if (its > 0) {
         Cc = GdA;
         if (Yortho) {
            /* X*Y*diag(1/D) (:1008-1012,1030-1034) */
            for (int j = 0; j < nb; j++)
               for (int i = 0; i < nb; i++) Yapply[(size_t)nb * j + i] = Y[(size_t)nb * j + i] * (1.0 / D[j]);
         } else if (nb == 1) {
            Yapply[0] = 1.0 / Y[0]; /* (:976-981) */
         } else {
            /* X*inv(U), U upper triangular Cholesky factor (:1035-1038) */
            hl_zero(Yapply, nb, nb, nb);
            for (int i = 0; i < nb; i++) Yapply[(size_t)nb * i + i] = 1.0;
            hl_trsm('R', 'U', 'N', 'N', nb, nb, 1.0, Y, nb, Yapply, nb);
         }
         Yc = Yapply;
      }
      if (its == 0 && P0 && numLocked == 0) {
         /* the sweep that produced the block already reduced this panel (candidates sweep) */
         hl_copy(P0, nVL + nb, nb, ldP0, A, ldG);
      } else {
         rc = pb200_dortho_sweep(S->dev, S->n, locked, numLocked, ldLocked, V, b1, ldV,
               V + (size_t)ldV * b1, nb, ldV, Cc, nVL > 0 ? nVL : 1, Yc, nb, 1, A, ldG);
         if (rc) goto done;
         rc = pb_reduce_panel(S, A, nVL + nb, nb, ldG);
         if (rc) goto done;
      }
      primme->stats.numOrthoInnerProds += (double)nVL * nb + (double)nb * nb;

      /* stop one sweep after the block first looks well conditioned (:652-665) */
      if (rank_estimation(G, nVL, numLocked + b2, maxRank, ldG) == numLocked + b2) {
         if (its >= plus1) {
            int i;
            for (i = b1; i < b2 && PB_ABS(G[(size_t)ldG * (numLocked + i) + numLocked + i] - 1.0) < .8; i++)
               ;
            if (i >= b2) break;
         } else
            plus1 = PB_MIN(its + 1, plus1);
      }

      /* overflowed norms: drop the inner products with the other vectors (:668-674) */
      for (int i = 0; i < nb; i++) {
         if (PB_ABS(A[(size_t)ldG * i + i]) < DBL_MAX) continue;
         hl_zero(&G[(size_t)ldG * (numLocked + i)], numLocked + i, 1, ldG);
         A[(size_t)ldG * i + i] = DBL_MAX;
      }

      /* C = X'X - (X'Vp) inv(Vp'Vp) (Vp'X), GdA = inv(Vp'Vp) Vp'X  (:682-691) */
      hl_copy(&A[nVL], nb, nb, ldG, C, nb);
      hl_copy(A, nVL, nb, ldG, GdA, nVL > 0 ? nVL : 1);
      hl_trsm('L', 'U', 'C', 'N', nVL, nb, 1.0, fG, ldG, GdA, nVL > 0 ? nVL : 1);
      hl_gemm('C', 'N', nb, nb, nVL, -1.0, GdA, nVL > 0 ? nVL : 1, GdA, nVL > 0 ? nVL : 1, 1.0, C, nb);
      hl_trsm('L', 'U', 'N', 'N', nVL, nb, 1.0, fG, ldG, GdA, nVL > 0 ? nVL : 1);

      for (int i = 0; i < nb; i++) N[i] = sqrt(PB_MAX(PB_ABS(C[(size_t)nb * i + i]), eps_orth));
      for (int i = 0; i < nb; i++)
         for (int j = 0; j <= i; j++) C[(size_t)nb * i + j] /= N[i] * N[j];

      rc = decomposition(C, nb, nb, Y, nb, D, &Yortho);
      if (rc) goto done;
      for (int i = 0; i < nb; i++) D[i] = sqrt(PB_MAX(D[i], eps_orth * nb));

      /* accumulate the rotations applied to the block (:718-757) */
      if (RLocked) hl_gemm('N', 'N', numLocked, nb, nb, 1.0, GdA, nVL > 0 ? nVL : 1, r, ldr, 1.0, RLocked, ldRLocked);
      if (R) hl_gemm('N', 'N', b1, nb, nb, 1.0, &GdA[numLocked], nVL > 0 ? nVL : 1, r, ldr, 1.0, &R[(size_t)ldR * b1], ldR);
      if (r) {
         for (int i = 0; i < nb; i++)
            for (int j = 0; j < nb; j++) r[(size_t)ldr * i + j] *= N[j];
         if (Yortho)
            hl_gemm('C', 'N', nb, nb, nb, 1.0, Y, nb, r, ldr, 0.0, C, nb);
         else {
            /* C = U * r */
            for (int j = 0; j < nb; j++)
               for (int i = 0; i < nb; i++) {
                  SCALAR s = 0.0;
                  for (int l = i; l < nb; l++) s += Y[(size_t)nb * l + i] * r[(size_t)ldr * j + l];
                  C[(size_t)nb * j + i] = s;
               }
         }
         for (int i = 0; i < nb; i++)
            for (int j = 0; j < nb; j++) r[(size_t)ldr * i + j] = D[j] * C[(size_t)nb * i + j];
      }

      /* fold the column scaling into Y (:760-772) */
      if (Yortho) {
         for (int i = 0; i < nb; i++)
            for (int j = 0; j < nb; j++) Y[(size_t)nb * i + j] /= N[j];
      } else {
         for (int i = 0; i < nb; i++)
            for (int j = 0; j < nb; j++) Y[(size_t)nb * i + j] *= N[i];
      }
   }

   b2 = rank_estimation(G, nVL, numLocked + b2, maxRank, ldG) - numLocked;
   *b2_out = b2;
   rc = pb_update_cholesky_gram(G, fG, ldG, nVL, numLocked + b2);
done:
   free(D), free(N), free(GdA), free(Y), free(Yapply), free(C), free(r_own);
   primme->stats.timeOrtho += hl_wtime() - t0;
   return rc;
}

/* X(:,inX) <- (I - Q inv(Q'Q) Q') X(:,inX), norms of the results (:826-934). */
int pb_ortho_single_iteration(pb_solver *S, const SCALAR *Q, int nQ, int64_t ldQ,
      const SCALAR *QtQ, int ldQtQ, SCALAR *X, const int *inX, int nX, int64_t ldX,
      double *norms) {
   primme_params *primme = S->primme;
   const double t0 = hl_wtime();
   (void)QtQ, (void)ldQtQ; /* reference computes z = QtQ\y but applies y (:885-906): same here */
   SCALAR *panel = (SCALAR *)malloc(sizeof(SCALAR) * (size_t)(nQ + 1));
   if (!panel) return PRIMME_MALLOC_FAILURE;
   for (int j = 0; j < nX; j++) {
      SCALAR *x = X + (size_t)ldX * (inX ? inX[j] : j);
      CHK(pb200_dortho_sweep(S->dev, S->n, Q, nQ, ldQ, NULL, 0, 0, x, 1, ldX, NULL, 0, NULL, 0, 0,
            panel, nQ + 1));
      CHK(pb_reduce_panel(S, panel, nQ, 1, nQ + 1));
      CHK(pb200_dortho_sweep(S->dev, S->n, Q, nQ, ldQ, NULL, 0, 0, x, 1, ldX, panel, nQ + 1, NULL,
            0, 1, norms ? panel : NULL, nQ + 1));
      if (norms) {
         CHK(pb_reduce_panel(S, panel, nQ + 1, 1, nQ + 1));
         norms[j] = sqrt(PB_REAL(panel[nQ]));
      }
   }
   primme->stats.numOrthoInnerProds += (double)nQ * nX + (norms ? nX : 0);
   free(panel);
   primme->stats.timeOrtho += hl_wtime() - t0;
   return 0;
}

/* ------------------------------------------------------------------------------------------
 * Host-only Gram-Schmidt on small coefficient vectors with inner product B (upper stored) or
 * identity (reference Bortho_local :395-413 -> Bortho_gen with primme == NULL: up to 7
 * passes, always reorthogonalise).  Orthonormalises V(:,b1..b2) against locked(:,0:numLocked)
 * and V(:,0:b1).  R (column stride 1 x 1 use only) receives the final norm or 0 if the vector
 * had to be replaced by a random one.  Returns 0 or -3.
 * ---------------------------------------------------------------------------------------- */
static int ortho_local_gen(SCALAR *V, int ldV, SCALAR *R, int ldR, int b1, int b2, SCALAR *locked, int ldLocked,
      int numLocked, int n, const SCALAR *B, int ldB, long long *iseed) {
   const int maxNumOrthos = 7, maxNumRandoms = 10;
   const double tol = sqrt(2.0) / 2.0, eps_orth = PB_EPS;
   SCALAR *overlaps = (SCALAR *)calloc((size_t)(b2 + 2 + numLocked), sizeof(SCALAR));
   SCALAR *Bx = (SCALAR *)malloc(sizeof(SCALAR) * (n > 0 ? n : 1));
   int ok = 1;
   if (R && b2 >= b1) hl_zero(&R[(size_t)ldR * b1], b2 + 1, b2 - b1 + 1, ldR); /* (:146-149) */
   for (int i = b1; i <= b2 && ok; i++) {
      SCALAR *v = &V[(size_t)ldV * i];
      int nOrth = 0, randomizations = 0, updateR = R ? 1 : 0, Bx_update = 0;
      double s0 = 0, s02 = 0, s1 = 0, s12 = 0;
      for (;;) {
         if (nOrth >= maxNumOrthos) {
            if (updateR) {
               if (R) R[(size_t)ldR * i + i] = 0.0;
               updateR = 0;
            }
            if (randomizations >= maxNumRandoms) {
               ok = 0;
               break;
            }
            hl_larnv2(iseed, n, v);
            randomizations++;
            nOrth = 0;
            Bx_update = 0;
         }
         nOrth++;
         const SCALAR *bx = v;
         if (B) {
            if (!Bx_update) {
               hl_zero(Bx, n, 1, n);
               hl_symm_lu(n, 1, 1.0, B, ldB, v, ldV > n ? ldV : n, 0.0, Bx, n);
            }
            bx = Bx;
         }
         if (nOrth == 1) s02 = PB_REAL(hl_dot(n, v, bx));
         if (i > 0) hl_gemm('C', 'N', i, 1, n, 1.0, V, ldV, bx, n, 0.0, overlaps, i);
         if (numLocked > 0)
            hl_gemm('C', 'N', numLocked, 1, n, 1.0, locked, ldLocked, bx, n, 0.0, &overlaps[i], numLocked);
         overlaps[i + numLocked] = s02;
         if (updateR && R)
            for (int j = 0; j < i; j++) R[(size_t)ldR * i + j] += overlaps[j]; /* (:226-229) */
         if (numLocked > 0)
            hl_gemm('N', 'N', n, 1, numLocked, -1.0, locked, ldLocked, &overlaps[i], numLocked, 1.0, v, n);
         if (i > 0) hl_gemm('N', 'N', n, 1, i, -1.0, V, ldV, overlaps, i, 1.0, v, n);
         Bx_update = 0;
         if (nOrth == 1) s0 = sqrt(s02 = PB_REAL(overlaps[i + numLocked]));
         if (B) {
            hl_zero(Bx, n, 1, n);
            hl_symm_lu(n, 1, 1.0, B, ldB, v, ldV > n ? ldV : n, 0.0, Bx, n);
            Bx_update = 1;
            bx = Bx;
         } else
            bx = v;
         s12 = PB_REAL(hl_dot(n, v, bx));
         s1 = sqrt(s12);
         if (!isfinite(s0) || !isfinite(s1) || s1 <= eps_orth * s0) {
            nOrth = maxNumOrthos;
         } else if (s1 <= tol * s0 || nOrth < maxNumOrthos) {
            s0 = s1, s02 = s12;
         } else {
            if (updateR && R) R[(size_t)ldR * i + i] = s1;
            double inv = 1.0 / s1;
            if (isfinite(inv)) {
               for (int t = 0; t < n; t++) v[t] *= inv;
               break;
            }
            nOrth = maxNumOrthos;
         }
      }
   }
   free(overlaps), free(Bx);
   return ok ? 0 : -3;
}

int pb_ortho_local(SCALAR *V, int ldV, SCALAR *R, int b1, int b2, SCALAR *locked, int ldLocked,
      int numLocked, int n, const SCALAR *B, int ldB, long long *iseed) {
   return ortho_local_gen(V, ldV, R, 1, b1, b2, locked, ldLocked, numLocked, n, B, ldB, iseed);
}

/* same with the full factor: V(:,b1..b2) = V_out * R(:,b1..b2), R with leading dimension ldR */
int pb_ortho_local_R(SCALAR *V, int ldV, SCALAR *R, int ldR, int b1, int b2, int n, const SCALAR *B, int ldB,
      long long *iseed) {
   if (b2 < b1) return 0;
   return ortho_local_gen(V, ldV, R, ldR, b1, b2, NULL, 0, 0, n, B, ldB, iseed);
}

/* dav_jdqmr.c -- inner solver of the Jacobi-Davidson correction equation: block symmetric QMR with
 * adaptive stopping.
 *
 * Restates reference src/eigs/inner_solve.c:132-636 (inner_solve_Sprimme) and its helpers
 * apply_projected_preconditioner :714-744, apply_skew_projector :769-812, apply_projected_matrix
 * :838-890: the operator is (I - QQ')(I - xx')(A - shift I) with the left projectors the method selects,
 * the preconditioner is followed by the right projectors (I - RX x'/x'RX)(I - QQ') when the method asks for
 * them (PRIMME_JDQR, user settings), orthogonal or skew (the skew-Q projector with a preconditioner applies
 * evecsHat = K^{-1}Q and the factors of M = Q'K^{-1}Q maintained by pb_update_XKinvBX); B = I.
 *
 * Every n-long operation is a kernel of the C-ABI (column dots, column axpy / scale, permutes, the
 * fused ortho sweep for Q'v and v - Q(Q'v)); the scalar recurrences stay on the host exactly as in
 * the reference.  Arrays indexed [p[i]] follow the ORIGINAL position of a right-hand side, arrays
 * indexed [i] travel with the permutation that moves finished systems to the end of the block.
 */
#include "pb_host.h"
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* v(:,0:bs) <- (I - Q Q') v: overlaps through the fused sweep (one Gram pass, one update pass) */
static int project_out(pb_solver *S, const SCALAR *Q, int64_t ldQ, int nQ, SCALAR *v, int64_t ldv, int bs) {
   primme_params *primme = S->primme;
   if (nQ <= 0 || bs <= 0) return 0;
   const double t0 = hl_wtime();
   SCALAR *ov = (SCALAR *)malloc(sizeof(SCALAR) * (size_t)nQ * bs);
   if (!ov) return PRIMME_MALLOC_FAILURE;
   int rc = pb200_dortho_sweep(S->dev, S->n, Q, nQ, ldQ, NULL, 0, 0, v, bs, ldv, NULL, 0, NULL, 0, 0, ov, nQ);
   primme->stats.numOrthoInnerProds += (double)nQ * bs;
   if (!rc) rc = pb_reduce_panel(S, ov, nQ, bs, nQ);
   if (!rc) rc = pb200_dortho_sweep(S->dev, S->n, Q, nQ, ldQ, NULL, 0, 0, v, bs, ldv, ov, nQ, NULL, 0, 0, NULL, 0);
   free(ov);
   primme->stats.timeOrtho += hl_wtime() - t0;
   return rc;
}

/* v_i <- (I - x_i x_i') v_i for every column: one batch of dots, one batch of axpys */
static int project_out_each(pb_solver *S, const SCALAR *x, int64_t ldx, SCALAR *v, int64_t ldv, int bs) {
   primme_params *primme = S->primme;
   if (bs <= 0) return 0;
   const double t0 = hl_wtime();
   SCALAR ov[8];
   int rc = pb200_dcolumn_dots(S->dev, S->n, x, ldx, v, ldv, bs, ov);
   primme->stats.numOrthoInnerProds += bs;
   if (!rc) rc = pb_reduce_panel(S, ov, bs, 1, bs);
   for (int i = 0; i < bs; i++) ov[i] = -ov[i];
   if (!rc) rc = pb200_daxpy_columns(S->dev, S->n, ov, x, ldx, v, ldv, bs);
   primme->stats.timeOrtho += hl_wtime() - t0;
   return rc;
}

/* result = (I - x x')(I - Q Q')(A - shift) v   (inner_solve.c:838-890 with B = I) */
static int apply_projected_matrix(pb_solver *S, SCALAR *v, int64_t ldv, const double *shift, const SCALAR *Q,
      int64_t ldQ, int nQ, const SCALAR *X, int64_t ldX, int nX, int bs, SCALAR *result, int64_t ldr) {
   CHK(pb_apply_matvec(S, v, ldv, result, ldr, bs));
   SCALAR ms[8];
   for (int i = 0; i < bs; i++) ms[i] = -shift[i];
   CHK(pb200_daxpy_columns(S->dev, S->n, ms, v, ldv, result, ldr, bs));
   CHK(project_out(S, Q, ldQ, nQ, result, ldr, bs));
   if (nX > 0) CHK(project_out_each(S, X, ldX, result, ldr, bs));
   return 0;
}

/* right projectors of the correction equation (setup_JD_projectors, correction.c:942-980): applied after the
 * preconditioner,  result <- (I - RX_i x_i'/xKinvBx_i)(I - RQ RQ') K^{-1} v  (inner_solve.c:714-812 with no
 * skew-Q factorisation: RQ = the locked vectors themselves) */
typedef struct right_projectors {
   const SCALAR *RQ; /* evecs, nRQ columns */
   int64_t ldRQ;
   int nRQ;
   SCALAR *RX;       /* x (orthogonal) or K^{-1}x (skew), n x blockSize; permuted with the systems */
   int64_t ldRX;
   int nRX;          /* 0 or the current block size */
   SCALAR *xKinvBx;  /* x_i' K^{-1} x_i, or ones; indexed by POSITION (permuted with the systems) */
   /* skew-Q projector (I - RQ (Q'RQ)^{-1} Q') with RQ = K^{-1} Q: Q, and the Bunch-Kaufman factors of Q'RQ packed
    * with leading dimension nRQ (pb_update_XKinvBX); skewQ == NULL: orthogonal projector (I - RQ RQ') */
   const SCALAR *skewQ;
   int64_t ldskewQ;
   const SCALAR *Mfact;
   const int *ipivot;
} right_projectors;

/* v <- (I - Qhat (Q'Qhat)^{-1} Q') v  (apply_skew_projector with Mfact, inner_solve.c:769-812; MSolve,
 * factorize.c:268-297) */
static int skew_project_out(pb_solver *S, const right_projectors *rp, SCALAR *v, int64_t ldv, int bs) {
   primme_params *primme = S->primme;
   const int nQ = rp->nRQ;
   if (nQ <= 0 || bs <= 0) return 0;
   const double t0 = hl_wtime();
   SCALAR *ov = (SCALAR *)malloc(sizeof(SCALAR) * (size_t)nQ * bs);
   if (!ov) return PRIMME_MALLOC_FAILURE;
   int rc = pb200_dortho_sweep(S->dev, S->n, rp->skewQ, nQ, rp->ldskewQ, NULL, 0, 0, v, bs, ldv, NULL, 0, NULL, 0, 0, ov, nQ);
   primme->stats.numOrthoInnerProds += (double)nQ * bs;
   if (!rc) rc = pb_reduce_panel(S, ov, nQ, bs, nQ);
   if (!rc) {
      if (nQ == 1)
         for (int i = 0; i < bs; i++) ov[i] = ov[i] / rp->Mfact[0];
      else if (hl_hetrs_upper(nQ, bs, rp->Mfact, nQ, rp->ipivot, ov, nQ))
         rc = PRIMME_LAPACK_FAILURE;
   }
   if (!rc) rc = pb200_dortho_sweep(S->dev, S->n, rp->RQ, nQ, rp->ldRQ, NULL, 0, 0, v, bs, ldv, ov, nQ, NULL, 0, 0, NULL, 0);
   free(ov);
   primme->stats.timeOrtho += hl_wtime() - t0;
   return rc;
}

static int apply_projected_preconditioner(pb_solver *S, right_projectors *rp, SCALAR *v, int64_t ldv, const SCALAR *x,
      int64_t ldx, SCALAR *result, int64_t ldr, int bs) {
   primme_params *primme = S->primme;
   CHK(pb_apply_precond(S, v, ldv, result, ldr, bs));
   if (!rp) return 0;
   if (rp->skewQ)
      CHK(skew_project_out(S, rp, result, ldr, bs));
   else
      CHK(project_out(S, rp->RQ, rp->ldRQ, rp->nRQ, result, ldr, bs));
   if (rp->nRX <= 0) return 0;
   const double t0 = hl_wtime();
   SCALAR ov[8];
   CHK(pb200_dcolumn_dots(S->dev, S->n, x, ldx, result, ldr, bs, ov));
   primme->stats.numOrthoInnerProds += bs;
   CHK(pb_reduce_panel(S, ov, bs, 1, bs));
   for (int i = 0; i < bs; i++) ov[i] = -ov[i] / rp->xKinvBx[i];
   CHK(pb200_daxpy_columns(S->dev, S->n, ov, rp->RX, rp->ldRX, result, ldr, bs));
   primme->stats.timeOrtho += hl_wtime() - t0;
   return 0;
}

/* sums over the processes still to be done through the host callback (no multi-rank kernel context) */
#define primme_host_sums(S) ((S)->primme->numProcs > 1 && pb200_ctx_nranks((S)->dev) <= 1)

static void perm_set_value_on_pos(int *p, int val, int pos, int n) {
   for (int i = 0; i < n; i++)
      if (p[i] == val) {
         p[i] = p[pos];
         p[pos] = val;
         return;
      }
}

/* real parts of the column dots (Num_dist_dots_real, auxiliary_eigs.c:662-700) */
static int dots_real(pb_solver *S, const SCALAR *a, int64_t lda, const SCALAR *b, int64_t ldb, int bs, double *out) {
   SCALAR tmp[8];
   CHK(pb200_dcolumn_dots(S->dev, S->n, a, lda, b, ldb, bs, tmp));
   CHK(pb_reduce_panel(S, tmp, bs, 1, bs));
   for (int i = 0; i < bs; i++) out[i] = PB_REAL(tmp[i]);
   return 0;
}
/* the recurrences' scalars are real in every precision (inner_solve.c:155-182) */
static int axpy_real(pb_solver *S, int64_t n, const double *alpha, const SCALAR *X, int64_t ldx, SCALAR *Y, int64_t ldy,
      int bs) {
   SCALAR a[8];
   for (int i = 0; i < bs; i++) a[i] = alpha[i];
   return pb200_daxpy_columns(S->dev, n, a, X, ldx, Y, ldy, bs);
}
static int scale_real(pb_solver *S, int64_t n, const double *alpha, SCALAR *X, int64_t ldx, int bs) {
   SCALAR a[8];
   for (int i = 0; i < bs; i++) a[i] = alpha[i];
   return pb200_dscale_columns(S->dev, n, a, X, ldx, bs);
}

/* x, r, sol: n x blockSize device blocks (leading dimension S->ld for x and r, ldsol for sol);
 * Q: the left projector (nQ columns, may be NULL); useX: project against every x_i as well.
 * rnorm / eval are indexed by original position, shift travels with the permutation. */
int pb_inner_solve(pb_solver *S, int blockSize, SCALAR *x, int64_t ldx, SCALAR *r, int64_t ldr, const double *rnorm,
      const SCALAR *Q, int64_t ldQ, int nQ, int useX, SCALAR *sol, int64_t ldsol, const double *eval, double *shift,
      int *touch, SCALAR *work, const SCALAR *RQ, int64_t ldRQ, int nRQ, SCALAR *RX, int64_t ldRX, SCALAR *xKinvBx,
      const SCALAR *skewQ, int64_t ldskewQ, const SCALAR *Mfact, const int *ipivot) {
   right_projectors rpv = {RQ, ldRQ, nRQ, RX, ldRX, RX ? blockSize : 0, xKinvBx, nRQ > 0 ? skewQ : NULL, ldskewQ, Mfact, ipivot};
   right_projectors *rp = (nRQ > 0 || RX) ? &rpv : NULL;
   primme_params *primme = S->primme;
   const correction_params *cp = &primme->correctionParams;
   const int64_t n = S->n, ldw = S->ld;
   const int bs0 = blockSize;
   int sizeX = useX ? blockSize : 0;
   SCALAR *g = work, *d = g + (size_t)ldw * bs0, *delta = d + (size_t)ldw * bs0, *w = delta + (size_t)ldw * bs0;
   double sigma_prev[8], rho_prev[8], rho[8], alpha_prev[8], Theta_prev[8], Theta[8], tau_init[8], tau_prev[8],
         tau[8], Beta_prev[8], Delta_prev[8], Psi_prev[8], eta[8], eval_prev[8], eres_updated[8], Gamma_prev[8],
         Phi_prev[8], gamma[8], dot_sol[8], one[8];
   int p[8], p0[8];
   const int adaptive = cp->convTest == primme_adaptive || cp->convTest == primme_adaptive_ETolerance;
   int rc = 0, i;

   for (i = 0; i < blockSize; i++) tau_prev[i] = tau_init[i] = rnorm[i]; /* zero initial guess */
   /* in any case stop when the linear residual is below max(machEps, eps) * |A| (:217-222) */
   double LTolerance = PB_EPS * pb_problem_norm(1, primme);
   double LTolerance_factor = 1.0, ETolerance = 0.0, ETolerance_factor = 0.0;
   switch (cp->convTest) {
   case primme_full_LTolerance: break;
   case primme_decreasing_LTolerance:
      LTolerance = PB_MAX(LTolerance, pow(cp->relTolBase, -(double)*touch));
      (*touch)++;
      break;
   case primme_adaptive:
      LTolerance_factor = pow(1.8, -(double)*touch);
      ETolerance_factor = pow(1.8, -(double)*touch);
      break;
   case primme_adaptive_ETolerance:
      LTolerance_factor = pow(1.8, -(double)*touch);
      ETolerance_factor = pow(1.8, -(double)*touch);
      ETolerance = 0.1;
      break;
   }
   PRIMME_INT maxIterations = primme->maxMatvecs > 0 ? primme->maxMatvecs - primme->stats.numMatvecs : INT_MAX;
   if (maxIterations > INT_MAX) maxIterations = INT_MAX;
   if (cp->maxInnerIterations > 0) maxIterations = PB_MIN((PRIMME_INT)cp->maxInnerIterations, maxIterations);

   /* g = r, d = (right projectors) K^{-1} g */
   CHK(pb200_copy_d2d(S->dev, r, ldr, g, ldw, n, blockSize, PB_ES));
   CHK(apply_projected_preconditioner(S, rp, g, ldw, x, ldx, d, ldw, blockSize));
   for (i = 0; i < blockSize; i++) Theta_prev[i] = 0.0, eval_prev[i] = eval[i];
   CHK(dots_real(S, g, ldw, d, ldw, blockSize, rho_prev));
   for (i = 0; i < blockSize; i++)
      Beta_prev[i] = Delta_prev[i] = Psi_prev[i] = Gamma_prev[i] = Phi_prev[i] = eres_updated[i] = 0.0, one[i] = 1.0;
   CHK(pb200_memset0(S->dev, delta, sizeof(SCALAR) * (size_t)ldw * blockSize));
   for (i = 0; i < blockSize; i++) CHK(pb200_memset0(S->dev, sol + (size_t)ldsol * i, sizeof(SCALAR) * (size_t)n));

   for (i = 0; i < blockSize; i++) p[i] = i;
   for (PRIMME_INT numIts = 0; numIts < maxIterations && blockSize > 0; numIts++) {
      CHK(apply_projected_matrix(S, d, ldw, shift, Q, ldQ, nQ, x, ldx, sizeX, blockSize, w, ldw));
      /* NOTE (parity): the reference lets Num_dist_dots write the block's dots in POSITION order and
       * then reads them as [p[i]] (:330-333, :388-390, :607-611).  The two orders differ only after a
       * system of a block > 1 has finished early; the indexing is restated as it is. */
      CHK(dots_real(S, d, ldw, w, ldw, blockSize, sigma_prev));
      int conv = 0;
      double alpha_neg[8];
      for (i = 0; i < blockSize; i++) p0[i] = i, alpha_neg[i] = 0.0;
      for (i = 0; i < blockSize; i++) {
         if (!isfinite(sigma_prev[p[i]]) || sigma_prev[p[i]] == 0.0) {
            if (numIts == 0) CHK(pb200_copy_d2d(S->dev, r + (size_t)ldr * i, ldr, sol + (size_t)ldsol * i, ldsol, n, 1, PB_ES));
            perm_set_value_on_pos(p0, i, blockSize - ++conv, blockSize);
            continue;
         }
         alpha_prev[p[i]] = rho_prev[p[i]] / sigma_prev[p[i]];
         if (!isfinite(alpha_prev[p[i]]) || fabs(alpha_prev[p[i]]) < PB_EPS || fabs(alpha_prev[p[i]]) > 1.0 / PB_EPS) {
            if (numIts == 0) CHK(pb200_copy_d2d(S->dev, r + (size_t)ldr * i, ldr, sol + (size_t)ldsol * i, ldsol, n, 1, PB_ES));
            perm_set_value_on_pos(p0, i, blockSize - ++conv, blockSize);
            continue;
         }
         alpha_neg[i] = -alpha_prev[p[i]];
      }
      /* g_i -= alpha_i w_i for the systems still running (alpha = 0 leaves the others untouched); with no
       * system leaving the block the same pass also delivers |g_i|^2 (one launch instead of two) */
      int have_theta = 0;
      if (conv == 0) {
         double alpha_pos[8];
         for (i = 0; i < blockSize; i++) alpha_pos[i] = -alpha_neg[i];
         CHK(pb200_dresidual_inplace(S->dev, n, alpha_pos, w, ldw, g, ldw, blockSize, Theta));
         if (primme_host_sums(S)) CHK(pb_global_sum(S, Theta, blockSize));
         have_theta = 1;
      } else
         CHK(axpy_real(S, n, alpha_neg, w, ldw, g, ldw, blockSize));

#define PB_PERMUTE_BLOCK()                                                                        \
   do {                                                                                           \
      hl_permute_ints(p, blockSize, p0);                                                          \
      hl_permute_reals(shift, blockSize, p0);                                                     \
      CHK(pb200_dpermute_columns(S->dev, n, g, ldw, p0, blockSize));                              \
      CHK(pb200_dpermute_columns(S->dev, n, d, ldw, p0, blockSize));                              \
      CHK(pb200_dpermute_columns(S->dev, n, delta, ldw, p0, blockSize));                          \
      CHK(pb200_dpermute_columns(S->dev, n, r, ldr, p0, blockSize));                              \
      CHK(pb200_dpermute_columns(S->dev, n, x, ldx, p0, blockSize));                              \
      CHK(pb200_dpermute_columns(S->dev, n, sol, ldsol, p0, blockSize));                          \
      if (rp && rp->nRX) {                                                                        \
         hl_permute_cols(rp->xKinvBx, 1, blockSize, 1, p0);                                       \
         if (rp->RX != x) CHK(pb200_dpermute_columns(S->dev, n, rp->RX, rp->ldRX, p0, blockSize)); \
         rp->nRX -= conv;                                                                         \
      }                                                                                           \
      blockSize -= conv;                                                                          \
      if (sizeX) sizeX -= conv;                                                                   \
   } while (0)

      if (conv > 0) PB_PERMUTE_BLOCK();
      if (blockSize <= 0) break;

      if (!have_theta) CHK(dots_real(S, g, ldw, g, ldw, blockSize, Theta));
      double gam_pos[8], eta_pos[8];
      for (i = 0; i < blockSize; i++) {
         Theta[p[i]] = sqrt(Theta[p[i]]) / tau_prev[p[i]];
         const double c = 1.0 / sqrt(1 + Theta[p[i]] * Theta[p[i]]);
         tau[p[i]] = tau_prev[p[i]] * Theta[p[i]] * c;
         gamma[p[i]] = c * c * Theta_prev[p[i]] * Theta_prev[p[i]];
         eta[p[i]] = alpha_prev[p[i]] * c * c;
         gam_pos[i] = gamma[p[i]], eta_pos[i] = eta[p[i]];
      }
      /* delta = gamma delta + eta d; sol += delta; |sol|^2 (device flavour of :395-413) */
      CHK(pb200_dqmr_update(S->dev, n, gam_pos, eta_pos, d, ldw, delta, ldw, sol, ldsol, blockSize, adaptive ? dot_sol : NULL));
      if (adaptive && primme_host_sums(S)) CHK(pb_global_sum(S, dot_sol, blockSize));

      conv = 0;
      for (i = 0; i < blockSize; i++) p0[i] = i;
      for (i = 0; i < blockSize; i++) {
         const int pi = p[i];
         int isConv = 0;
         if (fabs(rho_prev[pi]) == 0.0) {
            perm_set_value_on_pos(p0, i, blockSize - ++conv, blockSize);
            continue;
         }
         if (numIts > 0 && tau[pi] < LTolerance) {
            perm_set_value_on_pos(p0, i, blockSize - ++conv, blockSize);
            continue;
         }
         if (ETolerance > 0.0 || ETolerance_factor > 0.0) {
            /* adaptive stopping: recurrences for the Ritz value and eigenresidual of x + sol */
            const double Delta = gamma[pi] * Delta_prev[pi] + eta[pi] * rho_prev[pi];
            const double Beta = Beta_prev[pi] - Delta;
            const double Phi = gamma[pi] * gamma[pi] * Phi_prev[pi] + eta[pi] * eta[pi] * sigma_prev[pi];
            const double Psi = gamma[pi] * Psi_prev[pi] + gamma[pi] * Phi_prev[pi];
            const double Gamma = Gamma_prev[pi] + 2.0 * Psi + Phi;
            const double nrm = 1.0 + dot_sol[i];
            const double eval_updated = shift[i] + (eval[pi] - shift[i] + 2 * Beta + Gamma) / nrm;
            const double eres2_updated = (tau[pi] * tau[pi]) / nrm +
                                         ((eval[pi] - shift[i] + Beta) * (eval[pi] - shift[i] + Beta)) / nrm -
                                         (eval_updated - shift[i]) * (eval_updated - shift[i]);
            const double eres_prev = eres_updated[pi];
            eres_updated[pi] = eres2_updated < 0 ? sqrt((tau[pi] * tau[pi]) / nrm) : sqrt(eres2_updated);
            Delta_prev[pi] = Delta, Beta_prev[pi] = Beta, Phi_prev[pi] = Phi, Psi_prev[pi] = Psi, Gamma_prev[pi] = Gamma;

            if (numIts > 0 && (tau_prev[pi] <= eres_updated[pi] || eres_prev <= tau[pi])) {
               perm_set_value_on_pos(p0, i, blockSize - ++conv, blockSize);
               continue;
            }
            if (primme->target == primme_smallest && eval_updated > eval_prev[pi]) {
               perm_set_value_on_pos(p0, i, blockSize - ++conv, blockSize);
               continue;
            } else if (primme->target == primme_largest && eval_updated < eval_prev[pi]) {
               perm_set_value_on_pos(p0, i, blockSize - ++conv, blockSize);
               continue;
            } else if (primme->target == primme_closest_abs &&
                       fabs(eval[pi] - eval_updated) > tau_init[pi] + eres_updated[pi]) {
               perm_set_value_on_pos(p0, i, blockSize - ++conv, blockSize);
               continue;
            }
            if (numIts > 0 && eres_updated[pi] < ETolerance * tau_init[pi]) {
               perm_set_value_on_pos(p0, i, blockSize - ++conv, blockSize);
               continue;
            }
            const double tol = PB_MIN(tau[pi] / LTolerance_factor, eres_updated[pi] / ETolerance_factor);
            CHK(pb_conv_test(S, eval_updated, NULL, tol, &isConv));
            if (numIts > 0 && isConv) {
               (*touch)++;
               perm_set_value_on_pos(p0, i, blockSize - ++conv, blockSize);
               continue;
            }
            eval_prev[pi] = eval_updated;
            if (primme->monitorFun) {
               int zero = 0, unco = UNCONVERGED;
               double evalr = eval_updated, resr = eres_updated[pi];
               CHK(pb_monitor(S, &evalr, 1, &unco, &zero, 1, &resr, -1, NULL, -1, NULL, NULL, (int)numIts, tau[pi], NULL,
                     0.0, primme_event_inner_iteration));
            }
         } else {
            /* the QMR residual can be sqrt(iterations) away from the true one (:574-580) */
            CHK(pb_conv_test(S, eval[pi], NULL, tau[pi] / LTolerance_factor * sqrt((double)numIts), &isConv));
            if (numIts > 0 && isConv) {
               perm_set_value_on_pos(p0, i, blockSize - ++conv, blockSize);
               continue;
            } else if (primme->monitorFun) {
               int zero = 0, unco = UNCONVERGED;
               double evalr = eval[pi], resr = rnorm[pi];
               CHK(pb_monitor(S, &evalr, 1, &unco, &zero, 1, &resr, 0, NULL, 0, NULL, NULL, (int)numIts, tau[pi], NULL,
                     0.0, primme_event_inner_iteration));
            }
         }
      }
      if (conv > 0) PB_PERMUTE_BLOCK();
      if (blockSize <= 0) break;

      if (numIts + 1 < maxIterations) {
         CHK(apply_projected_preconditioner(S, rp, g, ldw, x, ldx, w, ldw, blockSize));
         double beta_pos[8];
         CHK(dots_real(S, g, ldw, w, ldw, blockSize, rho));
         for (i = 0; i < blockSize; i++) {
            beta_pos[i] = rho[p[i]] / rho_prev[p[i]];
            rho_prev[p[i]] = rho[p[i]];
            tau_prev[p[i]] = tau[p[i]];
            Theta_prev[p[i]] = Theta[p[i]];
         }
         CHK(axpy_real(S, n, beta_pos, d, ldw, w, ldw, blockSize));
         SCALAR *ptmp = d; /* alternate the buffers instead of copying (:618-622) */
         d = w, w = ptmp;
      }
   }
#undef PB_PERMUTE_BLOCK
   return rc;
}

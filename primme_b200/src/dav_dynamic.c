/* dav_dynamic.c -- PRIMME_DYNAMIC: run-time choice between GD+k and JDQMR from a cost model fed by
 * wall-clock measurements of the running solve.
 *
 * Restates reference src/eigs/main_iter.c:1943-2011 (switch_from_JDQMR), :2050-2108
 * (switch_from_GDpk), :2186-2330 (update_statistics), :2340-2356 (ratio_JDQMR_GDpk), :2366-2398
 * (update_slowdown), :2403-2440 (initializeModel) and the structure main_iter_private.h:60-110.
 * The model compares  JDQMR_time / GDpk_time = slowdown * (q+mv+pr + (gd - 2q - pr) kout/nMV) /
 * (gd+mv+pr)  with thresholds 0.95 / 1.05.  Its inputs are timings, so which method runs is a
 * property of the machine, in the reference as here; on a B200 the outer iteration is a handful of
 * latency-bound kernel round trips while an inner QMR step costs about as many, which the model
 * sees through the same measurements.
 */
#include "pb_host.h"
#include <math.h>

#ifndef PB_COMPLEX
void pb_dyn_init(pb_cost_model *m, primme_params *primme) {
   m->MV_PR = m->MV = m->PR = m->qmr_only = m->qmr_plus_MV_PR = 0.0;
   m->gdk_plus_MV_PR = m->gdk_plus_MV = m->project_locked = m->reortho_locked = 0.0;
   m->gdk_conv_rate = m->jdq_conv_rate = 0.0001;
   m->JDQMR_slowdown = 1.5;
   m->ratio_MV_outer = 0.0;
   m->nextReset = 1;
   m->gdk_sum_logResReductions = m->jdq_sum_logResReductions = 0.0;
   m->gdk_sum_MV = m->jdq_sum_MV = 0.0;
   m->nevals_by_gdk = m->nevals_by_jdq = 0;
   m->numMV_0 = primme->stats.numMatvecs;
   m->numIt_0 = primme->stats.numOuterIterations + 1;
   m->timer_0 = hl_wtime();
   m->time_in_inner = 0.0;
   m->resid_0 = -1.0;
   m->accum_jdq = m->accum_gdk = 0.0;
   m->accum_jdq_gdk = 1.0;
}

/* first sample, then the mean of the previous estimate and the new sample */
#endif
static double blend(double estimate, double sample) { return estimate == 0.0 ? sample : (estimate + sample) / 2.0; }

static double ratio_jdqmr_gdpk(const pb_cost_model *m, int numLocked, double slowdown, double ratio_MV_outer) {
   return slowdown *
          (m->qmr_plus_MV_PR + m->project_locked * numLocked +
                (m->gdk_plus_MV - m->qmr_only - m->qmr_plus_MV_PR +
                      (m->reortho_locked - m->project_locked) * numLocked) /
                      ratio_MV_outer) /
          (m->gdk_plus_MV_PR + m->reortho_locked * numLocked);
}

/* log(GD+k rate) / log(JDQMR rate), bounded by the inner iteration count and by [1.1, 2.5] */
static void update_slowdown(pb_cost_model *m) {
   double s;
   if (m->gdk_conv_rate < 1.0) {
      if (m->jdq_conv_rate < 1.0) s = log(m->gdk_conv_rate) / log(m->jdq_conv_rate);
      else if (m->jdq_conv_rate == 1.0) s = 2.5;
      else s = -log(m->gdk_conv_rate) / log(m->jdq_conv_rate);
   } else if (m->gdk_conv_rate == 1.0)
      s = 1.1;
   else {
      if (m->jdq_conv_rate < 1.0) s = log(m->gdk_conv_rate) / log(m->jdq_conv_rate);
      else if (m->jdq_conv_rate == 1.0) s = 1.1;
      else s = log(m->jdq_conv_rate) / log(m->gdk_conv_rate);
   }
   s = PB_MAX(m->ratio_MV_outer / (m->ratio_MV_outer - 1.0), PB_MIN(s, m->ratio_MV_outer));
   m->JDQMR_slowdown = PB_MAX(1.1, PB_MIN(s, 2.5));
}

/* returns 1 when the model was updated and the methods can be compared */
#ifndef PB_COMPLEX
int pb_dyn_update_statistics(pb_cost_model *m, primme_params *primme, double current_time, int recentConv,
      int calledAtRestart, int numConverged, double currentResNorm) {
   const double elapsed = current_time - m->timer_0;
   const double time_in_outer = elapsed - m->time_in_inner;
   int kout = (int)(primme->stats.numOuterIterations - m->numIt_0);
   const int nMV = (int)(primme->stats.numMatvecs - m->numMV_0);
   if (calledAtRestart) kout++;
   if (kout == 0) return 0;
   const double kinn = ((double)nMV) / kout - 2;
   if (primme->correctionParams.maxInnerIterations == -1 && (kinn < 1.0 && m->qmr_only == 0.0)) return 0;

   double low_res;
   if (recentConv > 0) {
      low_res = primme->stats.maxConvTol;
      if (primme->correctionParams.maxInnerIterations == -1) m->nevals_by_jdq += recentConv;
      else m->nevals_by_gdk += recentConv;
   } else
      low_res = currentResNorm;

   m->gdk_plus_MV = blend(m->gdk_plus_MV, time_in_outer / kout);

   /* the averaging window of the convergence rates restarts every 10 converged pairs (:2239-2258) */
   if (numConverged / 10 >= m->nextReset) {
      m->gdk_sum_logResReductions /= m->nevals_by_gdk;
      m->gdk_sum_MV /= m->nevals_by_gdk;
      m->jdq_sum_logResReductions /= m->nevals_by_jdq;
      m->jdq_sum_MV /= m->nevals_by_jdq;
      m->nextReset = numConverged / 10 + 1;
      m->nevals_by_gdk = 1;
      m->nevals_by_jdq = 1;
   }

   switch (primme->dynamicMethodSwitch) {
   case 1:
   case 3: /* GD+k is running */
      m->PR = blend(m->PR, m->time_in_inner / kout);
      m->gdk_plus_MV_PR = m->gdk_plus_MV + m->PR;
      m->MV_PR = m->MV + m->PR;
      if (low_res <= m->resid_0) m->gdk_sum_logResReductions += log(low_res / m->resid_0);
      m->gdk_sum_MV += nMV;
      m->gdk_conv_rate = exp(m->gdk_sum_logResReductions / m->gdk_sum_MV);
      break;
   case 2:
   case 4: /* JDQMR is running */
      if (m->qmr_plus_MV_PR == 0.0) {
         m->qmr_plus_MV_PR = (m->time_in_inner / kout - m->MV_PR) / kinn;
         m->ratio_MV_outer = ((double)nMV) / kout;
      } else {
         if (kinn != 0.0) m->qmr_plus_MV_PR = (m->qmr_plus_MV_PR + (m->time_in_inner / kout - m->MV_PR) / kinn) / 2.0;
         m->ratio_MV_outer = (m->ratio_MV_outer + ((double)nMV) / kout) / 2;
      }
      m->qmr_only = m->qmr_plus_MV_PR - m->MV_PR;
      m->gdk_plus_MV_PR = m->gdk_plus_MV + m->PR;
      if (low_res <= m->resid_0) m->jdq_sum_logResReductions += log(low_res / m->resid_0);
      m->jdq_sum_MV += nMV;
      m->jdq_conv_rate = exp(m->jdq_sum_logResReductions / m->jdq_sum_MV);
      break;
   }
   update_slowdown(m);

   m->numIt_0 = primme->stats.numOuterIterations;
   if (calledAtRestart) m->numIt_0++;
   m->numMV_0 = primme->stats.numMatvecs;
   m->timer_0 = current_time;
   m->time_in_inner = 0.0;
   m->resid_0 = currentResNorm;
   return 1;
}

#endif
static int average_over_ranks(pb_solver *S, double *ratio) {
   CHK(pb_global_sum(S, ratio, 1));
   *ratio /= (double)S->primme->numProcs;
   return 0;
}

/* book-keeping shared by both directions: expected accumulated times for the final recommendation */
static void account(pb_cost_model *m, double ratio) {
   m->accum_jdq += m->gdk_plus_MV_PR * ratio;
   m->accum_gdk += m->gdk_plus_MV_PR;
   m->accum_jdq_gdk = m->accum_jdq / m->accum_gdk;
}

static void run_gdpk(primme_params *primme, int state) {
   primme->dynamicMethodSwitch = state;
   primme->correctionParams.maxInnerIterations = 0;
   primme->correctionParams.projectors.RightX = 1;
}

static void run_jdqmr(primme_params *primme, int state) {
   primme->dynamicMethodSwitch = state;
   primme->correctionParams.maxInnerIterations = -1;
   primme->correctionParams.projectors.RightX = 0;
}

/* states: 1 GD+k (few eigenvalues, evaluated at restarts), 2 its JDQMR counterpart (evaluated every
 * iteration), 3 GD+k for many eigenvalues (evaluated when a pair converges), 4 its JDQMR counterpart */
int pb_dyn_switch_from_jdqmr(pb_solver *S, pb_cost_model *m) {
   primme_params *primme = S->primme;
   double ratio;
   if (primme->dynamicMethodSwitch == 2) {
      /* few eigenvalues: with the first timings, decide whether JDQMR can ever pay off, assuming its
       * best case -- slowdown 1.1, all the time in inner iterations (:1955-1973) */
      ratio = ratio_jdqmr_gdpk(m, 0, 1.1, 1000);
      CHK(average_over_ranks(S, &ratio));
      if (ratio > 1.05) {
         run_gdpk(primme, -1); /* GD+k for good: no further model updates */
         return 0;
      }
   }
   const int back_to = primme->dynamicMethodSwitch == 2 ? 1 : primme->dynamicMethodSwitch == 4 ? 3 : 0;
   ratio = ratio_jdqmr_gdpk(m, 0, m->JDQMR_slowdown, m->ratio_MV_outer);
   CHK(average_over_ranks(S, &ratio));
   if (ratio > 1.05) run_gdpk(primme, back_to);
   account(m, ratio);
   return 0;
}

int pb_dyn_switch_from_gdpk(pb_solver *S, pb_cost_model *m) {
   primme_params *primme = S->primme;
   /* timings without a restart are incomplete; a basis that saturates the space stays on GD (:2058-2065) */
   if (primme->stats.numRestarts == 0 ||
         primme->maxBasisSize + (primme->locking ? primme->numEvals : 0) >= primme->n)
      return 0;
   const int forward_to = primme->dynamicMethodSwitch == 1 ? 2 : primme->dynamicMethodSwitch == 3 ? 4 : 0;
   if (m->qmr_only == 0.0) { /* JDQMR never ran: take first measurements */
      run_jdqmr(primme, forward_to);
      return 0;
   }
   double ratio = ratio_jdqmr_gdpk(m, 0, m->JDQMR_slowdown, m->ratio_MV_outer);
   CHK(average_over_ranks(S, &ratio));
   if (ratio < 0.95) run_jdqmr(primme, forward_to);
   account(m, ratio);
   return 0;
}

#ifndef PB_COMPLEX
/* recommendation for future runs left in primme.dynamicMethodSwitch (main_iter.c:1221-1228) */
void pb_dyn_recommend(primme_params *primme, const pb_cost_model *m) {
   if (primme->dynamicMethodSwitch <= 0) return;
   if (m->accum_jdq_gdk < 0.96) primme->dynamicMethodSwitch = -2;      /* JDQMR_ETol */
   else if (m->accum_jdq_gdk > 1.04) primme->dynamicMethodSwitch = -1; /* GD+k */
   else primme->dynamicMethodSwitch = -3;                              /* close call: dynamic */
}
#endif

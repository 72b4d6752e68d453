/* params.c -- primme_params handling: defaults, preset methods, reflection, printing.
 *
 * Behavioural restatement of reference src/eigs/primme_interface.c (primme_initialize :101-217,
 * primme_set_method :293-531, primme_set_defaults :543-617, display :629-750, get/set/member
 * info :776-1837).  The reflective functions are generated from the single X-macro table
 * PRIMME_PARAM_TABLE in include/primme_eigs.h instead of three 90-case switches.
 */
#include "../../include/primme.h"
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

void primme_set_defaults(primme_params *primme);
void primme_display_params_prefix(const char *prefix, primme_params primme);

#define PB_MAX3(a, b, c) ((a) > (b) ? ((a) > (c) ? (a) : (c)) : ((b) > (c) ? (b) : (c)))

primme_params *primme_params_create(void) {
   primme_params *p = (primme_params *)malloc(sizeof(primme_params));
   if (p) primme_initialize(p);
   return p;
}

int primme_params_destroy(primme_params *primme) {
   free(primme);
   return 0;
}

void primme_free(primme_params *primme) { (void)primme; /* nothing is owned (ref :228-231) */ }

/* Sentinel values mean "decide later" (primme_set_method / primme_set_defaults). */
void primme_initialize(primme_params *p) {
   memset(p, 0, sizeof(*p));
   p->numEvals = 1;
   p->target = primme_smallest;
   p->numProcs = 1;
   p->nLocal = -1;
   p->locking = -1;
   p->dynamicMethodSwitch = -1;
   p->maxMatvecs = INT_MAX;
   p->maxOuterIterations = INT_MAX;
   p->restartingParams.maxPrevRetain = -1;
   p->correctionParams.precondition = -1;
   p->correctionParams.maxInnerIterations = -INT_MAX;
   p->correctionParams.convTest = primme_adaptive_ETolerance;
   p->outputFile = stdout;
   p->printLevel = 1;
   p->stats.estimateMinEVal = -HUGE_VAL;
   p->stats.estimateMaxEVal = HUGE_VAL;
   p->stats.estimateLargestSVal = -HUGE_VAL;
   p->stats.estimateBNorm = -HUGE_VAL;
   p->stats.estimateInvBNorm = -HUGE_VAL;
   for (int i = 0; i < 4; i++) p->iseed[i] = -1; /* seeded from procID at solve time */
   p->ldevecs = -1;
   p->ldOPs = -1;
   /* every *_type is primme_op_default (0), projection/initBasisMode/orth are *_default (0) */
}

static void default_prev_retain(primme_params *p) {
   if (p->restartingParams.maxPrevRetain <= 0) {
      int few = (p->maxBlockSize == 1 && p->numEvals > 1) || p->massMatrixMatvec;
      p->restartingParams.maxPrevRetain = few ? 2 : p->maxBlockSize;
   }
}

static void set_projectors(primme_params *p, int lq, int lx, int rq, int rx, int sq, int sx) {
   JD_projectors *j = &p->correctionParams.projectors;
   j->LeftQ = lq, j->LeftX = lx, j->RightQ = rq, j->RightX = rx, j->SkewQ = sq, j->SkewX = sx;
}

int primme_set_method(primme_preset_method method, primme_params *p) {
   correction_params *cp = &p->correctionParams;
   const int extremal = p->target == primme_smallest || p->target == primme_largest;

   if (method == PRIMME_DEFAULT_METHOD) method = PRIMME_DYNAMIC;
   if (method == PRIMME_DEFAULT_MIN_MATVECS) method = PRIMME_GD_Olsen_plusK;
   if (method == PRIMME_DEFAULT_MIN_TIME) method = extremal ? PRIMME_JDQMR_ETol : PRIMME_JDQMR;
   p->dynamicMethodSwitch = method == PRIMME_DYNAMIC ? 1 : 0;

   if (p->maxBlockSize == 0) p->maxBlockSize = 1;
   if (cp->precondition == -1) cp->precondition = p->applyPreconditioner ? 1 : 0;

   switch (method) {
   case PRIMME_Arnoldi:
      p->restartingParams.maxPrevRetain = 0;
      cp->precondition = 0;
      cp->maxInnerIterations = 0;
      break;
   case PRIMME_GD:
      p->restartingParams.maxPrevRetain = 0;
      cp->robustShifts = 1;
      cp->maxInnerIterations = 0;
      cp->projectors.RightX = 0;
      cp->projectors.SkewX = 0;
      break;
   case PRIMME_GD_plusK:
      default_prev_retain(p);
      cp->maxInnerIterations = 0;
      cp->projectors.RightX = 0;
      cp->projectors.SkewX = 0;
      break;
   case PRIMME_GD_Olsen_plusK:
      default_prev_retain(p);
      cp->maxInnerIterations = 0;
      cp->projectors.RightX = 1;
      cp->projectors.SkewX = 0;
      break;
   case PRIMME_JD_Olsen_plusK:
      default_prev_retain(p);
      cp->robustShifts = 1;
      cp->maxInnerIterations = 0;
      cp->projectors.RightX = 1;
      cp->projectors.SkewX = 1;
      break;
   case PRIMME_RQI:
      p->locking = 1;
      p->restartingParams.maxPrevRetain = 0;
      cp->robustShifts = 1;
      cp->maxInnerIterations = -1;
      set_projectors(p, 1, 1, 0, 1, 0, 0);
      cp->convTest = primme_full_LTolerance;
      break;
   case PRIMME_JDQR:
      p->locking = 1;
      p->restartingParams.maxPrevRetain = 1;
      cp->robustShifts = 0;
      if (cp->maxInnerIterations == -INT_MAX) cp->maxInnerIterations = 10;
      set_projectors(p, 0, 1, 1, 1, 1, 1);
      cp->relTolBase = 1.5;
      cp->convTest = primme_full_LTolerance;
      break;
   case PRIMME_JDQMR:
   case PRIMME_JDQMR_ETol:
      if (p->restartingParams.maxPrevRetain < 0) p->restartingParams.maxPrevRetain = 1;
      cp->maxInnerIterations = -1;
      set_projectors(p, cp->precondition ? 1 : 0, 1, 0, 0, 0, method == PRIMME_JDQMR ? 1 : 0);
      cp->convTest = method == PRIMME_JDQMR ? primme_adaptive : primme_adaptive_ETolerance;
      break;
   case PRIMME_STEEPEST_DESCENT:
      p->locking = 1;
      p->maxBasisSize = p->numEvals * 2;
      p->minRestartSize = p->numEvals;
      p->maxBlockSize = p->numEvals;
      p->restartingParams.maxPrevRetain = 0;
      cp->robustShifts = 0;
      cp->maxInnerIterations = 0;
      cp->projectors.RightX = 1;
      cp->projectors.SkewX = 0;
      break;
   case PRIMME_LOBPCG_OrthoBasis:
      p->maxBasisSize = p->numEvals * 3;
      p->minRestartSize = p->numEvals;
      p->maxBlockSize = p->numEvals;
      p->restartingParams.maxPrevRetain = p->numEvals;
      cp->robustShifts = 0;
      cp->maxInnerIterations = 0;
      cp->projectors.RightX = 1;
      cp->projectors.SkewX = 0;
      p->initBasisMode = primme_init_random;
      break;
   case PRIMME_LOBPCG_OrthoBasis_Window:
      if (p->maxBlockSize == 1 &&
            (p->target == primme_closest_leq || p->target == primme_closest_geq)) {
         p->maxBasisSize = 4;
         p->minRestartSize = 2;
         p->restartingParams.maxPrevRetain = 1;
      } else {
         p->maxBasisSize = p->maxBlockSize * 3;
         p->minRestartSize = p->maxBlockSize;
         p->restartingParams.maxPrevRetain = p->maxBlockSize;
      }
      cp->robustShifts = 0;
      cp->maxInnerIterations = 0;
      cp->projectors.RightX = 1;
      cp->projectors.SkewX = 0;
      p->initBasisMode = primme_init_random;
      break;
   case PRIMME_DYNAMIC:
      default_prev_retain(p);
      cp->maxInnerIterations = -1;
      set_projectors(p, cp->precondition ? 1 : 0, 1, 0, 0, 0, 0);
      cp->convTest = extremal ? primme_adaptive_ETolerance : primme_adaptive;
      break;
   default: return -1;
   }

   primme_set_defaults(p);
   return 0;
}

/* Fill every member that still carries its sentinel (reference :543-617). */
void primme_set_defaults(primme_params *p) {
   const int extremal = p->target == primme_smallest || p->target == primme_largest;
   const int keep = p->restartingParams.maxPrevRetain;

   if (p->dynamicMethodSwitch < 0) {
      primme_set_method(PRIMME_DYNAMIC, p); /* re-enters this function once */
      return;
   }
   if (p->ldevecs == -1 && p->nLocal != -1) p->ldevecs = p->nLocal;
   if (p->projectionParams.projection == primme_proj_default)
      p->projectionParams.projection = primme_proj_RR;
   if (p->initBasisMode == primme_init_default) p->initBasisMode = primme_init_krylov;

   if (p->maxBasisSize == 0) {
      /* note: the reference's "(int)2.5 * minRestartSize" casts the constant (2, resp. 1) */
      int floor_sz = extremal ? 15 : 35;
      int per_block = (extremal ? 4 : 5) * p->maxBlockSize + keep;
      int per_restart = (extremal ? 2 : 1) * p->minRestartSize + keep;
      int want = PB_MAX3(floor_sz, per_block, per_restart);
      PRIMME_INT room = p->n - p->numOrthoConst;
      p->maxBasisSize = (int)(room < want ? room : want);
   }

   if (p->minRestartSize == 0) {
      if (p->n <= 3)
         p->minRestartSize = (int)(p->n - p->numOrthoConst);
      else
         p->minRestartSize = (int)(0.5 + (extremal ? 0.4 : 0.6) * p->maxBasisSize);

      /* make the basis grow by whole blocks between restarts */
      if (p->maxBlockSize > 1) {
         int bs = p->maxBlockSize;
         int gap = p->maxBasisSize - p->minRestartSize - 1 - (keep > 0 ? keep : 0);
         int blocks = 1 + (int)(gap / (double)bs);
         p->minRestartSize = p->maxBasisSize - bs * blocks - (keep > 0 ? keep : 0);
      }
   }

   if (p->locking < 0) {
      if (!extremal)
         p->locking = 1;
      else
         p->locking = p->numEvals > p->minRestartSize ? 1 : 0;
   }
}

/* ------------------------------------------------------------------------- printing --- */
#define NAME_OF_(E) \
   case E: return #E;
static const char *target_name(int v) {
   switch (v) {
      NAME_OF_(primme_smallest)
      NAME_OF_(primme_largest)
      NAME_OF_(primme_closest_geq)
      NAME_OF_(primme_closest_leq)
      NAME_OF_(primme_closest_abs)
      NAME_OF_(primme_largest_abs)
   }
   return NULL;
}
static const char *projection_name(int v) {
   switch (v) {
      NAME_OF_(primme_proj_default)
      NAME_OF_(primme_proj_RR)
      NAME_OF_(primme_proj_harmonic)
      NAME_OF_(primme_proj_refined)
   }
   return NULL;
}
static const char *init_name(int v) {
   switch (v) {
      NAME_OF_(primme_init_default)
      NAME_OF_(primme_init_krylov)
      NAME_OF_(primme_init_random)
      NAME_OF_(primme_init_user)
   }
   return NULL;
}
static const char *convtest_name(int v) {
   switch (v) {
      NAME_OF_(primme_full_LTolerance)
      NAME_OF_(primme_decreasing_LTolerance)
      NAME_OF_(primme_adaptive_ETolerance)
      NAME_OF_(primme_adaptive)
   }
   return NULL;
}
static const char *orth_name(int v) {
   switch (v) {
      NAME_OF_(primme_orth_default)
      NAME_OF_(primme_orth_implicit_I)
      NAME_OF_(primme_orth_explicit_I)
   }
   return NULL;
}
static const char *optype_name(int v) {
   switch (v) {
      NAME_OF_(primme_op_default)
      NAME_OF_(primme_op_half)
      NAME_OF_(primme_op_float)
      NAME_OF_(primme_op_double)
      NAME_OF_(primme_op_quad)
      NAME_OF_(primme_op_int)
   }
   return NULL;
}

/* Same "prefix.member = value" lines as the reference printer (:655-750), so config files
 * written from this output are read back by the reference's test driver. */
void primme_display_params_prefix(const char *pre, primme_params primme) {
   FILE *f = primme.outputFile;
   const correction_params *cp = &primme.correctionParams;
   fprintf(f, "%s.n = %" PRIMME_INT_P "\n", pre, primme.n);
   fprintf(f, "%s.nLocal = %" PRIMME_INT_P "\n", pre, primme.nLocal);
   fprintf(f, "%s.numProcs = %d\n%s.procID = %d\n", pre, primme.numProcs, pre, primme.procID);
   fprintf(f, "\n// Output and reporting\n%s.printLevel = %d\n", pre, primme.printLevel);
   fprintf(f, "\n// Solver parameters\n%s.numEvals = %d\n", pre, primme.numEvals);
   fprintf(f, "%s.aNorm = %e\n%s.BNorm = %e\n%s.invBNorm = %e\n%s.eps = %e\n", pre, primme.aNorm,
         pre, primme.BNorm, pre, primme.invBNorm, pre, primme.eps);
   fprintf(f, "%s.maxBasisSize = %d\n%s.minRestartSize = %d\n%s.maxBlockSize = %d\n", pre,
         primme.maxBasisSize, pre, primme.minRestartSize, pre, primme.maxBlockSize);
   fprintf(f, "%s.maxOuterIterations = %" PRIMME_INT_P "\n", pre, primme.maxOuterIterations);
   fprintf(f, "%s.maxMatvecs = %" PRIMME_INT_P "\n", pre, primme.maxMatvecs);
   if (target_name(primme.target)) fprintf(f, "%s.target = %s\n", pre, target_name(primme.target));
   if (projection_name(primme.projectionParams.projection))
      fprintf(f, "%s.projection.projection = %s\n", pre,
            projection_name(primme.projectionParams.projection));
   if (init_name(primme.initBasisMode))
      fprintf(f, "%s.initBasisMode = %s\n", pre, init_name(primme.initBasisMode));
   fprintf(f, "%s.numTargetShifts = %d\n", pre, primme.numTargetShifts);
   if (primme.numTargetShifts > 0 && primme.targetShifts) {
      fprintf(f, "%s.targetShifts =", pre);
      for (int i = 0; i < primme.numTargetShifts; i++) fprintf(f, " %e", primme.targetShifts[i]);
      fprintf(f, "\n");
   }
   fprintf(f, "%s.dynamicMethodSwitch = %d\n%s.locking = %d\n%s.initSize = %d\n", pre,
         primme.dynamicMethodSwitch, pre, primme.locking, pre, primme.initSize);
   fprintf(f, "%s.numOrthoConst = %d\n", pre, primme.numOrthoConst);
   fprintf(f, "%s.ldevecs = %" PRIMME_INT_P "\n", pre, primme.ldevecs);
   fprintf(f, "%s.ldOPs = %" PRIMME_INT_P "\n", pre, primme.ldOPs);
   fprintf(f, "%s.iseed =", pre);
   for (int i = 0; i < 4; i++) fprintf(f, " %" PRIMME_INT_P, primme.iseed[i]);
   fprintf(f, "\n");
   if (primme.orth != primme_orth_default) fprintf(f, "%s.orth = %s\n", pre, orth_name(primme.orth));
   if (primme.internalPrecision != primme_op_default && optype_name(primme.internalPrecision) &&
         primme.internalPrecision != primme_op_int)
      fprintf(f, "%s.internalPrecision = %s\n", pre, optype_name(primme.internalPrecision));
   fprintf(f, "%s.restarting.maxPrevRetain = %d\n", pre, primme.restartingParams.maxPrevRetain);
   fprintf(f, "\n// Correction parameters\n");
   fprintf(f, "%s.correction.precondition = %d\n", pre, cp->precondition);
   fprintf(f, "%s.correction.robustShifts = %d\n", pre, cp->robustShifts);
   fprintf(f, "%s.correction.maxInnerIterations = %d\n", pre, cp->maxInnerIterations);
   fprintf(f, "%s.correction.relTolBase = %g\n", pre, cp->relTolBase);
   if (convtest_name(cp->convTest))
      fprintf(f, "%s.correction.convTest = %s\n", pre, convtest_name(cp->convTest));
   fprintf(f, "\n// projectors for JD cor.eq.\n");
   fprintf(f, "%s.correction.projectors.LeftQ = %d\n", pre, cp->projectors.LeftQ);
   fprintf(f, "%s.correction.projectors.LeftX = %d\n", pre, cp->projectors.LeftX);
   fprintf(f, "%s.correction.projectors.RightQ = %d\n", pre, cp->projectors.RightQ);
   fprintf(f, "%s.correction.projectors.SkewQ = %d\n", pre, cp->projectors.SkewQ);
   fprintf(f, "%s.correction.projectors.RightX = %d\n", pre, cp->projectors.RightX);
   fprintf(f, "%s.correction.projectors.SkewX = %d\n", pre, cp->projectors.SkewX);
   fprintf(f, "// ---------------------------------------------------\n");
}

void primme_display_params(primme_params primme) {
   fprintf(primme.outputFile, "// ---------------------------------------------------\n"
                              "//                 primme configuration               \n"
                              "// ---------------------------------------------------\n");
   primme_display_params_prefix("primme", primme);
   fflush(primme.outputFile);
}

/* ------------------------------------------------------------------------ reflection --- */
/* Value conventions (reference :776-1290): integer-like members travel as PRIMME_INT, doubles as
 * double; get writes pointers through *value, set takes the pointer itself as `value`. */
typedef void (*pb_anyfn)(void);

#define GET_I(path) *(PRIMME_INT *)value = (PRIMME_INT)primme->path
#define GET_D(path) *(double *)value = (double)primme->path
#define GET_P(path) *(void **)value = (void *)primme->path
#define GET_S(path) *(const char **)value = primme->path
#define GET_F(path) *(pb_anyfn *)value = (pb_anyfn)primme->path
#define GET_A4(path) \
   for (int i_ = 0; i_ < 4; i_++) ((PRIMME_INT *)value)[i_] = primme->path[i_]

int primme_get_member(primme_params *primme, primme_params_label label, void *value) {
   switch (label) {
#define X(name, id, path, kind) \
   case PRIMME_##name: GET_##kind(path); return 0;
      PRIMME_PARAM_TABLE(X)
#undef X
   default: return 1;
   }
}

/* store an integer into a member of width w (int, enum or PRIMME_INT) */
static int pb_store_int(void *dst, size_t w, PRIMME_INT v) {
   if (w == sizeof(PRIMME_INT))
      *(PRIMME_INT *)dst = v;
   else if (w == sizeof(int)) {
      if (v > INT_MAX || v < INT_MIN) return 1;
      *(int *)dst = (int)v;
   } else
      return 1;
   return 0;
}

int primme_set_member(primme_params *primme, primme_params_label label, void *value) {
   switch (label) {
#define SETK_I(path) return pb_store_int(&primme->path, sizeof(primme->path), *(PRIMME_INT *)value)
#define SETK_D(path) primme->path = *(double *)value; return 0
#define SETK_P(path) memcpy(&primme->path, &value, sizeof(void *)); return 0
#define SETK_S(path) primme->path = (const char *)value; return 0
#define SETK_F(path) memcpy(&primme->path, &value, sizeof(void *)); return 0
#define SETK_A4(path) \
   for (int i_ = 0; i_ < 4; i_++) primme->path[i_] = ((PRIMME_INT *)value)[i_]; \
   return 0
#define X(name, id, path, kind) \
   case PRIMME_##name: SETK_##kind(path);
      PRIMME_PARAM_TABLE(X)
#undef X
   default: return 1;
   }
}

int primme_member_info(
      primme_params_label *label, const char **label_name, primme_type *type, int *arity) {
   static const struct {
      int id;
      const char *name;
      char kind; /* I D P F S A */
   } tab[] = {
#define KIND_I 'I'
#define KIND_D 'D'
#define KIND_P 'P'
#define KIND_F 'F'
#define KIND_S 'S'
#define KIND_A4 'A'
#define X(name, id, path, kind) {id, #name, KIND_##kind},
         PRIMME_PARAM_TABLE(X)
#undef X
   };
   const int ntab = (int)(sizeof(tab) / sizeof(tab[0]));
   int hit = -1;
   for (int i = 0; i < ntab && hit < 0; i++) {
      if (label_name && *label_name) {
         if (strcmp(tab[i].name, *label_name) == 0) hit = i;
      } else if (label && tab[i].id == (int)*label)
         hit = i;
   }
   if (hit < 0) return 1;
   if (label) *label = (primme_params_label)tab[hit].id;
   if (label_name) *label_name = tab[hit].name;
   primme_type t = primme_pointer;
   int ar = 1;
   switch (tab[hit].kind) {
   case 'I': t = primme_int; break;
   case 'A': t = primme_int, ar = 4; break;
   case 'D': t = primme_double; break;
   case 'S': t = primme_string; break;
   default: t = primme_pointer; break;
   }
   /* arrays of doubles of run-time length (reference :1591-1595) */
   if (tab[hit].id == PRIMME_targetShifts || tab[hit].id == PRIMME_ShiftsForPreconditioner)
      t = primme_double, ar = 0;
   if (type) *type = t;
   if (arity) *arity = ar;
   return 0;
}

/* all named integer constants of the public API */
#define PB_CONSTANTS(C) \
   C(PRIMME_DEFAULT_METHOD) C(PRIMME_DYNAMIC) C(PRIMME_DEFAULT_MIN_TIME) \
   C(PRIMME_DEFAULT_MIN_MATVECS) C(PRIMME_Arnoldi) C(PRIMME_GD) C(PRIMME_GD_plusK) \
   C(PRIMME_GD_Olsen_plusK) C(PRIMME_JD_Olsen_plusK) C(PRIMME_RQI) C(PRIMME_JDQR) \
   C(PRIMME_JDQMR) C(PRIMME_JDQMR_ETol) C(PRIMME_STEEPEST_DESCENT) \
   C(PRIMME_LOBPCG_OrthoBasis) C(PRIMME_LOBPCG_OrthoBasis_Window) \
   C(primme_smallest) C(primme_largest) C(primme_closest_geq) C(primme_closest_leq) \
   C(primme_closest_abs) C(primme_largest_abs) \
   C(primme_proj_default) C(primme_proj_RR) C(primme_proj_harmonic) C(primme_proj_refined) \
   C(primme_init_default) C(primme_init_krylov) C(primme_init_random) C(primme_init_user) \
   C(primme_full_LTolerance) C(primme_decreasing_LTolerance) C(primme_adaptive_ETolerance) \
   C(primme_adaptive) \
   C(primme_event_outer_iteration) C(primme_event_inner_iteration) C(primme_event_restart) \
   C(primme_event_reset) C(primme_event_converged) C(primme_event_locked) \
   C(primme_event_message) C(primme_event_profile) \
   C(primme_orth_default) C(primme_orth_explicit_I) C(primme_orth_implicit_I) \
   C(primme_op_default) C(primme_op_quad) C(primme_op_double) C(primme_op_float) \
   C(primme_op_half) C(primme_op_int)

int primme_constant_info(const char *label_name, int *value) {
#define C(name) \
   if (strcmp(#name, label_name) == 0) { \
      *value = (int)name; \
      return 0; \
   }
   PB_CONSTANTS(C)
#undef C
   return 1;
}

int primme_enum_member_info(primme_params_label label, int *value, const char **value_name) {
   if (!value || !value_name || (*value >= 0 && *value_name) || (*value < 0 && !*value_name))
      return -1;
   const char *(*namer)(int) = NULL;
   int lo = 0, hi = -1;
   switch ((int)label) {
   case PRIMME_target: namer = target_name, hi = primme_largest_abs; break;
   case PRIMME_projectionParams_projection: namer = projection_name, hi = primme_proj_refined; break;
   case PRIMME_initBasisMode: namer = init_name, hi = primme_init_user; break;
   case PRIMME_correctionParams_convTest: namer = convtest_name, hi = primme_adaptive; break;
   case PRIMME_orth: namer = orth_name, hi = primme_orth_explicit_I; break;
   case PRIMME_matrixMatvec_type:
   case PRIMME_applyPreconditioner_type:
   case PRIMME_globalSumReal_type:
   case PRIMME_broadcastReal_type:
   case PRIMME_massMatrixMatvec_type: namer = optype_name, hi = primme_op_int; break;
   case PRIMME_commInfo: {
      /* the reference's "hack": the label of commInfo stands for the preset-method enum */
      static const char *const methods[] = {"PRIMME_DEFAULT_METHOD", "PRIMME_DYNAMIC",
            "PRIMME_DEFAULT_MIN_TIME", "PRIMME_DEFAULT_MIN_MATVECS", "PRIMME_Arnoldi",
            "PRIMME_GD", "PRIMME_GD_plusK", "PRIMME_GD_Olsen_plusK", "PRIMME_JD_Olsen_plusK",
            "PRIMME_RQI", "PRIMME_JDQR", "PRIMME_JDQMR", "PRIMME_JDQMR_ETol",
            "PRIMME_STEEPEST_DESCENT", "PRIMME_LOBPCG_OrthoBasis",
            "PRIMME_LOBPCG_OrthoBasis_Window"};
      for (int i = 0; i < 16; i++)
         if (*value == i || (*value_name && strcmp(methods[i], *value_name) == 0)) {
            *value = i, *value_name = methods[i];
            return 0;
         }
      return -2;
   }
   default: return -2;
   }
   for (int v = lo; v <= hi; v++) {
      const char *nm = namer(v);
      if (!nm) continue;
      if (*value == v || (*value_name && strcmp(nm, *value_name) == 0)) {
         *value = v, *value_name = nm;
         return 0;
      }
   }
   return -2;
}

/* svds.c -- singular value front end, normal-equations path.
 *
 * dprimme_svds / cublas_dprimme_svds with primme_svds_normalequations: the singular triplets of A
 * (m x n) are the eigenpairs of A'A (n <= m) or AA' (n > m); the eigenproblem runs through the
 * same Davidson hot path as dprimme with a matvec that applies the user's operator twice.
 * Restates, for one stage and fp64:
 *    primme_svds_initialize / set_method / set_defaults   src/svds/primme_svds_interface.c:107-420
 *    wrapper_svds                                         src/svds/primme_svds_c.c:388-540
 *    copy_last_params_from_svds / _to_svds                :551-1000
 *    matrixMatvec_eigs (A'A, AA')                         :1323-1383
 *    convTestFunATA / default_convTestFun                 :1594-1690
 * Out of scope here (refused with PRIMME_FUNCTION_UNAVAILABLE = -44, like a reference build
 * without the feature): the augmented operator and therefore primme_svds_hybrid / the second
 * stage (needs refined extraction and JDQMR), primme_svds_closest_abs, preconditioning, every
 * precision but double.
 */
#include "pb_host.h"
#include "../../include/primme_svds.h"
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ parameter interface ---- */
primme_svds_params *primme_svds_params_create(void) {
   primme_svds_params *p = (primme_svds_params *)malloc(sizeof(*p));
   if (p) primme_svds_initialize(p);
   return p;
}

int primme_svds_params_destroy(primme_svds_params *primme_svds) {
   free(primme_svds);
   return 0;
}

void primme_svds_initialize(primme_svds_params *s) {
   memset(s, 0, sizeof(*s));
   s->numSvals = 1;
   s->target = primme_svds_largest;
   s->method = primme_svds_op_none, s->methodStage2 = primme_svds_op_none;
   s->numProcs = 1;
   s->mLocal = -1, s->nLocal = -1;
   s->globalSumReal_type = s->broadcastReal_type = primme_op_default;
   s->internalPrecision = primme_op_default;
   s->matrixMatvec_type = s->applyPreconditioner_type = primme_op_default;
   s->precondition = -1;
   s->maxMatvecs = INT_MAX;
   s->printLevel = 1;
   s->outputFile = stdout;
   s->locking = -1;
   for (int i = 0; i < 4; i++) s->iseed[i] = -1;
   s->convTestFun_type = s->monitorFun_type = primme_op_default;
   primme_initialize(&s->primme);
   primme_initialize(&s->primmeStage2);
}

void primme_svds_free(primme_svds_params *primme_svds) {
   /* nothing is kept between solves (interface.c:527-531) */
   (void)primme_svds;
}

/* collectives of the eigensolver forwarded to the SVD callbacks (interface.c:1250-1290) */
static void global_sum_svds(void *sendBuf, void *recvBuf, int *count, primme_params *primme, int *ierr) {
   primme_svds_params *s = (primme_svds_params *)primme->matrix;
   s->globalSumReal(sendBuf, recvBuf, count, s, ierr);
}
static void broadcast_svds(void *buffer, int *count, primme_params *primme, int *ierr) {
   primme_svds_params *s = (primme_svds_params *)primme->matrix;
   s->broadcastReal(buffer, count, s, ierr);
}

/* options of primme_svds handed to the eigensolver of one stage (interface.c:296-408) */
static void copy_params_from_svds(primme_svds_params *s, int stage) {
   primme_params *primme = stage == 0 ? &s->primme : &s->primmeStage2;
   const primme_svds_operator method = stage == 0 ? s->method : s->methodStage2;
   if (method == primme_svds_op_none) {
      primme->maxMatvecs = 1;
      return;
   }
   primme->numEvals = s->numSvals;
   if (s->aNorm > 0.0)
      primme->aNorm = method == primme_svds_op_augmented ? s->aNorm * sqrt(2.0) : s->aNorm * s->aNorm;
   primme->eps = s->eps;
   primme->initSize = s->initSize;
   if (s->maxBasisSize > 0) primme->maxBasisSize = s->maxBasisSize;
   if (s->maxBlockSize > 0) primme->maxBlockSize = s->maxBlockSize;
   primme->maxMatvecs = s->maxMatvecs;
   primme->printLevel = s->printLevel;
   primme->outputFile = s->outputFile;
   primme->numOrthoConst = s->numOrthoConst;
   if (s->numProcs > 1) {
      primme->procID = s->procID;
      primme->numProcs = s->numProcs;
      primme->commInfo = s->commInfo;
   }
   if (s->globalSumReal) primme->globalSumReal = global_sum_svds;
   if (s->broadcastReal) primme->broadcastReal = broadcast_svds;
   switch (method) {
   case primme_svds_op_AtA:
      primme->n = s->n;
      if (primme->nLocal == -1 && s->nLocal != -1) primme->nLocal = s->nLocal;
      break;
   case primme_svds_op_AAt:
      primme->n = s->m;
      if (primme->nLocal == -1 && s->mLocal != -1) primme->nLocal = s->mLocal;
      break;
   default:
      primme->n = s->m + s->n;
      if (primme->nLocal == -1 && s->mLocal != -1 && s->nLocal != -1) primme->nLocal = s->mLocal + s->nLocal;
      break;
   }
   switch (s->target) {
   case primme_svds_largest: primme->target = primme_largest; break;
   case primme_svds_smallest:
      primme->target = method == primme_svds_op_augmented ? primme_closest_geq : primme_smallest;
      break;
   default:
      primme->target = primme_closest_abs;
      primme->numTargetShifts = s->numTargetShifts;
      break;
   }
   if (stage == 1 && primme->initBasisMode == primme_init_default) primme->initBasisMode = primme_init_user;
   if (((method == primme_svds_op_augmented && s->target != primme_svds_largest) ||
             s->target == primme_svds_closest_abs) &&
         primme->projectionParams.projection == primme_proj_default)
      primme->projectionParams.projection = primme_proj_refined;
   if (s->locking >= 0) primme->locking = s->locking;
   if (s->precondition >= 0)
      primme->correctionParams.precondition = s->precondition;
   else if (primme->correctionParams.precondition < 0)
      primme->correctionParams.precondition = s->applyPreconditioner ? 1 : 0;
}

static void svds_set_defaults(primme_svds_params *s);

int primme_svds_set_method(primme_svds_preset_method method, primme_preset_method methodStage1,
      primme_preset_method methodStage2, primme_svds_params *s) {
   switch (method) {
   case primme_svds_default:
   case primme_svds_hybrid:
      s->method = s->n <= s->m ? primme_svds_op_AtA : primme_svds_op_AAt;
      s->methodStage2 = primme_svds_op_augmented;
      break;
   case primme_svds_normalequations:
      s->method = s->n <= s->m ? primme_svds_op_AtA : primme_svds_op_AAt;
      s->methodStage2 = primme_svds_op_none;
      break;
   case primme_svds_augmented:
      s->method = primme_svds_op_augmented;
      s->methodStage2 = primme_svds_op_none;
      break;
   }
   svds_set_defaults(s);
   primme_set_method(methodStage1, &s->primme);
   if (methodStage2 == PRIMME_DEFAULT_METHOD && s->target != primme_svds_largest) methodStage2 = PRIMME_JDQMR;
   if (s->methodStage2 != primme_svds_op_none) primme_set_method(methodStage2, &s->primmeStage2);
   return 0;
}

/* interface.c:268-282 */
static void svds_set_defaults(primme_svds_params *s) {
   if (s->method == primme_svds_op_none) {
      primme_svds_set_method(primme_svds_default, PRIMME_DEFAULT_METHOD, PRIMME_DEFAULT_METHOD, s);
      return; /* set_method came back through here with the method set */
   }
   copy_params_from_svds(s, 0);
   if (s->methodStage2 != primme_svds_op_none) copy_params_from_svds(s, 1);
}

void primme_svds_display_params(primme_svds_params s) {
   FILE *f = s.outputFile ? s.outputFile : stdout;
   fprintf(f, "// ---------------------------------------------------\n");
   fprintf(f, "//            primme_svds configuration               \n");
   fprintf(f, "// ---------------------------------------------------\n");
   fprintf(f, "primme_svds.m = %" PRIMME_INT_P "\n", s.m);
   fprintf(f, "primme_svds.n = %" PRIMME_INT_P "\n", s.n);
   fprintf(f, "primme_svds.mLocal = %" PRIMME_INT_P "\n", s.mLocal);
   fprintf(f, "primme_svds.nLocal = %" PRIMME_INT_P "\n", s.nLocal);
   fprintf(f, "primme_svds.numProcs = %d\n", s.numProcs);
   fprintf(f, "primme_svds.procID = %d\n", s.procID);
   fprintf(f, "primme_svds.numSvals = %d\n", s.numSvals);
   fprintf(f, "primme_svds.aNorm = %e\n", s.aNorm);
   fprintf(f, "primme_svds.eps = %e\n", s.eps);
   fprintf(f, "primme_svds.maxBasisSize = %d\n", s.maxBasisSize);
   fprintf(f, "primme_svds.maxBlockSize = %d\n", s.maxBlockSize);
   fprintf(f, "primme_svds.maxMatvecs = %" PRIMME_INT_P "\n", s.maxMatvecs);
   fprintf(f, "primme_svds.target = %s\n", s.target == primme_svds_largest    ? "primme_svds_largest"
                                            : s.target == primme_svds_smallest ? "primme_svds_smallest"
                                                                               : "primme_svds_closest_abs");
   fprintf(f, "primme_svds.numTargetShifts = %d\n", s.numTargetShifts);
   fprintf(f, "primme_svds.locking = %d\n", s.locking);
   fprintf(f, "primme_svds.initSize = %d\n", s.initSize);
   fprintf(f, "primme_svds.numOrthoConst = %d\n", s.numOrthoConst);
   fprintf(f, "primme_svds.printLevel = %d\n", s.printLevel);
   static const char *ops[] = {"primme_svds_op_none", "primme_svds_op_AtA", "primme_svds_op_AAt", "primme_svds_op_augmented"};
   fprintf(f, "primme_svds.method = %s\n", ops[s.method & 3]);
   fprintf(f, "primme_svds.methodStage2 = %s\n", ops[s.methodStage2 & 3]);
   if (s.method != primme_svds_op_none) primme_display_params_prefix("primme", s.primme);
   if (s.methodStage2 != primme_svds_op_none) primme_display_params_prefix("primmeStage2", s.primmeStage2);
   fflush(f);
}

typedef void (*pb_svds_anyfn)(void);
int primme_svds_get_member(primme_svds_params *p, primme_svds_params_label label, void *value) {
   switch (label) {
#define SG_I(path) *(PRIMME_INT *)value = (PRIMME_INT)p->path
#define SG_D(path) *(double *)value = (double)p->path
#define SG_P(path) *(void **)value = (void *)p->path
#define SG_S(path) *(const char **)value = p->path
#define SG_F(path) *(pb_svds_anyfn *)value = (pb_svds_anyfn)p->path
#define SG_Z(path) *(void **)value = (void *)&p->path
#define SG_A4(path) \
   for (int i_ = 0; i_ < 4; i_++) ((PRIMME_INT *)value)[i_] = p->path[i_]
#define X(name, id, path, kind) \
   case PRIMME_SVDS_##name: SG_##kind(path); return 0;
      PRIMME_SVDS_PARAM_TABLE(X)
#undef X
   default: return 1;
   }
}

static int svds_store_int(void *dst, size_t w, PRIMME_INT v) {
   if (w == sizeof(PRIMME_INT))
      *(PRIMME_INT *)dst = v;
   else if (w == sizeof(int)) {
      if (v > INT_MAX || v < INT_MIN) return 1;
      *(int *)dst = (int)v;
   } else
      return 1;
   return 0;
}

int primme_svds_set_member(primme_svds_params *p, primme_svds_params_label label, void *value) {
   switch (label) {
#define SS_I(path) return svds_store_int(&p->path, sizeof(p->path), *(PRIMME_INT *)value)
#define SS_D(path) p->path = *(double *)value; return 0
#define SS_P(path) memcpy(&p->path, &value, sizeof(void *)); return 0
#define SS_S(path) p->path = (const char *)value; return 0
#define SS_F(path) memcpy(&p->path, &value, sizeof(void *)); return 0
#define SS_Z(path) p->path = *(primme_params *)value; return 0
#define SS_A4(path) \
   for (int i_ = 0; i_ < 4; i_++) p->path[i_] = ((PRIMME_INT *)value)[i_]; \
   return 0
#define X(name, id, path, kind) \
   case PRIMME_SVDS_##name: SS_##kind(path);
      PRIMME_SVDS_PARAM_TABLE(X)
#undef X
   default: return 1;
   }
}

int primme_svds_member_info(primme_svds_params_label *label, const char **label_name, primme_type *type, int *arity) {
   static const struct {
      int id;
      const char *name;
      char kind;
   } tab[] = {
#define SK_I 'I'
#define SK_D 'D'
#define SK_P 'P'
#define SK_F 'F'
#define SK_S 'S'
#define SK_Z 'P'
#define SK_A4 'A'
#define X(name, id, path, kind) {id, #name, SK_##kind},
         PRIMME_SVDS_PARAM_TABLE(X)
#undef X
   };
   int hit = -1;
   for (int i = 0; i < (int)(sizeof(tab) / sizeof(tab[0])) && hit < 0; i++) {
      if (label_name && *label_name) {
         if (strcmp(tab[i].name, *label_name) == 0) hit = i;
      } else if (label && tab[i].id == (int)*label)
         hit = i;
   }
   if (hit < 0) return 1;
   if (label) *label = (primme_svds_params_label)tab[hit].id;
   if (label_name) *label_name = tab[hit].name;
   primme_type t = primme_pointer;
   int ar = 1;
   switch (tab[hit].kind) {
   case 'I': t = primme_int; break;
   case 'A': t = primme_int, ar = 4; break;
   case 'D': t = primme_double; break;
   case 'S': t = primme_string; break;
   default: break;
   }
   if (tab[hit].id == PRIMME_SVDS_targetShifts) t = primme_double, ar = 0;
   if (type) *type = t;
   if (arity) *arity = ar;
   return 0;
}

int primme_svds_constant_info(const char *label_name, int *value) {
   static const struct {
      const char *name;
      int v;
   } tab[] = {{"primme_svds_largest", primme_svds_largest}, {"primme_svds_smallest", primme_svds_smallest},
         {"primme_svds_closest_abs", primme_svds_closest_abs}, {"primme_svds_default", primme_svds_default},
         {"primme_svds_hybrid", primme_svds_hybrid}, {"primme_svds_normalequations", primme_svds_normalequations},
         {"primme_svds_augmented", primme_svds_augmented}, {"primme_svds_op_none", primme_svds_op_none},
         {"primme_svds_op_AtA", primme_svds_op_AtA}, {"primme_svds_op_AAt", primme_svds_op_AAt},
         {"primme_svds_op_augmented", primme_svds_op_augmented}};
   for (size_t i = 0; i < sizeof(tab) / sizeof(tab[0]); i++)
      if (strcmp(tab[i].name, label_name) == 0) {
         *value = tab[i].v;
         return 0;
      }
   return primme_constant_info(label_name, value);
}

int primme_svds_enum_member_info(primme_svds_params_label label, int *value, const char **value_name) {
   static const struct {
      int label;
      const char *name;
      int v;
   } tab[] = {{PRIMME_SVDS_target, "primme_svds_largest", primme_svds_largest},
         {PRIMME_SVDS_target, "primme_svds_smallest", primme_svds_smallest},
         {PRIMME_SVDS_target, "primme_svds_closest_abs", primme_svds_closest_abs},
         {PRIMME_SVDS_method, "primme_svds_op_none", primme_svds_op_none},
         {PRIMME_SVDS_method, "primme_svds_op_AtA", primme_svds_op_AtA},
         {PRIMME_SVDS_method, "primme_svds_op_AAt", primme_svds_op_AAt},
         {PRIMME_SVDS_method, "primme_svds_op_augmented", primme_svds_op_augmented},
         {PRIMME_SVDS_methodStage2, "primme_svds_op_none", primme_svds_op_none},
         {PRIMME_SVDS_methodStage2, "primme_svds_op_AtA", primme_svds_op_AtA},
         {PRIMME_SVDS_methodStage2, "primme_svds_op_AAt", primme_svds_op_AAt},
         {PRIMME_SVDS_methodStage2, "primme_svds_op_augmented", primme_svds_op_augmented}};
   for (size_t i = 0; i < sizeof(tab) / sizeof(tab[0]); i++) {
      if (tab[i].label != (int)label) continue;
      if (value_name && *value_name) {
         if (strcmp(tab[i].name, *value_name) == 0) {
            if (value) *value = tab[i].v;
            return 0;
         }
      } else if (value && tab[i].v == *value) {
         if (value_name) *value_name = tab[i].name;
         return 0;
      }
   }
   return 1;
}

/* ------------------------------------------------------------------------ the solver ---- */
/* state of a running SVD solve, found from the eigensolver's callbacks through primme->matrix */
typedef struct svds_run {
   primme_svds_params *svds;
   int device_mode; /* callbacks and svecs live in device memory (cublas_dprimme_svds) */
   double *aux;     /* m x maxBlockSize (A'A) or n x maxBlockSize (AA') intermediate block */
   int64_t aux_rows;
   int aux_cols;
} svds_run;
#define SVDS_MAX_RUNS 16
static svds_run runs[SVDS_MAX_RUNS];

static svds_run *find_run(const primme_svds_params *s) {
   for (int i = 0; i < SVDS_MAX_RUNS; i++)
      if (runs[i].svds == s) return &runs[i];
   return NULL;
}

static int call_svds_matvec(primme_svds_params *s, double *x, PRIMME_INT ldx, double *y, PRIMME_INT ldy,
      int bs, int trans) {
   int ierr = 0;
   s->matrixMatvec(x, &ldx, y, &ldy, &bs, &trans, s, &ierr);
   return ierr;
}

/* y = A'(A x) or A(A' x) in blocks of maxBlockSize columns (primme_svds_c.c:1337-1371) */
static void matvec_normal_equations(void *x_, PRIMME_INT *ldx, void *y_, PRIMME_INT *ldy, int *blockSize,
      primme_params *primme, int *ierr) {
   primme_svds_params *s = (primme_svds_params *)primme->matrix;
   svds_run *run = find_run(s);
   double *x = (double *)x_, *y = (double *)y_;
   *ierr = 1;
   if (!run) return;
   const int ata = s->method == primme_svds_op_AtA;
   const PRIMME_INT rows = ata ? s->mLocal : s->nLocal;
   const int cap = PB_MAX(1, PB_MIN(primme->maxBlockSize, *blockSize));
   if (!run->aux || run->aux_cols < cap || run->aux_rows < rows) {
      pb200_ctx *ctx = primme_b200_solver_ctx(primme);
      if (run->aux) {
         if (run->device_mode) pb200_free(ctx, run->aux);
         else free(run->aux);
         run->aux = NULL;
      }
      const size_t bytes = sizeof(double) * (size_t)PB_MAX(rows, 1) * cap;
      if (run->device_mode) {
         if (!ctx || pb200_malloc(ctx, bytes, (void **)&run->aux)) return;
      } else if (!(run->aux = (double *)malloc(bytes)))
         return;
      run->aux_cols = cap, run->aux_rows = rows;
   }
   for (int i = 0; i < *blockSize; i += cap) {
      const int bs = PB_MIN(cap, *blockSize - i);
      int e = call_svds_matvec(s, x + (size_t)*ldx * i, *ldx, run->aux, PB_MAX(rows, 1), bs, ata ? 0 : 1);
      if (!e) e = call_svds_matvec(s, run->aux, PB_MAX(rows, 1), y + (size_t)*ldy * i, *ldy, bs, ata ? 1 : 0);
      if (e) {
         *ierr = e;
         return;
      }
   }
   *ierr = 0;
}

/* primme_svds_c.c:1594-1620 (the augmented re-check does not apply to the normal equations) */
static void default_conv_test_svds(double *sval, void *leftsvec, void *rightsvec, double *rNorm, int *method,
      int *isConv, primme_svds_params *s, int *ierr) {
   (void)sval, (void)leftsvec, (void)rightsvec, (void)method;
   *isConv = *rNorm < PB_MAX(s->eps, PB_EPS * 3.16) * s->aNorm;
   *ierr = 0;
}

/* convergence of an eigenpair of A'A / AA' judged as a singular triplet (primme_svds_c.c:1640-1690) */
static void conv_test_normal_equations(double *eval, void *evec, double *rNorm, int *isConv, primme_params *primme,
      int *ierr) {
   primme_svds_params *s = (primme_svds_params *)primme->matrix;
   const double aNorm = primme->aNorm > 0.0 ? primme->aNorm : primme->stats.estimateLargestSVal;
   const double maxaNorm = PB_MAX(primme->aNorm, primme->stats.estimateLargestSVal);
   if (rNorm && *rNorm < PB_EPS * maxaNorm * 3.16) {
      *isConv = 1, *ierr = 0;
      return;
   }
   const double oldaNorm = s->aNorm;
   if (s->aNorm <= 0.0) s->aNorm = sqrt(aNorm);
   double sval = eval ? sqrt(fabs(*eval)) : 0.0;
   double srNorm = (rNorm && eval) ? *rNorm / sval : 0.0;
   int method = (int)s->method;
   const int aat = s->method == primme_svds_op_AAt;
   s->convTestFun(&sval, aat ? evec : NULL, aat ? NULL : evec, &srNorm, &method, isConv, s, ierr);
   s->aNorm = oldaNorm;
}

static int svds_check_input(void *svals, void *svecs, void *resNorms, primme_svds_params *s) {
   if (!s) return -4;
   if (s->n < 0 || s->m < 0 || s->nLocal < 0 || s->mLocal < 0 || s->nLocal > s->n || s->mLocal > s->m) return -5;
   if (s->numProcs < 1) return -6;
   if (!s->matrixMatvec) return -7;
   if (!s->applyPreconditioner && s->precondition == 1) return -8;
   if (s->numProcs > 1 && !s->globalSumReal) return -9;
   if (s->numSvals > PB_MIN(s->n, s->m)) return -10;
   if (s->numSvals < 1) return -11;
   if (s->target != primme_svds_smallest && s->target != primme_svds_largest && s->target != primme_svds_closest_abs)
      return -13;
   if (s->method != primme_svds_op_AtA && s->method != primme_svds_op_AAt && s->method != primme_svds_op_augmented)
      return -14;
   if ((s->method == primme_svds_op_augmented && s->methodStage2 != primme_svds_op_none) ||
         (s->method != primme_svds_op_augmented && s->methodStage2 != primme_svds_op_augmented &&
               s->methodStage2 != primme_svds_op_none))
      return -15;
   if (s->printLevel < 0 || s->printLevel > 5) return -16;
   if (!svals) return -17;
   if (!svecs) return -18;
   if (!resNorms) return -19;
   return 0;
}

static int svds_out_of_scope(primme_svds_params *s) {
   const char *why = NULL;
   if (s->method == primme_svds_op_augmented || s->methodStage2 != primme_svds_op_none)
      why = "the augmented operator / second stage (primme_svds_hybrid, primme_svds_augmented)";
   else if (s->target == primme_svds_closest_abs)
      why = "primme_svds_closest_abs (needs refined extraction)";
   else if (s->applyPreconditioner || s->precondition > 0)
      why = "preconditioning of the normal equations";
   else if ((s->matrixMatvec_type != primme_op_default && s->matrixMatvec_type != primme_op_double) ||
            (s->internalPrecision != primme_op_default && s->internalPrecision != primme_op_double))
      why = "callbacks / internal precision other than double";
   else if (s->numProcs > 1)
      why = "the distributed SVD front end";
   if (!why) return 0;
   if (s->outputFile && s->printLevel > 0)
      fprintf(s->outputFile, "PRIMME-B200: %s is outside the scope of this build\n", why);
   return PRIMME_FUNCTION_UNAVAILABLE;
}

/* device or host copy / scale helpers for the post-processing of the vectors */
static int vec_copy(pb200_ctx *ctx, int dev, const double *src, int64_t lds, double *dst, int64_t ldd, int64_t rows,
      int cols) {
   if (rows <= 0 || cols <= 0 || src == dst) return 0;
   if (dev) return pb200_copy_d2d(ctx, src, lds, dst, ldd, rows, cols, 8);
   /* columns may overlap when packing to the left: move column by column in increasing order */
   for (int j = 0; j < cols; j++) memmove(dst + (size_t)ldd * j, src + (size_t)lds * j, sizeof(double) * rows);
   return 0;
}

static int svds_solve(double *svals, double *svecs, double *resNorms, primme_svds_params *s, int device_mode) {
   if (!s) return -4;
   /* defaults of a sequential run (primme_svds_c.c:402-409) */
   if (s->numProcs <= 1 && svals && svecs && resNorms) {
      s->mLocal = s->m, s->nLocal = s->n;
      s->procID = 0, s->numProcs = 1;
   }
   svds_set_defaults(s);
   if (!svals && !svecs && !resNorms) return 0;
   int rc = svds_check_input(svals, svecs, resNorms, s);
   if (rc) return rc;
   rc = svds_out_of_scope(s);
   if (rc) {
      s->initSize = 0;
      return rc;
   }
   if (!s->convTestFun) {
      s->convTestFun = default_conv_test_svds;
      s->convTestFun_type = primme_op_double;
      if (s->eps == 0.0) s->eps = PB_EPS * 1e4; /* after set_defaults, as in the reference (:427-434) */
   }
   memset(&s->stats, 0, sizeof(s->stats));

   primme_params *primme = &s->primme;
   const int ata = s->method == primme_svds_op_AtA;
   const PRIMME_INT mL = s->mLocal, nL = s->nLocal;
   const int nMax = PB_MAX(s->initSize, s->numSvals) + s->numOrthoConst;
   int n0 = s->initSize + s->numOrthoConst;

   svds_run *run = find_run(NULL);
   if (!run) return PRIMME_MALLOC_FAILURE;
   memset(run, 0, sizeof(*run));
   run->svds = s, run->device_mode = device_mode;

   /* ---- copy_last_params_from_svds, stage 0 (primme_svds_c.c:551-830) ---- */
   if (!primme->matrixMatvec) {
      primme->matrixMatvec = matvec_normal_equations;
      primme->matrixMatvec_type = primme_op_double;
      primme->matrix = s;
   }
   if (s->aNorm > 0.0) primme->aNorm = s->aNorm * s->aNorm;
   primme->convTestFun = conv_test_normal_equations;
   primme->convTestFun_type = primme_op_double;
   primme->initSize = s->initSize;
   primme->numOrthoConst = s->numOrthoConst;
   /* a private context for the whole call when the caller attached none: the vector
    * post-processing below needs it after the eigensolver returns */
   pb200_ctx *ctx = primme_b200_attached_ctx(primme);
   int own_ctx = 0;
   if (!ctx && device_mode) {
      if (pb200_ctx_create(&ctx, -1)) {
         run->svds = NULL;
         return PRIMME_FUNCTION_UNAVAILABLE;
      }
      own_ctx = 1;
      primme_b200_attach_ctx(primme, ctx);
   }
   /* the right vectors [Vc V0] move to the rightmost position of svecs, where the eigensolver
    * works in place; with AA' only Vc moves and the eigensolver works on the left block */
   double *aux = svecs + (size_t)nMax * mL;
   rc = vec_copy(ctx, device_mode, svecs + (size_t)mL * n0, nL, aux, nL, nL, ata ? n0 : s->numOrthoConst);
   double *evecs = ata ? aux : svecs;
   for (int i = 0; i < 4; i++) primme->iseed[i] = s->iseed[i];
   primme->maxMatvecs = s->maxMatvecs / 2;
   if (s->locking >= 0) primme->locking = s->locking;
   primme->queue = s->queue;
   primme->profile = s->profile;
   primme->ldevecs = ata ? nL : mL;

   int ret = rc;
   if (!ret) ret = device_mode ? cublas_dprimme(svals, evecs, resNorms, primme) : dprimme(svals, evecs, resNorms, primme);

   /* ---- copy_last_params_to_svds, stage 0 (primme_svds_c.c:838-1000) ---- */
   s->stats.numOuterIterations += primme->stats.numOuterIterations;
   s->stats.numRestarts += primme->stats.numRestarts;
   s->stats.numMatvecs += primme->stats.numMatvecs * 2; /* every eigensolver matvec is A and A' (:56-58) */
   s->stats.numPreconds += primme->stats.numPreconds;
   s->stats.numGlobalSum += primme->stats.numGlobalSum;
   s->stats.volumeGlobalSum += primme->stats.volumeGlobalSum;
   s->stats.numBroadcast += primme->stats.numBroadcast;
   s->stats.volumeBroadcast += primme->stats.volumeBroadcast;
   s->stats.numOrthoInnerProds += primme->stats.numOrthoInnerProds;
   s->stats.elapsedTime += primme->stats.elapsedTime;
   s->stats.timeMatvec += primme->stats.timeMatvec;
   s->stats.timePrecond += primme->stats.timePrecond;
   s->stats.timeOrtho += primme->stats.timeOrtho;
   s->stats.timeGlobalSum += primme->stats.timeGlobalSum;
   s->stats.timeBroadcast += primme->stats.timeBroadcast;
   s->stats.lockingIssue += primme->stats.lockingIssue;
   if (primme->aNorm > 0.0) s->aNorm = sqrt(primme->aNorm);
   const int nconv = primme->initSize > 0 ? primme->initSize : 0;
   for (int i = 0; i < nconv; i++) svals[i] = sqrt(PB_MAX(0.0, svals[i]));
   s->initSize = nconv;
   n0 = s->initSize + s->numOrthoConst;
   int rc2 = 0;
   if (nconv > 0) {
      double *inv = (double *)malloc(sizeof(double) * nconv);
      for (int i = 0; i < nconv; i++) inv[i] = 1.0 / svals[i];
      if (ata) {
         /* U = A V diag(1/sigma), then V packed right after the n0 left vectors */
         double *U = svecs + (size_t)mL * s->numOrthoConst;
         double *V = aux + (size_t)nL * s->numOrthoConst;
         rc2 = call_svds_matvec(s, V, nL, U, mL, nconv, 0) ? PRIMME_USER_FAILURE : 0;
         s->stats.numMatvecs += nconv;
         if (!rc2) {
            if (device_mode)
               rc2 = pb200_dscale_columns(ctx, mL, inv, U, mL, nconv);
            else
               for (int j = 0; j < nconv; j++)
                  for (PRIMME_INT r = 0; r < mL; r++) U[r + (size_t)mL * j] *= inv[j];
         }
         if (!rc2) rc2 = vec_copy(ctx, device_mode, aux, nL, svecs + (size_t)mL * n0, nL, nL, n0);
      } else {
         /* the constraints Vc first, then V = A' U diag(1/sigma) */
         rc2 = vec_copy(ctx, device_mode, aux, nL, svecs + (size_t)mL * n0, nL, nL, s->numOrthoConst);
         double *U = svecs + (size_t)mL * s->numOrthoConst;
         double *V = svecs + (size_t)mL * n0 + (size_t)nL * s->numOrthoConst;
         if (!rc2) rc2 = call_svds_matvec(s, U, mL, V, nL, nconv, 1) ? PRIMME_USER_FAILURE : 0;
         s->stats.numMatvecs += nconv;
         if (!rc2) {
            if (device_mode)
               rc2 = pb200_dscale_columns(ctx, nL, inv, V, nL, nconv);
            else
               for (int j = 0; j < nconv; j++)
                  for (PRIMME_INT r = 0; r < nL; r++) V[r + (size_t)nL * j] *= inv[j];
         }
      }
      free(inv);
      if (device_mode && ctx) pb200_ctx_sync(ctx);
   }
   for (int i = 0; i < 4; i++) s->iseed[i] = primme->iseed[i];
   for (int i = 0; i < nconv; i++) resNorms[i] = PB_MIN(resNorms[i] / svals[i], s->aNorm);

   if (run->aux) {
      if (device_mode) pb200_free(ctx, run->aux);
      else free(run->aux);
   }
   run->svds = NULL, run->aux = NULL;
   if (own_ctx) {
      primme_b200_attach_ctx(primme, NULL);
      pb200_ctx_destroy(ctx);
   }
   if (ret != 0) return ret - 100; /* errors of the first stage (primme_svds_c.c:497) */
   return rc2;
}

int dprimme_svds(double *svals, double *svecs, double *resNorms, primme_svds_params *primme_svds) {
   return svds_solve(svals, svecs, resNorms, primme_svds, 0);
}

int cublas_dprimme_svds(double *svals, double *svecs, double *resNorms, primme_svds_params *primme_svds) {
   return svds_solve(svals, svecs, resNorms, primme_svds, 1);
}

/* every other precision / back end of the SVD front end: not built (primme_svds_c.c:219-260 with the
 * type disabled) */
#define SVDS_UNAVAILABLE(name, SV, VEC)                                                        \
   int name(SV *svals, VEC *svecs, SV *resNorms, primme_svds_params *primme_svds) {           \
      (void)svals, (void)svecs, (void)resNorms;                                                \
      if (primme_svds) primme_svds->initSize = 0;                                              \
      return PRIMME_FUNCTION_UNAVAILABLE;                                                      \
   }
#define SVDS_UNAVAILABLE_ALL(name, SV, VEC) \
   SVDS_UNAVAILABLE(name, SV, VEC) SVDS_UNAVAILABLE(magma_##name, SV, VEC) SVDS_UNAVAILABLE(cublas_##name, SV, VEC)
SVDS_UNAVAILABLE_ALL(hprimme_svds, PRIMME_HALF, PRIMME_HALF)
SVDS_UNAVAILABLE_ALL(kprimme_svds, PRIMME_HALF, PRIMME_COMPLEX_HALF)
SVDS_UNAVAILABLE_ALL(sprimme_svds, float, float)
SVDS_UNAVAILABLE_ALL(cprimme_svds, float, PRIMME_COMPLEX_FLOAT)
SVDS_UNAVAILABLE_ALL(zprimme_svds, double, PRIMME_COMPLEX_DOUBLE)
SVDS_UNAVAILABLE_ALL(hsprimme_svds, float, PRIMME_HALF)
SVDS_UNAVAILABLE_ALL(ksprimme_svds, float, PRIMME_COMPLEX_HALF)
SVDS_UNAVAILABLE(magma_dprimme_svds, double, double)

/* built-in operator: primme_svds.matrix = pb200_csr* with its transposed copy
 * (pb200_csr_build_transpose); device blocks.  The kernel context is the one attached to the
 * first-stage eigensolver (primme_b200_attach_ctx(&primme_svds.primme, ctx)). */
void primme_b200_svds_csr_matvec(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy, int *blockSize, int *transpose,
      primme_svds_params *primme_svds, int *ierr) {
   pb200_ctx *ctx = primme_b200_solver_ctx(&primme_svds->primme);
   if (!ctx) ctx = primme_b200_attached_ctx(&primme_svds->primme);
   pb200_csr *A = (pb200_csr *)primme_svds->matrix;
   if (!ctx || !A) {
      *ierr = -1;
      return;
   }
   *ierr = *transpose ? pb200_dspmm_t(ctx, A, (const double *)x, *ldx, (double *)y, *ldy, *blockSize)
                      : pb200_dspmm(ctx, A, (const double *)x, *ldx, (double *)y, *ldy, *blockSize);
}

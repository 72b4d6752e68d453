/* svds.c -- singular value front end.
 *
 * dprimme_svds / cublas_dprimme_svds: the singular triplets of A (m x n) through eigenproblems run by
 * the same Davidson hot path as dprimme --
 *    primme_svds_normalequations   A'A (n <= m) or AA', the matvec applies the user's operator twice;
 *    primme_svds_augmented         [0 A'; A 0] on vectors [v; u];
 *    primme_svds_hybrid (default)  normal equations first, then the augmented operator started from the
 *                                  first stage's triplets, the ones that already pass the test kept as
 *                                  orthogonality constraints.
 * Restates, for fp64:
 *    primme_svds_initialize / set_method / set_defaults   src/svds/primme_svds_interface.c:107-420
 *    wrapper_svds (both stages)                           src/svds/primme_svds_c.c:388-540
 *    copy_last_params_from_svds / _to_svds                :551-1030
 *    matrixMatvec_eigs, applyPreconditioner_eigs          :1323-1417
 *    compute_resNorm, default_convTestFun, convTestFunATA, convTestFunAug   :1512-1745
 * Row-partitioned runs (numProcs > 1, mLocal / nLocal rows of the left / right vectors per process): the
 * eigensolver's collectives and the triplet-level sums go through primme_svds.globalSumReal or the
 * communicator of the attached kernel context; the user's matrixMatvec owns the exchange of the operator.
 * Refused with PRIMME_FUNCTION_UNAVAILABLE = -44 like a reference build without the feature: every
 * precision but double; and whatever the eigensolver underneath refuses (front.c:check_scope).  The
 * refined extraction that the augmented operator selects for smallest / closest_abs targets
 * (primme_svds_interface.c:385-391) runs (dav_refined.c).
 */
#include "pb_host.h"
#include "../../include/primme_svds.h"
#include <float.h>
#include <limits.h>
#include <pthread.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHK_RC(call)            \
   do {                         \
      int chk_rc_ = (call);     \
      if (chk_rc_) return chk_rc_; \
   } while (0)

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

/* the configuration in the reference's text format (primme_svds_interface.c:420-512): the test
 * drivers print it and read it back (tests/COMMON/shared_utils.c) */
void primme_svds_display_params(primme_svds_params s) {
   FILE *f = s.outputFile ? s.outputFile : stdout;
   static const char *ops[] = {"primme_svds_op_none", "primme_svds_op_AtA", "primme_svds_op_AAt", "primme_svds_op_augmented"};
   fprintf(f, "// ---------------------------------------------------\n"
              "//            primme_svds configuration               \n"
              "// ---------------------------------------------------\n");
   fprintf(f, "primme_svds.m = %" PRIMME_INT_P "\n", s.m);
   fprintf(f, "primme_svds.n = %" PRIMME_INT_P "\n", s.n);
   fprintf(f, "primme_svds.mLocal = %" PRIMME_INT_P "\n", s.mLocal);
   fprintf(f, "primme_svds.nLocal = %" PRIMME_INT_P "\n", s.nLocal);
   fprintf(f, "primme_svds.numProcs = %d\n", s.numProcs);
   fprintf(f, "primme_svds.procID = %d\n", s.procID);
   fprintf(f, "\n// Output and reporting\n");
   fprintf(f, "primme_svds.printLevel = %d\n", s.printLevel);
   fprintf(f, "\n// Solver parameters\n");
   fprintf(f, "primme_svds.numSvals = %d\n", s.numSvals);
   fprintf(f, "primme_svds.aNorm = %e\n", s.aNorm);
   fprintf(f, "primme_svds.eps = %e\n", s.eps);
   fprintf(f, "primme_svds.maxBasisSize = %d\n", s.maxBasisSize);
   fprintf(f, "primme_svds.maxBlockSize = %d\n", s.maxBlockSize);
   fprintf(f, "primme_svds.maxMatvecs = %" PRIMME_INT_P "\n", s.maxMatvecs);
   if (s.target == primme_svds_smallest) fprintf(f, "primme_svds.target = primme_svds_smallest\n");
   if (s.target == primme_svds_largest) fprintf(f, "primme_svds.target = primme_svds_largest\n");
   if (s.target == primme_svds_closest_abs) fprintf(f, "primme_svds.target = primme_svds_closest_abs\n");
   fprintf(f, "primme_svds.numTargetShifts = %d\n", s.numTargetShifts);
   if (s.numTargetShifts > 0) {
      fprintf(f, "primme_svds.targetShifts =");
      for (int i = 0; i < s.numTargetShifts; i++) fprintf(f, " %e", s.targetShifts[i]);
      fprintf(f, "\n");
   }
   fprintf(f, "primme_svds.locking = %d\n", s.locking);
   fprintf(f, "primme_svds.initSize = %d\n", s.initSize);
   fprintf(f, "primme_svds.numOrthoConst = %d\n", s.numOrthoConst);
   fprintf(f, "primme_svds.iseed =");
   for (int i = 0; i < 4; i++) fprintf(f, " %" PRIMME_INT_P, s.iseed[i]);
   fprintf(f, "\n");
   fprintf(f, "primme_svds.precondition = %d\n", s.precondition);
   if ((unsigned)s.method < 4) fprintf(f, "primme_svds.method = %s\n", ops[s.method]);
   if ((unsigned)s.methodStage2 < 4) fprintf(f, "primme_svds.methodStage2 = %s\n", ops[s.methodStage2]);
   if (s.internalPrecision == primme_op_half) fprintf(f, "primme_svds.internalPrecision = primme_op_half\n");
   if (s.internalPrecision == primme_op_float) fprintf(f, "primme_svds.internalPrecision = primme_op_float\n");
   if (s.internalPrecision == primme_op_double) fprintf(f, "primme_svds.internalPrecision = primme_op_double\n");
   if (s.internalPrecision == primme_op_quad) fprintf(f, "primme_svds.internalPrecision = primme_op_quad\n");
   if (s.method != primme_svds_op_none) {
      fprintf(f, "\n"
                 "// ---------------------------------------------------\n"
                 "//            1st stage primme configuration          \n"
                 "// ---------------------------------------------------\n");
      s.primme.outputFile = f;
      primme_display_params_prefix("primme", s.primme);
   }
   if (s.methodStage2 != primme_svds_op_none) {
      fprintf(f, "\n"
                 "// ---------------------------------------------------\n"
                 "//            2st stage primme configuration          \n"
                 "// ---------------------------------------------------\n");
      s.primmeStage2.outputFile = f;
      primme_display_params_prefix("primmeStage2", s.primmeStage2);
   }
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
   pb200_ctx *ctx;  /* kernel context of the call (device mode) */
   double *aux;     /* m x maxBlockSize (A'A) or n x maxBlockSize (AA') intermediate block */
   int64_t aux_rows;
   int aux_cols;
} svds_run;
#define SVDS_MAX_RUNS 16
static svds_run runs[SVDS_MAX_RUNS];

/* concurrent solves on different primme_svds_params: claiming and releasing a slot is serialised */
static pthread_mutex_t runs_lock = PTHREAD_MUTEX_INITIALIZER;

static svds_run *find_run(const primme_svds_params *s) {
   svds_run *hit = NULL;
   pthread_mutex_lock(&runs_lock);
   for (int i = 0; i < SVDS_MAX_RUNS && !hit; i++)
      if (runs[i].svds == s) hit = &runs[i];
   pthread_mutex_unlock(&runs_lock);
   return hit;
}

/* a free slot bound to s */
static svds_run *claim_run(primme_svds_params *s, int device_mode) {
   svds_run *hit = NULL;
   pthread_mutex_lock(&runs_lock);
   for (int i = 0; i < SVDS_MAX_RUNS && !hit; i++)
      if (runs[i].svds == NULL) hit = &runs[i];
   if (hit) {
      memset(hit, 0, sizeof(*hit));
      hit->svds = s, hit->device_mode = device_mode;
   }
   pthread_mutex_unlock(&runs_lock);
   return hit;
}

static void release_run(svds_run *run) {
   pthread_mutex_lock(&runs_lock);
   run->svds = NULL, run->aux = NULL;
   pthread_mutex_unlock(&runs_lock);
}

static primme_svds_operator stage_method(const primme_svds_params *s, const primme_params *primme) {
   return &s->primme == primme ? s->method : s->methodStage2;
}

static int call_svds_matvec(primme_svds_params *s, double *x, PRIMME_INT ldx, double *y, PRIMME_INT ldy,
      int bs, int trans) {
   int ierr = 0;
   s->matrixMatvec(x, &ldx, y, &ldy, &bs, &trans, s, &ierr);
   return ierr;
}

/* the operator applied outside the eigensolver: counted in the SVD statistics (matrixMatvecSVDS,
 * primme_svds_c.c:1118-1171) */
static int counted_svds_matvec(primme_svds_params *s, double *x, PRIMME_INT ldx, double *y, PRIMME_INT ldy,
      int bs, int trans) {
   if (bs <= 0) return 0;
   const double t0 = hl_wtime();
   if (call_svds_matvec(s, x, ldx, y, ldy, bs, trans)) return PRIMME_USER_FAILURE;
   s->stats.timeMatvec += hl_wtime() - t0;
   s->stats.numMatvecs += bs;
   return 0;
}

/* ---- n-long helpers on the singular vectors: host loops (dprimme_svds) or kernels of the C-ABI
 * (cublas_dprimme_svds) ---- */
static int vec_alloc(svds_run *run, size_t elems, double **p) {
   const size_t bytes = sizeof(double) * PB_MAX(elems, (size_t)1);
   if (run->device_mode) return pb200_malloc(run->ctx, bytes, (void **)p) ? PRIMME_MALLOC_FAILURE : 0;
   *p = (double *)malloc(bytes);
   return *p ? 0 : PRIMME_MALLOC_FAILURE;
}

static void vec_free(svds_run *run, double *p) {
   if (!p) return;
   if (run->device_mode) pb200_free(run->ctx, p);
   else free(p);
}

static int vec_copy(svds_run *run, const double *src, int64_t lds, double *dst, int64_t ldd, int64_t rows, int cols) {
   if (rows <= 0 || cols <= 0 || src == dst) return 0;
   /* source and destination may overlap (the right vectors slide inside svecs): like the reference's
    * Num_copy_matrix, never read a column after it has been overwritten */
   const double *s_end = src + (size_t)lds * (cols - 1) + rows, *d_end = dst + (size_t)ldd * (cols - 1) + rows;
   const int overlap = src < d_end && dst < s_end;
   if (run->device_mode) {
      if (!overlap) return pb200_copy_d2d(run->ctx, src, lds, dst, ldd, rows, cols, 8);
      double *tmp = NULL;
      if (pb200_malloc(run->ctx, sizeof(double) * (size_t)rows * cols, (void **)&tmp)) return PRIMME_MALLOC_FAILURE;
      int rc = pb200_copy_d2d(run->ctx, src, lds, tmp, rows, rows, cols, 8);
      if (!rc) rc = pb200_copy_d2d(run->ctx, tmp, rows, dst, ldd, rows, cols, 8);
      pb200_ctx_sync(run->ctx);
      pb200_free(run->ctx, tmp);
      return rc;
   }
   if (dst < src)
      for (int j = 0; j < cols; j++) memmove(dst + (size_t)ldd * j, src + (size_t)lds * j, sizeof(double) * rows);
   else
      for (int j = cols - 1; j >= 0; j--) memmove(dst + (size_t)ldd * j, src + (size_t)lds * j, sizeof(double) * rows);
   return 0;
}

/* X(:,j) *= alpha[j] */
static int vec_scale(svds_run *run, double *X, int64_t ldx, int64_t rows, int cols, const double *alpha) {
   if (rows <= 0 || cols <= 0) return 0;
   if (run->device_mode) {
      for (int j = 0; j < cols; j += 8)
         if (pb200_dscale_columns(run->ctx, rows, alpha + j, X + (size_t)ldx * j, ldx, PB_MIN(8, cols - j)))
            return PRIMME_UNEXPECTED_FAILURE;
      return 0;
   }
   for (int j = 0; j < cols; j++)
      for (int64_t r = 0; r < rows; r++) X[r + (size_t)ldx * j] *= alpha[j];
   return 0;
}

/* out[j] = X(:,j)' Y(:,j) */
static int vec_dots(svds_run *run, const double *X, int64_t ldx, const double *Y, int64_t ldy, int64_t rows, int cols,
      double *out) {
   if (run->device_mode) {
      for (int j = 0; j < cols; j += 8)
         if (pb200_dcolumn_dots(run->ctx, rows, X + (size_t)ldx * j, ldx, Y + (size_t)ldy * j, ldy, PB_MIN(8, cols - j),
                   out + j))
            return PRIMME_UNEXPECTED_FAILURE;
      return 0;
   }
   for (int j = 0; j < cols; j++) out[j] = hl_dot((int)rows, X + (size_t)ldx * j, Y + (size_t)ldy * j);
   return 0;
}

/* y += alpha x (one column) */
static int vec_axpy(svds_run *run, int64_t rows, double alpha, const double *x, double *y) {
   if (rows <= 0) return 0;
   if (run->device_mode) return pb200_daxpy_columns(run->ctx, rows, &alpha, x, rows, y, rows, 1) ? PRIMME_UNEXPECTED_FAILURE : 0;
   for (int64_t r = 0; r < rows; r++) y[r] += alpha * x[r];
   return 0;
}

/* uniform(-1,1) column from the eigensolver's evolving seed (Num_larnv, always generated on the host) */
static int vec_random(svds_run *run, primme_params *primme, int64_t rows, double *x) {
   long long seed[4];
   for (int i = 0; i < 4; i++) seed[i] = primme->iseed[i];
   int rc = 0;
   if (run->device_mode) {
      double *h = (double *)malloc(sizeof(double) * PB_MAX(rows, 1));
      if (!h) return PRIMME_MALLOC_FAILURE;
      hl_larnv2(seed, rows, h);
      rc = pb200_copy_h2d(run->ctx, h, PB_MAX(rows, 1), x, PB_MAX(rows, 1), rows, 1, 8) ? PRIMME_UNEXPECTED_FAILURE : 0;
      if (!rc) pb200_ctx_sync(run->ctx);
      free(h);
   } else
      hl_larnv2(seed, rows, x);
   for (int i = 0; i < 4; i++) primme->iseed[i] = seed[i];
   return rc;
}

/* sum over the processes of a row-partitioned run (globalSum_Rprimme_svds, primme_svds_c.c:1760-1800):
 * the user's globalSumReal on host buffers, else the communicator of the kernel context */
static int svds_global_sum(svds_run *run, double *buf, int count) {
   primme_svds_params *s = run->svds;
   if (s->numProcs <= 1 || count <= 0) return 0;
   /* every buffer summed here comes from vec_dots; with a multi-rank kernel context those dots were
    * already all-reduced on the device (pb200_dcolumn_dots), as in the eigensolver's pb_reduce_panel */
   if (run->device_mode && run->ctx && pb200_ctx_nranks(run->ctx) > 1) return 0;
   const double t0 = hl_wtime();
   int ierr = 0;
   if (s->globalSumReal) {
      s->globalSumReal(buf, buf, &count, s, &ierr);
      if (ierr) return PRIMME_USER_FAILURE;
   } else if (run->ctx && pb200_ctx_nranks(run->ctx) > 1) {
      if (pb200_allreduce_host(run->ctx, buf, count)) return PRIMME_PARALLEL_FAILURE;
   } else
      return PRIMME_PARALLEL_FAILURE;
   s->stats.numGlobalSum++;
   s->stats.volumeGlobalSum += count;
   s->stats.timeGlobalSum += hl_wtime() - t0;
   return 0;
}

/* y = A'(A x), A(A' x) in blocks of maxBlockSize columns, or [0 A'; A 0] x (primme_svds_c.c:1323-1383) */
static void matvec_eigs(void *x_, PRIMME_INT *ldx, void *y_, PRIMME_INT *ldy, int *blockSize,
      primme_params *primme, int *ierr) {
   primme_svds_params *s = (primme_svds_params *)primme->matrix;
   svds_run *run = find_run(s);
   double *x = (double *)x_, *y = (double *)y_;
   *ierr = 1;
   if (!run) return;
   const primme_svds_operator method = stage_method(s, primme);
   if (method == primme_svds_op_augmented) {
      int e = call_svds_matvec(s, x + s->nLocal, *ldx, y, *ldy, *blockSize, 1);
      if (!e) e = call_svds_matvec(s, x, *ldx, y + s->nLocal, *ldy, *blockSize, 0);
      *ierr = e;
      return;
   }
   const int ata = method == primme_svds_op_AtA;
   const PRIMME_INT rows = ata ? s->mLocal : s->nLocal;
   const int cap = PB_MAX(1, PB_MIN(primme->maxBlockSize, *blockSize));
   if (!run->aux || run->aux_cols < cap || run->aux_rows < rows) {
      vec_free(run, run->aux);
      run->aux = NULL;
      if (vec_alloc(run, (size_t)PB_MAX(rows, 1) * cap, &run->aux)) return;
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

/* the user's preconditioner told which operator it is approximating (primme_svds_c.c:1404-1417) */
static void precond_eigs(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy, int *blockSize, primme_params *primme,
      int *ierr) {
   primme_svds_params *s = (primme_svds_params *)primme->preconditioner;
   int method = (int)stage_method(s, primme);
   s->applyPreconditioner(x, ldx, y, ldy, blockSize, &method, s, ierr);
}

/* primme_svds_c.c:1200-1252 */
static int conv_test_svds(primme_svds_params *s, double sval, double *left, double *right, double rNorm, int method,
      int *isConv) {
   int ierr = 0;
   s->convTestFun(&sval, left, right, &rNorm, &method, isConv, s, &ierr);
   return ierr ? -1 : 0;
}

/* residual norm of the triplet (u, v): sqrt(|A v/|v| - s u/|u||^2 + |A' u/|u| - s v/|v||^2) with
 * s = u'Av / (|u| |v|)   (primme_svds_c.c:1512-1572) */
static int compute_res_norm(svds_run *run, double *left, double *right, double *rNorm) {
   primme_svds_params *s = run->svds;
   const PRIMME_INT mL = s->mLocal, nL = s->nLocal;
   double *Atu = NULL;
   int rc = vec_alloc(run, (size_t)(mL + nL), &Atu);
   if (rc) return rc;
   double *Av = Atu + nL;
   double ip[3];
   rc = counted_svds_matvec(s, left, mL, Atu, nL, 1, 1);
   if (!rc) rc = counted_svds_matvec(s, right, nL, Av, mL, 1, 0);
   if (!rc) rc = vec_dots(run, right, nL, right, nL, nL, 1, &ip[0]);
   if (!rc) rc = vec_dots(run, left, mL, left, mL, mL, 1, &ip[1]);
   if (!rc) rc = vec_dots(run, left, mL, Av, mL, mL, 1, &ip[2]);
   if (!rc) rc = svds_global_sum(run, ip, 3);
   if (!rc) {
      ip[0] = sqrt(ip[0]), ip[1] = sqrt(ip[1]);
      const double sval = ip[2] / ip[0] / ip[1];
      if (sval < -0.0) {
         *rNorm = DBL_MAX; /* u'Av negative: not a triplet */
      } else {
         double a = 1.0 / ip[1], b = 1.0 / ip[0], nrm2 = 0.0;
         rc = vec_scale(run, Atu, nL, nL, 1, &a);
         if (!rc) rc = vec_axpy(run, nL, -sval / ip[0], right, Atu);
         if (!rc) rc = vec_scale(run, Av, mL, mL, 1, &b);
         if (!rc) rc = vec_axpy(run, mL, -sval / ip[1], left, Av);
         if (!rc) rc = vec_dots(run, Atu, mL + nL, Atu, mL + nL, mL + nL, 1, &nrm2);
         if (!rc) rc = svds_global_sum(run, &nrm2, 1);
         *rNorm = sqrt(nrm2);
      }
   }
   if (run->device_mode) pb200_ctx_sync(run->ctx);
   vec_free(run, Atu);
   return rc;
}

/* primme_svds_c.c:1594-1620: with the augmented operator the eigensolver's residual is only an
 * estimate of the triplet's, so a pair that passes is re-checked with the actual residual */
static void default_conv_test_svds(double *sval, void *leftsvec, void *rightsvec, double *rNorm, int *method,
      int *isConv, primme_svds_params *s, int *ierr) {
   (void)sval;
   const double aNorm = s->aNorm;
   *isConv = *rNorm < PB_MAX(s->eps, PB_EPS * 3.16) * aNorm;
   if (*isConv && *method == primme_svds_op_augmented && leftsvec && rightsvec) {
      svds_run *run = find_run(s);
      double rn = 0.0;
      if (!run || compute_res_norm(run, (double *)leftsvec, (double *)rightsvec, &rn)) {
         *ierr = 1;
         return;
      }
      *isConv = rn < PB_MAX(s->eps, PB_EPS * 3.16) * aNorm;
   }
   *ierr = 0;
}

/* convergence of an eigenpair of A'A / AA' judged as a singular triplet (primme_svds_c.c:1640-1690) */
static void conv_test_normal_equations(double *eval, void *evec, double *rNorm, int *isConv, primme_params *primme,
      int *ierr) {
   primme_svds_params *s = (primme_svds_params *)primme->matrix;
   const primme_svds_operator method = stage_method(s, primme);
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
   const int aat = method == primme_svds_op_AAt;
   *ierr = conv_test_svds(s, sval, aat ? (double *)evec : NULL, aat ? NULL : (double *)evec, srNorm, (int)method, isConv)
                 ? 1
                 : 0;
   s->aNorm = oldaNorm;
}

/* convergence of an eigenpair [v; u] of the augmented operator (primme_svds_c.c:1710-1745); the
 * machine-precision shortcut is deliberately absent: null-space pairs must never pass */
static void conv_test_augmented(double *eval, void *evec, double *rNorm, int *isConv, primme_params *primme,
      int *ierr) {
   primme_svds_params *s = (primme_svds_params *)primme->matrix;
   const double aNorm = primme->aNorm > 0.0 ? primme->aNorm : primme->stats.estimateLargestSVal;
   const double oldaNorm = s->aNorm;
   if (s->aNorm <= 0.0) s->aNorm = aNorm;
   double sval = eval ? fabs(*eval) : 0.0;
   double srNorm = rNorm ? *rNorm * sqrt(2.0) : 0.0;
   *ierr = conv_test_svds(s, sval, evec ? (double *)evec + s->nLocal : NULL, (double *)evec, srNorm,
                 (int)primme_svds_op_augmented, isConv)
                 ? 1
                 : 0;
   s->aNorm = oldaNorm;
}

/* ---- progress reports: the eigensolver's events translated to singular-value terms and handed to
 * primme_svds.monitorFun (monitorFunSVDS :1240-1320, default_monitor_svds :1775-1880,
 * monitor_single_stage :1897-2020, monitor_stage1 :2035-2160, monitor_stage2 :2171-2280) ---- */
static int svds_monitor(primme_svds_params *s, double *basisSvals, int basisSize, int *basisFlags, int *iblock,
      int blockSize, double *basisNorms, int numConverged, double *lockedSvals, int numLocked, int *lockedFlags,
      double *lockedNorms, int inner_its, double LSRes, const char *msg, double time, primme_event event, int stage) {
   if (!s->monitorFun) return 0;
   int err = 0;
   s->monitorFun(basisSvals, &basisSize, basisFlags, iblock, &blockSize, basisNorms, &numConverged, lockedSvals,
         &numLocked, lockedFlags, lockedNorms, inner_its >= 0 ? &inner_its : NULL, LSRes >= 0 ? &LSRes : NULL, msg, &time,
         &event, &stage, s, &err);
   return err ? -1 : 0;
}

static void default_monitor_svds(void *basisSvals_, int *basisSize, int *basisFlags, int *iblock, int *blockSize,
      void *basisNorms_, int *numConverged, void *lockedSvals_, int *numLocked, int *lockedFlags, void *lockedNorms_,
      int *inner_its, void *LSRes_, const char *msg, double *time, primme_event *event, int *stage,
      primme_svds_params *s, int *err) {
   (void)basisSize, (void)basisFlags, (void)inner_its;
   double *basisSvals = (double *)basisSvals_, *basisNorms = (double *)basisNorms_;
   double *lockedSvals = (double *)lockedSvals_, *lockedNorms = (double *)lockedNorms_, *LSRes = (double *)LSRes_;
   FILE *f = s->outputFile;
   *err = 0;
   if (!f || !(s->procID == 0 || *event == primme_event_profile)) return;
   switch (*event) {
   case primme_event_outer_iteration:
      if (s->printLevel >= 3)
         for (int i = 0; i < *blockSize; i++)
            fprintf(f, "OUT %" PRIMME_INT_P " conv %d blk %d MV %" PRIMME_INT_P " Sec %E SV %13E |r| %.3E stage %d\n",
                  s->stats.numOuterIterations, *numConverged, i, s->stats.numMatvecs, s->stats.elapsedTime,
                  basisSvals[iblock[i]], basisNorms[iblock[i]], *stage + 1);
      break;
   case primme_event_inner_iteration:
      if (s->printLevel >= 4)
         fprintf(f, "INN MV %" PRIMME_INT_P " Sec %e Sval %e Lin|r| %.3e SV|r| %.3e stage %d\n", s->stats.numMatvecs,
               s->stats.elapsedTime, basisSvals[iblock[0]], *LSRes, basisNorms[iblock[0]], *stage + 1);
      break;
   case primme_event_converged:
      if ((*stage == 0 && s->printLevel >= 2) || s->printLevel >= 5)
         fprintf(f, "#Converged %d sval[ %d ]= %e norm %e Mvecs %" PRIMME_INT_P " Time %g stage %d\n", *numConverged,
               iblock[0], basisSvals[iblock[0]], basisNorms[iblock[0]], s->stats.numMatvecs, s->stats.elapsedTime,
               *stage + 1);
      break;
   case primme_event_locked:
      if (s->printLevel >= 2)
         fprintf(f, "Lock striplet[ %d ]= %e norm %.4e Mvecs %" PRIMME_INT_P " Time %.4e Flag %d stage %d\n",
               *numLocked - 1, lockedSvals[*numLocked - 1], lockedNorms[*numLocked - 1], s->stats.numMatvecs,
               s->stats.elapsedTime, lockedFlags[*numLocked - 1], *stage + 1);
      break;
   case primme_event_message:
      if (s->printLevel >= 2 && msg) fprintf(f, "%s\n", msg);
      break;
   case primme_event_profile:
      if (msg && time) {
         if (s->printLevel >= 3 && *time < 0.0) fprintf(f, "entering in %s proc %d\n", msg, s->procID);
         if (s->printLevel >= 2 && *time >= 0.0) fprintf(f, "time for %s : %g proc %d\n", msg, *time, s->procID);
      }
      break;
   default: break;
   }
   fflush(f);
}

/* the running totals a report shows include the stage in progress (UPDATE_STATS around the call) */
static primme_svds_stats stats_with_stage(primme_svds_params *s, const primme_params *primme) {
   primme_svds_stats saved = s->stats;
   s->stats.numOuterIterations += primme->stats.numOuterIterations;
   s->stats.numRestarts += primme->stats.numRestarts;
   s->stats.numMatvecs += primme->stats.numMatvecs * 2;
   s->stats.numPreconds += primme->stats.numPreconds;
   s->stats.numGlobalSum += primme->stats.numGlobalSum;
   s->stats.numBroadcast += primme->stats.numBroadcast;
   s->stats.volumeGlobalSum += primme->stats.volumeGlobalSum;
   s->stats.volumeBroadcast += primme->stats.volumeBroadcast;
   s->stats.numOrthoInnerProds += primme->stats.numOrthoInnerProds;
   s->stats.elapsedTime += primme->stats.elapsedTime;
   s->stats.timeMatvec += primme->stats.timeMatvec;
   s->stats.timePrecond += primme->stats.timePrecond;
   s->stats.timeOrtho += primme->stats.timeOrtho;
   s->stats.timeGlobalSum += primme->stats.timeGlobalSum;
   s->stats.timeBroadcast += primme->stats.timeBroadcast;
   s->stats.lockingIssue += primme->stats.lockingIssue;
   return saved;
}

static const char *profile_msg(primme_event event, const char *msg, int stage, char **owned) {
   *owned = NULL;
   if (event != primme_event_profile || !msg) return msg;
   const size_t len = 12 + strlen(msg);
   *owned = (char *)malloc(len);
   if (!*owned) return msg;
   snprintf(*owned, len, "~Sprimme%d%s", stage, msg);
   return *owned;
}

#define MON_ARGS                                                                                                   \
   void *basisEvals_, int *basisSize, int *basisFlags, int *iblock, int *blockSize, void *basisNorms_,             \
         int *numConverged, void *lockedEvals_, int *numLocked, int *lockedFlags, void *lockedNorms_, int *inner_its, \
         void *LSRes_, const char *msg, double *time, primme_event *event, primme_params *primme, int *err

/* one-stage runs: values and norms of the basis and of the locked pairs in singular-value terms */
static void monitor_single_stage(MON_ARGS) {
   primme_svds_params *s = (primme_svds_params *)primme->matrix;
   double *basisEvals = (double *)basisEvals_, *basisNorms = (double *)basisNorms_;
   double *lockedEvals = (double *)lockedEvals_, *lockedNorms = (double *)lockedNorms_, *LSRes = (double *)LSRes_;
   const int nb = basisEvals && basisSize ? *basisSize : 0, nl = lockedEvals && numLocked ? *numLocked : 0;
   double *bs = (double *)calloc(2 * (size_t)PB_MAX(nb, 1), sizeof(double)), *bn = bs + PB_MAX(nb, 1);
   double *ls = (double *)calloc(2 * (size_t)PB_MAX(nl, 1), sizeof(double)), *ln = ls + PB_MAX(nl, 1);
   if (s->method != primme_svds_op_augmented) {
      for (int i = 0; i < nb; i++) {
         bs[i] = sqrt(fabs(basisEvals[i]));
         bn[i] = bs[i] > 0.0 ? basisNorms[i] / bs[i] : basisNorms[i];
      }
      for (int i = 0; i < nl; i++) {
         ls[i] = sqrt(fabs(lockedEvals[i]));
         ln[i] = ls[i] > 0.0 ? lockedNorms[i] / ls[i] : lockedNorms[i];
      }
   } else {
      /* the reference leaves the values of the augmented operator unset here (:1960-1975); they are
       * the singular values themselves */
      for (int i = 0; i < nb; i++) bs[i] = basisEvals[i], bn[i] = basisNorms ? basisNorms[i] / sqrt(2.0) : 0.0;
      for (int i = 0; i < nl; i++) ls[i] = lockedEvals[i], ln[i] = lockedNorms[i] / sqrt(2.0);
   }
   char *owned;
   msg = profile_msg(*event, msg, 0, &owned);
   primme_svds_stats saved = stats_with_stage(s, primme);
   *err = svds_monitor(s, bs, nb, basisFlags, iblock, blockSize ? *blockSize : 0, bn, numConverged ? *numConverged : 0, ls,
                numLocked ? *numLocked : 0, lockedFlags, ln, inner_its ? *inner_its : 0, LSRes ? *LSRes : 0.0, msg,
                time ? *time : 0.0, *event, 0)
                ? 1
                : 0;
   s->stats = saved;
   free(bs), free(ls), free(owned);
}

/* first stage of a two-stage run: locked pairs shown as converged pairs of the basis */
static void monitor_stage1(MON_ARGS) {
   primme_svds_params *s = (primme_svds_params *)primme->matrix;
   double *basisEvals = (double *)basisEvals_, *basisNorms = (double *)basisNorms_;
   double *lockedEvals = (double *)lockedEvals_, *lockedNorms = (double *)lockedNorms_, *LSRes = (double *)LSRes_;
   *err = 0;
   if (*event == primme_event_converged && primme->locking && primme->printLevel <= 4) return;
   const int nl = lockedEvals && numLocked ? *numLocked : 0;
   const int nb = (basisEvals && basisSize ? *basisSize : 0) + nl;
   double *sv = (double *)calloc(2 * (size_t)PB_MAX(nb, 1), sizeof(double)), *sn = sv + PB_MAX(nb, 1);
   int *fl = (int *)calloc((size_t)PB_MAX(nb, 1), sizeof(int));
   int *ib = (int *)calloc((size_t)(blockSize && *blockSize > 0 ? *blockSize : 1), sizeof(int));
   int j = 0;
   for (int i = 0; i < nl; i++, j++) {
      sv[j] = sqrt(fabs(lockedEvals[i]));
      sn[j] = sv[i] > 0.0 ? lockedNorms[i] / sv[i] : lockedNorms[i];
      fl[j] = lockedFlags[i];
   }
   for (int i = 0; i < nb - nl; i++, j++) {
      sv[j] = sqrt(fabs(basisEvals[i]));
      sn[j] = sv[i] > 0.0 ? basisNorms[i] / sv[i] : basisNorms[i]; /* sv[i], not sv[j]: as the reference (:2098) */
      fl[j] = basisFlags ? basisFlags[i] : UNCONVERGED;
   }
   if (iblock && blockSize)
      for (int i = 0; i < *blockSize; i++) ib[i] = iblock[i] + nl;
   primme_event ev = *event;
   if (ev == primme_event_locked) ev = primme_event_converged, ib[0] = *numLocked - 1;
   char *owned;
   msg = profile_msg(*event, msg, 0, &owned);
   primme_svds_stats saved = stats_with_stage(s, primme);
   *err = svds_monitor(s, sv, nb, fl, ib, blockSize ? *blockSize : 0, sn, numConverged ? *numConverged : nl, NULL, 0, NULL,
                NULL, inner_its ? *inner_its : 0, LSRes ? *LSRes : 0.0, msg, time ? *time : 0.0, ev, 0)
                ? 1
                : 0;
   s->stats = saved;
   free(sv), free(fl), free(ib), free(owned);
}

/* second stage: the triplets the first stage converged count as locked */
static void monitor_stage2(MON_ARGS) {
   primme_svds_params *s = (primme_svds_params *)primme->matrix;
   double *basisEvals = (double *)basisEvals_, *basisNorms = (double *)basisNorms_;
   double *lockedEvals = (double *)lockedEvals_, *lockedNorms = (double *)lockedNorms_, *LSRes = (double *)LSRes_;
   const int extra = lockedEvals && numLocked ? s->numSvals - primme->numEvals : 0;
   const int nl = (lockedEvals && numLocked ? *numLocked : 0) + extra;
   const int nb = basisEvals && basisSize ? *basisSize : 0;
   double *bs = (double *)calloc(2 * (size_t)PB_MAX(nb, 1), sizeof(double)), *bn = bs + PB_MAX(nb, 1);
   double *ls = (double *)calloc(2 * (size_t)PB_MAX(nl, 1), sizeof(double)), *ln = ls + PB_MAX(nl, 1);
   int *lf = (int *)calloc((size_t)PB_MAX(nl, 1), sizeof(int));
   for (int i = 0; i < nb; i++) bs[i] = basisEvals[i], bn[i] = basisNorms[i] / sqrt(2.0);
   /* lockedEvals / lockedNorms point into the caller's svals / resNorms past the first stage's triplets */
   if (lockedEvals) lockedEvals -= extra, lockedNorms -= extra;
   for (int i = 0; i < extra; i++) ls[i] = lockedEvals[i], ln[i] = lockedNorms[i], lf[i] = CONVERGED;
   for (int i = extra; i < nl; i++) ls[i] = lockedEvals[i], ln[i] = lockedNorms[i] / sqrt(2.0), lf[i] = lockedFlags[i - extra];
   char *owned;
   msg = profile_msg(*event, msg, 1, &owned);
   primme_svds_stats saved = stats_with_stage(s, primme);
   *err = svds_monitor(s, bs, nb, basisFlags, iblock, blockSize ? *blockSize : 0, bn, numConverged ? *numConverged : 0, ls, nl,
                lf, ln, inner_its ? *inner_its : 0, LSRes ? *LSRes : 0.0, msg, time ? *time : 0.0, *event, 1)
                ? 1
                : 0;
   s->stats = saved;
   free(bs), free(ls), free(lf), free(owned);
}

static int svds_check_input(void *svals, void *svecs, void *resNorms, primme_svds_params *s) {
   if (!s) return -4;
   if (s->n < 0 || s->m < 0 || s->nLocal < 0 || s->mLocal < 0 || s->nLocal > s->n || s->mLocal > s->m) return -5;
   if (s->numProcs < 1) return -6;
   if (!s->matrixMatvec) return -7;
   if (!s->applyPreconditioner && s->precondition == 1) return -8;
   /* the sums of a row-partitioned run: the user's callback, or the communicator of the attached kernel
    * context (every device panel is then all-reduced by the kernel that produces it) */
   if (s->numProcs > 1 && !s->globalSumReal) {
      pb200_ctx *c = primme_b200_attached_ctx(&s->primme);
      if (!c || pb200_ctx_nranks(c) <= 1) return -9;
   }
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
   if ((s->matrixMatvec_type != primme_op_default && s->matrixMatvec_type != primme_op_double) ||
         (s->applyPreconditioner && s->applyPreconditioner_type != primme_op_default &&
               s->applyPreconditioner_type != primme_op_double) ||
         (s->internalPrecision != primme_op_default && s->internalPrecision != primme_op_double))
      why = "callbacks / internal precision other than double";
   else if ((s->method != primme_svds_op_none && s->primme.projectionParams.projection == primme_proj_harmonic) ||
            (s->methodStage2 != primme_svds_op_none && s->primmeStage2.projectionParams.projection == primme_proj_harmonic))
      why = "harmonic extraction"; /* known before the first stage runs */
   if (!why) return 0;
   if (s->outputFile && s->printLevel > 0)
      fprintf(s->outputFile, "PRIMME-B200: %s is outside the scope of this build\n", why);
   return PRIMME_FUNCTION_UNAVAILABLE;
}

static int cmp_double(const void *a, const void *b) { return *(const double *)a <= *(const double *)b ? -1 : 1; }

/* Prepare the eigensolver of one stage: callbacks, norms, vectors in the layout it works on, seeds,
 * shifts (copy_last_params_from_svds, primme_svds_c.c:551-830).  *out_svecs is the evecs argument. */
static int stage_begin(svds_run *run, int stage, double *svals, double *svecs, double *rnorms, int *allocatedShifts,
      double **out_svecs) {
   primme_svds_params *s = run->svds;
   primme_params *primme = stage == 0 ? &s->primme : &s->primmeStage2;
   const primme_svds_operator method = stage == 0 ? s->method : s->methodStage2;
   const PRIMME_INT mL = s->mLocal, nL = s->nLocal;
   *out_svecs = svecs;
   *allocatedShifts = 0;
   if (method == primme_svds_op_none) {
      primme->maxMatvecs = 0;
      return 0;
   }
   if (!primme->matrixMatvec) {
      primme->matrixMatvec = matvec_eigs;
      primme->matrixMatvec_type = primme_op_double;
      primme->matrix = s;
   }
   if (s->applyPreconditioner && !primme->applyPreconditioner) {
      primme->applyPreconditioner = precond_eigs;
      primme->applyPreconditioner_type = primme_op_double;
      primme->preconditioner = s;
   }
   if (s->aNorm > 0.0) primme->aNorm = method == primme_svds_op_augmented ? s->aNorm : s->aNorm * s->aNorm;
   primme->convTestFun = method == primme_svds_op_augmented ? conv_test_augmented : conv_test_normal_equations;
   primme->convTestFun_type = primme_op_double;

   /* svecs = [Uc U0 Vc V0]: constraints and initial guesses, left then right */
   primme->initSize = s->initSize;
   primme->numOrthoConst = s->numOrthoConst;
   const int n0 = s->initSize + s->numOrthoConst;
   const int nMax = PB_MAX(s->initSize, s->numSvals) + s->numOrthoConst;
   if (method != primme_svds_op_augmented) {
      /* the right vectors [Vc V0] move to the rightmost position of svecs, where the eigensolver
       * works in place; with AA' only Vc moves and the eigensolver works on the left block */
      const int ata = method == primme_svds_op_AtA;
      double *aux = svecs + (size_t)nMax * mL;
      CHK_RC(vec_copy(run, svecs + (size_t)mL * n0, nL, aux, nL, nL, ata ? n0 : s->numOrthoConst));
      if (ata) *out_svecs = aux;
      primme->ldevecs = ata ? nL : mL;
   } else {
      /* shuffle to [V; U] columns of height nL + mL; constraints get norm 1 (each half has norm 1) */
      const PRIMME_INT N = mL + nL;
      if (n0 > 0) {
         double *aux = NULL;
         CHK_RC(vec_alloc(run, (size_t)N * n0, &aux));
         int rc = vec_copy(run, svecs, N * n0, aux, N * n0, N * n0, 1);
         if (!rc) rc = vec_copy(run, aux + (size_t)mL * n0, nL, svecs, N, nL, n0);
         if (!rc) rc = vec_copy(run, aux, mL, svecs + nL, N, mL, n0);
         if (run->device_mode) pb200_ctx_sync(run->ctx);
         vec_free(run, aux);
         CHK_RC(rc);
      }
      double *isq2 = (double *)malloc(sizeof(double) * PB_MAX(s->numOrthoConst, 1));
      for (int i = 0; i < s->numOrthoConst; i++) isq2[i] = 1. / sqrt(2.);
      int rc = vec_scale(run, svecs, N, N, s->numOrthoConst, isq2);
      free(isq2);
      CHK_RC(rc);
      primme->ldevecs = N;
   }
   for (int i = 0; i < 4; i++) primme->iseed[i] = s->iseed[i];
   primme->maxMatvecs = stage == 0 ? s->maxMatvecs / 2 : s->maxMatvecs / 2 - s->primme.stats.numMatvecs;

   if ((stage == 0 && s->numTargetShifts > 0) ||
         (stage == 1 && primme->targetShifts == NULL && s->target == primme_svds_closest_abs)) {
      primme->numTargetShifts = s->numTargetShifts;
      if (stage == 0 && method != primme_svds_op_augmented) {
         *allocatedShifts = 1;
         primme->targetShifts = (double *)malloc(sizeof(double) * PB_MAX(s->numSvals, s->numTargetShifts));
         if (!primme->targetShifts) return PRIMME_MALLOC_FAILURE;
         for (int i = 0; i < primme->numTargetShifts; i++) primme->targetShifts[i] = s->targetShifts[i] * s->targetShifts[i];
      } else
         primme->targetShifts = s->targetShifts;
   } else if (stage == 1 && primme->targetShifts == NULL && s->target == primme_svds_smallest) {
      /* closest_geq to lower bounds of the values found by the first stage: the |m - n| null
       * eigenvalues of the augmented operator are not singular values (:700-735) */
      *allocatedShifts = 1;
      primme->targetShifts = (double *)malloc(sizeof(double) * s->numSvals);
      if (!primme->targetShifts) return PRIMME_MALLOC_FAILURE;
      const double min_val = s->aNorm * PB_EPS;
      int i;
      for (i = 0; i < s->initSize; i++)
         primme->targetShifts[i] = PB_MAX(sqrt(fabs(PB_MAX(svals[i] - rnorms[i], 0.0) * svals[i])), min_val);
      for (; i < s->numSvals; i++) primme->targetShifts[i] = min_val;
      qsort(primme->targetShifts, s->numSvals, sizeof(double), cmp_double);
      primme->numTargetShifts = s->numSvals;
   } else if (method == primme_svds_op_augmented && s->target == primme_svds_smallest && primme->targetShifts == NULL) {
      *allocatedShifts = 1;
      primme->targetShifts = (double *)malloc(sizeof(double));
      if (!primme->targetShifts) return PRIMME_MALLOC_FAILURE;
      primme->targetShifts[0] = 0.0;
      primme->numTargetShifts = 1;
   }

   /* no guess for the augmented operator: start from [A'x; x] or [x; Ax], x random (:752-790) */
   if (method == primme_svds_op_augmented && primme->initSize <= 0) {
      const PRIMME_INT N = mL + nL;
      double *v0 = svecs + (size_t)primme->numOrthoConst * N, n2[2];
      if (s->m >= s->n) {
         CHK_RC(vec_random(run, primme, mL, v0 + nL));
         CHK_RC(counted_svds_matvec(s, v0 + nL, mL, v0, nL, 1, 1));
      } else {
         CHK_RC(vec_random(run, primme, nL, v0));
         CHK_RC(counted_svds_matvec(s, v0, nL, v0 + nL, mL, 1, 0));
      }
      CHK_RC(vec_dots(run, v0, nL, v0, nL, nL, 1, &n2[0]));
      CHK_RC(vec_dots(run, v0 + nL, mL, v0 + nL, mL, mL, 1, &n2[1]));
      CHK_RC(svds_global_sum(run, n2, 2));
      n2[0] = 1.0 / sqrt(n2[0]), n2[1] = 1.0 / sqrt(n2[1]);
      CHK_RC(vec_scale(run, v0, nL, nL, 1, &n2[0]));
      CHK_RC(vec_scale(run, v0 + nL, mL, mL, 1, &n2[1]));
      primme->initSize = 1;
      if (rnorms) rnorms[0] = DBL_MAX;
      primme->initBasisMode = primme_init_user;
   }

   /* second stage: the leading triplets that already pass the convergence test become
    * orthogonality constraints (:795-826) */
   if (stage == 1) {
      const PRIMME_INT N = mL + nL;
      for (int i = 0; primme->initSize > 0; i++) {
         int isConv = 0;
         if (conv_test_svds(s, svals[i], svecs + (size_t)N * primme->numOrthoConst + nL,
                   svecs + (size_t)N * primme->numOrthoConst, rnorms[i], (int)method, &isConv))
            return -1;
         if (!isConv) break;
         /* reported as locked by the first stage */
         {
            int *flags = (int *)malloc(sizeof(int) * (size_t)(i + 1));
            for (int t = 0; t <= i; t++) flags[t] = CONVERGED;
            int e = svds_monitor(s, NULL, 0, NULL, NULL, 0, NULL, 0, svals, i + 1, flags, rnorms, 0, 0.0, NULL, 0.0,
                  primme_event_locked, 0);
            free(flags);
            if (e) return -1;
         }
         primme->numOrthoConst++;
         primme->initSize--;
         primme->numEvals--;
      }
   }
   if (s->locking >= 0) primme->locking = s->locking;
   if (!primme->monitorFun) {
      primme->monitorFun = s->methodStage2 == primme_svds_op_none ? monitor_single_stage : stage == 0 ? monitor_stage1 : monitor_stage2;
      primme->monitorFun_type = primme_op_double;
   }
   primme->queue = s->queue;
   primme->profile = s->profile;
   return 0;
}

/* Results of one stage back into primme_svds and svecs = [Uc U Vc V] (copy_last_params_to_svds,
 * primme_svds_c.c:838-1030). */
static int stage_end(svds_run *run, int stage, double *svals, double *svecs, double *rnorms, int allocatedShifts) {
   primme_svds_params *s = run->svds;
   primme_params *primme = stage == 0 ? &s->primme : &s->primmeStage2;
   const primme_svds_operator method = stage == 0 ? s->method : s->methodStage2;
   const PRIMME_INT mL = s->mLocal, nL = s->nLocal;
   if (method == primme_svds_op_none) {
      primme->maxMatvecs = 0;
      return 0;
   }
   if (primme->initSize < 0) primme->initSize = 0; /* failed solve: nothing to return */
   if (stage == 1) {
      const int nconv = s->numSvals - primme->numEvals;
      primme->initSize += nconv;
      primme->numOrthoConst -= nconv;
      primme->numEvals += nconv;
   }
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
   if (primme->aNorm > 0.0) s->aNorm = method == primme_svds_op_augmented ? primme->aNorm : sqrt(primme->aNorm);
   if (method != primme_svds_op_augmented)
      for (int i = 0; i < primme->initSize; i++) svals[i] = sqrt(PB_MAX(0.0, svals[i]));

   const int nMax = PB_MAX(s->initSize, s->numSvals) + s->numOrthoConst;
   s->initSize = primme->initSize;
   const int nconv = s->initSize;
   const int n0 = s->initSize + s->numOrthoConst;
   int rc = 0;
   if (method != primme_svds_op_augmented) {
      const int ata = method == primme_svds_op_AtA;
      double *aux = svecs + (size_t)nMax * mL;
      double *inv = (double *)malloc(sizeof(double) * PB_MAX(nconv, 1));
      for (int i = 0; i < nconv; i++) inv[i] = 1.0 / svals[i];
      if (ata) {
         /* U = A V diag(1/sigma), then [Vc V] packed right after the n0 left vectors */
         double *U = svecs + (size_t)mL * s->numOrthoConst;
         double *V = aux + (size_t)nL * s->numOrthoConst;
         rc = counted_svds_matvec(s, V, nL, U, mL, nconv, 0);
         if (!rc) rc = vec_scale(run, U, mL, mL, nconv, inv);
         if (!rc) rc = vec_copy(run, aux, nL, svecs + (size_t)mL * n0, nL, nL, n0);
      } else {
         /* the constraints Vc first, then V = A' U diag(1/sigma) */
         rc = vec_copy(run, aux, nL, svecs + (size_t)mL * n0, nL, nL, s->numOrthoConst);
         double *U = svecs + (size_t)mL * s->numOrthoConst;
         double *V = svecs + (size_t)mL * n0 + (size_t)nL * s->numOrthoConst;
         if (!rc) rc = counted_svds_matvec(s, U, mL, V, nL, nconv, 1);
         if (!rc) rc = vec_scale(run, V, nL, nL, nconv, inv);
      }
      free(inv);
   } else {
      const PRIMME_INT N = mL + nL;
      double *sq2 = (double *)malloc(sizeof(double) * PB_MAX(2 * n0, 1));
      for (int i = 0; i < s->numOrthoConst; i++) sq2[i] = sqrt(2.);
      rc = vec_scale(run, svecs, N, N, s->numOrthoConst, sq2);
      /* [Vc V; Uc U] back to [Uc U Vc V], every column of U and of V normalised */
      if (!rc && n0 > 0) {
         double *aux = NULL;
         rc = vec_alloc(run, (size_t)N * n0, &aux);
         if (!rc) rc = vec_copy(run, svecs, N * n0, aux, N * n0, N * n0, 1);
         if (!rc) rc = vec_copy(run, aux, N, svecs + (size_t)mL * n0, nL, nL, n0);
         if (!rc) rc = vec_copy(run, aux + nL, N, svecs, mL, mL, n0);
         if (run->device_mode) pb200_ctx_sync(run->ctx);
         vec_free(run, aux);
         double *U = svecs, *V = svecs + (size_t)mL * n0;
         if (!rc) rc = vec_dots(run, U, mL, U, mL, mL, n0, sq2);
         if (!rc) rc = vec_dots(run, V, nL, V, nL, nL, n0, sq2 + n0);
         if (!rc) rc = svds_global_sum(run, sq2, 2 * n0);
         for (int i = 0; i < 2 * n0; i++) sq2[i] = 1.0 / sqrt(sq2[i]);
         if (!rc) rc = vec_scale(run, U, mL, mL, n0, sq2);
         if (!rc) rc = vec_scale(run, V, nL, nL, n0, sq2 + n0);
      }
      free(sq2);
   }
   if (run->device_mode && run->ctx) pb200_ctx_sync(run->ctx);
   for (int i = 0; i < 4; i++) s->iseed[i] = primme->iseed[i];
   if (allocatedShifts) {
      free(primme->targetShifts);
      primme->targetShifts = NULL;
   }
   /* normal equations: the residual of the triplet is the eigen-residual over sigma; augmented: the
    * convergence test already replaced it by the actual one up to the sqrt(2) of the normalisation */
   if (method != primme_svds_op_augmented)
      for (int i = 0; i < nconv; i++) rnorms[i] = PB_MIN(rnorms[i] / svals[i], s->aNorm);
   else
      for (int i = 0; i < nconv; i++) rnorms[i] *= sqrt(2.0);
   return rc;
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
   if (!s->monitorFun) {
      s->monitorFun = default_monitor_svds;
      s->monitorFun_type = primme_op_double;
   }
   memset(&s->stats, 0, sizeof(s->stats));

   svds_run *run = claim_run(s, device_mode);
   if (!run) return PRIMME_MALLOC_FAILURE;

   /* one kernel context for the whole call: the caller's (attached to the first-stage eigensolver)
    * or a private one; the vector post-processing needs it outside the eigensolver */
   pb200_ctx *ctx = primme_b200_attached_ctx(&s->primme);
   int own_ctx = 0;
   if (!ctx && device_mode) {
      if (pb200_ctx_create(&ctx, -1)) {
         release_run(run);
         s->initSize = 0;
         return PRIMME_FUNCTION_UNAVAILABLE;
      }
      own_ctx = 1;
      primme_b200_attach_ctx(&s->primme, ctx);
   }
   run->ctx = ctx;
   const int attach2 = ctx && s->methodStage2 != primme_svds_op_none && !primme_b200_attached_ctx(&s->primmeStage2);
   if (attach2) primme_b200_attach_ctx(&s->primmeStage2, ctx);

   int ret = 0;
   for (int stage = 0; stage < 2 && ret == 0 && rc == 0; stage++) {
      primme_params *primme = stage == 0 ? &s->primme : &s->primmeStage2;
      if (stage == 1 && s->methodStage2 == primme_svds_op_none) break;
      int allocatedShifts = 0;
      double *evecs = NULL;
      rc = stage_begin(run, stage, svals, svecs, resNorms, &allocatedShifts, &evecs);
      if (rc) {
         if (allocatedShifts) free(primme->targetShifts), primme->targetShifts = NULL;
         break;
      }
      /* numSvals - numEvals triplets of the first stage are already converged (:505-508) */
      const int skip = stage == 1 ? s->numSvals - primme->numEvals : 0;
      ret = device_mode ? cublas_dprimme(svals + skip, evecs, resNorms + skip, primme)
                        : dprimme(svals + skip, evecs, resNorms + skip, primme);
      rc = stage_end(run, stage, svals, svecs, resNorms, allocatedShifts);
      if (ret != 0) ret -= 100 * (stage + 1); /* errors of the first / second stage (:497,515) */
   }

   vec_free(run, run->aux);
   release_run(run);
   if (attach2) primme_b200_attach_ctx(&s->primmeStage2, NULL);
   if (own_ctx) {
      primme_b200_attach_ctx(&s->primme, NULL);
      pb200_ctx_destroy(ctx);
   }
   if (ret != 0) return ret;
   return rc;
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

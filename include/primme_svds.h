/* primme_svds.h -- public SVD API (ABI mirror of the reference's include/primme_svds.h:42-260).
 *
 * The SVD front end is an adjacent ("next") row of the hot-path scope:
 * dprimme_svds with primme_svds_normalequations drives the same Davidson inner loop through
 * the normal-equations operator (reference src/svds/primme_svds_c.c:1323-1383).  The struct and
 * enum layout is kept byte-compatible so reference callers compile unchanged.
 */
#ifndef PRIMME_SVDS_H
#define PRIMME_SVDS_H

#include "primme_eigs.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
   primme_svds_largest,
   primme_svds_smallest,
   primme_svds_closest_abs
} primme_svds_target;

typedef enum {
   primme_svds_default,
   primme_svds_hybrid,
   primme_svds_normalequations,
   primme_svds_augmented
} primme_svds_preset_method;

typedef enum {
   primme_svds_op_none,
   primme_svds_op_AtA,
   primme_svds_op_AAt,
   primme_svds_op_augmented
} primme_svds_operator;

typedef struct primme_svds_stats {
   PRIMME_INT numOuterIterations;
   PRIMME_INT numRestarts;
   PRIMME_INT numMatvecs;
   PRIMME_INT numPreconds;
   PRIMME_INT numGlobalSum;
   PRIMME_INT numBroadcast;
   PRIMME_INT volumeGlobalSum;
   PRIMME_INT volumeBroadcast;
   double numOrthoInnerProds;
   double elapsedTime;
   double timeMatvec;
   double timePrecond;
   double timeOrtho;
   double timeGlobalSum;
   double timeBroadcast;
   PRIMME_INT lockingIssue;
} primme_svds_stats;

typedef struct primme_svds_params {
   primme_params primme;       /* stage-1 eigensolver configuration (must stay first) */
   primme_params primmeStage2; /* stage-2 (hybrid) */

   PRIMME_INT m; /* rows of A */
   PRIMME_INT n; /* columns of A */

   void (*matrixMatvec)(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy, int *blockSize,
         int *transpose, struct primme_svds_params *primme_svds, int *ierr);
   primme_op_datatype matrixMatvec_type;
   void (*applyPreconditioner)(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy,
         int *blockSize, int *transpose, struct primme_svds_params *primme_svds, int *ierr);
   primme_op_datatype applyPreconditioner_type;

   int numProcs;
   int procID;
   PRIMME_INT mLocal;
   PRIMME_INT nLocal;
   void *commInfo;
   void (*globalSumReal)(void *sendBuf, void *recvBuf, int *count,
         struct primme_svds_params *primme_svds, int *ierr);
   primme_op_datatype globalSumReal_type;
   void (*broadcastReal)(
         void *buffer, int *count, struct primme_svds_params *primme_svds, int *ierr);
   primme_op_datatype broadcastReal_type;

   int numSvals;
   primme_svds_target target;
   int numTargetShifts;
   double *targetShifts;
   primme_svds_operator method;
   primme_svds_operator methodStage2;

   void *matrix;
   void *preconditioner;

   int locking;
   int numOrthoConst;
   double aNorm;
   double eps;

   int precondition;
   int initSize;
   int maxBasisSize;
   int maxBlockSize;
   PRIMME_INT maxMatvecs;
   PRIMME_INT iseed[4];
   int printLevel;
   primme_op_datatype internalPrecision;
   FILE *outputFile;
   struct primme_svds_stats stats;

   void (*convTestFun)(double *sval, void *leftsvec, void *rightsvec, double *rNorm,
         int *method, int *isconv, struct primme_svds_params *primme, int *ierr);
   primme_op_datatype convTestFun_type;
   void *convtest;
   void (*monitorFun)(void *basisSvals, int *basisSize, int *basisFlags, int *iblock,
         int *blockSize, void *basisNorms, int *numConverged, void *lockedSvals,
         int *numLocked, int *lockedFlags, void *lockedNorms, int *inner_its, void *LSRes,
         const char *msg, double *time, primme_event *event, int *stage,
         struct primme_svds_params *primme_svds, int *err);
   primme_op_datatype monitorFun_type;
   void *monitor;
   void *queue;
   const char *profile;
} primme_svds_params;

/* X(label suffix, id, lvalue path, kind) -- ids as reference primme_svds.h:166-232.
 * kind Z = nested primme_params (get returns its address). */
#define PRIMME_SVDS_PARAM_TABLE(X) \
   X(primme, 1, primme, Z) \
   X(primmeStage2, 2, primmeStage2, Z) \
   X(m, 3, m, I) \
   X(n, 4, n, I) \
   X(matrixMatvec, 5, matrixMatvec, F) \
   X(matrixMatvec_type, 6, matrixMatvec_type, I) \
   X(applyPreconditioner, 7, applyPreconditioner, F) \
   X(applyPreconditioner_type, 8, applyPreconditioner_type, I) \
   X(numProcs, 9, numProcs, I) \
   X(procID, 10, procID, I) \
   X(mLocal, 11, mLocal, I) \
   X(nLocal, 12, nLocal, I) \
   X(commInfo, 13, commInfo, P) \
   X(globalSumReal, 14, globalSumReal, F) \
   X(globalSumReal_type, 15, globalSumReal_type, I) \
   X(broadcastReal, 16, broadcastReal, F) \
   X(broadcastReal_type, 17, broadcastReal_type, I) \
   X(numSvals, 18, numSvals, I) \
   X(target, 19, target, I) \
   X(numTargetShifts, 20, numTargetShifts, I) \
   X(targetShifts, 21, targetShifts, P) \
   X(method, 22, method, I) \
   X(methodStage2, 23, methodStage2, I) \
   X(matrix, 24, matrix, P) \
   X(preconditioner, 25, preconditioner, P) \
   X(locking, 26, locking, I) \
   X(numOrthoConst, 27, numOrthoConst, I) \
   X(aNorm, 28, aNorm, D) \
   X(eps, 29, eps, D) \
   X(precondition, 30, precondition, I) \
   X(initSize, 31, initSize, I) \
   X(maxBasisSize, 32, maxBasisSize, I) \
   X(maxBlockSize, 33, maxBlockSize, I) \
   X(maxMatvecs, 34, maxMatvecs, I) \
   X(iseed, 35, iseed, A4) \
   X(printLevel, 36, printLevel, I) \
   X(internalPrecision, 37, internalPrecision, I) \
   X(outputFile, 38, outputFile, P) \
   X(stats_numOuterIterations, 39, stats.numOuterIterations, I) \
   X(stats_numRestarts, 40, stats.numRestarts, I) \
   X(stats_numMatvecs, 41, stats.numMatvecs, I) \
   X(stats_numPreconds, 42, stats.numPreconds, I) \
   X(stats_numGlobalSum, 43, stats.numGlobalSum, I) \
   X(stats_volumeGlobalSum, 44, stats.volumeGlobalSum, I) \
   X(stats_numBroadcast, 45, stats.numBroadcast, I) \
   X(stats_volumeBroadcast, 46, stats.volumeBroadcast, I) \
   X(stats_numOrthoInnerProds, 47, stats.numOrthoInnerProds, D) \
   X(stats_elapsedTime, 48, stats.elapsedTime, D) \
   X(stats_timeMatvec, 49, stats.timeMatvec, D) \
   X(stats_timePrecond, 50, stats.timePrecond, D) \
   X(stats_timeOrtho, 51, stats.timeOrtho, D) \
   X(stats_timeGlobalSum, 52, stats.timeGlobalSum, D) \
   X(stats_timeBroadcast, 53, stats.timeBroadcast, D) \
   X(stats_lockingIssue, 54, stats.lockingIssue, I) \
   X(convTestFun, 55, convTestFun, F) \
   X(convTestFun_type, 56, convTestFun_type, I) \
   X(convtest, 57, convtest, P) \
   X(monitorFun, 58, monitorFun, F) \
   X(monitorFun_type, 59, monitorFun_type, I) \
   X(monitor, 60, monitor, P) \
   X(queue, 61, queue, P) \
   X(profile, 62, profile, S)

typedef enum {
   PRIMME_SVDS_invalid_label = 0,
#define PRIMME_SVDS_LABEL_ENUM_(name, id, path, kind) PRIMME_SVDS_##name = id,
   PRIMME_SVDS_PARAM_TABLE(PRIMME_SVDS_LABEL_ENUM_)
#undef PRIMME_SVDS_LABEL_ENUM_
   PRIMME_SVDS_params_label_end_ = 63
} primme_svds_params_label;

#define PRIMME_DECLARE_SVDS_(prefix, name, SV, VEC, RN) \
   int prefix##name(SV *svals, VEC *svecs, RN *resNorms, primme_svds_params *primme_svds);
#define PRIMME_DECLARE_SVDS_ALL_(name, SV, VEC, RN) \
   PRIMME_DECLARE_SVDS_(, name, SV, VEC, RN) \
   PRIMME_DECLARE_SVDS_(magma_, name, SV, VEC, RN) \
   PRIMME_DECLARE_SVDS_(cublas_, name, SV, VEC, RN)

PRIMME_DECLARE_SVDS_ALL_(hprimme_svds, PRIMME_HALF, PRIMME_HALF, PRIMME_HALF)
PRIMME_DECLARE_SVDS_ALL_(kprimme_svds, PRIMME_HALF, PRIMME_COMPLEX_HALF, PRIMME_HALF)
PRIMME_DECLARE_SVDS_ALL_(sprimme_svds, float, float, float)
PRIMME_DECLARE_SVDS_ALL_(cprimme_svds, float, PRIMME_COMPLEX_FLOAT, float)
PRIMME_DECLARE_SVDS_ALL_(dprimme_svds, double, double, double)
PRIMME_DECLARE_SVDS_ALL_(zprimme_svds, double, PRIMME_COMPLEX_DOUBLE, double)
PRIMME_DECLARE_SVDS_ALL_(hsprimme_svds, float, PRIMME_HALF, float)
PRIMME_DECLARE_SVDS_ALL_(ksprimme_svds, float, PRIMME_COMPLEX_HALF, float)

primme_svds_params *primme_svds_params_create(void);
int primme_svds_params_destroy(primme_svds_params *primme_svds);
void primme_svds_initialize(primme_svds_params *primme_svds);
int primme_svds_set_method(primme_svds_preset_method method,
      primme_preset_method methodStage1, primme_preset_method methodStage2,
      primme_svds_params *primme_svds);
void primme_svds_display_params(primme_svds_params primme_svds);
void primme_svds_free(primme_svds_params *primme_svds);
int primme_svds_get_member(
      primme_svds_params *primme_svds, primme_svds_params_label label, void *value);
int primme_svds_set_member(
      primme_svds_params *primme_svds, primme_svds_params_label label, void *value);
int primme_svds_member_info(primme_svds_params_label *label, const char **label_name,
      primme_type *type, int *arity);
int primme_svds_constant_info(const char *label_name, int *value);
int primme_svds_enum_member_info(
      primme_svds_params_label label, int *value, const char **value_name);

#ifdef __cplusplus
}
#endif

#endif /* PRIMME_SVDS_H */

/* primme_eigs.h -- public eigensolver API of the B200-native PRIMME hot-path library.
 *
 * ABI contract: every enum value, struct field (name, order, type) and entry-point signature
 * below equals the reference's include/primme_eigs.h (enums :47-107, primme_stats :109-135,
 * primme_params :166-253, preset methods :256-273, labels :286-378, entry points :382-477);
 * the reference's tests even fwrite() primme_params (tests/COMMON/ioandtest.c:254-256).
 * The file itself is written from scratch: the label enum and the reflective
 * get/set/member_info tables are generated from one X-macro list (PRIMME_PARAM_TABLE).
 *
 * Which entry points do work here:
 *   dprimme / zprimme                 host arrays + host callbacks; the basis lives in HBM, the
 *                                     callbacks are fed through pinned staging buffers.
 *   cublas_dprimme / cublas_zprimme   the reference's device-pointer contract
 *                                     (examples/ex_eigs_dcublas.c:173-181): evecs and the
 *                                     matrixMatvec/applyPreconditioner blocks are DEVICE pointers,
 *                                     evals/resNorms are host.  primme.queue is ignored.
 *   every other precision / magma_*   return PRIMME_FUNCTION_UNAVAILABLE, like a reference
 *                                     build without that type (src/eigs/primme_c.c:233-242).
 */
#ifndef PRIMME_EIGS_H
#define PRIMME_EIGS_H

#include <stdio.h>
#include "primme.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Which end of the spectrum (or which neighbourhood of targetShifts[]) is wanted. */
typedef enum {
   primme_smallest,
   primme_largest,
   primme_closest_geq,
   primme_closest_leq,
   primme_closest_abs,
   primme_largest_abs
} primme_target;

typedef enum {
   primme_proj_default,
   primme_proj_RR,
   primme_proj_harmonic,
   primme_proj_refined
} primme_projection;

typedef enum {
   primme_init_default,
   primme_init_krylov,
   primme_init_random,
   primme_init_user
} primme_init;

typedef enum {
   primme_full_LTolerance,
   primme_decreasing_LTolerance,
   primme_adaptive_ETolerance,
   primme_adaptive
} primme_convergencetest;

typedef enum {
   primme_event_outer_iteration,
   primme_event_inner_iteration,
   primme_event_restart,
   primme_event_reset,
   primme_event_converged,
   primme_event_locked,
   primme_event_message,
   primme_event_profile
} primme_event;

typedef enum {
   primme_orth_default,
   primme_orth_implicit_I, /* trust V'V = I */
   primme_orth_explicit_I  /* carry V'V explicitly and solve the pencil (H, V'V) */
} primme_orth;

typedef enum {
   primme_op_default,
   primme_op_half,
   primme_op_float,
   primme_op_double,
   primme_op_quad,
   primme_op_int
} primme_op_datatype;

typedef struct primme_stats {
   PRIMME_INT numOuterIterations;
   PRIMME_INT numRestarts;
   PRIMME_INT numMatvecs;
   PRIMME_INT numPreconds;
   PRIMME_INT numGlobalSum;
   PRIMME_INT numBroadcast;
   PRIMME_INT volumeGlobalSum;
   PRIMME_INT volumeBroadcast;
   double flopsDense;
   double numOrthoInnerProds;
   double elapsedTime;
   double timeMatvec;
   double timePrecond;
   double timeOrtho;
   double timeGlobalSum;
   double timeBroadcast;
   double timeDense;
   double estimateMinEVal;
   double estimateMaxEVal;
   double estimateLargestSVal;
   double estimateBNorm;
   double estimateInvBNorm;
   double maxConvTol;
   double estimateResidualError;
   PRIMME_INT lockingIssue;
} primme_stats;

typedef struct JD_projectors {
   int LeftQ;
   int LeftX;
   int RightQ;
   int RightX;
   int SkewQ;
   int SkewX;
} JD_projectors;

typedef struct projection_params {
   primme_projection projection;
} projection_params;

typedef struct correction_params {
   int precondition;
   int robustShifts;
   int maxInnerIterations;
   struct JD_projectors projectors;
   primme_convergencetest convTest;
   double relTolBase;
} correction_params;

typedef struct restarting_params {
   int maxPrevRetain;
} restarting_params;

struct primme_params;

/* Block operator callback: y(:,0:blockSize) = Op * x(:,0:blockSize), column-major. */
typedef void (*primme_block_op_fn)(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy,
      int *blockSize, struct primme_params *primme, int *ierr);

typedef struct primme_params {
   PRIMME_INT n;
   void (*matrixMatvec)(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy, int *blockSize,
         struct primme_params *primme, int *ierr);
   primme_op_datatype matrixMatvec_type;

   void (*applyPreconditioner)(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy,
         int *blockSize, struct primme_params *primme, int *ierr);
   primme_op_datatype applyPreconditioner_type;

   void (*massMatrixMatvec)(void *x, PRIMME_INT *ldx, void *y, PRIMME_INT *ldy,
         int *blockSize, struct primme_params *primme, int *ierr);
   primme_op_datatype massMatrixMatvec_type;

   /* SPMD row partition */
   int numProcs;
   int procID;
   PRIMME_INT nLocal;
   void *commInfo;
   void (*globalSumReal)(void *sendBuf, void *recvBuf, int *count,
         struct primme_params *primme, int *ierr);
   primme_op_datatype globalSumReal_type;
   void (*broadcastReal)(void *buffer, int *count, struct primme_params *primme, int *ierr);
   primme_op_datatype broadcastReal_type;

   int numEvals;
   primme_target target;
   int numTargetShifts;
   double *targetShifts;

   int dynamicMethodSwitch;
   int locking;
   int initSize;
   int numOrthoConst;
   int maxBasisSize;
   int minRestartSize;
   int maxBlockSize;
   PRIMME_INT maxMatvecs;
   PRIMME_INT maxOuterIterations;
   PRIMME_INT iseed[4];
   double aNorm;
   double BNorm;
   double invBNorm;
   double eps;
   primme_orth orth;
   primme_op_datatype internalPrecision;

   int printLevel;
   FILE *outputFile;

   void *matrix;
   void *preconditioner;
   void *massMatrix;
   double *ShiftsForPreconditioner;
   primme_init initBasisMode;
   PRIMME_INT ldevecs;
   PRIMME_INT ldOPs;

   struct projection_params projectionParams;
   struct restarting_params restartingParams;
   struct correction_params correctionParams;
   struct primme_stats stats;

   void (*convTestFun)(double *eval, void *evec, double *rNorm, int *isconv,
         struct primme_params *primme, int *ierr);
   primme_op_datatype convTestFun_type;
   void *convtest;
   void (*monitorFun)(void *basisEvals, int *basisSize, int *basisFlags, int *iblock,
         int *blockSize, void *basisNorms, int *numConverged, void *lockedEvals,
         int *numLocked, int *lockedFlags, void *lockedNorms, int *inner_its, void *LSRes,
         const char *msg, double *time, primme_event *event, struct primme_params *primme,
         int *err);
   primme_op_datatype monitorFun_type;
   void *monitor;
   void *queue;
   const char *profile;
} primme_params;

typedef enum {
   PRIMME_DEFAULT_METHOD,
   PRIMME_DYNAMIC,
   PRIMME_DEFAULT_MIN_TIME,
   PRIMME_DEFAULT_MIN_MATVECS,
   PRIMME_Arnoldi,
   PRIMME_GD,
   PRIMME_GD_plusK,
   PRIMME_GD_Olsen_plusK,
   PRIMME_JD_Olsen_plusK,
   PRIMME_RQI,
   PRIMME_JDQR,
   PRIMME_JDQMR,
   PRIMME_JDQMR_ETol,
   PRIMME_STEEPEST_DESCENT,
   PRIMME_LOBPCG_OrthoBasis,
   PRIMME_LOBPCG_OrthoBasis_Window
} primme_preset_method;

typedef enum { primme_int, primme_double, primme_pointer, primme_string } primme_type;

/* One row per reflective member: X(label suffix, numeric id, lvalue path, kind).
 * kind: I = integer-like (read/written through PRIMME_INT), D = double, P = data pointer,
 *       F = function pointer, S = string, A4 = PRIMME_INT[4].
 * Ids equal the reference's primme_params_label (include/primme_eigs.h:286-378). */
#define PRIMME_PARAM_TABLE(X) \
   X(n, 1, n, I) \
   X(matrixMatvec, 2, matrixMatvec, F) \
   X(matrixMatvec_type, 3, matrixMatvec_type, I) \
   X(applyPreconditioner, 4, applyPreconditioner, F) \
   X(applyPreconditioner_type, 5, applyPreconditioner_type, I) \
   X(massMatrixMatvec, 6, massMatrixMatvec, F) \
   X(massMatrixMatvec_type, 7, massMatrixMatvec_type, I) \
   X(numProcs, 8, numProcs, I) \
   X(procID, 9, procID, I) \
   X(commInfo, 10, commInfo, P) \
   X(nLocal, 11, nLocal, I) \
   X(globalSumReal, 12, globalSumReal, F) \
   X(globalSumReal_type, 13, globalSumReal_type, I) \
   X(broadcastReal, 14, broadcastReal, F) \
   X(broadcastReal_type, 15, broadcastReal_type, I) \
   X(numEvals, 16, numEvals, I) \
   X(target, 17, target, I) \
   X(numTargetShifts, 18, numTargetShifts, I) \
   X(targetShifts, 19, targetShifts, P) \
   X(locking, 20, locking, I) \
   X(initSize, 21, initSize, I) \
   X(numOrthoConst, 22, numOrthoConst, I) \
   X(maxBasisSize, 23, maxBasisSize, I) \
   X(minRestartSize, 24, minRestartSize, I) \
   X(maxBlockSize, 25, maxBlockSize, I) \
   X(maxMatvecs, 26, maxMatvecs, I) \
   X(maxOuterIterations, 27, maxOuterIterations, I) \
   X(iseed, 28, iseed, A4) \
   X(aNorm, 29, aNorm, D) \
   X(BNorm, 30, BNorm, D) \
   X(invBNorm, 31, invBNorm, D) \
   X(eps, 32, eps, D) \
   X(orth, 33, orth, I) \
   X(internalPrecision, 34, internalPrecision, I) \
   X(printLevel, 35, printLevel, I) \
   X(outputFile, 36, outputFile, P) \
   X(matrix, 37, matrix, P) \
   X(massMatrix, 38, massMatrix, P) \
   X(preconditioner, 39, preconditioner, P) \
   X(ShiftsForPreconditioner, 40, ShiftsForPreconditioner, P) \
   X(initBasisMode, 41, initBasisMode, I) \
   X(projectionParams_projection, 42, projectionParams.projection, I) \
   X(restartingParams_maxPrevRetain, 43, restartingParams.maxPrevRetain, I) \
   X(correctionParams_precondition, 44, correctionParams.precondition, I) \
   X(correctionParams_robustShifts, 45, correctionParams.robustShifts, I) \
   X(correctionParams_maxInnerIterations, 46, correctionParams.maxInnerIterations, I) \
   X(correctionParams_projectors_LeftQ, 47, correctionParams.projectors.LeftQ, I) \
   X(correctionParams_projectors_LeftX, 48, correctionParams.projectors.LeftX, I) \
   X(correctionParams_projectors_RightQ, 49, correctionParams.projectors.RightQ, I) \
   X(correctionParams_projectors_RightX, 50, correctionParams.projectors.RightX, I) \
   X(correctionParams_projectors_SkewQ, 51, correctionParams.projectors.SkewQ, I) \
   X(correctionParams_projectors_SkewX, 52, correctionParams.projectors.SkewX, I) \
   X(correctionParams_convTest, 53, correctionParams.convTest, I) \
   X(correctionParams_relTolBase, 54, correctionParams.relTolBase, D) \
   X(stats_numOuterIterations, 55, stats.numOuterIterations, I) \
   X(stats_numRestarts, 56, stats.numRestarts, I) \
   X(stats_numMatvecs, 57, stats.numMatvecs, I) \
   X(stats_numPreconds, 58, stats.numPreconds, I) \
   X(stats_numGlobalSum, 59, stats.numGlobalSum, I) \
   X(stats_volumeGlobalSum, 60, stats.volumeGlobalSum, I) \
   X(stats_numBroadcast, 61, stats.numBroadcast, I) \
   X(stats_volumeBroadcast, 62, stats.volumeBroadcast, I) \
   X(stats_flopsDense, 63, stats.flopsDense, D) \
   X(stats_numOrthoInnerProds, 64, stats.numOrthoInnerProds, D) \
   X(stats_elapsedTime, 65, stats.elapsedTime, D) \
   X(stats_timeMatvec, 66, stats.timeMatvec, D) \
   X(stats_timePrecond, 67, stats.timePrecond, D) \
   X(stats_timeOrtho, 68, stats.timeOrtho, D) \
   X(stats_timeGlobalSum, 69, stats.timeGlobalSum, D) \
   X(stats_timeBroadcast, 70, stats.timeBroadcast, D) \
   X(stats_timeDense, 71, stats.timeDense, D) \
   X(stats_estimateMinEVal, 72, stats.estimateMinEVal, D) \
   X(stats_estimateMaxEVal, 73, stats.estimateMaxEVal, D) \
   X(stats_estimateLargestSVal, 74, stats.estimateLargestSVal, D) \
   X(stats_estimateBNorm, 75, stats.estimateBNorm, D) \
   X(stats_estimateInvBNorm, 76, stats.estimateInvBNorm, D) \
   X(stats_maxConvTol, 77, stats.maxConvTol, D) \
   X(stats_lockingIssue, 78, stats.lockingIssue, I) \
   X(dynamicMethodSwitch, 79, dynamicMethodSwitch, I) \
   X(convTestFun, 80, convTestFun, F) \
   X(convTestFun_type, 81, convTestFun_type, I) \
   X(convtest, 82, convtest, P) \
   X(ldevecs, 83, ldevecs, I) \
   X(ldOPs, 84, ldOPs, I) \
   X(monitorFun, 85, monitorFun, F) \
   X(monitorFun_type, 86, monitorFun_type, I) \
   X(monitor, 87, monitor, P) \
   X(queue, 88, queue, P) \
   X(profile, 89, profile, S)

typedef enum {
   PRIMME_invalid_label = 0,
#define PRIMME_LABEL_ENUM_(name, id, path, kind) PRIMME_##name = id,
   PRIMME_PARAM_TABLE(PRIMME_LABEL_ENUM_)
#undef PRIMME_LABEL_ENUM_
   PRIMME_params_label_end_ = 90
} primme_params_label;

/* ---- solvers ---- */
#define PRIMME_DECLARE_SOLVER_(prefix, name, EV, VEC, RN) \
   int prefix##name(EV *evals, VEC *evecs, RN *resNorms, primme_params *primme);
#define PRIMME_DECLARE_SOLVERS_(name, EV, VEC, RN) \
   PRIMME_DECLARE_SOLVER_(, name, EV, VEC, RN) \
   PRIMME_DECLARE_SOLVER_(magma_, name, EV, VEC, RN) \
   PRIMME_DECLARE_SOLVER_(cublas_, name, EV, VEC, RN)

/* Hermitian problems */
PRIMME_DECLARE_SOLVERS_(hprimme, PRIMME_HALF, PRIMME_HALF, PRIMME_HALF)
PRIMME_DECLARE_SOLVERS_(kprimme, PRIMME_HALF, PRIMME_COMPLEX_HALF, PRIMME_HALF)
PRIMME_DECLARE_SOLVERS_(sprimme, float, float, float)
PRIMME_DECLARE_SOLVERS_(cprimme, float, PRIMME_COMPLEX_FLOAT, float)
PRIMME_DECLARE_SOLVERS_(dprimme, double, double, double)
PRIMME_DECLARE_SOLVERS_(zprimme, double, PRIMME_COMPLEX_DOUBLE, double)
PRIMME_DECLARE_SOLVERS_(hsprimme, float, PRIMME_HALF, float)
PRIMME_DECLARE_SOLVERS_(ksprimme, float, PRIMME_COMPLEX_HALF, float)
/* normal (non-Hermitian) problems: declared for link compatibility, unavailable here */
PRIMME_DECLARE_SOLVERS_(kprimme_normal, PRIMME_COMPLEX_HALF, PRIMME_COMPLEX_HALF, PRIMME_HALF)
PRIMME_DECLARE_SOLVERS_(cprimme_normal, PRIMME_COMPLEX_FLOAT, PRIMME_COMPLEX_FLOAT, float)
PRIMME_DECLARE_SOLVERS_(zprimme_normal, PRIMME_COMPLEX_DOUBLE, PRIMME_COMPLEX_DOUBLE, double)
PRIMME_DECLARE_SOLVERS_(kcprimme_normal, PRIMME_COMPLEX_FLOAT, PRIMME_COMPLEX_HALF, float)

/* ---- parameter handling ---- */
primme_params *primme_params_create(void);
int primme_params_destroy(primme_params *primme);
void primme_initialize(primme_params *primme);
int primme_set_method(primme_preset_method method, primme_params *params);
void primme_display_params(primme_params primme);
void primme_free(primme_params *primme);
int primme_get_member(primme_params *primme, primme_params_label label, void *value);
int primme_set_member(primme_params *primme, primme_params_label label, void *value);
int primme_member_info(
      primme_params_label *label, const char **label_name, primme_type *type, int *arity);
int primme_constant_info(const char *label_name, int *value);
int primme_enum_member_info(primme_params_label label, int *value, const char **value_name);

#ifdef __cplusplus
}
#endif

#endif /* PRIMME_EIGS_H */

// pb200_internal.cuh -- shared internals of the sm_100a kernel layer (context, error macros,
// deterministic two-stage panel reduction).  Everything public is declared in
// include/primme_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/primme_b200.h"

#define PB_CUDA(call)                                                                  \
   do {                                                                                \
      cudaError_t e_ = (call);                                                         \
      if (e_ != cudaSuccess) {                                                         \
         fprintf(stderr, "primme_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), \
               __FILE__, __LINE__);                                                    \
         return PB200_ERR_CUDA;                                                        \
      }                                                                                \
   } while (0)

#define PB_CHK(call)            \
   do {                         \
      int r_ = (call);          \
      if (r_ != 0) return r_;   \
   } while (0)

// capacity (doubles) of the coefficient block embedded in the kernel-argument structs
#ifndef PB_COEF_MAX
#define PB_COEF_MAX 1024
#endif

// small coefficient matrices (C|Y of the ortho update, h|theta of VWXR) passed by value in the
// kernel parameter space: a separate __grid_constant__ parameter, so that only this block has its
// address taken and the scalar arguments stay plain constant-bank operands
struct PbCoef {
   double v[PB_COEF_MAX];
};

struct pb200_ctx {
   int device;
   int num_sms;
   cudaStream_t stream;
   // staging for small operands and panels
   double *h_pinned;      // pinned host buffer (mapped: kernels may write panels into it directly)
   double *d_hpinned;     // device alias of h_pinned
   double *h_tagged;      // mapped pinned (value, sequence number) pairs of the polled panels
   double *d_htagged;     // device alias of h_tagged
   size_t tagged_cap;     // capacity in pairs
   long long seq;
   int no_poll;           // 1: classic memcpy + stream synchronise for panels
   double *d_small;       // device buffer for coefficient blocks (h, C, Y, theta, perms)
   double *d_panel;       // device buffer for reduced panels
   double *d_partials;    // per-CTA partial panels
   size_t small_cap;      // capacity (doubles) of h_pinned / d_small / d_panel
   size_t partials_cap;   // capacity (doubles) of d_partials
   void *d_scratch;       // growable scratch (permute etc.)
   size_t scratch_cap;
   int64_t launches;
   int use_ws;            // warp-specialised ortho sweep (v3) where eligible
   int use_tma_vwxr;      // same switch for the VWXR kernel alone
   int use_tma_spmm;      // persistent bulk-copy SpMM (v2)
   int use_wide;          // v3 wide VWXR kernel for the restart sweep
   int vwxr_cand_tma;     // 1: TMA-staged (v2) kernel also for the candidates sweep
   int coef_inline;       // small coefficient matrices travel as kernel parameters (no H2D copy)
   PbCoef coef;           // host staging of that block for the next launch
   int use_narrow;        // v3 narrow VWXR kernel for the candidates sweep (default off: the LDG kernel is faster)
   int ortho_2cta;        // prefer 2 CTAs/SM x 2 stages over 1 CTA/SM x 4 stages in the ortho sweep
   int use_tma;           // 1: TMA-staged kernels where eligible (default), 0: LDG kernels only
   // optional per-kernel-kind CUDA-event timing (bench.py's roofline numbers)
   int prof_on;
   int prof_pending;
   cudaEvent_t *prof_ev;   // 2 * PB_PROF_RING events
   int *prof_kind;
   double prof_ms[8];
   double prof_bytes[8];
   int64_t prof_cnt[8];
   // NCCL (dlopen'ed lazily)
   void *comm;
   int nranks, rank;
   int owns_comm;
};

// grow helpers (host side)
int pb_ensure_partials(pb200_ctx *ctx, size_t doubles);
int pb_ensure_scratch(pb200_ctx *ctx, size_t bytes);
int pb_ensure_small(pb200_ctx *ctx, size_t doubles);
int pb_ensure_tagged(pb200_ctx *ctx, size_t elems);
// reduce ctx->d_partials [nparts x cnt] -> ctx->d_panel [cnt] (fixed order), optional NCCL
// allreduce, copy to h_pinned and synchronize.  Result readable at ctx->h_pinned[0..cnt).
int pb_finish_panel(pb200_ctx *ctx, int nparts, int cnt);
int pb_nccl_allreduce_dev(pb200_ctx *ctx, double *dbuf, int count);
int pb_nccl_allgatherv_cols(pb200_ctx *ctx, const double *X, int64_t ldx, double *Y, int64_t ldy,
      const int64_t *counts, const int64_t *displs, int ncols);

// kernel kinds for profiling
enum { PB_K_SPMM = 0, PB_K_ORTHO = 1, PB_K_VWXR = 2, PB_K_UTIL = 3, PB_K_REDUCE = 4, PB_K_NKINDS = 5 };
#define PB_PROF_RING 2048
int pb_prof_begin(pb200_ctx *ctx, int kind);              // returns slot or -1
void pb_prof_end(pb200_ctx *ctx, int slot, double bytes); // algorithmic bytes of the launch
int pb_prof_flush(pb200_ctx *ctx);

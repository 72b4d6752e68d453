// pb200_internal.cuh -- shared internals of the sm_100a kernel layer (context, error macros,
// deterministic two-stage panel reduction).  Everything public is declared in
// include/primme_b200.h.
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is resolved at run time, no libcuda link)
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

#define PB_WS_SLOTS 4
#define PB_MAX_PEERS 8
#define PB_XCHG_SLOTS 4
#define PB_XCHG_CAP 4096   // pairs per (slot, rank)

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
   double *d_gpart;       // per-CTA-group partial panels (second level of the in-kernel reduction)
   size_t gpart_cap;
   unsigned int *d_counters;  // arrival counters of the in-kernel reduction (all zero between launches)
   int fused_finish;      // 1 (default): the sweep kernels reduce and deliver their own panel
   size_t small_cap;      // capacity (doubles) of h_pinned / d_small / d_panel
   size_t partials_cap;   // capacity (doubles) of d_partials
   void *d_scratch;       // growable scratch (permute etc.)
   size_t scratch_cap;
   int64_t launches;
   int use_tma_vwxr;      // same switch for the VWXR kernel alone
   int use_tma_spmm;      // persistent bulk-copy SpMM (v2)
   int l2_window;         // a persisting access-policy window is set on the stream
   int sweep_alternate;   // 1 (default): consecutive TMA sweeps walk the rows in opposite directions (L2 reuse)
   int sweep_rev;         // direction of the last sweep
   int spmm_evict_first;  // L2 evict_first policy on the CSR stream of the row-major SpMM (default 1)
   int spmm_v3;           // gather layout for b >= 2: 0 column-major (v2), 1 row-major copy (v3), 2 timed once per matrix (default)
   int use_wide;          // v3 wide VWXR kernel for the restart sweep
   int use_mma_vwxr;      // tensor-map TMA + DMMA VWXR kernel (default)
   int fuse_gram;         // candidates sweep also delivers the first Gram panel of the block ortho
   int coef_inline;       // small coefficient matrices travel as kernel parameters (no H2D copy)
   PbCoef coef;           // host staging of that block for the next launch
   int ortho_exact;       // specialised (exact tile count) instances of the ortho sweep
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
   // basis workspace kept between solves (pb200_ctx_workspace)
   void *ws_ptr[PB_WS_SLOTS];
   size_t ws_bytes[PB_WS_SLOTS];
   // peer-memory panel exchange (see PbFin)
   double2 *xchg_local;
   double2 *xchg_peer[PB_MAX_PEERS];
   int peer_on;
   // NCCL (dlopen'ed lazily)
   void *comm;
   int nranks, rank;
   int owns_comm;
};

// ---------------------------------------------------------------- in-kernel panel finish ----
// Every sweep kernel leaves one partial panel per CTA (or per tile group of a CTA).  Instead of a
// second launch, the CTAs finish the panel themselves in two fixed-order levels: the last CTA of
// each group of PB_FIN_GROUP consecutive CTAs to arrive sums its group's partials (slot order),
// the last group to arrive sums the group panels (group order) and delivers the result -- either
// (value, sequence number) pairs straight into mapped pinned host memory, or plain doubles into a
// device buffer in front of an NCCL all-reduce.  The order of every sum depends on the launch
// shape only, never on the arrival order => bitwise reproducible panels.
#define PB_FIN_GROUP 16
#define PB_FIN_MAXGROUPS 64
// Peer-memory panel exchange (one node, one process per GPU): the all-reduce of a panel is part of
// the kernel that produced it.  The CTA that finishes the local panel stores every element as a
// (value, sequence number) pair straight into the exchange buffer of EVERY rank over NVLink (one
// 16-byte store per peer), waits until the pairs of all ranks for this sequence number have landed
// in its own buffer, sums them in rank order (identical bits on every rank) and delivers the result
// to the host like a single-GPU panel.  Slots rotate with the sequence number; a rank cannot get
// two panels ahead of a peer (it needs the peer's previous panel to proceed), so 4 slots suffice.
// ASSUMPTION (platform contract of this protocol, x86-64 host + NVLink / PCIe on one node, checked by every
// multi-GPU and host-contract test run): a naturally aligned 16-byte store (st.volatile.global.v2.f64 to peer or
// mapped pinned memory) is observed as a whole -- the reader never sees the new tag with the old value.  CUDA
// guarantees single-copy atomicity for 8 bytes only; the 16-byte pair travels as one NVLink / PCIe write of an
// aligned 16-byte quantity on these platforms.  A platform without that property needs an LL-style layout (the
// tag repeated next to every 8-byte word, as NCCL's LL protocol does); the host reader already orders the value
// load after the tag load with an acquire fence (ctx.cu:pb_poll_tagged) for weakly ordered hosts.
struct PbFin {
   double *partials;         // [nparts][cnt], CTA c owns slots [c*ppc, (c+1)*ppc)
   double *gpart;            // [ngroups][cnt]
   unsigned int *counters;   // [0]: groups done, [1+g]: CTAs of group g done
   double *out;              // tagged pairs (tag != 0) or plain doubles
   long long tag;
   int cnt;                  // 0: no in-kernel finish
   int ppc;                  // partial slots per CTA
   int nranks, rank, slot;   // nranks > 1: peer-memory exchange before the delivery
   double2 *peer[PB_MAX_PEERS];  // exchange buffers of all ranks (peer[rank] is the local one)
};

#ifdef __CUDACC__
__device__ __forceinline__ void pb_fin_bar(int id, int nthreads) {
   asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void pb_fin_store(const PbFin &f, int e, double s) {
   if (f.nranks > 1) {
      const double tg = __longlong_as_double(f.tag);
      const size_t mine = ((size_t)f.slot * PB_MAX_PEERS + f.rank) * PB_XCHG_CAP + e;
      for (int p = 0; p < f.nranks; p++)
         asm volatile("st.volatile.global.v2.f64 [%0], {%1, %2};" ::"l"(f.peer[p] + mine), "d"(s), "d"(tg) : "memory");
      double tot = 0.0;
      for (int r = 0; r < f.nranks; r++) {
         const double2 *src = f.peer[f.rank] + ((size_t)f.slot * PB_MAX_PEERS + r) * PB_XCHG_CAP + e;
         double vx, vy;
         unsigned long long spins = 0;
         do {
            asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(vx), "=d"(vy) : "l"(src) : "memory");
         } while (__double_as_longlong(vy) != f.tag && ++spins < (1ull << 27));  // bounded: a lost peer must not hang the GPU
         tot += (__double_as_longlong(vy) == f.tag) ? vx : __longlong_as_double(0x7ff8000000000000ll);
      }
      s = tot;
   }
   if (f.tag)
      reinterpret_cast<double2 *>(f.out)[e] = make_double2(s, __longlong_as_double(f.tag));
   else
      f.out[e] = s;
}
// sum of np slots of one element, 8 independent chains in a fixed interleaving
__device__ __forceinline__ double pb_fin_sum(const double *src, int np, int cnt) {
   double s[8];
#pragma unroll
   for (int i = 0; i < 8; i++) s[i] = 0.0;
   int p = 0;
   for (; p + 8 <= np; p += 8) {
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] += __ldcg(src + (size_t)(p + i) * cnt);
   }
#pragma unroll
   for (int i = 0; i < 8; i++)
      if (p + i < np) s[i] += __ldcg(src + (size_t)(p + i) * cnt);
   return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}
// Called by threads tid = 0..nthr-1 of EVERY CTA (nthr a multiple of 32, the same set of threads
// that wrote the CTA's partial slots) after those stores; bar_id: a named barrier free for them;
// flag: one int of shared memory (dynamic: a static variable would change the kernels' occupancy).
__device__ __forceinline__ void pb_finish_device(const PbFin &f, int tid, int nthr, int bar_id, volatile int *flag) {
   if (f.cnt <= 0) return;
   const int ncta = gridDim.x;
   const int ngroups = (ncta + PB_FIN_GROUP - 1) / PB_FIN_GROUP;
   const int g = blockIdx.x / PB_FIN_GROUP;
   const int gsize = (ncta - g * PB_FIN_GROUP) < PB_FIN_GROUP ? (ncta - g * PB_FIN_GROUP) : PB_FIN_GROUP;
   __threadfence();  // this thread's partial stores are visible device-wide
   pb_fin_bar(bar_id, nthr);
   if (tid == 0) {
      __threadfence();
      const unsigned int t = atomicAdd(&f.counters[1 + g], 1u);
      *flag = (t == (unsigned int)(gsize - 1));
   }
   pb_fin_bar(bar_id, nthr);
   if (!*flag) return;
   __threadfence();
   {
      const double *src = f.partials + (size_t)g * PB_FIN_GROUP * f.ppc * f.cnt;
      const int np = gsize * f.ppc;
      for (int e = tid; e < f.cnt; e += nthr) {
         const double s = pb_fin_sum(src + e, np, f.cnt);
         if (ngroups == 1)
            pb_fin_store(f, e, s);
         else
            f.gpart[(size_t)g * f.cnt + e] = s;
      }
   }
   if (ngroups == 1) {
      if (tid == 0) f.counters[1 + g] = 0u;
      return;
   }
   __threadfence();
   pb_fin_bar(bar_id, nthr);
   if (tid == 0) {
      __threadfence();
      f.counters[1 + g] = 0u;
      const unsigned int t = atomicAdd(&f.counters[0], 1u);
      *flag = (t == (unsigned int)(ngroups - 1));
   }
   pb_fin_bar(bar_id, nthr);
   if (!*flag) return;
   __threadfence();
   for (int e = tid; e < f.cnt; e += nthr) pb_fin_store(f, e, pb_fin_sum(f.gpart + e, ngroups, f.cnt));
   if (tid == 0) f.counters[0] = 0u;
}
#endif

// ------------------------------------------------------------------ CSR matrix (spmm.cu) ----
struct pb200_csr {
   int64_t nrows, ncols, nnz;
   int64_t *d_rowptr;  // nrows+1, 0-based
   int32_t *d_colind;  // 0-based
   double *d_vals;
   int is_complex;
   // schedule
   int nblocks;
   int64_t *d_blk_row0;  // nblocks+1 : first row of each block (long-row chunks repeat the row)
   int64_t *d_blk_nz0;   // nblocks   : first nonzero handled by the block
   int32_t *d_blk_nnz;   // nblocks   : nonzeros handled by the block
   int32_t *d_blk_kind;  // 0 = group of whole rows, 1 = chunk of one long row
   int nlong;            // number of long-row chunks
   int lpr;              // lanes per row for group blocks
   double *d_long_part;  // partial sums of long-row chunks [nlongchunks][16]
   int32_t *d_long_slot; // per block: slot index in d_long_part (kind 1) else -1
   // fix-up list for long rows
   int nlongrows;
   int64_t *d_lr_row;    // row id
   int32_t *d_lr_slot0;  // first slot
   int32_t *d_lr_nslots; // number of chunks
   pb200_csr *T;
   // row-major gather buffer of the v3 kernel: G[col][bp] = X(col, 0:b) (packed per call)
   double *d_G;
   size_t G_cap;  // doubles
   int pooled;            // device arrays live in the context's matrix pool (not freed by pb200_csr_destroy)
   int layout_choice[4];  // per block width 1/2/4/8: 0 not timed yet, 1 column-major gathers (v2), 2 row-major (v3), 3 windowed (v4)
   // windowed right-hand sides (v4, spmm.cu): row blocks of <= 512 rows, per block the runs of 32-column
   // segments its nonzeros touch and 16-bit window offsets instead of column indices
   int win_state;         // 0 not analysed, 1 usable, -1 not usable (no locality, long rows, complex)
   int win_nblocks, win_maxseg;
   int64_t *d_w_row0;     // win_nblocks + 1
   int32_t *d_w_nrun;     // win_nblocks
   void *d_w_run;         // win_nblocks x 32 (int2: first segment, window slot << 16 | segments)
   uint16_t *d_wcol;      // nnz
   int *d_w_state;        // [0] not usable, [1] max segments per block
};

// halo protocol of the row-sharded operator (dist.cu) seen by the SpMM kernel: wait until every
// source rank in src_mask has flagged block `seq` in this rank's flags[] before the first gather;
// when the grid is done, store `seq` to ack[r] (peer memory) for every r in ack_mask
struct PbSpSync {
   const unsigned long long *flags;
   unsigned long long seq;
   unsigned int src_mask, ack_mask;
   unsigned int *counter;
   unsigned long long *ack[PB_MAX_PEERS];
};
int pb_spmm_bp(int b);
int pb_spmm_pack(pb200_ctx *ctx, const void *X, int64_t ldx, int64_t nrows, int b, int bp, int is_complex, double *G);
int pb_spmm_gathered(pb200_ctx *ctx, const pb200_csr *A, const double *G, int bp, void *Y, int64_t ldy, int b,
      const PbSpSync *sync);

// grow helpers (host side)
int pb_ensure_partials(pb200_ctx *ctx, size_t doubles);
int pb_ensure_scratch(pb200_ctx *ctx, size_t bytes);
int pb_ensure_small(pb200_ctx *ctx, size_t doubles);
int pb_ensure_tagged(pb200_ctx *ctx, size_t elems);
// reduce ctx->d_partials [nparts x cnt] -> ctx->d_panel [cnt] (fixed order), optional NCCL
// allreduce, copy to h_pinned and synchronize.  Result readable at ctx->h_pinned[0..cnt).
int pb_finish_panel(pb200_ctx *ctx, int nparts, int cnt);
// in-kernel finish: fill `f` for a launch of `grid` CTAs with `ppc` partial slots of `cnt` doubles
// each (allocates partials / group panels, picks the destination, takes a sequence number); after
// the launch pb_collect_panel waits for the panel (NCCL all-reduce first when sharded) and leaves
// it in ctx->h_pinned[0..cnt).  Returns 1 from pb_fin_prepare when the launch shape is not
// covered (fall back to pb_finish_panel).
int pb_fin_prepare(pb200_ctx *ctx, int grid, int ppc, int cnt, PbFin *f);
int pb_collect_panel(pb200_ctx *ctx, const PbFin *f);
// all-zero contribution of a rank without local rows; returns 1 when the peer path is not active
// (the caller then uses the NCCL path), 0 with the reduced panel in ctx->h_pinned[0..cnt)
int pb_fin_contribute_zeros(pb200_ctx *ctx, int cnt);
int pb_nccl_allreduce_dev(pb200_ctx *ctx, double *dbuf, int count);
int pb_nccl_allgather(pb200_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank);
int pb_nccl_group(pb200_ctx *ctx, int start);
int pb_nccl_send(pb200_ctx *ctx, const void *buf, size_t bytes, int peer);
int pb_nccl_recv(pb200_ctx *ctx, void *buf, size_t bytes, int peer);
int pb_nccl_allgatherv_cols(pb200_ctx *ctx, const double *X, int64_t ldx, double *Y, int64_t ldy,
      const int64_t *counts, const int64_t *displs, int ncols);

// 2-D tensor map over a column-major fp64 matrix (rows x cols, leading dimension ld): dimension 0
// = rows (contiguous), dimension 1 = columns; box = box_rows x box_cols.  Returns 1 when the
// driver entry point is unavailable or the shape is not encodable (callers fall back to the LDG
// kernels).
int pb_tensor_map_2d(CUtensorMap *tm, const double *base, int64_t rows, int cols, int64_t ld,
      int box_rows, int box_cols);

// kernel kinds for profiling
enum { PB_K_SPMM = 0, PB_K_ORTHO = 1, PB_K_VWXR = 2, PB_K_UTIL = 3, PB_K_REDUCE = 4, PB_K_NKINDS = 5 };
#define PB_PROF_RING 2048
int pb_prof_begin(pb200_ctx *ctx, int kind);              // returns slot or -1
void pb_prof_end(pb200_ctx *ctx, int slot, double bytes); // algorithmic bytes of the launch
int pb_prof_flush(pb200_ctx *ctx);

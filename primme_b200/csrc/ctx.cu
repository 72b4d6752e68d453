// ctx.cu -- device context, memory helpers, deterministic panel reduction and the NCCL hook.
// Replaces the per-call cudaMalloc/cudaMemcpy2D/cudaDeviceSynchronize pattern of the reference's
// cuBLAS back end (reference src/linalg/cublas_wrapper.c:187-232,335-392): one pinned staging
// buffer, one device coefficient buffer and one partials buffer live for the whole solve.
#include "pb200_internal.cuh"
#include <time.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

extern "C" int pb200_device_count(void) {
   int n = 0;
   if (cudaGetDeviceCount(&n) != cudaSuccess) {
      cudaGetLastError();
      return 0;
   }
   return n;
}

extern "C" int pb200_ctx_create(pb200_ctx **out, int device) {
   *out = NULL;
   if (pb200_device_count() <= 0) {
      fprintf(stderr, "primme_b200: no CUDA device visible -- this library has no CPU path\n");
      return PB200_ERR_NO_DEVICE;
   }
   if (device >= 0) PB_CUDA(cudaSetDevice(device));
   pb200_ctx *ctx = (pb200_ctx *)calloc(1, sizeof(pb200_ctx));
   if (!ctx) return PB200_ERR_ALLOC;
   PB_CUDA(cudaGetDevice(&ctx->device));
   cudaDeviceProp prop;
   PB_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
   ctx->num_sms = prop.multiProcessorCount;
   PB_CUDA(cudaStreamCreate(&ctx->stream));  // blocking: ordered with default-stream work of callbacks
   ctx->small_cap = 1 << 16; // 64 Ki doubles = 512 KB
   PB_CUDA(cudaHostAlloc((void **)&ctx->h_pinned, ctx->small_cap * sizeof(double), cudaHostAllocMapped));
   PB_CUDA(cudaHostGetDevicePointer((void **)&ctx->d_hpinned, ctx->h_pinned, 0));
   ctx->h_tagged = NULL, ctx->d_htagged = NULL, ctx->tagged_cap = 0;
   PB_CHK(pb_ensure_tagged(ctx, 4096));
   PB_CUDA(cudaMalloc((void **)&ctx->d_small, ctx->small_cap * sizeof(double)));
   PB_CUDA(cudaMalloc((void **)&ctx->d_panel, ctx->small_cap * sizeof(double)));
   ctx->partials_cap = 0;
   ctx->d_partials = NULL;
   ctx->nranks = 1;
   ctx->no_poll = getenv("PB200_NO_POLL") ? 1 : 0;
   ctx->fused_finish = getenv("PB200_NO_FUSED_FINISH") ? 0 : 1;
   PB_CUDA(cudaMalloc((void **)&ctx->d_counters, (1 + PB_FIN_MAXGROUPS) * sizeof(unsigned int)));
   PB_CUDA(cudaMemset(ctx->d_counters, 0, (1 + PB_FIN_MAXGROUPS) * sizeof(unsigned int)));
   ctx->coef_inline = getenv("PB200_NO_INLINE_COEF") ? 0 : 1;
   ctx->use_wide = getenv("PB200_NO_WIDE") ? 0 : 1;
   ctx->use_mma_vwxr = getenv("PB200_NO_MMA_VWXR") ? 0 : 1;
   ctx->fuse_gram = getenv("PB200_NO_FUSE_GRAM") ? 0 : 1;
   ctx->ortho_2cta = getenv("PB200_ORTHO_2CTA") ? atoi(getenv("PB200_ORTHO_2CTA")) : 0;
   ctx->use_tma = getenv("PB200_NO_TMA") ? 0 : 1;
   ctx->ortho_exact = getenv("PB200_NO_ORTHO_EXACT") ? 0 : 1;
   ctx->use_tma_vwxr = (getenv("PB200_NO_TMA") || getenv("PB200_NO_TMA_VWXR")) ? 0 : 1;
   ctx->use_tma_spmm = (getenv("PB200_NO_TMA") || getenv("PB200_NO_TMA_SPMM")) ? 0 : 1;
   ctx->sweep_alternate = getenv("PB200_NO_ALTERNATE") ? 0 : 1;
   ctx->spmm_evict_first = getenv("PB200_NO_EVICT_FIRST") ? 0 : 1;
   ctx->spmm_v3 = getenv("PB200_SPMM_V3") ? atoi(getenv("PB200_SPMM_V3")) : 2;
   *out = ctx;
   return 0;
}

extern "C" int pb200_ctx_destroy(pb200_ctx *ctx) {
   if (!ctx) return 0;
   cudaStreamSynchronize(ctx->stream);
   if (ctx->comm && ctx->owns_comm) pb200_ctx_comm_free(ctx);
   cudaFreeHost(ctx->h_pinned);
   cudaFreeHost(ctx->h_tagged);
   cudaFree(ctx->d_small);
   cudaFree(ctx->d_panel);
   cudaFree(ctx->d_partials);
   cudaFree(ctx->d_gpart);
   cudaFree(ctx->d_counters);
   cudaFree(ctx->d_scratch);
   for (int s = 0; s < PB_WS_SLOTS; s++) cudaFree(ctx->ws_ptr[s]);
   for (int p = 0; p < PB_MAX_PEERS; p++)
      if (ctx->peer_on && p != ctx->rank && ctx->xchg_peer[p]) cudaIpcCloseMemHandle(ctx->xchg_peer[p]);
   cudaFree(ctx->xchg_local);
   if (ctx->prof_ev) {
      for (int i = 0; i < 2 * PB_PROF_RING; i++) cudaEventDestroy(ctx->prof_ev[i]);
      free(ctx->prof_ev), free(ctx->prof_kind);
   }
   cudaStreamDestroy(ctx->stream);
   free(ctx);
   return 0;
}

// Start of a solve on a long-lived context: the alternating sweep direction restarts, so that repeated
// solves of the same problem visit the rows in the same order (bitwise identical panels, same counts)
extern "C" int pb200_ctx_begin_solve(pb200_ctx *ctx) {
   ctx->sweep_rev = 0;
   return 0;
}

// Keep the head of the basis resident in L2 across sweeps: every panel kernel re-reads V(:,0:m) from its
// first column on, so the leading part of the array (whole columns) is given a persisting access-policy
// window on the context's stream (cudaAccessPropertyPersisting; TMA loads honour it).  `bytes` = size of the
// array (0 removes the window and releases the carve-out).  Returns the number of bytes set aside
// (PB200_L2_PERSIST_MB overrides the size, 0 switches it off).
extern "C" int64_t pb200_ctx_l2_persist(pb200_ctx *ctx, const void *ptr, size_t bytes) {
   static int max_persist = -1, max_window = -1;
   if (max_persist < 0) {
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
      cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
      if (getenv("PB200_DEBUG"))
         fprintf(stderr, "primme_b200: L2 persisting max %d MB, access-policy window max %d MB\n", max_persist >> 20,
               max_window >> 20);
   }
   // default: 48 MB of a basis that does not fit the L2 anyway (measured on C2, n = 10^6: ortho sweeps 4.44 ->
   // 4.65 TB/s, solve 635 -> 622 ms; 79 MB, the maximum, starves the SpMM's gathers: 644 ms)
   size_t want = getenv("PB200_L2_PERSIST_MB") ? (size_t)atoi(getenv("PB200_L2_PERSIST_MB")) << 20
                                               : (bytes > ((size_t)96 << 20) ? (size_t)48 << 20 : 0);
   if (bytes == 0 || !ptr) want = 0;
   if (want > bytes) want = bytes;
   if (want > (size_t)max_persist) want = (size_t)max_persist;
   if (want > (size_t)max_window) want = (size_t)max_window;
   cudaStreamAttrValue attr;
   memset(&attr, 0, sizeof(attr));
   if (want == 0) {
      if (!ctx->l2_window) return 0;
      attr.accessPolicyWindow.num_bytes = 0;
      cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
      cudaCtxResetPersistingL2Cache();
      ctx->l2_window = 0;
      cudaGetLastError();
      return 0;
   }
   if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) {
      cudaGetLastError();
      return 0;
   }
   attr.accessPolicyWindow.base_ptr = const_cast<void *>(ptr);
   attr.accessPolicyWindow.num_bytes = want;
   attr.accessPolicyWindow.hitRatio = 1.0f;
   attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
   attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
   if (cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) {
      cudaGetLastError();
      return 0;
   }
   ctx->l2_window = 1;
   return (int64_t)want;
}

extern "C" int pb200_ctx_sync(pb200_ctx *ctx) {
   PB_CUDA(cudaStreamSynchronize(ctx->stream));
   return 0;
}
extern "C" void *pb200_ctx_stream(pb200_ctx *ctx) { return (void *)ctx->stream; }
extern "C" int64_t pb200_ctx_launches(pb200_ctx *ctx) { return ctx->launches; }
extern "C" int pb200_ctx_nranks(pb200_ctx *ctx) { return ctx->nranks; }

int pb_ensure_partials(pb200_ctx *ctx, size_t doubles) {
   if (doubles <= ctx->partials_cap) return 0;
   PB_CUDA(cudaStreamSynchronize(ctx->stream));
   if (ctx->d_partials) PB_CUDA(cudaFree(ctx->d_partials));
   size_t cap = doubles + doubles / 2;
   PB_CUDA(cudaMalloc((void **)&ctx->d_partials, cap * sizeof(double)));
   ctx->partials_cap = cap;
   return 0;
}
int pb_ensure_scratch(pb200_ctx *ctx, size_t bytes) {
   if (bytes <= ctx->scratch_cap) return 0;
   PB_CUDA(cudaStreamSynchronize(ctx->stream));
   if (ctx->d_scratch) PB_CUDA(cudaFree(ctx->d_scratch));
   PB_CUDA(cudaMalloc(&ctx->d_scratch, bytes));
   ctx->scratch_cap = bytes;
   return 0;
}
int pb_ensure_small(pb200_ctx *ctx, size_t doubles) {
   if (doubles <= ctx->small_cap) return 0;
   PB_CUDA(cudaStreamSynchronize(ctx->stream));
   size_t cap = doubles * 2;
   cudaFreeHost(ctx->h_pinned);
   cudaFree(ctx->d_small);
   cudaFree(ctx->d_panel);
   PB_CUDA(cudaHostAlloc((void **)&ctx->h_pinned, cap * sizeof(double), cudaHostAllocMapped));
   PB_CUDA(cudaHostGetDevicePointer((void **)&ctx->d_hpinned, ctx->h_pinned, 0));
   PB_CUDA(cudaMalloc((void **)&ctx->d_small, cap * sizeof(double)));
   PB_CUDA(cudaMalloc((void **)&ctx->d_panel, cap * sizeof(double)));
   ctx->small_cap = cap;
   return 0;
}

// mapped pinned buffer of (value, tag) pairs for polled panels
int pb_ensure_tagged(pb200_ctx *ctx, size_t elems) {
   if (elems <= ctx->tagged_cap) return 0;
   PB_CUDA(cudaStreamSynchronize(ctx->stream));
   size_t cap = elems * 2 > 8192 ? elems * 2 : 8192;
   if (ctx->h_tagged) cudaFreeHost(ctx->h_tagged);
   PB_CUDA(cudaHostAlloc((void **)&ctx->h_tagged, cap * 2 * sizeof(double), cudaHostAllocMapped));
   memset(ctx->h_tagged, 0, cap * 2 * sizeof(double));
   PB_CUDA(cudaHostGetDevicePointer((void **)&ctx->d_htagged, ctx->h_tagged, 0));
   ctx->tagged_cap = cap;
   return 0;
}

// ---------------------------------------------------------------------- tensor maps ----
typedef CUresult (*pb_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
      const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int pb_tensor_map_2d(CUtensorMap *tm, const double *base, int64_t rows, int cols, int64_t ld,
      int box_rows, int box_cols) {
   static pb_encode_tiled_fn encode = NULL;
   static int tried = 0;
   if (!tried) {
      tried = 1;
      void *fn = NULL;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
         encode = (pb_encode_tiled_fn)fn;
      else
         cudaGetLastError();
   }
   if (!encode || rows <= 0 || cols <= 0 || box_cols > 256 || box_rows > 256) return 1;
   if ((((uintptr_t)base) & 15) != 0 || (ld % 2) != 0 || rows > 0x7fffffffLL) return 1;
   // A descriptor is a pure function of its six arguments, and a solve asks for the same few dozen over and over
   // (the basis grows by a block per outer iteration and every restart starts the cycle again): a small
   // direct-mapped cache takes the driver call out of the launch path of every sweep.
   struct MapKey {
      const double *base;
      int64_t rows, ld;
      int cols, box_rows, box_cols, valid;
   };
   static thread_local MapKey keys[64];
   static thread_local CUtensorMap vals[64];
   static const int use_cache = getenv("PB200_TMAP_CACHE") ? atoi(getenv("PB200_TMAP_CACHE")) : 1;
   const unsigned h = (unsigned)((((uintptr_t)base >> 4) * 0x9E3779B97F4A7C15ull + (uint64_t)cols * 0x85EBCA6Bu + (uint64_t)box_rows * 31u +
                                  (uint64_t)box_cols * 131u + (uint64_t)rows * 0xC2B2AE35u) >> 58) & 63;
   if (use_cache) {
      const MapKey &k = keys[h];
      if (k.valid && k.base == base && k.rows == rows && k.ld == ld && k.cols == cols && k.box_rows == box_rows &&
            k.box_cols == box_cols) {
         *tm = vals[h];
         return 0;
      }
   }
   const cuuint64_t gdim[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
   const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(double)};
   const cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)box_cols};
   const cuuint32_t estr[2] = {1, 1};
   CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, gdim, gstride, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
   if (r == CUDA_SUCCESS && use_cache) {
      keys[h] = MapKey{base, rows, ld, cols, box_rows, box_cols, 1};
      vals[h] = *tm;
   }
   return r == CUDA_SUCCESS ? 0 : 1;
}

// ------------------------------------------------------------------- event profiling ----
int pb_prof_flush(pb200_ctx *ctx) {
   if (!ctx->prof_ev || ctx->prof_pending == 0) return 0;
   PB_CUDA(cudaStreamSynchronize(ctx->stream));
   for (int i = 0; i < ctx->prof_pending; i++) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]) == cudaSuccess)
         ctx->prof_ms[ctx->prof_kind[i]] += ms;
   }
   ctx->prof_pending = 0;
   return 0;
}
int pb_prof_begin(pb200_ctx *ctx, int kind) {
   if (!ctx->prof_on) return -1;
   if (!ctx->prof_ev) {
      ctx->prof_ev = (cudaEvent_t *)calloc(2 * PB_PROF_RING, sizeof(cudaEvent_t));
      ctx->prof_kind = (int *)calloc(PB_PROF_RING, sizeof(int));
      for (int i = 0; i < 2 * PB_PROF_RING; i++) cudaEventCreate(&ctx->prof_ev[i]);
   }
   if (ctx->prof_pending >= PB_PROF_RING) pb_prof_flush(ctx);
   int slot = ctx->prof_pending++;
   ctx->prof_kind[slot] = kind;
   cudaEventRecord(ctx->prof_ev[2 * slot], ctx->stream);
   return slot;
}
void pb_prof_end(pb200_ctx *ctx, int slot, double bytes) {
   if (slot < 0) return;
   cudaEventRecord(ctx->prof_ev[2 * slot + 1], ctx->stream);
   int kind = ctx->prof_kind[slot];
   ctx->prof_bytes[kind] += bytes;
   ctx->prof_cnt[kind] += 1;
}
extern "C" int pb200_ctx_set_profiling(pb200_ctx *ctx, int on) {
   pb_prof_flush(ctx);
   ctx->prof_on = on;
   if (on) {
      for (int k = 0; k < 8; k++) ctx->prof_ms[k] = ctx->prof_bytes[k] = 0.0, ctx->prof_cnt[k] = 0;
   }
   return 0;
}
extern "C" int pb200_ctx_get_profile(pb200_ctx *ctx, int kind, int64_t *count, double *ms, double *bytes) {
   if (kind < 0 || kind >= 8) return PB200_ERR_ARG;
   pb_prof_flush(ctx);
   *count = ctx->prof_cnt[kind], *ms = ctx->prof_ms[kind], *bytes = ctx->prof_bytes[kind];
   return 0;
}

// out[e] = sum over the per-CTA partial panels, in a FIXED order => bitwise reproducible.
// One CTA of 8 warps per 32 consecutive elements: lane <-> element (coalesced rows of the
// partials array), warp w sums parts w, w+8, w+16, ... (4 independent chains), the 8 warp sums
// are combined in warp order through shared memory.
__global__ void __launch_bounds__(256) pb_reduce_partials_kernel(
      const double *__restrict__ partials, int nparts, int cnt, double *__restrict__ out,
      long long tag) {
   __shared__ double red[8][33];
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   const int e = blockIdx.x * 32 + lane;
   double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
   if (e < cnt) {
      int p = warp;
      for (; p + 24 < nparts; p += 32) {
         s0 += partials[(size_t)(p + 0) * cnt + e];
         s1 += partials[(size_t)(p + 8) * cnt + e];
         s2 += partials[(size_t)(p + 16) * cnt + e];
         s3 += partials[(size_t)(p + 24) * cnt + e];
      }
      for (; p < nparts; p += 8) s0 += partials[(size_t)p * cnt + e];
   }
   red[warp][lane] = (s0 + s1) + (s2 + s3);
   __syncthreads();
   if (warp == 0 && e < cnt) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < 8; w++) s += red[w][lane];
      if (tag) {
         // `out` is mapped host memory: every element travels with the sequence number of this
         // panel in ONE 16-byte store, so the host can tell a fresh value from a stale one without
         // any fence, ticket or flag on the device side (the stores of all CTAs fly in parallel)
         reinterpret_cast<double2 *>(out)[e] = make_double2(s, __longlong_as_double(tag));
      } else {
         out[e] = s;
      }
   }
}

// wait until the (value, sequence number) pairs of panel `seq` have landed in mapped host memory
// PB200_HOST_PROFILE: wall time the host spends waiting for panels (front.c prints it next to the solve time)
static double g_wait_s = 0.0;
static long g_wait_n = 0;
extern "C" double pb200_debug_wait_seconds(long *calls) {
   if (calls) *calls = g_wait_n;
   return g_wait_s;
}
static int pb_poll_tagged_(pb200_ctx *ctx, int cnt, long long seq);
static int pb_poll_tagged(pb200_ctx *ctx, int cnt, long long seq) {
   static const int prof = getenv("PB200_HOST_PROFILE") != NULL;
   if (!prof) return pb_poll_tagged_(ctx, cnt, seq);
   struct timespec a, b;
   clock_gettime(CLOCK_MONOTONIC, &a);
   const int rc = pb_poll_tagged_(ctx, cnt, seq);
   clock_gettime(CLOCK_MONOTONIC, &b);
   g_wait_s += (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
   g_wait_n++;
   return rc;
}
static int pb_poll_tagged_(pb200_ctx *ctx, int cnt, long long seq) {
   volatile const long long *tg = (volatile const long long *)ctx->h_tagged;
   volatile const double *tv = (volatile const double *)ctx->h_tagged;
   unsigned long spins = 0;
   time_t t0 = 0;
   for (int e = 0; e < cnt; e++) {
      while (tg[2 * e + 1] != seq) {
         if ((++spins & 0xfffff) == 0) {
            // watchdog: a panel never takes seconds; give up instead of spinning for ever
            const time_t now = time(NULL);
            if (!t0) t0 = now;
            if (now - t0 > 120) {
               fprintf(stderr, "primme_b200: timed out waiting for a panel (kernel hung?)\n");
               return PB200_ERR_CUDA;
            }
            cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) {
               fprintf(stderr, "primme_b200: CUDA error %s while waiting for a panel\n", cudaGetErrorString(q));
               return PB200_ERR_CUDA;
            }
         }
      }
      // the value must not be read before the tag was seen (weakly ordered hosts, e.g. aarch64)
      __atomic_thread_fence(__ATOMIC_ACQUIRE);
      ctx->h_pinned[e] = tv[2 * e];
   }
   return 0;
}

int pb_fin_prepare(pb200_ctx *ctx, int grid, int ppc, int cnt, PbFin *f) {
   memset(f, 0, sizeof(*f));
   const int ngroups = (grid + PB_FIN_GROUP - 1) / PB_FIN_GROUP;
   if (!ctx->fused_finish || cnt <= 0 || ngroups > PB_FIN_MAXGROUPS) return 1;
   PB_CHK(pb_ensure_small(ctx, (size_t)cnt));
   PB_CHK(pb_ensure_partials(ctx, (size_t)grid * ppc * cnt + 16));
   if ((size_t)ngroups * cnt > ctx->gpart_cap) {
      PB_CUDA(cudaStreamSynchronize(ctx->stream));
      if (ctx->d_gpart) PB_CUDA(cudaFree(ctx->d_gpart));
      const size_t cap = (size_t)PB_FIN_MAXGROUPS * cnt;
      PB_CUDA(cudaMalloc((void **)&ctx->d_gpart, cap * sizeof(double)));
      ctx->gpart_cap = cap;
   }
   f->partials = ctx->d_partials, f->gpart = ctx->d_gpart, f->counters = ctx->d_counters;
   f->cnt = cnt, f->ppc = ppc;
   f->nranks = 1;
   if (ctx->nranks > 1 && ctx->peer_on && !ctx->no_poll && cnt <= PB_XCHG_CAP) {
      // the kernel all-reduces over peer memory and delivers like a single-GPU panel (the choice
      // depends on cnt only: every rank takes the same path)
      PB_CHK(pb_ensure_tagged(ctx, (size_t)cnt));
      f->out = ctx->d_htagged, f->tag = ++ctx->seq;
      f->nranks = ctx->nranks, f->rank = ctx->rank, f->slot = (int)(f->tag % PB_XCHG_SLOTS);
      for (int p = 0; p < ctx->nranks; p++) f->peer[p] = ctx->xchg_peer[p];
   } else if (ctx->nranks > 1 || ctx->no_poll) {
      f->out = ctx->d_panel, f->tag = 0;
   } else {
      PB_CHK(pb_ensure_tagged(ctx, (size_t)cnt));
      f->out = ctx->d_htagged, f->tag = ++ctx->seq;
   }
   return 0;
}

int pb_collect_panel(pb200_ctx *ctx, const PbFin *f) {
   if (f->cnt <= 0) return 0;
   if (f->tag == 0) {
      if (ctx->nranks > 1) PB_CHK(pb_nccl_allreduce_dev(ctx, ctx->d_panel, f->cnt));
      PB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_panel, (size_t)f->cnt * sizeof(double),
            cudaMemcpyDeviceToHost, ctx->stream));
      PB_CUDA(cudaStreamSynchronize(ctx->stream));
      return 0;
   }
   return pb_poll_tagged(ctx, f->cnt, f->tag);
}

int pb_finish_panel(pb200_ctx *ctx, int nparts, int cnt) {
   if (cnt <= 0) return 0;
   PB_CHK(pb_ensure_small(ctx, (size_t)cnt));
   int ps = pb_prof_begin(ctx, PB_K_REDUCE);
   if (ctx->nranks > 1 || ctx->no_poll) {
      pb_reduce_partials_kernel<<<(cnt + 31) / 32, 256, 0, ctx->stream>>>(
            ctx->d_partials, nparts, cnt, ctx->d_panel, 0);
      pb_prof_end(ctx, ps, 8.0 * nparts * cnt);
      ctx->launches++;
      PB_CUDA(cudaGetLastError());
      if (ctx->nranks > 1) PB_CHK(pb_nccl_allreduce_dev(ctx, ctx->d_panel, cnt));
      PB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_panel, (size_t)cnt * sizeof(double),
            cudaMemcpyDeviceToHost, ctx->stream));
      PB_CUDA(cudaStreamSynchronize(ctx->stream));
      return 0;
   }
   // single rank: the reduction writes (value, sequence number) pairs straight into mapped pinned
   // memory; the host spins on the sequence numbers (a few microseconds less than memcpy + stream
   // synchronisation, 10+ times per outer iteration)
   PB_CHK(pb_ensure_tagged(ctx, (size_t)cnt));
   const long long seq = ++ctx->seq;
   pb_reduce_partials_kernel<<<(cnt + 31) / 32, 256, 0, ctx->stream>>>(
         ctx->d_partials, nparts, cnt, ctx->d_htagged, seq);
   pb_prof_end(ctx, ps, 8.0 * nparts * cnt);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return pb_poll_tagged(ctx, cnt, seq);
}

// ---------------------------------------------------------------------------- memory ------
extern "C" int pb200_malloc(pb200_ctx *ctx, size_t bytes, void **dptr) {
   (void)ctx;
   cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 16);
   if (e != cudaSuccess) {
      cudaGetLastError();
      *dptr = NULL;
      return PB200_ERR_ALLOC;
   }
   return 0;
}
extern "C" int pb200_free(pb200_ctx *ctx, void *dptr) {
   if (!dptr) return 0;
   if (ctx) PB_CUDA(cudaStreamSynchronize(ctx->stream));
   PB_CUDA(cudaFree(dptr));
   return 0;
}
// Basis workspace that survives between solves on the same context: V and W (2 x ld x maxBasisSize,
// hundreds of MB) are not re-allocated and freed by every dprimme call (cudaFree synchronises the
// device; on a busy host both calls showed up as 0.1-1 s outliers of the end-to-end time).
extern "C" int pb200_ctx_workspace(pb200_ctx *ctx, int slot, size_t bytes, void **dptr) {
   if (slot < 0 || slot >= PB_WS_SLOTS) return PB200_ERR_ARG;
   if (bytes > ctx->ws_bytes[slot]) {
      if (ctx->ws_ptr[slot]) {
         PB_CUDA(cudaStreamSynchronize(ctx->stream));
         PB_CUDA(cudaFree(ctx->ws_ptr[slot]));
         ctx->ws_ptr[slot] = NULL, ctx->ws_bytes[slot] = 0;
      }
      cudaError_t e = cudaMalloc(&ctx->ws_ptr[slot], bytes ? bytes : 16);
      if (e != cudaSuccess) {
         cudaGetLastError();
         *dptr = NULL;
         return PB200_ERR_ALLOC;
      }
      ctx->ws_bytes[slot] = bytes;
   }
   *dptr = ctx->ws_ptr[slot];
   return 0;
}

extern "C" int pb200_memset0(pb200_ctx *ctx, void *dptr, size_t bytes) {
   PB_CUDA(cudaMemsetAsync(dptr, 0, bytes, ctx->stream));
   return 0;
}
static int copy2d(pb200_ctx *ctx, const void *s, int64_t lds, void *d, int64_t ldd, int64_t rows,
      int cols, int es, cudaMemcpyKind kind) {
   if (rows <= 0 || cols <= 0) return 0;
   PB_CUDA(cudaMemcpy2DAsync(d, (size_t)ldd * es, s, (size_t)lds * es, (size_t)rows * es,
         (size_t)cols, kind, ctx->stream));
   // host-visible copies complete before returning (pageable host memory may be reused)
   if (kind != cudaMemcpyDeviceToDevice) PB_CUDA(cudaStreamSynchronize(ctx->stream));
   return 0;
}
extern "C" int pb200_copy_h2d(pb200_ctx *ctx, const void *s, int64_t lds, void *d, int64_t ldd,
      int64_t rows, int cols, int es) {
   return copy2d(ctx, s, lds, d, ldd, rows, cols, es, cudaMemcpyHostToDevice);
}
extern "C" int pb200_copy_d2h(pb200_ctx *ctx, const void *s, int64_t lds, void *d, int64_t ldd,
      int64_t rows, int cols, int es) {
   return copy2d(ctx, s, lds, d, ldd, rows, cols, es, cudaMemcpyDeviceToHost);
}
extern "C" int pb200_copy_d2d(pb200_ctx *ctx, const void *s, int64_t lds, void *d, int64_t ldd,
      int64_t rows, int cols, int es) {
   return copy2d(ctx, s, lds, d, ldd, rows, cols, es, cudaMemcpyDeviceToDevice);
}
extern "C" int pb200_is_device_pointer(const void *p) {
   cudaPointerAttributes a;
   if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
      cudaGetLastError();
      return 0;
   }
   return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? 1 : 0;
}

// ------------------------------------------------------------ peer-memory panel exchange ----
extern "C" int pb200_ctx_peer_export(pb200_ctx *ctx, void *handle64) {
   if (!ctx->xchg_local) {
      const size_t bytes = sizeof(double2) * PB_XCHG_SLOTS * PB_MAX_PEERS * PB_XCHG_CAP;
      PB_CUDA(cudaMalloc((void **)&ctx->xchg_local, bytes));
      PB_CUDA(cudaMemset(ctx->xchg_local, 0, bytes));
      PB_CUDA(cudaDeviceSynchronize());
   }
   cudaIpcMemHandle_t h;
   PB_CUDA(cudaIpcGetMemHandle(&h, ctx->xchg_local));
   memcpy(handle64, &h, sizeof(h));
   return 0;
}

extern "C" int pb200_ctx_peer_attach(pb200_ctx *ctx, int nranks, int rank, const void *handles) {
   if (nranks < 1 || nranks > PB_MAX_PEERS || rank < 0 || rank >= nranks || !ctx->xchg_local) return PB200_ERR_ARG;
   // The decision is COLLECTIVE: a rank that cannot open a peer's buffer (or has the exchange switched
   // off) must not leave the others spinning for its pairs, so the local outcome is summed over the
   // communicator (pb200_ctx_comm_init / pb200_ctx_set_comm must have run) and the exchange is
   // enabled only if every rank succeeded.  Returns 0 (exchange on, everywhere) or 1 (panels stay on
   // NCCL, everywhere).
   int failed = getenv("PB200_NO_PEER_EXCHANGE") ? 1 : 0;
   for (int p = 0; p < nranks && !failed; p++) {
      if (p == rank) {
         ctx->xchg_peer[p] = ctx->xchg_local;
         continue;
      }
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char *)handles + (size_t)p * sizeof(h), sizeof(h));
      void *ptr = NULL;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
         fprintf(stderr, "primme_b200: cudaIpcOpenMemHandle(rank %d) failed: %s -- panels stay on NCCL\n", p,
               cudaGetErrorString(e));
         cudaGetLastError();
         failed = 1;
         break;
      }
      ctx->xchg_peer[p] = (double2 *)ptr;
   }
   if (nranks > 1) {
      if (!ctx->comm || ctx->nranks != nranks || ctx->rank != rank) {
         // no communicator to agree over: refuse rather than risk a one-sided decision
         failed = 1;
      } else {
         double f = (double)failed;
         PB_CHK(pb200_allreduce_host(ctx, &f, 1));
         failed = f > 0.0;
      }
   }
   if (failed) {
      for (int q = 0; q < nranks; q++)
         if (q != rank && ctx->xchg_peer[q]) cudaIpcCloseMemHandle(ctx->xchg_peer[q]), ctx->xchg_peer[q] = NULL;
      ctx->peer_on = 0;
      return 1;
   }
   ctx->nranks = nranks, ctx->rank = rank;
   ctx->peer_on = 1;
   return 0;
}

extern "C" int pb200_ctx_peer_active(pb200_ctx *ctx) { return ctx->peer_on; }

// a rank without local rows still takes part in the exchange: it contributes zeros
__global__ void pb_fin_zero_kernel(PbFin f) {
   for (int e = threadIdx.x; e < f.cnt; e += blockDim.x) pb_fin_store(f, e, 0.0);
}
int pb_fin_contribute_zeros(pb200_ctx *ctx, int cnt) {
   PbFin f;
   const int r = pb_fin_prepare(ctx, 1, 1, cnt, &f);
   if (r < 0) return r;
   if (r == 1 || f.nranks <= 1) return 1;  // not on the peer path
   pb_fin_zero_kernel<<<1, 256, 0, ctx->stream>>>(f);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return pb_collect_panel(ctx, &f);
}

// ------------------------------------------------------------------------------ NCCL ------
// NCCL is resolved at run time (dlopen) so that single-GPU users need no libnccl at link time.
typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_bcast_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_void_fn)(void);
typedef int (*nccl_uid_fn)(void *);
typedef struct { char internal[128]; } pb_nccl_uid;
typedef int (*nccl_init_fn)(void **, int, pb_nccl_uid, int);
typedef int (*nccl_destroy_fn)(void *);
static nccl_allreduce_fn p_ncclAllReduce = NULL;
static nccl_bcast_fn p_ncclBroadcast = NULL;
static nccl_void_fn p_ncclGroupStart = NULL, p_ncclGroupEnd = NULL;
static nccl_uid_fn p_ncclGetUniqueId = NULL;
static nccl_init_fn p_ncclCommInitRank = NULL;
static nccl_destroy_fn p_ncclCommDestroy = NULL;
typedef int (*nccl_allgather_fn)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*nccl_send_fn)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_recv_fn)(void *, size_t, int, int, void *, cudaStream_t);
static nccl_allgather_fn p_ncclAllGather = NULL;
static nccl_send_fn p_ncclSend = NULL;
static nccl_recv_fn p_ncclRecv = NULL;
enum { PB_NCCL_FLOAT64 = 8, PB_NCCL_SUM = 0 };

static int pb_nccl_load(void) {
   if (p_ncclAllReduce) return 0;
   void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
   if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
   if (!h) {
      fprintf(stderr, "primme_b200: cannot dlopen libnccl: %s\n", dlerror());
      return PB200_ERR_CUDA;
   }
   p_ncclAllReduce = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
   p_ncclBroadcast = (nccl_bcast_fn)dlsym(h, "ncclBroadcast");
   p_ncclGroupStart = (nccl_void_fn)dlsym(h, "ncclGroupStart");
   p_ncclGroupEnd = (nccl_void_fn)dlsym(h, "ncclGroupEnd");
   p_ncclGetUniqueId = (nccl_uid_fn)dlsym(h, "ncclGetUniqueId");
   p_ncclCommInitRank = (nccl_init_fn)dlsym(h, "ncclCommInitRank");
   p_ncclCommDestroy = (nccl_destroy_fn)dlsym(h, "ncclCommDestroy");
   p_ncclAllGather = (nccl_allgather_fn)dlsym(h, "ncclAllGather");
   p_ncclSend = (nccl_send_fn)dlsym(h, "ncclSend");
   p_ncclRecv = (nccl_recv_fn)dlsym(h, "ncclRecv");
   return (p_ncclAllReduce && p_ncclBroadcast && p_ncclGroupStart && p_ncclGroupEnd &&
                p_ncclGetUniqueId && p_ncclCommInitRank && p_ncclCommDestroy && p_ncclAllGather && p_ncclSend &&
                p_ncclRecv)
                ? 0
                : PB200_ERR_CUDA;
}

// 128-byte NCCL unique id, to be created on rank 0 and shared by the launcher
extern "C" int pb200_comm_unique_id(void *id128) {
   PB_CHK(pb_nccl_load());
   return p_ncclGetUniqueId(id128) == 0 ? 0 : PB200_ERR_CUDA;
}
// Create this rank's communicator and attach it to the context
extern "C" int pb200_ctx_comm_init(pb200_ctx *ctx, int nranks, int rank, const void *id128) {
   PB_CHK(pb_nccl_load());
   pb_nccl_uid uid;
   memcpy(&uid, id128, sizeof(uid));
   void *comm = NULL;
   PB_CUDA(cudaSetDevice(ctx->device));
   if (p_ncclCommInitRank(&comm, nranks, uid, rank) != 0) return PB200_ERR_CUDA;
   ctx->comm = comm, ctx->nranks = nranks, ctx->rank = rank;
   ctx->owns_comm = 1;
   return 0;
}
extern "C" int pb200_ctx_comm_free(pb200_ctx *ctx) {
   if (ctx->comm && ctx->owns_comm) {
      cudaStreamSynchronize(ctx->stream);
      p_ncclCommDestroy(ctx->comm);
   }
   ctx->comm = NULL, ctx->nranks = 1, ctx->rank = 0, ctx->owns_comm = 0;
   return 0;
}
// Y(displs[r] .. +counts[r]) on every rank <- X of rank r, for each of ncols columns:
// all-gather with unequal counts as grouped broadcasts (halo exchange of the SpMV block).
int pb_nccl_allgatherv_cols(pb200_ctx *ctx, const double *X, int64_t ldx, double *Y, int64_t ldy,
      const int64_t *counts, const int64_t *displs, int ncols) {
   if (ctx->nranks <= 1) return 0;
   if (p_ncclGroupStart() != 0) return PB200_ERR_CUDA;
   for (int c = 0; c < ncols; c++)
      for (int r = 0; r < ctx->nranks; r++) {
         const double *src = (r == ctx->rank) ? X + (size_t)c * ldx : Y + (size_t)c * ldy + displs[r];
         if (p_ncclBroadcast(src, Y + (size_t)c * ldy + displs[r], (size_t)counts[r], PB_NCCL_FLOAT64, r,
                   ctx->comm, ctx->stream) != 0)
            return PB200_ERR_CUDA;
      }
   if (p_ncclGroupEnd() != 0) return PB200_ERR_CUDA;
   return 0;
}

// byte-wise helpers for the setup of the row-sharded operator (dist.cu); device buffers, in-stream
int pb_nccl_allgather(pb200_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank) {
   if (ctx->nranks <= 1 || !ctx->comm) return PB200_ERR_ARG;
   return p_ncclAllGather(send, recv, bytes_per_rank, 0 /* ncclInt8 */, ctx->comm, ctx->stream) == 0 ? 0 : PB200_ERR_CUDA;
}
int pb_nccl_group(pb200_ctx *ctx, int start) {
   (void)ctx;
   return (start ? p_ncclGroupStart() : p_ncclGroupEnd()) == 0 ? 0 : PB200_ERR_CUDA;
}
int pb_nccl_send(pb200_ctx *ctx, const void *buf, size_t bytes, int peer) {
   return p_ncclSend(buf, bytes, 0, peer, ctx->comm, ctx->stream) == 0 ? 0 : PB200_ERR_CUDA;
}
int pb_nccl_recv(pb200_ctx *ctx, void *buf, size_t bytes, int peer) {
   return p_ncclRecv(buf, bytes, 0, peer, ctx->comm, ctx->stream) == 0 ? 0 : PB200_ERR_CUDA;
}

extern "C" int pb200_ctx_set_comm(pb200_ctx *ctx, void *nccl_comm, int nranks, int rank) {
   if (nranks <= 1) {
      ctx->comm = NULL, ctx->nranks = 1, ctx->rank = 0;
      return 0;
   }
   PB_CHK(pb_nccl_load());
   ctx->comm = nccl_comm, ctx->nranks = nranks, ctx->rank = rank;
   return 0;
}

int pb_nccl_allreduce_dev(pb200_ctx *ctx, double *dbuf, int count) {
   if (ctx->nranks <= 1) return 0;
   int r = p_ncclAllReduce(dbuf, dbuf, (size_t)count, PB_NCCL_FLOAT64, PB_NCCL_SUM, ctx->comm,
         ctx->stream);
   if (r != 0) {
      fprintf(stderr, "primme_b200: ncclAllReduce failed (%d)\n", r);
      return PB200_ERR_CUDA;
   }
   return 0;
}

extern "C" int pb200_allreduce_host(pb200_ctx *ctx, double *buf, int count) {
   if (ctx->nranks <= 1 || count <= 0) return 0;
   PB_CHK(pb_ensure_small(ctx, (size_t)count));
   memcpy(ctx->h_pinned, buf, sizeof(double) * count);
   PB_CUDA(cudaMemcpyAsync(ctx->d_panel, ctx->h_pinned, sizeof(double) * count,
         cudaMemcpyHostToDevice, ctx->stream));
   PB_CHK(pb_nccl_allreduce_dev(ctx, ctx->d_panel, count));
   PB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_panel, sizeof(double) * count,
         cudaMemcpyDeviceToHost, ctx->stream));
   PB_CUDA(cudaStreamSynchronize(ctx->stream));
   memcpy(buf, ctx->h_pinned, sizeof(double) * count);
   return 0;
}

extern "C" int pb200_bcast_host(pb200_ctx *ctx, double *buf, int count, int root) {
   if (ctx->nranks <= 1 || count <= 0) return 0;
   PB_CHK(pb_ensure_small(ctx, (size_t)count));
   memcpy(ctx->h_pinned, buf, sizeof(double) * count);
   PB_CUDA(cudaMemcpyAsync(ctx->d_panel, ctx->h_pinned, sizeof(double) * count,
         cudaMemcpyHostToDevice, ctx->stream));
   int r = p_ncclBroadcast(ctx->d_panel, ctx->d_panel, (size_t)count, PB_NCCL_FLOAT64, root,
         ctx->comm, ctx->stream);
   if (r != 0) return PB200_ERR_CUDA;
   PB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_panel, sizeof(double) * count,
         cudaMemcpyDeviceToHost, ctx->stream));
   PB_CUDA(cudaStreamSynchronize(ctx->stream));
   memcpy(buf, ctx->h_pinned, sizeof(double) * count);
   return 0;
}

// dist.cu -- row-sharded operator for one-process-per-GPU runs.
//
// PRIMME's SPMD model (reference include/primme_eigs.h:187-198, examples/ex_eigs_mpi.c:106-112):
// rank r owns a contiguous block of rows of A and of every multivector.  The only data-path
// exchange of the Davidson loop besides the small panel reductions is the SpMV halo: every rank
// needs the entries of the block x that its off-diagonal columns reference.
//
// Compacted halo, pushed over NVLink peer memory:
//   setup   every rank lists the DISTINCT remote columns its rows reference, per owner; the lists
//           are exchanged once (NCCL send/recv) so that each owner knows which of its rows every
//           peer needs; the column indices of the local CSR are remapped to
//           [0, nLocal) = own rows, [nLocal, nLocal + nHalo) = halo rows in (owner, column) order.
//   per block of b columns (sequence number seq):
//     1. ONE kernel (dist_push_kernel) packs the local block into the row-major gather buffer
//        G[seq & 1] of the v3 SpMM (spmm.cu) AND writes, for every peer, the rows that peer needs
//        straight into the peer's G[seq & 1] over NVLink (coalesced 8*bp-byte rows); its last CTA
//        releases flags[me] = seq in every destination.
//     2. the SpMM kernel's consumers wait (ld.acquire.sys) until every source has flagged seq, gather
//        from G with 32-byte loads, and the last CTA acknowledges seq to the sources, which will not
//        overwrite G[seq & 1] before (double buffering: a rank may run one block ahead of a peer).
//   No NCCL call, no host synchronisation and no copy of the whole vector on the data path.  Without
//   peer access the same compacted rows travel through grouped ncclSend/ncclRecv.
#include "pb200_internal.cuh"
#include "../../include/primme.h"
#include "../../include/primme_svds.h"
#include <stdlib.h>
#include <string.h>
#include <vector>

#define PB_DIST_HDR 4096  // flags[8] at byte 0, done[8] at byte 64 of the exchange region

struct pb200_dist_csr {
   pb200_csr *A;       // nLocal rows; column indices remapped by pb200_dist_csr_create
   int nranks, rank;
   int64_t nglobal, nloc, nhalo;
   int64_t *counts, *displs;  // host
   int64_t recv_cnt[PB_MAX_PEERS], recv_off[PB_MAX_PEERS];  // rows received from owner o / their offset in the halo
   int64_t send_cnt[PB_MAX_PEERS], send_off[PB_MAX_PEERS];  // rows sent to d / offset of my segment in d's halo
   int64_t peer_nloc[PB_MAX_PEERS], peer_nhalo[PB_MAX_PEERS];
   int32_t *d_send_rows[PB_MAX_PEERS];                      // local row ids needed by d (device)
   char *region;                                            // exchange region of this rank (header + 2 G buffers)
   char *peer_region[PB_MAX_PEERS];
   size_t rowbytes_max;                                     // 64 (real) or 128 (complex): 8 columns per row of G
   int peer_on;
   unsigned long long seq;
   unsigned int *d_counters;                                // [0] push kernel, [1] SpMM kernel
   double *d_stage;                                         // send staging of the NCCL fallback
   int64_t stat_blocks, stat_rows_sent;
   double dbg_push_ms, dbg_spmm_ms;                         // PB200_DIST_TIMING: event times of the two kernels
};

namespace {

size_t gbuf_bytes(int64_t nloc, int64_t nhalo, size_t rowbytes) {
   size_t b = (size_t)(nloc + nhalo > 0 ? nloc + nhalo : 1) * rowbytes;
   return (b + 255) / 256 * 256;
}

struct PbPushArgs {
   const void *X;
   int64_t ldx;
   int b, bp;
   int64_t nloc;
   void *Glocal;
   int ndst;
   int dst_rank[PB_MAX_PEERS];
   const int32_t *rows[PB_MAX_PEERS];
   int64_t start[PB_MAX_PEERS + 1];  // prefix sums of the rows per destination
   void *dstG[PB_MAX_PEERS];         // first row of my segment in the destination's G
   unsigned long long *flag[PB_MAX_PEERS];
   const unsigned long long *done;   // local done[] (written by the peers)
   unsigned long long seq;
   unsigned int *counter;
};

template <typename VT>
__global__ void __launch_bounds__(256) dist_push_kernel(const PbPushArgs a) {
   // a destination's buffer seq & 1 is free once it has acknowledged block seq - 2
   if (a.seq > 2 && a.done && threadIdx.x < a.ndst) {
      unsigned long long f, spins = 0;
      const unsigned long long *p = a.done + a.dst_rank[threadIdx.x];
      do {
         asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(p) : "memory");
      } while (f + 2 < a.seq && ++spins < (1ull << 31));
   }
   __syncthreads();
   const VT *X = reinterpret_cast<const VT *>(a.X);
   const int sh = a.bp == 1 ? 0 : a.bp == 2 ? 1 : a.bp == 4 ? 2 : 3;
   const int64_t nitems = a.nloc + a.start[a.ndst];
   const int64_t total = nitems << sh;
   for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
      const int64_t q = e >> sh;
      const int c = (int)(e & (a.bp - 1));
      int64_t r;
      VT *dst;
      if (q < a.nloc) {
         r = q;
         dst = reinterpret_cast<VT *>(a.Glocal) + e;
      } else {
         const int64_t h = q - a.nloc;
         int d = 0;
         while (d + 1 < a.ndst && h >= a.start[d + 1]) d++;
         const int64_t i = h - a.start[d];
         r = a.rows[d][i];
         dst = reinterpret_cast<VT *>(a.dstG[d]) + ((i << sh) + c);
      }
      VT v;
      if (c < a.b) v = X[r + (size_t)c * a.ldx];
      else memset(&v, 0, sizeof(v));
      *dst = v;
   }
   if (a.counter) {
      __syncthreads();
      if (threadIdx.x == 0) {
         __threadfence_system();
         const unsigned int t = atomicAdd(a.counter, 1u);
         if (t == gridDim.x - 1) {
            *a.counter = 0u;
            __threadfence_system();
            for (int d = 0; d < a.ndst; d++)
               if (a.flag[d]) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.flag[d]), "l"(a.seq) : "memory");
         }
      }
   }
}

}  // namespace

extern "C" int pb200_dist_csr_create(pb200_ctx *ctx, pb200_csr *A_local, const int64_t *counts_host,
      int nranks, pb200_dist_csr **out) {
   if (nranks != ctx->nranks || nranks > PB_MAX_PEERS) return PB200_ERR_ARG;
   pb200_dist_csr *D = (pb200_dist_csr *)calloc(1, sizeof(*D));
   if (!D) return PB200_ERR_ALLOC;
   D->A = A_local, D->nranks = nranks, D->rank = nranks > 1 ? ctx->rank : 0;
   D->counts = (int64_t *)malloc(sizeof(int64_t) * nranks);
   D->displs = (int64_t *)malloc(sizeof(int64_t) * nranks);
   int64_t off = 0;
   for (int r = 0; r < nranks; r++) D->counts[r] = counts_host[r], D->displs[r] = off, off += counts_host[r];
   D->nglobal = off;
   D->nloc = D->counts[D->rank];
   D->rowbytes_max = A_local->is_complex ? 128 : 64;
   *out = D;
   if (nranks <= 1) return 0;
   // counts partitions the COLUMN space (the rows of the block x this operator is applied to); the
   // local matrix may have any number of rows (square operators: nrows == counts[rank]; the two
   // operators of the row-partitioned SVD, below, are rectangular)
   const int me = D->rank;

   // ---- distinct remote columns per owner, remapped column indices ----
   std::vector<int32_t> ci((size_t)(A_local->nnz ? A_local->nnz : 1));
   PB_CUDA(cudaStreamSynchronize(ctx->stream));
   if (A_local->nnz)
      PB_CUDA(cudaMemcpy(ci.data(), A_local->d_colind, sizeof(int32_t) * A_local->nnz, cudaMemcpyDeviceToHost));
   std::vector<int32_t> hmap((size_t)(D->nglobal ? D->nglobal : 1), -1);
   const int64_t mylo = D->displs[me], myhi = mylo + D->nloc;
   for (int64_t k = 0; k < A_local->nnz; k++) {
      const int64_t c = ci[k];
      if (c < 0 || c >= D->nglobal) return PB200_ERR_ARG;
      if (c < mylo || c >= myhi) hmap[c] = 0;
   }
   std::vector<std::vector<int32_t>> need(nranks);
   int64_t nh = 0;
   for (int o = 0; o < nranks; o++) {
      D->recv_off[o] = nh;
      if (o == me) continue;
      for (int64_t c = D->displs[o]; c < D->displs[o] + D->counts[o]; c++)
         if (hmap[c] == 0) {
            hmap[c] = (int32_t)(D->nloc + nh);
            need[o].push_back((int32_t)(c - D->displs[o]));
            nh++;
         }
      D->recv_cnt[o] = (int64_t)need[o].size();
   }
   D->nhalo = nh;
   if (D->nloc + nh > 0x7fffffffLL) return PB200_ERR_ARG;
   for (int64_t k = 0; k < A_local->nnz; k++) {
      const int64_t c = ci[k];
      ci[k] = (c >= mylo && c < myhi) ? (int32_t)(c - mylo) : hmap[c];
   }
   if (A_local->nnz)
      PB_CUDA(cudaMemcpy(A_local->d_colind, ci.data(), sizeof(int32_t) * A_local->nnz, cudaMemcpyHostToDevice));
   A_local->ncols = D->nloc + D->nhalo;
   std::vector<int32_t>().swap(hmap);
   std::vector<int32_t>().swap(ci);

   // ---- who needs what from whom: count matrix M[r][o] = rows r receives from o ----
   std::vector<int64_t> M((size_t)nranks * nranks, 0);
   {
      int64_t *d_m = NULL;
      PB_CUDA(cudaMalloc((void **)&d_m, sizeof(int64_t) * nranks * (nranks + 1)));
      PB_CUDA(cudaMemcpy(d_m + (size_t)nranks * nranks, D->recv_cnt, sizeof(int64_t) * nranks, cudaMemcpyHostToDevice));
      PB_CHK(pb_nccl_allgather(ctx, d_m + (size_t)nranks * nranks, d_m, sizeof(int64_t) * nranks));
      PB_CUDA(cudaStreamSynchronize(ctx->stream));
      PB_CUDA(cudaMemcpy(M.data(), d_m, sizeof(int64_t) * nranks * nranks, cudaMemcpyDeviceToHost));
      cudaFree(d_m);
   }
   for (int d = 0; d < nranks; d++) {
      D->peer_nloc[d] = D->counts[d];
      int64_t tot = 0, before = 0;
      for (int o = 0; o < nranks; o++) {
         if (o == d) continue;
         if (o < me) before += M[(size_t)d * nranks + o];
         tot += M[(size_t)d * nranks + o];
      }
      D->peer_nhalo[d] = tot;
      D->send_cnt[d] = d == me ? 0 : M[(size_t)d * nranks + me];
      D->send_off[d] = before;
   }
   // ---- need lists travel to the owners ----
   {
      std::vector<int32_t *> d_need(nranks, (int32_t *)NULL);
      for (int o = 0; o < nranks; o++) {
         if (D->recv_cnt[o] > 0) {
            PB_CUDA(cudaMalloc((void **)&d_need[o], sizeof(int32_t) * D->recv_cnt[o]));
            PB_CUDA(cudaMemcpy(d_need[o], need[o].data(), sizeof(int32_t) * D->recv_cnt[o], cudaMemcpyHostToDevice));
         }
         if (D->send_cnt[o] > 0) PB_CUDA(cudaMalloc((void **)&D->d_send_rows[o], sizeof(int32_t) * D->send_cnt[o]));
      }
      PB_CHK(pb_nccl_group(ctx, 1));
      for (int o = 0; o < nranks; o++) {
         if (D->recv_cnt[o] > 0) PB_CHK(pb_nccl_send(ctx, d_need[o], sizeof(int32_t) * D->recv_cnt[o], o));
         if (D->send_cnt[o] > 0) PB_CHK(pb_nccl_recv(ctx, D->d_send_rows[o], sizeof(int32_t) * D->send_cnt[o], o));
      }
      PB_CHK(pb_nccl_group(ctx, 0));
      PB_CUDA(cudaStreamSynchronize(ctx->stream));
      for (int o = 0; o < nranks; o++) cudaFree(d_need[o]);
   }
   // ---- exchange region: header + two gather buffers; IPC handles all-gathered over NCCL ----
   const size_t gb = gbuf_bytes(D->nloc, D->nhalo, D->rowbytes_max);
   PB_CUDA(cudaMalloc((void **)&D->region, PB_DIST_HDR + 2 * gb));
   PB_CUDA(cudaMemset(D->region, 0, PB_DIST_HDR));
   PB_CUDA(cudaMalloc((void **)&D->d_counters, 2 * sizeof(unsigned int)));
   PB_CUDA(cudaMemset(D->d_counters, 0, 2 * sizeof(unsigned int)));
   PB_CUDA(cudaDeviceSynchronize());
   int failed = getenv("PB200_NO_PEER_HALO") ? 1 : 0;
   {
      cudaIpcMemHandle_t h;
      PB_CUDA(cudaIpcGetMemHandle(&h, D->region));
      char *d_h = NULL;
      PB_CUDA(cudaMalloc((void **)&d_h, sizeof(h) * (nranks + 1)));
      PB_CUDA(cudaMemcpy(d_h + sizeof(h) * nranks, &h, sizeof(h), cudaMemcpyHostToDevice));
      PB_CHK(pb_nccl_allgather(ctx, d_h + sizeof(h) * nranks, d_h, sizeof(h)));
      PB_CUDA(cudaStreamSynchronize(ctx->stream));
      std::vector<cudaIpcMemHandle_t> all(nranks);
      PB_CUDA(cudaMemcpy(all.data(), d_h, sizeof(h) * nranks, cudaMemcpyDeviceToHost));
      cudaFree(d_h);
      for (int p = 0; p < nranks && !failed; p++) {
         if (p == me) {
            D->peer_region[p] = D->region;
            continue;
         }
         void *ptr = NULL;
         cudaError_t e = cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess);
         if (e != cudaSuccess) {
            fprintf(stderr, "primme_b200: cudaIpcOpenMemHandle(halo, rank %d) failed: %s -- halo goes through NCCL\n", p,
                  cudaGetErrorString(e));
            cudaGetLastError();
            failed = 1;
            break;
         }
         D->peer_region[p] = (char *)ptr;
      }
      double f = (double)failed;  // collective decision, like pb200_ctx_peer_attach
      PB_CHK(pb200_allreduce_host(ctx, &f, 1));
      failed = f > 0.0;
   }
   if (failed) {
      for (int p = 0; p < nranks; p++)
         if (p != me && D->peer_region[p]) cudaIpcCloseMemHandle(D->peer_region[p]), D->peer_region[p] = NULL;
      int64_t tot = 0;
      for (int d = 0; d < nranks; d++) tot += D->send_cnt[d];
      PB_CUDA(cudaMalloc((void **)&D->d_stage, (size_t)(tot > 0 ? tot : 1) * D->rowbytes_max));
      D->peer_on = 0;
   } else {
      D->peer_on = 1;
   }
   return 0;
}

extern "C" int pb200_dist_csr_destroy(pb200_ctx *ctx, pb200_dist_csr *D) {
   if (!D) return 0;
   if (ctx) cudaStreamSynchronize(ctx->stream);
   if (D->stat_blocks > 0 && getenv("PB200_DIST_TIMING"))
      fprintf(stderr, "primme_b200 dist timing, rank %d: %lld blocks, push kernel %.3f ms per block, gather kernel (incl. waiting for the peers' rows) %.3f ms per block, %lld rows pushed per block\n",
            D->rank, (long long)D->stat_blocks, D->dbg_push_ms / (double)D->stat_blocks, D->dbg_spmm_ms / (double)D->stat_blocks,
            (long long)(D->stat_rows_sent / D->stat_blocks));
   if (D->nranks > 1 && D->peer_on && ctx) {
      // nobody unmaps a region a peer may still be writing to
      double z = 0.0;
      pb200_allreduce_host(ctx, &z, 1);
   }
   for (int p = 0; p < D->nranks; p++) {
      if (D->peer_on && p != D->rank && D->peer_region[p]) cudaIpcCloseMemHandle(D->peer_region[p]);
      cudaFree(D->d_send_rows[p]);
   }
   cudaFree(D->region), cudaFree(D->d_counters), cudaFree(D->d_stage);
   free(D->counts), free(D->displs), free(D);
   return 0;
}

extern "C" int pb200_dist_csr_info(const pb200_dist_csr *D, int64_t *nloc, int64_t *nhalo, int64_t *rows_sent_per_block,
      int *peer_halo) {
   if (nloc) *nloc = D->nloc;
   if (nhalo) *nhalo = D->nhalo;
   if (rows_sent_per_block) {
      int64_t t = 0;
      for (int d = 0; d < D->nranks; d++) t += D->send_cnt[d];
      *rows_sent_per_block = t;
   }
   if (peer_halo) *peer_halo = D->peer_on;
   return 0;
}

static int dist_spmm_any(pb200_ctx *ctx, pb200_dist_csr *D, const void *X, int64_t ldx, void *Y, int64_t ldy, int ncols) {
   const int cplx = D->A->is_complex;
   const size_t es = cplx ? 16 : 8;
   if (D->nranks <= 1)
      return cplx ? pb200_zspmm(ctx, D->A, X, ldx, Y, ldy, ncols)
                  : pb200_dspmm(ctx, D->A, (const double *)X, ldx, (double *)Y, ldy, ncols);
   const int me = D->rank;
   const size_t gb = gbuf_bytes(D->nloc, D->nhalo, D->rowbytes_max);
   for (int c0 = 0; c0 < ncols; c0 += 8) {
      const int b = ncols - c0 < 8 ? ncols - c0 : 8;
      const int bp = pb_spmm_bp(b);
      const size_t rowb = (size_t)bp * es;
      const char *Xc = (const char *)X + (size_t)c0 * ldx * es;
      char *Yc = (char *)Y + (size_t)c0 * ldy * es;
      const unsigned long long seq = ++D->seq;
      char *G = D->region + PB_DIST_HDR + (seq & 1) * gb;
      const double abytes = (double)(es + 4) * (double)D->A->nnz + 8.0 * (double)(D->A->nrows + 1) +
                            (double)es * (double)b * (double)(D->A->nrows + D->A->ncols);
      int ps = pb_prof_begin(ctx, PB_K_SPMM);
      PbPushArgs a;
      memset(&a, 0, sizeof(a));
      a.X = Xc, a.ldx = ldx, a.b = b, a.bp = bp, a.nloc = D->nloc, a.Glocal = G, a.seq = seq;
      int64_t tot = 0;
      // destinations in ROTATED order (me + 1, me + 2, ...): the push kernel walks its segments one after the other,
      // so with the natural order every rank would write to rank 0 first, then all to rank 1, ... -- one GPU's
      // NVLink ingress shared by seven senders while the other links idle (measured at N = 8: 0.90 ms per push of
      // 273 MB on seven of the eight ranks against 0.47 ms for the same volume at N = 2)
      for (int dk = 1; dk < D->nranks; dk++) {
         const int d = (me + dk) % D->nranks;
         if (D->send_cnt[d] == 0) continue;
         const int k = a.ndst++;
         a.dst_rank[k] = d;
         a.rows[k] = D->d_send_rows[d];
         a.start[k] = tot;
         if (D->peer_on) {
            const size_t gbd = gbuf_bytes(D->peer_nloc[d], D->peer_nhalo[d], D->rowbytes_max);
            char *Gd = D->peer_region[d] + PB_DIST_HDR + (seq & 1) * gbd;
            a.dstG[k] = Gd + (size_t)(D->peer_nloc[d] + D->send_off[d]) * rowb;
            a.flag[k] = (unsigned long long *)(D->peer_region[d]) + me;
         } else {
            a.dstG[k] = (char *)D->d_stage + (size_t)tot * rowb;
            a.flag[k] = NULL;
         }
         tot += D->send_cnt[d];
      }
      a.start[a.ndst] = tot;
      if (D->peer_on) {
         a.done = (const unsigned long long *)(D->region + 64);
         a.counter = D->d_counters;
      }
      int64_t blocks = ((D->nloc + tot) * bp + 255) / 256 / 4;
      const int64_t cap = (int64_t)ctx->num_sms * 8;
      if (blocks > cap) blocks = cap;
      if (blocks < 1) blocks = 1;
      static const int dbg_timing = getenv("PB200_DIST_TIMING") != NULL;
      cudaEvent_t dbg_ev[3] = {NULL, NULL, NULL};
      if (dbg_timing) {
         for (int i = 0; i < 3; i++) cudaEventCreate(&dbg_ev[i]);
         cudaEventRecord(dbg_ev[0], ctx->stream);
      }
      if (cplx) dist_push_kernel<double2><<<(int)blocks, 256, 0, ctx->stream>>>(a);
      else dist_push_kernel<double><<<(int)blocks, 256, 0, ctx->stream>>>(a);
      ctx->launches++;
      PB_CUDA(cudaGetLastError());
      if (dbg_timing) cudaEventRecord(dbg_ev[1], ctx->stream);
      D->stat_blocks++, D->stat_rows_sent += tot;

      PbSpSync sy;
      memset(&sy, 0, sizeof(sy));
      if (D->peer_on) {
         sy.flags = (const unsigned long long *)D->region;
         sy.seq = seq;
         sy.counter = D->d_counters + 1;
         for (int o = 0; o < D->nranks; o++) {
            if (o == me || D->recv_cnt[o] == 0) continue;
            sy.src_mask |= 1u << o, sy.ack_mask |= 1u << o;
            sy.ack[o] = (unsigned long long *)(D->peer_region[o] + 64) + me;
         }
      } else {
         // compacted rows through NCCL point-to-point, straight into the halo part of G
         PB_CHK(pb_nccl_group(ctx, 1));
         int64_t so = 0;
         for (int dk = 1; dk < D->nranks; dk++) {  // same order as the staging segments above
            const int d = (me + dk) % D->nranks;
            if (D->send_cnt[d] > 0) {
               PB_CHK(pb_nccl_send(ctx, (char *)D->d_stage + (size_t)so * rowb, (size_t)D->send_cnt[d] * rowb, d));
               so += D->send_cnt[d];
            }
            if (D->recv_cnt[d] > 0)
               PB_CHK(pb_nccl_recv(ctx, G + (size_t)(D->nloc + D->recv_off[d]) * rowb, (size_t)D->recv_cnt[d] * rowb, d));
         }
         PB_CHK(pb_nccl_group(ctx, 0));
      }
      int rc = pb_spmm_gathered(ctx, D->A, (const double *)G, bp, Yc, ldy, b, D->peer_on ? &sy : NULL);
      pb_prof_end(ctx, ps, abytes);
      if (dbg_timing) {
         cudaEventRecord(dbg_ev[2], ctx->stream);
         cudaEventSynchronize(dbg_ev[2]);
         float t01 = 0.f, t12 = 0.f;
         cudaEventElapsedTime(&t01, dbg_ev[0], dbg_ev[1]), cudaEventElapsedTime(&t12, dbg_ev[1], dbg_ev[2]);
         D->dbg_push_ms += t01, D->dbg_spmm_ms += t12;
         for (int i = 0; i < 3; i++) cudaEventDestroy(dbg_ev[i]);
      }
      PB_CHK(rc);
   }
   return 0;
}

// Y(local rows, 0:b) = A_local * [X_local; halo(X)]
extern "C" int pb200_ddist_spmm(pb200_ctx *ctx, pb200_dist_csr *D, const double *X, int64_t ldx,
      double *Y, int64_t ldy, int ncols) {
   if (D->A->is_complex) return PB200_ERR_ARG;
   return dist_spmm_any(ctx, D, X, ldx, Y, ldy, ncols);
}
extern "C" int pb200_zdist_spmm(pb200_ctx *ctx, pb200_dist_csr *D, const void *X, int64_t ldx, void *Y,
      int64_t ldy, int ncols) {
   if (!D->A->is_complex) return PB200_ERR_ARG;
   return dist_spmm_any(ctx, D, X, ldx, Y, ldy, ncols);
}

// primme.matrix = pb200_dist_csr*; primme.matrixMatvec = primme_b200_dist_csr_matvec
extern "C" void primme_b200_dist_csr_matvec(void *x, int64_t *ldx, void *y, int64_t *ldy,
      int *blockSize, struct primme_params *primme, int *ierr) {
   pb200_ctx *ctx = primme_b200_solver_ctx(primme);
   pb200_dist_csr *D = (pb200_dist_csr *)primme->matrix;
   if (!ctx || !D) {
      *ierr = -1;
      return;
   }
   *ierr = dist_spmm_any(ctx, D, x, *ldx, y, *ldy, *blockSize);
}

// ---- row-partitioned operator of the SVD front end (config C4: primme_svds with numProcs > 1) ----
// PRIMME's SPMD model for svds (reference include/primme_svds.h: mLocal rows of the left vectors, nLocal rows
// of the right vectors per process; src/svds/primme_svds_c.c:1323-1383 applies A then A^T): rank r owns the
// rows [m-range r] of A -- `A` below, applied to an n-side block gathered from all ranks -- and the rows
// [n-range r] of A^T -- `At`, applied to an m-side block.  Both directions are gather-type products over
// the compacted peer-memory halo; A^T y needs no reduce-scatter and stays deterministic.
extern "C" void primme_b200_svds_dist_csr_matvec(void *x, int64_t *ldx, void *y, int64_t *ldy, int *blockSize,
      int *transpose, struct primme_svds_params *primme_svds, int *ierr) {
   pb200_ctx *ctx = primme_b200_solver_ctx(&primme_svds->primme);
   if (!ctx) ctx = primme_b200_attached_ctx(&primme_svds->primme);
   const primme_b200_svds_dist *op = (const primme_b200_svds_dist *)primme_svds->matrix;
   if (!ctx || !op || !op->A || !op->At) {
      *ierr = -1;
      return;
   }
   *ierr = dist_spmm_any(ctx, *transpose ? op->At : op->A, x, *ldx, y, *ldy, *blockSize);
}

// vwxr.cu -- K5: fused basis update / Ritz vectors / residuals / Gram blocks (fp64).
//
// One sweep over the rows of V and W (reference src/eigs/auxiliary_eigs_normal.c:155-388,
// Num_update_VWXR_Sprimme with B = I):
//     xrow = V(r,:) * h,  yrow = W(r,:) * h           (m x nh coefficient block h, host-provided)
//     scatter column ranges of xrow / yrow to up to 3 + 1 destinations (may alias V / W)
//     R(r,:) = yrow - xrow .* theta, accumulate ||R_j||^2 (+ extra norms without storing)
//     G += xrow(0:nG)^T xrow(0:nG),  H += xrow(0:nH)^T yrow(0:nH)
//
// HBM-bound.  thread <-> row: V and W are read column by column, coalesced across the threads of
// a warp; h is broadcast from shared memory; the nh (<= NT) running sums of both products live
// in registers, so every row of V and W is read exactly once and only then written (in-place
// restart V <- V*h is safe: a row is owned by one thread).  G/H: the CTA parks its tile of
// xrow/yrow in shared memory and accumulates 2x2 register blocks per thread across all tiles;
// per-CTA partials are reduced in fixed order (pb_finish_panel).
#include "pb200_internal.cuh"
#include <math.h>
#include <string.h>

namespace {

constexpr int VT = 128;  // threads per CTA == rows per tile

struct VwxrArgs {
   const double *V;
   const double *W;
   int64_t n, ld;
   int m, nh;
   const double *hdev;      // m x nh, column stride m
   const double *thetadev;  // nh
   pb200_cols X[3];
   pb200_cols Wo;
   pb200_cols R;
   int want_Rnorms;
   int rb, re;
   int nG, nH;
   int need_y;              // whether W*h is needed at all
   double *partials;        // [grid][cnt] : Rnorms(nR) | rnorms(nr) | G(nG*nG) | H(nH*nH)
};

template <int NT>
__global__ void __launch_bounds__(VT) vwxr_kernel(VwxrArgs a) {
   extern __shared__ double smem[];
   double *hs = smem;                          // m * NT   ([k][c], zero padded)
   double *th = hs + (size_t)a.m * NT;         // NT
   const int ngh = a.nG > a.nH ? a.nG : a.nH;  // columns parked for G/H
   double *xs = th + NT;                       // ngh * VT
   double *ys = xs + (size_t)ngh * VT;         // nH * VT
   const int tid = threadIdx.x;

   for (int i = tid; i < a.m * NT; i += VT) {
      int k = i / NT, c = i % NT;
      hs[i] = c < a.nh ? a.hdev[k + (size_t)c * a.m] : 0.0;
   }
   for (int i = tid; i < NT; i += VT) th[i] = i < a.nh ? a.thetadev[i] : 0.0;
   __syncthreads();

   const int nR = a.R.ptr ? a.R.ce - a.R.cb : 0;
   const int nr = a.re - a.rb;
   // per-thread running squared norms (nR + nr <= NT)
   double nrm[NT];
#pragma unroll
   for (int c = 0; c < NT; c++) nrm[c] = 0.0;

   // G/H accumulators: 2x2 blocks, block id = tid + VT*t
   constexpr int MAXBLK = 4;  // up to 4 blocks of 2x2 per thread => covers 2*(nGb^2) <= 512 blocks
   double gacc[MAXBLK][4];
#pragma unroll
   for (int t = 0; t < MAXBLK; t++)
#pragma unroll
      for (int e = 0; e < 4; e++) gacc[t][e] = 0.0;
   const int nGb = (a.nG + 1) / 2, nHb = (a.nH + 1) / 2;
   const int nblkG = nGb * nGb, nblk = nblkG + nHb * nHb;

   const int64_t ntiles = (a.n + VT - 1) / VT;
   for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t r = tile * VT + tid;
      double x[NT], y[NT];
#pragma unroll
      for (int c = 0; c < NT; c++) x[c] = 0.0, y[c] = 0.0;
      if (r < a.n) {
         const double *vp = a.V + r;
         const double *wp = a.W + r;
         int k = 0;
         for (; k + 4 <= a.m; k += 4) {
            double v0 = vp[(size_t)(k + 0) * a.ld], v1 = vp[(size_t)(k + 1) * a.ld];
            double v2 = vp[(size_t)(k + 2) * a.ld], v3 = vp[(size_t)(k + 3) * a.ld];
            double w0 = 0, w1 = 0, w2 = 0, w3 = 0;
            if (a.need_y) {
               w0 = wp[(size_t)(k + 0) * a.ld], w1 = wp[(size_t)(k + 1) * a.ld];
               w2 = wp[(size_t)(k + 2) * a.ld], w3 = wp[(size_t)(k + 3) * a.ld];
            }
            const double *h0 = hs + (size_t)k * NT;
#pragma unroll
            for (int c = 0; c < NT; c++) {
               double c0 = h0[c], c1 = h0[NT + c], c2 = h0[2 * NT + c], c3 = h0[3 * NT + c];
               x[c] += v0 * c0;
               x[c] += v1 * c1;
               x[c] += v2 * c2;
               x[c] += v3 * c3;
               y[c] += w0 * c0;
               y[c] += w1 * c1;
               y[c] += w2 * c2;
               y[c] += w3 * c3;
            }
         }
         for (; k < a.m; k++) {
            double v0 = vp[(size_t)k * a.ld];
            double w0 = a.need_y ? wp[(size_t)k * a.ld] : 0.0;
            const double *h0 = hs + (size_t)k * NT;
#pragma unroll
            for (int c = 0; c < NT; c++) {
               x[c] += v0 * h0[c];
               y[c] += w0 * h0[c];
            }
         }
         // ---- scatter (all reads of this row are done: aliasing V/W is safe) ----
#pragma unroll
         for (int t = 0; t < 3; t++) {
            if (a.X[t].ptr) {
#pragma unroll
               for (int c = 0; c < NT; c++)
                  if (c >= a.X[t].cb && c < a.X[t].ce)
                     a.X[t].ptr[r + (size_t)(c - a.X[t].cb) * a.X[t].ld] = x[c];
            }
         }
         if (a.Wo.ptr) {
#pragma unroll
            for (int c = 0; c < NT; c++)
               if (c >= a.Wo.cb && c < a.Wo.ce)
                  a.Wo.ptr[r + (size_t)(c - a.Wo.cb) * a.Wo.ld] = y[c];
         }
#pragma unroll
         for (int c = 0; c < NT; c++) {
            double res = y[c] - x[c] * th[c];
            const bool inR = a.R.ptr && c >= a.R.cb && c < a.R.ce;
            if (inR) a.R.ptr[r + (size_t)(c - a.R.cb) * a.R.ld] = res;
            if (inR || (c >= a.rb && c < a.re)) nrm[c] += res * res;
         }
      }
      // ---- G / H ----
      if (nblk > 0) {
         // park (zero rows beyond n contribute nothing)
#pragma unroll
         for (int c = 0; c < NT; c++) {
            if (c < ngh) xs[(size_t)c * VT + tid] = x[c];
            if (c < a.nH) ys[(size_t)c * VT + tid] = y[c];
         }
         __syncthreads();
#pragma unroll
         for (int t = 0; t < MAXBLK; t++) {
            int blk = tid + VT * t;
            if (blk < nblk) {
               const double *L, *Rr;
               int bi, bj, nn;
               if (blk < nblkG) {
                  bi = blk % nGb, bj = blk / nGb, nn = a.nG, L = xs, Rr = xs;
               } else {
                  int bb = blk - nblkG;
                  bi = bb % nHb, bj = bb / nHb, nn = a.nH, L = xs, Rr = ys;
               }
               // G is symmetric: only blocks on/above the diagonal are computed
               if (blk >= nblkG || bi <= bj) {
                  int i0 = 2 * bi, j0 = 2 * bj;
                  int i1 = i0 + 1 < nn ? i0 + 1 : i0, j1 = j0 + 1 < nn ? j0 + 1 : j0;
                  const double *li0 = L + (size_t)i0 * VT, *li1 = L + (size_t)i1 * VT;
                  const double *rj0 = Rr + (size_t)j0 * VT, *rj1 = Rr + (size_t)j1 * VT;
                  double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
#pragma unroll 4
                  for (int rr = 0; rr < VT; rr++) {
                     double l0 = li0[rr], l1 = li1[rr], q0 = rj0[rr], q1 = rj1[rr];
                     s00 += l0 * q0;
                     s01 += l0 * q1;
                     s10 += l1 * q0;
                     s11 += l1 * q1;
                  }
                  gacc[t][0] += s00, gacc[t][1] += s01, gacc[t][2] += s10, gacc[t][3] += s11;
               }
            }
         }
         __syncthreads();
      }
   }

   // ---------------- epilogue: per-CTA partials ----------------
   // slots: squared norm of column c at out[c] for c < nn (nn = a.nh if any norm is wanted), G, H
   const int nn = (nR + nr) > 0 ? a.nh : 0;
   const int cnt = nn + a.nG * a.nG + a.nH * a.nH;
   double *out = a.partials + (size_t)blockIdx.x * cnt;
   // norms: block reduction through shared memory (reuse xs region is unsafe if nblk==0: use hs)
   __syncthreads();
   double *red = smem;  // VT doubles needed; hs no longer used
#pragma unroll
   for (int c = 0; c < NT; c++) {
      bool isR = a.R.ptr && c >= a.R.cb && c < a.R.ce;
      bool isr = !isR && c >= a.rb && c < a.re;
      if (!(isR || isr)) continue;  // uniform across the CTA
      double v = nrm[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((tid & 31) == 0) red[tid >> 5] = v;
      __syncthreads();
      if (tid == 0) {
         double s = 0.0;
         for (int w = 0; w < VT / 32; w++) s += red[w];
         out[c] = s;
      }
      __syncthreads();
   }
   if (nblk > 0) {
      double *Gout = out + nn;
      double *Hout = Gout + a.nG * a.nG;
#pragma unroll
      for (int t = 0; t < MAXBLK; t++) {
         int blk = tid + VT * t;
         if (blk >= nblk) continue;
         if (blk < nblkG) {
            int bi = blk % nGb, bj = blk / nGb;
            if (bi > bj) continue;
            for (int e = 0; e < 4; e++) {
               int i = 2 * bi + (e >> 1), j = 2 * bj + (e & 1);
               if (i < a.nG && j < a.nG) {
                  Gout[i + (size_t)j * a.nG] = gacc[t][e];
                  if (bi != bj) Gout[j + (size_t)i * a.nG] = gacc[t][e];
               }
            }
         } else {
            int bb = blk - nblkG;
            int bi = bb % nHb, bj = bb / nHb;
            for (int e = 0; e < 4; e++) {
               int i = 2 * bi + (e >> 1), j = 2 * bj + (e & 1);
               if (i < a.nH && j < a.nH) Hout[i + (size_t)j * a.nH] = gacc[t][e];
            }
         }
      }
   }
}

template <int NT>
int launch_vwxr(pb200_ctx *ctx, const VwxrArgs &a, int grid, size_t shmem) {
   auto kern = vwxr_kernel<NT>;
   if (shmem > 48 * 1024)
      PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
   kern<<<grid, VT, shmem, ctx->stream>>>(a);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return 0;
}

}  // namespace

extern "C" int pb200_dvwxr(pb200_ctx *ctx, int64_t n, const double *V, const double *W, int m,
      int64_t ld, const double *h_host, int ldh, int nh, const double *theta_host,
      const pb200_vwxr_out *o) {
   if (nh <= 0 || m < 0) return 0;
   if (nh > 64) return PB200_ERR_ARG;
   VwxrArgs a;
   memset(&a, 0, sizeof(a));
   a.V = V, a.W = W, a.n = n, a.ld = ld, a.m = m, a.nh = nh;
   for (int t = 0; t < 3; t++)
      if (o->X[t].ptr && o->X[t].ce > o->X[t].cb) a.X[t] = o->X[t];
   if (o->Wo.ptr && o->Wo.ce > o->Wo.cb) a.Wo = o->Wo;
   if (o->R.ptr && o->R.ce > o->R.cb) a.R = o->R;
   a.want_Rnorms = o->Rnorms_host != NULL;
   if (o->rnorms_host && o->re > o->rb) a.rb = o->rb, a.re = o->re;
   a.nG = o->G_host ? o->nG : 0;
   a.nH = o->H_host ? o->nH : 0;
   a.need_y = (a.Wo.ptr || a.R.ptr || a.re > a.rb || a.nH > 0) ? 1 : 0;
   const int nR = a.R.ptr ? a.R.ce - a.R.cb : 0, nr = a.re - a.rb;
   const int nn = (nR + nr) > 0 ? nh : 0;
   const int cnt = nn + a.nG * a.nG + a.nH * a.nH;
   {
      int gb = (a.nG + 1) / 2, hb = (a.nH + 1) / 2;
      if (gb * gb + hb * hb > 4 * VT) return PB200_ERR_ARG;  // nG,nH <= 32 each
   }

   // stage h (compacted to ld m) and theta
   size_t need = (size_t)m * nh + nh;
   PB_CHK(pb_ensure_small(ctx, need > (size_t)cnt ? need : (size_t)cnt));
   // the pinned staging buffer may still feed an earlier async copy
   PB_CUDA(cudaStreamSynchronize(ctx->stream));
   double *hp = ctx->h_pinned;
   for (int c = 0; c < nh; c++)
      for (int k = 0; k < m; k++) hp[k + (size_t)c * m] = h_host[k + (size_t)c * ldh];
   for (int c = 0; c < nh; c++) hp[(size_t)m * nh + c] = theta_host ? theta_host[c] : 0.0;
   PB_CUDA(cudaMemcpyAsync(
         ctx->d_small, hp, need * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
   a.hdev = ctx->d_small;
   a.thetadev = ctx->d_small + (size_t)m * nh;

   int grid = 1;
   if (n > 0) {
      const int64_t ntiles = (n + VT - 1) / VT;
      int64_t g = (int64_t)ctx->num_sms * 4;
      grid = (int)(ntiles < g ? ntiles : g);
   }
   if (cnt > 0) {
      PB_CHK(pb_ensure_partials(ctx, (size_t)grid * cnt));
      a.partials = ctx->d_partials;
      // G partial slots below the diagonal blocks are written by mirroring; zero everything
      PB_CUDA(cudaMemsetAsync(ctx->d_partials, 0, (size_t)grid * cnt * sizeof(double), ctx->stream));
   } else {
      PB_CHK(pb_ensure_partials(ctx, 16));
      a.partials = ctx->d_partials;
   }
   const int NT = nh <= 4 ? 4 : nh <= 8 ? 8 : nh <= 16 ? 16 : nh <= 24 ? 24 : nh <= 32 ? 32
                 : nh <= 40 ? 40 : nh <= 48 ? 48 : 64;
   const int ngh = a.nG > a.nH ? a.nG : a.nH;
   size_t shd = (size_t)m * NT + NT + (size_t)ngh * VT + (size_t)a.nH * VT;
   if (shd < VT) shd = VT;
   size_t shmem = shd * sizeof(double);
   int rc;
   // algorithmic bytes: V (and W) read once + every output column written once
   double ocols = 0;
   for (int t = 0; t < 3; t++) if (a.X[t].ptr) ocols += a.X[t].ce - a.X[t].cb;
   if (a.Wo.ptr) ocols += a.Wo.ce - a.Wo.cb;
   if (a.R.ptr) ocols += a.R.ce - a.R.cb;
   const double abytes = 8.0 * (double)n * ((a.need_y ? 2.0 : 1.0) * m + ocols);
   int ps = pb_prof_begin(ctx, PB_K_VWXR);
   switch (NT) {
   case 4: rc = launch_vwxr<4>(ctx, a, grid, shmem); break;
   case 8: rc = launch_vwxr<8>(ctx, a, grid, shmem); break;
   case 16: rc = launch_vwxr<16>(ctx, a, grid, shmem); break;
   case 24: rc = launch_vwxr<24>(ctx, a, grid, shmem); break;
   case 32: rc = launch_vwxr<32>(ctx, a, grid, shmem); break;
   case 40: rc = launch_vwxr<40>(ctx, a, grid, shmem); break;
   case 48: rc = launch_vwxr<48>(ctx, a, grid, shmem); break;
   default: rc = launch_vwxr<64>(ctx, a, grid, shmem); break;
   }
   pb_prof_end(ctx, ps, abytes);
   PB_CHK(rc);
   if (cnt > 0) {
      PB_CHK(pb_finish_panel(ctx, grid, cnt));
      const double *p = ctx->h_pinned;
      if (o->Rnorms_host)
         for (int c = 0; c < nR; c++) o->Rnorms_host[c] = sqrt(p[a.R.cb + c]);
      for (int c = 0; c < nr; c++) o->rnorms_host[c] = sqrt(p[a.rb + c]);
      const double *pg = p + nn;
      for (int j = 0; j < a.nG; j++)
         for (int i = 0; i < a.nG; i++) o->G_host[i + (size_t)j * o->ldG] = pg[i + (size_t)j * a.nG];
      const double *ph = pg + a.nG * a.nG;
      for (int j = 0; j < a.nH; j++)
         for (int i = 0; i < a.nH; i++) o->H_host[i + (size_t)j * o->ldH] = ph[i + (size_t)j * a.nH];
   }
   return 0;
}

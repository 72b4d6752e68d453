// vwxr.cu -- K5: fused basis update / Ritz vectors / residuals / Gram blocks (fp64).
//
// One sweep over the rows of V and W (reference src/eigs/auxiliary_eigs_normal.c:155-388,
// Num_update_VWXR_Sprimme with B = I):
//     xrow = V(r,:) * h,  yrow = W(r,:) * h           (m x nh coefficient block h, host-provided)
//     scatter column ranges of xrow / yrow to up to 3 + 1 destinations (may alias V / W)
//     R(r,:) = yrow - xrow .* theta, accumulate ||R_j||^2 (+ extra norms without storing)
//     G += xrow(0:nG)^T xrow(0:nG),  H += xrow(0:nH)^T yrow(0:nH)
//
// HBM-bound for the candidates sweep, instruction-bound for the wide restart sweep.  Three kernels:
//   vwxr_mma_kernel<NT8,MT,NW>   default: persistent CTAs, one producer thread streams row tiles of V and W with
//            2-D tensor-map TMA, each consumer warp owns 8 rows and forms V h and W h with DMMA m8n8k4 (h fragments
//            in registers / shared memory), residuals and norms in registers, G / H through a per-warp transposing
//            scratch, optionally P = [V R]^T R (the first Gram panel of the next block orthogonalisation);
//            a row is owned by one warp, so the in-place restart V <- V h is safe;
//   vwxr_wide_kernel<NTH,NG>     bulk-copy staged FMA kernel for wide coefficient blocks the MMA instances do not take;
//   vwxr_kernel<NT>              LDG fallback (thread <-> row), any alignment, n < 256.
// Per-CTA partials (norms, G, H, P) are finished inside the kernel in a fixed order (pb_finish_device).
#include "pb200_internal.cuh"
#include "tma_pipe.cuh"
#include <math.h>
#include <string.h>
#include <vector>
#include <stdlib.h>
#include <type_traits>

namespace {

constexpr int VT = 128;  // threads per CTA == rows per tile

#define PB_TILE(t) (a.rev ? ntiles - 1 - (t) : (t))  // see ortho_sweep.cu
struct VwxrArgs {
   int rev;
   const double *V;
   const double *W;
   int64_t n, ld;
   int m, nh;
   const double *hdev;      // m x nh, column stride m
   const double *thetadev;  // nh
   int coef_inline;         // 1: h and theta travel in `coef` below (kernel parameter space)
   pb200_cols X[3];
   pb200_cols Wo;
   pb200_cols R;
   int want_Rnorms;
   int rb, re;
   int nG, nH;
   int need_y;              // whether W*h is needed at all
   double *R2;              // second destination of the residual columns (MMA kernel only)
   int64_t ldR2;
   int nP;                  // > 0: also P = [V R]^T R, (m + nP) x nP with nP = R.ce - R.cb (MMA kernel only)
   int mpad;                // MMA kernel: m rounded up to a multiple of 4 (k-steps of the DMMA shape)
   int stage_doubles;       // MMA kernel: size of the stage ring (>= the end-of-kernel reduction scratch)
   double *partials;        // [grid][cnt] : Rnorms(nR) | rnorms(nr) | G(nG*nG) | H(nH*nH)
   PbFin fin;               // in-kernel panel finish (fin.cnt == 0: the host launches the reduction)
};

template <int NT>
__global__ void __launch_bounds__(VT) vwxr_kernel(VwxrArgs a, const __grid_constant__ PbCoef coef) {
   extern __shared__ double smem[];
   double *hs = smem;                          // m * NT   ([k][c], zero padded)
   double *th = hs + (size_t)a.m * NT;         // NT
   const int ngh = a.nG > a.nH ? a.nG : a.nH;  // columns parked for G/H
   double *xs = th + NT;                       // ngh * VT
   double *ys = xs + (size_t)ngh * VT;         // nH * VT
   const int tid = threadIdx.x;

   for (int i = tid; i < a.m * NT; i += VT) {
      int k = i / NT, c = i % NT;
      hs[i] = c < a.nh ? (a.coef_inline ? coef.v : a.hdev)[k + (size_t)c * a.m] : 0.0;
   }
   for (int i = tid; i < NT; i += VT) th[i] = i < a.nh ? (a.coef_inline ? coef.v + (size_t)a.m * a.nh : a.thetadev)[i] : 0.0;
   __syncthreads();

   const int nR = a.R.ptr ? a.R.ce - a.R.cb : 0;
   const int nr = a.re - a.rb;
   // per-thread running squared norms (nR + nr <= NT)
   double nrm[NT];
#pragma unroll
   for (int c = 0; c < NT; c++) nrm[c] = 0.0;

   // G/H accumulators: 2x2 blocks, block id = tid + VT*t
   constexpr int MAXBLK = 6;  // 2x2 blocks per thread: nGb^2 + nHb^2 <= 768 (nG, nH <= 38)
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
      if (!(isR || isr)) {  // uniform across the CTA; unused slots are still defined
         if (tid == 0 && c < nn) out[c] = 0.0;
         continue;
      }
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
   if (cnt > 0) pb_finish_device(a.fin, tid, VT, 15, reinterpret_cast<int *>(smem + 64));
}


// two consecutive rows with one 16-byte store; the very last row of an odd-length vector alone
__device__ __forceinline__ void st2(double *p, double v0, double v1, int64_t r, int64_t n) {
   if (r + 1 < n)
      *reinterpret_cast<double2 *>(p) = make_double2(v0, v1);
   else if (r < n)
      *p = v0;
}

// ------------------------------------------------------------------------------------------
// v3 for wide coefficient blocks (restart: V <- V*h, W <- W*h, next block, G = X'X, H = X'Y).
// fp64-FLOP bound on B200 (5.3 flop/B), so the goal is to keep the DFMA pipe busy:
//   * producer warp streams 64-row tiles of V and W (2m bulk copies) through a stage ring;
//   * the 8 consumer warps form 8/NG groups of NG warps; a group owns every (8/NG)-th tile;
//     inside a group warp cg owns NTH columns of h and lane <-> two consecutive rows, so a thread
//     accumulates 2 rows x NTH columns x {V*h, W*h} from 16-byte shared loads with h broadcast,
//     writes its outputs with 16-byte stores, keeps residual norms in registers and parks x/y
//     for the Gram phase;
//   * Gram phase per group: one 4x4 block of G (upper) or H per thread, row pairs visited in a
//     lane-rotated order (conflict-free 16-byte loads), accumulators live across all tiles.
template <int NTH, int NG>
__global__ void __launch_bounds__(256 + 32) vwxr_wide_kernel(VwxrArgs a, const __grid_constant__ PbCoef coef, int nstages, int park_cols) {
   constexpr int TR = 64;
   constexpr int NT = NTH * NG;
   constexpr int NTG = 8 / NG;      // tile groups
   constexpr int GS = 32 * NG;      // threads per group
   extern __shared__ __align__(128) unsigned char smraw[];
   const int m = a.m;
   double *stage0 = reinterpret_cast<double *>(smraw);                 // nstages * 2m * TR
   double *hs = stage0 + (size_t)nstages * 2 * m * TR;                 // m * NT
   double *th = hs + (size_t)m * NT;                                   // NT
   double *park0 = th + NT;                                            // NTG * 2 * park_cols * TR
   const int NB = NTG * nstages;  // barrier slots, one per (stage, tile group) pair (see narrow kernel)
   uint64_t *full = reinterpret_cast<uint64_t *>(park0 + (size_t)NTG * 2 * park_cols * TR);
   uint64_t *empty = full + NB;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

   if (tid == 0) {
      for (int u = 0; u < NB; u++) {
         pbtma::mbar_init(&full[u], 1);
         pbtma::mbar_init(&empty[u], NG);
      }
      pbtma::fence_barrier_init();
   }
   for (int i = tid; i < m * NT; i += 288) {
      int k = i / NT, c = i % NT;
      hs[i] = c < a.nh ? (a.coef_inline ? coef.v : a.hdev)[k + (size_t)c * m] : 0.0;
   }
   for (int i = tid; i < NT; i += 288) th[i] = i < a.nh ? (a.coef_inline ? coef.v + (size_t)a.m * a.nh : a.thetadev)[i] : 0.0;
   __syncthreads();

   const int64_t ntiles = (a.n + TR - 1) / TR;  // the last tile may be partial
   if (warp == 8) {
      int64_t i = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, i++) {
         const int s = (int)(i % nstages);
         const int u = (int)(i % NB);
         if (lane == 0 && i >= nstages) {
            const int64_t j = i - nstages;  // previous user of this stage
            pbtma::mbar_wait(&empty[j % NB], (uint32_t)((j / NB) & 1));
         }
         __syncwarp();
         double *dst = stage0 + (size_t)s * 2 * m * TR;
         const int64_t r0 = tile * TR;
         const int rows = (int)((a.n - r0) < TR ? (a.n - r0) : TR);
         const int rows_even = rows & ~1;
         if (rows < TR) {
            // partial tile: odd last row and zero padding by plain stores (see ortho_sweep.cu)
            for (int c = lane; c < 2 * m; c += 32) {
               const double *src = (c < m ? a.V + (size_t)c * a.ld : a.W + (size_t)(c - m) * a.ld) + r0;
               double *d = dst + (size_t)c * TR;
               for (int rr = rows_even; rr < TR; rr++) d[rr] = rr < rows ? src[rr] : 0.0;
            }
            __syncwarp();
         }
         if (lane == 0)
            pbtma::mbar_arrive_expect_tx(&full[u], (uint32_t)(2 * m * rows_even * sizeof(double)));
         __syncwarp();
         if (rows_even > 0)
            for (int c = lane; c < 2 * m; c += 32) {
               const double *src = (c < m ? a.V + (size_t)c * a.ld : a.W + (size_t)(c - m) * a.ld) + r0;
               pbtma::bulk_g2s(dst + (size_t)c * TR, src, rows_even * sizeof(double), &full[u]);
            }
      }
      return;
   }

   const int grp = warp / NG, cg = warp % NG;  // tile group, column group
   const int tg = cg * 32 + lane;              // thread index inside the group
   const int p2 = 2 * lane;                    // first of this thread's two rows
   const int c0 = cg * NTH;
   double *xs = park0 + (size_t)grp * 2 * park_cols * TR;
   double *ys = xs + (size_t)park_cols * TR;

   double nrm[NTH];
#pragma unroll
   for (int c = 0; c < NTH; c++) nrm[c] = 0.0;

   // Gram block of this thread: G upper blocks first, then all H blocks
   const int nb4g = (a.nG + 3) / 4, nb4h = (a.nH + 3) / 4;
   const int nblkG = nb4g * (nb4g + 1) / 2, nblk = nblkG + nb4h * nb4h;
   int bi = 0, bj = 0, isH = 0;
   const bool has_blk = tg < nblk;
   if (has_blk) {
      if (tg < nblkG) {
         int t = tg;  // enumerate (bi <= bj) column by column
         bj = 0;
         while (t > bj) t -= bj + 1, bj++;
         bi = t;
      } else {
         isH = 1;
         bi = (tg - nblkG) % nb4h, bj = (tg - nblkG) / nb4h;
      }
   }
   double gacc[4][4];
#pragma unroll
   for (int e = 0; e < 4; e++)
#pragma unroll
      for (int f = 0; f < 4; f++) gacc[e][f] = 0.0;

   int64_t i = grp;
   for (int64_t tile = blockIdx.x + (int64_t)grp * gridDim.x; tile < ntiles;
         tile += (int64_t)NTG * gridDim.x, i += NTG) {
      const int s = (int)(i % nstages);
      const int u = (int)(i % NB);
      pbtma::mbar_wait(&full[u], (uint32_t)((i / NB) & 1));
      const double *sv = stage0 + (size_t)s * 2 * m * TR + p2;
      const double *sw = sv + (size_t)m * TR;
      double x0[NTH], x1[NTH], y0[NTH], y1[NTH];
#pragma unroll
      for (int c = 0; c < NTH; c++) x0[c] = x1[c] = y0[c] = y1[c] = 0.0;
#pragma unroll 2
      for (int k = 0; k < m; k++) {
         const double2 v = *reinterpret_cast<const double2 *>(sv + (size_t)k * TR);
         const double2 w = *reinterpret_cast<const double2 *>(sw + (size_t)k * TR);
         const double *hk = hs + (size_t)k * NT + c0;
#pragma unroll
         for (int c = 0; c < NTH; c++) {
            const double hc = hk[c];
            x0[c] += v.x * hc, x1[c] += v.y * hc;
            y0[c] += w.x * hc, y1[c] += w.y * hc;
         }
      }
      __syncwarp();
      if (lane == 0) pbtma::mbar_arrive(&empty[u]);

      const int64_t r = tile * TR + p2;
#pragma unroll
      for (int c = 0; c < NTH; c++) {
         const int cc = c0 + c;
#pragma unroll
         for (int t = 0; t < 3; t++)
            if (a.X[t].ptr && cc >= a.X[t].cb && cc < a.X[t].ce)
               st2(a.X[t].ptr + r + (size_t)(cc - a.X[t].cb) * a.X[t].ld, x0[c], x1[c], r, a.n);
         if (a.Wo.ptr && cc >= a.Wo.cb && cc < a.Wo.ce)
            st2(a.Wo.ptr + r + (size_t)(cc - a.Wo.cb) * a.Wo.ld, y0[c], y1[c], r, a.n);
         const bool inR = a.R.ptr && cc >= a.R.cb && cc < a.R.ce;
         if (inR || (cc >= a.rb && cc < a.re)) {
            const double r0 = y0[c] - x0[c] * th[cc], r1 = y1[c] - x1[c] * th[cc];
            if (inR)
               st2(a.R.ptr + r + (size_t)(cc - a.R.cb) * a.R.ld, r0, r1, r, a.n);
            nrm[c] += r0 * r0 + r1 * r1;
         }
      }
      if (nblk > 0) {
#pragma unroll
         for (int c = 0; c < NTH; c++) {
            const int cc = c0 + c;
            if (cc < park_cols) {
               *reinterpret_cast<double2 *>(xs + (size_t)cc * TR + p2) = make_double2(x0[c], x1[c]);
               *reinterpret_cast<double2 *>(ys + (size_t)cc * TR + p2) = make_double2(y0[c], y1[c]);
            }
         }
         pbtma::named_bar_sync(1 + grp, GS);
         if (has_blk) {
            const double *L = xs + (size_t)(4 * bi) * TR;
            const double *Rr = (isH ? ys : xs) + (size_t)(4 * bj) * TR;
            // entries past the matrix edge are discarded in the epilogue: clamp their column
            // index to a valid parked column so that only initialised data is read
            const int nn = isH ? a.nH : a.nG;
            int li[4], rj[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
               li[e] = (4 * bi + e < nn ? e : 0) * TR;
               rj[e] = (4 * bj + e < nn ? e : 0) * TR;
            }
#pragma unroll 4
            for (int it = 0; it < TR / 2; it++) {
               const int q2 = 2 * ((it + lane) & (TR / 2 - 1));  // lane-rotated row pair
               double2 l[4], rr[4];
#pragma unroll
               for (int e = 0; e < 4; e++) {
                  l[e] = *reinterpret_cast<const double2 *>(L + li[e] + q2);
                  rr[e] = *reinterpret_cast<const double2 *>(Rr + rj[e] + q2);
               }
#pragma unroll
               for (int e = 0; e < 4; e++)
#pragma unroll
                  for (int f = 0; f < 4; f++) gacc[e][f] += l[e].x * rr[f].x + l[e].y * rr[f].y;
            }
         }
         pbtma::named_bar_sync(1 + grp, GS);  // the parked tile is rewritten by the group's next tile
      }
   }

   // ---------------- epilogue: one partial slot per (CTA, tile group) ----------------
   const int nR = a.R.ptr ? a.R.ce - a.R.cb : 0;
   const int nn = (nR + (a.re - a.rb)) > 0 ? a.nh : 0;
   const int cnt = nn + a.nG * a.nG + a.nH * a.nH;
   double *out = a.partials + ((size_t)blockIdx.x * NTG + grp) * cnt;
   if (nn > 0) {
#pragma unroll
      for (int c = 0; c < NTH; c++) {
         double v = nrm[c];
#pragma unroll
         for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
         if (lane == 0 && c0 + c < a.nh) out[c0 + c] = v;
      }
   }
   if (has_blk) {
      double *Gout = out + nn, *Hout = Gout + a.nG * a.nG;
#pragma unroll
      for (int e = 0; e < 4; e++)
#pragma unroll
         for (int f = 0; f < 4; f++) {
            const int ii = 4 * bi + e, jj = 4 * bj + f;
            if (isH) {
               if (ii < a.nH && jj < a.nH) Hout[ii + (size_t)jj * a.nH] = gacc[e][f];
            } else if (ii < a.nG && jj < a.nG) {
               Gout[ii + (size_t)jj * a.nG] = gacc[e][f];
               if (bi != bj) Gout[jj + (size_t)ii * a.nG] = gacc[e][f];
            }
         }
   }
   if (cnt > 0) pb_finish_device(a.fin, tid, 256, 15, reinterpret_cast<int *>(empty + NB));
}


// ------------------------------------------------------------------------------------------
// Main kernel: tensor-map TMA stages + fp64 tensor-core (DMMA m8n8k4) contractions, warp-local
// rows -- the same skeleton as ortho_sweep_mma_kernel.  A producer thread streams TR-row tiles of
// [V | W] (column stride TR+4 doubles: conflict-free fragment loads; columns m..mpad and rows past
// n are zero-filled by the TMA unit) into a ring of stages.  Each consumer warp owns 8 rows:
//   X(8 x 8 NT8) = V(8 x m) h,  Y = W h            2 * NT8 * mpad/4 DMMAs, h fragments from shared
//   R = Y - X diag(theta), residual norms, scatter of X / Y / R column ranges (fragment layout:
//        a warp store covers 4 columns x 8 consecutive rows)
//   G += X^T X (upper tiles), H += X^T Y            through a per-warp transposing scratch
//   P += [V R]^T R                                  (nP > 0) V^T fragments straight from the stage
// and keeps every accumulator in registers across all tiles of the CTA.  The warps meet once at
// the end (per-warp panels summed in warp order through the stage memory), then the in-kernel
// finish delivers the reduced panel.  fp64 has no tcgen05 kind: mma.sync DMMA is the tensor path.
__device__ __forceinline__ void vdmma884(double &d0, double &d1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(d0), "+d"(d1)
                : "d"(a), "d"(b));
}

struct VwxrMaps {
   CUtensorMap v, w;  // boxes of (TR+4) rows x mpad columns
};

constexpr int XW_LD = 12;  // row stride of the per-warp transposing scratch ([column][row])
// rows of columns 4..7 of every 8-column tile are XORed with 4: both the fragment-layout stores
// (columns 2t, 2t+1, row g) and the operand loads (column g, row t + 4 ks) are bank-conflict free
__device__ __forceinline__ int xw_at(int col, int row) { return col * XW_LD + (row ^ (col & 4)); }

template <int NT8, int MT, int NW>
__global__ void __launch_bounds__(NW * 32 + 32) vwxr_mma_kernel(VwxrArgs a, const __grid_constant__ PbCoef coef,
      const __grid_constant__ VwxrMaps maps, int nstages) {
   constexpr int TR = 8 * NW, S = TR + 4, NCT = NW * 32;
   constexpr int NC = 8 * NT8;       // padded column count of h
   constexpr int HS_LD = NC + 4;     // row stride of h in shared memory (= 4 or 12 mod 16)
   constexpr int NTG = NT8;          // 8-column tiles of G / H
   extern __shared__ __align__(128) unsigned char smraw[];
   const int m = a.m, mpad = a.mpad;
   const int stage_sz = 2 * mpad * S;
   double *stage0 = reinterpret_cast<double *>(smraw);   // ring (later: reduction scratch)
   double *hs = stage0 + a.stage_doubles;                // mpad * HS_LD
   double *th = hs + (size_t)mpad * HS_LD;               // NC
   double *xw0 = th + NC;                                // NW * 2 * NC * XW_LD
   uint64_t *full = reinterpret_cast<uint64_t *>(xw0 + (size_t)NW * 2 * NC * XW_LD);
   uint64_t *empty = full + nstages;
   int *flag = reinterpret_cast<int *>(empty + nstages);
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

   if (tid == 0) {
      for (int s = 0; s < nstages; s++) {
         pbtma::mbar_init(&full[s], 1);
         pbtma::mbar_init(&empty[s], NW);
      }
      pbtma::fence_barrier_init();
   }
   {
      const double *hg = a.coef_inline ? coef.v : a.hdev;
      const double *tg = a.coef_inline ? coef.v + (size_t)m * a.nh : a.thetadev;
      for (int i = tid; i < mpad * NC; i += NCT + 32) {
         const int k = i / NC, c = i % NC;
         hs[k * HS_LD + c] = (k < m && c < a.nh) ? hg[k + (size_t)c * m] : 0.0;
      }
      for (int i = tid; i < NC; i += NCT + 32) th[i] = i < a.nh ? tg[i] : 0.0;
   }
   __syncthreads();

   const int64_t ntiles = (a.n + TR - 1) / TR;
   if (warp == NW) {
      if (lane != 0) return;
      int s = 0;
      uint32_t ph = 0;
      const uint32_t bytes = (uint32_t)(stage_sz * sizeof(double));
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
         pbtma::mbar_wait(&empty[s], ph ^ 1);
         double *dst = stage0 + (size_t)s * stage_sz;
         const int r0 = (int)(PB_TILE(tile) * TR);
         pbtma::mbar_arrive_expect_tx(&full[s], bytes);
         pbtma::tensor_g2s_2d(dst, &maps.v, r0, 0, &full[s]);
         pbtma::tensor_g2s_2d(dst + mpad * S, &maps.w, r0, 0, &full[s]);
         if (++s == nstages) s = 0, ph ^= 1;
      }
      return;
   }

   // ------------------------------ consumer warps ------------------------------
   const int g = lane >> 2, t = lane & 3;
   const int r0w = warp * 8;
   const int nks = mpad >> 2;
   const int nR = a.R.ptr ? a.R.ce - a.R.cb : 0;
   const bool gh = a.nG > 0 || a.nH > 0;
   const int nmtv = a.nP > 0 ? (m + 7) >> 3 : 0;     // tiles of V in the P panel
   double *xw = xw0 + (size_t)warp * 2 * NC * XW_LD;  // X (or R in P mode), [column][row]
   double *yw = xw + NC * XW_LD;                      // Y

   // Loop-invariant store plan of this thread's 2 NT8 columns (slot q = 2 i + j <-> column
   // 8 i + 2 t + j): one destination for the V h value, one for the W h value or the residual.
   // A column with more destinations than that (locking: X[2] overlapping X[0]) takes the
   // generic path.
   constexpr int NQ = 2 * NT8;
   double *px[NQ], *py[NQ];
   double thv[NQ];
   unsigned yres = 0, xres = 0, wnorm = 0, slow = 0;
#pragma unroll
   for (int q = 0; q < NQ; q++) {
      const int cc = 8 * (q >> 1) + 2 * t + (q & 1);
      px[q] = py[q] = nullptr;
      thv[q] = th[cc];
      int nx = 0, ny = 0;
#pragma unroll
      for (int u = 0; u < 3; u++)
         if (a.X[u].ptr && cc >= a.X[u].cb && cc < a.X[u].ce) {
            if (!nx) px[q] = a.X[u].ptr + (size_t)(cc - a.X[u].cb) * a.X[u].ld;
            nx++;
         }
      if (a.Wo.ptr && cc >= a.Wo.cb && cc < a.Wo.ce) py[q] = a.Wo.ptr + (size_t)(cc - a.Wo.cb) * a.Wo.ld, ny++;
      const bool inR = nR > 0 && cc >= a.R.cb && cc < a.R.ce;
      if (inR) {
         if (!ny) py[q] = a.R.ptr + (size_t)(cc - a.R.cb) * a.R.ld, yres |= 1u << q;
         ny++;
         if (a.R2) {
            if (!nx) px[q] = a.R2 + (size_t)(cc - a.R.cb) * a.ldR2, xres |= 1u << q;
            nx++;
         }
      }
      if (inR || (cc >= a.rb && cc < a.re)) wnorm |= 1u << q;
      if (nx > 1 || ny > 1) slow |= 1u << q;
   }
   slow = __any_sync(0xffffffffu, slow != 0) ? 1u : 0u;  // warp-uniform choice of the path

   double nrm[NT8][2];
   double Gacc[NTG * (NTG + 1) / 2][2], Hacc[NTG * NTG][2];
   double Pacc[MT > 0 ? MT : 1][2], RRacc[2];
#pragma unroll
   for (int i = 0; i < NT8; i++) nrm[i][0] = nrm[i][1] = 0.0;
#pragma unroll
   for (int i = 0; i < NTG * (NTG + 1) / 2; i++) Gacc[i][0] = Gacc[i][1] = 0.0;
#pragma unroll
   for (int i = 0; i < NTG * NTG; i++) Hacc[i][0] = Hacc[i][1] = 0.0;
#pragma unroll
   for (int i = 0; i < (MT > 0 ? MT : 1); i++) Pacc[i][0] = Pacc[i][1] = 0.0;
   RRacc[0] = RRacc[1] = 0.0;

   const int offu = r0w + t * S + g;  // A fragment of V h: column t of a k-step, row g
   const int offg = r0w + g * S + t;  // A fragment of V^T R: column g of a tile, row t of a k-step
   const double *hb = hs + t * HS_LD + g;
   // scratch offsets: fragment-layout store of column 2 t + j (tile i adds 8 i XW_LD), operand
   // load of column g, row t + 4 ks
   const int so0 = xw_at(2 * t, g), so1 = xw_at(2 * t + 1, g);
   const int lo0 = xw_at(g, t), lo1 = xw_at(g, t + 4);
   const int pcol0 = xw_at((2 * t - a.R.cb) & 7, g), pcol1 = xw_at((2 * t + 1 - a.R.cb) & 7, g);

   int s = 0;
   uint32_t ph = 0;
   for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      pbtma::mbar_wait(&full[s], ph);
      const double *st = stage0 + (size_t)s * stage_sz;
      double X[NT8][2], Y[NT8][2];
#pragma unroll
      for (int i = 0; i < NT8; i++) X[i][0] = X[i][1] = Y[i][0] = Y[i][1] = 0.0;
      {
         const double *pv = st + offu, *pw = pv + mpad * S, *pb = hb;
#pragma unroll 2
         for (int ks = 0; ks < nks; ks++) {
            const double av = pv[0], aw = pw[0];
#pragma unroll
            for (int i = 0; i < NT8; i++) {
               const double bf = pb[8 * i];
               vdmma884(X[i][0], X[i][1], av, bf);
               vdmma884(Y[i][0], Y[i][1], aw, bf);
            }
            pv += 4 * S, pw += 4 * S, pb += 4 * HS_LD;
         }
      }
      if (MT == 0) {  // the stage is not needed any more (P mode reads V^T from it below)
         __syncwarp();
         if (lane == 0) pbtma::mbar_arrive(&empty[s]);
      }
      // ---- residuals, norms, scatter (fragment layout: row g, columns 8 i + 2 t + j) ----
      const int64_t r = PB_TILE(tile) * TR + r0w + g;
      const bool rok = r < a.n;
      if (!slow) {
#pragma unroll
         for (int q = 0; q < NQ; q++) {
            const double x = X[q >> 1][q & 1], y = Y[q >> 1][q & 1];
            const double res = y - x * thv[q];
            if (wnorm & (1u << q)) nrm[q >> 1][q & 1] += res * res;
            if (rok) {
               if (px[q]) px[q][r] = (xres & (1u << q)) ? res : x;
               if (py[q]) py[q][r] = (yres & (1u << q)) ? res : y;
            }
            if (MT > 0)
               xw[(q & 1) ? pcol1 : pcol0] = (yres & (1u << q)) ? res : 0.0;
         }
      } else {
#pragma unroll
         for (int i = 0; i < NT8; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
               const int cc = 8 * i + 2 * t + j;
               const double x = X[i][j], y = Y[i][j];
               const double res = y - x * th[cc];
               const bool inR = nR > 0 && cc >= a.R.cb && cc < a.R.ce;
               if (inR || (cc >= a.rb && cc < a.re)) nrm[i][j] += res * res;
               if (rok) {
#pragma unroll
                  for (int u = 0; u < 3; u++)
                     if (a.X[u].ptr && cc >= a.X[u].cb && cc < a.X[u].ce)
                        a.X[u].ptr[r + (size_t)(cc - a.X[u].cb) * a.X[u].ld] = x;
                  if (a.Wo.ptr && cc >= a.Wo.cb && cc < a.Wo.ce)
                     a.Wo.ptr[r + (size_t)(cc - a.Wo.cb) * a.Wo.ld] = y;
                  if (inR) a.R.ptr[r + (size_t)(cc - a.R.cb) * a.R.ld] = res;
                  if (inR && a.R2) a.R2[r + (size_t)(cc - a.R.cb) * a.ldR2] = res;
               }
               if (MT > 0) xw[j ? pcol1 : pcol0] = inR ? res : 0.0;
            }
      }
      if (MT > 0) {
         // P mode (NT8 == 1): the residual block sits compacted in scratch columns 0..nR-1 (zeros
         // in the others: cc -> (cc - R.cb) mod 8 is a bijection of the 8 columns)
         __syncwarp();
         const double *pg = st + offg;
#pragma unroll
         for (int ks = 0; ks < 2; ks++) {
            const double bf = xw[ks ? lo1 : lo0];
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
               if (mt < nmtv) vdmma884(Pacc[mt][0], Pacc[mt][1], pg[mt * 8 * S + 4 * ks], bf);
            vdmma884(RRacc[0], RRacc[1], bf, bf);
         }
         __syncwarp();
         if (lane == 0) pbtma::mbar_arrive(&empty[s]);
      } else if (gh) {
         // every tile of G (upper) and H is accumulated, whatever nG and nH: columns past them
         // are simply not read back
#pragma unroll
         for (int i = 0; i < NT8; i++) {
            xw[8 * i * XW_LD + so0] = X[i][0], xw[8 * i * XW_LD + so1] = X[i][1];
            yw[8 * i * XW_LD + so0] = Y[i][0], yw[8 * i * XW_LD + so1] = Y[i][1];
         }
         __syncwarp();
#pragma unroll
         for (int ks = 0; ks < 2; ks++) {
            double xf[NTG], yf[NTG];
#pragma unroll
            for (int i = 0; i < NTG; i++) {
               xf[i] = xw[8 * i * XW_LD + (ks ? lo1 : lo0)];
               yf[i] = yw[8 * i * XW_LD + (ks ? lo1 : lo0)];
            }
            int ig = 0;
#pragma unroll
            for (int tj = 0; tj < NTG; tj++)
#pragma unroll
               for (int ti = 0; ti < NTG; ti++) {
                  if (ti <= tj) vdmma884(Gacc[ig][0], Gacc[ig][1], xf[ti], xf[tj]), ig++;
                  vdmma884(Hacc[ti + NTG * tj][0], Hacc[ti + NTG * tj][1], xf[ti], yf[tj]);
               }
         }
         __syncwarp();  // the scratch is rewritten by the next tile
      }
      if (++s == nstages) s = 0, ph ^= 1;
   }

   // ---- per-warp panels -> one partial panel per CTA (warp order), through the stage memory ----
   const int nn = (nR + (a.re - a.rb)) > 0 ? a.nh : 0;
   const int prow = m + a.nP;
   const int cnt = nn + a.nG * a.nG + a.nH * a.nH + prow * a.nP;
   if (cnt <= 0) return;
   constexpr int NGT = NTG * (NTG + 1) / 2, NHT = NTG * NTG, NPT = MT > 0 ? MT + 1 : 0;
   constexpr int WSZ = (NGT + NHT + NPT) * 64 + NC;  // doubles per warp
   pbtma::named_bar_sync(1, NCT);  // every warp is done with the stages
   double *red = stage0 + (size_t)warp * WSZ;
   auto put = [&](int tile_id, const double *acc2) {
      *reinterpret_cast<double2 *>(red + ((size_t)tile_id * 8 + g) * 8 + 2 * t) = make_double2(acc2[0], acc2[1]);
   };
   if (MT == 0) {
#pragma unroll
      for (int i = 0; i < NGT; i++) put(i, Gacc[i]);
#pragma unroll
      for (int i = 0; i < NHT; i++) put(NGT + i, Hacc[i]);
   } else {
#pragma unroll
      for (int i = 0; i < (MT > 0 ? MT : 1); i++) put(NGT + NHT + i, Pacc[i]);
      put(NGT + NHT + MT, RRacc);
   }
#pragma unroll
   for (int i = 0; i < NT8; i++)
#pragma unroll
      for (int j = 0; j < 2; j++) {
         double v = nrm[i][j];
         v += __shfl_xor_sync(0xffffffffu, v, 4);
         v += __shfl_xor_sync(0xffffffffu, v, 8);
         v += __shfl_xor_sync(0xffffffffu, v, 16);
         if (g == 0) red[(NGT + NHT + NPT) * 64 + 8 * i + 2 * t + j] = v;
      }
   pbtma::named_bar_sync(1, NCT);
   double *out = a.partials + (size_t)blockIdx.x * cnt;
   const int offG = nn, offH = offG + a.nG * a.nG, offP = offH + a.nH * a.nH;
   for (int e = tid; e < cnt; e += NCT) {
      int idx;  // index inside a warp's block of `red`
      if (e < offG) {
         idx = (NGT + NHT + NPT) * 64 + e;
      } else if (e < offH) {
         int i = (e - offG) % a.nG, j = (e - offG) / a.nG;
         if (i > j) { const int sw = i; i = j, j = sw; }  // G is symmetric: upper tiles only
         const int ti = i >> 3, tj = j >> 3;
         idx = ((tj * (tj + 1) / 2 + ti) * 8 + (i & 7)) * 8 + (j & 7);
      } else if (e < offP) {
         const int i = (e - offH) % a.nH, j = (e - offH) / a.nH;
         idx = ((NGT + (i >> 3) + NTG * (j >> 3)) * 8 + (i & 7)) * 8 + (j & 7);
      } else {
         const int i = (e - offP) % prow, j = (e - offP) / prow;
         idx = i < m ? ((NGT + NHT + (i >> 3)) * 8 + (i & 7)) * 8 + j : ((NGT + NHT + MT) * 8 + (i - m)) * 8 + j;
      }
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < NW; w++) sum += stage0[(size_t)w * WSZ + idx];
      out[e] = sum;
   }
   pb_finish_device(a.fin, tid, NCT, 15, flag);
}

template <int NT8, int MT, int NW>
int launch_vwxr_mma(pb200_ctx *ctx, const VwxrArgs &a, const VwxrMaps &maps, int grid, size_t shmem, int nstages) {
   auto kern = vwxr_mma_kernel<NT8, MT, NW>;
   static size_t attr_shmem = 0;
   if (shmem > attr_shmem) {
      PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
      attr_shmem = shmem;
   }
   if (getenv("PB200_DEBUG")) {
      static size_t seen = 0;
      if (seen != shmem) {
         seen = shmem;
         int occ = 0;
         cudaFuncAttributes fa;
         cudaFuncGetAttributes(&fa, kern);
         cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NW * 32 + 32, shmem);
         fprintf(stderr, "primme_b200: vwxr_mma<%d,%d,%d> dyn smem %zu regs %d stages %d -> %d CTA/SM\n", NT8, MT, NW,
               shmem, fa.numRegs, nstages, occ);
      }
   }
   kern<<<grid, NW * 32 + 32, shmem, ctx->stream>>>(a, ctx->coef, maps, nstages);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return 0;
}

// ------------------------------------------------------------------------------------------
// Restart kernel (8 < nh <= 48): the columns of h are split over the warps, so the register file
// of a warp holds one 8-column tile of X = V h / Y = W h and a share of the Gram tiles instead of
// all of them (the one-warp-per-8-rows kernel above needs 168 registers at nh = 24 and cannot
// hold nh > 32 at all: 8 warps per SM, instruction-latency bound).
//   warp (rg, CG): row group rg (16 rows of the tile), column tile CG (columns 8 CG .. 8 CG + 7)
//   phase 1   X, Y for 2 x 8 rows: per k-step TWO 16-byte fragment loads (rows 2g, 2g+1 of column
//             t + 4 ks of V and of W: fragment row g of the first / second 8-row half) feed FOUR DMMAs
//             whose B fragment (h) is one 8-byte load; residuals, norms and the column scatter from
//             the accumulator layout with 16-byte stores (two consecutive rows of a column)
//   phase 2   the warps of a row group exchange their tiles through a double-buffered transposing
//             scratch ([column][16 rows], one named barrier per tile) and each accumulates every
//             NT8-th tile of G = X^T X (upper tiles) and H = X^T Y; the k index of these products
//             is the row inside the group in the order 4 t + ks, so a fragment for all four k-steps
//             is two 16-byte loads
// 12-16 consumer warps per SM; a row is written by the warps of ONE row group after the whole
// tile has landed in shared memory, so the in-place restart V <- V h stays safe.
constexpr int CG_LD = 18;  // row stride (doubles) of the exchange scratch: conflict-free 16-byte stores and loads

template <int NT8>
struct CgTiles {
   static constexpr int NGT = NT8 * (NT8 + 1) / 2, NHT = NT8 * NT8, NTILES = NGT + NHT;
   // tile id -> (ti, tj): G ids tj (tj + 1) / 2 + ti (ti <= tj), H ids NGT + ti + NT8 tj
   __host__ __device__ static constexpr int is_h(int e) { return e >= NGT; }
   __host__ __device__ static constexpr int tj(int e) {
      if (e >= NGT) return (e - NGT) / NT8;
      int j = 0;
      while ((j + 1) * (j + 2) / 2 <= e) j++;
      return j;
   }
   __host__ __device__ static constexpr int ti(int e) {
      if (e >= NGT) return (e - NGT) % NT8;
      return e - tj(e) * (tj(e) + 1) / 2;
   }
};

template <int I, int N, class F>
__device__ __forceinline__ void cg_static_for(F &&f) {
   if constexpr (I < N) {
      f(std::integral_constant<int, I>{});
      cg_static_for<I + 1, N>(f);
   }
}

template <int NT8, int NRG, int CG>
__device__ __forceinline__ void vwxr_cg_consumer(const VwxrArgs &a, const double *stage0, int stage_sz, const double *hs,
      const double *th, double *scr0, uint64_t *full, uint64_t *empty, int nstages, int rgw, int lane) {
   using T = CgTiles<NT8>;
   constexpr int TR = 16 * NRG, S = TR + 4, NC = 8 * NT8, HS_LD = NC + 4;
   constexpr int NA = (T::NTILES - CG + NT8 - 1) / NT8;  // tiles of G / H this warp accumulates
   const int g = lane >> 2, t = lane & 3;
   const int mpad = a.mpad, nks = mpad >> 2;
   const int64_t ntiles = (a.n + TR - 1) / TR;
   const bool gh = a.nG > 0 || a.nH > 0;
   const int nR = a.R.ptr ? a.R.ce - a.R.cb : 0;

   // store plan of this thread's two columns 8 CG + 2 t + j
   double *px[2][3], *py[2], *pr[2];
   double thv[2];
   bool wn[2];
#pragma unroll
   for (int j = 0; j < 2; j++) {
      const int cc = 8 * CG + 2 * t + j;
      thv[j] = th[cc];
#pragma unroll
      for (int u = 0; u < 3; u++)
         px[j][u] = (a.X[u].ptr && cc >= a.X[u].cb && cc < a.X[u].ce) ? a.X[u].ptr + (size_t)(cc - a.X[u].cb) * a.X[u].ld : nullptr;
      py[j] = (a.Wo.ptr && cc >= a.Wo.cb && cc < a.Wo.ce) ? a.Wo.ptr + (size_t)(cc - a.Wo.cb) * a.Wo.ld : nullptr;
      const bool inR = nR > 0 && cc >= a.R.cb && cc < a.R.ce;
      pr[j] = inR ? a.R.ptr + (size_t)(cc - a.R.cb) * a.R.ld : nullptr;
      wn[j] = inR || (cc >= a.rb && cc < a.re);
   }
   double nrm[2] = {0.0, 0.0};
   double acc[NA][2];
#pragma unroll
   for (int i = 0; i < NA; i++) acc[i][0] = acc[i][1] = 0.0;

   const int offu = t * S + 16 * rgw + 2 * g;
   const double *hb = hs + t * HS_LD + 8 * CG + g;
   double *scr = scr0 + (size_t)rgw * 2 * 2 * NC * CG_LD;       // [buffer][X | Y][column][CG_LD]
   const int so = (8 * CG + 2 * t) * CG_LD + 2 * g;             // own tile, column 2 t, rows 2 g, 2 g + 1
   const int lo = g * CG_LD + 4 * t;                            // operand: column g of a tile, rows 4 t .. 4 t + 3

   int s = 0, it = 0;
   uint32_t ph = 0;
   for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it ^= 1) {
      pbtma::mbar_wait(&full[s], ph);
      const double *st = stage0 + (size_t)s * stage_sz;
      double X[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, Y[2][2] = {{0.0, 0.0}, {0.0, 0.0}};  // [row half][column 2 t + j]
      {
         const double *pv = st + offu, *pw = pv + mpad * S, *pb = hb;
#pragma unroll 4
         for (int ks = 0; ks < nks; ks++) {
            const double2 av = *reinterpret_cast<const double2 *>(pv);
            const double2 aw = *reinterpret_cast<const double2 *>(pw);
            const double bf = pb[0];
            vdmma884(X[0][0], X[0][1], av.x, bf);
            vdmma884(X[1][0], X[1][1], av.y, bf);
            vdmma884(Y[0][0], Y[0][1], aw.x, bf);
            vdmma884(Y[1][0], Y[1][1], aw.y, bf);
            pv += 4 * S, pw += 4 * S, pb += 4 * HS_LD;
         }
      }
      __syncwarp();
      if (lane == 0) pbtma::mbar_arrive(&empty[s]);
      // ---- residuals, norms, scatter: rows r, r + 1 of columns 8 CG + 2 t + j ----
      const int64_t r = PB_TILE(tile) * TR + 16 * rgw + 2 * g;
      const bool ok2 = r + 1 < a.n, ok1 = r < a.n;
#pragma unroll
      for (int j = 0; j < 2; j++) {
         const double x0 = X[0][j], x1 = X[1][j], y0 = Y[0][j], y1 = Y[1][j];
         const double e0 = y0 - x0 * thv[j], e1 = y1 - x1 * thv[j];
         if (wn[j]) nrm[j] += e0 * e0 + e1 * e1;
         if (ok2) {
#pragma unroll
            for (int u = 0; u < 3; u++)
               if (px[j][u]) *reinterpret_cast<double2 *>(px[j][u] + r) = make_double2(x0, x1);
            if (py[j]) *reinterpret_cast<double2 *>(py[j] + r) = make_double2(y0, y1);
            if (pr[j]) *reinterpret_cast<double2 *>(pr[j] + r) = make_double2(e0, e1);
         } else if (ok1) {
#pragma unroll
            for (int u = 0; u < 3; u++)
               if (px[j][u]) px[j][u][r] = x0;
            if (py[j]) py[j][r] = y0;
            if (pr[j]) pr[j][r] = e0;
         }
      }
      if (gh) {
         double *xs = scr + (size_t)it * 2 * NC * CG_LD, *ys = xs + NC * CG_LD;
         *reinterpret_cast<double2 *>(xs + so) = make_double2(X[0][0], X[1][0]);
         *reinterpret_cast<double2 *>(xs + so + CG_LD) = make_double2(X[0][1], X[1][1]);
         *reinterpret_cast<double2 *>(ys + so) = make_double2(Y[0][0], Y[1][0]);
         *reinterpret_cast<double2 *>(ys + so + CG_LD) = make_double2(Y[0][1], Y[1][1]);
         pbtma::named_bar_sync(1 + rgw, 32 * NT8);
#pragma unroll
         for (int half = 0; half < 2; half++) {
            double2 xf[NT8], yf[NT8];
#pragma unroll
            for (int i = 0; i < NT8; i++) {
               xf[i] = *reinterpret_cast<const double2 *>(xs + 8 * i * CG_LD + lo + 2 * half);
               yf[i] = *reinterpret_cast<const double2 *>(ys + 8 * i * CG_LD + lo + 2 * half);
            }
            cg_static_for<0, NA>([&](auto qc) {
               constexpr int q = decltype(qc)::value, e = CG + q * NT8;
               if constexpr (e < T::NTILES) {
                  constexpr int ti = T::ti(e), tj = T::tj(e);
                  const double2 af = xf[ti];
                  const double2 bf = T::is_h(e) ? yf[tj] : xf[tj];
                  vdmma884(acc[q][0], acc[q][1], af.x, bf.x);
                  vdmma884(acc[q][0], acc[q][1], af.y, bf.y);
               }
            });
         }
      }
      if (++s == nstages) s = 0, ph ^= 1;
   }

   // ---- per-warp tiles -> one partial panel per CTA: row groups summed in order, through the stage memory ----
   const int nn = (nR + (a.re - a.rb)) > 0 ? a.nh : 0;
   const int cnt = nn + a.nG * a.nG + a.nH * a.nH;
   if (cnt <= 0) return;
   constexpr int NCT = NT8 * NRG * 32;
   constexpr int WSZ = T::NTILES * 64 + NC;  // doubles per row group
   pbtma::named_bar_sync(8, NCT);  // every warp is done with the stages
   double *red = const_cast<double *>(stage0) + (size_t)rgw * WSZ;
   cg_static_for<0, NA>([&](auto qc) {
      constexpr int q = decltype(qc)::value, e = CG + q * NT8;
      if constexpr (e < T::NTILES)
         *reinterpret_cast<double2 *>(red + ((size_t)e * 8 + g) * 8 + 2 * t) = make_double2(acc[q][0], acc[q][1]);
   });
#pragma unroll
   for (int j = 0; j < 2; j++) {
      double v = nrm[j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (g == 0) red[T::NTILES * 64 + 8 * CG + 2 * t + j] = v;
   }
}

template <int NT8, int NRG>
__global__ void __launch_bounds__(NT8 * NRG * 32 + 32) vwxr_cg_kernel(VwxrArgs a, const __grid_constant__ PbCoef coef,
      const __grid_constant__ VwxrMaps maps, int nstages) {
   using T = CgTiles<NT8>;
   constexpr int NW = NT8 * NRG, TR = 16 * NRG, S = TR + 4, NCT = NW * 32;
   constexpr int NC = 8 * NT8, HS_LD = NC + 4;
   extern __shared__ __align__(128) unsigned char smraw[];
   const int m = a.m, mpad = a.mpad;
   const int stage_sz = 2 * mpad * S;
   double *stage0 = reinterpret_cast<double *>(smraw);
   double *hs = stage0 + a.stage_doubles;    // mpad * HS_LD
   double *th = hs + (size_t)mpad * HS_LD;   // NC
   double *scr0 = th + NC;                   // NRG * 2 * 2 * NC * CG_LD
   uint64_t *full = reinterpret_cast<uint64_t *>(scr0 + (size_t)NRG * 4 * NC * CG_LD);
   uint64_t *empty = full + nstages;
   int *flag = reinterpret_cast<int *>(empty + nstages);
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

   if (tid == 0) {
      for (int s = 0; s < nstages; s++) {
         pbtma::mbar_init(&full[s], 1);
         pbtma::mbar_init(&empty[s], NW);
      }
      pbtma::fence_barrier_init();
   }
   {
      const double *hg = a.coef_inline ? coef.v : a.hdev;
      const double *tg = a.coef_inline ? coef.v + (size_t)m * a.nh : a.thetadev;
      for (int i = tid; i < mpad * NC; i += NCT + 32) {
         const int k = i / NC, c = i % NC;
         hs[k * HS_LD + c] = (k < m && c < a.nh) ? hg[k + (size_t)c * m] : 0.0;
      }
      for (int i = tid; i < NC; i += NCT + 32) th[i] = i < a.nh ? tg[i] : 0.0;
   }
   __syncthreads();

   const int64_t ntiles = (a.n + TR - 1) / TR;
   if (warp == NW) {
      if (lane != 0) return;
      int s = 0;
      uint32_t ph = 0;
      const uint32_t bytes = (uint32_t)(stage_sz * sizeof(double));
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
         pbtma::mbar_wait(&empty[s], ph ^ 1);
         double *dst = stage0 + (size_t)s * stage_sz;
         const int r0 = (int)(PB_TILE(tile) * TR);
         pbtma::mbar_arrive_expect_tx(&full[s], bytes);
         pbtma::tensor_g2s_2d(dst, &maps.v, r0, 0, &full[s]);
         pbtma::tensor_g2s_2d(dst + mpad * S, &maps.w, r0, 0, &full[s]);
         if (++s == nstages) s = 0, ph ^= 1;
      }
      return;
   }
   // warp -> (row group, column tile): the warps of a row group are consecutive
   const int rgw = warp / NT8, cg = warp % NT8;
#define PB_CG_CASE(C_) \
   case C_: \
      if constexpr (C_ < NT8) vwxr_cg_consumer<NT8, NRG, (C_ < NT8 ? C_ : 0)>(a, stage0, stage_sz, hs, th, scr0, full, empty, nstages, rgw, lane); \
      break;
   switch (cg) {
      PB_CG_CASE(0) PB_CG_CASE(1) PB_CG_CASE(2) PB_CG_CASE(3) PB_CG_CASE(4) PB_CG_CASE(5)
   }
#undef PB_CG_CASE

   const int nR = a.R.ptr ? a.R.ce - a.R.cb : 0;
   const int nn = (nR + (a.re - a.rb)) > 0 ? a.nh : 0;
   const int cnt = nn + a.nG * a.nG + a.nH * a.nH;
   if (cnt <= 0) return;
   constexpr int WSZ = T::NTILES * 64 + NC;
   pbtma::named_bar_sync(8, NCT);
   double *out = a.partials + (size_t)blockIdx.x * cnt;
   const int offG = nn, offH = offG + a.nG * a.nG;
   for (int e = tid; e < cnt; e += NCT) {
      int idx;
      if (e < offG) {
         idx = T::NTILES * 64 + e;
      } else if (e < offH) {
         int i = (e - offG) % a.nG, j = (e - offG) / a.nG;
         if (i > j) { const int sw = i; i = j, j = sw; }  // G is symmetric: upper tiles only
         const int ti = i >> 3, tj = j >> 3;
         idx = ((tj * (tj + 1) / 2 + ti) * 8 + (i & 7)) * 8 + (j & 7);
      } else {
         const int i = (e - offH) % a.nH, j = (e - offH) / a.nH;
         idx = ((T::NGT + (i >> 3) + NT8 * (j >> 3)) * 8 + (i & 7)) * 8 + (j & 7);
      }
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < NRG; w++) sum += stage0[(size_t)w * WSZ + idx];
      out[e] = sum;
   }
   pb_finish_device(a.fin, tid, NCT, 15, flag);
}

template <int NT8, int NRG>
int launch_vwxr_cg(pb200_ctx *ctx, const VwxrArgs &a, const VwxrMaps &maps, int grid, size_t shmem, int nstages) {
   auto kern = vwxr_cg_kernel<NT8, NRG>;
   static size_t attr_shmem = 0;
   if (shmem > attr_shmem) {
      PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
      attr_shmem = shmem;
   }
   if (getenv("PB200_DEBUG")) {
      static size_t seen = 0;
      if (seen != shmem) {
         seen = shmem;
         cudaFuncAttributes fa;
         cudaFuncGetAttributes(&fa, kern);
         fprintf(stderr, "primme_b200: vwxr_cg<%d,%d> dyn smem %zu regs %d stages %d\n", NT8, NRG, shmem, fa.numRegs, nstages);
      }
   }
   kern<<<grid, NT8 * NRG * 32 + 32, shmem, ctx->stream>>>(a, ctx->coef, maps, nstages);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return 0;
}

template <int NTH, int NG>
int launch_vwxr_wide(pb200_ctx *ctx, const VwxrArgs &a, int grid, size_t shmem, int nstages, int park_cols) {
   auto kern = vwxr_wide_kernel<NTH, NG>;
   static size_t attr_shmem = 0;
   if (shmem > attr_shmem) {
      PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
      attr_shmem = shmem;
   }
   kern<<<grid, 288, shmem, ctx->stream>>>(a, ctx->coef, nstages, park_cols);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return 0;
}

template <int NT>
int launch_vwxr(pb200_ctx *ctx, const VwxrArgs &a, int grid, size_t shmem) {
   auto kern = vwxr_kernel<NT>;
   static size_t attr_shmem = 48 * 1024;
   if (shmem > attr_shmem) {
      PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
      attr_shmem = shmem;
   }
   kern<<<grid, VT, shmem, ctx->stream>>>(a, ctx->coef);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return 0;
}

}  // namespace

// single-launch path: nh <= 64 columns of h, Gram blocks up to 38 x 38
static int vwxr_fast(pb200_ctx *ctx, int64_t n, const double *V, const double *W, int m,
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
   int cnt = nn + a.nG * a.nG + a.nH * a.nH;
   if (o->P_host && nR <= 0) return PB200_ERR_ARG;
   if (o->R2 && !o->P_host) return PB200_ERR_ARG;  // second residual destination: fused sweep only
   if (o->R2) {
      if ((((uintptr_t)o->R2) & 15) != 0 || o->ldR2 % 2 != 0) return PB200_ERR_ARG;
      a.R2 = o->R2, a.ldR2 = o->ldR2;
   }
   {
      int gb = (a.nG + 1) / 2, hb = (a.nH + 1) / 2;
      if (gb * gb + hb * hb > 768) return PB200_ERR_ARG;
   }

   // h (compacted to ld m) and theta: inside the kernel parameters when they fit, else staged
   size_t need = (size_t)m * nh + nh;
   PB_CHK(pb_ensure_small(ctx, need > (size_t)cnt ? need : (size_t)cnt));
   double *hp;
   if (need <= PB_COEF_MAX && ctx->coef_inline) {
      a.coef_inline = 1;
      hp = ctx->coef.v;
   } else {
      // the pinned staging buffer may still feed an earlier async copy
      PB_CUDA(cudaStreamSynchronize(ctx->stream));
      hp = ctx->h_pinned;
   }
   for (int c = 0; c < nh; c++)
      for (int k = 0; k < m; k++) hp[k + (size_t)c * m] = h_host[k + (size_t)c * ldh];
   for (int c = 0; c < nh; c++) hp[(size_t)m * nh + c] = theta_host ? theta_host[c] : 0.0;
   if (!a.coef_inline) {
      PB_CUDA(cudaMemcpyAsync(
            ctx->d_small, hp, need * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      a.hdev = ctx->d_small;
      a.thetadev = ctx->d_small + (size_t)m * nh;
   }

   // algorithmic bytes: V (and W) read once + every output column written once
   double ocols = 0;
   for (int t = 0; t < 3; t++)
      if (a.X[t].ptr) ocols += a.X[t].ce - a.X[t].cb;
   if (a.Wo.ptr) ocols += a.Wo.ce - a.Wo.cb;
   if (a.R.ptr) ocols += (a.R.ce - a.R.cb) * (a.R2 ? 2 : 1);
   const double abytes = 8.0 * (double)n * ((a.need_y ? 2.0 : 1.0) * m + ocols);

   // ---- wide plan: restart sweep (TMA-staged, Gram blocks) ----
   int wide_nst = 0, wide_nth = 0, wide_ng = 0, wide_park = 0;
   size_t wide_shm = 0;
   const bool cand_shape = nh <= 8 && a.nG == 0 && a.nH == 0;  // candidates sweep: LDG kernel (fastest measured)
   if (ctx->use_tma_vwxr && ctx->use_wide && !cand_shape && nh <= 48 && m > 0 && n >= 4 * 64 &&
         (((uintptr_t)V) & 15) == 0 && (((uintptr_t)W) & 15) == 0 && ld % 2 == 0) {
      bool ok = true;
      auto al = [](const pb200_cols &c) { return !c.ptr || ((((uintptr_t)c.ptr) & 15) == 0 && c.ld % 2 == 0); };
      for (int t = 0; t < 3; t++) ok = ok && al(a.X[t]);
      ok = ok && al(a.Wo) && al(a.R);
      const int cfg[][2] = {{8, 1}, {8, 2}, {12, 2}, {8, 4}, {12, 4}};
      const int park = a.nG > a.nH ? a.nG : a.nH;
      const int nb4g = (a.nG + 3) / 4, nb4h = (a.nH + 3) / 4;
      const int nblk = nb4g * (nb4g + 1) / 2 + nb4h * nb4h;
      for (int i = 0; i < 5 && ok && !wide_nst; i++) {
         const int nth = cfg[i][0], ng = cfg[i][1];
         if (nh > nth * ng || nblk > 32 * ng || park > nth * ng) continue;
         const size_t fixed = ((size_t)m * nth * ng + nth * ng + (size_t)(8 / ng) * 2 * park * 64) * sizeof(double) + 640;
         const size_t stage = (size_t)2 * m * 64 * sizeof(double);
         if (fixed + 2 * stage > 227 * 1024) continue;
         int st = (int)((227 * 1024 - fixed) / stage);
         if (st > 4) st = 4;
         wide_nst = st, wide_nth = nth, wide_ng = ng, wide_park = park, wide_shm = fixed + st * stage;
      }
   }

   // ---- main plan: tensor-map TMA + DMMA kernel (candidates with nh <= 8, restart with nh <= 32) ----
   int mma_nt8 = 0, mma_mt = 0, mma_nw = 0, mma_nst = 0;
   size_t mma_shm = 0;
   VwxrMaps maps;
   {
      const int nt8 = (nh + 7) / 8;
      const int ngh = a.nG > a.nH ? a.nG : a.nH;
      const bool wantP = o->P_host != NULL;
      bool ok = ctx->use_tma_vwxr && ctx->use_mma_vwxr && a.need_y && nt8 <= 4 && ngh <= 8 * nt8 && m > 0 && m <= 96 &&
                n >= 256 && (((uintptr_t)V) & 15) == 0 && (((uintptr_t)W) & 15) == 0 && ld % 2 == 0 &&
                (!wantP || (nt8 == 1 && ngh == 0 && m <= 72));
      if (ok) {
         const int mpad = (m + 3) & ~3;
         const int mt = wantP ? ((m + 7) / 8 <= 5 ? 5 : 9) : 0;
         const int NC = 8 * nt8;
         auto plan = [&](int nw, int *nst, size_t *shm) {
            const size_t stage = (size_t)2 * mpad * (8 * nw + 4);
            const size_t wsz = (size_t)(nt8 * (nt8 + 1) / 2 + nt8 * nt8 + (mt ? mt + 1 : 0)) * 64 + NC;
            const size_t fixed = ((size_t)mpad * (NC + 4) + NC + (size_t)nw * 2 * NC * XW_LD) * sizeof(double) + 2 * 8 * sizeof(uint64_t) + 64;
            int st = (int)((227 * 1024 - fixed) / (stage * sizeof(double)));
            if (st > 6) st = 6;
            size_t ring = (size_t)st * stage;
            if (ring < wsz * nw) ring = wsz * nw;
            *nst = st, *shm = ring * sizeof(double) + fixed;
            a.stage_doubles = (int)ring;
            return st >= 2 && *shm <= 227 * 1024;
         };
         int nw = (nt8 == 1 && mt <= 5) ? 16 : 8;
         static const int min_st16 = getenv("PB200_VWXR_MIN_STAGES16") ? atoi(getenv("PB200_VWXR_MIN_STAGES16")) : 2;
         if (nw == 16 && !(plan(16, &mma_nst, &mma_shm) && mma_nst >= min_st16)) nw = 8;
         // restart shape (3 tiles of h): optional 10 consumer warps (80-row tiles); measured equal to
         // 8 warps (397 vs 390 us at C2), so 8 warps with the deeper ring stay the default
         static const int nw_restart = getenv("PB200_VWXR_RESTART_WARPS") ? atoi(getenv("PB200_VWXR_RESTART_WARPS")) : 8;
         if (nt8 == 3 && nw == 8 && nw_restart == 10 && plan(10, &mma_nst, &mma_shm) && mma_nst >= 3) nw = 10;
         ok = plan(nw, &mma_nst, &mma_shm);
         if (ok) {
            memset(&maps, 0, sizeof(maps));
            ok = !pb_tensor_map_2d(&maps.v, V, n, m, ld, 8 * nw + 4, mpad) && !pb_tensor_map_2d(&maps.w, W, n, m, ld, 8 * nw + 4, mpad);
         }
         if (ok) mma_nt8 = nt8, mma_mt = mt, mma_nw = nw, a.mpad = mpad, a.nP = wantP ? nR : 0;
      }
      if (wantP && !mma_nt8) return PB200_ERR_ARG;  // callers ask for P only after pb200_dvwxr_can_fuse_gram()
      if (mma_nt8) cnt += (m + a.nP) * a.nP;
   }

   // ---- restart plan: columns of h split over the warps (8 < nh <= 48, no P panel) ----
   int cg_nt8 = 0, cg_nrg = 0, cg_nst = 0;
   size_t cg_shm = 0;
   {
      const int nt8 = (nh + 7) / 8;
      const int ngh = a.nG > a.nH ? a.nG : a.nH;
      static const int use_cg = getenv("PB200_VWXR_CG") ? atoi(getenv("PB200_VWXR_CG")) : 1;
      auto al = [](const pb200_cols &c) { return !c.ptr || ((((uintptr_t)c.ptr) & 15) == 0 && c.ld % 2 == 0); };
      bool ok = use_cg && ctx->use_tma_vwxr && ctx->use_mma_vwxr && a.need_y && !o->P_host && !a.R2 && nt8 >= 2 && nt8 <= 6 &&
                ngh <= 8 * nt8 && m > 0 && m <= 96 && n >= 256 && (((uintptr_t)V) & 15) == 0 && (((uintptr_t)W) & 15) == 0 &&
                ld % 2 == 0 && al(a.X[0]) && al(a.X[1]) && al(a.X[2]) && al(a.Wo) && al(a.R);
      if (ok) {
         static const int nrg_tab[7] = {0, 0, 4, 4, 3, 3, 2};
         const int nrg = nrg_tab[nt8], NC = 8 * nt8, mpad = (m + 3) & ~3;
         const size_t stage = (size_t)2 * mpad * (16 * nrg + 4);
         const size_t wsz = (size_t)(nt8 * (nt8 + 1) / 2 + nt8 * nt8) * 64 + NC;
         const size_t fixed = ((size_t)mpad * (NC + 4) + NC + (size_t)nrg * 4 * NC * CG_LD) * sizeof(double) + 2 * 8 * sizeof(uint64_t) + 64;
         int st = (int)((227 * 1024 - fixed) / (stage * sizeof(double)));
         if (st > 6) st = 6;
         size_t ring = (size_t)(st > 0 ? st : 0) * stage;
         if (ring < wsz * nrg) ring = wsz * nrg;
         const size_t shm = ring * sizeof(double) + fixed;
         ok = st >= 2 && shm <= 227 * 1024;
         VwxrMaps cmaps;
         memset(&cmaps, 0, sizeof(cmaps));
         ok = ok && !pb_tensor_map_2d(&cmaps.v, V, n, m, ld, 16 * nrg + 4, mpad) && !pb_tensor_map_2d(&cmaps.w, W, n, m, ld, 16 * nrg + 4, mpad);
         if (ok) {
            cg_nt8 = nt8, cg_nrg = nrg, cg_nst = st, cg_shm = shm;
            a.mpad = mpad, a.stage_doubles = (int)ring, a.nP = 0;
            maps = cmaps;
            mma_nt8 = 0;  // (no P panel in either plan here: cnt is unchanged)
         }
      }
   }

   int grid = 1, ppc = 1, rc = 0;
   if (cg_nt8) {
      a.rev = ctx->sweep_alternate ? (ctx->sweep_rev ^= 1) : 0;
      const int tr = 16 * cg_nrg;
      const int64_t ntiles = (n + tr - 1) / tr;
      grid = (int)(ntiles < (int64_t)ctx->num_sms ? ntiles : (int64_t)ctx->num_sms);
   } else if (mma_nt8) {
      a.rev = ctx->sweep_alternate ? (ctx->sweep_rev ^= 1) : 0;
      const int tr = 8 * mma_nw;
      const int64_t ntiles = (n + tr - 1) / tr;
      grid = (int)(ntiles < (int64_t)ctx->num_sms ? ntiles : (int64_t)ctx->num_sms);
   } else if (wide_nst >= 2) {
      const int64_t ntiles = (n + 63) / 64;
      grid = (int)(ntiles < (int64_t)ctx->num_sms ? ntiles : (int64_t)ctx->num_sms);
      ppc = 8 / wide_ng;
   } else if (n > 0) {
      const int64_t ntiles = (n + VT - 1) / VT;
      const int64_t g = (int64_t)ctx->num_sms * 4;
      grid = (int)(ntiles < g ? ntiles : g);
   }
   // partial panels: every slot is written by the kernel (unused norm slots as zeros, G below the
   // diagonal by mirroring), so no memset in front of the launch
   {
      const int r = pb_fin_prepare(ctx, grid, ppc, cnt, &a.fin);
      if (r < 0) return r;
      if (r == 1) PB_CHK(pb_ensure_partials(ctx, (size_t)grid * ppc * (cnt > 0 ? cnt : 1) + 16));
      a.partials = ctx->d_partials;
   }
   int ps = pb_prof_begin(ctx, PB_K_VWXR);
   if (cg_nt8) {
#define VC(NT8_, NRG_) \
   if (cg_nt8 == NT8_ && cg_nrg == NRG_) rc = launch_vwxr_cg<NT8_, NRG_>(ctx, a, maps, grid, cg_shm, cg_nst);
      VC(2, 4) VC(3, 4) VC(4, 3) VC(5, 3) VC(6, 2)
#undef VC
   } else if (mma_nt8) {
#define VM(NT8_, MT_, NW_) \
   if (mma_nt8 == NT8_ && mma_mt == MT_ && mma_nw == NW_) rc = launch_vwxr_mma<NT8_, MT_, NW_>(ctx, a, maps, grid, mma_shm, mma_nst);
      VM(1, 0, 16) VM(1, 5, 16) VM(1, 0, 8) VM(1, 5, 8) VM(1, 9, 8) VM(2, 0, 8) VM(3, 0, 8) VM(3, 0, 10) VM(4, 0, 8)
#undef VM
   } else if (wide_nst >= 2) {
#define VW3(NTH_, NG_) \
   if (wide_nth == NTH_ && wide_ng == NG_) rc = launch_vwxr_wide<NTH_, NG_>(ctx, a, grid, wide_shm, wide_nst, wide_park);
      VW3(8, 1) VW3(8, 2) VW3(12, 2) VW3(8, 4) VW3(12, 4)
#undef VW3
   } else {
      const int NT = nh <= 4 ? 4 : nh <= 8 ? 8 : nh <= 16 ? 16 : nh <= 24 ? 24 : nh <= 32 ? 32
                    : nh <= 40 ? 40 : nh <= 48 ? 48 : 64;
      const int ngh = a.nG > a.nH ? a.nG : a.nH;
      size_t shd = (size_t)m * NT + NT + (size_t)ngh * VT + (size_t)a.nH * VT;
      if (shd < VT) shd = VT;
      size_t shmem = shd * sizeof(double);
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
   }
   pb_prof_end(ctx, ps, abytes);
   PB_CHK(rc);
   if (cnt > 0) {
      if (a.fin.cnt > 0)
         PB_CHK(pb_collect_panel(ctx, &a.fin));
      else
         PB_CHK(pb_finish_panel(ctx, grid * ppc, cnt));
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
      if (a.nP > 0) {
         const double *pp = ph + a.nH * a.nH;
         for (int j = 0; j < a.nP; j++)
            for (int i = 0; i < m + a.nP; i++) o->P_host[i + (size_t)j * o->ldP] = pp[i + (size_t)j * (m + a.nP)];
      }
   }
   return 0;
}

// General path for wide coefficient blocks (more than 64 columns of h, or Gram blocks larger
// than the single-launch kernel holds; e.g. a restart that keeps an almost full basis):
// the products P = V*h and Q = W*h go to scratch in chunks of 32 columns through the fast kernel,
// then every output is derived from P and Q with the other fused kernels.  Rare and off the
// steady-state path (it runs once per solve for the benchmark configurations).
static int vwxr_general(pb200_ctx *ctx, int64_t n, const double *V, const double *W, int m,
      int64_t ld, const double *h_host, int ldh, int nh, const double *theta_host,
      const pb200_vwxr_out *o) {
   const int64_t lds = (n + 15) / 16 * 16 > 0 ? (n + 15) / 16 * 16 : 16;
   const bool need_y = (o->Wo.ptr && o->Wo.ce > o->Wo.cb) || (o->R.ptr && o->R.ce > o->R.cb) ||
                       (o->rnorms_host && o->re > o->rb) || (o->H_host && o->nH > 0);
   PB_CHK(pb_ensure_scratch(ctx, sizeof(double) * (size_t)lds * nh * (need_y ? 2 : 1)));
   double *P = (double *)ctx->d_scratch, *Q = need_y ? P + (size_t)lds * nh : NULL;
   for (int c0 = 0; c0 < nh; c0 += 32) {
      const int nc = nh - c0 < 32 ? nh - c0 : 32;
      pb200_vwxr_out oc;
      memset(&oc, 0, sizeof(oc));
      oc.X[0].ptr = P + (size_t)lds * c0, oc.X[0].ld = lds, oc.X[0].cb = 0, oc.X[0].ce = nc;
      if (need_y) oc.Wo.ptr = Q + (size_t)lds * c0, oc.Wo.ld = lds, oc.Wo.cb = 0, oc.Wo.ce = nc;
      PB_CHK(vwxr_fast(ctx, n, V, W, m, ld, h_host + (size_t)ldh * c0, ldh, nc,
            theta_host ? theta_host + c0 : NULL, &oc));
   }
   // Gram blocks G = P(:,0:nG)' P(:,0:nG), H = P(:,0:nH)' Q(:,0:nH): panels of <= 8 columns
   if (o->G_host && o->nG > 0)
      for (int c0 = 0; c0 < o->nG; c0 += 8) {
         const int nc = o->nG - c0 < 8 ? o->nG - c0 : 8;
         PB_CHK(pb200_dortho_sweep(ctx, n, NULL, 0, 0, P, o->nG, lds, P + (size_t)lds * c0, nc, lds,
               NULL, 0, NULL, 0, 0, o->G_host + (size_t)o->ldG * c0, o->ldG));
      }
   if (o->H_host && o->nH > 0)
      for (int c0 = 0; c0 < o->nH; c0 += 8) {
         const int nc = o->nH - c0 < 8 ? o->nH - c0 : 8;
         PB_CHK(pb200_dortho_sweep(ctx, n, NULL, 0, 0, P, o->nH, lds, Q + (size_t)lds * c0, nc, lds,
               NULL, 0, NULL, 0, 0, o->H_host + (size_t)o->ldH * c0, o->ldH));
      }
   // column-range outputs (may alias V / W: every product has been formed by now)
   for (int t = 0; t < 3; t++)
      if (o->X[t].ptr && o->X[t].ce > o->X[t].cb)
         PB_CHK(pb200_copy_d2d(ctx, P + (size_t)lds * o->X[t].cb, lds, o->X[t].ptr, o->X[t].ld, n,
               o->X[t].ce - o->X[t].cb, 8));
   if (o->Wo.ptr && o->Wo.ce > o->Wo.cb)
      PB_CHK(pb200_copy_d2d(ctx, Q + (size_t)lds * o->Wo.cb, lds, o->Wo.ptr, o->Wo.ld, n,
            o->Wo.ce - o->Wo.cb, 8));
   // residuals: Q_j <- Q_j - theta_j P_j in scratch, squared norms from the same pass
   std::vector<double> n2(nh, -1.0);
   auto residual_cols = [&](int cb, int ce) -> int {
      for (int c0 = cb; c0 < ce; c0 += 8) {
         int c1 = c0 + 8 < ce ? c0 + 8 : ce, lo = c0;
         while (lo < c1 && n2[lo] >= 0.0) lo++;  // already done (overlapping ranges)
         if (lo >= c1) continue;
         PB_CHK(pb200_dresidual_inplace(ctx, n, theta_host + lo, P + (size_t)lds * lo, lds,
               Q + (size_t)lds * lo, lds, c1 - lo, &n2[lo]));
      }
      return 0;
   };
   if (o->R.ptr && o->R.ce > o->R.cb) {
      PB_CHK(residual_cols(o->R.cb, o->R.ce));
      PB_CHK(pb200_copy_d2d(ctx, Q + (size_t)lds * o->R.cb, lds, o->R.ptr, o->R.ld, n, o->R.ce - o->R.cb, 8));
      if (o->Rnorms_host)
         for (int c = o->R.cb; c < o->R.ce; c++) o->Rnorms_host[c - o->R.cb] = sqrt(n2[c]);
   }
   if (o->rnorms_host && o->re > o->rb) {
      PB_CHK(residual_cols(o->rb, o->re));
      for (int c = o->rb; c < o->re; c++) o->rnorms_host[c - o->rb] = sqrt(n2[c]);
   }
   return 0;
}

// whether a sweep of this shape can also deliver P = [V R]^T R (out->P_host)
extern "C" int pb200_dvwxr_can_fuse_gram(pb200_ctx *ctx, int64_t n, const double *V, const double *W, int m,
      int64_t ld, int nh, const pb200_vwxr_out *o) {
   const int nG = o->G_host ? o->nG : 0, nH = o->H_host ? o->nH : 0;
   const int nR = (o->R.ptr && o->R.ce > o->R.cb) ? o->R.ce - o->R.cb : 0;
   if (!ctx->use_tma_vwxr || !ctx->use_mma_vwxr || !ctx->fuse_gram) return 0;
   if (nh <= 0 || nh > 8 || nG > 0 || nH > 0 || nR <= 0 || m <= 0 || m > 72 || n < 256) return 0;
   if ((((uintptr_t)V) & 15) != 0 || (((uintptr_t)W) & 15) != 0 || ld % 2 != 0) return 0;
   // the stage ring must hold >= 2 stages of 64-row tiles
   const int mpad = (m + 3) & ~3;
   const size_t stage = (size_t)2 * mpad * 68 * sizeof(double);
   return 2 * stage + 40 * 1024 <= 227 * 1024;
}

extern "C" int pb200_dvwxr(pb200_ctx *ctx, int64_t n, const double *V, const double *W, int m,
      int64_t ld, const double *h_host, int ldh, int nh, const double *theta_host,
      const pb200_vwxr_out *o) {
   if (nh <= 0 || m < 0) return 0;
   const int nG = o->G_host ? o->nG : 0, nH = o->H_host ? o->nH : 0;
   const int gb = (nG + 1) / 2, hb = (nH + 1) / 2;
   if (o->P_host && !pb200_dvwxr_can_fuse_gram(ctx, n, V, W, m, ld, nh, o)) return PB200_ERR_ARG;
   if (nh <= 64 && gb * gb + hb * hb <= 768)
      return vwxr_fast(ctx, n, V, W, m, ld, h_host, ldh, nh, theta_host, o);
   return vwxr_general(ctx, n, V, W, m, ld, h_host, ldh, nh, theta_host, o);
}

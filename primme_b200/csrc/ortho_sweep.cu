// ortho_sweep.cu -- K2/K3/K4: fused block-orthogonalisation row sweep (fp64).
//
//    X <- (X - [Q V] C) Y        (optional update;  reference src/eigs/ortho.c:1017-1038)
//    P  = [Q V (X)]^H X          (optional Gram panel; ortho.c:1043-1059, and
//                                 update_projection.c:99-102 when X is the new W block)
//
// HBM-bound (0.5-2 flop/B): the basis is streamed exactly once per sweep.  Three kernels, one entry point:
//   ortho_sweep_mma_exact_kernel<NMT,16,UPD,XX>   the steady-state instances (no locked vectors, 3-5 tiles of
//            8 basis columns): persistent CTAs, one producer thread streams 128-row tiles of [V X] with 2-D
//            tensor-map TMA into a ring of stages, 16 consumer warps own 8 rows each and do the update and
//            the Gram with DMMA m8n8k4, all MMAs unconditional, accumulators in registers across tiles;
//   ortho_sweep_mma_kernel<MT,NW>                 the same pipeline for any (q, mv, b) that fits the stages
//            (locked vectors, wide C5 panels: 64-row tiles and 8 warps when the ring would be < 3 deep);
//   ortho_sweep_kernel<BT,CPW>                    LDG fallback (unaligned columns, n < 256): thread <-> row
//            update, warp <-> CPW basis columns Gram.
// Every CTA leaves one partial panel; the CTAs finish it themselves in two fixed-order levels
// (pb_finish_device) and the last one delivers (value, sequence number) pairs to mapped host memory, or
// all-reduces over peer memory first => bitwise reproducible panels, no second launch.
// Consecutive sweeps walk the row tiles in opposite directions (PB_TILE) for L2 reuse.
#include "pb200_internal.cuh"
#include "tma_pipe.cuh"
#include <string.h>
#include <stdlib.h>
#include <time.h>

namespace {

constexpr int TILE = 256;  // rows per tile == threads per CTA
constexpr int NWARP = TILE / 32;

// tile visited at step `t` of a CTA's walk: every other sweep runs back to front (a.rev) so that it starts on
// the rows the previous sweep left in the 126 MB L2
#define PB_TILE(t) (a.rev ? ntiles - 1 - (t) : (t))
struct SweepArgs {
   int rev;
   const double *Q;
   const double *V;
   double *X;
   int64_t n, ldq, ldv, ldx;
   int q, mv, b;
   int do_update, has_Y, do_gram, xx;
   int bt;              // leading dimension of Y / column count of the packed C (b padded to 1,2,4,8)
   int stage_doubles;   // size of the stage ring of the MMA kernel (>= the end-of-kernel panel scratch)
   int qpad, xcol0;     // MMA kernel: virtual column of the first V / first X column
   unsigned long long *trace;  // PB200_TRACE: [cta][8] %globaltimer stamps
   const double *Cdev;  // (q+mv) x BT, column stride = (q+mv)
   const double *Ydev;  // BT x BT, column stride BT
   int coef_inline;     // 1: C and Y travel in `coef` below (kernel parameter space), no H2D copy
   double *partials;    // [gridDim.x][(k + xx*b) * b]
   PbFin fin;           // in-kernel panel finish (fin.cnt == 0: the host launches the reduction)
};

__device__ __forceinline__ const double *col_ptr(const SweepArgs &a, int j) {
   return j < a.q ? a.Q + (size_t)j * a.ldq : a.V + (size_t)(j - a.q) * a.ldv;
}

template <int BT, int CPW>
__global__ void __launch_bounds__(TILE) ortho_sweep_kernel(SweepArgs a, const __grid_constant__ PbCoef coef) {
   extern __shared__ double smem[];
   const int k = a.q + a.mv;
   double *Cs = smem;                 // k * BT
   double *Ys = Cs + (size_t)k * BT;  // BT * BT
   double *xs = Ys + BT * BT;         // BT * TILE  (new X tile, [c][row])
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

   if (a.do_update) {
      for (int i = tid; i < k * BT; i += TILE) {
         int j = i % k, c = i / k;
         Cs[j * BT + c] = (c < a.b) ? (a.coef_inline ? coef.v : a.Cdev)[j + (size_t)c * k] : 0.0;
      }
      for (int i = tid; i < BT * BT; i += TILE) {
         int r = i % BT, c = i / BT;
         double y = (r == c) ? 1.0 : 0.0;
         if (a.has_Y) y = (r < a.b && c < a.b) ? (a.coef_inline ? coef.v + (size_t)k * BT : a.Ydev)[r + c * BT] : 0.0;
         Ys[r * BT + c] = y;
      }
   }
   __syncthreads();

   // Gram accumulators: this warp owns basis columns [warp*CPW, warp*CPW+CPW)
   double acc[CPW][BT];
   double accx[BT];
#pragma unroll
   for (int j = 0; j < CPW; j++)
#pragma unroll
      for (int c = 0; c < BT; c++) acc[j][c] = 0.0;
#pragma unroll
   for (int c = 0; c < BT; c++) accx[c] = 0.0;

   const double *cols[CPW];
#pragma unroll
   for (int j = 0; j < CPW; j++) {
      int jj = warp * CPW + j;
      cols[j] = jj < k ? col_ptr(a, jj) : nullptr;
   }

   const int64_t ntiles = (a.n + TILE - 1) / TILE;
   for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t r0 = tile * TILE;
      // ---------------- phase 1: per-row update, park new X rows in shared memory -------
      {
         const int64_t r = r0 + tid;
         double x[BT];
#pragma unroll
         for (int c = 0; c < BT; c++)
            x[c] = (r < a.n && c < a.b) ? a.X[r + (size_t)c * a.ldx] : 0.0;
         if (a.do_update && r < a.n) {
            int j = 0;
            for (; j + 4 <= k; j += 4) {  // 4 independent loads in flight per step
               double v0 = col_ptr(a, j)[r], v1 = col_ptr(a, j + 1)[r];
               double v2 = col_ptr(a, j + 2)[r], v3 = col_ptr(a, j + 3)[r];
#pragma unroll
               for (int c = 0; c < BT; c++) {
                  x[c] -= v0 * Cs[(j + 0) * BT + c];
                  x[c] -= v1 * Cs[(j + 1) * BT + c];
                  x[c] -= v2 * Cs[(j + 2) * BT + c];
                  x[c] -= v3 * Cs[(j + 3) * BT + c];
               }
            }
            for (; j < k; j++) {
               double v0 = col_ptr(a, j)[r];
#pragma unroll
               for (int c = 0; c < BT; c++) x[c] -= v0 * Cs[j * BT + c];
            }
            if (a.has_Y) {
               double y[BT];
#pragma unroll
               for (int c = 0; c < BT; c++) {
                  double s = 0.0;
#pragma unroll
                  for (int cc = 0; cc < BT; cc++) s += x[cc] * Ys[cc * BT + c];
                  y[c] = s;
               }
#pragma unroll
               for (int c = 0; c < BT; c++) x[c] = y[c];
            }
#pragma unroll
            for (int c = 0; c < BT; c++)
               if (c < a.b) a.X[r + (size_t)c * a.ldx] = x[c];
         }
#pragma unroll
         for (int c = 0; c < BT; c++) xs[c * TILE + tid] = x[c];
      }
      __syncthreads();
      // ---------------- phase 2: Gram accumulation, columns split over warps ------------
      if (a.do_gram) {
#pragma unroll 2
         for (int i = 0; i < TILE / 32; i++) {
            const int row = i * 32 + lane;
            const int64_t r = r0 + row;
            double xv[BT];
#pragma unroll
            for (int c = 0; c < BT; c++) xv[c] = xs[c * TILE + row];
            double av[CPW];
#pragma unroll
            for (int j = 0; j < CPW; j++) av[j] = (cols[j] && r < a.n) ? cols[j][r] : 0.0;
#pragma unroll
            for (int j = 0; j < CPW; j++)
#pragma unroll
               for (int c = 0; c < BT; c++) acc[j][c] += av[j] * xv[c];
            if (a.xx && warp < a.b) {
               double xw = xs[warp * TILE + row];
#pragma unroll
               for (int c = 0; c < BT; c++) accx[c] += xw * xv[c];
            }
         }
      }
      __syncthreads();
   }

   if (!a.do_gram) return;
   // ---------------- epilogue: lane reduction, one partial panel per CTA ----------------
   const int rows = k + (a.xx ? a.b : 0);
   double *out = a.partials + (size_t)blockIdx.x * rows * a.b;
#pragma unroll
   for (int j = 0; j < CPW; j++) {
#pragma unroll
      for (int c = 0; c < BT; c++) {
         double v = acc[j][c];
#pragma unroll
         for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
         int jj = warp * CPW + j;
         if (lane == 0 && jj < k && c < a.b) out[jj + (size_t)c * rows] = v;
      }
   }
   if (a.xx) {
#pragma unroll
      for (int c = 0; c < BT; c++) {
         double v = accx[c];
#pragma unroll
         for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
         if (lane == 0 && warp < a.b && c < a.b) out[k + warp + (size_t)c * rows] = v;
      }
   }
   pb_finish_device(a.fin, tid, TILE, 15, reinterpret_cast<int *>(xs + BT * TILE));
}


// ------------------------------------------------------------------------------------------
// Main kernel: TMA-staged tiles + fp64 tensor-core (DMMA m8n8k4) contractions, warp-local rows.
//
// A producer warp streams TR-row tiles of [Q V X] into a ring of shared-memory stages, one bulk
// copy (cp.async.bulk, SASS UBLKCP) per column segment, completion tracked by mbarriers.  Each of
// the NW consumer warps owns 8 rows of the tile and does everything for them without any
// CTA-wide barrier:
//   update  D(8 x 8)   = X - A(8 x k) * C(k x 8)       k/4 DMMAs, A fragments straight from the
//                                                      stage (column stride TR+4 doubles: the
//                                                      fragment loads are bank-conflict free)
//           D          = D * Y                         1-2 DMMAs after a quad shuffle
//           new rows -> global and back into the stage's X columns
//   Gram    P(kc x 8) += A(8 x kc)^T * X(8 x 8)        kc/8 accumulator tiles x 2 DMMAs, kept in
//                                                      registers across all tiles of the CTA
// The warps only meet at the end: per-warp panels are summed in warp order through shared
// memory (one partial panel per CTA), then the in-kernel finish delivers the reduced panel.
// Block widths b < 8 are zero-padded to the 8 columns of the MMA shape (the kernel is HBM-bound:
// the fp64 pipe runs at 20-45 % of its 37 TFLOP/s).  Requires 16-byte aligned columns.
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(d0), "+d"(d1)
                : "d"(a), "d"(b));
}

__device__ __forceinline__ void pb_stamp(unsigned long long *trace, int slot) {
   if (trace) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      trace[(size_t)blockIdx.x * 8 + slot] = t;
   }
}

constexpr int CS_LD = 12;  // row stride (doubles) of the coefficient blocks in shared memory:
                           // B-fragment loads (row = lane%4, column = lane/4) hit 32 distinct banks

struct SweepMaps {
   CUtensorMap q, v, x;  // boxes of (TR+4) rows x (qpad | xcol0-qpad | 8) columns
};

// Stage layout ("virtual columns", each S = TR+4 doubles): [Q | pad to a multiple of 4][V | pad so
// that X starts at a multiple of 8 = xcol0][X padded to 8 columns].  The pads are columns past the
// end of the respective tensor map, which the TMA unit zero-fills, so every region starts
// 128-byte aligned (4 columns = 33 x 128 B), column addresses are linear in the virtual index and
// no fragment load needs a bounds check.
template <int MT, int NW>
__global__ void __launch_bounds__(NW * 32 + 32) ortho_sweep_mma_kernel(SweepArgs a, const __grid_constant__ PbCoef coef,
      const __grid_constant__ SweepMaps maps, int nstages) {
   constexpr int TR = 8 * NW;     // rows per tile
   constexpr int S = TR + 4;      // column stride of a staged tile
   constexpr int NCT = NW * 32;   // consumer threads
   extern __shared__ __align__(128) unsigned char smraw[];
   const int k = a.q + a.mv;
   const int qpad = a.qpad, xcol0 = a.xcol0;
   const int stage_sz = (xcol0 + 8) * S;
   double *stage0 = reinterpret_cast<double *>(smraw);             // nstages * stage_sz
   double *Cs = stage0 + a.stage_doubles;                          // xcol0 * CS_LD   (holds -C)
   double *Ys = Cs + (size_t)xcol0 * CS_LD;                        // 8 * CS_LD
   double *xw0 = Ys + 8 * CS_LD;                                   // NW * 8 * CS_LD: updated rows, per warp
   uint64_t *full = reinterpret_cast<uint64_t *>(xw0 + NW * 8 * CS_LD);
   uint64_t *empty = full + nstages;
   int *flag = reinterpret_cast<int *>(empty + nstages);
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

   if (tid == 0) {
      pb_stamp(a.trace, 0);
      for (int s = 0; s < nstages; s++) {
         pbtma::mbar_init(&full[s], 1);
         pbtma::mbar_init(&empty[s], NW);
      }
      pbtma::fence_barrier_init();
   }
   if (a.do_update) {
      const double *cg = a.coef_inline ? coef.v : a.Cdev;
      const double *yg = a.coef_inline ? coef.v + (size_t)k * a.bt : a.Ydev;
      for (int i = tid; i < NW * 8 * CS_LD; i += NCT + 32) xw0[i] = 0.0;
      for (int i = tid; i < xcol0 * 8; i += NCT + 32) {
         const int vc = i >> 3, c = i & 7;
         int j = -1;  // real column of [Q V] behind virtual column vc
         if (vc < qpad) {
            if (vc < a.q) j = vc;
         } else if (vc - qpad < a.mv)
            j = a.q + vc - qpad;
         Cs[vc * CS_LD + c] = (j >= 0 && c < a.b) ? -cg[j + (size_t)c * k] : 0.0;
      }
      for (int i = tid; i < 64; i += NCT + 32) {
         const int r = i >> 3, c = i & 7;
         Ys[r * CS_LD + c] = (a.has_Y && r < a.b && c < a.b) ? yg[r + c * a.bt] : 0.0;
      }
   }
   __syncthreads();

   const int64_t ntiles = (a.n + TR - 1) / TR;  // the last tile may be partial
   if (tid == 0) pb_stamp(a.trace, 1);
   if (warp == NW) {
      // ------------------------------ producer: one elected lane ------------------------------
      // one tensor-tile copy per operand and tile; rows past n and the pad columns are zero-filled
      // by the TMA unit, the 4 extra rows of a box make the fragment loads bank-conflict free
      if (lane != 0) return;
      int s = 0;
      uint32_t ph = 0;
      const uint32_t bytes = (uint32_t)(stage_sz * sizeof(double));
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
         pbtma::mbar_wait(&empty[s], ph ^ 1);
         double *dst = stage0 + (size_t)s * stage_sz;
         const int r0 = (int)(PB_TILE(tile) * TR);
         pbtma::mbar_arrive_expect_tx(&full[s], bytes);
         if (qpad > 0) pbtma::tensor_g2s_2d(dst, &maps.q, r0, 0, &full[s]);
         if (xcol0 > qpad) pbtma::tensor_g2s_2d(dst + qpad * S, &maps.v, r0, 0, &full[s]);
         pbtma::tensor_g2s_2d(dst + xcol0 * S, &maps.x, r0, 0, &full[s]);
         if (++s == nstages) s = 0, ph ^= 1;
      }
      return;
   }

   // ------------------------------ consumer warps ------------------------------
   const int g = lane >> 2, t = lane & 3;  // fragment coordinates
   const int r0w = warp * 8;               // first row of this warp inside the tile
   const int nmtv = a.do_gram ? xcol0 >> 3 : 0;   // accumulator tiles over [Q V]
   const int nks2 = xcol0 >> 3;                   // pairs of k-steps of the update
   double acc[MT][2], accx[2];
#pragma unroll
   for (int mt = 0; mt < MT; mt++) acc[mt][0] = 0.0, acc[mt][1] = 0.0;
   accx[0] = accx[1] = 0.0;
   double *xw = xw0 + warp * 8 * CS_LD;
   const int offu = r0w + t * S + g;   // update A fragment: column t of a k-step, row g
   const int offg = r0w + g * S + t;   // Gram A fragment: column g of a tile, row t of a k-step
   const double *cpu = Cs + t * CS_LD + g;

   int s = 0;
   uint32_t ph = 0;
   for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      pbtma::mbar_wait(&full[s], ph);
      if (tid == 0 && tile == blockIdx.x) pb_stamp(a.trace, 2);
      const double *st = stage0 + (size_t)s * stage_sz;
      const double *xb = st + xcol0 * S + offg;  // Gram X operand: X(row t of a k-step, column g)
      if (a.do_update) {
         const double *sx = st + xcol0 * S + r0w + g;
         double d0 = sx[(2 * t) * S], d1 = sx[(2 * t + 1) * S];
         double e0 = 0.0, e1 = 0.0;
         const double *pa = st + offu;
         const double *pc = cpu;
#pragma unroll 2
         for (int i = 0; i < nks2; i++) {
            const double a0 = pa[0], a1 = pa[4 * S];
            dmma884(d0, d1, a0, pc[0]);
            dmma884(e0, e1, a1, pc[4 * CS_LD]);
            pa += 8 * S, pc += 8 * CS_LD;
         }
         d0 += e0, d1 += e1;
         if (a.has_Y) {
            // D <- D * Y: the accumulator tile becomes the A operand after a shuffle inside each
            // quad (A wants column t of row g, the accumulators hold columns 2t, 2t+1)
            double y0 = 0.0, y1 = 0.0;
            {
               const int src = (lane & ~3) | (t >> 1);
               const double v0 = __shfl_sync(0xffffffffu, d0, src), v1 = __shfl_sync(0xffffffffu, d1, src);
               dmma884(y0, y1, (t & 1) ? v1 : v0, Ys[t * CS_LD + g]);
            }
            if (a.b > 4) {
               const int src = (lane & ~3) | 2 | (t >> 1);
               const double v0 = __shfl_sync(0xffffffffu, d0, src), v1 = __shfl_sync(0xffffffffu, d1, src);
               dmma884(y0, y1, (t & 1) ? v1 : v0, Ys[(4 + t) * CS_LD + g]);
            }
            d0 = y0, d1 = y1;
         }
         // the new rows go to global memory and to this warp's scratch ([column][row], stride
         // CS_LD): the Gram below takes its X operand from there (the stage itself is only ever
         // written by the TMA unit, so no proxy fence is needed before it is refilled)
         const int64_t r = PB_TILE(tile) * TR + r0w + g;
         if (2 * t < a.b && r < a.n) a.X[r + (size_t)(2 * t) * a.ldx] = d0;
         if (2 * t + 1 < a.b && r < a.n) a.X[r + (size_t)(2 * t + 1) * a.ldx] = d1;
         xw[(2 * t) * CS_LD + g] = d0;
         xw[(2 * t + 1) * CS_LD + g] = d1;
         __syncwarp();
         xb = xw + g * CS_LD + t;
      }
      if (a.do_gram) {
         const double *pg = st + offg;
#pragma unroll
         for (int ks = 0; ks < 2; ks++) {
            const double bf = xb[4 * ks];
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
               if (mt < nmtv) dmma884(acc[mt][0], acc[mt][1], pg[mt * 8 * S + 4 * ks], bf);
            // X'X block: A(column g, row t) and B(row t, column g) are the same element
            if (a.xx) dmma884(accx[0], accx[1], bf, bf);
         }
      }
      __syncwarp();
      if (lane == 0) pbtma::mbar_arrive(&empty[s]);
      if (++s == nstages) s = 0, ph ^= 1;
   }

   if (tid == 0) pb_stamp(a.trace, 3);
   if (!a.do_gram) return;
   // ---- per-warp panels -> one partial panel per CTA (warp order), through the stage memory ----
   pbtma::named_bar_sync(1, NCT);  // every warp is done with the stages (all issued tiles consumed)
   double *red = stage0;           // [NW][(nmtv + 1) * 8][8], virtual rows
   const int nv = (nmtv + 1) * 8;
#pragma unroll
   for (int mt = 0; mt < MT; mt++)
      if (mt < nmtv)
         *reinterpret_cast<double2 *>(red + ((size_t)warp * nv + mt * 8 + g) * 8 + 2 * t) =
               make_double2(acc[mt][0], acc[mt][1]);
   *reinterpret_cast<double2 *>(red + ((size_t)warp * nv + nmtv * 8 + g) * 8 + 2 * t) = make_double2(accx[0], accx[1]);
   pbtma::named_bar_sync(1, NCT);
   const int rows_out = k + (a.xx ? a.b : 0);
   double *out = a.partials + (size_t)blockIdx.x * rows_out * a.b;
   for (int e = tid; e < rows_out * a.b; e += NCT) {
      const int j = e % rows_out, c = e / rows_out;
      const int i = j < a.q ? j : j < k ? qpad + (j - a.q) : nmtv * 8 + (j - k);  // virtual row
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < NW; w++) sum += red[((size_t)w * nv + i) * 8 + c];
      out[e] = sum;
   }
   if (tid == 0) pb_stamp(a.trace, 4);
   pb_finish_device(a.fin, tid, NCT, 15, flag);
   if (tid == 0) pb_stamp(a.trace, 5);
}

// ------------------------------------------------------------------------------------------
// Specialised instances of the main kernel for the steady-state shapes (no locked vectors, the
// stage holds exactly NMT 8-column tiles of V, Gram panel always wanted): every DMMA is
// unconditional (a predicated mma.sync costs a branch and a WARPSYNC), the update coefficients live
// in registers as B fragments (shared-memory instructions compete with DMMA for the same issue
// queue: +3 cycles per LDS on top of 16 per DMMA, scripts/micro/dmma_issue.cu), the transposing
// scratch is swizzled so that its stores and loads are conflict free.
template <int NMT, int NW, bool UPD, bool XX>
__global__ void __launch_bounds__(NW * 32 + 32) ortho_sweep_mma_exact_kernel(SweepArgs a, const __grid_constant__ PbCoef coef,
      const __grid_constant__ SweepMaps maps, int nstages) {
   constexpr int TR = 8 * NW, S = TR + 4, NCT = NW * 32;
   constexpr int XC0 = 8 * NMT;              // first X column of a stage
   constexpr int STAGE = (XC0 + 8) * S;
   extern __shared__ __align__(128) unsigned char smraw[];
   const int k = a.mv;  // q == 0 here
   double *stage0 = reinterpret_cast<double *>(smraw);
   double *Cs = stage0 + a.stage_doubles;                 // XC0 * CS_LD   (holds -C)
   double *Ys = Cs + (size_t)XC0 * CS_LD;                 // 8 * CS_LD
   double *xw0 = Ys + 8 * CS_LD;                          // NW * 8 * CS_LD
   uint64_t *full = reinterpret_cast<uint64_t *>(xw0 + NW * 8 * CS_LD);
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
   if (UPD) {
      const double *cg = a.coef_inline ? coef.v : a.Cdev;
      const double *yg = a.coef_inline ? coef.v + (size_t)k * a.bt : a.Ydev;
      for (int i = tid; i < XC0 * 8; i += NCT + 32) {
         const int vc = i >> 3, c = i & 7;
         Cs[vc * CS_LD + c] = (vc < k && c < a.b) ? -cg[vc + (size_t)c * k] : 0.0;
      }
      for (int i = tid; i < 64; i += NCT + 32) {
         const int r = i >> 3, c = i & 7;
         Ys[r * CS_LD + c] = (r < a.b && c < a.b) ? (a.has_Y ? yg[r + c * a.bt] : (r == c ? 1.0 : 0.0)) : 0.0;
      }
   }
   __syncthreads();

   const int64_t ntiles = (a.n + TR - 1) / TR;
   if (warp == NW) {
      if (lane != 0) return;
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
         pbtma::mbar_wait(&empty[s], ph ^ 1);
         double *dst = stage0 + (size_t)s * STAGE;
         const int r0 = (int)(PB_TILE(tile) * TR);
         pbtma::mbar_arrive_expect_tx(&full[s], (uint32_t)(STAGE * sizeof(double)));
         pbtma::tensor_g2s_2d(dst, &maps.v, r0, 0, &full[s]);
         pbtma::tensor_g2s_2d(dst + XC0 * S, &maps.x, r0, 0, &full[s]);
         if (++s == nstages) s = 0, ph ^= 1;
      }
      return;
   }

   const int g = lane >> 2, t = lane & 3;
   const int r0w = warp * 8;
   double acc[NMT][2], accx[2] = {0.0, 0.0};
#pragma unroll
   for (int mt = 0; mt < NMT; mt++) acc[mt][0] = acc[mt][1] = 0.0;
   double *xw = xw0 + warp * 8 * CS_LD;
   const int offu = r0w + t * S + g;   // update A fragment: column t of a k-step, row g
   const int offg = r0w + g * S + t;   // Gram A fragment: column g of a tile, row t of a k-step
   // B fragments of -C for every k-step and of Y, loaded once
   double cfr[UPD ? 2 * NMT : 1], yfr0 = 0.0, yfr1 = 0.0;
   if (UPD) {
#pragma unroll
      for (int i = 0; i < 2 * NMT; i++) cfr[i] = Cs[(4 * i + t) * CS_LD + g];
      yfr0 = Ys[t * CS_LD + g], yfr1 = Ys[(4 + t) * CS_LD + g];
   }
   // scratch [column][row]: rows of columns 4..7 XORed with 4 (conflict-free stores and loads)
   const int xs0 = (2 * t) * CS_LD + (g ^ ((2 * t) & 4)), xs1 = xs0 + CS_LD;
   const int xl0 = g * CS_LD + (t ^ (g & 4)), xl1 = g * CS_LD + ((t + 4) ^ (g & 4));
   const bool wide = a.b > 4;
   const bool st0 = 2 * t < a.b, st1 = 2 * t + 1 < a.b;
   double *gx0 = a.X + (size_t)(2 * t) * a.ldx + r0w + g, *gx1 = gx0 + a.ldx;

   int s = 0;
   uint32_t ph = 0;
   for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      pbtma::mbar_wait(&full[s], ph);
      const double *st = stage0 + (size_t)s * STAGE;
      double bf0, bf1;  // Gram X operand: X(row t (+4) of the warp's 8 rows, column g)
      if (UPD) {
         const double *sx = st + XC0 * S + r0w + g;
         double d0 = sx[(2 * t) * S], d1 = sx[(2 * t + 1) * S], e0 = 0.0, e1 = 0.0;
         const double *pa = st + offu;
#pragma unroll
         for (int i = 0; i < NMT; i++) {
            dmma884(d0, d1, pa[(8 * i) * S], cfr[2 * i]);
            dmma884(e0, e1, pa[(8 * i + 4) * S], cfr[2 * i + 1]);
         }
         d0 += e0, d1 += e1;
         // D <- D * Y: the accumulator tile becomes the A operand after a shuffle inside each quad
         double y0 = 0.0, y1 = 0.0;
         {
            const int src = (lane & ~3) | (t >> 1);
            const double v0 = __shfl_sync(0xffffffffu, d0, src), v1 = __shfl_sync(0xffffffffu, d1, src);
            dmma884(y0, y1, (t & 1) ? v1 : v0, yfr0);
         }
         if (wide) {
            const int src = (lane & ~3) | 2 | (t >> 1);
            const double v0 = __shfl_sync(0xffffffffu, d0, src), v1 = __shfl_sync(0xffffffffu, d1, src);
            dmma884(y0, y1, (t & 1) ? v1 : v0, yfr1);
         }
         const int64_t roff = PB_TILE(tile) * TR;
         if (roff + r0w + g < a.n) {
            if (st0) gx0[roff] = y0;
            if (st1) gx1[roff] = y1;
         }
         xw[xs0] = y0, xw[xs1] = y1;
         __syncwarp();
         bf0 = xw[xl0], bf1 = xw[xl1];
      } else {
         const double *xb = st + XC0 * S + offg;
         bf0 = xb[0], bf1 = xb[4];
      }
      {
         const double *pg = st + offg;
#pragma unroll
         for (int mt = 0; mt < NMT; mt++) dmma884(acc[mt][0], acc[mt][1], pg[mt * 8 * S], bf0);
         if (XX) dmma884(accx[0], accx[1], bf0, bf0);
#pragma unroll
         for (int mt = 0; mt < NMT; mt++) dmma884(acc[mt][0], acc[mt][1], pg[mt * 8 * S + 4], bf1);
         if (XX) dmma884(accx[0], accx[1], bf1, bf1);
      }
      __syncwarp();
      if (lane == 0) pbtma::mbar_arrive(&empty[s]);
      if (++s == nstages) s = 0, ph ^= 1;
   }

   // ---- per-warp panels -> one partial panel per CTA (warp order), through the stage memory ----
   pbtma::named_bar_sync(1, NCT);
   double *red = stage0;  // [NW][(NMT + 1) * 8][8]
   constexpr int nv = (NMT + 1) * 8;
#pragma unroll
   for (int mt = 0; mt < NMT; mt++)
      *reinterpret_cast<double2 *>(red + ((size_t)warp * nv + mt * 8 + g) * 8 + 2 * t) = make_double2(acc[mt][0], acc[mt][1]);
   *reinterpret_cast<double2 *>(red + ((size_t)warp * nv + NMT * 8 + g) * 8 + 2 * t) = make_double2(accx[0], accx[1]);
   pbtma::named_bar_sync(1, NCT);
   const int rows_out = k + (XX ? a.b : 0);
   double *out = a.partials + (size_t)blockIdx.x * rows_out * a.b;
   for (int e = tid; e < rows_out * a.b; e += NCT) {
      const int j = e % rows_out, c = e / rows_out;
      const int i = j < k ? j : NMT * 8 + (j - k);
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < NW; w++) sum += red[((size_t)w * nv + i) * 8 + c];
      out[e] = sum;
   }
   pb_finish_device(a.fin, tid, NCT, 15, flag);
}

// partial-panel storage + in-kernel finish for a launch of `grid` CTAs (cnt = 0: no panel)
static int sweep_panel_setup(pb200_ctx *ctx, SweepArgs &a, int grid, int cnt) {
   if (cnt <= 0) return 0;
   const int r = pb_fin_prepare(ctx, grid, 1, cnt, &a.fin);
   if (r < 0) return r;
   if (r == 1) PB_CHK(pb_ensure_partials(ctx, (size_t)grid * cnt));
   a.partials = ctx->d_partials;
   return 0;
}

// shared memory of the MMA kernel besides the stage ring
static size_t mma_fixed_smem(int xcol0) {
   return ((size_t)xcol0 * CS_LD + 8 * CS_LD + 16 * 8 * CS_LD) * sizeof(double) + 2 * 8 * sizeof(uint64_t) + 64;
}

template <int MT, int NW>
int launch_sweep_mma(pb200_ctx *ctx, SweepArgs &a, const SweepMaps &maps, int grid, size_t shmem, int nstages, int cnt) {
   auto kern = ortho_sweep_mma_kernel<MT, NW>;
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
         fprintf(stderr, "primme_b200: ortho_sweep_mma<%d,%d> dyn smem %zu regs %d stages %d -> %d CTA/SM\n",
               MT, NW, shmem, fa.numRegs, nstages, occ);
      }
   }
   static const int tracing = getenv("PB200_TRACE") ? atoi(getenv("PB200_TRACE")) : 0;
   struct timespec tA, tB, tC;
   if (tracing > 1) clock_gettime(CLOCK_MONOTONIC, &tA);
   PB_CHK(sweep_panel_setup(ctx, a, grid, cnt));
   if (tracing > 1) clock_gettime(CLOCK_MONOTONIC, &tB);
   kern<<<grid, NW * 32 + 32, shmem, ctx->stream>>>(a, ctx->coef, maps, nstages);
   if (tracing > 1) {
      clock_gettime(CLOCK_MONOTONIC, &tC);
      fprintf(stderr, "TRACE2 setup %.2f us, <<<>>> %.2f us\n", (tB.tv_sec - tA.tv_sec) * 1e6 + (tB.tv_nsec - tA.tv_nsec) * 1e-3,
            (tC.tv_sec - tB.tv_sec) * 1e6 + (tC.tv_nsec - tB.tv_nsec) * 1e-3);
   }
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return 0;
}


template <int NMT, int NW, bool UPD, bool XX>
int launch_sweep_exact(pb200_ctx *ctx, SweepArgs &a, const SweepMaps &maps, int grid, size_t shmem, int nstages, int cnt) {
   auto kern = ortho_sweep_mma_exact_kernel<NMT, NW, UPD, XX>;
   static size_t attr_shmem = 0;
   if (shmem > attr_shmem) {
      PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
      attr_shmem = shmem;
   }
   PB_CHK(sweep_panel_setup(ctx, a, grid, cnt));
   kern<<<grid, NW * 32 + 32, shmem, ctx->stream>>>(a, ctx->coef, maps, nstages);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return 0;
}

template <int NMT>
int dispatch_exact(pb200_ctx *ctx, SweepArgs &a, const SweepMaps &maps, int grid, size_t shmem, int nstages, int cnt) {
   if (a.do_update) return launch_sweep_exact<NMT, 16, true, true>(ctx, a, maps, grid, shmem, nstages, cnt);
   if (a.xx) return launch_sweep_exact<NMT, 16, false, true>(ctx, a, maps, grid, shmem, nstages, cnt);
   return launch_sweep_exact<NMT, 16, false, false>(ctx, a, maps, grid, shmem, nstages, cnt);
}

template <int BT, int CPW>
int launch_sweep(pb200_ctx *ctx, const SweepArgs &a, int grid, size_t shmem) {
   auto kern = ortho_sweep_kernel<BT, CPW>;
   static size_t attr_shmem = 48 * 1024;
   if (shmem > attr_shmem) {
      PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
      attr_shmem = shmem;
   }
   kern<<<grid, TILE, shmem, ctx->stream>>>(a, ctx->coef);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return 0;
}

template <int BT>
int dispatch_cpw(pb200_ctx *ctx, const SweepArgs &a, int cpw, int grid, size_t shmem) {
   switch (cpw) {
   case 1: return launch_sweep<BT, 1>(ctx, a, grid, shmem);
   case 2: return launch_sweep<BT, 2>(ctx, a, grid, shmem);
   case 3: return launch_sweep<BT, 3>(ctx, a, grid, shmem);
   case 4: return launch_sweep<BT, 4>(ctx, a, grid, shmem);
   case 5: return launch_sweep<BT, 5>(ctx, a, grid, shmem);
   case 6: return launch_sweep<BT, 6>(ctx, a, grid, shmem);
   case 7:
   case 8: return launch_sweep<BT, 8>(ctx, a, grid, shmem);
   case 9:
   case 10: return launch_sweep<BT, 10>(ctx, a, grid, shmem);
   case 11:
   case 12: return launch_sweep<BT, 12>(ctx, a, grid, shmem);
   default: return PB200_ERR_ARG;
   }
}

}  // namespace

// Maximum number of basis columns one launch handles (NWARP * 12); more are chunked.
static const int KMAX = NWARP * 12;

static int sweep_once(pb200_ctx *ctx, int64_t n, const double *Q, int q, int64_t ldq,
      const double *V, int mv, int64_t ldv, double *X, int b, int64_t ldx, const double *C_host,
      int ldc, const double *Y_host, int ldy, int xx, double *P_host, int ldp) {
   const int k = q + mv;
   const int BT = b <= 1 ? 1 : b <= 2 ? 2 : b <= 4 ? 4 : 8;
   SweepArgs a;
   memset(&a, 0, sizeof(a));
   a.Q = Q, a.V = V, a.X = X, a.n = n, a.ldq = ldq, a.ldv = ldv, a.ldx = ldx;
   a.q = q, a.mv = mv, a.b = b;
   a.do_update = C_host != NULL || Y_host != NULL;
   a.has_Y = Y_host != NULL;
   a.do_gram = P_host != NULL;
   a.xx = xx ? 1 : 0;
   if (!a.do_update && !a.do_gram) return 0;

   // C (k x b, compacted to ld k) and Y (b x b, ld BT): inside the kernel parameters when they fit
   // (no H2D copy in front of the launch), else staged through the pinned buffer
   if (a.do_update) {
      size_t need = (size_t)k * BT + BT * BT;
      double *hp;
      if (need <= PB_COEF_MAX && ctx->coef_inline) {
         a.coef_inline = 1;
         hp = ctx->coef.v;
      } else {
         PB_CHK(pb_ensure_small(ctx, need));
         // the pinned staging buffer may still feed an earlier async copy
         PB_CUDA(cudaStreamSynchronize(ctx->stream));
         hp = ctx->h_pinned;
      }
      for (int c = 0; c < b; c++)
         for (int j = 0; j < k; j++) hp[j + (size_t)c * k] = C_host ? C_host[j + (size_t)c * ldc] : 0.0;
      for (size_t i = (size_t)k * b; i < (size_t)k * BT; i++) hp[i] = 0.0;
      double *hy = hp + (size_t)k * BT;
      for (int i = 0; i < BT * BT; i++) hy[i] = 0.0;
      if (Y_host)
         for (int c = 0; c < b; c++)
            for (int r = 0; r < b; r++) hy[r + c * BT] = Y_host[r + (size_t)c * ldy];
      if (!a.coef_inline) {
         PB_CUDA(cudaMemcpyAsync(ctx->d_small, hp, need * sizeof(double), cudaMemcpyHostToDevice,
               ctx->stream));
         a.Cdev = ctx->d_small;
         a.Ydev = ctx->d_small + (size_t)k * BT;
      }
   }

   const int rows = k + (a.xx ? b : 0);
   int cpw = (k + NWARP - 1) / NWARP;
   if (cpw < 1) cpw = 1;
   // algorithmic bytes: basis read once, X read (+written when updated)  
   const double abytes = 8.0 * (double)n * (k + b * (a.do_update ? 2 : 1));

   // ---- main (TMA + DMMA) kernel eligibility: 16-byte aligned columns (tensor maps), enough
   // rows, a ring of >= 2 stages in shared memory ----
   a.bt = BT;
   // virtual column layout of a stage (see the kernel): Q padded to 4, V padded so that X starts
   // at a multiple of 8, X padded to 8
   const int qpad = mv > 0 ? (q + 3) & ~3 : (q + 7) & ~7;
   const int xcol0 = mv > 0 ? (qpad + mv + 7) & ~7 : qpad;
   a.qpad = qpad, a.xcol0 = xcol0;
   bool mma_ok = ctx->use_tma && n >= 256 && xcol0 <= 104;
   int nstages = 0, nw = 16, ctas_per_sm = 1;
   const size_t fixed_sm = mma_fixed_smem(xcol0);
   const size_t smem_cap = 227 * 1024;
   size_t stage_d = 0;
   if (mma_ok) {
      // 128-row tiles and 16 consumer warps; wide panels (C5: 64 + 8 columns) take 64-row tiles
      // and 8 consumer warps so that the ring stays >= 3 deep.  One CTA per SM.
      stage_d = (size_t)(xcol0 + 8) * (128 + 4);
      nstages = (int)((smem_cap - fixed_sm) / (stage_d * sizeof(double)));
      if (nstages < 3) {
         nw = 8;
         stage_d = (size_t)(xcol0 + 8) * (64 + 4);
         nstages = (int)((smem_cap - fixed_sm) / (stage_d * sizeof(double)));
      }
      if (nstages > 8) nstages = 8;
      if (nw == 16 && ctx->ortho_2cta == 2) {
         // two CTAs of 16 consumer warps per SM, 2-3 stages each (twice the resident warps)
         const int ns = (int)((113 * 1024 - fixed_sm) / (stage_d * sizeof(double)));
         if (ns >= 2) nstages = ns > 3 ? 3 : ns, ctas_per_sm = 2;
      } else if (nw == 16 && ctx->ortho_2cta) {
         // two CTAs of 8 consumer warps (64-row tiles) per SM: the CTAs drift apart, so the MMA
         // bursts of one overlap the shared-memory / store phases of the other
         const size_t sd = (size_t)(xcol0 + 8) * (64 + 4);
         const int ns = (int)((112 * 1024 - fixed_sm) / (sd * sizeof(double)));
         if (ns >= 3) nw = 8, stage_d = sd, nstages = ns > 4 ? 4 : ns, ctas_per_sm = 2;
      }
      if (nstages < 2) mma_ok = false;
   }
   SweepMaps maps;
   if (mma_ok) {
      const int box = 8 * nw + 4;
      memset(&maps, 0, sizeof(maps));
      if (qpad > 0 && pb_tensor_map_2d(&maps.q, Q, n, q, ldq, box, qpad)) mma_ok = false;
      if (mma_ok && xcol0 > qpad && pb_tensor_map_2d(&maps.v, V, n, mv, ldv, box, xcol0 - qpad)) mma_ok = false;
      if (mma_ok && pb_tensor_map_2d(&maps.x, X, n, b, ldx, box, 8)) mma_ok = false;
   }

   int grid = 0, rc = 0;
   const int cnt = a.do_gram ? rows * b : 0;
   static unsigned long long *d_trace = NULL;
   static const int tracing = getenv("PB200_TRACE") ? atoi(getenv("PB200_TRACE")) : 0;
   struct timespec th0, th1, th2, th3;
   if (tracing) {
      if (!d_trace) cudaMalloc((void **)&d_trace, 8 * 1024 * sizeof(unsigned long long));
      cudaMemsetAsync(d_trace, 0, 8 * 1024 * sizeof(unsigned long long), ctx->stream);
      cudaStreamSynchronize(ctx->stream);
      a.trace = mma_ok ? d_trace : NULL;
      clock_gettime(CLOCK_MONOTONIC, &th0);
   }
   int ps = pb_prof_begin(ctx, PB_K_ORTHO);
   if (mma_ok) {
      a.rev = ctx->sweep_alternate ? (ctx->sweep_rev ^= 1) : 0;
      const int tr = 8 * nw;
      const int64_t ntiles = (n + tr - 1) / tr;  // a partial last tile is zero-filled by the TMA unit
      grid = (int)(ntiles < (int64_t)ctx->num_sms * ctas_per_sm ? ntiles : (int64_t)ctx->num_sms * ctas_per_sm);
      const int nmtv = xcol0 / 8;
      size_t ring = (size_t)nstages * stage_d;
      if (ring < (size_t)nw * (nmtv + 1) * 64) ring = (size_t)nw * (nmtv + 1) * 64;
      a.stage_doubles = (int)ring;
      const size_t shmem = ring * sizeof(double) + fixed_sm;
      if (tracing) clock_gettime(CLOCK_MONOTONIC, &th1);
#define PB_MMA(MT_, NW_) rc = launch_sweep_mma<MT_, NW_>(ctx, a, maps, grid, shmem, nstages, cnt)
      // steady-state shapes (no locked vectors, Gram wanted, update always with xx): specialised
      const bool exact = ctx->ortho_exact && nw == 16 && q == 0 && mv > 0 && a.do_gram && !a.trace &&
                         (!a.do_update || a.xx) && nmtv >= 3 && nmtv <= 5;
      if (exact) {
         rc = nmtv == 3 ? dispatch_exact<3>(ctx, a, maps, grid, shmem, nstages, cnt)
            : nmtv == 4 ? dispatch_exact<4>(ctx, a, maps, grid, shmem, nstages, cnt)
                        : dispatch_exact<5>(ctx, a, maps, grid, shmem, nstages, cnt);
      } else if (nw == 16) {
         if (nmtv <= 5) PB_MMA(5, 16);
         else if (nmtv <= 9) PB_MMA(9, 16);
         else PB_MMA(13, 16);
      } else {
         if (nmtv <= 5) PB_MMA(5, 8);
         else if (nmtv <= 9) PB_MMA(9, 8);
         else PB_MMA(13, 8);
      }
#undef PB_MMA
   } else {
      const int64_t ntiles = (n + TILE - 1) / TILE;
      grid = (int)(ntiles < (int64_t)ctx->num_sms * 3 ? ntiles : (int64_t)ctx->num_sms * 3);
      if (grid < 1) grid = 1;
      rc = sweep_panel_setup(ctx, a, grid, cnt);
      size_t shmem = ((size_t)k * BT + BT * BT + (size_t)BT * TILE) * sizeof(double) + 16;
      if (!rc) switch (BT) {
      case 1: rc = dispatch_cpw<1>(ctx, a, cpw, grid, shmem); break;
      case 2: rc = dispatch_cpw<2>(ctx, a, cpw, grid, shmem); break;
      case 4: rc = dispatch_cpw<4>(ctx, a, cpw, grid, shmem); break;
      default: rc = dispatch_cpw<8>(ctx, a, cpw, grid, shmem); break;
      }
   }
   pb_prof_end(ctx, ps, abytes);
   PB_CHK(rc);
   if (tracing) clock_gettime(CLOCK_MONOTONIC, &th2);
   if (a.do_gram) {
      if (a.fin.cnt > 0)
         PB_CHK(pb_collect_panel(ctx, &a.fin));
      else
         PB_CHK(pb_finish_panel(ctx, grid, cnt));
      if (tracing && a.trace) {
         clock_gettime(CLOCK_MONOTONIC, &th3);
         cudaStreamSynchronize(ctx->stream);
         static unsigned long long ht[8 * 1024];
         cudaMemcpy(ht, d_trace, sizeof(unsigned long long) * 8 * grid, cudaMemcpyDeviceToHost);
         unsigned long long mn[6], mx[6];
         for (int s = 0; s < 6; s++) mn[s] = ~0ULL, mx[s] = 0;
         for (int c = 0; c < grid; c++)
            for (int s = 0; s < 6; s++) {
               const unsigned long long t = ht[c * 8 + s];
               if (t == 0) continue;
               if (t < mn[s]) mn[s] = t;
               if (t > mx[s]) mx[s] = t;
            }
         auto us = [](const timespec &x, const timespec &y) { return (y.tv_sec - x.tv_sec) * 1e6 + (y.tv_nsec - x.tv_nsec) * 1e-3; };
         const double t0 = (double)mn[0];
         fprintf(stderr, "TRACE n=%lld upd=%d host: prep %.1f launch %.1f wait %.1f total %.1f us | gpu (us from first CTA entry) "
                         "entry<=%.1f setup<=%.1f first_tile %.1f..%.1f loop_done %.1f..%.1f partial %.1f..%.1f finish %.1f..%.1f\n",
               (long long)n, a.do_update, us(th0, th1), us(th1, th2), us(th2, th3), us(th0, th3),
               (mx[0] - t0) * 1e-3, (mx[1] - t0) * 1e-3, (mn[2] - t0) * 1e-3, (mx[2] - t0) * 1e-3, (mn[3] - t0) * 1e-3,
               (mx[3] - t0) * 1e-3, (mn[4] - t0) * 1e-3, (mx[4] - t0) * 1e-3, (mn[5] - t0) * 1e-3, (mx[5] - t0) * 1e-3);
      }
      for (int c = 0; c < b; c++)
         for (int j = 0; j < rows; j++) P_host[j + (size_t)c * ldp] = ctx->h_pinned[j + (size_t)c * rows];
   }
   return 0;
}

extern "C" int pb200_dortho_sweep(pb200_ctx *ctx, int64_t n, const double *Q, int q, int64_t ldq,
      const double *V, int mv, int64_t ldv, double *X, int b, int64_t ldx, const double *C_host,
      int ldc, const double *Y_host, int ldy, int xx, double *P_host, int ldp) {
   if (b <= 0) return 0;
   if (b > 8 || q < 0 || mv < 0) return PB200_ERR_ARG;
   if (n <= 0) {
      // empty local part (a rank may own no rows): the panel is all zeros
      if (P_host) {
         int rows = q + mv + (xx ? b : 0);
         for (int c = 0; c < b; c++)
            for (int j = 0; j < rows; j++) P_host[j + (size_t)c * ldp] = 0.0;
         if (ctx->nranks > 1) {
            /* still take part in the collective */
            PB_CHK(pb_ensure_small(ctx, (size_t)rows * b));
            const int zr = pb_fin_contribute_zeros(ctx, rows * b);
            if (zr < 0) return zr;
            if (zr == 1) {
               PB_CUDA(cudaMemsetAsync(ctx->d_panel, 0, sizeof(double) * rows * b, ctx->stream));
               PB_CHK(pb_nccl_allreduce_dev(ctx, ctx->d_panel, rows * b));
               PB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_panel, sizeof(double) * rows * b,
                     cudaMemcpyDeviceToHost, ctx->stream));
               PB_CUDA(cudaStreamSynchronize(ctx->stream));
            }
            for (int c = 0; c < b; c++)
               for (int j = 0; j < rows; j++)
                  P_host[j + (size_t)c * ldp] = ctx->h_pinned[j + (size_t)c * rows];
         }
      }
      return 0;
   }
   const int k = q + mv;
   if (k <= KMAX)
      return sweep_once(ctx, n, Q, q, ldq, V, mv, ldv, X, b, ldx, C_host, ldc, Y_host, ldy, xx,
            P_host, ldp);

   // More basis columns than one launch covers: (1) apply the update chunk by chunk (Y only
   // with the last chunk), (2) compute the Gram panel chunk by chunk.
   const bool upd = C_host != NULL || Y_host != NULL;
   if (upd) {
      for (int j0 = 0; j0 < k; j0 += KMAX) {
         int j1 = j0 + KMAX < k ? j0 + KMAX : k;
         int qa = j0 < q ? (j1 < q ? j1 : q) - j0 : 0;
         int va0 = j0 > q ? j0 - q : 0;
         int va = j1 > q ? (j1 - q) - va0 : 0;
         bool last = j1 == k;
         PB_CHK(sweep_once(ctx, n, Q + (size_t)j0 * ldq * (qa > 0), qa, ldq,
               V + (size_t)va0 * ldv, va, ldv, X, b, ldx, C_host ? C_host + j0 : NULL, ldc,
               last ? Y_host : NULL, ldy, 0, NULL, 0));
         if (!C_host && !last) continue;
      }
   }
   if (P_host) {
      for (int j0 = 0; j0 < k; j0 += KMAX) {
         int j1 = j0 + KMAX < k ? j0 + KMAX : k;
         int qa = j0 < q ? (j1 < q ? j1 : q) - j0 : 0;
         int va0 = j0 > q ? j0 - q : 0;
         int va = j1 > q ? (j1 - q) - va0 : 0;
         bool last = j1 == k;
         PB_CHK(sweep_once(ctx, n, Q + (size_t)j0 * ldq * (qa > 0), qa, ldq,
               V + (size_t)va0 * ldv, va, ldv, X, b, ldx, NULL, 0, NULL, 0, last ? xx : 0,
               P_host + j0, ldp));
      }
   }
   return 0;
}

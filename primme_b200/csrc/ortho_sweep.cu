// ortho_sweep.cu -- K2/K3/K4: fused block-orthogonalisation row sweep (fp64).
//
//    X <- (X - [Q V] C) Y        (optional update;  reference src/eigs/ortho.c:1017-1038)
//    P  = [Q V (X)]^H X          (optional Gram panel; ortho.c:1043-1059, and
//                                 update_projection.c:99-102 when X is the new W block)
//
// HBM-bound (0.5-2 flop/B): the basis is streamed exactly once per sweep.  Work decomposition
// per CTA (256 threads, persistent over 256-row tiles):
//   phase 1 (update): thread <-> row.  Coalesced column reads of [Q V], C and Y broadcast from
//            shared memory, new X row written back and parked in shared memory.
//   phase 2 (Gram):   warp <-> CPW basis columns, lane <-> rows of the tile.  k*b accumulators
//            are spread over the CTA (CPW*BT per thread) and live in registers across all
//            tiles; the tile of [Q V] touched in phase 1 is re-read through L1/L2, not HBM.
//   epilogue: warp-shuffle reduction, one partial panel per CTA, fixed-order second stage
//            (pb_finish_panel) => bitwise reproducible panels.
#include "pb200_internal.cuh"
#include "tma_pipe.cuh"
#include <string.h>

namespace {

constexpr int TILE = 256;  // rows per tile == threads per CTA
constexpr int NWARP = TILE / 32;

struct SweepArgs {
   const double *Q;
   const double *V;
   double *X;
   int64_t n, ldq, ldv, ldx;
   int q, mv, b;
   int do_update, has_Y, do_gram, xx;
   const double *Cdev;  // (q+mv) x BT, column stride = (q+mv)
   const double *Ydev;  // BT x BT, column stride BT
   int coef_inline;     // 1: C and Y travel in `coef` below (kernel parameter space), no H2D copy
   double *partials;    // [gridDim.x][(k + xx*b) * b]
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
}


// ------------------------------------------------------------------------------------------
// v2: TMA-staged variant.  A dedicated producer warp streams 128-row tiles of [Q V X] into a
// ring of shared-memory stages with one bulk copy (cp.async.bulk, SASS UBLKCP) per column
// segment, completion tracked by mbarriers; 8 consumer warps do the update and the Gram from
// shared memory.  Memory-level parallelism no longer depends on registers/occupancy: up to
// 4 tiles (~200 KB) are in flight per SM.  Requires 16-byte aligned columns (even leading
// dimensions); a partial last tile is completed with plain stores by the producer warp.
constexpr int TR = 128;          // rows per tile
constexpr int NCW = 8;           // consumer warps
constexpr int NCT = NCW * 32;    // consumer threads

template <int BT, int CPW>
__global__ void __launch_bounds__(NCT + 32) ortho_sweep_tma_kernel(SweepArgs a, const __grid_constant__ PbCoef coef, int nstages) {
   extern __shared__ __align__(128) unsigned char smraw[];
   const int k = a.q + a.mv;
   const int kc = k + a.b;
   double *stage0 = reinterpret_cast<double *>(smraw);
   double *Cs = stage0 + (size_t)nstages * kc * TR;  // k * BT
   double *Ys = Cs + (size_t)k * BT;                 // BT * BT
   double *xs = Ys + BT * BT;                        // BT * TR (updated X tile)
   double *exch = xs + BT * TR;                      // 2 * TR * BT (partial updates)
   uint64_t *full = reinterpret_cast<uint64_t *>(exch + 2 * TR * BT);
   uint64_t *empty = full + nstages;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

   if (tid == 0) {
      for (int s = 0; s < nstages; s++) {
         pbtma::mbar_init(&full[s], 1);
         pbtma::mbar_init(&empty[s], NCW);
      }
      pbtma::fence_barrier_init();
   }
   if (a.do_update) {
      for (int i = tid; i < k * BT; i += NCT + 32) {
         int j = i % k, c = i / k;
         Cs[j * BT + c] = (c < a.b) ? (a.coef_inline ? coef.v : a.Cdev)[j + (size_t)c * k] : 0.0;
      }
      for (int i = tid; i < BT * BT; i += NCT + 32) {
         int r = i % BT, c = i / BT;
         double y = (r == c) ? 1.0 : 0.0;
         if (a.has_Y) y = (r < a.b && c < a.b) ? (a.coef_inline ? coef.v + (size_t)k * BT : a.Ydev)[r + c * BT] : 0.0;
         Ys[r * BT + c] = y;
      }
   }
   __syncthreads();

   const int64_t ntiles = (a.n + TR - 1) / TR;  // the last tile may be partial
   if (warp == NCW) {
      // ------------------------------ producer warp ------------------------------
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
         if (lane == 0) pbtma::mbar_wait(&empty[s], ph ^ 1);
         __syncwarp();
         double *dst = stage0 + (size_t)s * kc * TR;
         const int64_t r0 = tile * TR;
         const int rows = (int)((a.n - r0) < TR ? (a.n - r0) : TR);
         const int rows_even = rows & ~1;
         if (rows < TR) {
            // partial tile: bulk copies take the even part, the odd last row and the zero
            // padding are written with plain stores (padded rows then contribute nothing)
            for (int c = lane; c < kc; c += 32) {
               const double *src = c < k ? col_ptr(a, c) + r0 : a.X + (size_t)(c - k) * a.ldx + r0;
               double *d = dst + (size_t)c * TR;
               for (int rr = rows_even; rr < TR; rr++) d[rr] = rr < rows ? src[rr] : 0.0;
            }
            __syncwarp();
         }
         if (lane == 0)
            pbtma::mbar_arrive_expect_tx(&full[s], (uint32_t)(kc * rows_even * sizeof(double)));
         __syncwarp();
         if (rows_even > 0)
            for (int c = lane; c < kc; c += 32) {
               const double *src = c < k ? col_ptr(a, c) + r0 : a.X + (size_t)(c - k) * a.ldx + r0;
               pbtma::bulk_g2s(dst + (size_t)c * TR, src, rows_even * sizeof(double), &full[s]);
            }
         if (++s == nstages) s = 0, ph ^= 1;
      }
      return;
   }

   // ------------------------------ consumer warps ------------------------------
   double acc[CPW][BT];
   double accx[BT];
#pragma unroll
   for (int j = 0; j < CPW; j++)
#pragma unroll
      for (int c = 0; c < BT; c++) acc[j][c] = 0.0;
#pragma unroll
   for (int c = 0; c < BT; c++) accx[c] = 0.0;

   const int row1 = tid & (TR - 1), half = tid >> 7;  // phase-1 mapping: 2 threads per row
   const int kh = (k + 1) / 2, j0 = half * kh, j1 = (j0 + kh < k) ? j0 + kh : k;

   int s = 0;
   uint32_t ph = 0;
   for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      pbtma::mbar_wait(&full[s], ph);
      const double *st = stage0 + (size_t)s * kc * TR;
      const double *xsrc = st + (size_t)k * TR;  // X columns of this tile, [c][row]
      if (a.do_update) {
         // phase 1: each half of the CTA applies half of the basis columns to its row
         double t[BT];
#pragma unroll
         for (int c = 0; c < BT; c++) t[c] = 0.0;
         int j = j0;
         for (; j + 2 <= j1; j += 2) {
            double v0 = st[(size_t)j * TR + row1], v1 = st[(size_t)(j + 1) * TR + row1];
#pragma unroll
            for (int c = 0; c < BT; c++) t[c] += v0 * Cs[j * BT + c] + v1 * Cs[(j + 1) * BT + c];
         }
         if (j < j1) {
            double v0 = st[(size_t)j * TR + row1];
#pragma unroll
            for (int c = 0; c < BT; c++) t[c] += v0 * Cs[j * BT + c];
         }
#pragma unroll
         for (int c = 0; c < BT; c++) exch[(size_t)(half * TR + row1) * BT + c] = t[c];
         pbtma::named_bar_sync(1, NCT);
         if (half == 0) {
            double x[BT];
#pragma unroll
            for (int c = 0; c < BT; c++)
               x[c] = (c < a.b ? xsrc[(size_t)c * TR + row1] : 0.0) - exch[(size_t)row1 * BT + c] -
                      exch[(size_t)(TR + row1) * BT + c];
            if (a.has_Y) {
               double y[BT];
#pragma unroll
               for (int c = 0; c < BT; c++) {
                  double sum = 0.0;
#pragma unroll
                  for (int cc = 0; cc < BT; cc++) sum += x[cc] * Ys[cc * BT + c];
                  y[c] = sum;
               }
#pragma unroll
               for (int c = 0; c < BT; c++) x[c] = y[c];
            }
            const int64_t r = tile * TR + row1;
#pragma unroll
            for (int c = 0; c < BT; c++) {
               if (c < a.b && r < a.n) a.X[r + (size_t)c * a.ldx] = x[c];
               xs[(size_t)c * TR + row1] = x[c];
            }
         }
         pbtma::named_bar_sync(1, NCT);
         xsrc = xs;
      }
      if (a.do_gram) {
#pragma unroll
         for (int i = 0; i < TR / 32; i++) {
            const int row = i * 32 + lane;
            double xv[BT];
#pragma unroll
            for (int c = 0; c < BT; c++) xv[c] = (c < a.b) ? xsrc[(size_t)c * TR + row] : 0.0;
#pragma unroll
            for (int j = 0; j < CPW; j++) {
               const int jj = warp * CPW + j;
               const double av = jj < k ? st[(size_t)jj * TR + row] : 0.0;
#pragma unroll
               for (int c = 0; c < BT; c++) acc[j][c] += av * xv[c];
            }
            if (a.xx && warp < a.b) {
               const double xw = xsrc[(size_t)warp * TR + row];
#pragma unroll
               for (int c = 0; c < BT; c++) accx[c] += xw * xv[c];
            }
         }
      }
      __syncwarp();
      if (lane == 0) pbtma::mbar_arrive(&empty[s]);  // this warp is done with the stage
      if (++s == nstages) s = 0, ph ^= 1;
   }

   if (!a.do_gram) return;
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
}

template <int BT, int CPW>
int launch_sweep_tma(pb200_ctx *ctx, const SweepArgs &a, int &grid, size_t shmem, int nstages) {
   auto kern = ortho_sweep_tma_kernel<BT, CPW>;
   // persistent grid = resident CTAs only (a second wave would serialise half of the tiles);
   // attribute + occupancy are queried once per (instantiation, shared-memory size)
   static size_t cached_shmem = 0;
   static int cached_occ = 0;
   if (cached_shmem != shmem || cached_occ == 0) {
      PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
      int o = 1;
      PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, NCT + 32, shmem));
      cached_occ = o < 1 ? 1 : o;
      cached_shmem = shmem;
      if (getenv("PB200_DEBUG")) {
         cudaFuncAttributes fa;
         cudaFuncGetAttributes(&fa, kern);
         fprintf(stderr, "primme_b200: ortho_sweep_tma<%d,%d> dyn smem %zu static %zu regs %d stages %d -> %d CTA/SM\n",
               BT, CPW, shmem, fa.sharedSizeBytes, fa.numRegs, nstages, cached_occ);
      }
   }
   const int occ = cached_occ;
   if (grid > occ * ctx->num_sms) grid = occ * ctx->num_sms;
   kern<<<grid, NCT + 32, shmem, ctx->stream>>>(a, ctx->coef, nstages);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return 0;
}

template <int BT>
int dispatch_cpw_tma(pb200_ctx *ctx, const SweepArgs &a, int cpw, int &grid, size_t shmem, int nst) {
   switch (cpw) {
   case 1: return launch_sweep_tma<BT, 1>(ctx, a, grid, shmem, nst);
   case 2: return launch_sweep_tma<BT, 2>(ctx, a, grid, shmem, nst);
   case 3: return launch_sweep_tma<BT, 3>(ctx, a, grid, shmem, nst);
   case 4: return launch_sweep_tma<BT, 4>(ctx, a, grid, shmem, nst);
   case 5: return launch_sweep_tma<BT, 5>(ctx, a, grid, shmem, nst);
   case 6: return launch_sweep_tma<BT, 6>(ctx, a, grid, shmem, nst);
   case 7:
   case 8: return launch_sweep_tma<BT, 8>(ctx, a, grid, shmem, nst);
   case 9:
   case 10: return launch_sweep_tma<BT, 10>(ctx, a, grid, shmem, nst);
   case 11:
   case 12: return launch_sweep_tma<BT, 12>(ctx, a, grid, shmem, nst);
   default: return PB200_ERR_ARG;
   }
}


// ------------------------------------------------------------------------------------------
// v3: warp-specialised TMA pipeline.  Roles inside a CTA of 9 warps:
//   warp 8      producer: one bulk copy per column segment into the stage ring (as v2)
//   warps 0-1   update (only when C/Y are given): thread <-> two consecutive rows, 16-byte
//               shared-memory loads, X <- (X - [Q V] C) Y, new rows to global and to a double-
//               buffered shared tile
//   others      Gram: warp <-> CPW basis columns, lane <-> row pairs, 16-byte loads
// The roles are chained with mbarriers only (no CTA-wide barrier): the update of tile t+1
// overlaps the Gram of tile t, and the bulk copies of tiles t+2.. are in flight meanwhile.
template <int BT, int CPW>
__global__ void __launch_bounds__(NCT + 32) ortho_sweep_ws_kernel(SweepArgs a, const __grid_constant__ PbCoef coef, int nstages) {
   extern __shared__ __align__(128) unsigned char smraw[];
   const int k = a.q + a.mv;
   const int kc = k + a.b;
   double *stage0 = reinterpret_cast<double *>(smraw);
   double *xs = stage0 + (size_t)nstages * kc * TR;  // 2 * BT * TR, 16-byte aligned (double2 access)
   double *Cs = xs + 2 * BT * TR;                    // k * BT
   double *Ys = Cs + (size_t)k * BT;                 // BT * BT
   uint64_t *full = reinterpret_cast<uint64_t *>(Ys + BT * BT);
   uint64_t *empty = full + nstages;
   uint64_t *xfull = empty + nstages;  // [2]
   uint64_t *xempty = xfull + 2;       // [2]
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int nupd = a.do_update ? 2 : 0;                 // update warps
   const int ngw = a.do_gram ? NCW - nupd : 0;           // Gram warps

   if (tid == 0) {
      for (int s = 0; s < nstages; s++) {
         pbtma::mbar_init(&full[s], 1);
         pbtma::mbar_init(&empty[s], nupd + ngw);
      }
      for (int i = 0; i < 2; i++) {
         pbtma::mbar_init(&xfull[i], 2);
         pbtma::mbar_init(&xempty[i], ngw > 0 ? ngw : 1);
      }
      pbtma::fence_barrier_init();
   }
   if (a.do_update) {
      for (int i = tid; i < k * BT; i += NCT + 32) {
         int j = i % k, c = i / k;
         Cs[j * BT + c] = (c < a.b) ? (a.coef_inline ? coef.v : a.Cdev)[j + (size_t)c * k] : 0.0;
      }
      for (int i = tid; i < BT * BT; i += NCT + 32) {
         int r = i % BT, c = i / BT;
         double y = (r == c) ? 1.0 : 0.0;
         if (a.has_Y) y = (r < a.b && c < a.b) ? (a.coef_inline ? coef.v + (size_t)k * BT : a.Ydev)[r + c * BT] : 0.0;
         Ys[r * BT + c] = y;
      }
   }
   __syncthreads();

   const int64_t ntiles = (a.n + TR - 1) / TR;  // the last tile may be partial
   if (warp == NCW) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
         if (lane == 0) pbtma::mbar_wait(&empty[s], ph ^ 1);
         __syncwarp();
         double *dst = stage0 + (size_t)s * kc * TR;
         const int64_t r0 = tile * TR;
         const int rows = (int)((a.n - r0) < TR ? (a.n - r0) : TR);
         const int rows_even = rows & ~1;
         if (rows < TR) {
            // partial tile: bulk copies take the even part, the odd last row and the zero
            // padding are written with plain stores (padded rows then contribute nothing)
            for (int c = lane; c < kc; c += 32) {
               const double *src = c < k ? col_ptr(a, c) + r0 : a.X + (size_t)(c - k) * a.ldx + r0;
               double *d = dst + (size_t)c * TR;
               for (int rr = rows_even; rr < TR; rr++) d[rr] = rr < rows ? src[rr] : 0.0;
            }
            __syncwarp();
         }
         if (lane == 0)
            pbtma::mbar_arrive_expect_tx(&full[s], (uint32_t)(kc * rows_even * sizeof(double)));
         __syncwarp();
         if (rows_even > 0)
            for (int c = lane; c < kc; c += 32) {
               const double *src = c < k ? col_ptr(a, c) + r0 : a.X + (size_t)(c - k) * a.ldx + r0;
               pbtma::bulk_g2s(dst + (size_t)c * TR, src, rows_even * sizeof(double), &full[s]);
            }
         if (++s == nstages) s = 0, ph ^= 1;
      }
      return;
   }

   if (warp < nupd) {
      // ------------------------------ update warps ------------------------------
      const int p2 = 2 * tid;  // first of this thread's two rows (tid < 64)
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
         const int buf = it & 1;
         const uint32_t xph = (it >> 1) & 1;
         pbtma::mbar_wait(&full[s], ph);
         const double *st = stage0 + (size_t)s * kc * TR;
         double t0[BT], t1[BT];
#pragma unroll
         for (int c = 0; c < BT; c++) t0[c] = 0.0, t1[c] = 0.0;
#pragma unroll 4
         for (int j = 0; j < k; j++) {
            const double2 v = *reinterpret_cast<const double2 *>(st + (size_t)j * TR + p2);
#pragma unroll
            for (int c = 0; c < BT; c++) {
               const double cj = Cs[j * BT + c];
               t0[c] += v.x * cj;
               t1[c] += v.y * cj;
            }
         }
         double x0[BT], x1[BT];
#pragma unroll
         for (int c = 0; c < BT; c++) {
            double2 xv = make_double2(0.0, 0.0);
            if (c < a.b) xv = *reinterpret_cast<const double2 *>(st + (size_t)(k + c) * TR + p2);
            x0[c] = xv.x - t0[c], x1[c] = xv.y - t1[c];
         }
         if (a.has_Y) {
            double y0[BT], y1[BT];
#pragma unroll
            for (int c = 0; c < BT; c++) {
               double s0 = 0.0, s1 = 0.0;
#pragma unroll
               for (int cc = 0; cc < BT; cc++) {
                  const double ycc = Ys[cc * BT + c];
                  s0 += x0[cc] * ycc, s1 += x1[cc] * ycc;
               }
               y0[c] = s0, y1[c] = s1;
            }
#pragma unroll
            for (int c = 0; c < BT; c++) x0[c] = y0[c], x1[c] = y1[c];
         }
         // the stage is no longer needed by this warp
         __syncwarp();
         if (lane == 0) pbtma::mbar_arrive(&empty[s]);
         if (++s == nstages) s = 0, ph ^= 1;

         const int64_t r = tile * TR + p2;
#pragma unroll
         for (int c = 0; c < BT; c++)
            if (c < a.b) {
               if (r + 1 < a.n)
                  *reinterpret_cast<double2 *>(a.X + r + (size_t)c * a.ldx) = make_double2(x0[c], x1[c]);
               else if (r < a.n)
                  a.X[r + (size_t)c * a.ldx] = x0[c];
            }
         if (ngw > 0) {
            pbtma::mbar_wait(&xempty[buf], xph ^ 1);  // Gram warps are done with this buffer
            double *xb = xs + (size_t)buf * BT * TR;
#pragma unroll
            for (int c = 0; c < BT; c++)
               *reinterpret_cast<double2 *>(xb + (size_t)c * TR + p2) = make_double2(x0[c], x1[c]);
            __syncwarp();
            if (lane == 0) pbtma::mbar_arrive(&xfull[buf]);
         }
      }
      return;
   }

   if (ngw == 0) {
      return;  // update-only sweep: the remaining warps have nothing to do
   }

   // ------------------------------ Gram warps ------------------------------
   const int gw = warp - nupd;
   double acc[CPW][BT];
   double accx[2][BT];
#pragma unroll
   for (int j = 0; j < CPW; j++)
#pragma unroll
      for (int c = 0; c < BT; c++) acc[j][c] = 0.0;
#pragma unroll
   for (int c = 0; c < BT; c++) accx[0][c] = 0.0, accx[1][c] = 0.0;
   const int xr0 = gw, xr1 = gw + ngw;  // rows of the X'X block owned by this warp

   {
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
         const int buf = it & 1;
         const uint32_t xph = (it >> 1) & 1;
         pbtma::mbar_wait(&full[s], ph);
         const double *st = stage0 + (size_t)s * kc * TR;
         const double *xsrc = st + (size_t)k * TR;
         if (nupd) {
            pbtma::mbar_wait(&xfull[buf], xph);
            xsrc = xs + (size_t)buf * BT * TR;
         }
#pragma unroll
         for (int i2 = 0; i2 < TR / 64; i2++) {
            const int p2 = 2 * (i2 * 32 + lane);
            double xv0[BT], xv1[BT];
#pragma unroll
            for (int c = 0; c < BT; c++) {
               double2 xv = make_double2(0.0, 0.0);
               if (c < a.b) xv = *reinterpret_cast<const double2 *>(xsrc + (size_t)c * TR + p2);
               xv0[c] = xv.x, xv1[c] = xv.y;
            }
#pragma unroll
            for (int j = 0; j < CPW; j++) {
               const int jj = gw * CPW + j;
               double2 av = make_double2(0.0, 0.0);
               if (jj < k) av = *reinterpret_cast<const double2 *>(st + (size_t)jj * TR + p2);
#pragma unroll
               for (int c = 0; c < BT; c++) acc[j][c] += av.x * xv0[c] + av.y * xv1[c];
            }
            if (a.xx) {
               if (xr0 < a.b) {
                  const double2 xw = *reinterpret_cast<const double2 *>(xsrc + (size_t)xr0 * TR + p2);
#pragma unroll
                  for (int c = 0; c < BT; c++) accx[0][c] += xw.x * xv0[c] + xw.y * xv1[c];
               }
               if (xr1 < a.b) {
                  const double2 xw = *reinterpret_cast<const double2 *>(xsrc + (size_t)xr1 * TR + p2);
#pragma unroll
                  for (int c = 0; c < BT; c++) accx[1][c] += xw.x * xv0[c] + xw.y * xv1[c];
               }
            }
         }
         __syncwarp();
         if (lane == 0) {
            if (nupd) pbtma::mbar_arrive(&xempty[buf]);
            pbtma::mbar_arrive(&empty[s]);
         }
         if (++s == nstages) s = 0, ph ^= 1;
      }
   }

   const int rows = k + (a.xx ? a.b : 0);
   double *out = a.partials + (size_t)blockIdx.x * rows * a.b;
#pragma unroll
   for (int j = 0; j < CPW; j++) {
#pragma unroll
      for (int c = 0; c < BT; c++) {
         double v = acc[j][c];
#pragma unroll
         for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
         const int jj = gw * CPW + j;
         if (lane == 0 && jj < k && c < a.b) out[jj + (size_t)c * rows] = v;
      }
   }
   if (a.xx) {
#pragma unroll
      for (int h2 = 0; h2 < 2; h2++) {
         const int xr = h2 == 0 ? xr0 : xr1;
#pragma unroll
         for (int c = 0; c < BT; c++) {
            double v = accx[h2][c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && xr < a.b && c < a.b) out[k + xr + (size_t)c * rows] = v;
         }
      }
   }
}

template <int BT, int CPW>
int launch_sweep_ws(pb200_ctx *ctx, const SweepArgs &a, int &grid, size_t shmem, int nstages) {
   auto kern = ortho_sweep_ws_kernel<BT, CPW>;
   static size_t cached_shmem = 0;
   static int cached_occ = 0;
   if (cached_shmem != shmem || cached_occ == 0) {
      PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
      int o = 1;
      PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, NCT + 32, shmem));
      cached_occ = o < 1 ? 1 : o;
      cached_shmem = shmem;
   }
   if (grid > cached_occ * ctx->num_sms) grid = cached_occ * ctx->num_sms;
   kern<<<grid, NCT + 32, shmem, ctx->stream>>>(a, ctx->coef, nstages);
   ctx->launches++;
   PB_CUDA(cudaGetLastError());
   return 0;
}

template <int BT>
int dispatch_cpw_ws(pb200_ctx *ctx, const SweepArgs &a, int cpw, int &grid, size_t shmem, int nst) {
   switch (cpw) {
   case 1: return launch_sweep_ws<BT, 1>(ctx, a, grid, shmem, nst);
   case 2: return launch_sweep_ws<BT, 2>(ctx, a, grid, shmem, nst);
   case 3: return launch_sweep_ws<BT, 3>(ctx, a, grid, shmem, nst);
   case 4: return launch_sweep_ws<BT, 4>(ctx, a, grid, shmem, nst);
   case 5: return launch_sweep_ws<BT, 5>(ctx, a, grid, shmem, nst);
   case 6: return launch_sweep_ws<BT, 6>(ctx, a, grid, shmem, nst);
   case 7:
   case 8: return launch_sweep_ws<BT, 8>(ctx, a, grid, shmem, nst);
   case 9:
   case 10: return launch_sweep_ws<BT, 10>(ctx, a, grid, shmem, nst);
   default: return PB200_ERR_ARG;
   }
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

   // ---- v2 (TMA-staged) eligibility: 16-byte aligned column segments, enough rows, tile ring
   // of >= 2 stages in shared memory ----
   const int kc = k + b;
   auto aligned16 = [](const void *p) { return (((uintptr_t)p) & 15) == 0; };
   bool tma_ok = ctx->use_tma && n >= 4 * TR && kc <= 96 && aligned16(X) && (ldx % 2 == 0) &&
                 (q == 0 || (aligned16(Q) && ldq % 2 == 0)) && (mv == 0 || (aligned16(V) && ldv % 2 == 0));
   int nstages = 0, ctas_per_sm = 1;
   size_t fixed_sm = ((size_t)k * BT + BT * BT + (size_t)BT * TR + 2 * (size_t)TR * BT) * sizeof(double) + 128;
   if (tma_ok) {
      const size_t stage_b = (size_t)kc * TR * sizeof(double);
      if (3 * stage_b + fixed_sm <= 110 * 1024) {
         nstages = 3, ctas_per_sm = 2;
      } else if (ctx->ortho_2cta && 2 * stage_b + fixed_sm <= 110 * 1024) {
         // two resident CTAs (16 consumer warps per SM) with a 2-deep ring each hide the
         // shared-memory latency of the Gram warps better than one CTA with a 4-deep ring
         nstages = 2, ctas_per_sm = 2;
      } else {
         nstages = (int)((220 * 1024 - fixed_sm) / stage_b);
         if (nstages > 4) nstages = 4;
      }
      if (nstages < 2) tma_ok = false;
   }

   int grid = 0, nparts = 0, rc = 0;
   int ps = pb_prof_begin(ctx, PB_K_ORTHO);
   if (tma_ok) {
      // v3 (warp-specialised, opt-in) when its column split fits the register budget
      const int ngw = a.do_gram ? NCW - (a.do_update ? 2 : 0) : 0;
      const int cpw_ws = ngw > 0 ? (k + ngw - 1) / ngw : 1;
      const bool ws = ctx->use_ws && cpw_ws <= 10 && (ngw == 0 || 2 * ngw >= b);
      const int64_t ntiles = (n + TR - 1) / TR;  // both kernels handle a partial last tile
      const int64_t n_main = n;
      grid = (int)(ntiles < (int64_t)ctx->num_sms * ctas_per_sm ? ntiles : (int64_t)ctx->num_sms * ctas_per_sm);
      const int tail = n_main < n ? 1 : 0;
      if (a.do_gram) {
         PB_CHK(pb_ensure_partials(ctx, (size_t)(grid + tail) * rows * b));
         a.partials = ctx->d_partials;
      }
      SweepArgs am = a;
      am.n = n_main;
      size_t shmem = (size_t)nstages * kc * TR * sizeof(double) + fixed_sm + 2 * nstages * sizeof(uint64_t);
      if (ws) {
         const int cw = cpw_ws < 1 ? 1 : cpw_ws;
         switch (BT) {
         case 1: rc = dispatch_cpw_ws<1>(ctx, am, cw, grid, shmem, nstages); break;
         case 2: rc = dispatch_cpw_ws<2>(ctx, am, cw, grid, shmem, nstages); break;
         case 4: rc = dispatch_cpw_ws<4>(ctx, am, cw, grid, shmem, nstages); break;
         default: rc = dispatch_cpw_ws<8>(ctx, am, cw, grid, shmem, nstages); break;
         }
      } else {
         switch (BT) {
         case 1: rc = dispatch_cpw_tma<1>(ctx, am, cpw, grid, shmem, nstages); break;
         case 2: rc = dispatch_cpw_tma<2>(ctx, am, cpw, grid, shmem, nstages); break;
         case 4: rc = dispatch_cpw_tma<4>(ctx, am, cpw, grid, shmem, nstages); break;
         default: rc = dispatch_cpw_tma<8>(ctx, am, cpw, grid, shmem, nstages); break;
         }
      }
      nparts = grid;
      if (!rc && tail) {
         // rows [n_main, n): one CTA of the v1 kernel, its partial panel goes to slot `grid`
         SweepArgs at = a;
         at.n = n - n_main;
         at.Q = Q ? Q + n_main : Q, at.V = V ? V + n_main : V, at.X = X + n_main;
         if (a.do_gram) at.partials = ctx->d_partials + (size_t)grid * rows * b;
         size_t shmem1 = ((size_t)k * BT + BT * BT + (size_t)BT * TILE) * sizeof(double);
         switch (BT) {
         case 1: rc = dispatch_cpw<1>(ctx, at, cpw, 1, shmem1); break;
         case 2: rc = dispatch_cpw<2>(ctx, at, cpw, 1, shmem1); break;
         case 4: rc = dispatch_cpw<4>(ctx, at, cpw, 1, shmem1); break;
         default: rc = dispatch_cpw<8>(ctx, at, cpw, 1, shmem1); break;
         }
         nparts = grid + 1;
      }
   } else {
      const int64_t ntiles = (n + TILE - 1) / TILE;
      grid = (int)(ntiles < (int64_t)ctx->num_sms * 3 ? ntiles : (int64_t)ctx->num_sms * 3);
      if (grid < 1) grid = 1;
      if (a.do_gram) {
         PB_CHK(pb_ensure_partials(ctx, (size_t)grid * rows * b));
         a.partials = ctx->d_partials;
      }
      size_t shmem = ((size_t)k * BT + BT * BT + (size_t)BT * TILE) * sizeof(double);
      switch (BT) {
      case 1: rc = dispatch_cpw<1>(ctx, a, cpw, grid, shmem); break;
      case 2: rc = dispatch_cpw<2>(ctx, a, cpw, grid, shmem); break;
      case 4: rc = dispatch_cpw<4>(ctx, a, cpw, grid, shmem); break;
      default: rc = dispatch_cpw<8>(ctx, a, cpw, grid, shmem); break;
      }
      nparts = grid;
   }
   pb_prof_end(ctx, ps, abytes);
   PB_CHK(rc);
   grid = nparts;
   if (a.do_gram) {
      PB_CHK(pb_finish_panel(ctx, grid, rows * b));
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
            PB_CUDA(cudaMemsetAsync(ctx->d_panel, 0, sizeof(double) * rows * b, ctx->stream));
            PB_CHK(pb_nccl_allreduce_dev(ctx, ctx->d_panel, rows * b));
            PB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_panel, sizeof(double) * rows * b,
                  cudaMemcpyDeviceToHost, ctx->stream));
            PB_CUDA(cudaStreamSynchronize(ctx->stream));
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

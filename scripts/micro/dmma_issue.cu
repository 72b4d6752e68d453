// dmma_issue.cu -- does a DMMA m8n8k4 block the issue port of its SM sub-partition?  K independent
// integer (or shared-memory load) instructions are interleaved with every DMMA; if the DMMA rate
// stays at its peak until K ~ 15 the other instructions issue "for free" in the 16-cycle shadow of
// a DMMA, otherwise time ~ 16 + K cycles per DMMA.  Also times the larger fp64 MMA shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/dmma_issue scripts/micro/dmma_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
   asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
template <int K, int MODE>  // MODE 0: integer ALU filler, 1: shared-memory load filler
__global__ void mix_kernel(double *out, int iters, double a, double b, int seed) {
   __shared__ double sm[1024];
   for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
   __syncthreads();
   double c0[4], c1[4];
#pragma unroll
   for (int i = 0; i < 4; i++) c0[i] = threadIdx.x + i, c1[i] = i;
   unsigned x[K > 0 ? K : 1];
   double ld[K > 0 ? K : 1];
#pragma unroll
   for (int i = 0; i < (K > 0 ? K : 1); i++) x[i] = seed + i + threadIdx.x, ld[i] = 0;
   for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
         dmma(c0[i], c1[i], a, b);
#pragma unroll
         for (int k = 0; k < K; k++) {
            if (MODE == 0) x[k] = x[k] * 1664525u + 1013904223u;
            else ld[k] += sm[(threadIdx.x + 32 * k + it) & 1023];
         }
      }
   }
   double s = 0;
#pragma unroll
   for (int i = 0; i < 4; i++) s += c0[i] + c1[i];
#pragma unroll
   for (int i = 0; i < (K > 0 ? K : 1); i++) s += x[i] + ld[i];
   if (s == 12345.678) out[0] = s;
}
__global__ void big_kernel(double *out, int iters, double av, double bv) {
   double c[4][4], a[8], b[4];
   for (int i = 0; i < 8; i++) a[i] = av + i;
   for (int i = 0; i < 4; i++) b[i] = bv + i;
   for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) c[j][i] = threadIdx.x + i;
   for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int j = 0; j < 4; j++) dmma16816(c[j], a, b);
   }
   double s = 0;
   for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) s += c[j][i];
   if (s == 12345.678) out[0] = s;
}
template <typename F> static float time_ms(F f) {
   cudaEvent_t e0, e1; cudaEventCreate(&e0), cudaEventCreate(&e1);
   f(); cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
   float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
template <int K, int MODE> void run(double *out, int sms, int wps) {
   const int iters = 5000, threads = 32 * wps, blocks = sms;
   float ms = time_ms([&] { mix_kernel<K, MODE><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, 3); });
   double ndmma = 4.0 * iters * wps / 4;  // per SM sub-partition
   printf("%s filler K=%2d, %2d warps/SM: %7.3f ms  %6.1f cycles per DMMA per sub-partition (1.965 GHz)  %6.2f TFLOP/s\n",
         MODE ? "LDS" : "ALU", K, wps, ms, ms * 1e-3 * 1.965e9 / ndmma, 512.0 * 4 * iters * wps * sms / ms / 1e9);
}
int main() {
   double *out; cudaMalloc(&out, 8);
   cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
   const int sms = p.multiProcessorCount;
   for (int wps : {4, 16}) {
      run<0, 0>(out, sms, wps); run<2, 0>(out, sms, wps); run<4, 0>(out, sms, wps); run<8, 0>(out, sms, wps);
      run<12, 0>(out, sms, wps); run<16, 0>(out, sms, wps); run<24, 0>(out, sms, wps);
      run<1, 1>(out, sms, wps); run<2, 1>(out, sms, wps); run<4, 1>(out, sms, wps); run<8, 1>(out, sms, wps);
   }
   for (int wps : {4, 8, 16}) {
      const int iters = 2000, threads = 32 * wps;
      float ms = time_ms([&] { big_kernel<<<sms, threads>>>(out, iters, 1.0000001, 1e-9); });
      printf("DMMA m16n8k16, %2d warps/SM: %7.3f ms  %6.2f TFLOP/s\n", wps, ms, 2.0 * 16 * 8 * 16 * 4 * iters * wps * sms / ms / 1e9);
   }
   printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
   return 0;
}

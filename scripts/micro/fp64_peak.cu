// fp64_peak.cu -- measured fp64 throughput of one B200: DFMA (CUDA cores) vs DMMA m8n8k4 (tensor
// pipe), register-resident operands, 8..16 independent chains per thread.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_peak scripts/micro/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double *out, int iters, double a, double b) {
   double acc[16];
#pragma unroll
   for (int i = 0; i < 16; i++) acc[i] = threadIdx.x + i;
   for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
   }
   double s = 0;
#pragma unroll
   for (int i = 0; i < 16; i++) s += acc[i];
   if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(c0), "+d"(c1)
                : "d"(a), "d"(b));
}

template <int NCH>
__global__ void dmma_kernel(double *out, int iters, double a, double b) {
   double c0[NCH], c1[NCH];
#pragma unroll
   for (int i = 0; i < NCH; i++) c0[i] = threadIdx.x + i, c1[i] = i;
   for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < NCH; i++) dmma(c0[i], c1[i], a, b);
   }
   double s = 0;
#pragma unroll
   for (int i = 0; i < NCH; i++) s += c0[i] + c1[i];
   if (s == 12345.678) out[0] = s;
}

template <typename F>
static float time_ms(F f) {
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0), cudaEventCreate(&e1);
   f();
   cudaDeviceSynchronize();
   cudaEventRecord(e0);
   f();
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms;
   cudaEventElapsedTime(&ms, e0, e1);
   return ms;
}

int main() {
   double *out;
   cudaMalloc(&out, 8);
   cudaDeviceProp p;
   cudaGetDeviceProperties(&p, 0);
   const int sms = p.multiProcessorCount;
   const int iters = 20000;
   for (int wps = 4; wps <= 32; wps *= 2) {  // warps per SM
      const int threads = 32 * (wps > 8 ? 8 : wps), blocks = sms * (wps > 8 ? wps / 8 : 1);
      float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 16 * iters * (double)threads * blocks;
      printf("DFMA  %2d warps/SM: %8.3f ms  %7.2f TFLOP/s\n", wps, ms, fl / ms / 1e9);
   }
   for (int wps = 4; wps <= 32; wps *= 2) {
      const int threads = 32 * (wps > 8 ? 8 : wps), blocks = sms * (wps > 8 ? wps / 8 : 1);
      float ms = time_ms([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 256 * 8 * iters * (double)(threads / 32) * blocks;
      printf("DMMA8 %2d warps/SM: %8.3f ms  %7.2f TFLOP/s\n", wps, ms, fl / ms / 1e9);
      ms = time_ms([&] { dmma_kernel<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      fl = 2.0 * 256 * 2 * iters * (double)(threads / 32) * blocks;
      printf("DMMA2 %2d warps/SM: %8.3f ms  %7.2f TFLOP/s\n", wps, ms, fl / ms / 1e9);
   }
   printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
   return 0;
}

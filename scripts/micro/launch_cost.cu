// launch_cost.cu -- host-side cost of one kernel launch as a function of the kernel-parameter size,
// of a cudaEventRecord, and the launch->completion->host-visible round trip through mapped memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_cost launch_cost.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <time.h>
template <int N> struct P { double v[N]; };
template <int N> __global__ void k(const __grid_constant__ P<N> p, volatile long long *flag, long long seq) {
   if (threadIdx.x == 0 && blockIdx.x == 0) *flag = seq + (p.v[0] > 1e300);
}
static double now() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }
template <int N> void run(cudaStream_t s, volatile long long *hflag, long long *dflag, int grid, int threads, size_t smem) {
   static P<N> p;
   cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
   long long seq = 1000 * N;
   for (int i = 0; i < 50; i++) k<N><<<grid, threads, smem, s>>>(p, dflag, ++seq);
   cudaStreamSynchronize(s);
   const int reps = 500;
   double tl = 0, tt = 0;
   for (int i = 0; i < reps; i++) {
      double t0 = now();
      k<N><<<grid, threads, smem, s>>>(p, dflag, ++seq);
      double t1 = now();
      while (*hflag != seq) {}
      double t2 = now();
      tl += t1 - t0, tt += t2 - t0;
   }
   printf("params %5zu B grid %3d x %3d smem %6zu: launch call %.2f us, launch->flag visible %.2f us\n", sizeof(P<N>), grid, threads, smem, tl / reps, tt / reps);
}
int main() {
   cudaStream_t s; cudaStreamCreate(&s);
   long long *h, *d;
   cudaHostAlloc((void **)&h, 64, cudaHostAllocMapped); *h = 0;
   cudaHostGetDevicePointer((void **)&d, h, 0);
   run<1>(s, h, d, 1, 32, 0);
   run<1>(s, h, d, 148, 544, 0);
   run<1>(s, h, d, 148, 544, 227 * 1024);
   run<64>(s, h, d, 148, 544, 227 * 1024);
   run<256>(s, h, d, 148, 544, 227 * 1024);
   run<496>(s, h, d, 148, 544, 227 * 1024);
   run<512>(s, h, d, 148, 544, 227 * 1024);
   run<1024>(s, h, d, 148, 544, 227 * 1024);
   run<1100>(s, h, d, 148, 544, 227 * 1024);
   run<2048>(s, h, d, 148, 544, 227 * 1024);
   cudaEvent_t e; cudaEventCreate(&e);
   double t0 = now();
   for (int i = 0; i < 1000; i++) cudaEventRecord(e, s);
   printf("cudaEventRecord %.2f us\n", (now() - t0) / 1000);
   cudaStreamSynchronize(s);
   // alternating shared-memory configurations (carve-out switch)
   {
      static P<1> p; long long seq = 5;
      cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      double t0 = now();
      for (int i = 0; i < 500; i++) {
         k<1><<<148, 544, 227 * 1024, s>>>(p, d, ++seq);
         k<1><<<592, 128, 2048, s>>>(p, d, ++seq);
      }
      cudaStreamSynchronize(s);
      printf("alternating 227 KB / 2 KB smem kernels: %.2f us per pair (device-throughput bound)\n", (now() - t0) / 500);
      t0 = now();
      for (int i = 0; i < 500; i++) {
         k<1><<<148, 544, 227 * 1024, s>>>(p, d, ++seq);
         k<1><<<148, 544, 227 * 1024, s>>>(p, d, ++seq);
      }
      cudaStreamSynchronize(s);
      printf("same-config pairs: %.2f us per pair\n", (now() - t0) / 500);
   }
   return 0;
}

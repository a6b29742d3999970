// Dependent-issue latency of DFMA and DMMA.8x8x4 on B200 (sm_100a): ONE warp per SM sub-partition (128-thread blocks, one per SM),
// ILP independent accumulation chains per warp.  cycles per instruction = elapsed SM cycles / (iters * ILP); with ILP = 1 that is the
// latency a dependent chain pays, and the ILP at which it stops improving tells how many chains a warp needs to fill the pipe alone.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, bool MMA>
__global__ void __launch_bounds__(128) chain_kernel(double* out, long long* cyc, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (MMA)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
      else
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c0[i]) : "d"(a), "d"(b));
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP, bool MMA>
void run(double* out, long long* cyc, int sms) {
  const int iters = 4000;
  chain_kernel<ILP, MMA><<<sms, 128>>>(out, cyc, iters, 1.0000001, 1e-9);
  chain_kernel<ILP, MMA><<<sms, 128>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; i++) avg += (double)h[i];
  avg /= sms;
  printf("%s ilp %d: %.1f cycles per instruction per warp (%.1f per chain step)\n", MMA ? "DMMA.8x8x4" : "DFMA      ", ILP, avg / (iters * (double)ILP),
         avg / iters);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("device %s sms %d\n", p.name, sms);
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * sms * 128);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  run<1, false>(out, cyc, sms); run<2, false>(out, cyc, sms); run<4, false>(out, cyc, sms); run<8, false>(out, cyc, sms);
  run<1, true>(out, cyc, sms); run<2, true>(out, cyc, sms); run<4, true>(out, cyc, sms); run<8, true>(out, cyc, sms); run<16, true>(out, cyc, sms);
  return 0;
}

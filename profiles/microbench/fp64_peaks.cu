// FP64 pipe microbenchmark for B200 (sm_100a): DFMA vs DMMA.8x8x4 issue rates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

template<int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int ILP>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// exp() throughput (library fp64 exp)
__global__ void __launch_bounds__(256) exp_kernel(double* out, int iters, double a) {
  double x0 = -1e-3 * threadIdx.x, x1 = x0 - 0.5, x2 = x0 - 1.0, x3 = x0 - 1.5;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int it = 0; it < iters; it++) {
    s0 += exp(x0); s1 += exp(x1); s2 += exp(x2); s3 += exp(x3);
    x0 -= a; x1 -= a; x2 -= a; x3 -= a;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}

template<typename F> float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s sms %d\n", p.name, sms);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256 * 4));
  const int iters = 20000;
  for (int bps = 1; bps <= 8; bps *= 2) {
    int grid = sms * bps;
    float ms = time_ms([&]{ dfma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double fl = 2.0 * 8 * iters * 256.0 * grid;
    printf("DFMA  ilp8 blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
  }
  for (int bps = 1; bps <= 8; bps *= 2) {
    int grid = sms * bps;
    float ms = time_ms([&]{ dmma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double fl = 2.0 * 256 * 8 * iters * 8.0 * grid;  // 8 warps/block, 256 MAC per mma
    printf("DMMA884 ilp8 blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
  }
  {
    int grid = sms * 2;
    float ms = time_ms([&]{ dmma_kernel<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double fl = 2.0 * 256 * 16 * iters * 8.0 * grid;
    printf("DMMA884 ilp16 blocks/SM 2: %.3f ms  %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
  }
  {
    int grid = sms * 4;
    float ms = time_ms([&]{ exp_kernel<<<grid, 256>>>(out, 2000, 1e-4); }, 5);
    double n = 4.0 * 2000 * 256.0 * grid;
    printf("exp(fp64): %.3f ms  %.2f Gexp/s\n", ms, n / ms * 1e-6);
  }
  // sustained DMMA for ~3 s to see the power-capped rate
  {
    int grid = sms * 2;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    int n = 0; float ms = 0;
    do { for (int i = 0; i < 10; i++) dmma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); n += 10;
         cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); } while (ms < 3000);
    double fl = 2.0 * 256 * 8 * iters * 8.0 * grid * n;
    printf("DMMA884 sustained %.0f ms: %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
  }
  return 0;
}

// Are DMMA.8x8x4 and DFMA separate pipes on B200?  Half of the warps issue DMMA, the other half DFMA; if the pipes were
// independent the combined rate would approach the sum of the two peaks (37 + 34 TFLOP/s).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

// mode 0: all warps DMMA; 1: all warps DFMA; 2: even warps DMMA, odd warps DFMA; 3: every warp interleaves both
__global__ void __launch_bounds__(256) mix_kernel(double* out, int iters, int mode, int fma_per_mma, double a, double b) {
  double c0[8], c1[8], f[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; f[i] = 1e-3 * i; }
  const int warp = threadIdx.x >> 5;
  const bool do_mma = mode == 0 || mode == 3 || (mode == 2 && (warp & 1) == 0);
  const bool do_fma = mode == 1 || mode == 3 || (mode == 2 && (warp & 1) == 1);
  for (int it = 0; it < iters; it++) {
    if (do_mma) {
#pragma unroll
      for (int i = 0; i < 8; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    if (do_fma) {
      for (int r = 0; r < fma_per_mma; r++) {
#pragma unroll
        for (int i = 0; i < 8; i++) f[i] = fma(f[i], a, b);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c0[i] + c1[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount, grid = sms * 4, iters = 4000;
  double* out; cudaMalloc(&out, sizeof(double) * grid * 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int fpm = 2; fpm <= 8; fpm *= 2) {
    for (int mode = 0; mode < 4; mode++) {
      mix_kernel<<<grid, 256>>>(out, iters, mode, fpm, 1.0000001, 1e-9);
      cudaDeviceSynchronize();
      float best = 1e30f;
      for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); mix_kernel<<<grid, 256>>>(out, iters, mode, fpm, 1.0000001, 1e-9); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      double warps = 8.0 * grid;
      double mma_w = mode == 0 || mode == 3 ? warps : (mode == 2 ? warps / 2 : 0), fma_w = mode == 1 || mode == 3 ? warps : (mode == 2 ? warps / 2 : 0);
      double fl_mma = mma_w * iters * 8 * 512.0, fl_fma = fma_w * iters * 8.0 * fpm * 64.0;
      printf("fma_per_mma %d mode %d: %.3f ms  DMMA %.2f + DFMA %.2f = %.2f TFLOP/s\n", fpm, mode, best, fl_mma / best * 1e-9,
             fl_fma / best * 1e-9, (fl_mma + fl_fma) / best * 1e-9);
    }
  }
  return 0;
}

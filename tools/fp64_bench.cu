// dev microbenchmark: FP64 DADD/DMUL and F2F conversion issue rate on the device (not product code)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dp(double *out, int iters, double k) {
  double a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = __dadd_rn(__dmul_rn(a[i], k), 1e-9);
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_cvt(float *out, int iters, double k) {
  float a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = (float)__dmul_rn((double)a[i], k);
  }
  float s = 0;
  for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double *d; cudaMalloc(&d, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0);
    k_dp<<<148 * 8, 256>>>(d, iters, 0.999999);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 8 * 256 * iters * 16;
    printf("DADD+DMUL: %.2f T instr/s (%.1f per clk per SM at 1.965 GHz)\n", ops / ms / 1e9, ops / ms / 1e9 * 1e3 / 148 / 1.965);
    cudaEventRecord(e0);
    k_cvt<<<148 * 8, 256>>>((float *)d, iters, 0.999999);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    ops = 148.0 * 8 * 256 * iters * 8;
    printf("F2F.64.32 + DMUL + F2F.32.64 triplets: %.2f T/s (%.1f per clk per SM) => %.1f fp64-pipe ops per clk per SM if all three share the pipe\n",
           ops / ms / 1e9, ops / ms / 1e9 * 1e3 / 148 / 1.965, 3 * ops / ms / 1e9 * 1e3 / 148 / 1.965);
  }
  return 0;
}

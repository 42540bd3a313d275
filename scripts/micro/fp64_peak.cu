// FP64 FMA throughput of the device: 8 independent chains per thread, no memory traffic.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, int iters, double a, double b) {
  double x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void kf(float* out, int iters, float a, float b) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  const int grid = 148 * 8, block = 256, iters = 20000;
  double* d; cudaMalloc(&d, sizeof(double) * grid * block);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<grid, block>>>(d, 100, 1.0000001, 1e-9); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<<<grid, block>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("FP64: %.3f ms, %.2f TFLOP/s\n", ms, 2.0 * grid * block * 8.0 * iters / ms / 1e9);
  kf<<<grid, block>>>((float*)d, 100, 1.0000001f, 1e-9f); cudaDeviceSynchronize();
  cudaEventRecord(e0); kf<<<grid, block>>>((float*)d, iters, 1.0000001f, 1e-9f); cudaEventRecord(e1); cudaDeviceSynchronize();
  cudaEventElapsedTime(&ms, e0, e1);
  printf("FP32: %.3f ms, %.2f TFLOP/s\n", ms, 2.0 * grid * block * 8.0 * iters / ms / 1e9);
  return 0;
}

// FP64 ceilings of the device, for the roofline of the frontal GEMMs ("DGEMM peak measured on the box"):
//   1. FP64 FMA on the CUDA cores (8 independent chains per thread, no memory traffic)
//   2. FP64 tensor path: mma.sync m8n8k4 and m16n8k8 (DMMA), accumulators chained, no memory traffic
//   3. cuBLAS DGEMM 4096^3 and 8192 x 8192 x 128 (the shape of a frontal Schur update) — yardstick only
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/micro/fp64_peak.cu -lcublas -o gpurun_out/fp64_peak
#include <cstdio>
#include <cublas_v2.h>
#include <cuda_runtime.h>
__global__ void k_fma(double* out, int iters, double a, double b) {
  double x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma884(double* out, int iters) {
  double c[4][2];
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0; for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma1688(double* out, int iters) {
  double c[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
  const double a0 = 1.0 + threadIdx.x * 1e-6, a1 = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(a0), "d"(a1), "d"(a1), "d"(a0));
  }
  double s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timed(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  const int grid = 148 * 8, block = 256, iters = 20000;
  double* d; cudaMalloc(&d, sizeof(double) * grid * block);
  float ms = timed([&] { k_fma<<<grid, block>>>(d, iters, 1.0000001, 1e-9); });
  printf("FP64 FMA (CUDA cores)      : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * grid * block * 8.0 * iters / ms / 1e9);
  ms = timed([&] { k_dmma884<<<grid, block>>>(d, iters); });
  printf("DMMA mma.sync m8n8k4  f64  : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * 8 * 8 * 4 * 4.0 * iters * (grid * block / 32) / ms / 1e9);
  ms = timed([&] { k_dmma1688<<<grid, block>>>(d, iters); });
  printf("DMMA mma.sync m16n8k8 f64  : %8.3f ms  %7.2f TFLOP/s\n", ms, 2.0 * 16 * 8 * 8 * 4.0 * iters * (grid * block / 32) / ms / 1e9);
  cublasHandle_t h; cublasCreate(&h);
  const double one = 1.0, zero = 0.0;
  for (int shape = 0; shape < 3; ++shape) {
    const int M = shape == 0 ? 4096 : (shape == 1 ? 8192 : 2048), N = M, K = shape == 0 ? 4096 : 128;
    double *A, *B, *C; cudaMalloc(&A, sizeof(double) * M * K); cudaMalloc(&B, sizeof(double) * K * N); cudaMalloc(&C, sizeof(double) * M * N);
    cudaMemset(A, 0, sizeof(double) * M * K); cudaMemset(B, 0, sizeof(double) * K * N);
    ms = timed([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, M, N, K, &one, A, M, B, N, &zero, C, M); });
    printf("cuBLAS DGEMM %5d x %5d x %4d: %8.3f ms  %7.2f TFLOP/s\n", M, N, K, ms, 2.0 * M * N * K / ms / 1e9);
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  return 0;
}

// Micro-benchmark of the sweep tasks' inner loop: one warp multiplies a panel in shared memory (n columns of ld doubles,
// lanes over rows, R rows per lane) with NR right-hand sides.  Prints cycles per column for several formulations, with 1 and
// with 3 warps per SM sub-partition resident.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o panel_fma panel_fma.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NR = 4;
constexpr int NCOL = 128;

template <int R, int VARIANT>
__global__ void __launch_bounds__(32) k(const double* __restrict__ gpanel, const double* __restrict__ gy, double* out, long long* cyc, int ld, int reps) {
  extern __shared__ double sm[];
  double* panel = sm;                     // NCOL * ld
  double* yv = sm + NCOL * 128;           // NCOL * NR
  const int lane = threadIdx.x;
  for (int i = lane; i < NCOL * ld; i += 32) panel[i] = gpanel[i];
  for (int i = lane; i < NCOL * NR; i += 32) yv[i] = gy[i];
  __syncwarp();
  double acc[4][R][NR];
  for (int u = 0; u < 4; ++u) for (int q = 0; q < R; ++q) for (int r = 0; r < NR; ++r) acc[u][q][r] = 0.0;
  bool valid[R];
  for (int q = 0; q < R; ++q) valid[q] = lane + 32 * q < ld;
  const long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    if (VARIANT == 0) {          // the product's first form: one accumulator set, index arithmetic per load
      const double* base = panel + lane;
#pragma unroll 4
      for (int c = 0; c < NCOL; ++c) {
        const double2 y01 = *reinterpret_cast<const double2*>(yv + c * NR), y23 = *reinterpret_cast<const double2*>(yv + c * NR + 2);
#pragma unroll
        for (int q = 0; q < R; ++q) {
          const double m = valid[q] ? base[c * ld + 32 * q] : 0.0;
          acc[0][q][0] = fma(m, y01.x, acc[0][q][0]); acc[0][q][1] = fma(m, y01.y, acc[0][q][1]);
          acc[0][q][2] = fma(m, y23.x, acc[0][q][2]); acc[0][q][3] = fma(m, y23.y, acc[0][q][3]);
        }
      }
    } else if (VARIANT == 1) {   // four accumulator sets
      const double* base = panel + lane;
#pragma unroll 2
      for (int c = 0; c < NCOL; c += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double2 y01 = *reinterpret_cast<const double2*>(yv + (c + u) * NR), y23 = *reinterpret_cast<const double2*>(yv + (c + u) * NR + 2);
#pragma unroll
          for (int q = 0; q < R; ++q) {
            const double m = valid[q] ? base[(c + u) * ld + 32 * q] : 0.0;
            acc[u][q][0] = fma(m, y01.x, acc[u][q][0]); acc[u][q][1] = fma(m, y01.y, acc[u][q][1]);
            acc[u][q][2] = fma(m, y23.x, acc[u][q][2]); acc[u][q][3] = fma(m, y23.y, acc[u][q][3]);
          }
        }
      }
    } else if (VARIANT == 2) {   // pointer increments, loads of 8 columns first, two accumulator sets
      const double* p = panel + lane;
      const double* y = yv;
      for (int c = 0; c < NCOL; c += 8) {
        double m[8][R]; double2 ya[8], yb[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          ya[u] = *reinterpret_cast<const double2*>(y + u * NR); yb[u] = *reinterpret_cast<const double2*>(y + u * NR + 2);
#pragma unroll
          for (int q = 0; q < R; ++q) m[u][q] = valid[q] ? p[32 * q] : 0.0;
          p += ld;
        }
        y += 8 * NR;
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int q = 0; q < R; ++q) {
            acc[u & 1][q][0] = fma(m[u][q], ya[u].x, acc[u & 1][q][0]); acc[u & 1][q][1] = fma(m[u][q], ya[u].y, acc[u & 1][q][1]);
            acc[u & 1][q][2] = fma(m[u][q], yb[u].x, acc[u & 1][q][2]); acc[u & 1][q][3] = fma(m[u][q], yb[u].y, acc[u & 1][q][3]);
          }
      }
    } else if (VARIANT == 3) {   // no FMAs on the FP64 pipe: loads only (sum of loaded words), to see what the loads alone cost
      const double* p = panel + lane;
      const double* y = yv;
      for (int c = 0; c < NCOL; c += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double2 a = *reinterpret_cast<const double2*>(y + u * NR), b = *reinterpret_cast<const double2*>(y + u * NR + 2);
          double s = a.x;
#pragma unroll
          for (int q = 0; q < R; ++q) s = valid[q] ? p[32 * q] : s;
          acc[u & 3][0][0] = __longlong_as_double(__double_as_longlong(acc[u & 3][0][0]) ^ __double_as_longlong(s) ^ __double_as_longlong(b.y));
          p += ld;
        }
        y += 8 * NR;
      }
    }
  }
  const long long t1 = clock64();
  double s = 0.0;
  for (int u = 0; u < 4; ++u) for (int q = 0; q < R; ++q) for (int r = 0; r < NR; ++r) s += acc[u][q][r];
  out[blockIdx.x * 32 + lane] = s;
  if (lane == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int R, int V>
void run(const char* name, int ld, int ctas_per_sm) {
  int dev = 0; cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
  const int grid = pr.multiProcessorCount * ctas_per_sm, reps = 20;
  double *gp, *gy, *out; long long* cyc;
  cudaMalloc(&gp, NCOL * 128 * 8); cudaMalloc(&gy, NCOL * NR * 8); cudaMalloc(&out, grid * 32 * 8); cudaMalloc(&cyc, grid * 8);
  cudaMemset(gp, 0, NCOL * 128 * 8); cudaMemset(gy, 0, NCOL * NR * 8);
  const size_t smem = (NCOL * 128 + NCOL * NR) * 8;
  cudaFuncSetAttribute(k<R, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<R, V><<<grid, 32, smem>>>(gp, gy, out, cyc, ld, reps);
  k<R, V><<<grid, 32, smem>>>(gp, gy, out, cyc, ld, reps);
  cudaDeviceSynchronize();
  long long h[4096]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  printf("%-34s R=%d ld=%3d warps/SM=%d : %6.1f cycles per column, %5.2f cycles per FMA instruction (%s)\n", name, R, ld, ctas_per_sm,
         avg / (reps * NCOL), avg / (reps * NCOL) / (R * NR), cudaGetErrorString(cudaGetLastError()));
  cudaFree(gp); cudaFree(gy); cudaFree(out); cudaFree(cyc);
}

int main() {
  // shared memory per CTA is 135 KB: one CTA per SM; to get more warps per sub-partition the grid must use smaller panels —
  // here residency is what the driver can fit (1 CTA/SM with this size), so the multi-warp rows repeat the run with a 4-warp... (not done)
  run<1, 0>("one set, indexed", 32, 1);
  run<1, 1>("four sets, indexed", 32, 1);
  run<1, 2>("pointer, 8-column batches", 32, 1);
  run<1, 3>("loads only", 32, 1);
  run<2, 0>("one set, indexed", 64, 1);
  run<2, 2>("pointer, 8-column batches", 64, 1);
  run<4, 0>("one set, indexed", 128, 1);
  run<4, 1>("four sets, indexed", 128, 1);
  run<4, 2>("pointer, 8-column batches", 128, 1);
  run<4, 3>("loads only", 128, 1);
  return 0;
}

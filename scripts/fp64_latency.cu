// Dependent-chain latencies of the FP64 instructions on the Cholesky pivot chain (one warp, clock64).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o conicip.jl_b200/csrc/build_prof/fp64_latency scripts/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed) {
  const int N = 1024;
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0, t1;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
  t1 = clock64(); cyc[0] = t1 - t0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
  t1 = clock64(); cyc[1] = t1 - t0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31);
  t1 = clock64(); cyc[2] = t1 - t0;
  __shared__ double sm[64];
  sm[threadIdx.x] = x; sm[threadIdx.x + 32] = x;
  __syncwarp();
  t0 = clock64();
  int idx = threadIdx.x;
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double v = sm[idx & 63]; idx = (int)(v * 0.0) + idx + 1; }   // LDS -> D2I -> IADD chain
  t1 = clock64(); cyc[3] = t1 - t0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { sm[threadIdx.x] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31] + 1e-9; __syncwarp(); }
  t1 = clock64(); cyc[4] = t1 - t0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * y;
  t1 = clock64(); cyc[5] = t1 - t0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = 1.0 / sqrt(x + 2.0);
  t1 = clock64(); cyc[6] = t1 - t0;
  double c0 = x, c1 = y;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(y), "d"(y));
  t1 = clock64(); cyc[7] = t1 - t0;
  out[threadIdx.x] = x + idx + c0 + c1;
}
int main() {
  double* o; long long* c;
  cudaMalloc(&o, 256); cudaMalloc(&c, 64);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(o, c, 1.5);
  long long h[8];
  cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
  const char* names[] = {"DFMA", "MUFU.RCP64H(rcp.approx.f64)", "SHFL f64 (2 x SHFL)", "LDS.64 + D2I + IADD", "STS+syncwarp+LDS+DADD+syncwarp", "DMUL", "1/sqrt (full)", "DMMA.8x8x4 (dependent)"};
  for (int i = 0; i < 8; ++i) printf("%-34s %.1f cycles per dependent op\n", names[i], h[i] / 1024.0);
  return 0;
}

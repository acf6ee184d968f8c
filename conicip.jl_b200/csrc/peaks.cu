// Register-only FP64 pipe microbenchmarks: what DMMA.8x8x4 and DFMA can sustain on this
// device.  Used by bench.py to state the FP64 ceiling next to the cuBLAS DGEMM roof
// (MEASURED_PEAKS.json has no FP64 entry).
#include "kernels.cuh"

namespace cip {
namespace {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}

}  // namespace

int measure_fp64_peaks(double* dmma_tflops, double* dfma_tflops) {
  // Best of five short (~4 ms, burst clocks) launches of each loop.  A long (35 ms) register-only DMMA
  // loop was measured too: it settles at ~29.7 TFLOP/s under the power cap and leaves the chip throttled
  // for the next measurement, so only the burst figure (~37.0 TFLOP/s) -- the pipe ceiling a kernel can
  // be compared with -- is taken here.
  double* d = nullptr;
  CIP_CUDA(cudaMalloc(&d, 8));
  cudaEvent_t e0, e1;
  CIP_CUDA(cudaEventCreate(&e0));
  CIP_CUDA(cudaEventCreate(&e1));
  const int blocks = sm_count() * 4;
  double best_mma = 0.0, best_fma = 0.0;
  for (int iters : {4096}) {
    for (int rep = 0; rep < 5; ++rep) {
      float ms = 0;
      CIP_CUDA(cudaEventRecord(e0));
      dmma_peak_kernel<<<blocks, 256>>>(d, iters);
      CIP_CHECK_LAUNCH();
      CIP_CUDA(cudaEventRecord(e1));
      CIP_CUDA(cudaEventSynchronize(e1));
      CIP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      // per warp-instruction: 8*8*4 FMA = 512 flop
      best_mma = fmax(best_mma, (double)blocks * 8 * (double)iters * 16 * 512.0 / (ms * 1e-3) / 1e12);
      CIP_CUDA(cudaEventRecord(e0));
      dfma_peak_kernel<<<blocks, 256>>>(d, iters);
      CIP_CHECK_LAUNCH();
      CIP_CUDA(cudaEventRecord(e1));
      CIP_CUDA(cudaEventSynchronize(e1));
      CIP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      best_fma = fmax(best_fma, (double)blocks * 256 * (double)iters * 16 * 2.0 / (ms * 1e-3) / 1e12);
    }
  }
  *dmma_tflops = best_mma;
  *dfma_tflops = best_fma;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return 0;
}

}  // namespace cip

// Shared helpers for the conicip_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

namespace cip {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;   // kernels launched by this library (cip_stats.kernel_launches)
// Number of SMs of the current device (cached per device; grids are sized from it, never from a literal).
int sm_count();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: `done` is a per-kernel bit mask over device
// ordinals, so a process that drives several GPUs (single-process multi-GPU handles) sets it once on each.
int ensure_dyn_smem(const void* func, int bytes, std::atomic<unsigned long long>* done);

#define CIP_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      cip::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -1;                                                                    \
    }                                                                               \
  } while (0)

#define CIP_CHECK_LAUNCH()                                                          \
  do {                                                                              \
    cip::g_launches++;                                                              \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      cip::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return -1;                                                                    \
    }                                                                               \
  } while (0)

#define CIP_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != 0) return _r;      \
  } while (0)

// ---------------------------------------------------------------- Q4 layout
// A logical R x K matrix X (R "rows" = the non-contracted index, K = the
// contracted index) is stored quad-interleaved:
//     X[r, k]  at  ((k >> 2) * ld + r) * 4 + (k & 3)         (ld >= R, ld % 128 == 0)
// so that (a) the 8x4 DMMA operand fragment {rows g..g+7, k = 4q..4q+3} is one
// contiguous 256-byte run (conflict-free LDS.64, TMA-friendly 2 KB rows) and
// (b) every sub-panel (row range, k range) of a matrix is directly a K-major
// GEMM operand.  The symmetric matrix H / its Cholesky factor, A' (rows = columns
// of A, k = rows of A), G (rows = rows of G, k = columns) all live in this layout.
__host__ __device__ __forceinline__ size_t q4_index(int r, int k, int ld) {
  return ((size_t)(k >> 2) * (size_t)ld + (size_t)r) * 4 + (size_t)(k & 3);
}

constexpr int TILE = 128;  // CTA tile edge, Cholesky panel width, padding quantum

inline int round_up(int x, int q) { return (x + q - 1) / q * q; }

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 2-D tiled TMA load global -> shared, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// FP64 tensor-core MMA, D(8x8) += A(8x4,row) * B(4x8,col)   (SASS: DMMA.8x8x4)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace cip

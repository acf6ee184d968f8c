// K2/K3: blocked right-looking Cholesky of a Q4-layout SPD matrix and the triangular sweeps.
//   per 128-wide panel:  potrf_diag (one CTA: L11 and inv(L11) in shared memory)
//                        L21  = A21 * inv(L11)'      -> gemm_nt (DMMA tiles)
//                        A22 -= L21 * L21'           -> gemm_nt (lower, same tiles as the SYRK)
// Replaces LAPACK qr/lu at src/kktsolvers.jl:35,:295 and the solves at :39-48,:299.
#include "kernels.cuh"

namespace cip {

namespace {
constexpr int NB = 128;
constexpr int SLD = 132;  // padded row stride of the shared 128x128 block

__global__ void __launch_bounds__(512, 1)
potrf_diag_kernel(double* __restrict__ H, int ld, int j0, double* __restrict__ W, int* info) {
  extern __shared__ double S[];  // S[r * SLD + c]
  const int tid = threadIdx.x;
  {
    const int r = tid & 127;
    for (int q = tid >> 7; q < 32; q += 4) {
      const double2* p = reinterpret_cast<const double2*>(H + ((size_t)(j0 / 4 + q) * ld + j0 + r) * 4);
      const double2 v0 = p[0], v1 = p[1];
      double* d = S + r * SLD + 4 * q;
      d[0] = v0.x; d[1] = v0.y; d[2] = v1.x; d[3] = v1.y;
    }
  }
  __syncthreads();

  // ---- unblocked right-looking Cholesky (lower)
  {
    const int c = tid & 127, rg = tid >> 7;
    for (int k = 0; k < NB; ++k) {
      double d = S[k * SLD + k];
      if (!(d > 0.0)) {
        if (tid == 0) atomicCAS(info, 0, j0 + k + 1);
        d = 1.0;
      }
      const double lkk = sqrt(d);
      if (tid > k && tid < NB) S[tid * SLD + k] = S[tid * SLD + k] / lkk;
      __syncthreads();
      if (tid == k) S[k * SLD + k] = lkk;
      if (c > k) {
        const double lck = S[c * SLD + k];
        int r = c + ((rg - c) & 3);          // first row >= c with r % 4 == rg
        for (; r < NB; r += 4) S[r * SLD + c] -= S[r * SLD + k] * lck;
      }
      __syncthreads();
    }
  }
  // zero the strict upper triangle, write L back
  {
    const int r = tid & 127;
    for (int q = tid >> 7; q < 32; q += 4) {
      double* d = S + r * SLD + 4 * q;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (4 * q + t > r) d[t] = 0.0;
      double2* p = reinterpret_cast<double2*>(H + ((size_t)(j0 / 4 + q) * ld + j0 + r) * 4);
      p[0] = make_double2(d[0], d[1]);
      p[1] = make_double2(d[2], d[3]);
    }
  }
  __syncthreads();

  // ---- in-place inverse of the lower-triangular factor, row by row
  {
    const int j = tid >> 2, part = tid & 3;
    for (int i = 0; i < NB; ++i) {
      double sum = 0.0;
      if (j < i) {
        const double* Li = S + i * SLD;
        for (int k = j + part; k < i; k += 4) sum += Li[k] * S[k * SLD + j];
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const double lii = S[i * SLD + i];
      __syncthreads();
      if (part == 0) {
        if (j < i) S[i * SLD + j] = -sum / lii;
        else if (j == i) S[i * SLD + i] = 1.0 / lii;
      }
      __syncthreads();
    }
  }
  {
    const int r = tid & 127;
    for (int q = tid >> 7; q < 32; q += 4) {
      const double* d = S + r * SLD + 4 * q;
      double2* p = reinterpret_cast<double2*>(W + ((size_t)q * NB + r) * 4);
      p[0] = make_double2(d[0], d[1]);
      p[1] = make_double2(d[2], d[3]);
    }
  }
}

// forward sweep step for panel jb:  y_j = inv(L_jj) b_j ;  b_i -= L_ij y_j  (i > j)
__global__ void __launch_bounds__(128)
trsv_fwd_kernel(const double* __restrict__ H, int ld, const double* __restrict__ Winv, int jb,
                double* __restrict__ b, double* __restrict__ y) {
  __shared__ double bj[NB], yj[NB];
  const int tid = threadIdx.x, j0 = jb * NB;
  bj[tid] = b[j0 + tid];
  __syncthreads();
  const double* W = Winv + (size_t)jb * NB * NB;
  double acc = 0.0;
#pragma unroll 4
  for (int q = 0; q < 32; ++q) {
    const double2* p = reinterpret_cast<const double2*>(W + ((size_t)q * NB + tid) * 4);
    const double2 w0 = __ldg(p), w1 = __ldg(p + 1);
    acc = fma(w0.x, bj[4 * q], acc); acc = fma(w0.y, bj[4 * q + 1], acc);
    acc = fma(w1.x, bj[4 * q + 2], acc); acc = fma(w1.y, bj[4 * q + 3], acc);
  }
  yj[tid] = acc;
  __syncthreads();
  if (blockIdx.x == 0) {
    y[j0 + tid] = acc;
    return;
  }
  const int row = (jb + blockIdx.x) * NB + tid;
  double s = 0.0;
#pragma unroll 4
  for (int q = 0; q < 32; ++q) {
    const double2* p = reinterpret_cast<const double2*>(H + ((size_t)(j0 / 4 + q) * ld + row) * 4);
    const double2 l0 = __ldg(p), l1 = __ldg(p + 1);
    s = fma(l0.x, yj[4 * q], s); s = fma(l0.y, yj[4 * q + 1], s);
    s = fma(l1.x, yj[4 * q + 2], s); s = fma(l1.y, yj[4 * q + 3], s);
  }
  b[row] -= s;
}

// backward sweep step for panel jb:  x_j = inv(L_jj)' y_j ;  y_i -= L_ji' x_j  (i < j)
__global__ void __launch_bounds__(128)
trsv_bwd_kernel(const double* __restrict__ H, int ld, const double* __restrict__ Winv, int jb,
                double* __restrict__ y, double* __restrict__ x) {
  __shared__ double yj[NB], xj[NB];
  const int tid = threadIdx.x, j0 = jb * NB;
  const int q = tid >> 2, part = tid & 3;
  yj[tid] = y[j0 + tid];
  __syncthreads();
  {
    const double* W = Winv + (size_t)jb * NB * NB;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int r = part; r < NB; r += 4) {
      const double2* p = reinterpret_cast<const double2*>(W + ((size_t)q * NB + r) * 4);
      const double2 w0 = __ldg(p), w1 = __ldg(p + 1);
      const double yr = yj[r];
      a0 = fma(w0.x, yr, a0); a1 = fma(w0.y, yr, a1); a2 = fma(w1.x, yr, a2); a3 = fma(w1.y, yr, a3);
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
    }
    if (part == 0) { xj[4 * q] = a0; xj[4 * q + 1] = a1; xj[4 * q + 2] = a2; xj[4 * q + 3] = a3; }
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    x[j0 + tid] = xj[tid];
    return;
  }
  const int ib = blockIdx.x - 1;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int r = part; r < NB; r += 4) {
    const double2* p = reinterpret_cast<const double2*>(H + ((size_t)(ib * 32 + q) * ld + j0 + r) * 4);
    const double2 l0 = __ldg(p), l1 = __ldg(p + 1);
    const double xr = xj[r];
    a0 = fma(l0.x, xr, a0); a1 = fma(l0.y, xr, a1); a2 = fma(l1.x, xr, a2); a3 = fma(l1.y, xr, a3);
  }
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
  }
  if (part == 0) {
    double* yo = y + ib * NB + 4 * q;
    yo[0] -= a0; yo[1] -= a1; yo[2] -= a2; yo[3] -= a3;
  }
}

bool g_potrf_attr = false;
}  // namespace

int chol_make_plan(CholPlan* p, double* H, int n_pad, double* Winv, int* info) {
  p->H = H;
  p->ld = n_pad;
  p->npanels = n_pad / NB;
  p->Winv = Winv;
  p->info = info;
  CIP_TRY(make_q4_tensor_map(&p->mapH.map, H, n_pad, n_pad / 4));
  CIP_TRY(make_q4_tensor_map(&p->mapWinv.map, Winv, NB, 32LL * p->npanels));
  return 0;
}

int chol_factor(const CholPlan& p, cudaStream_t s) {
  const int smem = NB * SLD * (int)sizeof(double);
  if (!g_potrf_attr) {
    CIP_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    g_potrf_attr = true;
  }
  CIP_CUDA(cudaMemsetAsync(p.info, 0, sizeof(int), s));
  for (int jb = 0; jb < p.npanels; ++jb) {
    const int j0 = jb * NB;
    potrf_diag_kernel<<<1, 512, smem, s>>>(p.H, p.ld, j0, p.Winv + (size_t)jb * NB * NB, p.info);
    CIP_CHECK_LAUNCH();
    const int rem = p.npanels - jb - 1;
    if (rem == 0) break;
    GemmArgs t{};
    t.lower = 0; t.ntm = rem; t.ntn = 1; t.sym = 0;
    t.x_row0 = j0 + NB; t.y_row0 = 0; t.x_kq0 = j0 / 4; t.y_kq0 = 32 * jb; t.nk = NB / 32;
    t.Cin = nullptr; t.Cout = p.H; t.ldc = p.ld; t.c_row0 = j0 + NB; t.c_col0 = j0; t.alpha = 1.0;
    CIP_TRY(launch_gemm_nt(p.mapH, p.mapWinv, t, s));
    GemmArgs u{};
    u.lower = 1; u.ntm = rem; u.ntn = rem; u.sym = 1;
    u.x_row0 = j0 + NB; u.y_row0 = j0 + NB; u.x_kq0 = j0 / 4; u.y_kq0 = j0 / 4; u.nk = NB / 32;
    u.Cin = p.H; u.Cout = p.H; u.ldc = p.ld; u.c_row0 = j0 + NB; u.c_col0 = j0 + NB; u.alpha = -1.0;
    CIP_TRY(launch_gemm_nt(p.mapH, p.mapH, u, s));
  }
  return 0;
}

int chol_fwd(const CholPlan& p, double* b, double* y, cudaStream_t s) {
  for (int jb = 0; jb < p.npanels; ++jb) {
    trsv_fwd_kernel<<<p.npanels - jb, NB, 0, s>>>(p.H, p.ld, p.Winv, jb, b, y);
    CIP_CHECK_LAUNCH();
  }
  return 0;
}

int chol_bwd(const CholPlan& p, double* y, double* x, cudaStream_t s) {
  for (int jb = p.npanels - 1; jb >= 0; --jb) {
    trsv_bwd_kernel<<<jb + 1, NB, 0, s>>>(p.H, p.ld, p.Winv, jb, y, x);
    CIP_CHECK_LAUNCH();
  }
  return 0;
}

}  // namespace cip

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
  extern __shared__ double S[];        // S[r * SLD + c]; lower triangle: factor, upper triangle: inverse^T
  double* dinv = S + NB * SLD;         // [NB] 1/d_k during the factorisation, then 1/L_kk
  double* rsq = dinv + NB;             // [NB] 1/sqrt(d_k)
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

  // ---- right-looking Cholesky with deferred column scaling: column k keeps L[:,k]*L[k,k] until the
  //      end, the trailing update uses S[r][k]*S[c][k]/d_k, so each step needs a single barrier.
  {
    const int c = tid & 127, rg = tid >> 7;
    for (int k = 0; k < NB - 1; ++k) {
      double d = S[k * SLD + k];
      if (!(d > 0.0)) {
        if (tid == 0) atomicCAS(info, 0, j0 + k + 1);
        d = 1.0;
      }
      const double id = 1.0 / d;
      if (c > k) {
        const double t = S[c * SLD + k] * id;
        int r = c + ((rg - c) & 3);          // first row >= c with r % 4 == rg
#pragma unroll 4
        for (; r < NB; r += 4) S[r * SLD + c] = fma(-S[r * SLD + k], t, S[r * SLD + c]);
      }
      __syncthreads();
    }
    if (tid < NB) {
      double d = S[tid * SLD + tid];
      if (!(d > 0.0)) {
        atomicCAS(info, 0, j0 + tid + 1);
        d = 1.0;
      }
      const double l = sqrt(d);
      rsq[tid] = 1.0 / l;
      dinv[tid] = 1.0 / l;                    // 1 / L_kk
    }
    __syncthreads();
  }
  // scale the columns (L[r][k] = S[r][k] / sqrt(d_k)), zero the strict upper triangle, write L back
  {
    const int r = tid & 127;
    for (int q = tid >> 7; q < 32; q += 4) {
      double* d = S + r * SLD + 4 * q;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int cc = 4 * q + t;
        d[t] = (cc > r) ? 0.0 : ((cc == r) ? 1.0 / rsq[cc] : d[t] * rsq[cc]);
      }
      double2* p = reinterpret_cast<double2*>(H + ((size_t)(j0 / 4 + q) * ld + j0 + r) * 4);
      p[0] = make_double2(d[0], d[1]);
      p[1] = make_double2(d[2], d[3]);
    }
  }
  __syncthreads();

  // ---- inverse X = L^-1, column by column: a 4-lane group owns column j and walks down the rows
  //      with no block barrier; X[i][j] is kept transposed in the (free) upper triangle S[j][i].
  {
    const int j = tid >> 2, part = tid & 3;
    const unsigned gmask = 0xFu << ((tid & 31) & ~3);
    double* Xj = S + j * SLD;                 // Xj[i] = X[i][j]  (i > j), diagonal in xjj
    const double xjj = dinv[j];
    for (int i = j + 1; i < NB; ++i) {
      const double* Li = S + i * SLD;
      double sum = (part == 0) ? Li[j] * xjj : 0.0;
      for (int k = j + 1 + part; k < i; k += 4) sum = fma(Li[k], Xj[k], sum);
      sum += __shfl_xor_sync(gmask, sum, 1);
      sum += __shfl_xor_sync(gmask, sum, 2);
      if (part == 0) Xj[i] = -sum * dinv[i];
      __syncwarp(gmask);
    }
  }
  __syncthreads();
  {
    // W[r][c] = X[r][c] = (c < r) ? S[c][r] : (c == r ? dinv[r] : 0)
    const int r = tid & 127;
    for (int q = tid >> 7; q < 32; q += 4) {
      double o[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int cc = 4 * q + t;
        o[t] = (cc < r) ? S[cc * SLD + r] : ((cc == r) ? dinv[r] : 0.0);
      }
      double2* p = reinterpret_cast<double2*>(W + ((size_t)q * NB + r) * 4);
      p[0] = make_double2(o[0], o[1]);
      p[1] = make_double2(o[2], o[3]);
    }
  }
}

// forward sweep step for panel jb:  y_j = inv(L_jj) b_j ;  b_i -= L_ij y_j  (i > j)
// 512 threads: thread (row r, part) handles 8 of the 32 k-quads so that all 16 loads are in flight.
__global__ void __launch_bounds__(512)
trsv_fwd_kernel(const double* __restrict__ H, int ld, const double* __restrict__ Winv, int jb,
                double* __restrict__ b, double* __restrict__ y) {
  __shared__ double bj[NB], yj[NB], red[4][NB];
  const int tid = threadIdx.x, j0 = jb * NB;
  const int r = tid & 127, part = tid >> 7;
  if (tid < NB) bj[tid] = b[j0 + tid];
  __syncthreads();
  {
    const double* W = Winv + (size_t)jb * NB * NB;
    double2 w[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double2* p = reinterpret_cast<const double2*>(W + ((size_t)(part * 8 + i) * NB + r) * 4);
      w[2 * i] = __ldg(p);
      w[2 * i + 1] = __ldg(p + 1);
    }
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double* bq = bj + 4 * (part * 8 + i);
      acc = fma(w[2 * i].x, bq[0], acc); acc = fma(w[2 * i].y, bq[1], acc);
      acc = fma(w[2 * i + 1].x, bq[2], acc); acc = fma(w[2 * i + 1].y, bq[3], acc);
    }
    red[part][r] = acc;
  }
  __syncthreads();
  if (tid < NB) yj[tid] = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
  __syncthreads();
  if (blockIdx.x == 0) {
    if (tid < NB) y[j0 + tid] = yj[tid];
    return;
  }
  const int row = (jb + blockIdx.x) * NB + r;
  {
    double2 l[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double2* p = reinterpret_cast<const double2*>(H + ((size_t)(j0 / 4 + part * 8 + i) * ld + row) * 4);
      l[2 * i] = __ldg(p);
      l[2 * i + 1] = __ldg(p + 1);
    }
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double* yq = yj + 4 * (part * 8 + i);
      acc = fma(l[2 * i].x, yq[0], acc); acc = fma(l[2 * i].y, yq[1], acc);
      acc = fma(l[2 * i + 1].x, yq[2], acc); acc = fma(l[2 * i + 1].y, yq[3], acc);
    }
    __syncthreads();
    red[part][r] = acc;
  }
  __syncthreads();
  if (tid < NB) b[row] -= (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
}

// backward sweep step for panel jb:  x_j = inv(L_jj)' y_j ;  y_i -= L_ji' x_j  (i < j)
// 512 threads: 16 lanes share one k-quad (4 output columns) and split the 128 rows.
__device__ __forceinline__ void colsum4(const double* __restrict__ base, const double* __restrict__ vec, int part,
                                        double& a0, double& a1, double& a2, double& a3) {
  double2 m[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double2* p = reinterpret_cast<const double2*>(base + (size_t)(part + 16 * i) * 4);
    m[2 * i] = __ldg(p);
    m[2 * i + 1] = __ldg(p + 1);
  }
  a0 = a1 = a2 = a3 = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double vr = vec[part + 16 * i];
    a0 = fma(m[2 * i].x, vr, a0); a1 = fma(m[2 * i].y, vr, a1);
    a2 = fma(m[2 * i + 1].x, vr, a2); a3 = fma(m[2 * i + 1].y, vr, a3);
  }
#pragma unroll
  for (int o = 1; o <= 8; o <<= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
  }
}

__global__ void __launch_bounds__(512)
trsv_bwd_kernel(const double* __restrict__ H, int ld, const double* __restrict__ Winv, int jb,
                double* __restrict__ y, double* __restrict__ x) {
  __shared__ double yj[NB], xj[NB];
  const int tid = threadIdx.x, j0 = jb * NB;
  const int q = tid >> 4, part = tid & 15;
  if (tid < NB) yj[tid] = y[j0 + tid];
  __syncthreads();
  double a0, a1, a2, a3;
  colsum4(Winv + (size_t)jb * NB * NB + (size_t)q * NB * 4, yj, part, a0, a1, a2, a3);
  if (part == 0) { xj[4 * q] = a0; xj[4 * q + 1] = a1; xj[4 * q + 2] = a2; xj[4 * q + 3] = a3; }
  __syncthreads();
  if (blockIdx.x == 0) {
    if (tid < NB) x[j0 + tid] = xj[tid];
    return;
  }
  const int ib = blockIdx.x - 1;
  colsum4(H + ((size_t)(ib * 32 + q) * ld + j0) * 4, xj, part, a0, a1, a2, a3);
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
  int lo = 0, hi = 0;
  CIP_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CIP_CUDA(cudaStreamCreateWithPriority(&p->sc, cudaStreamNonBlocking, hi));
  for (auto* e : {&p->evT[0], &p->evT[1], &p->evR[0], &p->evR[1], &p->evS})
    CIP_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  return 0;
}

void chol_free_plan(CholPlan* p) {
  if (p->sc) cudaStreamDestroy(p->sc);
  for (auto e : {p->evT[0], p->evT[1], p->evR[0], p->evR[1], p->evS})
    if (e) cudaEventDestroy(e);
  p->sc = nullptr;
}

// Two-level blocked right-looking Cholesky with look-ahead.
//   outer panel = OUTER inner panels (512 columns).  Inside an outer panel every 128-wide inner
//   panel is factored (potrf_diag), solved against all rows below (TRSM as a GEMM with inv(L11))
//   and applied only to the remaining columns of the outer panel.  The trailing matrix then gets
//   ONE update with K = 512, which amortises the C-tile read/write and the pipeline fill that
//   dominate K = 128 tiles.  The block column of the next outer panel is updated first on the
//   high-priority stream `sc`, so its factorisation overlaps the bulk update running on `s`.
int chol_factor(const CholPlan& p, cudaStream_t s) {
  constexpr int OUTER = 4;
  const int smem = (NB * SLD + 2 * NB) * (int)sizeof(double);
  if (!g_potrf_attr) {
    CIP_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    g_potrf_attr = true;
  }
  cudaStream_t sc = p.sc;
  const int np = p.npanels;
  CIP_CUDA(cudaMemsetAsync(p.info, 0, sizeof(int), s));
  CIP_CUDA(cudaEventRecord(p.evS, s));
  CIP_CUDA(cudaStreamWaitEvent(sc, p.evS, 0));
  int outer = 0;
  for (int J0 = 0; J0 < np; J0 += OUTER, ++outer) {
    const int J1 = (J0 + OUTER < np) ? J0 + OUTER : np;
    for (int jb = J0; jb < J1; ++jb) {
      const int j0 = jb * NB;
      potrf_diag_kernel<<<1, 512, smem, sc>>>(p.H, p.ld, j0, p.Winv + (size_t)jb * NB * NB, p.info);
      CIP_CHECK_LAUNCH();
      const int rem = np - jb - 1;
      if (rem == 0) break;
      GemmArgs t{};   // L21 = A21 * inv(L11)'
      t.lower = 0; t.ntm = rem; t.ntn = 1; t.sym = 0;
      t.x_row0 = j0 + NB; t.y_row0 = 0; t.x_kq0 = j0 / 4; t.y_kq0 = 32 * jb; t.nk = NB / 32;
      t.Cin = nullptr; t.Cout = p.H; t.ldc = p.ld; t.c_row0 = j0 + NB; t.c_col0 = j0; t.alpha = 1.0;
      CIP_TRY(launch_gemm_nt(p.mapH, p.mapWinv, t, sc));
      const int inner_cols = J1 - jb - 1;
      if (inner_cols > 0) {   // update the rest of this outer panel's columns (K = 128)
        GemmArgs c{};
        c.lower = 0; c.ntm = rem; c.ntn = inner_cols; c.sym = 0;
        c.x_row0 = j0 + NB; c.y_row0 = j0 + NB; c.x_kq0 = j0 / 4; c.y_kq0 = j0 / 4; c.nk = NB / 32;
        c.Cin = p.H; c.Cout = p.H; c.ldc = p.ld; c.c_row0 = j0 + NB; c.c_col0 = j0 + NB; c.alpha = -1.0;
        CIP_TRY(launch_gemm_nt(p.mapH, p.mapH, c, sc));
      }
    }
    CIP_CUDA(cudaEventRecord(p.evT[outer & 1], sc));
    const int rem = np - J1;
    if (rem <= 0) break;
    const int k0 = J0 * NB, kw = (J1 - J0) * NB, r0 = J1 * NB;
    // next outer panel's block column first (critical path); it touches tiles the previous bulk
    // update also wrote, so it has to wait for that one.
    if (outer > 0) CIP_CUDA(cudaStreamWaitEvent(sc, p.evR[(outer - 1) & 1], 0));
    const int next_cols = rem < OUTER ? rem : OUTER;
    GemmArgs c{};
    c.lower = 0; c.ntm = rem; c.ntn = next_cols; c.sym = 0;
    c.x_row0 = r0; c.y_row0 = r0; c.x_kq0 = k0 / 4; c.y_kq0 = k0 / 4; c.nk = kw / 32;
    c.Cin = p.H; c.Cout = p.H; c.ldc = p.ld; c.c_row0 = r0; c.c_col0 = r0; c.alpha = -1.0;
    CIP_TRY(launch_gemm_nt(p.mapH, p.mapH, c, sc));
    // bulk trailing update on the caller's stream (K = 512)
    CIP_CUDA(cudaStreamWaitEvent(s, p.evT[outer & 1], 0));
    if (rem > OUTER) {
      const int b0 = r0 + OUTER * NB;
      GemmArgs u{};
      u.lower = 1; u.ntm = rem - OUTER; u.ntn = rem - OUTER; u.sym = 1;
      u.x_row0 = b0; u.y_row0 = b0; u.x_kq0 = k0 / 4; u.y_kq0 = k0 / 4; u.nk = kw / 32;
      u.Cin = p.H; u.Cout = p.H; u.ldc = p.ld; u.c_row0 = b0; u.c_col0 = b0; u.alpha = -1.0;
      CIP_TRY(launch_gemm_nt(p.mapH, p.mapH, u, s));
    }
    CIP_CUDA(cudaEventRecord(p.evR[outer & 1], s));
  }
  CIP_CUDA(cudaEventRecord(p.evS, sc));
  CIP_CUDA(cudaStreamWaitEvent(s, p.evS, 0));
  return 0;
}

int chol_fwd(const CholPlan& p, double* b, double* y, cudaStream_t s) {
  for (int jb = 0; jb < p.npanels; ++jb) {
    trsv_fwd_kernel<<<p.npanels - jb, 512, 0, s>>>(p.H, p.ld, p.Winv, jb, b, y);
    CIP_CHECK_LAUNCH();
  }
  return 0;
}

int chol_bwd(const CholPlan& p, double* y, double* x, cudaStream_t s) {
  for (int jb = p.npanels - 1; jb >= 0; --jb) {
    trsv_bwd_kernel<<<jb + 1, 512, 0, s>>>(p.H, p.ld, p.Winv, jb, y, x);
    CIP_CHECK_LAUNCH();
  }
  return 0;
}

}  // namespace cip

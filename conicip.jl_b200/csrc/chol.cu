// K2/K3: blocked right-looking Cholesky of a Q4-layout SPD matrix and the triangular sweeps.
//   per 128-wide panel:  potrf_diag (one CTA: L11 and inv(L11) in shared memory)
//                        L21  = A21 * inv(L11)'      -> gemm_nt (DMMA tiles)
//                        A22 -= L21 * L21'           -> gemm_nt (lower, same tiles as the SYRK)
// Replaces LAPACK qr/lu at src/kktsolvers.jl:35,:295 and the solves at :39-48,:299.
#include <stdlib.h>

#include "kernels.cuh"
#include "nccl_dl.h"

namespace cip {

namespace {
constexpr int NB = 128;
constexpr int SLD = 129;  // padded (odd) row stride of the shared 128x128 block: column walks are conflict-free

__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));   // ~20 bits; two Newton steps -> full double
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  return fma(x, e, x);
}

// Cholesky factor L and inverse inv(L) of one 128x128 diagonal block, one CTA of 512 threads.
//   * four 32-column sub-panels; inside a sub-panel every thread keeps its 8 matrix elements in
//     registers and only the current column travels through shared memory (double-buffered, one
//     barrier per column); the pivot scaling is deferred (S[i][j] -= S[i][k] S[j][k] / d_k) so the
//     dependent chain per column is LDS -> rcp -> FMA.  Rows below the 32x32 diagonal block are
//     eliminated in the same sweep, i.e. the in-block TRSM comes for free.
//   * rank-32 update of the remaining columns from a transposed copy of the sub-panel (4x4 register
//     tiles, conflict-free 32-byte loads).
//   * inv(L): the four 32x32 diagonal blocks by 4-lane column groups, then the off-diagonal blocks
//     X_ij = -X_ii * sum_k L_ik X_kj by block distance.  X is kept transposed in the free upper
//     triangle of S.
constexpr int SB = 32;           // sub-panel width
constexpr int PLD = 132;         // leading dimension of the transposed sub-panel copy P[k][i]
constexpr int SWEEP_WARPS = 16;  // warps that run the column sweep of a sub-panel (8 was measured slower: 118 vs 112 us)
constexpr int SWEEP_ROWS = NB / SWEEP_WARPS;   // matrix elements per thread during the sweep

__global__ void __launch_bounds__(512, 1)
potrf_diag_kernel(double* __restrict__ H, int ld, int j0, double* __restrict__ W, int* info) {
  extern __shared__ double S[];                  // S[r * SLD + c]
  double* dinv = S + NB * SLD;                   // [NB]   1 / L_kk
  double* colk = dinv + NB;                      // [2][NB] travelling column (double-buffered)
  double* P = colk + 2 * NB;                     // [SB][PLD] transposed sub-panel; later block temporaries
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
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

  for (int c0 = 0; c0 < NB; c0 += SB) {
    // ---- sub-panel sweep (column j = c0 + lane, rows i_e = w + SWEEP_WARPS e): the per-column chain is
    //      update -> publish -> barrier -> pivot reciprocal, so elements per thread are kept small.
    if (w < SWEEP_WARPS) {
      const int j = c0 + lane;
      double a[SWEEP_ROWS];
#pragma unroll
      for (int e = 0; e < SWEEP_ROWS; ++e) {
        const int i = w + SWEEP_WARPS * e;
        a[e] = (i >= j) ? S[i * SLD + j] : 0.0;
      }
      double dj = 1.0;
      for (int k = 0; k < SB; ++k) {
        double* ck = colk + (k & 1) * NB;
        if (lane == k) {
#pragma unroll
          for (int e = 0; e < SWEEP_ROWS; ++e) ck[w + SWEEP_WARPS * e] = a[e];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(SWEEP_WARPS * 32) : "memory");
        double d = ck[c0 + k];
        const bool bad = !(d > 0.0);
        if (bad && tid == 0) atomicCAS(info, 0, j0 + c0 + k + 1);
        d = bad ? 1.0 : d;
        if (lane == k) dj = d;
        if (lane > k) {
          const double t = ck[j] * fast_rcp(d);
#pragma unroll
          for (int e = 0; e < SWEEP_ROWS; ++e) {
            const int i = w + SWEEP_WARPS * e;
            if (i >= j) a[e] = fma(-ck[i], t, a[e]);
          }
        }
      }
      // scale: L[i][j] = a / sqrt(d_j); publish to S and to the transposed copy P[lane][i]
      const double rs = 1.0 / sqrt(dj);
      if (w == (j & (SWEEP_WARPS - 1))) dinv[j] = rs;     // exactly one thread per column (the owner of (j,j))
#pragma unroll
      for (int e = 0; e < SWEEP_ROWS; ++e) {
        const int i = w + SWEEP_WARPS * e;
        double v = 0.0;
        if (i > j) v = a[e] * rs;
        else if (i == j) v = dj * rs;
        if (i >= c0) {
          S[i * SLD + j] = v;
          P[lane * PLD + i] = v;
        }
      }
    }
    __syncthreads();
    // ---- rank-32 update of the columns right of the sub-panel: 4x4 register tiles, lower part
    const int R = NB - c0 - SB;                    // remaining order
    if (R > 0) {
      const int nt4 = R / 4;
      const int ntile = nt4 * (nt4 + 1) / 2;
      for (int t = tid; t < ntile; t += 512) {
        int ti = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        while (ti * (ti + 1) / 2 > t) --ti;
        const int tj = t - ti * (ti + 1) / 2;
        const int i0 = c0 + SB + 4 * ti, jj0 = c0 + SB + 4 * tj;
        double acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
#pragma unroll 4
        for (int k = 0; k < SB; ++k) {
          const double2 p0 = *reinterpret_cast<const double2*>(P + k * PLD + i0);
          const double2 p1 = *reinterpret_cast<const double2*>(P + k * PLD + i0 + 2);
          const double2 q0 = *reinterpret_cast<const double2*>(P + k * PLD + jj0);
          const double2 q1 = *reinterpret_cast<const double2*>(P + k * PLD + jj0 + 2);
          const double li[4] = {p0.x, p0.y, p1.x, p1.y}, lj[4] = {q0.x, q0.y, q1.x, q1.y};
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = fma(li[x], lj[y], acc[x][y]);
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y)
            if (i0 + x >= jj0 + y) S[(i0 + x) * SLD + jj0 + y] -= acc[x][y];
      }
    }
    __syncthreads();
  }

  // ---- write L back (strict upper triangle as zeros)
  {
    const int r = tid & 127;
    for (int q = tid >> 7; q < 32; q += 4) {
      double o[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) o[t] = (4 * q + t > r) ? 0.0 : S[r * SLD + 4 * q + t];
      double2* p = reinterpret_cast<double2*>(H + ((size_t)(j0 / 4 + q) * ld + j0 + r) * 4);
      p[0] = make_double2(o[0], o[1]);
      p[1] = make_double2(o[2], o[3]);
    }
  }
  __syncthreads();

  // ---- inverse.  X(r,c), r > c, lives at S[c][r]; X(r,r) = dinv[r].
#define XG(r, c) S[(c) * SLD + (r)]
  {
    // diagonal 32x32 blocks: 4 blocks x 32 columns x 4 lanes = 512 threads, no block barrier.  All
    // groups of a warp walk the same row index i (columns that start later are predicated off), so
    // the warp stays convergent instead of serialising eight 4-lane paths.
    const int blk = tid >> 7, jl = (tid & 127) >> 2, part = tid & 3;
    const int b0 = blk * SB, jc = b0 + jl;
    const double xjj = dinv[jc];
    for (int i = b0 + 1; i < b0 + SB; ++i) {
      const bool active = i > jc;
      const double* Li = S + i * SLD;
      double s0 = 0.0, s1 = 0.0;
      if (active) {
        if (part == 0) s0 = Li[jc] * xjj;
        int k = jc + 1 + part;
        for (; k + 4 < i; k += 8) {
          s0 = fma(Li[k], XG(k, jc), s0);
          s1 = fma(Li[k + 4], XG(k + 4, jc), s1);
        }
        if (k < i) s0 = fma(Li[k], XG(k, jc), s0);
      }
      double sum = s0 + s1;
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      if (active && part == 0) XG(i, jc) = -sum * dinv[i];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int dlt = 1; dlt < NB / SB; ++dlt) {
    const int nblk = NB / SB - dlt;
    // T_b = sum_{kk} L(ib, kk) * X(kk, jb)   for block pairs (ib, jb) = (b + dlt, b)
    for (int o = tid; o < nblk * SB * SB; o += 512) {
      const int b = o / (SB * SB), r = (o / SB) % SB, c = o % SB;
      const int ri = (b + dlt) * SB + r, cj = b * SB + c;
      const double* Li = S + ri * SLD;
      // four independent partial sums: the dependent-FMA chain, not the loads, bounds this loop
      double s0 = Li[cj] * dinv[cj], s1 = 0.0, s2 = 0.0, s3 = 0.0;
      const double* Xc = S + cj * SLD;                      // XG(kk, cj) = Xc[kk]
      const int kend = (b + dlt) * SB;
      int kk = cj + 1;
      for (; kk + 3 < kend; kk += 4) {
        s0 = fma(Li[kk], Xc[kk], s0);
        s1 = fma(Li[kk + 1], Xc[kk + 1], s1);
        s2 = fma(Li[kk + 2], Xc[kk + 2], s2);
        s3 = fma(Li[kk + 3], Xc[kk + 3], s3);
      }
      for (; kk < kend; ++kk) s0 = fma(Li[kk], Xc[kk], s0);
      P[(b * SB + r) * SB + c] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    // X_ib,jb = - X_ib,ib * T_b
    for (int o = tid; o < nblk * SB * SB; o += 512) {
      const int b = o / (SB * SB), r = (o / SB) % SB, c = o % SB;
      const int i0 = (b + dlt) * SB;
      double s0 = dinv[i0 + r] * P[(b * SB + r) * SB + c], s1 = 0.0, s2 = 0.0, s3 = 0.0;
      const double* Pc = P + (b * SB) * SB + c;             // Pc[k * SB] = T_b[k][c]
      int k = 0;
      for (; k + 3 < r; k += 4) {
        s0 = fma(XG(i0 + r, i0 + k), Pc[k * SB], s0);
        s1 = fma(XG(i0 + r, i0 + k + 1), Pc[(k + 1) * SB], s1);
        s2 = fma(XG(i0 + r, i0 + k + 2), Pc[(k + 2) * SB], s2);
        s3 = fma(XG(i0 + r, i0 + k + 3), Pc[(k + 3) * SB], s3);
      }
      for (; k < r; ++k) s0 = fma(XG(i0 + r, i0 + k), Pc[k * SB], s0);
      XG(i0 + r, b * SB + c) = -((s0 + s1) + (s2 + s3));
    }
    __syncthreads();
  }
  {
    // W[r][c] = X(r,c) = (c < r) ? S[c][r] : (c == r ? dinv[r] : 0)
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
#undef XG
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
  const int smem = (NB * SLD + 3 * NB + SB * PLD) * (int)sizeof(double);
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

// ---------------------------------------------------------------- distributed (block-cyclic) variant
namespace {
// inner factorisation of outer panel [J0, J1) on stream sc (the same sequence chol_factor uses)
int factor_outer_panel(const CholPlan& p, int J0, int J1, cudaStream_t sc, int smem) {
  const int np = p.npanels;
  for (int jb = J0; jb < J1; ++jb) {
    const int j0 = jb * NB;
    potrf_diag_kernel<<<1, 512, smem, sc>>>(p.H, p.ld, j0, p.Winv + (size_t)jb * NB * NB, p.info);
    CIP_CHECK_LAUNCH();
    const int rem = np - jb - 1;
    if (rem == 0) break;
    GemmArgs t{};
    t.lower = 0; t.ntm = rem; t.ntn = 1; t.sym = 0;
    t.x_row0 = j0 + NB; t.y_row0 = 0; t.x_kq0 = j0 / 4; t.y_kq0 = 32 * jb; t.nk = NB / 32;
    t.Cin = nullptr; t.Cout = p.H; t.ldc = p.ld; t.c_row0 = j0 + NB; t.c_col0 = j0; t.alpha = 1.0;
    CIP_TRY(launch_gemm_nt(p.mapH, p.mapWinv, t, sc));
    const int inner_cols = J1 - jb - 1;
    if (inner_cols > 0) {
      GemmArgs c{};
      c.lower = 0; c.ntm = rem; c.ntn = inner_cols; c.sym = 0;
      c.x_row0 = j0 + NB; c.y_row0 = j0 + NB; c.x_kq0 = j0 / 4; c.y_kq0 = j0 / 4; c.nk = NB / 32;
      c.Cin = p.H; c.Cout = p.H; c.ldc = p.ld; c.c_row0 = j0 + NB; c.c_col0 = j0 + NB; c.alpha = -1.0;
      CIP_TRY(launch_gemm_nt(p.mapH, p.mapH, c, sc));
    }
  }
  return 0;
}
// C[rows >= first row of outer panel Jc, columns of Jc] -= L[rows, panel J] * L[cols of Jc, panel J]'
int update_outer_panel(const CholPlan& p, int OUTER, int J, int Jc, cudaStream_t st) {
  const int np = p.npanels;
  const int J0 = J * OUTER, J1 = (J0 + OUTER < np) ? J0 + OUTER : np;
  const int C0 = Jc * OUTER, C1 = (C0 + OUTER < np) ? C0 + OUTER : np;
  GemmArgs c{};
  c.lower = 0; c.ntm = np - C0; c.ntn = C1 - C0; c.sym = 0;
  c.x_row0 = C0 * NB; c.y_row0 = C0 * NB; c.x_kq0 = J0 * NB / 4; c.y_kq0 = J0 * NB / 4; c.nk = (J1 - J0) * NB / 32;
  c.Cin = p.H; c.Cout = p.H; c.ldc = p.ld; c.c_row0 = C0 * NB; c.c_col0 = C0 * NB; c.alpha = -1.0;
  return launch_gemm_nt(p.mapH, p.mapH, c, st);
}
}  // namespace

int chol_factor_dist(const CholPlan& p, cudaStream_t s, const CholDist& d) {
  // outer panel width in 128-column panels (measured at C4 on 4 GPUs: 4 -> 39.3 ms, 2 -> 41.1 ms)
  int OUTER = 4;
  if (const char* env = getenv("CIP_DIST_OUTER")) { const int v = atoi(env); if (v >= 1 && v <= 8) OUTER = v; }
  const NcclApi* api = nccl_api();
  if (!api) return -1;
  const int smem = (NB * SLD + 3 * NB + SB * PLD) * (int)sizeof(double);
  if (!g_potrf_attr) {
    CIP_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    g_potrf_attr = true;
  }
  cudaStream_t sc = p.sc;
  const int np = p.npanels, NO = (np + OUTER - 1) / OUTER, N = d.nranks, me = d.rank;
  CIP_CUDA(cudaMemsetAsync(p.info, 0, sizeof(int), s));
  CIP_CUDA(cudaEventRecord(p.evS, s));
  CIP_CUDA(cudaStreamWaitEvent(sc, p.evS, 0));
  int myNext = me;                          // the next outer panel this rank will factor; its updates run on sc
  for (int J = 0; J < NO; ++J) {
    const int owner = J % N;
    const int J0 = J * OUTER, J1 = (J0 + OUTER < np) ? J0 + OUTER : np;
    if (owner == me) CIP_TRY(factor_outer_panel(p, J0, J1, sc, smem));
    // broadcast the factored block column (all rows: one contiguous Q4 region) and its inverse blocks
    {
      double* panel = p.H + (size_t)(J0 * NB / 4) * p.ld * 4;
      const size_t cnt = (size_t)(J1 - J0) * NB * p.ld;
      int r = api->Broadcast(panel, panel, cnt, kNcclFloat64, owner, d.comm, sc);
      if (r == 0) {
        double* w = p.Winv + (size_t)J0 * NB * NB;
        r = api->Broadcast(w, w, (size_t)(J1 - J0) * NB * NB, kNcclFloat64, owner, d.comm, sc);
      }
      if (r != 0) {
        set_error("ncclBroadcast failed in the distributed Cholesky: %s", api->GetErrorString(r));
        return -1;
      }
    }
    CIP_CUDA(cudaEventRecord(p.evT[J & 1], sc));
    if (owner == me) {
      myNext += N;
      // the new "next" panel has so far been updated on s: order sc after everything s has queued
      CIP_CUDA(cudaEventRecord(p.evR[0], s));
      CIP_CUDA(cudaStreamWaitEvent(sc, p.evR[0], 0));
    }
    if (myNext < NO && myNext > J) CIP_TRY(update_outer_panel(p, OUTER, J, myNext, sc));
    CIP_CUDA(cudaStreamWaitEvent(s, p.evT[J & 1], 0));
    for (int Jc = me; Jc < NO; Jc += N) {   // owned panels right of J, except the look-ahead one
      if (Jc <= J || Jc == myNext) continue;
      CIP_TRY(update_outer_panel(p, OUTER, J, Jc, s));
    }
  }
  CIP_CUDA(cudaEventRecord(p.evS, sc));
  CIP_CUDA(cudaStreamWaitEvent(s, p.evS, 0));
  // a failed pivot is only known to the owner of that panel
  int r = api->AllReduce(p.info, p.info, 1, kNcclInt32, kNcclMax, d.comm, s);
  if (r != 0) {
    set_error("ncclAllReduce(info) failed: %s", api->GetErrorString(r));
    return -1;
  }
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

// S (PSD) cone kernels: VecCongurance apply, nestod_sdc, maxstep_sdc, xsdc / dsdc and the scaled
// panel rows of S blocks.  One CTA per S cone; all k x k work (k <= 64) happens in shared memory.
// Replaces src/ConicIP.jl:35-40,69 (VecCongurance), :93-151 (mat/vecm), :196-210 (nestod_sdc),
// :272-303 (maxstep_sdc), :347-360 (dsdc!/xsdc!).  Dense factorizations that Julia takes from
// LAPACK (cholesky, svd, eigvals, lyap) are done here with an in-CTA Cholesky and a one-sided
// (Hestenes) Jacobi iteration, which has high relative accuracy for the SVD the NT scaling needs.
#include <math_constants.h>

#include <algorithm>
#include <stdlib.h>

#include "kernels.cuh"
#include "../../include/conicip_b200.h"

namespace cip {

namespace {

constexpr int KMAX = 64;               // up to this order the four k x k work matrices of a cone live in shared memory
constexpr int KCAP = 512;              // larger orders run the same code on a per-CTA workspace in global memory (L2)
constexpr int LDM = KMAX + 1;          // padded leading dimension of the shared k x k matrices
constexpr int MAT = KMAX * LDM;        // doubles per shared matrix
constexpr int NT = 256;                // threads per CTA
#define SQRT2 1.4142135623730951
#define M(A, i, j) (A)[(j) * ldm + (i)]   // column-major; `ldm` (LDM in shared memory, k + 1 in the workspace) is in scope everywhere

__device__ __forceinline__ int svec_index(int i, int j, int k) {   // i <= j, row-major upper triangle
  return i * k - (i * (i - 1)) / 2 + (j - i);
}

// X = mat(x)   (src/ConicIP.jl:93-119)
__device__ void load_mat(double* X, const double* __restrict__ x, int k, int ldm) {
  for (int e = threadIdx.x; e < k * k; e += NT) {
    const int i = e % k, j = e / k;
    const int a = i < j ? i : j, b = i < j ? j : i;
    const double v = x[svec_index(a, b, k)];
    M(X, i, j) = (i == j) ? v : v / SQRT2;
  }
}
// y = vecm(Y)  (src/ConicIP.jl:128-151); Y is symmetrised as (Y + Y')/2 to wash out rounding asymmetry
__device__ void store_vecm(double* __restrict__ y, const double* Y, int k, int ldm, bool accumulate = false) {
  for (int e = threadIdx.x; e < k * k; e += NT) {
    const int i = e % k, j = e / k;
    if (i <= j) {
      const double v = 0.5 * (M(Y, i, j) + M(Y, j, i));
      const double o = (i == j) ? v : v * SQRT2;
      const int idx = svec_index(i, j, k);
      y[idx] = accumulate ? y[idx] + o : o;
    }
  }
}
// C = op(A) * op(B), k x k, all in shared memory (C must not alias A or B)
template <bool TA, bool TB>
__device__ void matmul(double* C, const double* A, const double* B, int k, int ldm) {
  for (int e = threadIdx.x; e < k * k; e += NT) {
    const int i = e % k, j = e / k;
    double s = 0.0;
    for (int l = 0; l < k; ++l) s = fma(TA ? M(A, l, i) : M(A, i, l), TB ? M(B, j, l) : M(B, l, j), s);
    M(C, i, j) = s;
  }
}
// In-place lower Cholesky of a symmetric matrix; strict upper triangle zeroed.  Returns false in
// *ok (shared) when a pivot is not positive.
__device__ void cholesky(double* A, int k, int* ok, int ldm) {
  if (threadIdx.x == 0) *ok = 1;
  __syncthreads();
  for (int c = 0; c < k; ++c) {
    const double d = M(A, c, c);
    __syncthreads();
    if (!(d > 0.0)) {
      if (threadIdx.x == 0) *ok = 0;
      __syncthreads();
      return;
    }
    const double l = sqrt(d);
    for (int i = c + threadIdx.x; i < k; i += NT) M(A, i, c) = (i == c) ? l : M(A, i, c) / l;
    __syncthreads();
    for (int e = threadIdx.x; e < (k - c - 1) * (k - c - 1); e += NT) {
      const int i = c + 1 + e % (k - c - 1), j = c + 1 + e / (k - c - 1);
      if (i >= j) M(A, i, j) -= M(A, i, c) * M(A, j, c);
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < k * k; e += NT) {
    const int i = e % k, j = e / k;
    if (i < j) M(A, i, j) = 0.0;
  }
  __syncthreads();
}
// X = inv(L) for lower-triangular L (X lower-triangular, distinct buffer)
__device__ void tri_inverse(double* X, const double* L, int k, int ldm) {
  for (int e = threadIdx.x; e < k * k; e += NT) M(X, e % k, e / k) = 0.0;
  __syncthreads();
  // column j of X solves L x = e_j; one thread per column (k <= 64 columns, forward substitution)
  for (int j = threadIdx.x; j < k; j += NT) {
    for (int i = j; i < k; ++i) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int l = j; l < i; ++l) s -= M(L, i, l) * M(X, l, j);
      M(X, i, j) = s / M(L, i, i);
    }
  }
  __syncthreads();
}
// One-sided (Hestenes) Jacobi: rotates the columns of G (and of V, if non-null, starting from
// whatever V holds) until they are mutually orthogonal: G_out = G_in * J, V_out = V_in * J.
// Round-robin ordering; each of the k/2 disjoint pairs of a step is handled by 8 lanes.
__device__ void jacobi_onesided(double* G, double* V, int k, int* sh_flag, int ldm) {
  const int kk = (k + 1) & ~1;                 // even number of players (a phantom column if k is odd)
  const int pairs = kk / 2;
  const int lane8 = threadIdx.x & 7, grp = threadIdx.x >> 3;   // 32 groups of 8 lanes
  const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);   // groups of one warp may diverge
  for (int sweep = 0; sweep < 30; ++sweep) {
    if (threadIdx.x == 0) *sh_flag = 0;
    __syncthreads();
    for (int step = 0; step < kk - 1; ++step) {
      for (int pr = grp; pr < pairs; pr += NT / 8) {
        // round-robin tournament: player 0 fixed, others rotate
        int p = (pr == 0) ? 0 : 1 + (pr - 1 + step) % (kk - 1);
        int q = 1 + (kk - 2 - pr + step + (kk - 1)) % (kk - 1);
        if (p > q) { const int t = p; p = q; q = t; }
        const bool live = (q < k);
        double a = 0, b = 0, g = 0;
        if (live) {
          for (int i = lane8; i < k; i += 8) {
            const double gp = M(G, i, p), gq = M(G, i, q);
            a = fma(gp, gp, a); b = fma(gq, gq, b); g = fma(gp, gq, g);
          }
        }
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) {
          a += __shfl_xor_sync(gmask, a, o);
          b += __shfl_xor_sync(gmask, b, o);
          g += __shfl_xor_sync(gmask, g, o);
        }
        if (live && fabs(g) > 1e-16 * sqrt(a * b) && g != 0.0) {
          if (lane8 == 0) *sh_flag = 1;
          const double zeta = (b - a) / (2.0 * g);
          const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
          for (int i = lane8; i < k; i += 8) {
            const double gp = M(G, i, p), gq = M(G, i, q);
            M(G, i, p) = c * gp - s * gq;
            M(G, i, q) = s * gp + c * gq;
            if (V) {
              const double vp = M(V, i, p), vq = M(V, i, q);
              M(V, i, p) = c * vp - s * vq;
              M(V, i, q) = s * vp + c * vq;
            }
          }
        }
      }
      __syncthreads();
    }
    const int again = *sh_flag;
    __syncthreads();
    if (!again) break;
  }
}

struct SMem {
  double* A; double* B; double* C; double* D;   // four k x k matrices
  double* vec;                                  // k doubles
  int* flag;
  int ldm;
};
// k <= KMAX: everything in dynamic shared memory; above: this CTA's slice of the global workspace (the flag stays
// in shared memory).  `cta` = linear CTA index of the launch.
__device__ __forceinline__ SMem carve(double* base, int k, double* ws, long long ws_stride, int cta) {
  SMem s;
  s.flag = reinterpret_cast<int*>(base + 4 * MAT + KMAX);
  if (k <= KMAX) {
    s.ldm = LDM;
    s.A = base; s.B = base + MAT; s.C = base + 2 * MAT; s.D = base + 3 * MAT;
    s.vec = base + 4 * MAT;
  } else {
    s.ldm = k + 1;
    double* g = ws + (size_t)cta * (size_t)ws_stride;
    const size_t mat = (size_t)k * (k + 1);
    s.A = g; s.B = g + mat; s.C = g + 2 * mat; s.D = g + 3 * mat;
    s.vec = g + 4 * mat;
  }
  return s;
}
constexpr int SDP_SMEM = (4 * MAT + KMAX) * 8 + 16;

// signed eigen-decomposition of a symmetric matrix Msym (in s.A, destroyed): eigenvalues -> s.vec,
// eigenvectors -> s.B (if want_vectors).  Uses the shift M + c I (c = ||M||_F) so that the one-sided
// Jacobi sees a PSD matrix and singular values equal eigenvalues.
__device__ void sym_eigen(SMem s, int k, bool want_vectors) {
  const int ldm = s.ldm;
  double* A = s.A; double* V = s.B;
  double fro = 0.0;
  for (int e = threadIdx.x; e < k * k; e += NT) { const double v = M(A, e % k, e / k); fro = fma(v, v, fro); }
  fro = warp_sum(fro);
  __shared__ double red[NT / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = fro;
  __syncthreads();
  double c = 0.0;
  for (int i = 0; i < NT / 32; ++i) c += red[i];
  c = sqrt(c);
  for (int e = threadIdx.x; e < k * k; e += NT) {
    const int i = e % k, j = e / k;
    if (i == j) M(A, i, j) += c;
    M(V, i, j) = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  jacobi_onesided(A, V, k, s.flag, ldm);
  // columns of A are now (lambda_j + c) v_j: lambda_j = v_j . a_j - c
  for (int j = threadIdx.x; j < k; j += NT) {
    double d = 0.0;
    for (int i = 0; i < k; ++i) d = fma(M(V, i, j), M(A, i, j), d);
    s.vec[j] = d - c;
  }
  __syncthreads();
  (void)want_vectors;
}

// ------------------------------------------------------------------ kernels (grid = number of S cones)
struct SDesc {
  const int* slist;   // cone indices
  const int* off;     // cone row offsets
  const int* sord;    // order k per S cone (slist order)
  const int* roff;    // offset (doubles) into R / Ri per S cone
  double* ws;         // global workspace for orders above KMAX (nullptr if none), ws_stride doubles per CTA
  long long ws_stride;
};

// y_I = vecm(A' mat(x_I) A) with A = R, R', inv(R), inv(R)' selected by (use_inv, transpose)
__global__ void __launch_bounds__(NT)
sdp_apply_kernel(SDesc d, const int* __restrict__ kind, const double* __restrict__ R, const double* __restrict__ Ri,
                 int use_inv, int transpose, const double* __restrict__ x, double* __restrict__ y) {
  extern __shared__ double smem[];
  const int si = blockIdx.x, ci = d.slist[si];
  if (kind[ci] != CIP_BLK_VECCONG) return;
  const int k = d.sord[si], off = d.off[ci];
  SMem s = carve(smem, k, d.ws, d.ws_stride, blockIdx.y * gridDim.x + blockIdx.x);
  const int ldm = s.ldm;
  const double* src = (use_inv ? Ri : R) + d.roff[si];
  for (int e = threadIdx.x; e < k * k; e += NT) M(s.A, e % k, e / k) = src[e];
  load_mat(s.B, x + off, k, ldm);
  __syncthreads();
  if (!transpose) {
    matmul<false, false>(s.C, s.B, s.A, k, ldm);     // X A
    __syncthreads();
    matmul<true, false>(s.D, s.A, s.C, k, ldm);      // A' (X A)
  } else {
    matmul<false, true>(s.C, s.B, s.A, k, ldm);      // X A'
    __syncthreads();
    matmul<false, false>(s.D, s.A, s.C, k, ldm);     // A (X A')
  }
  __syncthreads();
  store_vecm(y + off, s.D, k, ldm);
}

// nestod_sdc (src/ConicIP.jl:196-210): R = inv(Lz)' U sqrt(Lambda), U Lambda V' = svd(Lz' Ls);
// also writes inv(R) = Lambda^-1/2 U' Lz' and lambda = F v = vecm(R' Z R).
__global__ void __launch_bounds__(NT)
sdp_nt_kernel(SDesc d, int* __restrict__ kindF, int* __restrict__ kindFi, double* __restrict__ R,
              double* __restrict__ Ri, const double* __restrict__ v, const double* __restrict__ sv,
              double* __restrict__ lambda, int* __restrict__ info) {
  extern __shared__ double smem[];
  const int si = blockIdx.x, ci = d.slist[si];
  const int k = d.sord[si], off = d.off[ci];
  SMem s = carve(smem, k, d.ws, d.ws_stride, blockIdx.y * gridDim.x + blockIdx.x);
  const int ldm = s.ldm;
  load_mat(s.A, sv + off, k, ldm);      // S
  load_mat(s.B, v + off, k, ldm);       // Z
  __syncthreads();
  cholesky(s.A, k, s.flag, ldm);        // Ls
  const int ok1 = *s.flag;
  __syncthreads();
  cholesky(s.B, k, s.flag, ldm);        // Lz
  const int ok2 = *s.flag;
  __syncthreads();
  if (!(ok1 && ok2)) {
    // mat(s) or mat(z) is not positive definite: the reference throws PosDefException from cholesky(mat(s))
    // here (src/ConicIP.jl:201-202).  Flag the cone (read back by cip_nt_scaling, which returns ci + 1), poison
    // lambda so that nothing downstream can use the half-factored matrices, and stop.
    if (threadIdx.x == 0) { atomicCAS(info, 0, ci + 1); kindF[ci] = CIP_BLK_VECCONG; kindFi[ci] = CIP_BLK_VECCONG; }
    const int dim = k * (k + 1) / 2;
    for (int e = threadIdx.x; e < dim; e += NT) lambda[off + e] = CUDART_NAN;
    return;
  }
  matmul<true, false>(s.C, s.B, s.A, k, ldm);   // G = Lz' Ls
  __syncthreads();
  jacobi_onesided(s.C, nullptr, k, s.flag, ldm);   // C = U Sigma (columns orthogonal)
  for (int j = threadIdx.x; j < k; j += NT) {
    double n2 = 0.0;
    for (int i = 0; i < k; ++i) n2 = fma(M(s.C, i, j), M(s.C, i, j), n2);
    s.vec[j] = sqrt(n2);                    // sigma_j
  }
  __syncthreads();
  tri_inverse(s.D, s.B, k, ldm);                 // D = inv(Lz)
  // R = inv(Lz)' * (C * Sigma^-1/2)   ;   inv(R) = Sigma^-3/2 C' Lz'
  for (int e = threadIdx.x; e < k * k; e += NT) { const int j = e / k; M(s.C, e % k, j) /= sqrt(s.vec[j]); }
  __syncthreads();
  matmul<true, false>(s.A, s.D, s.C, k, ldm);    // A = R
  __syncthreads();
  double* Rg = R + d.roff[si];
  double* Rig = Ri + d.roff[si];
  for (int e = threadIdx.x; e < k * k; e += NT) Rg[e] = M(s.A, e % k, e / k);
  // inv(R) = (C Sigma^-1/2)' Lz' / sigma  (C already scaled once by Sigma^-1/2)
  matmul<true, true>(s.D, s.C, s.B, k, ldm);     // D = C' Lz'
  __syncthreads();
  for (int e = threadIdx.x; e < k * k; e += NT) { const int i = e % k; Rig[e] = M(s.D, i, e / k) / s.vec[i]; }
  // lambda = vecm(R' Z R)
  load_mat(s.B, v + off, k, ldm);
  __syncthreads();
  matmul<false, false>(s.C, s.B, s.A, k, ldm);
  __syncthreads();
  matmul<true, false>(s.D, s.A, s.C, k, ldm);
  __syncthreads();
  store_vecm(lambda + off, s.D, k, ldm);
  if (threadIdx.x == 0) { kindF[ci] = CIP_BLK_VECCONG; kindFi[ci] = CIP_BLK_VECCONG; }
}

// inv(R) for user-supplied R (cip_factor / cip_set_scaling): Gauss-Jordan with partial pivoting
__global__ void __launch_bounds__(NT)
sdp_invert_kernel(SDesc d, const int* __restrict__ kind, const double* __restrict__ R, double* __restrict__ Ri) {
  extern __shared__ double smem[];
  const int si = blockIdx.x, ci = d.slist[si];
  if (kind[ci] != CIP_BLK_VECCONG) return;
  const int k = d.sord[si];
  SMem s = carve(smem, k, d.ws, d.ws_stride, blockIdx.y * gridDim.x + blockIdx.x);
  const int ldm = s.ldm;
  const double* Rg = R + d.roff[si];
  for (int e = threadIdx.x; e < k * k; e += NT) {
    const int i = e % k, j = e / k;
    M(s.A, i, j) = Rg[e];
    M(s.B, i, j) = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int c = 0; c < k; ++c) {
    if (threadIdx.x == 0) {
      int piv = c; double best = fabs(M(s.A, c, c));
      for (int i = c + 1; i < k; ++i) if (fabs(M(s.A, i, c)) > best) { best = fabs(M(s.A, i, c)); piv = i; }
      *s.flag = piv;
    }
    __syncthreads();
    const int piv = *s.flag;
    if (piv != c) {
      for (int j = threadIdx.x; j < k; j += NT) {
        double t = M(s.A, c, j); M(s.A, c, j) = M(s.A, piv, j); M(s.A, piv, j) = t;
        t = M(s.B, c, j); M(s.B, c, j) = M(s.B, piv, j); M(s.B, piv, j) = t;
      }
    }
    __syncthreads();
    const double pv = M(s.A, c, c);
    for (int i = threadIdx.x; i < k; i += NT) s.vec[i] = (i == c) ? 0.0 : M(s.A, i, c) / pv;
    __syncthreads();
    for (int e = threadIdx.x; e < k * k; e += NT) {
      const int i = e % k, j = e / k;
      const double f = s.vec[i];
      if (i != c) { M(s.A, i, j) -= f * M(s.A, c, j); M(s.B, i, j) -= f * M(s.B, c, j); }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += NT) { M(s.A, c, j) /= pv; M(s.B, c, j) /= pv; }
    __syncthreads();
  }
  double* Rig = Ri + d.roff[si];
  for (int e = threadIdx.x; e < k * k; e += NT) Rig[e] = M(s.B, e % k, e / k);
}

// o = vecm(XY + YX)  (xsdc!, :355-360)   /   o = vecm(O), Y O + O Y = X  (dsdc!, :347-353)
__global__ void __launch_bounds__(NT)
sdp_prod_div_kernel(SDesc d, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ o,
                    int divide) {
  extern __shared__ double smem[];
  const int si = blockIdx.x, ci = d.slist[si];
  const int k = d.sord[si], off = d.off[ci];
  SMem s = carve(smem, k, d.ws, d.ws_stride, blockIdx.y * gridDim.x + blockIdx.x);
  const int ldm = s.ldm;
  if (!divide) {
    load_mat(s.A, x + off, k, ldm);
    load_mat(s.B, y + off, k, ldm);
    __syncthreads();
    matmul<false, false>(s.C, s.A, s.B, k, ldm);
    __syncthreads();
    for (int e = threadIdx.x; e < k * k; e += NT) { const int i = e % k, j = e / k; M(s.D, i, j) = M(s.C, i, j) + M(s.C, j, i); }
    __syncthreads();
    store_vecm(o + off, s.D, k, ldm);
    return;
  }
  // Y = V diag(l) V';  T = V' X V;  T_ij /= (l_i + l_j);  O = V T V'
  load_mat(s.A, y + off, k, ldm);
  __syncthreads();
  sym_eigen(s, k, true);                       // eigenvalues s.vec, vectors s.B
  load_mat(s.A, x + off, k, ldm);
  __syncthreads();
  matmul<false, false>(s.C, s.A, s.B, k, ldm);      // X V
  __syncthreads();
  matmul<true, false>(s.D, s.B, s.C, k, ldm);       // V' X V
  __syncthreads();
  for (int e = threadIdx.x; e < k * k; e += NT) { const int i = e % k, j = e / k; M(s.D, i, j) /= (s.vec[i] + s.vec[j]); }
  __syncthreads();
  matmul<false, true>(s.C, s.D, s.B, k, ldm);       // T V'
  __syncthreads();
  matmul<false, false>(s.A, s.B, s.C, k, ldm);      // V T V'
  __syncthreads();
  store_vecm(o + off, s.A, k, ldm);
}

__device__ __forceinline__ unsigned long long dkey2(double x) {
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// maxstep_sdc (:272-303)
__global__ void __launch_bounds__(NT)
sdp_maxstep_kernel(SDesc d, const double* __restrict__ x, const double* __restrict__ dd, double d_scale,
                   unsigned long long* key) {
  extern __shared__ double smem[];
  const int si = blockIdx.x, ci = d.slist[si];
  const int k = d.sord[si], off = d.off[ci];
  SMem s = carve(smem, k, d.ws, d.ws_stride, blockIdx.y * gridDim.x + blockIdx.x);
  const int ldm = s.ldm;
  double res;
  if (!dd) {                                       // minimum eigenvalue of X
    load_mat(s.A, x + off, k, ldm);
    __syncthreads();
    sym_eigen(s, k, false);
    double mn = CUDART_INF;
    for (int j = 0; j < k; ++j) mn = fmin(mn, s.vec[j]);
    res = mn > 0 ? 0.0 : -1.0 + mn;
  } else {
    load_mat(s.C, x + off, k, ldm);
    __syncthreads();
    cholesky(s.C, k, s.flag, ldm);                      // X = L L'
    const int ok = *s.flag;
    __syncthreads();
    if (!ok) {
      res = CUDART_INF;                            // X not positive definite (:277-280)
    } else {
      tri_inverse(s.D, s.C, k, ldm);                    // inv(L)
      load_mat(s.A, dd + off, k, ldm);
      __syncthreads();
      for (int e = threadIdx.x; e < k * k; e += NT) M(s.A, e % k, e / k) /= d_scale;
      __syncthreads();
      matmul<false, true>(s.B, s.A, s.D, k, ldm);       // D inv(L)'
      __syncthreads();
      matmul<false, false>(s.A, s.D, s.B, k, ldm);      // inv(L) D inv(L)'  (similar to X^-1/2 D X^-1/2)
      __syncthreads();
      for (int e = threadIdx.x; e < k * k; e += NT) {   // symmetrise, :284
        const int i = e % k, j = e / k;
        if (i < j) { const double m = 0.5 * (M(s.A, i, j) + M(s.A, j, i)); M(s.A, i, j) = m; M(s.A, j, i) = m; }
      }
      __syncthreads();
      sym_eigen(s, k, false);
      double mx = -CUDART_INF;
      for (int j = 0; j < k; ++j) mx = fmax(mx, s.vec[j]);
      res = (mx < 0) ? CUDART_INF : 1.0 / mx;
    }
  }
  if (threadIdx.x == 0 && res < CUDART_INF) atomicMin(key, dkey2(res));
}

// Atil rows of an S block: column j of A restricted to the block, a_j -> vecm(inv(R) mat(a_j) inv(R)')
// grid (S cone, column chunk); each CTA keeps inv(R) resident and loops over its columns.
__global__ void __launch_bounds__(NT)
sdp_scale_panel_kernel(SDesc d, const int* __restrict__ kind, const double* __restrict__ Ri,
                       const double* __restrict__ At4, double* __restrict__ Atil4, int ld, int ncols,
                       int cols_per_cta, int skip_up_to) {
  extern __shared__ double smem[];
  const int si = blockIdx.x, ci = d.slist[si];
  if (kind[ci] != CIP_BLK_VECCONG) return;
  const int k = d.sord[si], off = d.off[ci];
  if (k <= skip_up_to) return;              // done by sdp_scale_panel_dmma_kernel
  SMem s = carve(smem, k, d.ws, d.ws_stride, blockIdx.y * gridDim.x + blockIdx.x);
  const int ldm = s.ldm;
  const int dim = k * (k + 1) / 2;
  const double* src = Ri + d.roff[si];
  for (int e = threadIdx.x; e < k * k; e += NT) M(s.A, e % k, e / k) = src[e];
  const int j0 = blockIdx.y * cols_per_cta;
  const int j1 = min(ncols, j0 + cols_per_cta);
  for (int j = j0; j < j1; ++j) {
    __syncthreads();
    for (int e = threadIdx.x; e < k * k; e += NT) {
      const int i = e % k, c = e / k;
      const int a = i < c ? i : c, b = i < c ? c : i;
      const double v = At4[q4_index(j, off + svec_index(a, b, k), ld)];
      M(s.B, i, c) = (i == c) ? v : v / SQRT2;
    }
    __syncthreads();
    matmul<false, true>(s.C, s.B, s.A, k, ldm);      // X inv(R)'
    __syncthreads();
    matmul<false, false>(s.D, s.A, s.C, k, ldm);     // inv(R) X inv(R)'
    __syncthreads();
    for (int e = threadIdx.x; e < k * k; e += NT) {
      const int i = e % k, c = e / k;
      if (i <= c) {
        const double v = 0.5 * (M(s.D, i, c) + M(s.D, c, i));
        Atil4[q4_index(j, off + svec_index(i, c, k), ld)] = (i == c) ? v : v * SQRT2;
      }
    }
  }
  (void)dim;
}


// ---- the same panel on the FP64 tensor cores (orders up to KMAX) -------------------------------------------------
// Per column of A the congruence inv(R) X inv(R)' is two k x k x k products; over the n columns they are two GEMMs
// with inv(R)' as the shared right operand, [X_1; X_2; ...] inv(R)' and (per column) inv(R) T_j.  A CTA keeps inv(R)
// resident in shared memory and walks over groups of PCOLS adjacent columns: the group's S rows are staged with
// coalesced 128-byte runs (in Q4 the PCOLS x 4 doubles of a quad row are contiguous), each column is expanded to the
// full symmetric matrix, multiplied twice on DMMA.8x8x4 -- warp w owns rows 8w .. 8w+7 of the product, all columns,
// 16 accumulators per lane -- and packed back into the staging slot, which is written out like it was read.
// Every k x k array has leading dimension PLD = 68 (= 4 mod 16): the fragment element of lane (g, t) sits at
// (4s + t) * PLD + row0 + g, sixteen distinct 8-byte banks per half warp, so the LDS.64 are conflict-free.
// The upper triangle of the product is stored as computed (the scalar kernel averages Y and Y'; the two differ by
// rounding only).  Replaces `F^-T * Matrix(A)` for VecCongurance blocks, src/kktsolvers.jl:33, src/ConicIP.jl:69.
constexpr int PLD = 68;
constexpr int PCOLS = 4;
constexpr int PANEL_SMEM = (3 * KMAX * PLD + PCOLS * (KMAX * (KMAX + 1) / 2 + 16)) * 8;
static_assert(PANEL_SMEM <= 227 * 1024, "panel kernel shared memory");

__global__ void __launch_bounds__(NT, 1)
sdp_scale_panel_dmma_kernel(SDesc d, const int* __restrict__ kind, const double* __restrict__ Ri,
                            const double* __restrict__ At4, double* __restrict__ Atil4, int ld, int ncols) {
  extern __shared__ double smem[];
  double* sR = smem;                       // inv(R)      [c * PLD + i] = inv(R)[i][c]
  double* sX = sR + KMAX * PLD;            // mat(a_j)    [l * PLD + i] = X[i][l]
  double* sT = sX + KMAX * PLD;            // X inv(R)'   [l * PLD + c] = T[l][c]
  double* stage = sT + KMAX * PLD;         // PCOLS packed columns, `sstr` doubles apart
  const int si = blockIdx.x, ci = d.slist[si];
  if (kind[ci] != CIP_BLK_VECCONG) return;
  const int k = d.sord[si];
  if (k > KMAX) return;                    // the workspace kernel takes those
  const int off = d.off[ci], dim = k * (k + 1) / 2;
  const int kp = (k + 7) & ~7;             // order padded to whole 8 x 8 tiles (zero rows / columns)
  // distance between the staged columns = 4 mod 16: the 4 columns x 4 rows of a quad row (16 consecutive threads of the
  // coalesced gather / scatter) then fall into 16 different 8-byte banks
  const int sstr = ((dim + 11) & ~15) + 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const double* src = Ri + d.roff[si];
  for (int e = tid; e < kp * kp; e += NT) {
    const int i = e % kp, c = e / kp;
    sR[c * PLD + i] = (i < k && c < k) ? src[c * k + i] : 0.0;
    sX[c * PLD + i] = 0.0;                 // the padding stays zero: the expansion below writes i, c < k only
  }
  const int groups = (ncols + PCOLS - 1) / PCOLS;
  const int q0 = off >> 2, nq = ((off + dim - 1) >> 2) - q0 + 1;     // quad rows that hold the block
  const bool active = warp * 8 < kp;
  for (int grp = blockIdx.y; grp < groups; grp += gridDim.y) {
    const int j0 = grp * PCOLS;
    __syncthreads();                       // the previous group has left the staging buffer
    for (int idx = tid; idx < nq * 16; idx += NT) {
      const int quad = q0 + (idx >> 4), c = (idx & 15) >> 2, r = idx & 3;
      const int e = quad * 4 + r - off;
      if (e >= 0 && e < dim) stage[c * sstr + e] = At4[((size_t)quad * ld + j0 + c) * 4 + r];   // (columns < ld always exist)
    }
    __syncthreads();
    for (int c = 0; c < PCOLS && j0 + c < ncols; ++c) {
      double* sv = stage + c * sstr;
      for (int e = tid; e < k * k; e += NT) {                       // X = mat(a_j)
        const int i = e % k, cc = e / k;
        const int a = i < cc ? i : cc, b = i < cc ? cc : i;
        const double v = sv[svec_index(a, b, k)];
        sX[cc * PLD + i] = (i == cc) ? v : v * (1.0 / SQRT2);
      }
      __syncthreads();
      if (active) {                                                 // T = X inv(R)'
        double acc[8][2];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
        for (int s4 = 0; s4 < kp / 4; ++s4) {
          const double a = sX[(4 * s4 + t) * PLD + 8 * warp + g];
#pragma unroll
          for (int nt = 0; nt < 8; ++nt)
            if (nt * 8 < kp) dmma884(acc[nt][0], acc[nt][1], a, sR[(4 * s4 + t) * PLD + nt * 8 + g]);
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
          if (nt * 8 < kp)
            *reinterpret_cast<double2*>(sT + (8 * warp + g) * PLD + nt * 8 + 2 * t) = make_double2(acc[nt][0], acc[nt][1]);
      }
      __syncthreads();
      if (active) {                                                 // Y = inv(R) T, packed into the column's slot
        double acc[8][2];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
        for (int s4 = 0; s4 < kp / 4; ++s4) {
          const double a = sR[(4 * s4 + t) * PLD + 8 * warp + g];
#pragma unroll
          for (int nt = 0; nt < 8; ++nt)
            if (nt * 8 < kp) dmma884(acc[nt][0], acc[nt][1], a, sT[(4 * s4 + t) * PLD + nt * 8 + g]);
        }
        const int i = 8 * warp + g;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int cc = nt * 8 + 2 * t + u;
            if (i < k && cc < k && i <= cc) sv[svec_index(i, cc, k)] = (i == cc) ? acc[nt][u] : acc[nt][u] * SQRT2;
          }
        }
      }
      __syncthreads();                                              // sX and sT are rewritten for the next column
    }
    for (int idx = tid; idx < nq * 16; idx += NT) {
      const int quad = q0 + (idx >> 4), c = (idx & 15) >> 2, r = idx & 3;
      const int e = quad * 4 + r - off;
      if (e >= 0 && e < dim && j0 + c < ncols) Atil4[((size_t)quad * ld + j0 + c) * 4 + r] = stage[c * sstr + e];
    }
  }
}

std::atomic<unsigned long long> g_attr[6];
std::atomic<unsigned long long> g_attr_panel{0};
int set_attrs() {
  const void* fn[6] = {(const void*)sdp_apply_kernel, (const void*)sdp_nt_kernel, (const void*)sdp_invert_kernel,
                       (const void*)sdp_prod_div_kernel, (const void*)sdp_maxstep_kernel,
                       (const void*)sdp_scale_panel_kernel};
  for (int i = 0; i < 6; ++i) CIP_TRY(ensure_dyn_smem(fn[i], SDP_SMEM, &g_attr[i]));
  return 0;
}
SDesc sdesc(const ConeDesc& c) { return SDesc{c.slist, c.off, c.sord, c.roff, c.sws, c.sws_stride}; }

}  // namespace

int sdp_max_order() { return KCAP; }
// doubles of global workspace one CTA needs for an S cone of order k (0 up to KMAX: shared memory only)
long long sdp_workspace_doubles(int k) { return k <= KMAX ? 0 : 4LL * k * (k + 1) + k + 8; }
int sdp_panel_chunks(int max_order, int ncols) { return max_order <= KMAX ? (ncols + 15) / 16 : std::min(32, (ncols + 15) / 16); }

int sdp_apply(const ConeDesc& c, const Scaling& F, int use_inv, int transpose, const double* x, double* y,
              cudaStream_t st) {
  if (c.ns == 0) return 0;
  CIP_TRY(set_attrs());
  sdp_apply_kernel<<<c.ns, NT, SDP_SMEM, st>>>(sdesc(c), F.kind, F.R, F.Ri, use_inv, transpose, x, y);
  CIP_CHECK_LAUNCH();
  return 0;
}
int sdp_nt_scaling(const ConeDesc& c, Scaling F, Scaling Fi, const double* v, const double* s, double* lambda,
                   int* info, cudaStream_t st) {
  if (c.ns == 0) return 0;
  CIP_TRY(set_attrs());
  sdp_nt_kernel<<<c.ns, NT, SDP_SMEM, st>>>(sdesc(c), F.kind, Fi.kind, F.R, F.Ri, v, s, lambda, info);
  CIP_CHECK_LAUNCH();
  return 0;
}
int sdp_invert(const ConeDesc& c, Scaling F, cudaStream_t st) {
  if (c.ns == 0) return 0;
  CIP_TRY(set_attrs());
  sdp_invert_kernel<<<c.ns, NT, SDP_SMEM, st>>>(sdesc(c), F.kind, F.R, F.Ri);
  CIP_CHECK_LAUNCH();
  return 0;
}
namespace {
__global__ void sdp_rows_rsub_kernel(SDesc d, const int* __restrict__ kind, const double* __restrict__ minus,
                                     double* __restrict__ y) {
  const int si = blockIdx.x, ci = d.slist[si];
  if (kind[ci] != CIP_BLK_VECCONG) return;      // a Diagonal block on an S cone (F = I at the initial point) was
                                                // handled, subtraction included, by the elementwise kernel
  const int k = d.sord[si], off = d.off[ci], dim = k * (k + 1) / 2;
  for (int e = threadIdx.x; e < dim; e += blockDim.x) y[off + e] = minus[off + e] - y[off + e];
}
}  // namespace
int sdp_rows_rsub(const ConeDesc& c, const Scaling& F, const double* minus, double* y, cudaStream_t st) {
  if (c.ns == 0) return 0;
  sdp_rows_rsub_kernel<<<c.ns, NT, 0, st>>>(sdesc(c), F.kind, minus, y);
  CIP_CHECK_LAUNCH();
  return 0;
}
int sdp_prod_div(const ConeDesc& c, const double* x, const double* y, double* o, int divide, cudaStream_t st) {
  if (c.ns == 0) return 0;
  CIP_TRY(set_attrs());
  sdp_prod_div_kernel<<<c.ns, NT, SDP_SMEM, st>>>(sdesc(c), x, y, o, divide);
  CIP_CHECK_LAUNCH();
  return 0;
}
int sdp_maxstep(const ConeDesc& c, const double* x, const double* d, double d_scale, unsigned long long* key,
                cudaStream_t st) {
  if (c.ns == 0) return 0;
  CIP_TRY(set_attrs());
  sdp_maxstep_kernel<<<c.ns, NT, SDP_SMEM, st>>>(sdesc(c), x, d, d_scale, key);
  CIP_CHECK_LAUNCH();
  return 0;
}
int sdp_scale_panel(const ConeDesc& c, const Scaling& Fi, const double* At4, double* Atil4, int ld, int ncols,
                    cudaStream_t st) {
  if (c.ns == 0) return 0;
  CIP_TRY(set_attrs());
  // orders up to KMAX: tensor-core kernel (one CTA per SM, inv(R) resident, a loop over groups of PCOLS columns);
  // larger orders (and CIP_SDP_PANEL_DMMA=0, the A/B switch): the scalar kernel on shared memory / the workspace
  static const bool use_dmma = [] { const char* e = getenv("CIP_SDP_PANEL_DMMA"); return !e || atoi(e) != 0; }();
  int skip_up_to = 0;
  if (use_dmma) {
    CIP_TRY(ensure_dyn_smem((const void*)sdp_scale_panel_dmma_kernel, PANEL_SMEM, &g_attr_panel));
    const int groups = (ncols + PCOLS - 1) / PCOLS;
    const int per_cone = std::max(1, (sm_count() + c.ns - 1) / c.ns);
    dim3 grid(c.ns, std::min(groups, per_cone));
    sdp_scale_panel_dmma_kernel<<<grid, NT, PANEL_SMEM, st>>>(sdesc(c), Fi.kind, Fi.Ri, At4, Atil4, ld, ncols);
    CIP_CHECK_LAUNCH();
    skip_up_to = KMAX;
    if (c.max_s_ord <= KMAX) return 0;
  }
  const int chunks = sdp_panel_chunks(c.max_s_ord, ncols);
  const int per = (ncols + chunks - 1) / chunks;
  dim3 grid(c.ns, chunks);
  sdp_scale_panel_kernel<<<grid, NT, SDP_SMEM, st>>>(sdesc(c), Fi.kind, Fi.Ri, At4, Atil4, ld, ncols, per, skip_up_to);
  CIP_CHECK_LAUNCH();
  return 0;
}

}  // namespace cip

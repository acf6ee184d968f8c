// Cone kernels K5-K8: Nesterov-Todd scaling, block apply, max step, Jordan product / division,
// and the scaled panel Atil = F^-T A.  R rows are elementwise (vectorised grid-stride); each
// Q cone is handled by one warp (or one 256-thread group for very large cones) with warp
// reductions.  Replaces src/ConicIP.jl:165-194 (nestod_soc), :212-270 (maxstep_rp/soc),
// :305-345 (drp/xrp/dsoc/xsoc), :571-665 (closures) and src/blockmatrices.jl:107-131 (Block*x).
#include <math_constants.h>

#include "kernels.cuh"
#include "../../include/conicip_b200.h"

namespace cip {

namespace {

// ---- group reductions: G == 32 -> one warp per cone; G == 256 -> one CTA per cone
template <int G>
__device__ __forceinline__ double group_sum(double v, double* sm) {
  v = warp_sum(v);
  if (G == 32) return v;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double t = 0;
#pragma unroll
  for (int i = 0; i < G / 32; ++i) t += sm[i];
  return t;
}

template <int G>
__device__ __forceinline__ bool q_cone_of(const ConeDesc& c, int& ci, int& off, int& dim, int& lid) {
  const int per_block = blockDim.x / G;
  const int qi = blockIdx.x * per_block + threadIdx.x / G;
  lid = threadIdx.x % G;
  if (qi >= c.nq) return false;   // G==256: whole block exits together; G==32: whole warp
  ci = c.qlist[qi];
  off = c.off[ci];
  dim = c.off[ci + 1] - off;
  return true;
}

// order-preserving map double -> uint64 for atomicMin
__device__ __forceinline__ unsigned long long dkey(double x) {
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// ================================================================= NT scaling
__global__ void nt_r_kernel(ConeDesc c, const double* __restrict__ v, const double* __restrict__ s, Scaling F,
                            Scaling Fi, double* __restrict__ lambda) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.m; i += gridDim.x * blockDim.x) {
    const int ty = c.type[c.row_cone[i]];
    if (ty == CIP_CONE_S) { F.a[i] = 0.0; F.b[i] = 0.0; Fi.a[i] = 0.0; Fi.b[i] = 0.0; }
    if (ty != CIP_CONE_R) continue;
    const double f = sqrt(s[i] / v[i]);          // Diagonal(sqrt.(yI./xI)), src/ConicIP.jl:598
    F.a[i] = f;
    F.b[i] = 0.0;
    Fi.a[i] = 1.0 / f;
    Fi.b[i] = 0.0;
    lambda[i] = f * v[i];
  }
}

// all rows in R cones (LP / QP): no per-row cone lookup, two rows per iteration with 16-byte accesses
__global__ void nt_r_all_kernel(int m, const double* __restrict__ v, const double* __restrict__ s, Scaling F,
                                Scaling Fi, double* __restrict__ lambda) {
  const int m2 = m >> 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m2; i += gridDim.x * blockDim.x) {
    const double2 vv = reinterpret_cast<const double2*>(v)[i], ss = reinterpret_cast<const double2*>(s)[i];
    const double2 f = make_double2(sqrt(ss.x / vv.x), sqrt(ss.y / vv.y));
    reinterpret_cast<double2*>(F.a)[i] = f;
    reinterpret_cast<double2*>(F.b)[i] = make_double2(0.0, 0.0);
    reinterpret_cast<double2*>(Fi.a)[i] = make_double2(1.0 / f.x, 1.0 / f.y);
    reinterpret_cast<double2*>(Fi.b)[i] = make_double2(0.0, 0.0);
    reinterpret_cast<double2*>(lambda)[i] = make_double2(f.x * vv.x, f.y * vv.y);
  }
  if ((m & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int i = m - 1;
    const double f = sqrt(s[i] / v[i]);
    F.a[i] = f; F.b[i] = 0.0; Fi.a[i] = 1.0 / f; Fi.b[i] = 0.0; lambda[i] = f * v[i];
  }
}

__global__ void nt_kind_kernel(ConeDesc c, Scaling F, Scaling Fi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.ncones) return;
  if (c.type[i] == CIP_CONE_R) {
    F.kind[i] = CIP_BLK_DIAG;  Fi.kind[i] = CIP_BLK_DIAG;  F.D[i] = 0.0;  Fi.D[i] = 0.0;
  } else if (c.type[i] == CIP_CONE_Q) {
    F.kind[i] = CIP_BLK_WOODBURY;  Fi.kind[i] = CIP_BLK_WOODBURY;   // D set by nt_q_kernel
  }
}

template <int G>
__global__ void nt_q_kernel(ConeDesc c, const double* __restrict__ v, const double* __restrict__ s, Scaling F,
                            Scaling Fi, double* __restrict__ lambda) {
  __shared__ double sm[8];
  int ci, off, dim, lid;
  if (!q_cone_of<G>(c, ci, off, dim, lid)) return;
  const double* z = v + off;     // nestod_soc(z = v_I, s = s_I), src/ConicIP.jl:599
  const double* sv = s + off;
  double zz = 0, ss = 0;
  for (int i = lid; i < dim; i += G) { zz += z[i] * z[i]; ss += sv[i] * sv[i]; }
  zz = group_sum<G>(zz, sm);
  ss = group_sum<G>(ss, sm);
  const double qfz = 2 * z[0] * z[0] - zz;       // QF, src/ConicIP.jl:160
  const double qfs = 2 * sv[0] * sv[0] - ss;
  const double beta = sqrt(sqrt(qfs / qfz));     // (QF(s)/QF(z))^(1/4)
  // normalisation by multiplication with the reciprocals (one division per cone instead of four per
  // element; differs from the reference's elementwise z/sqrt(QF(z)) in the last bit only)
  const double irz = 1.0 / sqrt(qfz), irs = 1.0 / sqrt(qfs);
  double zs = 0;
  for (int i = lid; i < dim; i += G) zs += (z[i] * irz) * (sv[i] * irs);
  zs = group_sum<G>(zs, sm);
  const double gamma = sqrt((1 + zs) / 2);
  const double inv2g = 1.0 / (2.0 * gamma);
  // w = (s + Jz)/(2 gamma); w1 += 1; w *= sqrt(2 beta)/sqrt(2 w1)
  const double w1 = inv2g * (sv[0] * irs + z[0] * irz) + 1.0;
  const double scal = sqrt(2 * beta) / sqrt(2 * w1);
  double wv = 0, bib = 0;
  for (int i = lid; i < dim; i += G) {
    const double zi = z[i] * irz, si = sv[i] * irs;
    const double w = (i == 0) ? w1 * scal : (inv2g * (si - zi)) * scal;
    const double a = (i == 0) ? -beta : beta;
    F.a[off + i] = a;
    F.b[off + i] = w;
    const double ia = 1.0 / a;                   // Woodbury inverse: W = inv(A), X = W*B
    Fi.a[off + i] = ia;
    Fi.b[off + i] = ia * w;
    wv += w * z[i];
    bib += w * (ia * w);
  }
  wv = group_sum<G>(wv, sm);
  bib = group_sum<G>(bib, sm);
  for (int i = lid; i < dim; i += G) {                             // lambda = F*v, w recomputed (no re-read)
    const double w = (i == 0) ? w1 * scal : (inv2g * (sv[i] * irs - z[i] * irz)) * scal;
    lambda[off + i] = ((i == 0) ? -beta : beta) * z[i] + w * wv;
  }
  if (lid == 0) {
    F.D[ci] = 1.0;
    Fi.D[ci] = 1.0 / (-1.0 - bib);               // Z = inv(-inv(D) - B'X), D = 1
  }
}

// generic inverse of a flattened scaling (used by cip_factor / cip_set_scaling)
__global__ void inv_diag_kernel(ConeDesc c, Scaling F, Scaling Fi) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.m; i += gridDim.x * blockDim.x) {
    const int kd = F.kind[c.row_cone[i]];
    if (kd == CIP_BLK_VECCONG) {          // handled by the S-cone kernels; keep the flat part inert
      F.a[i] = 0.0; F.b[i] = 0.0; Fi.a[i] = 0.0; Fi.b[i] = 0.0;
      continue;
    }
    const double ia = 1.0 / F.a[i];
    Fi.a[i] = ia;
    Fi.b[i] = (kd == CIP_BLK_WOODBURY) ? ia * F.b[i] : 0.0;
  }
}
__global__ void inv_wood_kernel(ConeDesc c, Scaling F, Scaling Fi) {
  // one warp per cone (any type); only kind-1 blocks do work
  const int ci = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (ci >= c.ncones) return;
  if (lane == 0) Fi.kind[ci] = F.kind[ci];
  if (F.kind[ci] != CIP_BLK_WOODBURY) {
    if (lane == 0) Fi.D[ci] = 0.0;
    return;
  }
  const int off = c.off[ci], dim = c.off[ci + 1] - off;
  double bib = 0;
  for (int i = lane; i < dim; i += 32) bib += F.b[off + i] * (F.b[off + i] / F.a[off + i]);
  bib = warp_sum(bib);
  if (lane == 0) Fi.D[ci] = 1.0 / (-1.0 / F.D[ci] - bib);
}

// ================================================================= block apply
__global__ void apply_diag_kernel(int m, const double* __restrict__ a, const double* __restrict__ x,
                                  double* __restrict__ y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) y[i] = a[i] * x[i];
}
__global__ void apply_wood_kernel(ConeDesc c, Scaling F, const double* __restrict__ x, double* __restrict__ y) {
  const int ci = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (ci >= c.ncones || F.kind[ci] != CIP_BLK_WOODBURY) return;
  const int off = c.off[ci], dim = c.off[ci + 1] - off;
  double bx = 0;
  for (int i = lane; i < dim; i += 32) bx += F.b[off + i] * x[off + i];
  bx = warp_sum(bx) * F.D[ci];
  for (int i = lane; i < dim; i += 32) y[off + i] += F.b[off + i] * bx;   // A*x + B*(D*(B'x))
}

// ================================================================= Jordan product / division
__global__ void prod_r_kernel(ConeDesc c, const double* __restrict__ x, const double* __restrict__ y,
                              double* __restrict__ o, int divide) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.m; i += gridDim.x * blockDim.x) {
    if (c.type[c.row_cone[i]] != CIP_CONE_R) continue;
    o[i] = divide ? x[i] / y[i] : x[i] * y[i];   // drp! / xrp!, src/ConicIP.jl:305-315
  }
}
__global__ void prod_r_all_kernel(int m, const double* __restrict__ x, const double* __restrict__ y,
                                  double* __restrict__ o, int divide) {
  const int m2 = m >> 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m2; i += gridDim.x * blockDim.x) {
    const double2 a = reinterpret_cast<const double2*>(x)[i], b = reinterpret_cast<const double2*>(y)[i];
    reinterpret_cast<double2*>(o)[i] = divide ? make_double2(a.x / b.x, a.y / b.y) : make_double2(a.x * b.x, a.y * b.y);
  }
  if ((m & 1) && blockIdx.x == 0 && threadIdx.x == 0) o[m - 1] = divide ? x[m - 1] / y[m - 1] : x[m - 1] * y[m - 1];
}

template <int G>
__global__ void prod_q_kernel(ConeDesc c, const double* __restrict__ x, const double* __restrict__ y,
                              double* __restrict__ o) {
  __shared__ double sm[8];
  int ci, off, dim, lid;
  if (!q_cone_of<G>(c, ci, off, dim, lid)) return;
  const double* xp = x + off;
  const double* yp = y + off;
  double d = 0;
  for (int i = lid; i < dim; i += G) d += xp[i] * yp[i];
  d = group_sum<G>(d, sm);
  const double x0 = xp[0], y0 = yp[0];
  for (int i = lid; i < dim; i += G) o[off + i] = (i == 0) ? d : x0 * yp[i] + y0 * xp[i];   // xsoc!, :340-345
}
template <int G>
__global__ void div_q_kernel(ConeDesc c, const double* __restrict__ x, const double* __restrict__ y,
                             double* __restrict__ o) {
  // o = arrow(y)^-1 x   (dsoc!(y=x_arg, x=y_arg, o), src/ConicIP.jl:317-338)
  __shared__ double sm[8];
  int ci, off, dim, lid;
  if (!q_cone_of<G>(c, ci, off, dim, lid)) return;
  const double* num = x + off;   // reference's "y" (x1, xb)
  const double* arr = y + off;   // reference's "x" (y1, yb)
  double ybyb = 0, ybxb = 0;
  for (int i = lid; i < dim; i += G) {
    if (i > 0) { ybyb += arr[i] * arr[i]; ybxb += arr[i] * num[i]; }
  }
  ybyb = group_sum<G>(ybyb, sm);
  ybxb = group_sum<G>(ybxb, sm);
  const double y1 = arr[0], x1 = num[0];
  const double alpha = y1 * y1 - ybyb;
  const double b1 = (-x1 / alpha) + ybxb / (y1 * alpha);
  const double b2 = 1.0 / y1;
  for (int i = lid; i < dim; i += G)
    o[off + i] = (i == 0) ? (y1 * x1 - ybxb) / alpha : arr[i] * b1 + num[i] * b2;
}

// ================================================================= max step
__global__ void maxstep_init_kernel(unsigned long long* key) { *key = dkey(CUDART_INF); }
__global__ void maxstep_r_kernel(ConeDesc c, const double* __restrict__ x, const double* __restrict__ d,
                                 double d_scale, unsigned long long* key) {
  double mn = CUDART_INF;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.m; i += gridDim.x * blockDim.x) {
    if (c.type[c.row_cone[i]] != CIP_CONE_R) continue;
    if (d) {
      const double di = d[i] / d_scale;
      if (di > 0) mn = fmin(mn, x[i] / di);                     // maxstep_rp, :212-225
    } else {
      mn = fmin(mn, x[i] > 0 ? 0.0 : -1.0 + x[i]);              // :227-240
    }
  }
  mn = warp_min(mn);
  if ((threadIdx.x & 31) == 0 && mn < CUDART_INF) atomicMin(key, dkey(mn));
}
__device__ __forceinline__ double rp_candidate(double x, const double* d, size_t i, double d_scale) {
  if (d) {
    const double di = d[i] / d_scale;
    return di > 0 ? x / di : CUDART_INF;                           // maxstep_rp, :212-225
  }
  return x > 0 ? 0.0 : -1.0 + x;                                   // :227-240
}
__global__ void maxstep_r_all_kernel(int m, const double* __restrict__ x, const double* __restrict__ d,
                                     double d_scale, unsigned long long* key) {
  double mn = CUDART_INF;
  const int m2 = m >> 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m2; i += gridDim.x * blockDim.x) {
    const double2 xx = reinterpret_cast<const double2*>(x)[i];
    if (d) {
      const double2 dd = reinterpret_cast<const double2*>(d)[i];
      const double d0 = dd.x / d_scale, d1 = dd.y / d_scale;
      if (d0 > 0) mn = fmin(mn, xx.x / d0);
      if (d1 > 0) mn = fmin(mn, xx.y / d1);
    } else {
      mn = fmin(mn, fmin(xx.x > 0 ? 0.0 : -1.0 + xx.x, xx.y > 0 ? 0.0 : -1.0 + xx.y));
    }
  }
  if ((m & 1) && blockIdx.x == 0 && threadIdx.x == 0) mn = fmin(mn, rp_candidate(x[m - 1], d, m - 1, d_scale));
  mn = warp_min(mn);
  if ((threadIdx.x & 31) == 0 && mn < CUDART_INF) atomicMin(key, dkey(mn));
}

// Grid-stride over the cones; every group keeps a running minimum and the CTA issues ONE atomicMin at
// the end (half a million same-address atomics serialise in L2 and used to be the whole kernel time).
template <int G>
__global__ void maxstep_q_kernel(ConeDesc c, const double* __restrict__ x, const double* __restrict__ d,
                                 double d_scale, unsigned long long* key) {
  __shared__ double sm[8];
  __shared__ double smin[8];
  const int per_block = blockDim.x / G, lid = threadIdx.x % G;
  const double ids = d ? -1.0 / d_scale : 0.0;                    // d <- -d / d_scale
  double best = CUDART_INF;
  for (int qi = blockIdx.x * per_block + threadIdx.x / G; qi < c.nq; qi += gridDim.x * per_block) {
    const int ci = c.qlist[qi], off = c.off[ci], dim = c.off[ci + 1] - off;
    const double* xp = x + off;
    double res;
    if (!d) {                                                      // maxstep_soc(x, nothing), :264-270
      double nn = 0;
      for (int i = lid; i < dim; i += G) if (i > 0) nn += xp[i] * xp[i];
      nn = group_sum<G>(nn, sm);
      const double al = sqrt(nn) - xp[0];
      res = al < 0 ? 0.0 : -1.0 - al;
    } else {                                                       // maxstep_soc(x, d), :242-262
      const double* dp = d + off;
      double xx = 0, xdr = 0;
      for (int i = lid; i < dim; i += G) { xx += xp[i] * xp[i]; }
      xx = group_sum<G>(xx, sm);
      const double gam = 2 * xp[0] * xp[0] - xx;
      const double irg = 1.0 / sqrt(gam);
      const double d0 = dp[0] * ids;
      for (int i = lid; i < dim; i += G) xdr += (xp[i] * irg) * (dp[i] * ids);
      const double xd = group_sum<G>(xdr, sm);
      const double xb0 = xp[0] * irg;
      const double beta = 2 * xb0 * d0 - xd;
      const double rho1 = beta * irg;
      const double mu = (beta + d0) / (xb0 + 1);
      double r2 = 0;
      for (int i = lid; i < dim; i += G) {
        if (i > 0) {
          const double r = dp[i] * ids - mu * (xp[i] * irg);
          r2 += r * r;
        }
      }
      r2 = group_sum<G>(r2, sm);
      const double al = sqrt(r2) * irg - rho1;
      res = al < 0 ? CUDART_INF : 1.0 / al;
    }
    best = fmin(best, res);
  }
  // every lane of a group holds the same `best`; CTA minimum through shared memory
  if ((threadIdx.x & 31) == 0) smin[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best = fmin(best, smin[w]);
    if (best < CUDART_INF) atomicMin(key, dkey(best));
  }
}

// ================================================================= scaled panel  Atil = F^-T A
__global__ void __launch_bounds__(256)
scale_panel_diag_kernel(const double* __restrict__ At4, double* __restrict__ Atil4, int ld,
                        const double* __restrict__ ia) {
  const int kq = blockIdx.x;
  const double2 s0 = *reinterpret_cast<const double2*>(ia + 4 * kq);
  const double2 s1 = *reinterpret_cast<const double2*>(ia + 4 * kq + 2);
  const size_t base = (size_t)kq * ld * 4;
  for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < ld; j += gridDim.y * blockDim.x) {
    const double2* in = reinterpret_cast<const double2*>(At4 + base + (size_t)j * 4);
    double2 v0 = __ldg(in), v1 = __ldg(in + 1);
    v0.x *= s0.x; v0.y *= s0.y; v1.x *= s1.x; v1.y *= s1.y;
    double2* out = reinterpret_cast<double2*>(Atil4 + base + (size_t)j * 4);
    out[0] = v0;
    out[1] = v1;
  }
}
__global__ void __launch_bounds__(128)
scale_panel_wood_kernel(ConeDesc c, Scaling Fi, const double* __restrict__ At4, double* __restrict__ Atil4,
                        int ld, int ncols) {
  // block (column chunk, cone): thread per column j of A; adds D * b (b' A[:, j]) over the cone's rows
  const int ci = c.qlist[blockIdx.x];
  if (Fi.kind[ci] != CIP_BLK_WOODBURY) return;
  const int j = blockIdx.y * blockDim.x + threadIdx.x;
  if (j >= ncols) return;
  const int off = c.off[ci], end = c.off[ci + 1];
  double dot = 0;
  for (int k = off; k < end; ++k) dot += Fi.b[k] * At4[q4_index(j, k, ld)];
  dot *= Fi.D[ci];
  for (int k = off; k < end; ++k) Atil4[q4_index(j, k, ld)] += Fi.b[k] * dot;
}

inline int nblocks(int n, int per) {
  int b = (n + per - 1) / per;
  const int cap = sm_count() * 16;
  return b < 1 ? 1 : (b > cap ? cap : b);
}

}  // namespace

#define Q_DISPATCH(kernel, ...)                                                   \
  do {                                                                            \
    if (c.nq > 0) {                                                               \
      if (c.max_q_dim > 1024) {                                                   \
        kernel<256><<<c.nq, 256, 0, st>>>(__VA_ARGS__);                           \
      } else {                                                                    \
        kernel<32><<<(c.nq + 7) / 8, 256, 0, st>>>(__VA_ARGS__);                  \
      }                                                                           \
      CIP_CHECK_LAUNCH();                                                         \
    }                                                                             \
  } while (0)

int cone_nt_scaling(const ConeDesc& c, const double* v, const double* s, Scaling F, Scaling Fi, double* lambda,
                    int* info, cudaStream_t st) {
  if (c.m == 0) return 0;
  nt_kind_kernel<<<(c.ncones + 255) / 256, 256, 0, st>>>(c, F, Fi);
  CIP_CHECK_LAUNCH();
  if (c.nq + c.ns == 0) nt_r_all_kernel<<<nblocks(c.m / 2 + 1, 256), 256, 0, st>>>(c.m, v, s, F, Fi, lambda);
  else if (c.nr_rows + c.ns > 0) nt_r_kernel<<<nblocks(c.m, 256), 256, 0, st>>>(c, v, s, F, Fi, lambda);
  CIP_CHECK_LAUNCH();
  Q_DISPATCH(nt_q_kernel, c, v, s, F, Fi, lambda);
  CIP_TRY(sdp_nt_scaling(c, F, Fi, v, s, lambda, info, st));
  return 0;
}

int cone_invert_scaling(const ConeDesc& c, Scaling F, Scaling Fi, cudaStream_t st) {
  if (c.m == 0) return 0;
  inv_diag_kernel<<<nblocks(c.m, 256), 256, 0, st>>>(c, F, Fi);
  CIP_CHECK_LAUNCH();
  inv_wood_kernel<<<(c.ncones + 7) / 8, 256, 0, st>>>(c, F, Fi);
  CIP_CHECK_LAUNCH();
  CIP_TRY(sdp_invert(c, F, st));
  return 0;
}

int cone_apply(const ConeDesc& c, const Scaling& F, const Scaling& Fi, int op, const double* x, double* y,
               cudaStream_t st) {
  if (c.m == 0) return 0;
  const bool inv = (op == CIP_OP_FINVT || op == CIP_OP_FINV);
  const Scaling& S = inv ? Fi : F;           // R / Q blocks are symmetric: F' = F
  apply_diag_kernel<<<nblocks(c.m, 256), 256, 0, st>>>(c.m, S.a, x, y);
  CIP_CHECK_LAUNCH();
  if (c.nq + c.ns > 0) {   // only non-R cones can carry a Woodbury block
    apply_wood_kernel<<<(c.ncones + 7) / 8, 256, 0, st>>>(c, S, x, y);
    CIP_CHECK_LAUNCH();
  }
  // VecCongurance blocks: F -> R, F' -> R', inv(F) -> inv(R), inv(F)' -> inv(R)'
  CIP_TRY(sdp_apply(c, F, inv ? 1 : 0, (op == CIP_OP_FT || op == CIP_OP_FINVT) ? 1 : 0, x, y, st));
  return 0;
}

int cone_prod(const ConeDesc& c, const double* x, const double* y, double* o, cudaStream_t st) {
  if (c.m == 0) return 0;
  if (c.nq + c.ns == 0) prod_r_all_kernel<<<nblocks(c.m / 2 + 1, 256), 256, 0, st>>>(c.m, x, y, o, 0);
  else if (c.nr_rows > 0) prod_r_kernel<<<nblocks(c.m, 256), 256, 0, st>>>(c, x, y, o, 0);
  CIP_CHECK_LAUNCH();
  Q_DISPATCH(prod_q_kernel, c, x, y, o);
  CIP_TRY(sdp_prod_div(c, x, y, o, 0, st));
  return 0;
}

int cone_div(const ConeDesc& c, const double* x, const double* y, double* o, cudaStream_t st) {
  if (c.m == 0) return 0;
  if (c.nq + c.ns == 0) prod_r_all_kernel<<<nblocks(c.m / 2 + 1, 256), 256, 0, st>>>(c.m, x, y, o, 1);
  else if (c.nr_rows > 0) prod_r_kernel<<<nblocks(c.m, 256), 256, 0, st>>>(c, x, y, o, 1);
  CIP_CHECK_LAUNCH();
  Q_DISPATCH(div_q_kernel, c, x, y, o);
  CIP_TRY(sdp_prod_div(c, x, y, o, 1, st));
  return 0;
}

int cone_maxstep(const ConeDesc& c, const double* x, const double* d, double d_scale, double* partial,
                 int npartial, double* result, cudaStream_t st) {
  (void)partial; (void)npartial;
  unsigned long long* key = reinterpret_cast<unsigned long long*>(result);
  maxstep_init_kernel<<<1, 1, 0, st>>>(key);
  CIP_CHECK_LAUNCH();
  if (c.m == 0) return 0;
  if (c.nq + c.ns == 0) maxstep_r_all_kernel<<<nblocks(c.m / 2 + 1, 256), 256, 0, st>>>(c.m, x, d, d_scale, key);
  else if (c.nr_rows > 0) maxstep_r_kernel<<<nblocks(c.m, 256), 256, 0, st>>>(c, x, d, d_scale, key);
  CIP_CHECK_LAUNCH();
  if (c.nq > 0) {
    if (c.max_q_dim > 1024) maxstep_q_kernel<256><<<std::min(c.nq, sm_count() * 8), 256, 0, st>>>(c, x, d, d_scale, key);
    else maxstep_q_kernel<32><<<std::min((c.nq + 7) / 8, sm_count() * 8), 256, 0, st>>>(c, x, d, d_scale, key);
    CIP_CHECK_LAUNCH();
  }
  CIP_TRY(sdp_maxstep(c, x, d, d_scale, key, st));
  return 0;
}

int cone_scale_panel(const ConeDesc& c, Scaling Fi, const double* At4, double* Atil4, int ld, int m_pad,
                     int ncols, cudaStream_t st) {
  if (m_pad == 0) return 0;
  dim3 grid(m_pad / 4, (ld + 255) / 256 > 8 ? 8 : (ld + 255) / 256);
  scale_panel_diag_kernel<<<grid, 256, 0, st>>>(At4, Atil4, ld, Fi.a);
  CIP_CHECK_LAUNCH();
  if (c.nq > 0) {
    dim3 g2(c.nq, (ncols + 127) / 128);
    scale_panel_wood_kernel<<<g2, 128, 0, st>>>(c, Fi, At4, Atil4, ld, ncols);
    CIP_CHECK_LAUNCH();
  }
  CIP_TRY(sdp_scale_panel(c, Fi, At4, Atil4, ld, ncols, st));
  return 0;
}

}  // namespace cip

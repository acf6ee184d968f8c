// Layout conversion (column-major <-> Q4) and the HBM-bound mat-vec kernels (K4) on Q4 matrices.
// Replaces the host-side `A'*t1` / `A*dy` of src/kktsolvers.jl:326-328 and the residual
// products of src/ConicIP.jl:747-750,:912-914 with kernels on the device-resident layout.
#include "kernels.cuh"

namespace cip {

namespace {

__global__ void pack_rows_kernel(double* __restrict__ dst, int ld, const double* __restrict__ src, int lds, int R,
                                 int K, int Rpad, int Kq) {
  // thread (r, kq): gathers 4 columns of one row (coalesced over r), writes one 32-byte quad
  const int r = blockIdx.y * blockDim.x + threadIdx.x;
  const int kq = blockIdx.x;
  if (r >= Rpad || kq >= Kq) return;
  double v[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int k = kq * 4 + t;
    v[t] = (r < R && k < K) ? src[(size_t)r + (size_t)k * lds] : 0.0;
  }
  double2* o = reinterpret_cast<double2*>(dst + ((size_t)kq * ld + r) * 4);
  o[0] = make_double2(v[0], v[1]);
  o[1] = make_double2(v[2], v[3]);
}

__global__ void pack_trans_kernel(double* __restrict__ dst, int ld, int r0, const double* __restrict__ src,
                                  int lds, int K, int Kq, int nc) {
  // tile: 32 src columns (dst rows) x 32 k-quads; read coalesced along k, write coalesced along r
  __shared__ double tile[32][129];
  const int rbase = blockIdx.y * 32, kqbase = blockIdx.x * 32;
  const int tx = threadIdx.x;  // 0..127
  for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
    const int r = rbase + rr;
    const int k = kqbase * 4 + tx;
    tile[rr][tx] = (r < nc && k < K) ? src[(size_t)k + (size_t)r * lds] : 0.0;
  }
  __syncthreads();
  // 32 rows x 32 quads = 1024 quads; 128*blockDim.y threads
  const int tid = threadIdx.y * 128 + tx;
  const int nthr = 128 * blockDim.y;
  for (int i = tid; i < 1024; i += nthr) {
    const int rr = i & 31, q = i >> 5;
    const int r = rbase + rr, kq = kqbase + q;
    if (r < nc && kq < Kq) {
      double2* o = reinterpret_cast<double2*>(dst + ((size_t)kq * ld + (r0 + r)) * 4);
      o[0] = make_double2(tile[rr][q * 4 + 0], tile[rr][q * 4 + 1]);
      o[1] = make_double2(tile[rr][q * 4 + 2], tile[rr][q * 4 + 3]);
    }
  }
}

__global__ void unpack_rows_kernel(double* __restrict__ dst, int ldd, const double* __restrict__ src, int ld, int R,
                                   int K) {
  const int r = blockIdx.y * blockDim.x + threadIdx.x;
  const int k = blockIdx.x;
  if (r < R && k < K) dst[(size_t)r + (size_t)k * ldd] = src[q4_index(r, k, ld)];
}

__global__ void add_diag_kernel(double* X, int ld, int from, int to, double val, int set) {
  const int i = from + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < to) {
    const size_t idx = q4_index(i, i, ld);
    X[idx] = set ? val : X[idx] + val;
  }
}

__global__ void set_diag_vec_kernel(double* X, int ld, int n, const double* __restrict__ v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) X[q4_index(i, i, ld)] = v[i];
}

// CSC scatter: one CTA per column j; entry (i, j, val) goes to X[r=i,k=j] (transpose=0: Q, G) or
// X[r=j,k=i] (transpose=1: A).  atomicAdd so duplicate entries sum, as in Julia's sparse().
__global__ void scatter_csc_kernel(double* __restrict__ dst, int ld, const long long* __restrict__ colptr,
                                   const long long* __restrict__ rowval, const double* __restrict__ nzval,
                                   int base, int transpose, int nrows, long long nnz, int* __restrict__ bad) {
  const int j = blockIdx.x;
  const long long e0 = colptr[j] - base, e1 = colptr[j + 1] - base;
  if (e0 < 0 || e1 < e0 || e1 > nnz) {
    if (threadIdx.x == 0) atomicCAS(bad, 0, j + 1);
    return;
  }
  for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    const long long il = rowval[e] - base;
    if (il < 0 || il >= nrows) { atomicCAS(bad, 0, j + 1); continue; }
    const int i = (int)il;
    const size_t idx = transpose ? q4_index(j, i, ld) : q4_index(i, j, ld);
    atomicAdd(dst + idx, nzval[e]);
  }
}

__global__ void transpose_scale_kernel(double* __restrict__ dst, int ldd, int k0, const double* __restrict__ src,
                                       int lds, int R, int K, double scale) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;   // column of src (row of dst)
  const int i = blockIdx.y;                              // row of src
  if (j < K && i < R) dst[q4_index(j, k0 + i, ldd)] = scale * src[q4_index(i, j, lds)];
}

// out-partials[split][r] = sum over this split's k-quads of X[r, k] v[k]
__global__ void __launch_bounds__(128)
mv_rows_kernel(double* __restrict__ partial, const double* __restrict__ X, int ld, int R, int Kq,
               const double* __restrict__ v, int kq_per_split, int K) {
  const int r = blockIdx.x * 128 + threadIdx.x;
  const int split = blockIdx.y;
  const int q0 = split * kq_per_split;
  const int q1 = min(Kq, q0 + kq_per_split);
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  if (r < R) {
    const double* xp = X + ((size_t)q0 * ld + r) * 4;
    const size_t step = (size_t)ld * 4;
    int q = q0;
    for (; q + 1 < q1; q += 2) {
      const double2 x0 = __ldg(reinterpret_cast<const double2*>(xp));
      const double2 x1 = __ldg(reinterpret_cast<const double2*>(xp) + 1);
      const double2 y0 = __ldg(reinterpret_cast<const double2*>(xp + step));
      const double2 y1 = __ldg(reinterpret_cast<const double2*>(xp + step) + 1);
      const double2 v0 = *reinterpret_cast<const double2*>(v + 4 * q);
      const double2 v1 = *reinterpret_cast<const double2*>(v + 4 * q + 2);
      const double2 w0 = *reinterpret_cast<const double2*>(v + 4 * q + 4);
      const double2 w1 = *reinterpret_cast<const double2*>(v + 4 * q + 6);
      a0 = fma(x0.x, v0.x, a0); a1 = fma(x0.y, v0.y, a1); a2 = fma(x1.x, v1.x, a2); a3 = fma(x1.y, v1.y, a3);
      a0 = fma(y0.x, w0.x, a0); a1 = fma(y0.y, w0.y, a1); a2 = fma(y1.x, w1.x, a2); a3 = fma(y1.y, w1.y, a3);
      xp += 2 * step;
    }
    for (; q < q1; ++q) {
      const double2 x0 = __ldg(reinterpret_cast<const double2*>(xp));
      const double2 x1 = __ldg(reinterpret_cast<const double2*>(xp) + 1);
      const double2 v0 = *reinterpret_cast<const double2*>(v + 4 * q);
      const double2 v1 = *reinterpret_cast<const double2*>(v + 4 * q + 2);
      a0 = fma(x0.x, v0.x, a0); a1 = fma(x0.y, v0.y, a1); a2 = fma(x1.x, v1.x, a2); a3 = fma(x1.y, v1.y, a3);
      xp += step;
    }
    partial[(size_t)split * R + r] = (a0 + a1) + (a2 + a3);
  }
}

__global__ void mv_rows_reduce_kernel(double* __restrict__ out, const double* __restrict__ partial, int R,
                                      int nsplit, const double* __restrict__ add) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  double s = 0;
  for (int i = 0; i < nsplit; ++i) s += partial[(size_t)i * R + r];
  out[r] = add ? add[r] + s : s;
}

// out[4kq..4kq+3] = sum_r X[r, 4kq..] u[r]; one warp per k-quad
__global__ void __launch_bounds__(256)
mv_k_kernel(double* __restrict__ out, const double* __restrict__ X, int ld, int R, int K,
            const double* __restrict__ u) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kq = blockIdx.x * 8 + warp;
  if (kq * 4 >= K) return;
  const double* xp = X + (size_t)kq * ld * 4;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int r = lane; r < R; r += 32) {
    const double2 x0 = __ldg(reinterpret_cast<const double2*>(xp + (size_t)r * 4));
    const double2 x1 = __ldg(reinterpret_cast<const double2*>(xp + (size_t)r * 4) + 1);
    const double ur = u[r];
    a0 = fma(x0.x, ur, a0); a1 = fma(x0.y, ur, a1); a2 = fma(x1.x, ur, a2); a3 = fma(x1.y, ur, a3);
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
  if (lane == 0) {
    const int k = kq * 4;
    if (k + 0 < K) out[k + 0] = a0;
    if (k + 1 < K) out[k + 1] = a1;
    if (k + 2 < K) out[k + 2] = a2;
    if (k + 3 < K) out[k + 3] = a3;
  }
}

// ---- two right-hand sides per pass over X (cip_solve_multi): the matrix is streamed ONCE, every element feeds two
//      accumulator sets.  Same summation order per vector as the single-vector kernels (bit-identical results).
__global__ void __launch_bounds__(128)
mv_rows2_kernel(double* __restrict__ partial, const double* __restrict__ X, int ld, int R, int Kq,
                const double* __restrict__ va, const double* __restrict__ vb, int kq_per_split, int nsplit) {
  const int r = blockIdx.x * 128 + threadIdx.x;
  const int split = blockIdx.y;
  const int q0 = split * kq_per_split;
  const int q1 = min(Kq, q0 + kq_per_split);
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
  if (r < R) {
    const double* xp = X + ((size_t)q0 * ld + r) * 4;
    const size_t step = (size_t)ld * 4;
    int q = q0;
    for (; q + 1 < q1; q += 2) {
      const double2 x0 = __ldg(reinterpret_cast<const double2*>(xp));
      const double2 x1 = __ldg(reinterpret_cast<const double2*>(xp) + 1);
      const double2 y0 = __ldg(reinterpret_cast<const double2*>(xp + step));
      const double2 y1 = __ldg(reinterpret_cast<const double2*>(xp + step) + 1);
      {
        const double2 v0 = *reinterpret_cast<const double2*>(va + 4 * q), v1 = *reinterpret_cast<const double2*>(va + 4 * q + 2);
        const double2 w0 = *reinterpret_cast<const double2*>(va + 4 * q + 4), w1 = *reinterpret_cast<const double2*>(va + 4 * q + 6);
        a0 = fma(x0.x, v0.x, a0); a1 = fma(x0.y, v0.y, a1); a2 = fma(x1.x, v1.x, a2); a3 = fma(x1.y, v1.y, a3);
        a0 = fma(y0.x, w0.x, a0); a1 = fma(y0.y, w0.y, a1); a2 = fma(y1.x, w1.x, a2); a3 = fma(y1.y, w1.y, a3);
      }
      {
        const double2 v0 = *reinterpret_cast<const double2*>(vb + 4 * q), v1 = *reinterpret_cast<const double2*>(vb + 4 * q + 2);
        const double2 w0 = *reinterpret_cast<const double2*>(vb + 4 * q + 4), w1 = *reinterpret_cast<const double2*>(vb + 4 * q + 6);
        b0 = fma(x0.x, v0.x, b0); b1 = fma(x0.y, v0.y, b1); b2 = fma(x1.x, v1.x, b2); b3 = fma(x1.y, v1.y, b3);
        b0 = fma(y0.x, w0.x, b0); b1 = fma(y0.y, w0.y, b1); b2 = fma(y1.x, w1.x, b2); b3 = fma(y1.y, w1.y, b3);
      }
      xp += 2 * step;
    }
    for (; q < q1; ++q) {
      const double2 x0 = __ldg(reinterpret_cast<const double2*>(xp));
      const double2 x1 = __ldg(reinterpret_cast<const double2*>(xp) + 1);
      const double2 v0 = *reinterpret_cast<const double2*>(va + 4 * q), v1 = *reinterpret_cast<const double2*>(va + 4 * q + 2);
      const double2 w0 = *reinterpret_cast<const double2*>(vb + 4 * q), w1 = *reinterpret_cast<const double2*>(vb + 4 * q + 2);
      a0 = fma(x0.x, v0.x, a0); a1 = fma(x0.y, v0.y, a1); a2 = fma(x1.x, v1.x, a2); a3 = fma(x1.y, v1.y, a3);
      b0 = fma(x0.x, w0.x, b0); b1 = fma(x0.y, w0.y, b1); b2 = fma(x1.x, w1.x, b2); b3 = fma(x1.y, w1.y, b3);
      xp += step;
    }
    partial[(size_t)split * R + r] = (a0 + a1) + (a2 + a3);
    partial[((size_t)nsplit + split) * R + r] = (b0 + b1) + (b2 + b3);
  }
}

__global__ void __launch_bounds__(256)
mv_k2_kernel(double* __restrict__ outa, double* __restrict__ outb, const double* __restrict__ X, int ld, int R, int K,
             const double* __restrict__ ua, const double* __restrict__ ub) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kq = blockIdx.x * 8 + warp;
  if (kq * 4 >= K) return;
  const double* xp = X + (size_t)kq * ld * 4;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
  for (int r = lane; r < R; r += 32) {
    const double2 x0 = __ldg(reinterpret_cast<const double2*>(xp + (size_t)r * 4));
    const double2 x1 = __ldg(reinterpret_cast<const double2*>(xp + (size_t)r * 4) + 1);
    const double ur = ua[r], wr = ub[r];
    a0 = fma(x0.x, ur, a0); a1 = fma(x0.y, ur, a1); a2 = fma(x1.x, ur, a2); a3 = fma(x1.y, ur, a3);
    b0 = fma(x0.x, wr, b0); b1 = fma(x0.y, wr, b1); b2 = fma(x1.x, wr, b2); b3 = fma(x1.y, wr, b3);
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
  b0 = warp_sum(b0); b1 = warp_sum(b1); b2 = warp_sum(b2); b3 = warp_sum(b3);
  if (lane == 0) {
    const int k = kq * 4;
    if (k + 0 < K) { outa[k + 0] = a0; outb[k + 0] = b0; }
    if (k + 1 < K) { outa[k + 1] = a1; outb[k + 1] = b1; }
    if (k + 2 < K) { outa[k + 2] = a2; outb[k + 2] = b2; }
    if (k + 3 < K) { outa[k + 3] = a3; outb[k + 3] = b3; }
  }
}

__global__ void axpby_kernel(double* out, double a, const double* x, double b, const double* y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double r = 0;
    if (x) r = a * x[i];
    if (y) r += b * y[i];
    out[i] = r;
  }
}

}  // namespace

int pack_rows_q4(double* dst, int ld, const double* src, int lds, int R, int K, int Rpad, int Kpad,
                 cudaStream_t s) {
  if (Rpad <= 0 || Kpad <= 0) return 0;
  dim3 grid(Kpad / 4, (Rpad + 127) / 128);
  pack_rows_kernel<<<grid, 128, 0, s>>>(dst, ld, src, lds, R, K, Rpad, Kpad / 4);
  CIP_CHECK_LAUNCH();
  return 0;
}

int pack_trans_q4(double* dst, int ld, int r0, const double* src, int lds, int K, int Kpad, int nc,
                  cudaStream_t s) {
  if (nc <= 0 || Kpad <= 0) return 0;
  const int Kq = Kpad / 4;
  dim3 grid((Kq + 31) / 32, (nc + 31) / 32);
  dim3 block(128, 4);
  pack_trans_kernel<<<grid, block, 0, s>>>(dst, ld, r0, src, lds, K, Kq, nc);
  CIP_CHECK_LAUNCH();
  return 0;
}

int unpack_rows_q4(double* dst, int ldd, const double* src, int ld, int R, int K, cudaStream_t s) {
  if (R <= 0 || K <= 0) return 0;
  dim3 grid(K, (R + 127) / 128);
  unpack_rows_kernel<<<grid, 128, 0, s>>>(dst, ldd, src, ld, R, K);
  CIP_CHECK_LAUNCH();
  return 0;
}

int add_diag_q4(double* X, int ld, int from, int to, double val, int set, cudaStream_t s) {
  if (to <= from) return 0;
  add_diag_kernel<<<(to - from + 127) / 128, 128, 0, s>>>(X, ld, from, to, val, set);
  CIP_CHECK_LAUNCH();
  return 0;
}

int set_diag_vec_q4(double* X, int ld, int n, const double* v, cudaStream_t s) {
  if (n <= 0) return 0;
  set_diag_vec_kernel<<<(n + 127) / 128, 128, 0, s>>>(X, ld, n, v);
  CIP_CHECK_LAUNCH();
  return 0;
}

int scatter_csc_q4(double* dst, int ld, int ncols, const long long* colptr, const long long* rowval,
                   const double* nzval, int base, int transpose, int nrows, long long nnz, int* bad,
                   cudaStream_t s) {
  if (ncols <= 0) return 0;
  scatter_csc_kernel<<<ncols, 128, 0, s>>>(dst, ld, colptr, rowval, nzval, base, transpose, nrows, nnz, bad);
  CIP_CHECK_LAUNCH();
  return 0;
}

int transpose_scale_q4(double* dst, int ldd, int k0, const double* src, int lds, int R, int K, double scale,
                       cudaStream_t s) {
  if (R <= 0 || K <= 0) return 0;
  dim3 grid((K + 127) / 128, R);
  transpose_scale_kernel<<<grid, 128, 0, s>>>(dst, ldd, k0, src, lds, R, K, scale);
  CIP_CHECK_LAUNCH();
  return 0;
}

int fill_zero(double* p, size_t n, cudaStream_t s) {
  if (n == 0) return 0;
  CIP_CUDA(cudaMemsetAsync(p, 0, n * sizeof(double), s));
  return 0;
}

int q4_mv_rows(double* out, const double* X, int ld, int R, int K, const double* v, double* partial,
               int partial_capacity, cudaStream_t s, const double* add) {
  if (R <= 0) return 0;
  const int Kq = (K + 3) / 4;  // v must be readable (zero-padded) up to 4*Kq
  const int rblocks = (R + 127) / 128;
  int nsplit = (sm_count() * 8 + rblocks - 1) / rblocks;
  if (nsplit > Kq) nsplit = Kq > 0 ? Kq : 1;
  if ((long long)nsplit * R > partial_capacity) nsplit = partial_capacity / R;
  if (nsplit < 1) {
    set_error("q4_mv_rows: partial buffer too small");
    return -1;
  }
  const int per = (Kq + nsplit - 1) / nsplit;
  nsplit = per > 0 ? (Kq + per - 1) / per : 1;
  dim3 grid(rblocks, nsplit);
  mv_rows_kernel<<<grid, 128, 0, s>>>(partial, X, ld, R, Kq, v, per, K);
  CIP_CHECK_LAUNCH();
  mv_rows_reduce_kernel<<<(R + 255) / 256, 256, 0, s>>>(out, partial, R, nsplit, add);
  CIP_CHECK_LAUNCH();
  return 0;
}

int q4_mv_k(double* out, const double* X, int ld, int R, int K, const double* u, cudaStream_t s) {
  if (K <= 0) return 0;
  const int Kq = (K + 3) / 4;
  mv_k_kernel<<<(Kq + 7) / 8, 256, 0, s>>>(out, X, ld, R, K, u);
  CIP_CHECK_LAUNCH();
  return 0;
}

int q4_mv_rows2(double* outa, double* outb, const double* X, int ld, int R, int K, const double* va, const double* vb,
                double* partial, int partial_capacity, cudaStream_t s, const double* adda, const double* addb) {
  if (R <= 0) return 0;
  const int Kq = (K + 3) / 4;
  const int rblocks = (R + 127) / 128;
  // the same split of the contraction as q4_mv_rows would choose for one vector (so that every column of a pair
  // gets bit for bit what cip_solve gives it), as long as the partial sums of both vectors fit the buffer
  int nsplit = (sm_count() * 8 + rblocks - 1) / rblocks;
  if (nsplit > Kq) nsplit = Kq > 0 ? Kq : 1;
  if ((long long)nsplit * R > partial_capacity) nsplit = partial_capacity / R;
  if (2LL * nsplit * R > partial_capacity) nsplit = partial_capacity / (2 * R);
  if (nsplit < 1) {
    set_error("q4_mv_rows2: partial buffer too small");
    return -1;
  }
  const int per = (Kq + nsplit - 1) / nsplit;
  nsplit = per > 0 ? (Kq + per - 1) / per : 1;
  dim3 grid(rblocks, nsplit);
  mv_rows2_kernel<<<grid, 128, 0, s>>>(partial, X, ld, R, Kq, va, vb, per, nsplit);
  CIP_CHECK_LAUNCH();
  mv_rows_reduce_kernel<<<(R + 255) / 256, 256, 0, s>>>(outa, partial, R, nsplit, adda);
  CIP_CHECK_LAUNCH();
  mv_rows_reduce_kernel<<<(R + 255) / 256, 256, 0, s>>>(outb, partial + (size_t)nsplit * R, R, nsplit, addb);
  CIP_CHECK_LAUNCH();
  return 0;
}

int q4_mv_k2(double* outa, double* outb, const double* X, int ld, int R, int K, const double* ua, const double* ub,
             cudaStream_t s) {
  if (K <= 0) return 0;
  const int Kq = (K + 3) / 4;
  mv_k2_kernel<<<(Kq + 7) / 8, 256, 0, s>>>(outa, outb, X, ld, R, K, ua, ub);
  CIP_CHECK_LAUNCH();
  return 0;
}

int vec_axpby(double* out, double a, const double* x, double b, const double* y, size_t n, cudaStream_t s) {
  if (n == 0) return 0;
  int blocks = (int)((n + 255) / 256);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  axpby_kernel<<<blocks, 256, 0, s>>>(out, a, x, b, y, n);
  CIP_CHECK_LAUNCH();
  return 0;
}

int vec_copy(double* out, const double* x, size_t n, cudaStream_t s) {
  if (n == 0) return 0;
  CIP_CUDA(cudaMemcpyAsync(out, x, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  return 0;
}

}  // namespace cip

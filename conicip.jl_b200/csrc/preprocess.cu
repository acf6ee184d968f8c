// the preprocessor's rank repair on the device.
//
// cip_imcols replaces `imcols(A, b, eps)` of the reference (src/preprocessor.jl:10-28): find a maximal set
// of linearly independent rows of A (the reference: sparse QR of A', rows with |R_ii| > eps after scaling
// by 1/||A||_F) and check that the equations A x = b are consistent on the dropped rows.
//
// Device algorithm: right-looking modified Gram-Schmidt with greedy (largest residual) row pivoting on a
// row-major working copy -- i.e. QR with column pivoting of A'.  Step k picks the unused row of largest
// residual norm r_kk, stops when r_kk <= eps ||A||_F, normalises it to q_k and removes the q_k component
// from every other unused row; the residual norms are recomputed from the updated rows in the same pass
// (no down-dating).  The whole loop runs without a host synchronisation: the pivot and the stop flag stay
// on the device and later kernels are no-ops once the flag is set.  The O(p^2 n) work streams the working
// copy (HBM-bound: 3 passes per step); what is left for the host are the p x k coefficients
// c_jk = <row_j, q_k>, from which the minimum-norm solution of the kept equations and the residual of
// every equation follow without touching A again:  A_j . x = sum_k c_jk y_k  with  L y = b_kept.
#include <math.h>

#include <algorithm>
#include <vector>

#include "../../include/conicip_b200.h"
#include "common.cuh"

namespace cip {
namespace {

struct McState {     // device-resident loop state
  int nsel;          // rows selected so far
  int done;          // set when the largest residual fell below the threshold
  int piv;           // pivot row of the current step
  double rkk;        // its residual norm
};

// Rm[j][c] = A[j + c * lda]  (column-major in, row-major out), 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) mc_transpose_kernel(const double* __restrict__ A, int lda, int p, int n,
                                                           double* __restrict__ Rm, size_t ldr) {
  __shared__ double t[32][33];
  const int j0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int cc = ty; cc < 32; cc += 8)
    if (j0 + tx < p && c0 + cc < n) t[cc][tx] = A[(size_t)(c0 + cc) * lda + j0 + tx];
  __syncthreads();
  for (int jj = ty; jj < 32; jj += 8)
    if (j0 + jj < p && c0 + tx < n) Rm[(size_t)(j0 + jj) * ldr + c0 + tx] = t[tx][jj];
}

__device__ __forceinline__ double block_sum(double v, double* sm) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sm[i];
  return s;
}

// norms2[j] = ||row j||^2
__global__ void __launch_bounds__(256) mc_rownorm_kernel(const double* __restrict__ Rm, size_t ldr, int n,
                                                         double* __restrict__ norms2) {
  __shared__ double sm[8];
  const double* r = Rm + (size_t)blockIdx.x * ldr;
  double s = 0.0;
  for (int c = threadIdx.x; c < n; c += blockDim.x) s = fma(r[c], r[c], s);
  s = block_sum(s, sm);
  if (threadIdx.x == 0) norms2[blockIdx.x] = s;
}

// one CTA: pivot = unused row of largest residual; stop when it is below the threshold
__global__ void __launch_bounds__(256) mc_pivot_kernel(const double* __restrict__ norms2, int* __restrict__ used, int p,
                                                       double thresh2, McState* st, int* __restrict__ order) {
  __shared__ double bv[256];
  __shared__ int bi[256];
  if (st->done) return;
  double best = -1.0;
  int idx = -1;
  for (int j = threadIdx.x; j < p; j += blockDim.x)
    if (!used[j] && norms2[j] > best) { best = norms2[j]; idx = j; }     // first index wins ties below
  bv[threadIdx.x] = best;
  bi[threadIdx.x] = idx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      const double v = bv[threadIdx.x + o];
      const int i = bi[threadIdx.x + o];
      if (i >= 0 && (v > bv[threadIdx.x] || (v == bv[threadIdx.x] && (bi[threadIdx.x] < 0 || i < bi[threadIdx.x])))) {
        bv[threadIdx.x] = v;
        bi[threadIdx.x] = i;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (bi[0] < 0 || !(bv[0] > thresh2)) {
      st->done = 1;
    } else {
      st->piv = bi[0];
      st->rkk = sqrt(bv[0]);
      used[bi[0]] = 1;
      order[st->nsel] = bi[0];
    }
  }
}

// q = row[piv] / r_kk
__global__ void __launch_bounds__(256) mc_normalize_kernel(const double* __restrict__ Rm, size_t ldr, int n,
                                                           const McState* st, double* __restrict__ q) {
  if (st->done) return;
  const double* r = Rm + (size_t)st->piv * ldr;
  const double inv = 1.0 / st->rkk;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) q[c] = r[c] * inv;
}

// every unused row j: c_jk = <r_j, q>, r_j -= c_jk q, norms2[j] = ||r_j||^2;  the pivot row: c = r_kk
__global__ void __launch_bounds__(256) mc_project_kernel(double* __restrict__ Rm, size_t ldr, int n,
                                                         const double* __restrict__ q, const int* __restrict__ used,
                                                         const McState* st, double* __restrict__ C, int kmax,
                                                         double* __restrict__ norms2) {
  __shared__ double sm[8];
  if (st->done) return;
  const int j = blockIdx.x, k = st->nsel;
  if (j == st->piv) {
    if (threadIdx.x == 0) C[(size_t)j * kmax + k] = st->rkk;
    return;
  }
  if (used[j]) return;
  double* r = Rm + (size_t)j * ldr;
  double d = 0.0;
  for (int c = threadIdx.x; c < n; c += blockDim.x) d = fma(r[c], q[c], d);
  d = block_sum(d, sm);
  double s = 0.0;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    const double v = fma(-d, q[c], r[c]);
    r[c] = v;
    s = fma(v, v, s);
  }
  s = block_sum(s, sm);
  if (threadIdx.x == 0) {
    C[(size_t)j * kmax + k] = d;
    norms2[j] = s;
  }
}

__global__ void mc_advance_kernel(McState* st) {
  if (!st->done) st->nsel += 1;
}

bool device_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

struct Buffers {
  double *A = nullptr, *Rm = nullptr, *q = nullptr, *C = nullptr, *norms2 = nullptr;
  int *used = nullptr, *order = nullptr;
  McState* st = nullptr;
  cudaStream_t s = nullptr;
  ~Buffers() {
    cudaFree(A); cudaFree(Rm); cudaFree(q); cudaFree(C); cudaFree(norms2); cudaFree(used); cudaFree(order); cudaFree(st);
    if (s) cudaStreamDestroy(s);
  }
};

int imcols_impl(int device, const double* A, int lda, int p, int n, const double* b, double eps, int* keep, int* nkeep,
                int* consistent) {
  if (p < 0 || n < 0 || (p > 0 && lda < p) || !keep || !nkeep || !consistent || (p > 0 && n > 0 && (!A || !b))) {
    set_error("cip_imcols: bad arguments (p=%d n=%d lda=%d)", p, n, lda);
    return -2;
  }
  *nkeep = 0;
  *consistent = 1;
  if (p == 0 || n == 0) {
    // the reference returns ([], true) for an empty matrix (src/preprocessor.jl:15)
    return 0;
  }
  if (device >= 0) CIP_CUDA(cudaSetDevice(device));
  {
    int dev = 0;
    cudaDeviceProp prop;
    CIP_CUDA(cudaGetDevice(&dev));
    CIP_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
      set_error("conicip_b200 requires an sm_100a device (found sm_%d%d); there is no fallback path", prop.major,
                prop.minor);
      return -3;
    }
  }
  Buffers B;
  CIP_CUDA(cudaStreamCreateWithFlags(&B.s, cudaStreamNonBlocking));
  const int kmax = std::min(p, n);
  const size_t ldr = ((size_t)n + 3) / 4 * 4;
  const double* dA = A;
  if (!device_ptr(A)) {
    CIP_CUDA(cudaMalloc(&B.A, (size_t)p * n * sizeof(double)));
    CIP_CUDA(cudaMemcpy2DAsync(B.A, (size_t)p * 8, A, (size_t)lda * 8, (size_t)p * 8, n, cudaMemcpyHostToDevice, B.s));
    dA = B.A;
    lda = p;
  }
  CIP_CUDA(cudaMalloc(&B.Rm, (size_t)p * ldr * sizeof(double)));
  CIP_CUDA(cudaMalloc(&B.q, ldr * sizeof(double)));
  CIP_CUDA(cudaMalloc(&B.C, (size_t)p * kmax * sizeof(double)));
  CIP_CUDA(cudaMalloc(&B.norms2, (size_t)p * sizeof(double)));
  CIP_CUDA(cudaMalloc(&B.used, (size_t)p * sizeof(int)));
  CIP_CUDA(cudaMalloc(&B.order, (size_t)p * sizeof(int)));
  CIP_CUDA(cudaMalloc(&B.st, sizeof(McState)));
  CIP_CUDA(cudaMemsetAsync(B.C, 0, (size_t)p * kmax * sizeof(double), B.s));
  CIP_CUDA(cudaMemsetAsync(B.used, 0, (size_t)p * sizeof(int), B.s));
  CIP_CUDA(cudaMemsetAsync(B.st, 0, sizeof(McState), B.s));
  mc_transpose_kernel<<<dim3((p + 31) / 32, (n + 31) / 32), 256, 0, B.s>>>(dA, lda, p, n, B.Rm, ldr);
  CIP_CHECK_LAUNCH();
  mc_rownorm_kernel<<<p, 256, 0, B.s>>>(B.Rm, ldr, n, B.norms2);
  CIP_CHECK_LAUNCH();
  // ||A||_F from the row norms (one small read-back; the threshold of the loop depends on it)
  std::vector<double> hn(p);
  CIP_CUDA(cudaMemcpyAsync(hn.data(), B.norms2, (size_t)p * 8, cudaMemcpyDeviceToHost, B.s));
  CIP_CUDA(cudaStreamSynchronize(B.s));
  double fro2 = 0.0;
  for (double v : hn) fro2 += v;
  const double fro = sqrt(fro2);
  if (!(fro > 0.0)) {
    // the zero matrix: no independent rows; consistent iff b = 0 (the reference divides by norm(A) = 0
    // and returns NaN comparisons, i.e. "inconsistent", unless the matrix is empty)
    std::vector<double> hb(p);
    CIP_CUDA(cudaMemcpy(hb.data(), b, (size_t)p * 8, cudaMemcpyDefault));
    for (double v : hb) if (v != 0.0) *consistent = 0;
    for (int j = 0; j < p; ++j) keep[j] = 0;
    return 0;
  }
  const double thresh = eps * fro;           // |R_kk| / ||A||_F > eps   <=>   r_kk > eps ||A||_F
  const int qblocks = std::max(1, std::min(sm_count() * 4, (n + 255) / 256));
  for (int k = 0; k < kmax; ++k) {
    mc_pivot_kernel<<<1, 256, 0, B.s>>>(B.norms2, B.used, p, thresh * thresh, B.st, B.order);
    CIP_CHECK_LAUNCH();
    mc_normalize_kernel<<<qblocks, 256, 0, B.s>>>(B.Rm, ldr, n, B.st, B.q);
    CIP_CHECK_LAUNCH();
    mc_project_kernel<<<p, 256, 0, B.s>>>(B.Rm, ldr, n, B.q, B.used, B.st, B.C, kmax, B.norms2);
    CIP_CHECK_LAUNCH();
    mc_advance_kernel<<<1, 1, 0, B.s>>>(B.st);
    CIP_CHECK_LAUNCH();
  }
  McState hst;
  CIP_CUDA(cudaMemcpyAsync(&hst, B.st, sizeof(McState), cudaMemcpyDeviceToHost, B.s));
  CIP_CUDA(cudaStreamSynchronize(B.s));
  const int r = hst.nsel;
  std::vector<int> order(r);
  std::vector<double> C((size_t)p * kmax), hb(p);
  if (r > 0) CIP_CUDA(cudaMemcpy(order.data(), B.order, (size_t)r * sizeof(int), cudaMemcpyDeviceToHost));
  CIP_CUDA(cudaMemcpy(C.data(), B.C, (size_t)p * kmax * 8, cudaMemcpyDeviceToHost));
  CIP_CUDA(cudaMemcpy(hb.data(), b, (size_t)p * 8, cudaMemcpyDefault));
  for (int j = 0; j < p; ++j) keep[j] = 0;
  for (int i = 0; i < r; ++i) keep[order[i]] = 1;
  *nkeep = r;
  if (r == 0) return 0;                      // the reference: isempty(R) -> ([], true)   (:24)
  // kept row order[i] = sum_{k <= i} C[order[i]][k] q_k: forward substitution for the minimum-norm solution
  // x = sum_k y_k q_k of the kept equations (the reference's A[R,:] \ b[R], :26), all in units of ||A||_F
  std::vector<double> y(r);
  for (int i = 0; i < r; ++i) {
    const double* ci = &C[(size_t)order[i] * kmax];
    double s = hb[order[i]];
    for (int k = 0; k < i; ++k) s -= ci[k] * y[k];
    y[i] = s / ci[i];
  }
  double worst = 0.0;
  for (int j = 0; j < p; ++j) {
    const double* cj = &C[(size_t)j * kmax];
    double s = -hb[j];
    for (int k = 0; k < r; ++k) s += cj[k] * y[k];
    worst = std::max(worst, fabs(s));
  }
  *consistent = (worst / fro < eps) ? 1 : 0;   // norm(A x - b, Inf) < eps after the 1/||A|| scaling (:12, :26)
  return 0;
}

}  // namespace
}  // namespace cip

extern "C" int cip_imcols(int device, const double* A, int lda, int p, int n, const double* b, double eps, int* keep,
                          int* nkeep, int* consistent) {
  return cip::imcols_impl(device, A, lda, p, n, b, eps, keep, nkeep, consistent);
}

// K2/K3: blocked right-looking Cholesky of a Q4-layout SPD matrix and the triangular sweeps.
//   per 128-wide panel:  potrf_diag (one CTA: L11 and inv(L11) in shared memory)
//                        L21  = A21 * inv(L11)'      -> gemm_nt (DMMA tiles)
//                        A22 -= L21 * L21'           -> gemm_nt (lower, same tiles as the SYRK)
// Replaces LAPACK qr/lu at src/kktsolvers.jl:35,:295 and the solves at :39-48,:299.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "kernels.cuh"
#include "nccl_dl.h"

namespace cip {

namespace {
constexpr int NB = 128;

// 1 / sqrt(d), branch-free (the IEEE sqrt and division sequences carry slow-path branches that cut the
// instruction stream of the sweep into pieces): 20-bit seed, one third-order step, |e|^3 < 2^-60.
__device__ __forceinline__ double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d * y, y, 1.0);
  const double p = fma(0.375, e, 0.5) * e;
  return fma(y, p, y);
}

// Cholesky factor L and inverse X = inv(L) of one 128x128 diagonal block, one CTA of 256 threads.
// The block lives in shared memory column-major, Lc[c * LD + r] = L(r, c) for r >= c; the free triangle
// (r < c) receives the strict lower triangle of X transposed, X(r, c) -> Lc[r * LD + c], and
// X(r, r) = dinv[r].  LD = 132 makes every access pattern below conflict-free: lanes along r are
// contiguous, and DMMA fragments (4 columns x 8 rows, or 8 columns x 4 rows) hit 16 distinct bank pairs
// per half-warp.
//   * four 32-column sub-panels.  Sweep: one thread per matrix row keeps its 32 sub-panel entries in
//     registers.  The warp that owns the 32x32 diagonal block eliminates it alone (pivot by shuffle,
//     column broadcast through a 32x32 scratch, __syncwarp only; scaling deferred so the chain per column
//     is SHFL -> rcp -> FMA); one named barrier later the warps below apply the 32 published columns
//     to their rows without further synchronisation (the in-block TRSM).
//   * rank-32 update of the trailing columns on the tensor pipe: 16x16 blocks per warp, DMMA.8x8x4
//     fragments read straight from the column-major block.
//   * inv(L): the four diagonal 32x32 blocks by one warp each (thread = column, forward substitution
//     in registers against broadcast reads of L), then the off-diagonal blocks
//     X_ij = -X_ii * sum_k L_ik X_kj by block distance as 8x8 DMMA tiles.
constexpr int SB = 32;           // sub-panel width
constexpr int LD = 132;          // column stride of the shared block
constexpr int TLD = 36;          // row stride of the 32x32 T blocks of the inverse
constexpr int PT = 256;          // threads of potrf_diag_kernel: 8 warps, so that a thread may hold ~250 registers (row + load batch)
constexpr int PW = PT / 32;
static_assert(PW == 8 && NB / SB == 4, "the named-barrier thread counts (192, 224) and the warp roles of potrf_diag_kernel assume 8 warps and 4 sub-panels");
constexpr int POTRF_SMEM = (NB * LD + NB + 2 * (SB * SB + 2 * SB) + 3 * SB * TLD) * (int)sizeof(double);

// X(rr, cc) of the (partially built) inverse: strict lower part from the free triangle, diagonal from dinv
__device__ __forceinline__ double xinv_at(const double* Lc, const double* dinv, int rr, int cc) {
  // both loads unconditional (any (rr, cc) is inside the block), the choice is a select: lanes of a DMMA
  // fragment straddle the diagonal and must not diverge
  const double v = Lc[rr * LD + cc], dv = dinv[rr];
  return rr > cc ? v : (rr == cc ? dv : 0.0);
}

// Look-ahead part of the rank-32 update: the 32x32 diagonal block (s, s) right of the sub-panel at c0, as ten
// lower 8x8 tiles t = idx, idx + nw (< 10), both advanced together.
__device__ __forceinline__ void own_block_update(double* Lc, int c0, int idx, int nw, int g, int t4) {
  const int base = c0 + SB;
  int tr[2], tc[2];
  bool on[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int t = idx + u * nw;
    on[u] = t < 10;
    int r = 0, c = on[u] ? t : 0;
    while (c > r) { c -= r + 1; ++r; }
    tr[u] = r; tc[u] = c;
  }
  double av[2][SB / 4], bv[2][SB / 4], c0v[2] = {0.0, 0.0}, c1v[2] = {0.0, 0.0};
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int kq = 0; kq < SB / 4; ++kq) {
      const double* colp = Lc + (c0 + 4 * kq + t4) * LD + base + g;
      av[u][kq] = colp[8 * tr[u]];
      bv[u][kq] = colp[8 * tc[u]];
    }
#pragma unroll
  for (int kq = 0; kq < SB / 4; ++kq)
#pragma unroll
    for (int u = 0; u < 2; ++u) dmma884(c0v[u], c1v[u], av[u][kq], bv[u][kq]);
  double cv[2][2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int row = base + 8 * tr[u] + g, col = base + 8 * tc[u] + 2 * t4;
    cv[u][0] = Lc[col * LD + row];
    cv[u][1] = Lc[(col + 1) * LD + row];
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int row = base + 8 * tr[u] + g, col = base + 8 * tc[u] + 2 * t4;
    if (on[u] && row >= col) Lc[col * LD + row] = cv[u][0] - c0v[u];
    if (on[u] && row >= col + 1) Lc[(col + 1) * LD + row] = cv[u][1] - c1v[u];
  }
}

// Off-diagonal blocks of the inverse at block distance DLT: X(b+DLT, b) = -X(b+DLT, b+DLT) T_b with
// T_b = sum_{kb = b .. b+DLT-1} L(b+DLT, kb) X(kb, b).  A 32x32 block is 16 tiles of 8x8; warp w owns the
// tiles (tr, tc) = (w >> 2, w & 3) and (tr + 2, tc) of every block and advances all of them together, one
// DMMA each per k step, so that the dependent-DMMA latency of one tile is hidden by the others.
template <int DLT>
__device__ __forceinline__ void inv_offdiag(double* Lc, const double* dinv, double* T, int w, int g, int t4) {
  constexpr int NT = 2 * (NB / SB - DLT);        // tiles of this warp
  const int tc = w & 3, tr0 = w >> 2;
  double c0v[NT], c1v[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) c0v[i] = c1v[i] = 0.0;
  // rows of X(., cj..cj+7) above cj are zero: the first k block starts at step 2 tc
#pragma unroll 2
  for (int st = 2 * tc; st < DLT * 8; ++st) {
    double av[NT], bv[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int b = i >> 1, tr = tr0 + 2 * (i & 1);
      const int kk = b * SB + 4 * st + t4;
      av[i] = Lc[kk * LD + (b + DLT) * SB + 8 * tr + g];
      bv[i] = (st < 8) ? xinv_at(Lc, dinv, kk, b * SB + 8 * tc + g) : Lc[kk * LD + b * SB + 8 * tc + g];
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) dmma884(c0v[i], c1v[i], av[i], bv[i]);
  }
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int b = i >> 1, tr = tr0 + 2 * (i & 1);
    *reinterpret_cast<double2*>(T + (b * SB + 8 * tr + g) * TLD + 8 * tc + 2 * t4) = make_double2(c0v[i], c1v[i]);
    c0v[i] = c1v[i] = 0.0;
  }
  __syncthreads();
  // X_ii is lower triangular: tile row tr needs k < 8 tr + 8
#pragma unroll
  for (int st = 0; st < 8; ++st) {
    double av[NT], bv[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int b = i >> 1, tr = tr0 + 2 * (i & 1), i0 = (b + DLT) * SB;
      av[i] = xinv_at(Lc, dinv, i0 + 8 * tr + g, i0 + 4 * st + t4);
      bv[i] = T[(b * SB + 4 * st + t4) * TLD + 8 * tc + g];
    }
#pragma unroll
    for (int i = 0; i < NT; ++i)
      if (st < 2 * (tr0 + 2 * (i & 1)) + 2) dmma884(c0v[i], c1v[i], av[i], bv[i]);
  }
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int b = i >> 1, tr = tr0 + 2 * (i & 1), i0 = (b + DLT) * SB;
    *reinterpret_cast<double2*>(Lc + (i0 + 8 * tr + g) * LD + b * SB + 8 * tc + 2 * t4) = make_double2(-c0v[i], -c1v[i]);
  }
  __syncthreads();
}

#ifdef CIP_POTRF_PROF
#define PROF_MARK(i) do { prof_t[i] = clock64(); } while (0)
#define DBG_MARK(s, k) do { if (lane == 0) dbg[s][k] = clock64(); } while (0)
#else
#define PROF_MARK(i) do { } while (0)
#define DBG_MARK(s, k) do { } while (0)
#endif

// Row application a[j] -= tt * row[j] for j in (K, 32), `row` being a broadcast row of 32 doubles in
// shared memory.  The loads are issued as one batch of 16-byte loads ahead of the arithmetic (with one
// warp per scheduler nothing else hides the LDS latency), split from the FMAs so that the caller can
// put the pivot reciprocal chain between the two.
template <int K>
struct RowBatch {
  static constexpr int P0 = (K + 1) >> 1;          // first pair that holds an index > K
  double2 u[16 - P0 > 0 ? 16 - P0 : 1];
  __device__ __forceinline__ void load(const double* row) {
    const double2* r2 = reinterpret_cast<const double2*>(row);
#pragma unroll
    for (int p = P0; p < 16; ++p) u[p - P0] = r2[p];
  }
  __device__ __forceinline__ double first() const { return ((K + 1) & 1) ? u[0].y : u[0].x; }   // row[K + 1]
  __device__ __forceinline__ void apply_rest(double (&a)[SB], double tt) const {                 // j >= K + 2
#pragma unroll
    for (int p = P0; p < 16; ++p) {
      if (2 * p > K + 1) a[2 * p] = fma(-tt, u[p - P0].x, a[2 * p]);
      if (2 * p + 1 > K + 1) a[2 * p + 1] = fma(-tt, u[p - P0].y, a[2 * p + 1]);
    }
  }
  __device__ __forceinline__ void apply(double (&a)[SB], double tt) const {
#pragma unroll
    for (int p = P0; p < 16; ++p) {
      if (2 * p > K) a[2 * p] = fma(-tt, u[p - P0].x, a[2 * p]);
      a[2 * p + 1] = fma(-tt, u[p - P0].y, a[2 * p + 1]);
    }
  }
};
// One column of the diagonal 32x32 block (warp-synchronous).  On entry column K is already published
// (Bs[K][lane] = a[K]) and `d` is its pivot; the stage finishes the next column first and publishes it
// before touching the other 30 entries, so the dependent chain per column is
// SHFL -> MUFU.RCP64H -> 3 DFMA -> DFMA.  The reciprocal is never formed on the chain:
// a[K] / d = t0 (1 + e + e^2) with r0 ~ 1/d to 20 bits, t0 = a[K] r0, e = 1 - d r0  (|e|^3 < 2^-60).
template <int K>
struct DiagSweep {
  static __device__ __forceinline__ void run(double (&a)[SB], double* Bs, double* rd, double& dj, int& badcol, int lane,
                                             double d, int nconsumers) {
    __syncwarp();
    RowBatch<K> rb;
    rb.load(Bs + K * SB);
    const bool bad = !(d > 0.0);
    badcol = (bad && badcol == 0) ? K + 1 : badcol;
    d = bad ? 1.0 : d;
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
    const double e = fma(-d, r0, 1.0), t0 = a[K] * r0;
    const double pp = fma(e, e, e);
    const double tt = fma(t0, pp, t0);
    double dn = 0.0;
    if (K + 1 < SB) {
      a[(K + 1) % SB] = fma(-tt, rb.first(), a[(K + 1) % SB]);
      Bs[((K + 1) % SB) * SB + lane] = a[(K + 1) % SB];
      dn = __shfl_sync(0xffffffffu, a[(K + 1) % SB], (K + 1) % SB);
    }
    if (lane == K) { dj = d; rd[K] = fma(r0, pp, r0); }
    rb.apply_rest(a, tt);
    if ((K & 7) == 7 && nconsumers > 0)           // columns K-7 .. K are published: release the warps below
      asm volatile("bar.arrive %0, %1;" ::"r"(2 + (K >> 3)), "r"((nconsumers + 1) * 32) : "memory");
    DiagSweep<K + 1>::run(a, Bs, rd, dj, badcol, lane, dn, nconsumers);
  }
};
template <>
struct DiagSweep<SB> {
  static __device__ __forceinline__ void run(double (&)[SB], double*, double*, double&, int&, int, double, int) {}
};
// Rows below the diagonal block: apply the published columns, eight at a time as the diagonal warp
// releases them (named barriers 2..5).
template <int K>
struct BelowSweep {
  static __device__ __forceinline__ void run(double (&a)[SB], const double* Bs, const double* rd, int nthreads) {
    if ((K & 7) == 0) asm volatile("bar.sync %0, %1;" ::"r"(2 + (K >> 3)), "r"(nthreads) : "memory");
    RowBatch<K> rb;
    rb.load(Bs + K * SB);
    rb.apply(a, a[K] * rd[K]);
    BelowSweep<K + 1>::run(a, Bs, rd, nthreads);
  }
};
template <>
struct BelowSweep<SB> {
  static __device__ __forceinline__ void run(double (&)[SB], const double*, const double*, int) {}
};

// Rank-32 update of the trailing matrix by the sub-panel at columns c0 .. c0+31, lower 16x16 blocks
// blk = first, first + stride, ..., on DMMA tiles.
// PANEL = true:  the blocks (bi >= 2, bj < 2), i.e. the columns of the next sub-panel below its diagonal block;
// PANEL = false: the blocks (bi >= bj >= 2), the trailing matrix behind the next sub-panel.
template <bool PANEL>
__device__ __forceinline__ void rank32_update(double* Lc, int c0, int first, int stride, int g, int t4) {
  const int base = c0 + SB;
  const int nb16 = (NB - base) / 16 - 2;         // block rows / columns behind the 32x32 block (s, s)
  const int nblk = PANEL ? 2 * nb16 : nb16 * (nb16 + 1) / 2;
  for (int blk = first; blk < nblk; blk += stride) {
    int bi, bj;
    if (PANEL) {
      bi = 2 + (blk >> 1); bj = blk & 1;
    } else {
      bi = 0; bj = blk;
      while (bj > bi) { bj -= bi + 1; ++bi; }
      bi += 2; bj += 2;
    }
    const int i0 = base + 16 * bi, jj0 = base + 16 * bj;
    double acc[2][2][2];
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int y = 0; y < 2; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
#pragma unroll
    for (int kq = 0; kq < SB / 4; ++kq) {
      const double* colp = Lc + (c0 + 4 * kq + t4) * LD;
      const double a0 = colp[i0 + g], a1 = colp[i0 + 8 + g];
      const double b0 = colp[jj0 + g], b1 = colp[jj0 + 8 + g];
      dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
      dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
      dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
      dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
    }
    // read-modify-write of the C block: all loads first (the compiler cannot reorder them past the stores itself)
    double cv[2][2][2];
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int y = 0; y < 2; ++y) {
        const int row = i0 + 8 * x + g, col = jj0 + 8 * y + 2 * t4;
        cv[x][y][0] = Lc[col * LD + row];
        cv[x][y][1] = Lc[(col + 1) * LD + row];
      }
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int y = 0; y < 2; ++y) {
        const int row = i0 + 8 * x + g, col = jj0 + 8 * y + 2 * t4;
        if (row >= col) Lc[col * LD + row] = cv[x][y][0] - acc[x][y][0];
        if (row >= col + 1) Lc[(col + 1) * LD + row] = cv[x][y][1] - acc[x][y][1];
      }
  }
}

__global__ void __launch_bounds__(PT, 1)
potrf_diag_kernel(double* __restrict__ H, int ld, int j0, double* __restrict__ W, int* info) {
  extern __shared__ __align__(16) double Lc[];   // Lc[c * LD + r]
  double* dinv = Lc + NB * LD;                   // [NB]      1 / L_kk
  double* Bs0 = dinv + NB;                       // 2 x { [SB][SB] unscaled diagonal block of the sub-panel, Bs[k][j] = S(c0+j, c0+k);
                                                 //       [SB] 1 / d_k (pivot reciprocals); [SB] 1 / sqrt(d_k) }, by parity of the sub-panel
  double* T = Bs0 + 2 * (SB * SB + 2 * SB);      // [3][SB][TLD] block temporaries of the inverse
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;        // DMMA fragment coordinates
#ifdef CIP_POTRF_PROF
  long long prof_t[16];
  __shared__ long long dbg[4][8];
#endif
  PROF_MARK(0);
  {
    // all 32-byte loads of a thread in flight before the first store (sixteen quads, two batches of eight)
    const int r = tid & 127, qh = tid >> 7;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      double2 v0[8], v1[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int q = qh + 2 * (8 * half + u);
        if (4 * q <= r) {                        // strictly-upper quads are never read
          const double2* p = reinterpret_cast<const double2*>(H + ((size_t)(j0 / 4 + q) * ld + j0 + r) * 4);
          v0[u] = p[0];
          v1[u] = p[1];
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int q = qh + 2 * (8 * half + u);
        if (4 * q <= r) {
          Lc[(4 * q + 0) * LD + r] = v0[u].x;
          Lc[(4 * q + 1) * LD + r] = v0[u].y;
          Lc[(4 * q + 2) * LD + r] = v1[u].x;
          Lc[(4 * q + 3) * LD + r] = v1[u].y;
        }
      }
    }
  }
  __syncthreads();
  PROF_MARK(1);

  // Sub-panel s = columns c0 .. c0+31.  Iteration s first applies the rank-32 update of sub-panel s-1 with
  // look-ahead, then sweeps:
  //   warp s           waits until block (s, s) has been updated (barrier 7), eliminates it, publishing its columns
  //                    eight at a time (bar.arrive 2..5) and the scaling (6); stores its rows after the iteration's
  //                    closing barrier, nobody needs them before the inverse
  //   warp s+4         idle: it shares the scheduler (and the FP64 pipe) with warp s
  //   the other six    update block (s, s) first (bar.arrive 7; not warp s-1, which is still storing), then the
  //                    columns of this sub-panel below it (barrier 1); then warps s+1 .. 3 apply the published
  //                    columns to their rows while the others update the trailing matrix behind the sub-panel
#pragma unroll 1
  for (int s = 0; s < NB / SB; ++s) {
    const int c0 = s * SB;
    const int nbelow = NB / SB - 1 - s;
    double* Bs = Bs0 + (s & 1) * (SB * SB + 2 * SB);
    double* rd = Bs + SB * SB;
    double* rsq = rd + SB;
    if (w == s) {
      DBG_MARK(s, 0);
      asm volatile("bar.sync 7, 192;" ::: "memory");   // block (s, s) is up to date
      const int i = tid;
      double a[SB];
#pragma unroll
      for (int j = 0; j < SB; ++j) {
        const double v = Lc[(c0 + j) * LD + i];        // unconditional load + select: no divergence
        a[j] = (j <= lane) ? v : 0.0;
      }
      double dj = 1.0;
      int badcol = 0;
      DBG_MARK(s, 1);
      Bs[lane] = a[0];
      DiagSweep<0>::run(a, Bs, rd, dj, badcol, lane, __shfl_sync(0xffffffffu, a[0], 0), nbelow);
      if (badcol && lane == 0) atomicCAS(info, 0, j0 + c0 + badcol);
      const double rsj = fast_rsqrt(dj);
      rsq[lane] = rsj;
      dinv[c0 + lane] = rsj;
      if (nbelow > 0) asm volatile("bar.arrive 6, %0;" ::"r"((nbelow + 1) * 32) : "memory");
      DBG_MARK(s, 2);
      __syncthreads();                                 // closes iteration s
      // L(i, c0 + j) = a[j] / sqrt(d_j), diagonal d_j / sqrt(d_j).  Entries right of the diagonal land in the
      // free triangle (garbage, overwritten by the inverse later): no predicate, no divergence
      double rq[SB];
#pragma unroll
      for (int j = 0; j < SB; ++j) rq[j] = rsq[j];
#pragma unroll
      for (int j = 0; j < SB; ++j) Lc[(c0 + j) * LD + i] = ((j == lane) ? dj : a[j]) * rq[j];
      DBG_MARK(s, 3);
    } else {
      // barrier 7 is unconditional (five arrivals + the diagonal warp, also when there is nothing to update at
      // s = 0): a conditional one makes the compiler peel the first iteration and duplicate the unrolled sweeps
      if (w != s + 4 && w != (s + NB / SB - 1) % (NB / SB)) {   // not warp s-1: it is still storing its rows of sub-panel s-1
        const int idx5 = w - (w > s - 1) - (w > s) - (w > s + 4);     // 0 .. 4
        if (s > 0) own_block_update(Lc, c0 - SB, idx5, PW - 3, g, t4);
        asm volatile("bar.arrive 7, 192;" ::: "memory");
      }
      if (s > 0 && w != s + 4) {
        const int idx6 = w - (w > s) - (w > s + 4);    // 0 .. 5
        rank32_update<true>(Lc, c0 - SB, idx6, PW - 2, g, t4);      // the columns of this sub-panel
      }
      if (w == 7) DBG_MARK(s, 4);
      asm volatile("bar.sync 1, %0;" ::"r"((PW - 1) * 32) : "memory");
      if (w == 3) DBG_MARK(s, 5);
      if (s > 0 && w != s + 4 && !(w > s && w < NB / SB)) {
        // the trailing matrix behind this sub-panel, by the warps that have no rows to sweep
        const int nfree = PW - 2 - nbelow;
        const int idxf = (w < s) ? w : w - (NB / SB) - (w > s + 4) + s;   // warps 0 .. s-1, then 4 .. 7 without s + 4
        rank32_update<false>(Lc, c0 - SB, idxf, nfree, g, t4);
      }
      if (w > s && w < NB / SB) {
        const int i = tid;
        double a[SB];
#pragma unroll
        for (int j = 0; j < SB; ++j) a[j] = Lc[(c0 + j) * LD + i];
        BelowSweep<0>::run(a, Bs, rd, (nbelow + 1) * 32);
        asm volatile("bar.sync 6, %0;" ::"r"((nbelow + 1) * 32) : "memory");
        double rq[SB];
#pragma unroll
        for (int j = 0; j < SB; ++j) rq[j] = rsq[j];
#pragma unroll
        for (int j = 0; j < SB; ++j) Lc[(c0 + j) * LD + i] = a[j] * rq[j];
        if (w == 3) DBG_MARK(s, 6);
      }
      __syncthreads();                                 // closes iteration s
    }
    if (w == 0) DBG_MARK(s, 7);
    PROF_MARK(2 + 2 * s);
    PROF_MARK(3 + 2 * s);
  }
  __syncthreads();                                     // the last diagonal warp has stored its rows

  // ---- write L back: the quads up to the diagonal (zeros right of it inside the last one; the quads
  //      beyond are never read by anyone).  Only reads the lower triangle, so it needs no barrier
  //      against the inverse below, which writes the free triangle
  {
    const int r = tid & 127;
    for (int q = tid >> 7; q < 32; q += PT / 128) {
      if (4 * q > r) continue;
      double o[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) o[t] = (4 * q + t > r) ? 0.0 : Lc[(4 * q + t) * LD + r];
      double2* p = reinterpret_cast<double2*>(H + ((size_t)(j0 / 4 + q) * ld + j0 + r) * 4);
      p[0] = make_double2(o[0], o[1]);
      p[1] = make_double2(o[2], o[3]);
    }
  }
  PROF_MARK(10);

  // ---- inverse, diagonal 32x32 blocks: warp b, lane = column.  x[i] first accumulates
  //      sum_k L(i, k) X(k, j), then becomes X(i, j) = -x[i] / L(i, i).
  if (w < NB / SB) {
    const int b0 = w * SB;
    double x[SB];
#pragma unroll
    for (int i = 0; i < SB; ++i) x[i] = 0.0;
#pragma unroll
    for (int k = 0; k < SB; ++k) {
      const double dk = dinv[b0 + k];
      x[k] = (k < lane) ? 0.0 : ((k == lane) ? dk : -dk * x[k]);
      const double* col = Lc + (b0 + k) * LD + b0;               // L(b0 + i, b0 + k), broadcast reads
#pragma unroll
      for (int i = k + 1; i < SB; ++i) x[i] = fma(col[i], x[k], x[i]);
    }
#pragma unroll
    for (int i = 1; i < SB; ++i)
      if (i > lane) Lc[(b0 + i) * LD + b0 + lane] = x[i];
  }
  __syncthreads();
  PROF_MARK(11);
  // ---- inverse, off-diagonal blocks by block distance (inv_offdiag below)
  inv_offdiag<1>(Lc, dinv, T, w, g, t4);
  PROF_MARK(12);
  inv_offdiag<2>(Lc, dinv, T, w, g, t4);
  PROF_MARK(13);
  inv_offdiag<3>(Lc, dinv, T, w, g, t4);
  PROF_MARK(14);
  {
    // W[r][c] = X(r, c); quads right of the diagonal stay zero (the buffer is zero-initialised and only
    // this kernel, or a broadcast of its output, ever writes it)
    const int r = tid & 127;
    for (int q = tid >> 7; q < 32; q += PT / 128) {
      if (4 * q > r) continue;
      const double2 x0 = *reinterpret_cast<const double2*>(Lc + r * LD + 4 * q);
      const double2 x1 = *reinterpret_cast<const double2*>(Lc + r * LD + 4 * q + 2);
      double o[4] = {x0.x, x0.y, x1.x, x1.y};
      if (4 * q + 3 >= r) {                       // the quad that holds the diagonal
        const double dr = dinv[r];
#pragma unroll
        for (int t = 0; t < 4; ++t) o[t] = (4 * q + t < r) ? o[t] : ((4 * q + t == r) ? dr : 0.0);
      }
      double2* p = reinterpret_cast<double2*>(W + ((size_t)q * NB + r) * 4);
      p[0] = make_double2(o[0], o[1]);
      p[1] = make_double2(o[2], o[3]);
    }
  }
#ifdef CIP_POTRF_PROF
  __syncthreads();
  PROF_MARK(15);
  if (tid == 0 && j0 == 256) {
    for (int s = 0; s < 4; ++s)
      printf("  s=%d: diag start %lld | own update %lld | load %lld... D %lld | store %lld || w7 update done %lld | w3 after bar1 %lld | w3 below done %lld | iteration end %lld\n", s,
             dbg[s][0] - prof_t[0], 0LL, dbg[s][1] - dbg[s][0], dbg[s][2] - dbg[s][1], dbg[s][3] - dbg[s][2], dbg[s][4] - dbg[s][0],
             dbg[s][5] - dbg[s][0], dbg[s][6] - dbg[s][0], dbg[s][7] - dbg[s][0]);
    printf("potrf prof (cycles): load %lld |", prof_t[1] - prof_t[0]);
    for (int i = 0; i < 4; ++i) printf(" sweep%d %lld upd%d %lld |", i, prof_t[2 + 2 * i] - prof_t[1 + 2 * i], i, prof_t[3 + 2 * i] - prof_t[2 + 2 * i]);
    printf(" writeL %lld | invdiag %lld | inv1 %lld inv2 %lld inv3 %lld | writeW %lld | total %lld\n", prof_t[10] - prof_t[9],
           prof_t[11] - prof_t[10], prof_t[12] - prof_t[11], prof_t[13] - prof_t[12], prof_t[14] - prof_t[13], prof_t[15] - prof_t[14],
           prof_t[15] - prof_t[0]);
  }
#endif
}

// ---------------------------------------------------------------- panel head (the chain between two potrf's)
// After potrf_diag(jb) the next potrf only waits for ONE block row of work: T = A(jb+1, jb) inv(L_jj)' (the TRSM of
// that row) and A(jb+1, jb+1) -= T T' (its diagonal update).  As 128x128 tiles of the general GEMM kernel these are
// two single-CTA launches of ~20 us each -- a 128^3 product is 17 us on one SM at the DMMA peak -- so the chain
// between two factorisations was longer than the factorisation itself.  This kernel spreads the block row over a
// cluster of eight CTAs (16 rows each): phase 1 the strip of T (triangular: only k <= c contributes; the column
// groups are paired so that every scheduler gets the same share), cluster barrier, phase 2 the strip's rows of the
// diagonal update against the strips of the lower-numbered CTAs, read back through L2.  DMMA.8x8x4 fragments come
// straight from k-major shared arrays whose leading dimensions are = 8 mod 16, which spreads the 32 lanes of a
// fragment over all banks (two wavefronts for 256 bytes: the minimum).
constexpr int HEAD_CTAS = 8;
constexpr int HEAD_THREADS = 256;
constexpr int HS = NB / HEAD_CTAS;           // 16 rows per CTA
constexpr int WLD = NB + 8;                  // k-major inverse:  Ws[k * WLD + c] = inv(L)(c, k);  later Tf[c * WLD + row]
constexpr int SLD = HS + 8;                  // k-major strip:    As[k * SLD + i]
constexpr int HEAD_SMEM = (NB * WLD + NB * SLD) * (int)sizeof(double);
static_assert(WLD % 16 == 8 && SLD % 16 == 8, "bank-conflict-free DMMA fragments need leading dimensions = 8 mod 16");

__global__ void __cluster_dims__(HEAD_CTAS, 1, 1) __launch_bounds__(HEAD_THREADS, 1)
chol_head_kernel(double* H, int ld, int j0, const double* __restrict__ W) {
  extern __shared__ __align__(16) double hsm[];
  double* Ws = hsm;                          // [128 k][WLD]
  double* As = Ws + NB * WLD;                // [128 k][SLD]
  double* Tf = Ws;                           // [128 c][WLD]: T of the block row (own strip + the lower ones); phase 2 only
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, t = lane & 3;     // DMMA fragment coordinates
  const int q = blockIdx.x % HEAD_CTAS;      // strip of this CTA
  const int r0 = j0 + NB + HS * q;           // first matrix row of the strip
  // ---- loads (every load of a batch is issued before the first store: a load -> store loop is a chain of global
  //      latencies): the strip of A(jb+1, jb) and inv(L_jj), both transposed to k-major
  {
    double2 a0[2], a1[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {                              // strip of A: 16 rows x 32 quads
      const int e = tid + u * HEAD_THREADS, i = e & (HS - 1), kq = e >> 4;
      const double2* p = reinterpret_cast<const double2*>(H + ((size_t)(j0 / 4 + kq) * ld + r0 + i) * 4);
      a0[u] = p[0]; a1[u] = p[1];
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {                     // inv(L_jj): 128 rows x 32 quads, two batches of eight
      double2 w0[8], w1[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = tid + (half * 8 + u) * HEAD_THREADS, c = e & (NB - 1), kq = e >> 7;   // quad kq of row c: inv(L)(c, 4kq ..)
        const double2* p = reinterpret_cast<const double2*>(W + ((size_t)kq * NB + c) * 4);
        w0[u] = __ldg(p); w1[u] = __ldg(p + 1);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = tid + (half * 8 + u) * HEAD_THREADS, c = e & (NB - 1), kq = e >> 7;
        Ws[(4 * kq + 0) * WLD + c] = w0[u].x; Ws[(4 * kq + 1) * WLD + c] = w0[u].y;
        Ws[(4 * kq + 2) * WLD + c] = w1[u].x; Ws[(4 * kq + 3) * WLD + c] = w1[u].y;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int e = tid + u * HEAD_THREADS, i = e & (HS - 1), kq = e >> 4;
      As[(4 * kq + 0) * SLD + i] = a0[u].x; As[(4 * kq + 1) * SLD + i] = a0[u].y;
      As[(4 * kq + 2) * SLD + i] = a1[u].x; As[(4 * kq + 3) * SLD + i] = a1[u].y;
    }
  }
  __syncthreads();
  // ---- phase 1: T(i, c) = sum_{k <= c} A(i, k) inv(L)(c, k).  Warp = 16 columns x 16 rows (2 x 2 tiles); the
  //      triangular k range grows with the column group, so warps w and w + 4 (same scheduler) take groups whose
  //      ranges add up to the same total.
  {
    const int cg = (w < 4) ? 2 * w : 2 * (7 - w) + 1;
    const int c0 = 16 * cg;
    double acc[2][2][2];
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int y = 0; y < 2; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
    const int kq_end = (c0 + 15) >> 2;                          // last quad that meets a non-zero of these columns
#pragma unroll 4
    for (int kq = 0; kq <= kq_end; ++kq) {
      const double* ap = As + (4 * kq + t) * SLD + g;
      const double* wp = Ws + (4 * kq + t) * WLD + c0 + g;
      const double a0 = ap[0], a1 = ap[8], b0 = wp[0], b1 = wp[8];
      dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
      dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
      dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
      dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
    }
    __syncthreads();                                            // Ws is dead from here: Tf takes its place
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int y = 0; y < 2; ++y) {
        const int i = 8 * x + g, c = c0 + 8 * y + 2 * t;
        // tile (jb+1, jb) of the factor itself ...
        *reinterpret_cast<double2*>(H + ((size_t)((j0 + c) >> 2) * ld + r0 + i) * 4 + (c & 3)) =
            make_double2(acc[x][y][0], acc[x][y][1]);
        // ... and the own rows of Tf
        Tf[c * WLD + HS * q + i] = acc[x][y][0];
        Tf[(c + 1) * WLD + HS * q + i] = acc[x][y][1];
      }
  }
  // ---- every strip of T visible to the cluster (it was written to global memory above)
  __threadfence();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  // ---- phase 2 operands: the rows of the lower-numbered strips back from L2 (never read by this SM before, so no
  //      stale L1 line), and -- requested now, needed at the very end -- this warp's tiles of S = A(jb+1, jb+1)
  const int nct = 2 * (q + 1);                                  // 8-column tiles of S left of / on the diagonal
  double2 cv[2][2];
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int ct = w + 8 * u;
      cv[x][u] = make_double2(0.0, 0.0);
      if (ct < nct) {
        const int i = 8 * x + g, j = 8 * ct + 2 * t;
        cv[x][u] = *reinterpret_cast<const double2*>(H + ((size_t)((j0 + NB + j) >> 2) * ld + r0 + i) * 4 + (j & 3));
      }
    }
  for (int b0 = 0; b0 < 2 * q; b0 += 2) {                       // HS * q rows x 32 quads = 2 q rounds of 256, two per batch
    double2 t0[2], t1[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int e = tid + (b0 + u) * HEAD_THREADS, i = e % (HS * q), cq = e / (HS * q);
      const double2* p = reinterpret_cast<const double2*>(H + ((size_t)(j0 / 4 + cq) * ld + j0 + NB + i) * 4);
      t0[u] = __ldcg(p); t1[u] = __ldcg(p + 1);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int e = tid + (b0 + u) * HEAD_THREADS, i = e % (HS * q), cq = e / (HS * q);
      Tf[(4 * cq + 0) * WLD + i] = t0[u].x; Tf[(4 * cq + 1) * WLD + i] = t0[u].y;
      Tf[(4 * cq + 2) * WLD + i] = t1[u].x; Tf[(4 * cq + 3) * WLD + i] = t1[u].y;
    }
  }
  __syncthreads();
  // ---- phase 2: S(i, j) -= sum_c T(i, c) T(j, c) for the strip's rows i and the columns j < 16 (q + 1)
  if (w < nct) {
    const bool two = (w + 8) < nct;
    double acc[2][2][2];
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int u = 0; u < 2; ++u) acc[x][u][0] = acc[x][u][1] = 0.0;
#pragma unroll 4
    for (int cq = 0; cq < NB / 4; ++cq) {
      const double* tp = Tf + (4 * cq + t) * WLD + g;
      const double a0 = tp[HS * q], a1 = tp[HS * q + 8];
      const double b0 = tp[8 * w], b1 = two ? tp[8 * (w + 8)] : 0.0;
      dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
      dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
      if (two) {
        dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
        dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
      }
    }
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int ct = w + 8 * u;
        if (ct < nct) {
          const int i = 8 * x + g, j = 8 * ct + 2 * t;
          *reinterpret_cast<double2*>(H + ((size_t)((j0 + NB + j) >> 2) * ld + r0 + i) * 4 + (j & 3)) =
              make_double2(cv[x][u].x - acc[x][u][0], cv[x][u].y - acc[x][u][1]);
        }
      }
  }
}
std::atomic<unsigned long long> g_head_attr{0};

// ---------------------------------------------------------------- triangular sweeps (K3)
// One persistent launch per direction (forward L y = b, backward L' x = y) instead of one launch per
// 128-column panel.  Block row b of the triangular matrix is owned by CTAs ("units"): its off-diagonal tiles
// are cut into segments of at most `seg` tiles, one unit per segment; every unit streams its 128x128 tiles
// (coalesced 32-byte loads, the next half tile always in flight) and multiplies them with the solution blocks
// as those are PUBLISHED by the units that own them; the unit that holds the last tile of the row (the
// "finisher") adds the partial sums of the other segments in segment order (deterministic), applies the stored
// inverse of the diagonal block (already waiting in shared memory) and publishes its solution block.
//   hand-off: every published double is self-validating.  Output and partial-sum buffers are preset to the
//   all-ones bit pattern (a NaN that no arithmetic produces: results are canonicalised before they are stored)
//   and consumers poll the value itself, so a hand-off costs ONE L2 round trip, no flag + data pair, no fence.
//   no deadlock: units are ordered by the last solution block they need, so a unit only ever waits for units
//   with a smaller index, which the hardware dispatches first; every poll also gives up after ~2^26 tries and
//   raises *err, so that a bug cannot hang the device.
// The chain per block row is: poll (1 L2 trip) -> 128x128 product from registers -> partials -> 128x128 product
// with the inverse from shared memory -> publish; everything else (HBM streaming of L, once per sweep) hides
// behind it.  Replaces `Z\[dy;dw]` at src/kktsolvers.jl:299 (LAPACK / UMFPACK triangular solves).
constexpr int SWEEP_THREADS = 512;
constexpr int SWEEP_SMEM = NB * NB * (int)sizeof(double);      // inverse of the diagonal block
constexpr unsigned long long SENTINEL = 0xFFFFFFFFFFFFFFFFull;
constexpr int SPIN_LIMIT = 1 << 26;

struct SweepUnit { int blk, d0, nd, seg, nseg, fin; };   // block row, first dependency, #dependencies (in sweep order)

__device__ __forceinline__ double canon(double v) {           // never store the sentinel bit pattern
  return (v != v) ? __longlong_as_double(0x7FF8000000000000ll) : v;
}
__device__ __forceinline__ void publish(double* p, double v) {
  asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(p), "d"(canon(v)) : "memory");
}
__device__ __forceinline__ double poll(const double* p, int* err) {
  unsigned long long v;
  int spins = 0;
  do {
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (v != SENTINEL) break;
    if (++spins > SPIN_LIMIT) { atomicExch(err, 1); v = 0x7FF8000000000000ull; break; }
  } while (true);
  return __longlong_as_double((long long)v);
}

// FWD: sol = y, rhs = b:  y_b = inv(L_bb) (b_b - sum_{d<b} L_bd y_d),        dependencies d ascending
// BWD: sol = x, rhs = y:  x_b = inv(L_bb)' (y_b - sum_{d>b} L_db' x_d),      dependencies d descending
template <bool FWD>
__global__ void __launch_bounds__(SWEEP_THREADS, 1)
trsv_sweep_kernel(const double* __restrict__ H, int ld, const double* __restrict__ Winv,
                  const SweepUnit* __restrict__ units, const double* __restrict__ rhs, double* sol,
                  double* part, int maxseg, int* err) {
  extern __shared__ __align__(16) double Ws[];                 // [32 quads][128 rows][4]: inv(L_bb), Q4 with ld = 128
  __shared__ double vs[NB];                                    // the dependency's solution block
  __shared__ double red[4][NB];
  __shared__ double tot[NB];
  const SweepUnit u = units[blockIdx.x];
  const int tid = threadIdx.x;
  if (u.fin) {
    // inverse of the diagonal block -> shared memory, asynchronously (needed only at the very end)
    const double* W = Winv + (size_t)u.blk * NB * NB;
#pragma unroll
    for (int i = 0; i < (NB * NB * 8) / (SWEEP_THREADS * 16); ++i) {
      const int o = (i * SWEEP_THREADS + tid) * 2;             // doubles
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(Ws + o)), "l"(W + o) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // thread -> (row or quad, part) of a 128x128 tile; `base` = first 32-byte element of this thread in tile 0 of
  // the row, `tstep` = distance between consecutive dependencies' tiles
  //   FWD: tile(b, d) = rows 128b.., k-quads 32d..      thread (r = tid & 127, pq = tid >> 7): quads pq*8 .. +7
  //   BWD: tile(d, b) = rows 128d.., k-quads 32b..      thread (q = tid >> 4, pr = tid & 15): rows pr + 16 i
  const int r = tid & 127, pq = tid >> 7, q = tid >> 4, pr = tid & 15;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  double2 cur[8], nxt[8];
  auto tile_ptr = [&](int d, int half) -> const double2* {
    if (FWD) return reinterpret_cast<const double2*>(H + ((size_t)(d * 32 + pq * 8 + half * 4) * ld + (size_t)u.blk * NB + r) * 4);
    return reinterpret_cast<const double2*>(H + ((size_t)(u.blk * 32 + q) * ld + (size_t)d * NB + pr + 64 * half) * 4);
  };
  auto load_half = [&](double2 (&buf)[8], int d, int half) {
    const double2* p = tile_ptr(d, half);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const size_t o = FWD ? (size_t)i * ld * 2 : (size_t)i * 32;      // in double2: next quad / 16 rows further
      buf[2 * i] = __ldg(p + o);
      buf[2 * i + 1] = __ldg(p + o + 1);
    }
  };
  auto fma_half = [&](const double2 (&buf)[8], int half) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (FWD) {
        const double* v = vs + 4 * (pq * 8 + half * 4 + i);
        a0 = fma(buf[2 * i].x, v[0], a0); a1 = fma(buf[2 * i].y, v[1], a1);
        a2 = fma(buf[2 * i + 1].x, v[2], a2); a3 = fma(buf[2 * i + 1].y, v[3], a3);
      } else {
        const double v = vs[pr + 64 * half + 16 * i];
        a0 = fma(buf[2 * i].x, v, a0); a1 = fma(buf[2 * i].y, v, a1);
        a2 = fma(buf[2 * i + 1].x, v, a2); a3 = fma(buf[2 * i + 1].y, v, a3);
      }
    }
  };
  const int dstep = FWD ? 1 : -1;
  // finisher, off the chain: its right-hand side block and (before the last dependency is awaited) the partial
  // sums of the other segments of the row, which were published long ago
  double myrhs = 0.0, sib = 0.0;
  if (u.fin && tid < NB) myrhs = rhs[(size_t)u.blk * NB + tid];
  auto gather_siblings = [&]() {
    if (!(u.fin && tid < NB)) return;
    const double* pp = part + (size_t)u.blk * maxseg * NB + tid;
    for (int sg = 0; sg < u.nseg - 1; sg += 4) {
      // four independent loads in flight, then the (normally not taken) wait for those that were not there yet
      unsigned long long w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        w[k] = 0;
        if (sg + k < u.nseg - 1)
          asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w[k]) : "l"(pp + (size_t)(sg + k) * NB) : "memory");
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (sg + k < u.nseg - 1) {
          const double v = (w[k] == SENTINEL) ? poll(pp + (size_t)(sg + k) * NB, err) : __longlong_as_double((long long)w[k]);
          sib += v;                                             // segment order: deterministic
        }
      }
    }
  };
  if (u.nd == 0) gather_siblings();
  if (u.nd > 0) load_half(cur, u.d0, 0);
  for (int t = 0; t < u.nd; ++t) {
    const int d = u.d0 + t * dstep;
    load_half(nxt, d, 1);
    if (t == u.nd - 1) gather_siblings();
    if (tid < NB) vs[tid] = poll(sol + (size_t)d * NB + tid, err);
    __syncthreads();
    fma_half(cur, 0);
    if (t + 1 < u.nd) load_half(cur, d + dstep, 0);
    fma_half(nxt, 1);
    __syncthreads();                                           // vs is rewritten in the next round
  }
  // ---- reduce the partial sums of this unit to one value per row (FWD) / column (BWD) of the block
  if (FWD) {
    red[pq][r] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (tid < NB) tot[tid] = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
    if (u.fin && tid < NB) vs[tid] = myrhs - (sib + tot[tid]);
  } else {
#pragma unroll
    for (int o = 1; o <= 8; o <<= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
    }
    if (pr == 0) { tot[4 * q] = a0; tot[4 * q + 1] = a1; tot[4 * q + 2] = a2; tot[4 * q + 3] = a3; }
  }
  if (u.fin) asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  double* mypart = part + ((size_t)u.blk * maxseg + u.seg) * NB;
  if (!u.fin) {
    if (tid < NB) publish(mypart + tid, tot[tid]);
    return;
  }
  // ---- finisher: rhs - (other segments in segment order + its own); inverse of the diagonal block
  if (!FWD) {
    if (tid < NB) vs[tid] = myrhs - (sib + tot[tid]);
    __syncthreads();
  }
  if (FWD) {
    // y[r] = sum_c W[r, c] v[c]
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int qq = pq * 8 + i;
      const double2 w0 = *reinterpret_cast<const double2*>(Ws + ((size_t)qq * NB + r) * 4);
      const double2 w1 = *reinterpret_cast<const double2*>(Ws + ((size_t)qq * NB + r) * 4 + 2);
      const double* v = vs + 4 * qq;
      s0 = fma(w0.x, v[0], s0); s1 = fma(w0.y, v[1], s1); s2 = fma(w1.x, v[2], s2); s3 = fma(w1.y, v[3], s3);
    }
    red[pq][r] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (tid < NB) publish(sol + (size_t)u.blk * NB + tid, (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]));
  } else {
    // x[c] = sum_r W[r, c] v[r]
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = pr + 16 * i;
      const double2 w0 = *reinterpret_cast<const double2*>(Ws + ((size_t)q * NB + rr) * 4);
      const double2 w1 = *reinterpret_cast<const double2*>(Ws + ((size_t)q * NB + rr) * 4 + 2);
      const double v = vs[rr];
      s0 = fma(w0.x, v, s0); s1 = fma(w0.y, v, s1); s2 = fma(w1.x, v, s2); s3 = fma(w1.y, v, s3);
    }
#pragma unroll
    for (int o = 1; o <= 8; o <<= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    }
    if (pr == 0) {
      double* o4 = sol + (size_t)u.blk * NB + 4 * q;
      publish(o4, s0); publish(o4 + 1, s1); publish(o4 + 2, s2); publish(o4 + 3, s3);
    }
  }
}

std::atomic<unsigned long long> g_sweep_attr[2];
std::atomic<unsigned long long> g_potrf_attr{0};
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
  CIP_CUDA(cudaStreamCreateWithPriority(&p->sd, cudaStreamNonBlocking, hi));
  CIP_CUDA(cudaStreamCreateWithPriority(&p->se, cudaStreamNonBlocking, hi));
  CIP_CUDA(cudaStreamCreateWithPriority(&p->sf, cudaStreamNonBlocking, hi));
  CIP_CUDA(cudaStreamCreateWithPriority(&p->sb, cudaStreamNonBlocking, hi));
  for (auto* e : {&p->evT[0], &p->evT[1], &p->evR[0], &p->evR[1], &p->evS, &p->evD[0], &p->evD[1], &p->evP, &p->evE,
                  &p->evC, &p->evH, &p->evUc[0], &p->evUc[1], &p->evUo[0], &p->evUo[1]})
    CIP_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  // ---- work units of the persistent triangular sweeps (see trsv_sweep_kernel)
  {
    const int np = p->npanels;
    const int seg = std::min(32, std::max(8, (np + 3) / 4));
    p->maxseg = std::max(1, (np - 1 + seg - 1) / seg);
    std::vector<SweepUnit> uf, ub;
    struct Keyed { int key, fin, blk; SweepUnit u; };
    for (int dir = 0; dir < 2; ++dir) {
      std::vector<Keyed> v;
      for (int t = 0; t < np; ++t) {                 // t = position of the block row in sweep order
        const int nseg = std::max(1, (t + seg - 1) / seg);
        for (int sg = 0; sg < nseg; ++sg) {
          const int t0 = sg * seg, t1 = std::min(t, (sg + 1) * seg);      // dependencies [t0, t1) in sweep order
          SweepUnit u;
          u.blk = dir == 0 ? t : np - 1 - t;
          u.d0 = dir == 0 ? t0 : np - 1 - t0;
          u.nd = t1 - t0;
          u.seg = sg; u.nseg = nseg; u.fin = (sg == nseg - 1);
          v.push_back(Keyed{t1 - 1, u.fin, t, u});   // ordered by the last solution block the unit waits for
        }
      }
      std::stable_sort(v.begin(), v.end(), [](const Keyed& a, const Keyed& b) {
        if (a.key != b.key) return a.key < b.key;
        if (a.fin != b.fin) return a.fin > b.fin;
        return a.blk < b.blk;
      });
      for (auto& k : v) (dir == 0 ? uf : ub).push_back(k.u);
    }
    p->nunits = (int)uf.size();
    CIP_CUDA(cudaMalloc(&p->units_fwd, sizeof(SweepUnit) * uf.size()));
    CIP_CUDA(cudaMalloc(&p->units_bwd, sizeof(SweepUnit) * ub.size()));
    CIP_CUDA(cudaMemcpy(p->units_fwd, uf.data(), sizeof(SweepUnit) * uf.size(), cudaMemcpyHostToDevice));
    CIP_CUDA(cudaMemcpy(p->units_bwd, ub.data(), sizeof(SweepUnit) * ub.size(), cudaMemcpyHostToDevice));
    CIP_CUDA(cudaMalloc(&p->sweep_part, sizeof(double) * (size_t)np * p->maxseg * NB));
    CIP_CUDA(cudaMalloc(&p->sweep_err, sizeof(int)));
    CIP_CUDA(cudaMemset(p->sweep_err, 0, sizeof(int)));
  }
  return 0;
}

void chol_free_plan(CholPlan* p) {
  if (p->sc) cudaStreamDestroy(p->sc);
  if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
  p->graph_exec = nullptr;
  if (p->sd) cudaStreamDestroy(p->sd);
  if (p->se) cudaStreamDestroy(p->se);
  if (p->sf) cudaStreamDestroy(p->sf);
  if (p->sb) cudaStreamDestroy(p->sb);
  for (auto e : {p->evT[0], p->evT[1], p->evR[0], p->evR[1], p->evS, p->evD[0], p->evD[1], p->evP, p->evE, p->evC, p->evH,
                 p->evUc[0], p->evUc[1], p->evUo[0], p->evUo[1]})
    if (e) cudaEventDestroy(e);
  p->sc = nullptr; p->sd = nullptr; p->se = nullptr; p->sb = nullptr; p->sf = nullptr;
  for (void* q : {(void*)p->units_fwd, (void*)p->units_bwd, (void*)p->sweep_part, (void*)p->sweep_err})
    if (q) cudaFree(q);
  p->units_fwd = p->units_bwd = nullptr; p->sweep_part = nullptr; p->sweep_err = nullptr;
}

namespace {
// Whether the next outer panel's block column is fed inner panel by inner panel (four K = 128 updates beside the
// chain) or gets one K = 512 update after the outer panel is complete.  Feeding shortens the panel chain but
// does the same flops on less efficient tiles; it pays when the factorisation is chain-bound (small n, or the
// trailing matrix shared by several ranks).  Both Cholesky variants take the same decision from the same
// inputs, so single-GPU and distributed factors stay bit-identical.
bool feed_next_panel(int npanels, int nranks) {
  if (const char* env = getenv("CIP_CHOL_PIPE")) return atoi(env) != 0;
  return npanels <= 64 || nranks >= 4;
}

// Factorisation of the outer panel [J0, J1) of 128-column panels on the chain stream sc.  Per inner panel:
// potrf_diag (L11 and its inverse), L21 = A21 inv(L11)' for all rows below, then the update of the NEXT inner
// panel's column only -- that is all the next potrf waits for.  The update of the remaining columns of the outer
// panel runs beside the chain on p.sd (inner look-ahead).  `hook(jb)` is called as soon as column jb is final
// (p.evP has just been recorded on sc behind it): the callers feed the next outer panel / the other ranks from it.
template <typename Hook>
int factor_outer_panel(const CholPlan& p, int J0, int J1, cudaStream_t sc, int smem, Hook hook) {
  const int np = p.npanels;
  static const bool use_head = [] { const char* e = getenv("CIP_CHOL_HEAD"); return !e || atoi(e) != 0; }();
  CIP_TRY(ensure_dyn_smem((const void*)chol_head_kernel, HEAD_SMEM, &g_head_attr));
  cudaStream_t sd = p.sd, sf = p.sf;
  auto trsm = [&](int jb, int first_row_panel, cudaStream_t st) -> int {     // L(i, jb) = A(i, jb) inv(L_jj)' for block rows i >= first
    const int j0 = jb * NB, ntm = np - first_row_panel;
    if (ntm <= 0) return 0;
    GemmArgs t{};
    t.lower = 0; t.ntm = ntm; t.ntn = 1; t.sym = 0;
    t.x_row0 = first_row_panel * NB; t.y_row0 = 0; t.x_kq0 = j0 / 4; t.y_kq0 = 32 * jb; t.nk = NB / 32;
    t.Cin = nullptr; t.Cout = p.H; t.ldc = p.ld; t.c_row0 = first_row_panel * NB; t.c_col0 = j0; t.alpha = 1.0;
    t.y_lower_tri = 1;
    return launch_gemm_nt(p.mapH, p.mapWinv, t, st);
  };
  auto update = [&](int jb, int row_panel, int col_panel0, int ncols, cudaStream_t st) -> int {   // rank-128 update by panel jb
    const int j0 = jb * NB, ntm = np - row_panel;
    if (ntm <= 0 || ncols <= 0) return 0;
    GemmArgs c{};
    c.lower = 0; c.ntm = ntm; c.ntn = ncols; c.sym = 0;
    c.x_row0 = row_panel * NB; c.y_row0 = col_panel0 * NB; c.x_kq0 = j0 / 4; c.y_kq0 = j0 / 4; c.nk = NB / 32;
    c.Cin = p.H; c.Cout = p.H; c.ldc = p.ld; c.c_row0 = row_panel * NB; c.c_col0 = col_panel0 * NB; c.alpha = -1.0;
    return launch_gemm_nt(p.mapH, p.mapH, c, st);
  };
  bool uc_prev = false, uo_prev = false;       // off-chain updates of the previous inner panel still to be awaited
  for (int jb = J0; jb < J1; ++jb) {
    const int j0 = jb * NB;
    const int rem = np - jb - 1;
    const int inner_cols = J1 - jb - 1;
    potrf_diag_kernel<<<1, PT, smem, sc>>>(p.H, p.ld, j0, p.Winv + (size_t)jb * NB * NB, p.info);
    CIP_CHECK_LAUNCH();
    CIP_CUDA(cudaEventRecord(p.evP, sc));
    // everything below reads tiles that the previous panel's off-chain updates wrote
    if (uc_prev) CIP_CUDA(cudaStreamWaitEvent(sc, p.evUc[(jb - 1) & 1], 0));
    if (uo_prev) CIP_CUDA(cudaStreamWaitEvent(sc, p.evUo[(jb - 1) & 1], 0));
    const bool uo_before = uo_prev;
    const bool head = use_head && inner_cols >= 1 && rem >= 1;
    if (!head) {
      // last inner panel of the outer panel (or head disabled): the whole column on the chain, as one TRSM
      CIP_TRY(trsm(jb, jb + 1, sc));
      CIP_CUDA(cudaEventRecord(p.evC, sc));
      CIP_TRY(hook(jb));
      uc_prev = uo_prev = false;
      if (rem == 0 || inner_cols <= 0) continue;
      if (inner_cols > 1) {                     // (head disabled) other columns beside the chain, next column on it
        CIP_CUDA(cudaStreamWaitEvent(sf, p.evC, 0));
        CIP_TRY(update(jb, jb + 2, jb + 2, inner_cols - 1, sf));
        CIP_CUDA(cudaEventRecord(p.evUo[jb & 1], sf));
        uo_prev = true;
      }
      CIP_TRY(update(jb, jb + 1, jb + 1, 1, sc));
      continue;
    }
    // ---- chain: block row jb+1 only (its TRSM tile and its diagonal update), on a cluster of four CTAs
    chol_head_kernel<<<HEAD_CTAS, HEAD_THREADS, HEAD_SMEM, sc>>>(p.H, p.ld, j0, p.Winv + (size_t)jb * NB * NB);
    CIP_CHECK_LAUNCH();
    CIP_CUDA(cudaEventRecord(p.evH, sc));
    // ---- beside the chain: the rest of the column, then the updates that need it
    CIP_CUDA(cudaStreamWaitEvent(sd, p.evP, 0));
    CIP_TRY(trsm(jb, jb + 2, sd));                                   // rows jb+2 ..
    CIP_CUDA(cudaStreamWaitEvent(sd, p.evH, 0));                     // + L(jb+1, jb) from the head: column jb is final
    CIP_CUDA(cudaEventRecord(p.evC, sd));
    CIP_TRY(hook(jb));
    uc_prev = uo_prev = false;
    if (rem >= 2) {
      if (uo_before) CIP_CUDA(cudaStreamWaitEvent(sd, p.evUo[(jb - 1) & 1], 0));   // column jb+1 was last written there
      CIP_TRY(update(jb, jb + 2, jb + 1, 1, sd));                    // column jb+1, rows jb+2 ..
      CIP_CUDA(cudaEventRecord(p.evUc[jb & 1], sd));
      uc_prev = true;
      if (inner_cols >= 2) {
        CIP_CUDA(cudaStreamWaitEvent(sf, p.evC, 0));
        CIP_TRY(update(jb, jb + 2, jb + 2, inner_cols - 1, sf));     // columns jb+2 .. J1-1, rows jb+2 ..
        CIP_CUDA(cudaEventRecord(p.evUo[jb & 1], sf));
        uo_prev = true;
      }
    }
  }
  return 0;
}
// C[rows >= C0 (first row of the column block), columns [C0, C1)] -= L[rows, K0..K1) * L[cols, K0..K1)'   (panel units)
int update_block_column(const CholPlan& p, int C0, int C1, int K0, int K1, cudaStream_t st) {
  const int np = p.npanels;
  GemmArgs c{};
  c.lower = 0; c.ntm = np - C0; c.ntn = C1 - C0; c.sym = 0;
  c.x_row0 = C0 * NB; c.y_row0 = C0 * NB; c.x_kq0 = K0 * NB / 4; c.y_kq0 = K0 * NB / 4; c.nk = (K1 - K0) * NB / 32;
  c.Cin = p.H; c.Cout = p.H; c.ldc = p.ld; c.c_row0 = C0 * NB; c.c_col0 = C0 * NB; c.alpha = -1.0;
  return launch_gemm_nt(p.mapH, p.mapH, c, st);
}
}  // namespace

// Two-level blocked right-looking Cholesky with look-ahead at both levels.
//   outer panel = OUTER inner panels (512 columns).  Inside an outer panel every 128-wide inner panel is
//   factored (potrf_diag), solved against all rows below (TRSM as a GEMM with inv(L11)) and applied to the
//   remaining columns of the outer panel (factor_outer_panel).
//   The block column of the NEXT outer panel receives the rank-128 update of every inner panel as soon as that
//   panel is final (stream se, beside the chain), so that when the outer panel is done only the last of those
//   small updates stands between it and the next potrf -- not a K = 512 update of the whole block column.
//   The rest of the trailing matrix gets ONE update with K = 512 per outer panel on the caller's stream, which
//   amortises the C-tile read/write and the pipeline fill that dominate K = 128 tiles.
int chol_factor(const CholPlan& p, cudaStream_t s) {
  // The launch sequence below is static for a plan (same kernels, same arguments, same stream fork / join
  // pattern every time): it is captured once into a CUDA graph and replayed, which takes the per-launch and
  // per-event-dependency latencies of ~600 stream operations off the panel chain.
  static const bool use_graph = [] { const char* e = getenv("CIP_CHOL_GRAPH"); return !e || atoi(e) != 0; }();
  if (use_graph && !p.capturing) {
    CholPlan& mp = const_cast<CholPlan&>(p);
    if (!mp.graph_exec && !mp.graph_failed) {
      cudaGraph_t g = nullptr;
      mp.capturing = true;
      cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
      int rc = -1;
      if (e == cudaSuccess) {
        rc = chol_factor(p, s);
        e = cudaStreamEndCapture(s, &g);
      }
      mp.capturing = false;
      if (e == cudaSuccess && rc == 0 && g) e = cudaGraphInstantiate(&mp.graph_exec, g, 0);
      if (g) cudaGraphDestroy(g);
      if (e != cudaSuccess || rc != 0 || !mp.graph_exec) {
        cudaGetLastError();
        mp.graph_exec = nullptr;
        mp.graph_failed = true;                 // fall back to plain stream launches for this plan
      }
    }
    if (mp.graph_exec) {
      CIP_CUDA(cudaGraphLaunch(mp.graph_exec, s));
      g_launches += mp.graph_kernels;
      return 0;
    }
  }
  const long long launches_before = g_launches.load();
  struct CountKernels { const CholPlan& p; long long before; ~CountKernels() { if (p.capturing) const_cast<CholPlan&>(p).graph_kernels = g_launches.load() - before; } } counter{p, launches_before};
  int OUTER = 4;
  if (const char* env = getenv("CIP_CHOL_OUTER")) { const int v = atoi(env); if (v >= 1 && v <= 8) OUTER = v; }
  const int smem = POTRF_SMEM;
  CIP_TRY(ensure_dyn_smem((const void*)potrf_diag_kernel, smem, &g_potrf_attr));
  cudaStream_t sc = p.sc, se = p.se;
  const int np = p.npanels;
  const bool pipe = feed_next_panel(np, p.nranks_hint);
  CIP_CUDA(cudaMemsetAsync(p.info, 0, sizeof(int), s));
  CIP_CUDA(cudaEventRecord(p.evS, s));
  CIP_CUDA(cudaStreamWaitEvent(sc, p.evS, 0));
  CIP_CUDA(cudaStreamWaitEvent(se, p.evS, 0));
  int outer = 0;
  for (int J0 = 0; J0 < np; J0 += OUTER, ++outer) {
    const int J1 = (J0 + OUTER < np) ? J0 + OUTER : np;
    const int rem = np - J1;
    const int N1 = (J1 + OUTER < np) ? J1 + OUTER : np;      // the next outer panel is [J1, N1)
    // its block column was last written by the bulk update of the previous outer panel
    if (outer > 0 && rem > 0) CIP_CUDA(cudaStreamWaitEvent(se, p.evR[(outer - 1) & 1], 0));
    CIP_TRY(factor_outer_panel(p, J0, J1, sc, smem, [&](int jb) -> int {
      if (rem <= 0 || !pipe) return 0;
      CIP_CUDA(cudaStreamWaitEvent(se, p.evC, 0));
      return update_block_column(p, J1, N1, jb, jb + 1, se);
    }));
    if (rem > 0) {
      if (!pipe) {                                                         // one K = 512 update instead
        CIP_CUDA(cudaEventRecord(p.evP, sc));
        CIP_CUDA(cudaStreamWaitEvent(se, p.evP, 0));
        CIP_TRY(update_block_column(p, J1, N1, J0, J1, se));
      }
      CIP_CUDA(cudaEventRecord(p.evE, se));
      CIP_CUDA(cudaStreamWaitEvent(sc, p.evE, 0));
    }
    CIP_CUDA(cudaEventRecord(p.evT[outer & 1], sc));
    if (rem <= 0) break;
    // bulk trailing update on the caller's stream (K = 512): everything right of the next outer panel
    CIP_CUDA(cudaStreamWaitEvent(s, p.evT[outer & 1], 0));
    if (rem > N1 - J1) {
      const int b0 = N1 * NB, k0 = J0 * NB, kw = (J1 - J0) * NB;
      GemmArgs u{};
      u.lower = 1; u.ntm = np - N1; u.ntn = np - N1; u.sym = 1;
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
// Every rank holds the full matrix buffer; outer panel J is factored by rank J % N and each of its inner panels
// is broadcast (stream sb) the moment it is final, while the owner's chain goes on.  The rank that factors the
// NEXT outer panel applies every inner panel to that block column as it arrives (rank-128 updates on its chain
// stream), exactly as the single-GPU schedule does on `se`; all other owned block columns get one K = 512 update
// per outer panel.  Every tile therefore receives the same sequence of updates as in chol_factor and the factor
// is bit-identical to the single-GPU one.
int chol_factor_dist(const CholPlan& p, cudaStream_t s, const CholDist& d) {
  // (Replaying this sequence as a CUDA graph, as chol_factor does, was tried in round 2: capturing the NCCL broadcasts
  //  fails with "operation not permitted when stream is capturing" under NCCL 2.28 and leaves a single-process handle
  //  hung, so the distributed variant stays on plain stream launches.)
  int OUTER = 4;
  if (const char* env = getenv("CIP_CHOL_OUTER")) { const int v = atoi(env); if (v >= 1 && v <= 8) OUTER = v; }
  const NcclApi* api = nccl_api();
  if (!api) return -1;
  const int smem = POTRF_SMEM;
  CIP_TRY(ensure_dyn_smem((const void*)potrf_diag_kernel, smem, &g_potrf_attr));
  cudaStream_t sc = p.sc, sb = p.sb;
  const int np = p.npanels, NO = (np + OUTER - 1) / OUTER, N = d.nranks, me = d.rank;
  const bool pipe = feed_next_panel(np, N);
  CIP_CUDA(cudaMemsetAsync(p.info, 0, sizeof(int), s));
  CIP_CUDA(cudaEventRecord(p.evS, s));
  CIP_CUDA(cudaStreamWaitEvent(sc, p.evS, 0));
  CIP_CUDA(cudaStreamWaitEvent(sb, p.evS, 0));
  auto bcast_inner = [&](int jb, int owner) -> int {        // column jb (all rows: contiguous in Q4) + inv(L_jj)
    double* col = p.H + (size_t)(jb * NB / 4) * p.ld * 4;
    double* w = p.Winv + (size_t)jb * NB * NB;
    int r = api->GroupStart();
    if (r == 0) r = api->Broadcast(col, col, (size_t)NB * p.ld, kNcclFloat64, owner, d.comm, sb);
    if (r == 0) r = api->Broadcast(w, w, (size_t)NB * NB, kNcclFloat64, owner, d.comm, sb);
    const int r2 = api->GroupEnd();
    if (r == 0) r = r2;
    if (r != 0) {
      set_error("ncclBroadcast failed in the distributed Cholesky: %s", api->GetErrorString(r));
      return -1;
    }
    return 0;
  };
  int myNext = me;                          // the next outer panel this rank will factor; its updates run on sc
  for (int J = 0; J < NO; ++J) {
    const int owner = J % N;
    const int J0 = J * OUTER, J1 = (J0 + OUTER < np) ? J0 + OUTER : np;
    const int N1 = (J1 + OUTER < np) ? J1 + OUTER : np;
    const bool feed_next = (myNext == J + 1) && J1 < np;     // this rank factors the next outer panel
    if (owner == me) {
      CIP_TRY(factor_outer_panel(p, J0, J1, sc, smem, [&](int jb) -> int {
        CIP_CUDA(cudaStreamWaitEvent(sb, p.evC, 0));
        return bcast_inner(jb, owner);
      }));
    } else {
      for (int jb = J0; jb < J1; ++jb) {
        CIP_TRY(bcast_inner(jb, owner));
        if (feed_next && pipe) {
          CIP_CUDA(cudaEventRecord(p.evE, sb));
          CIP_CUDA(cudaStreamWaitEvent(sc, p.evE, 0));
          CIP_TRY(update_block_column(p, J1, N1, jb, jb + 1, sc));
        }
      }
    }
    CIP_CUDA(cudaEventRecord(p.evT[J & 1], sb));            // the whole outer panel is here (and, on the owner, factored)
    CIP_CUDA(cudaStreamWaitEvent(sc, p.evT[J & 1], 0));
    if (owner == me) {
      myNext += N;
      // the new "next" panel has so far been updated on s: order sc after everything s has queued
      CIP_CUDA(cudaEventRecord(p.evR[0], s));
      CIP_CUDA(cudaStreamWaitEvent(sc, p.evR[0], 0));
    }
    // look-ahead block column further away than J+1: one K = 512 update, first, on the chain stream
    if (myNext < NO && myNext > J && !(feed_next && pipe))
      CIP_TRY(update_block_column(p, myNext * OUTER, std::min(np, (myNext + 1) * OUTER), J0, J1, sc));
    CIP_CUDA(cudaStreamWaitEvent(s, p.evT[J & 1], 0));
    for (int Jc = me; Jc < NO; Jc += N) {   // owned panels right of J, except the look-ahead one
      if (Jc <= J || Jc == myNext) continue;
      CIP_TRY(update_block_column(p, Jc * OUTER, std::min(np, (Jc + 1) * OUTER), J0, J1, s));
    }
  }
  CIP_CUDA(cudaEventRecord(p.evS, sc));
  CIP_CUDA(cudaStreamWaitEvent(s, p.evS, 0));
  CIP_CUDA(cudaEventRecord(p.evE, sb));
  CIP_CUDA(cudaStreamWaitEvent(s, p.evE, 0));
  // a failed pivot is only known to the owner of that panel
  int r = api->AllReduce(p.info, p.info, 1, kNcclInt32, kNcclMax, d.comm, s);
  if (r != 0) {
    set_error("ncclAllReduce(info) failed: %s", api->GetErrorString(r));
    return -1;
  }
  return 0;
}

namespace {
template <bool FWD>
int run_sweep(const CholPlan& p, const double* rhs, double* sol, cudaStream_t s) {
  CIP_TRY(ensure_dyn_smem((const void*)trsv_sweep_kernel<FWD>, SWEEP_SMEM, &g_sweep_attr[FWD ? 0 : 1]));
  // preset the hand-off buffers to the sentinel (all-ones bytes)
  CIP_CUDA(cudaMemsetAsync(sol, 0xFF, sizeof(double) * (size_t)p.npanels * NB, s));
  CIP_CUDA(cudaMemsetAsync(p.sweep_part, 0xFF, sizeof(double) * (size_t)p.npanels * p.maxseg * NB, s));
  trsv_sweep_kernel<FWD><<<p.nunits, SWEEP_THREADS, SWEEP_SMEM, s>>>(
      p.H, p.ld, p.Winv, reinterpret_cast<const SweepUnit*>(FWD ? p.units_fwd : p.units_bwd), rhs, sol, p.sweep_part,
      p.maxseg, p.sweep_err);
  CIP_CHECK_LAUNCH();
  return 0;
}
}  // namespace

// y receives the solution of L y = b (b is read only; the two must not alias)
int chol_fwd(const CholPlan& p, double* b, double* y, cudaStream_t s) { return run_sweep<true>(p, b, y, s); }
// x receives the solution of L' x = y
int chol_bwd(const CholPlan& p, double* y, double* x, cudaStream_t s) { return run_sweep<false>(p, y, x, s); }

}  // namespace cip

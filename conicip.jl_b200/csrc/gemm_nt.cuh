// gemm_nt: the one FP64 tensor-core tile kernel behind K1 (scaled SYRK), K2 (Cholesky
// panel TRSM + trailing update) and the Schur-complement block.
//
//     C[i, j] = Cin[i, j] + alpha * sum_k X[i, k] * Y[j, k]       ("NT": both operands K-major)
//
// X, Y, Cin, Cout are all in the Q4 quad-interleaved layout (common.cuh).
// Replaces the reference's BLAS calls at src/kktsolvers.jl:33-34 (`Atil'Atil` -> dsyrk),
// :290 (`A'*(F^-T'*(F^-T*A))`) and the LAPACK factorisation at :35 / :295.
#pragma once
#include "common.cuh"

namespace cip {

struct GemmOperand {
  CUtensorMap map;  // 2-D: dim0 = ld*4 doubles (quad-interleaved rows), dim1 = K/4 quads
};

struct GemmArgs {
  // tile grid
  int lower;     // 1: only tiles ti >= tj of an nt x nt grid (SYRK / trailing update); 0: ntm x ntn
  int ntm, ntn;  // tile counts (lower: ntm == ntn)
  int sym;       // 1: X and Y are the same operand with the same origin (diagonal tiles load once)
  // operand origins
  int x_row0, y_row0;  // first row of tile (0, .) in X / Y
  int x_kq0, y_kq0;    // first k-quad in X / Y
  int nk;              // number of KT-wide k tiles
  // C
  const double* Cin;   // may be nullptr (treated as 0)
  double* Cout;
  int ldc;             // Q4 ld of C
  int c_row0, c_col0;  // C origin
  double alpha;
};

// Creates a tensor map over a Q4 matrix (ld rows, kq_total quads).
int make_q4_tensor_map(CUtensorMap* out, const double* base, int ld, long long kq_total);

// Launch on `stream`.
int launch_gemm_nt(const GemmOperand& X, const GemmOperand& Y, const GemmArgs& a, cudaStream_t stream);

int gemm_nt_smem_bytes();

}  // namespace cip

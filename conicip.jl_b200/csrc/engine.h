// Private definition of the engine handle, shared by engine.cu (C ABI) and ipm.cu (native
// interior-point driver).
#pragma once
#include <vector>

#include "../../include/conicip_b200.h"
#include "kernels.cuh"

constexpr int NV = 8;   // n-length work vectors
constexpr int MV = 10;  // m-length work vectors
constexpr int PV = 6;   // p-length work vectors

struct cip_engine {
  int n = 0, m = 0, p = 0, n_pad = 0, m_pad = 0, p_pad = 0, ncones = 0;
  int device = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  cip_options opt{};
  // matrices (Q4 layout)
  double *At4 = nullptr, *Atil4 = nullptr, *Qq4 = nullptr, *H4 = nullptr, *Winv = nullptr;
  double *G4 = nullptr, *Z4 = nullptr, *S4 = nullptr, *Sbase4 = nullptr, *WinvS = nullptr;
  cip::GemmOperand mapAtil{}, mapZ{};
  cip::CholPlan cholH{}, cholS{};
  int* info = nullptr;  // [2] device
  // cones
  std::vector<int> h_type, h_off;
  int *d_type = nullptr, *d_off = nullptr, *d_rowcone = nullptr, *d_qlist = nullptr, *d_slist = nullptr;
  int *d_sord = nullptr, *d_roff = nullptr;
  std::vector<int> h_slist, h_sord, h_roff;
  size_t r_total = 0;
  cip::ConeDesc cd{};
  cip::Scaling F{}, Fi{};
  bool have_scaling = false, have_factor = false;
  double* gemm_ws = nullptr;   // split-K workspace of gemm_nt (GEMM_WS_DOUBLES)
  int aug_rows = 0;        // rows of Atil4 beyond m_pad holding sqrt(aug_rho) * G
  double aug_rho = 0.0;
  // work vectors
  double* nv[NV] = {};
  double* mv[MV] = {};
  double* pv[PV] = {};
  double* partial = nullptr;
  int partial_cap = 0;
  double* scalar = nullptr;  // device scratch scalars [8]
  // NCCL
  void* comm = nullptr;
  int nranks = 1, rank = 0;
  // stats
  cip_stats_t st{};
  cudaEvent_t ev[8] = {};
  size_t bytes = 0;
  bool need_sync = false;
};


namespace cip {
// sum-all-reduce `count` doubles in place across the row shards (no-op for a single GPU)
int engine_allreduce(cip_engine* h, double* buf, size_t count);
}  // namespace cip

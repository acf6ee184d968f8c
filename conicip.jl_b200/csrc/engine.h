// Private definition of the engine handle, shared by engine.cu (C ABI) and ipm.cu (native
// interior-point driver).
#pragma once
#include <vector>

#include "../../include/conicip_b200.h"
#include "kernels.cuh"

constexpr int NV = 8;   // n-length work vectors
constexpr int MV = 10;  // m-length work vectors
constexpr int PV = 6;   // p-length work vectors

namespace cip { struct Multi; }

struct cip_engine {
  // single-process multi-GPU: a handle created with opts.ngpus > 1 is only a front for `multi` (one shard
  // engine per device, each driven by its own host thread); a shard engine points back through `parent`
  cip::Multi* multi = nullptr;
  cip::Multi* parent = nullptr;
  int shard_index = 0;
  bool always_sync = false;   // shard engines: every call returns with its stream idle (outputs may live on another device)
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
  double* d_sws = nullptr;     // S-cone workspace (orders above 64)
  std::vector<int> h_slist, h_sord, h_roff;
  size_t r_total = 0;
  cip::ConeDesc cd{};
  cip::Scaling F{}, Fi{};
  bool have_scaling = false, have_factor = false;
  double* gemm_ws = nullptr;   // split-K workspace of gemm_nt (GEMM_WS_DOUBLES)
  bool fold = false;           // R-only handle without Atil4: the SYRK reads At4 and scales in registers (opts.fold_scaling)
  cip::GemmOperand mapAt{};
  int aug_rows = 0;        // rows of Atil4 beyond m_pad holding sqrt(aug_rho) * G
  double aug_rho = 0.0;
  // work vectors
  double* nv[NV] = {};
  double* mv[MV] = {};
  double* pv[PV] = {};
  // second set of the work vectors cip_solve uses (nv[0..5], mv[0..5], pv[0..4]): column b of a pair in
  // cip_solve_multi; allocated on the first call
  double* nv2[6] = {};
  double* mv2[6] = {};
  double* pv2[5] = {};
  double* partial = nullptr;
  int partial_cap = 0;
  double* scalar = nullptr;  // device scratch scalars [8]
  // NCCL
  void* comm = nullptr;
  int nranks = 1, rank = 0;
  // overlapped Gram reduction (row-sharded handles): tile-major partial Gram matrix, all-reduced range by range
  // on `cs` while the SYRK of the later ranges still runs (alternating between the handle's stream and `s2`)
  double* Hp = nullptr;
  cudaStream_t cs = nullptr, s2 = nullptr;
  cudaEvent_t evc[16] = {};
  // stats
  cip_stats_t st{};
  cudaEvent_t ev[8] = {};
  size_t bytes = 0;
  bool need_sync = false;
};


namespace cip {
// sum-all-reduce `count` doubles in place across the row shards (no-op for a single GPU)
int engine_allreduce(cip_engine* h, double* buf, size_t count);
// cip_create / cip_create_csc for one device (engine.cu)
int engine_create_single(cip_handle* out, int n, int m, int p, const double* Q, int ldq, const double* A, int lda,
                         const double* G, int ldg, const cip_csc* Qs, const cip_csc* As, const cip_csc* Gs,
                         int ncones, const int* cone_type, const int* cone_dim, const cip_options* opts);
const char* last_error_string();
// buffers / streams of the overlapped Gram reduction, once the communicator exists (engine.cu)
int engine_setup_comm(cip_engine* h);

// ---- single-process multi-GPU front (multi.cu): same contracts as the C ABI entry points, global vectors
int multi_create(cip_handle* out, int n, int m, int p, const double* Q, int ldq, const double* A, int lda,
                 const double* G, int ldg, const cip_csc* Qs, const cip_csc* As, const cip_csc* Gs, int ncones,
                 const int* cone_type, const int* cone_dim, const cip_options* opts);
int multi_destroy(cip_engine* h);
int multi_factor(cip_engine* h, const int* kind, const double* fa, const double* fb, const double* fD,
                 const double* fR, int factor);      // factor = 0: cip_set_scaling
int multi_get_scaling(cip_engine* h, int* kind, double* fa, double* fb, double* fD, double* fR);
int multi_nt_scaling(cip_engine* h, const double* v, const double* s, double* lambda_out, int factor);
int multi_solve(cip_engine* h, const double* ry, const double* rw, const double* rv, double* dy, double* dw,
                double* dv);
int multi_solve_multi(cip_engine* h, int nrhs, const double* ry, int ldy, const double* rw, int ldw, const double* rv,
                      int ldv, double* dy, double* dw, double* dv);
int multi_apply(cip_engine* h, int op, const double* x, double* y);
int multi_maxstep(cip_engine* h, const double* x, const double* d, double d_scale, double* alpha_out);
int multi_prod_div(cip_engine* h, const double* x, const double* y, double* o, int divide);
int multi_mul_A(cip_engine* h, int trans, const double* x, double* y);
int multi_mul_GQ(cip_engine* h, int which, int trans, const double* x, double* y);   // which: 0 = G, 1 = Q
int multi_ipm_solve(cip_engine* h, const double* c, const double* b, const double* d, const cip_ipm_options* opts,
                    double* y, double* w, double* v, cip_ipm_result* result);
int multi_stats(cip_engine* h, cip_stats_t* out);
int multi_simple(cip_engine* h, int what, double* out, int ldo);   // 0 form_H, 1 factor_H, 2 sync, 3 get_H
void multi_barrier(cip::Multi* m);   // rendezvous of the shard threads inside one dispatched call
}  // namespace cip

// Single-process multi-GPU front of the C ABI (cip_options.ngpus > 1; SURVEY 8b "Threading", 8e).
//
// The reference calls its kktsolver from ONE process and one task (src/ConicIP.jl:667,682,688), so the
// drop-in that a Julia user reaches with `kktsolver = kktsolver_b200(ngpus = 8)` must drive all devices
// behind that single call.  A handle created with opts.ngpus = N > 1 owns N shard engines, one per device:
// the rows of A are cut on cone boundaries (W is block-diagonal by cone; R cones may be cut anywhere), each
// shard has its own stream and its NCCL communicator (ncclCommInitAll), and every entry point takes and
// returns GLOBAL vectors, which are sliced (m-vectors) or replicated (n- / p-vectors) here.  Each shard is
// driven by its own host thread, parked between calls, so that the N per-device call sequences -- kernel
// launches, the NCCL all-reduce of the partial Gram matrices, the block-cyclic Cholesky -- are issued
// concurrently exactly as the one-process-per-GPU path issues them; the calling thread blocks until every
// shard has finished, which is the blocking `ccall` semantics the reference protocol expects.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <math.h>

#include "engine.h"
#include "nccl_dl.h"

namespace cip {

struct Multi {
  int N = 0;
  std::vector<int> dev;
  std::vector<cip_engine*> shard;
  std::vector<int> row_lo, row_hi;                // global row range of every shard
  std::vector<std::vector<int>> gcone;            // [shard][local cone] -> global cone
  std::vector<std::vector<int>> ltype, ldim;      // local cone lists
  std::vector<long long> sr_off;                  // [shard] doubles of R blocks (all S cones) before the shard
  int n = 0, m = 0, p = 0, ncones = 0;
  std::vector<int> type, dim, off;                // global cones
  // ---- worker pool
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::function<int(int)> job;
  long long gen = 0;
  int pending = 0;
  bool stop = false;
  std::vector<int> rc;
  std::vector<std::string> err;
  // ---- barrier usable from inside a job
  std::mutex bmu;
  std::condition_variable bcv;
  int bcount = 0;
  long long bgen = 0;

  void worker(int r) {
    cudaSetDevice(dev[r]);
    long long seen = 0;
    for (;;) {
      std::function<int(int)> f;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&] { return stop || gen != seen; });
        if (stop) return;
        seen = gen;
        f = job;
      }
      cudaSetDevice(dev[r]);
      const int code = f(r);
      {
        std::lock_guard<std::mutex> lk(mu);
        rc[r] = code;
        if (code != 0) err[r] = last_error_string();
        if (--pending == 0) cv_done.notify_all();
      }
    }
  }

  // run f(r) on every shard thread; < 0 wins, then shard 0's code, then the largest positive one
  int run(const std::function<int(int)>& f) {
    {
      std::lock_guard<std::mutex> lk(mu);
      job = f;
      pending = N;
      for (auto& c : rc) c = 0;
      ++gen;
    }
    cv_work.notify_all();
    {
      std::unique_lock<std::mutex> lk(mu);
      cv_done.wait(lk, [&] { return pending == 0; });
    }
    int pick = -1;
    for (int r = 0; r < N; ++r)
      if (rc[r] < 0) { pick = r; break; }
    if (pick < 0 && rc[0] > 0) pick = 0;
    if (pick < 0)
      for (int r = 0; r < N; ++r)
        if (rc[r] > 0 && (pick < 0 || rc[r] > rc[pick])) pick = r;
    if (pick < 0) return 0;
    set_error("[device %d] %s", dev[pick], err[pick].c_str());
    return rc[pick];
  }

  void barrier() {
    std::unique_lock<std::mutex> lk(bmu);
    const long long g = bgen;
    if (++bcount == N) {
      bcount = 0;
      ++bgen;
      bcv.notify_all();
    } else {
      bcv.wait(lk, [&] { return bgen != g; });
    }
  }
};

void multi_barrier(Multi* m) {
  if (m && m->N > 1) m->barrier();
}

namespace {

template <typename T>
const T* at(const T* p, long long o) { return p ? p + o : nullptr; }
template <typename T>
T* at(T* p, long long o) { return p ? p + o : nullptr; }

// small arrays that may live on the host or on a device -> host vector
template <typename T>
int to_host(std::vector<T>& out, const T* src, size_t count) {
  out.resize(count);
  if (count == 0 || !src) return 0;
  CIP_CUDA(cudaMemcpy(out.data(), src, count * sizeof(T), cudaMemcpyDefault));
  return 0;
}

// Cut the cone list into N contiguous row slabs of near-equal size: R cones are splittable, Q / S cones are
// atoms that go to the slab whose target they straddle least (the host mirror is dist.shard_cones).
void shard_cones(Multi* M) {
  const int N = M->N;
  M->row_lo.assign(N, 0); M->row_hi.assign(N, 0);
  M->gcone.assign(N, {}); M->ltype.assign(N, {}); M->ldim.assign(N, {});
  const long long m = M->m;
  auto target = [&](int r) { return (long long)llround((double)m * (r + 1) / N); };
  int r = 0;
  long long row = 0;
  auto close = [&]() { M->row_hi[r] = (int)row; ++r; M->row_lo[r] = (int)row; };
  for (int g = 0; g < M->ncones; ++g) {
    int k = M->dim[g];
    const int t = M->type[g];
    while (k > 0) {
      int take;
      if (r == N - 1) {
        take = k;
      } else if (t == CIP_CONE_R) {
        const long long room = target(r) - row;
        take = (int)std::min<long long>(k, room > 0 ? room : 0);
        if (take == 0) { close(); continue; }
      } else {
        const bool has = !M->gcone[r].empty();
        if (row >= target(r) || (row + k - target(r) > target(r) - row && has)) { close(); continue; }
        take = k;
      }
      if (!M->gcone[r].empty() && M->gcone[r].back() == g) {
        M->ldim[r].back() += take;                                 // (cannot happen: a cone is cut at most once per shard)
      } else {
        M->gcone[r].push_back(g); M->ltype[r].push_back(t); M->ldim[r].push_back(take);
      }
      row += take;
      k -= take;
    }
  }
  M->row_hi[r] = (int)row;
  for (int q = r + 1; q < N; ++q) { M->row_lo[q] = (int)row; M->row_hi[q] = (int)row; }
  // R blocks of the S cones that precede every shard (cip_get_scaling's fR layout: all S cones, cone order)
  M->sr_off.assign(N, 0);
  std::vector<long long> before(M->ncones + 1, 0);
  for (int g = 0; g < M->ncones; ++g) {
    long long kk = 0;
    if (M->type[g] == CIP_CONE_S) {
      const long long k = (long long)((sqrt(1.0 + 8.0 * M->dim[g]) - 1.0) / 2.0 + 0.5);
      kk = k * k;
    }
    before[g + 1] = before[g] + kk;
  }
  for (int q = 0; q < N; ++q) M->sr_off[q] = M->gcone[q].empty() ? before[M->ncones] : before[M->gcone[q][0]];
}

// rows [lo, hi) of a CSC matrix (host arrays), re-based to index 0
struct CscSlice {
  std::vector<int64_t> colptr, rowval;
  std::vector<double> nzval;
  cip_csc view{};
};
int slice_csc(const cip_csc* M, int lo, int hi, CscSlice* out) {
  std::vector<int64_t> cp, rv;
  std::vector<double> nz;
  CIP_TRY(to_host(cp, M->colptr, (size_t)M->ncols + 1));
  const long long nnz = cp[M->ncols] - cp[0];
  CIP_TRY(to_host(rv, M->rowval, (size_t)(nnz > 0 ? nnz : 0)));
  CIP_TRY(to_host(nz, M->nzval, (size_t)(nnz > 0 ? nnz : 0)));
  const int base = M->index_base;
  out->colptr.assign(M->ncols + 1, 0);
  out->rowval.clear(); out->nzval.clear();
  for (int j = 0; j < M->ncols; ++j) {
    for (long long e = cp[j] - base; e < cp[j + 1] - base; ++e) {
      const long long i = rv[e] - base;
      if (i >= lo && i < hi) { out->rowval.push_back(i - lo); out->nzval.push_back(nz[e]); }
      else if (i < 0 || i >= M->nrows) { set_error("malformed CSC input: row index outside [0, %d) in column %d", M->nrows, j); return -1; }
    }
    out->colptr[j + 1] = (int64_t)out->rowval.size();
  }
  out->view.nrows = hi - lo; out->view.ncols = M->ncols; out->view.index_base = 0;
  out->view.colptr = out->colptr.data();
  out->view.rowval = out->rowval.empty() ? out->colptr.data() : out->rowval.data();   // never dereferenced when nnz = 0
  out->view.nzval = out->nzval.empty() ? nullptr : out->nzval.data();
  return 0;
}

}  // namespace

}  // namespace cip

extern "C" int cip_shard_plan(int ncones, const int* cone_type, const int* cone_dim, int ngpus, int* row_lo,
                              int* row_hi) {
  using namespace cip;
  if (ncones < 0 || ngpus < 1 || !row_lo || !row_hi || (ncones > 0 && (!cone_type || !cone_dim))) {
    set_error("cip_shard_plan: bad arguments");
    return -1;
  }
  Multi M;
  M.N = ngpus; M.ncones = ncones;
  M.type.assign(cone_type, cone_type + ncones);
  M.dim.assign(cone_dim, cone_dim + ncones);
  long long rows = 0;
  for (int i = 0; i < ncones; ++i) {
    if (cone_dim[i] <= 0) { set_error("cone %d has non-positive dimension", i); return -1; }
    rows += cone_dim[i];
  }
  M.m = (int)rows;
  shard_cones(&M);
  for (int r = 0; r < ngpus; ++r) { row_lo[r] = M.row_lo[r]; row_hi[r] = M.row_hi[r]; }
  return 0;
}

namespace cip {

// ------------------------------------------------------------------------------------------------ LEVEL 1
int multi_create(cip_handle* out, int n, int m, int p, const double* Q, int ldq, const double* A, int lda,
                 const double* G, int ldg, const cip_csc* Qs, const cip_csc* As, const cip_csc* Gs, int ncones,
                 const int* cone_type, const int* cone_dim, const cip_options* opts) {
  *out = nullptr;
  const int N = opts->ngpus;
  int ndev = 0;
  CIP_CUDA(cudaGetDeviceCount(&ndev));
  const int base = opts->device < 0 ? 0 : opts->device;
  if (base + N > ndev) {
    set_error("cip_create: ngpus = %d from device %d, but only %d CUDA devices are visible", N, base, ndev);
    return -1;
  }
  if (N > 1 && !nccl_api()) return -1;
  long long rows = 0;
  for (int i = 0; i < ncones; ++i) rows += cone_dim[i];
  if (rows != m) {
    set_error("cone dimensions sum to %lld but A has %d rows", rows, m);
    return -1;
  }
  cip_engine* h = new cip_engine();
  Multi* M = new Multi();
  h->multi = M;
  h->n = n; h->m = m; h->p = p; h->ncones = ncones;
  M->N = N; M->n = n; M->m = m; M->p = p; M->ncones = ncones;
  M->type.assign(cone_type, cone_type + ncones);
  M->dim.assign(cone_dim, cone_dim + ncones);
  M->off.assign(ncones + 1, 0);
  for (int i = 0; i < ncones; ++i) M->off[i + 1] = M->off[i] + cone_dim[i];
  for (int r = 0; r < N; ++r) M->dev.push_back(base + r);
  shard_cones(M);
  M->shard.assign(N, nullptr);
  M->rc.assign(N, 0);
  M->err.assign(N, "");
  std::vector<CscSlice> slices(As ? N : 0);
  if (As)
    for (int r = 0; r < N; ++r)
      if (slice_csc(As, M->row_lo[r], M->row_hi[r], &slices[r]) != 0) { multi_destroy(h); return -1; }
  const bool chatty = opts->verbose > 0 || getenv("CIP_VERBOSE");
  auto say = [&](const char* what) { if (chatty) { fprintf(stderr, "[conicip_b200] multi_create(ngpus=%d): %s\n", N, what); fflush(stderr); } };
  for (int r = 0; r < N; ++r) M->th.emplace_back(&Multi::worker, M, r);
  say("creating the shard engines");
  int rc = M->run([&](int r) {
    // direct peer copies for slabs that arrive as device pointers on another GPU (ignored where unsupported)
    for (int q = 0; q < N; ++q) {
      if (q == r) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, M->dev[r], M->dev[q]) == cudaSuccess && can) cudaDeviceEnablePeerAccess(M->dev[q], 0);
      cudaGetLastError();
    }
    cip_options o{};
    memcpy(&o, opts, std::min<size_t>(sizeof(o), (size_t)opts->struct_size));
    o.struct_size = sizeof(o);
    o.device = M->dev[r];
    o.ngpus = 0;
    const int m_r = M->row_hi[r] - M->row_lo[r];
    cip_handle e = nullptr;
    const int code = engine_create_single(&e, n, m_r, p, Q, ldq, at(A, M->row_lo[r]), lda, G, ldg, Qs,
                                          As ? &slices[r].view : nullptr, Gs, (int)M->gcone[r].size(),
                                          M->ltype[r].data(), M->ldim[r].data(), &o);
    if (code != 0) return code;
    e->parent = M; e->shard_index = r; e->always_sync = true;
    M->shard[r] = e;
    return 0;
  });
  if (rc != 0) { multi_destroy(h); return rc; }
  say("shard engines ready");
  if (N > 1) {
    const NcclApi* api = nccl_api();
    std::vector<void*> comms(N, nullptr);
    say("ncclCommInitAll");
    const int r = api->CommInitAll(comms.data(), N, M->dev.data());
    say("ncclCommInitAll done");
    if (r != 0) {
      set_error("ncclCommInitAll failed: %s", api->GetErrorString(r));
      multi_destroy(h);
      return -1;
    }
    for (int q = 0; q < N; ++q) { M->shard[q]->comm = comms[q]; M->shard[q]->nranks = N; M->shard[q]->rank = q; }
    rc = M->run([&](int q) { return engine_setup_comm(M->shard[q]); });
    if (rc != 0) { multi_destroy(h); return rc; }
    say("communicators attached");
  }
  h->device = M->dev[0];
  *out = h;
  return 0;
}

int multi_destroy(cip_engine* h) {
  Multi* M = h->multi;
  if (M) {
    if (!M->th.empty()) {
      M->run([&](int r) {
        if (M->shard[r]) cip_destroy(M->shard[r]);
        M->shard[r] = nullptr;
        return 0;
      });
      {
        std::lock_guard<std::mutex> lk(M->mu);
        M->stop = true;
      }
      M->cv_work.notify_all();
      for (auto& t : M->th) t.join();
    }
    delete M;
  }
  delete h;
  return 0;
}

// ------------------------------------------------------------------------------------------------ LEVEL 2
int multi_factor(cip_engine* h, const int* kind, const double* fa, const double* fb, const double* fD,
                 const double* fR, int factor) {
  Multi* M = h->multi;
  if (!kind || !fa) { set_error("cip_factor/cip_set_scaling: kind and fa are required"); return -1; }
  std::vector<int> hk;
  std::vector<double> hD;
  CIP_TRY(to_host(hk, kind, (size_t)M->ncones));
  if (fD) CIP_TRY(to_host(hD, fD, (size_t)M->ncones));
  // fR is concatenated over the VECCONG blocks in cone order
  std::vector<long long> vbefore(M->ncones + 1, 0);
  for (int g = 0; g < M->ncones; ++g) {
    long long kk = 0;
    if (hk[g] == CIP_BLK_VECCONG && M->type[g] == CIP_CONE_S) {
      const long long k = (long long)((sqrt(1.0 + 8.0 * M->dim[g]) - 1.0) / 2.0 + 0.5);
      kk = k * k;
    }
    vbefore[g + 1] = vbefore[g] + kk;
  }
  std::vector<std::vector<int>> lk(M->N);
  std::vector<std::vector<double>> lD(M->N);
  for (int r = 0; r < M->N; ++r)
    for (int g : M->gcone[r]) { lk[r].push_back(hk[g]); lD[r].push_back(fD ? hD[g] : 0.0); }
  // three dispatches, not one cip_factor per shard: a shard that rejects its scaling must not leave the others
  // waiting in the all-reduce of form_H
  CIP_TRY(M->run([&](int r) {
    const long long lo = M->row_lo[r];
    const long long ro = M->gcone[r].empty() ? 0 : vbefore[M->gcone[r][0]];
    return cip_set_scaling(M->shard[r], lk[r].data(), fa + lo, at(fb, lo), fD ? lD[r].data() : nullptr, at(fR, ro));
  }));
  if (!factor) return 0;
  CIP_TRY(M->run([&](int r) { return cip_form_H(M->shard[r]); }));
  return M->run([&](int r) { return cip_factor_H(M->shard[r]); });
}

int multi_get_scaling(cip_engine* h, int* kind, double* fa, double* fb, double* fD, double* fR) {
  Multi* M = h->multi;
  std::vector<std::vector<int>> lk(M->N);
  std::vector<std::vector<double>> lD(M->N);
  for (int r = 0; r < M->N; ++r) { lk[r].assign(M->gcone[r].size() + 1, 0); lD[r].assign(M->gcone[r].size() + 1, 0.0); }
  CIP_TRY(M->run([&](int r) {
    const long long lo = M->row_lo[r];
    return cip_get_scaling(M->shard[r], lk[r].data(), at(fa, lo), at(fb, lo), lD[r].data(), at(fR, M->sr_off[r]));
  }));
  std::vector<int> hk(M->ncones, 0);
  std::vector<double> hD(M->ncones, 0.0);
  for (int r = 0; r < M->N; ++r)
    for (size_t i = 0; i < M->gcone[r].size(); ++i) { hk[M->gcone[r][i]] = lk[r][i]; hD[M->gcone[r][i]] = lD[r][i]; }
  if (kind && M->ncones) CIP_CUDA(cudaMemcpy(kind, hk.data(), sizeof(int) * M->ncones, cudaMemcpyDefault));
  if (fD && M->ncones) CIP_CUDA(cudaMemcpy(fD, hD.data(), sizeof(double) * M->ncones, cudaMemcpyDefault));
  return 0;
}

int multi_nt_scaling(cip_engine* h, const double* v, const double* s, double* lambda_out, int factor) {
  Multi* M = h->multi;
  CIP_TRY(M->run([&](int r) {
    const long long lo = M->row_lo[r];
    if (M->row_hi[r] == lo) { M->shard[r]->have_scaling = true; return 0; }
    int code = cip_nt_scaling(M->shard[r], at(v, lo), at(s, lo), at(lambda_out, lo));
    // a failed S cone comes back as its 1-based LOCAL index: report the global one
    if (code > 0 && code <= (int)M->gcone[r].size()) code = M->gcone[r][code - 1] + 1;
    return code;
  }));
  if (!factor) return 0;
  CIP_TRY(M->run([&](int r) { return cip_form_H(M->shard[r]); }));
  return M->run([&](int r) { return cip_factor_H(M->shard[r]); });
}

// ------------------------------------------------------------------------------------------------ LEVEL 3
int multi_solve(cip_engine* h, const double* ry, const double* rw, const double* rv, double* dy, double* dw,
                double* dv) {
  Multi* M = h->multi;
  return M->run([&](int r) {
    const long long lo = M->row_lo[r];
    return cip_solve(M->shard[r], ry, rw, at(rv, lo), r == 0 ? dy : nullptr, r == 0 ? dw : nullptr, at(dv, lo));
  });
}

int multi_solve_multi(cip_engine* h, int nrhs, const double* ry, int ldy, const double* rw, int ldw, const double* rv,
                      int ldv, double* dy, double* dw, double* dv) {
  Multi* M = h->multi;
  return M->run([&](int r) {
    const long long lo = M->row_lo[r];
    return cip_solve_multi(M->shard[r], nrhs, ry, ldy, rw, ldw, at(rv, lo), ldv, r == 0 ? dy : nullptr, r == 0 ? dw : nullptr,
                           at(dv, lo));
  });
}

int multi_apply(cip_engine* h, int op, const double* x, double* y) {
  Multi* M = h->multi;
  return M->run([&](int r) { return cip_apply(M->shard[r], op, at(x, M->row_lo[r]), at(y, M->row_lo[r])); });
}

int multi_maxstep(cip_engine* h, const double* x, const double* d, double d_scale, double* alpha_out) {
  Multi* M = h->multi;
  std::vector<double> a(M->N, INFINITY);
  CIP_TRY(M->run([&](int r) {
    if (M->row_hi[r] == M->row_lo[r]) return 0;
    return cip_maxstep(M->shard[r], at(x, M->row_lo[r]), at(d, M->row_lo[r]), d_scale, &a[r]);
  }));
  double best = INFINITY;
  for (double v : a) best = fmin(best, v);
  *alpha_out = best;
  return 0;
}

int multi_prod_div(cip_engine* h, const double* x, const double* y, double* o, int divide) {
  Multi* M = h->multi;
  return M->run([&](int r) {
    const long long lo = M->row_lo[r];
    return divide ? cip_cone_div(M->shard[r], at(x, lo), at(y, lo), at(o, lo))
                  : cip_cone_prod(M->shard[r], at(x, lo), at(y, lo), at(o, lo));
  });
}

int multi_mul_A(cip_engine* h, int trans, const double* x, double* y) {
  Multi* M = h->multi;
  return M->run([&](int r) {
    const long long lo = M->row_lo[r];
    if (!trans) return cip_mul_A(M->shard[r], 0, x, at(y, lo));
    return cip_mul_A(M->shard[r], 1, at(x, lo), r == 0 ? y : nullptr);     // A'x is all-reduced inside
  });
}

int multi_mul_GQ(cip_engine* h, int which, int trans, const double* x, double* y) {
  Multi* M = h->multi;   // G and Q are replicated: shard 0 answers
  return M->run([&](int r) {
    if (r != 0) return 0;
    return which == 0 ? cip_mul_G(M->shard[0], trans, x, y) : cip_mul_Q(M->shard[0], x, y);
  });
}

int multi_ipm_solve(cip_engine* h, const double* c, const double* b, const double* d, const cip_ipm_options* opts,
                    double* y, double* w, double* v, cip_ipm_result* result) {
  Multi* M = h->multi;
  std::vector<cip_ipm_result> res(M->N);
  const int rc = M->run([&](int r) {
    const long long lo = M->row_lo[r];
    return cip_ipm_solve(M->shard[r], c, at(b, lo), d, opts, r == 0 ? y : nullptr, r == 0 ? w : nullptr, at(v, lo),
                         &res[r]);
  });
  *result = res[0];
  return rc;
}

int multi_stats(cip_engine* h, cip_stats_t* out) {
  Multi* M = h->multi;
  std::vector<cip_stats_t> st(M->N);
  CIP_TRY(M->run([&](int r) { return cip_stats(M->shard[r], &st[r]); }));
  cip_stats_t s = st[0];
  s.m = M->m;
  for (int r = 1; r < M->N; ++r) {       // the step waits for the slowest shard; work and memory add up
    s.ms_scale = fmax(s.ms_scale, st[r].ms_scale); s.ms_syrk = fmax(s.ms_syrk, st[r].ms_syrk);
    s.ms_allreduce = fmax(s.ms_allreduce, st[r].ms_allreduce); s.ms_chol = fmax(s.ms_chol, st[r].ms_chol);
    s.ms_schur = fmax(s.ms_schur, st[r].ms_schur); s.ms_solve = fmax(s.ms_solve, st[r].ms_solve);
    s.syrk_flops += st[r].syrk_flops;
    s.device_bytes += st[r].device_bytes;
    s.m_pad += st[r].m_pad;
  }
  *out = s;
  return 0;
}

int multi_simple(cip_engine* h, int what, double* out, int ldo) {
  Multi* M = h->multi;
  return M->run([&](int r) {
    switch (what) {
      case 0: return cip_form_H(M->shard[r]);
      case 1: return cip_factor_H(M->shard[r]);
      case 2: return cip_sync(M->shard[r]);
      default: return r == 0 ? cip_get_H(M->shard[0], out, ldo) : 0;
    }
  });
}

}  // namespace cip

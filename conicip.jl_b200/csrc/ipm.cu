// cip_ipm_solve: the whole interior-point loop of ConicIP.jl's `conicIP`
// (/root/reference/src/ConicIP.jl:468-939) driven natively on device-resident vectors
// (SURVEY section 8f, rank 1).  Control flow and scalar arithmetic run on the host thread; every
// vector operation is one of the library's kernels; all inner products of a phase are fused into one
// launch + one (optional NCCL) reduction + one 8-byte-per-value read-back.  Row-sharded handles
// reduce the m-vector quantities across ranks.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#include "engine.h"

using namespace cip;

namespace {

constexpr int MAXP = 20;
struct DotPair { const double* x; const double* y; int n; };
struct DotBatch { DotPair p[MAXP]; };

__global__ void __launch_bounds__(1024) dots_kernel(DotBatch b, double* __restrict__ out) {
  const DotPair pr = b.p[blockIdx.x];
  double s = 0.0;
  for (int i = threadIdx.x; i < pr.n; i += 1024) s = fma(pr.x[i], pr.y[i], s);
  __shared__ double sm[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = sm[threadIdx.x];
    t = warp_sum(t);
    if (threadIdx.x == 0) out[blockIdx.x] = t;
  }
}
__global__ void fill_kernel(double* x, double v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = v;
}

struct V4 { double *y, *w, *v, *s; };

struct Ipm {
  cip_engine* h;
  cudaStream_t st;
  int n, m, p;
  std::vector<double*> pool;
  double* scratch = nullptr;   // device [4*MAXP]
  double* dslot = nullptr;     // device [2*nranks]: step-length exchange across the row shards
  std::vector<double> slot;
  int rc = 0;

  double* vec(size_t len) {
    double* q = nullptr;
    if (cudaMalloc(&q, (len ? len : 1) * sizeof(double)) != cudaSuccess) { rc = -1; set_error("cip_ipm_solve: out of device memory"); return nullptr; }
    cudaMemsetAsync(q, 0, (len ? len : 1) * sizeof(double), st);
    pool.push_back(q);
    return q;
  }
  V4 v4() { return V4{vec(n), vec(p), vec(m), vec(m)}; }
  ~Ipm() { for (double* q : pool) cudaFree(q); }

  int ck(int r) { if (r < 0 && rc == 0) rc = r; return r; }
  void axpby(double* out, double a, const double* x, double b, const double* y, size_t len) { ck(vec_axpby(out, a, x, b, y, len, st)); }
  void copy(double* out, const double* x, size_t len) { ck(vec_copy(out, x, len, st)); }
  void fill(double* x, double v, size_t len) {
    if (!len) return;
    fill_kernel<<<(unsigned)std::min<size_t>((len + 255) / 256, 1184), 256, 0, st>>>(x, v, len);
  }
  // local pairs first (n-/p-vectors, replicated), then shard pairs (m-vectors, summed across ranks)
  std::vector<double> dots(const std::vector<DotPair>& local, const std::vector<DotPair>& shard) {
    DotBatch b;
    const int nl = (int)local.size(), ns = (int)shard.size();
    for (int i = 0; i < nl; ++i) b.p[i] = local[i];
    for (int i = 0; i < ns; ++i) b.p[nl + i] = shard[i];
    std::vector<double> out(nl + ns, 0.0);
    if (nl + ns == 0) return out;
    dots_kernel<<<nl + ns, 1024, 0, st>>>(b, scratch);
    cip::g_launches++;
    if (ns) ck(engine_allreduce(h, scratch + nl, ns));
    if (cudaMemcpyAsync(out.data(), scratch, sizeof(double) * (nl + ns), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { ck(-1); set_error("cip_ipm_solve: reduction read-back failed"); }
    return out;
  }
  // min over ranks of two locally computed step lengths
  void maxstep2(const double* x1, const double* d1, const double* x2, const double* d2, double scale, double& a1, double& a2) {
    ck(cip_maxstep(h, x1, d1, scale, &a1));
    ck(cip_maxstep(h, x2, d2, scale, &a2));
    if (h->comm) {
      // global min through the engine's sum all-reduce: every rank writes its two values into its own
      // slots of a zero vector (an all-gather), +Inf included (Inf + 0 = Inf stays in its slot)
      // (dslot is allocated once per solve: a cudaMalloc / cudaFree pair here would be an implicit device
      //  synchronisation three times per iteration)
      slot.assign(2 * h->nranks, 0.0);
      slot[2 * h->rank] = a1; slot[2 * h->rank + 1] = a2;
      cudaMemcpyAsync(dslot, slot.data(), sizeof(double) * slot.size(), cudaMemcpyHostToDevice, st);
      ck(engine_allreduce(h, dslot, slot.size()));
      cudaMemcpyAsync(slot.data(), dslot, sizeof(double) * slot.size(), cudaMemcpyDeviceToHost, st);
      cudaStreamSynchronize(st);
      for (int r = 0; r < h->nranks; ++r) { a1 = std::min(a1, slot[2 * r]); a2 = std::min(a2, slot[2 * r + 1]); }
    }
  }
  double gsum(double x) {     // scalar sum over ranks (setup only)
    if (!h->comm) return x;
    cudaMemcpyAsync(scratch, &x, 8, cudaMemcpyHostToDevice, st);
    ck(engine_allreduce(h, scratch, 1));
    cudaMemcpyAsync(&x, scratch, 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    return x;
  }
};

double nanv() { return std::numeric_limits<double>::quiet_NaN(); }

}  // namespace

extern "C" int cip_ipm_solve(cip_handle h, const double* c_in, const double* b_in, const double* d_in,
                             const cip_ipm_options* opt_in, double* y_out, double* w_out, double* v_out,
                             cip_ipm_result* res) {
  if (!h || !res || !c_in || (h->m && !b_in) || (h->p && !d_in)) { set_error("cip_ipm_solve: null argument"); return -1; }
  if (h->multi) return multi_ipm_solve(h, c_in, b_in, d_in, opt_in, y_out, w_out, v_out, res);
  CIP_CUDA(cudaSetDevice(h->device));
  const auto t_start = std::chrono::steady_clock::now();
  cip_ipm_options o{};
  o.optTol = 1e-6; o.DTB = 0.01; o.maxRefinementSteps = 3; o.maxIters = 100; o.infeasTol = -1; o.refinementThreshold = -1;
  if (opt_in) memcpy(&o, opt_in, std::min<size_t>(sizeof(o), (size_t)opt_in->struct_size));
  if (o.infeasTol < 0) o.infeasTol = o.optTol;                         // src/ConicIP.jl:506
  if (o.refinementThreshold < 0) o.refinementThreshold = o.optTol / 1e7;   // :509
  memset(res, 0, sizeof(*res));
  res->prFeas = res->duFeas = res->muFeas = res->pobj = INFINITY; res->dobj = -INFINITY;

  Ipm I;
  I.h = h; I.st = h->stream; I.n = h->n; I.m = h->m; I.p = h->p;
  const int n = I.n, m = I.m, p = I.p;
  CIP_CUDA(cudaMalloc(&I.scratch, sizeof(double) * 4 * MAXP));
  I.pool.push_back(I.scratch);
  if (h->comm) I.dslot = I.vec(2 * h->nranks);

  // ---- problem vectors, e and conedim (src/ConicIP.jl:547-565)
  double *c = I.vec(n), *b = I.vec(m), *d = I.vec(p), *e = I.vec(m), *ones = I.vec(m);
  CIP_CUDA(cudaMemcpyAsync(c, c_in, sizeof(double) * n, cudaMemcpyDefault, I.st));
  if (m) CIP_CUDA(cudaMemcpyAsync(b, b_in, sizeof(double) * m, cudaMemcpyDefault, I.st));
  if (p) CIP_CUDA(cudaMemcpyAsync(d, d_in, sizeof(double) * p, cudaMemcpyDefault, I.st));
  std::vector<double> eh(m, 0.0);
  double conedim = 0;
  for (int i = 0; i < h->ncones; ++i) {
    const int off = h->h_off[i], k = h->h_off[i + 1] - off;
    if (h->h_type[i] == CIP_CONE_R) { conedim += k; for (int r = 0; r < k; ++r) eh[off + r] = 1.0; }
    else if (h->h_type[i] == CIP_CONE_Q) { conedim += 1; eh[off] = 1.0; }
    else {
      const int ks = (int)((std::sqrt(1.0 + 8.0 * k) - 1.0) / 2.0 + 0.5);
      conedim += ks;
      int idx = 0;
      for (int r = 0; r < ks; ++r) { eh[off + idx] = 1.0; idx += ks - r; }
    }
  }
  if (m) CIP_CUDA(cudaMemcpyAsync(e, eh.data(), sizeof(double) * m, cudaMemcpyHostToDevice, I.st));
  I.fill(ones, 1.0, m);
  conedim = I.gsum(conedim);
  const double m_glob = I.gsum((double)m);
  const bool any_s = I.gsum((double)h->cd.ns) > 0;     // some shard holds S cones: their NT scaling can fail

  V4 z = I.v4(), rl = I.v4(), r0 = I.v4(), r = I.v4(), daff = I.v4(), dz = I.v4(), dzr = I.v4(), rI = I.v4();
  double *lam = I.vec(m), *t1 = I.vec(m), *tm1 = I.vec(m), *tm2 = I.vec(m), *lc = I.vec(m);
  double *Qy = I.vec(n), *Gtw = I.vec(n), *Atv = I.vec(n), *tn1 = I.vec(n), *Ay = I.vec(m), *tp1 = I.vec(p);
  multi_barrier(h->parent);        // shard threads of one process: all allocations done before the first collective
  if (I.rc) return I.rc;

  const auto n0 = I.dots({{c, c, n}, {d, d, p}}, {{b, b, m}});
  const double normc = std::sqrt(n0[0]);
  const double normd = p ? std::sqrt(n0[1]) : -INFINITY;
  const double normb = std::sqrt(n0[2]);

  int factors = 0, solves = 0;
  auto solve4x4 = [&](const V4& rr, V4& out) {           // src/ConicIP.jl:684-692
    ++solves;
    I.ck(cip_cone_div(h, rr.s, lam, tm1));
    I.ck(cip_apply(h, CIP_OP_FT, tm1, t1));
    I.axpby(tm2, 1.0, rr.v, 1.0, t1, m);
    I.ck(cip_solve(h, rr.y, rr.w, tm2, out.y, p ? out.w : nullptr, out.v));
    I.ck(cip_apply(h, CIP_OP_F, out.v, tm1));
    I.ck(cip_apply(h, CIP_OP_FT, tm1, tm2));
    I.axpby(out.s, 1.0, t1, -1.0, tm2, m);
  };

  // ---- initial point (:704-713): F = I for every cone, lambda = e
  {
    std::vector<int> kind(h->ncones, CIP_BLK_DIAG);
    int st = cip_factor(h, kind.data(), ones, nullptr, nullptr, nullptr);
    ++factors;
    if (st < 0) return st;
    if (st > 0) { res->status = CIP_STATUS_ERROR; return 0; }
    I.copy(lam, e, m);
    V4 rr{c, d, b, tm1};               // r0 = (c, d, b, 0)
    I.fill(lc, 0.0, m);
    rr.s = lc;
    solve4x4(rr, z);
    double av, as;
    I.maxstep2(z.v, nullptr, z.s, nullptr, 1.0, av, as);
    I.axpby(z.v, 1.0, z.v, -av, e, m);
    I.axpby(z.s, 1.0, z.s, -as, e, m);
  }

  double optBest = INFINITY;
  int status = CIP_STATUS_NONE;
  double yscale = 1.0, vwscale = 1.0;

  for (int Iter = 1; Iter <= o.maxIters && I.rc == 0; ++Iter) {
    const int nst = cip_nt_scaling(h, z.v, z.s, lam);                 // :732-735
    if (nst < 0) { I.ck(nst); break; }
    if ((any_s ? I.gsum(nst > 0 ? 1.0 : 0.0) : (double)nst) > 0) { status = CIP_STATUS_ERROR; break; }                 // S-cone iterate not PD (PosDefException, :201)
    I.ck(cip_form_H(h));
    int fst = cip_factor_H(h);                                         // :737 -> :682
    ++factors;
    if (fst < 0) { I.ck(fst); break; }
    if (fst > 0) { status = CIP_STATUS_ERROR; break; }

    // residuals (:746-753)
    I.ck(cip_mul_Q(h, z.y, Qy));
    I.ck(cip_mul_G(h, 1, z.w, Gtw));
    I.ck(cip_mul_A(h, 1, z.v, Atv));
    I.ck(cip_mul_A(h, 0, z.y, Ay));
    if (p) I.ck(cip_mul_G(h, 0, z.y, rl.w));
    I.axpby(tn1, 1.0, Gtw, -1.0, Atv, n);                              // G'w - A'v
    I.axpby(rl.y, 1.0, Qy, 1.0, tn1, n);
    I.axpby(rl.v, 1.0, Ay, -1.0, z.s, m);                              // also A y - s for d_infeas1
    I.ck(cip_cone_prod(h, lam, lam, rl.s));
    I.axpby(r0.y, 1.0, rl.y, -1.0, c, n);
    I.axpby(r0.w, 1.0, rl.w, -1.0, d, p);
    I.axpby(r0.v, 1.0, rl.v, -1.0, b, m);
    I.copy(r0.s, rl.s, m);

    const auto s1 = I.dots({{c, z.y, n}, {r0.y, r0.y, n}, {z.y, Qy, n}, {z.w, r0.w, p}, {d, z.w, p}, {tn1, tn1, n},
                            {z.y, z.y, n}, {rl.w, rl.w, p}, {Qy, Qy, n}},
                           {{z.v, z.s, m}, {r0.v, r0.v, m}, {r0.s, r0.s, m}, {z.v, r0.v, m}, {b, z.v, m}, {z.v, z.v, m},
                            {rl.v, rl.v, m}});
    const double cty = s1[0], r0y2 = s1[1], yQy = s1[2], w_r0w = s1[3], dtw = s1[4], gta2 = s1[5], yy = s1[6],
                 gy2 = s1[7], qy2 = s1[8];
    const double mubar = s1[9], r0v2 = s1[10], r0s2 = s1[11], v_r0v = s1[12], btv = s1[13], vv = s1[14], ays2 = s1[15];
    const double mu = mubar / conedim;
    const double rDu = std::sqrt(r0y2) / (1 + normc), rPr = std::sqrt(r0v2) / (1 + normb),
                 rCp = std::sqrt(r0s2) / (1 + std::fabs(cty));
    const double worst = std::fmax(rDu, std::fmax(rPr, rCp));
    if (worst < optBest) {                                             // :768-773
      res->Iter = Iter; res->Mu = mu; res->duFeas = rDu; res->prFeas = rPr; res->muFeas = rCp;
      optBest = worst;
    }
    res->pobj = 0.5 * yQy - cty;
    res->dobj = res->pobj + w_r0w + v_r0v - mubar;
    if (worst < o.optTol) status = CIP_STATUS_OPTIMAL;                  // :786
    if (!(p == 0 && m_glob == 0)) {                                    // :790-852
      const double dty_btv = dtw - btv;
      double p_infeas = nanv();
      if (dty_btv < 0) {
        const double pu = std::sqrt(gta2);
        const double p_cvx = pu / (std::sqrt(yy) + std::sqrt(vv));
        const double p_ecos = pu / (std::fmax(1.0, normc) * std::fabs(dty_btv));
        p_infeas = (std::isnan(p_cvx) || std::isnan(p_ecos)) ? nanv() : std::fmax(p_cvx, p_ecos);
      }
      if (p_infeas < o.infeasTol) { status = CIP_STATUS_INFEASIBLE; vwscale = 1.0 / -dty_btv; }
      const double d1 = m_glob == 0 ? -INFINITY : std::sqrt(ays2);
      const double d2 = p == 0 ? -INFINITY : std::sqrt(gy2);
      const double d3 = std::isfinite(yy) ? std::sqrt(qy2) : nanv();
      double d_infeas = nanv();
      if (cty > 0) {
        const double a1 = std::fmax(d1 / std::fmax(1.0, normb), std::fmax(d2 / std::fmax(1.0, normd), d3 / std::fmax(1.0, normc))) / std::fabs(cty);
        const double a2 = std::fmax(d1, std::fmax(d2, d3)) / std::sqrt(yy);
        d_infeas = (std::isnan(a1) || std::isnan(a2) || std::isnan(d3)) ? nanv() : std::fabs(std::fmax(a1, a2));
      }
      if (d_infeas < o.infeasTol) { status = CIP_STATUS_UNBOUNDED; yscale = 1.0 / std::fabs(cty); }
    }
    if (o.verbose)
      printf(" %6d  | %8.1e %8.1e %8.1e | % 8.1e % 8.1e | mu %8.1e\n", Iter, rDu, rPr, rCp, res->pobj, res->dobj, mu);
    if (status != CIP_STATUS_NONE) break;
    if (!(std::isfinite(mu) && std::isfinite(rDu) && std::isfinite(rPr) && std::isfinite(rCp))) {   // :870-873
      status = CIP_STATUS_ERROR;
      break;
    }

    // predictor (:879-887)
    solve4x4(r0, daff);
    double a1, a2;
    I.maxstep2(z.v, daff.v, z.s, daff.s, 1.0, a1, a2);
    const double a_aff = std::fmin(std::fmin(a1, 1.0), std::fmin(a2, 1.0));
    const auto s2 = I.dots({}, {{z.v, daff.s, m}, {daff.v, z.s, m}, {daff.v, daff.s, m}});
    const double rho = (mubar - a_aff * s2[0] - a_aff * s2[1] + a_aff * a_aff * s2[2]) / mubar;
    const double sigma = std::pow(std::fmax(0.0, std::fmin(1.0, rho)), 3);

    // corrector (:893-901):  lc = -(F^-T ds o F dv) + sigma mu e ;  r.s = rleft.s - lc
    I.ck(cip_apply(h, CIP_OP_FINVT, daff.s, tm1));
    I.ck(cip_apply(h, CIP_OP_F, daff.v, tm2));
    I.ck(cip_cone_prod(h, tm1, tm2, lc));
    I.axpby(lc, -1.0, lc, sigma * mu, e, m);
    r.y = r0.y; r.w = r0.w; r.v = r0.v;
    I.axpby(r.s, 1.0, rl.s, -1.0, lc, m);

    // Newton step with iterative refinement (:907-921)
    solve4x4(r, dz);
    for (int rs = 0; rs < o.maxRefinementSteps; ++rs) {
      I.ck(cip_apply(h, CIP_OP_F, dz.v, tm1));
      I.ck(cip_cone_prod(h, lam, tm1, tm2));                           // lam o F dv
      I.ck(cip_apply(h, CIP_OP_FINVT, dz.s, tm1));
      I.ck(cip_cone_prod(h, lam, tm1, lc));                            // lam o F^-T ds
      I.axpby(tm2, 1.0, tm2, 1.0, lc, m);
      I.axpby(rI.s, 1.0, r.s, -1.0, tm2, m);
      I.ck(cip_mul_Q(h, dz.y, Qy));
      I.ck(cip_mul_G(h, 1, dz.w, Gtw));
      I.ck(cip_mul_A(h, 1, dz.v, Atv));
      I.axpby(tn1, 1.0, Qy, 1.0, Gtw, n);
      I.axpby(tn1, 1.0, tn1, -1.0, Atv, n);
      I.axpby(rI.y, 1.0, r.y, -1.0, tn1, n);
      if (p) { I.ck(cip_mul_G(h, 0, dz.y, tp1)); I.axpby(rI.w, 1.0, r.w, -1.0, tp1, p); }
      I.ck(cip_mul_A(h, 0, dz.y, Ay));
      I.axpby(tm1, 1.0, Ay, -1.0, dz.s, m);
      I.axpby(rI.v, 1.0, r.v, -1.0, tm1, m);
      const auto s3 = I.dots({{rI.y, rI.y, n}, {rI.w, rI.w, p}}, {{rI.v, rI.v, m}, {rI.s, rI.s, m}});
      const double rnorm = (std::sqrt(s3[0]) + std::sqrt(s3[1]) + std::sqrt(s3[2]) + std::sqrt(s3[3])) / (n + 2 * m_glob);
      if (rnorm < o.refinementThreshold) break;
      solve4x4(rI, dzr);
      I.axpby(dz.y, 1.0, dz.y, 1.0, dzr.y, n);
      I.axpby(dz.w, 1.0, dz.w, 1.0, dzr.w, p);
      I.axpby(dz.v, 1.0, dz.v, 1.0, dzr.v, m);
      I.axpby(dz.s, 1.0, dz.s, 1.0, dzr.s, m);
    }

    // step (:927-932)
    double av, as;
    I.maxstep2(z.v, dz.v, z.s, dz.s, 1.0 - o.DTB, av, as);
    const double alpha = std::fmin(std::fmin(av, 1.0), std::fmin(as, 1.0));
    I.axpby(z.y, 1.0, z.y, -alpha, dz.y, n);
    I.axpby(z.w, 1.0, z.w, -alpha, dz.w, p);
    I.axpby(z.v, 1.0, z.v, -alpha, dz.v, m);
    I.axpby(z.s, 1.0, z.s, -alpha, dz.s, m);
    if (Iter == o.maxIters) status = CIP_STATUS_ABANDONED;             // :936
  }
  if (I.rc) return I.rc;
  if (status == CIP_STATUS_NONE) status = CIP_STATUS_ABANDONED;

  // returned vectors: current iterate, or the certificates of :816 / :848
  if (status == CIP_STATUS_INFEASIBLE) {
    I.fill(z.y, nanv(), n);
    I.axpby(z.w, vwscale, z.w, 0.0, nullptr, p);
    I.axpby(z.v, vwscale, z.v, 0.0, nullptr, m);
  } else if (status == CIP_STATUS_UNBOUNDED) {
    I.axpby(z.y, yscale, z.y, 0.0, nullptr, n);
    I.fill(z.w, nanv(), p);
    I.fill(z.v, nanv(), m);
  }
  if (y_out) CIP_CUDA(cudaMemcpyAsync(y_out, z.y, sizeof(double) * n, cudaMemcpyDefault, I.st));
  if (w_out && p) CIP_CUDA(cudaMemcpyAsync(w_out, z.w, sizeof(double) * p, cudaMemcpyDefault, I.st));
  if (v_out && m) CIP_CUDA(cudaMemcpyAsync(v_out, z.v, sizeof(double) * m, cudaMemcpyDefault, I.st));
  CIP_CUDA(cudaStreamSynchronize(I.st));
  res->status = status;
  res->factors = factors;
  res->solves = solves;
  res->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  return 0;
}

"""Developer sanity run on a B200 (not part of the test suite): checks each kernel family
against the oracle at small sizes and prints timings at C2 size."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import conicip_b200 as cb
import oracle as O
from conicip_b200 import problems as P

def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

print("torch", torch.__version__, torch.cuda.get_device_name(0), flush=True)
print("peaks", cb.measure_fp64_peaks(), flush=True)

def oracle_F(prob, v, s):
    blocks = []
    off = 0
    for t, k in prob["cone_dims"]:
        if t == "R":
            blocks.append(O.Diag(np.sqrt(s[off:off+k] / v[off:off+k])))
        else:
            blocks.append(O.nestod_soc(v[off:off+k], s[off:off+k]))
        off += k
    return O.Block(blocks)

def interior_point(prob, rng):
    v = np.zeros(len(prob["b"])); s = np.zeros_like(v); off = 0
    for t, k in prob["cone_dims"]:
        if t == "R":
            v[off:off+k] = rng.uniform(0.5, 2, k); s[off:off+k] = rng.uniform(0.5, 2, k)
        else:
            for x in (v, s):
                u = rng.standard_normal(k - 1); x[off] = np.linalg.norm(u) + rng.uniform(0.1, 1); x[off+1:off+k] = u
        off += k
    return v, s

for prob in (P.mixed(), P.mixed(n=200, mr=333, ncones=5, k=33, p=0, seed=3), P.mixed(n=300, mr=1000, ncones=0, k=3, p=130, seed=4)):
    rng = np.random.default_rng(0)
    Q, A, G, cd = prob["Q"], prob["A"], prob["G"], prob["cone_dims"]
    n, m, p = len(prob["c"]), A.shape[0], G.shape[0]
    eng = cb.Engine(Q, A, G if p else None, cd)
    v, s = interior_point(prob, rng)
    lam = eng.nt_scaling(v, s)
    Fo = oracle_F(prob, v, s)
    print(prob["name"], n, m, p, "lambda", rel(lam, Fo.mul(v)), flush=True)
    x = rng.standard_normal(m)
    print("  apply F", rel(eng.apply(cb.OP_F, x), Fo.mul(x)), "FinvT", rel(eng.apply(cb.OP_FINVT, x), Fo.inv_adjoint().mul(x)))
    # maxstep / prod / div
    d = rng.standard_normal(m)
    def o_maxstep(x, d):
        mn = np.inf; off = 0
        for t, k in cd:
            xi = x[off:off+k]; di = None if d is None else d[off:off+k]
            mn = min(mn, O.maxstep_rp(xi, di) if t == "R" else O.maxstep_soc(xi, di)); off += k
        return mn
    print("  maxstep", eng.maxstep(v, d), o_maxstep(v, d), eng.maxstep(d, None), o_maxstep(d, None), eng.maxstep(v, None))
    def o_pd(fn_r, fn_q, x, y):
        o = np.zeros(m); off = 0
        for t, k in cd:
            o[off:off+k] = (fn_r if t == "R" else fn_q)(x[off:off+k], y[off:off+k]); off += k
        return o
    print("  prod", rel(eng.cone_prod(x, d), o_pd(O.xrp, O.xsoc, x, d)), "div", rel(eng.cone_div(x, v), o_pd(O.drp, O.dsoc, x, v)))
    print("  mulA", rel(eng.mul_A(rng.standard_normal(n)*0+1), A @ np.ones(n)), "mulAt", rel(eng.mul_A(x, trans=True), A.T @ x),
          "mulQ", rel(eng.mul_Q(np.arange(n, dtype=float)), Q @ np.arange(n)))
    if p:
        w = rng.standard_normal(p)
        print("  mulG", rel(eng.mul_G(np.ones(n)), G @ np.ones(n)), "mulGt", rel(eng.mul_G(w, trans=True), G.T @ w))
    eng.form_H()
    Fi = Fo.inv_adjoint()
    At = Fi.mul(A)
    Href = Q + At.T @ At
    Hg = eng.get_H()
    print("  H lower rel", rel(np.tril(Hg), np.tril(Href)), flush=True)
    st = eng.factor_H()
    Lg = np.tril(eng.get_H())
    Lref = np.linalg.cholesky(Href)
    print("  chol status", st, "L rel", rel(Lg, Lref))
    ry, rw, rv = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m)
    dy, dw, dv = eng.solve(ry, rw, rv)
    solve_o = O.kktsolver_chol(Q, A, G, cd)(Fo, Fi)
    oy, ow, ov = solve_o(ry, rw, rv)
    print("  solve rel", rel(dy, oy), rel(dw, ow) if p else 0, rel(dv, ov))
    # residual of the 3x3 system
    FtF = Fo.dense().T @ Fo.dense()
    r1 = Q @ dy + (G.T @ dw if p else 0) - A.T @ dv - ry
    r3 = A @ dy + FtF @ dv - rv
    print("  kkt resid", np.linalg.norm(r1), np.linalg.norm(G @ dy - rw) if p else 0, np.linalg.norm(r3), flush=True)
    eng.close()

# full driver on small problems
for prob in (P.sphere(), P.combined(), P.simplex(), P.mixed()):
    kw = dict(optTol=prob.get("optTol", 1e-7))
    t0 = time.time()
    s = cb.conicIP(prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"], prob["G"] if prob["G"].shape[0] else None,
                   prob["d"] if prob["G"].shape[0] else None, **kw)
    so = O.conicIP(prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"], prob["G"], prob["d"], kktsolver=O.kktsolver_chol, **kw)
    print(prob["name"], s.status, s.Iter, so.Iter, "Mu", s.Mu, so.Mu, "y rel", rel(s.y, so.y), "v rel", rel(s.v, so.v), f"{time.time()-t0:.2f}s", flush=True)

# timing at C2 size
n, m = 8192, 16384
g = torch.Generator(device="cuda"); g.manual_seed(0)
At = torch.randn((n, m), generator=g, dtype=torch.float64, device="cuda") / n**0.5
Qd = np.ones(n)
import scipy.sparse as sp
eng = cb.Engine(sp.diags(Qd).tocsr(), At.t(), None, [("R", m)])
v = torch.rand(m, dtype=torch.float64, device="cuda") + 0.5
s_ = torch.rand(m, dtype=torch.float64, device="cuda") + 0.5
for it in range(3):
    lam = eng.factor_from_point(v, s_)
    st = eng.stats()
    print({k: round(st[k], 3) for k in ("ms_scale", "ms_syrk", "ms_chol", "ms_schur")},
          "syrk TF", st["syrk_flops"] / st["ms_syrk"] / 1e9, "chol TF", st["chol_flops"] / st["ms_chol"] / 1e9, flush=True)
ry = torch.randn(n, dtype=torch.float64, device="cuda"); rv = torch.randn(m, dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t0 = time.time()
for _ in range(5):
    dy, dw, dv = eng.solve(ry, None, rv)
torch.cuda.synchronize(); print("solve ms", (time.time() - t0) / 5 * 1e3, eng.stats()["ms_solve"])
# correctness at size: residual of reduced system
f = torch.sqrt(s_ / v)
t1 = rv / (f * f)
Hy = eng.mul_Q(dy) + eng.mul_A(eng.mul_A(dy) / (f * f), trans=True)
rhs = ry + eng.mul_A(t1, trans=True)
print("C2 reduced-system rel resid", float(torch.linalg.vector_norm(Hy - rhs) / torch.linalg.vector_norm(rhs)))
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
for _ in range(2): c = a @ b
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); 
for _ in range(3): c = a @ b
e1.record(); torch.cuda.synchronize(); print("cuBLAS DGEMM 8192^3 TF", 3 * 2 * 8192**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)

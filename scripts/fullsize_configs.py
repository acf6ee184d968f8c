"""Full-size runs of the BASELINE.json parity configurations C1, C2, C3 and a C5-like LP with an
S(64) block, for the record under profiles/: time-to-1e-8, iteration counts, residuals,
per-phase timings, and parity against the oracle where the oracle finishes in reasonable time."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import conicip_b200 as cb
import oracle as O
from conicip_b200 import problems as P


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run(prob, with_oracle, oracle_solver=None):
    p = prob["G"].shape[0]
    out = {"name": prob["name"], "n": len(prob["c"]), "m": prob["A"].shape[0], "p": p,
           "cones": f"{len(prob['cone_dims'])} blocks"}
    holder = {}

    def kk(Q, A, G, cd):
        gen = cb.kktsolver_b200(Q, A, G, cd)
        holder["eng"] = gen.engine
        return gen

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s = cb.conicIP(prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"], prob["G"] if p else None,
                   prob["d"] if p else None, optTol=1e-8, kktsolver=kk)
    torch.cuda.synchronize()
    out["b200"] = dict(time_to_1e8_s=time.perf_counter() - t0, status=s.status, Iter=s.Iter, factors=s.factors,
                       solves=s.solves, prFeas=s.prFeas, duFeas=s.duFeas, muFeas=s.muFeas)
    st = holder["eng"].stats()
    out["last_factor_ms"] = {k: st[k] for k in ("ms_scale", "ms_syrk", "ms_chol", "ms_schur", "ms_solve")}
    if with_oracle:
        t0 = time.perf_counter()
        so = O.conicIP(prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"], prob["G"], prob["d"],
                       optTol=1e-8, kktsolver=oracle_solver or O.kktsolver_chol)
        out["oracle"] = dict(time_s=time.perf_counter() - t0, status=so.status, Iter=so.Iter)
        out["parity"] = dict(iter_diff=s.Iter - so.Iter, rel_y=rel(s.y, so.y), rel_v=rel(s.v, so.v),
                             rel_w=rel(s.w, so.w) if p else 0.0)
    print(json.dumps(out), flush=True)
    return out


def config5(n=20000, k=64, p=1000, seed=5):
    """C5-like: LP (Q = 0), x >= 0 on all n variables, one PSD block of order k on the first
    k(k+1)/2 variables, p sparse equality rows (~10 nnz/row); strictly feasible primal and dual."""
    rng = np.random.default_rng(seed)
    dim = k * (k + 1) // 2
    m = n + dim
    A = np.zeros((m, n))
    A[np.arange(n), np.arange(n)] = 1.0
    A[n + np.arange(dim), np.arange(dim)] = 1.0
    B = rng.standard_normal((k, k)) / np.sqrt(k)
    X0 = B @ B.T + 0.5 * np.eye(k)
    y0 = rng.uniform(0.5, 1.5, n)
    y0[:dim] = np.abs(O.vecm(X0)) + 0.05            # >0 entrywise and keep mat() PD via diagonal dominance
    Xs = O.mat(y0[:dim])
    Xs += (0.1 - min(0.0, np.linalg.eigvalsh(Xs).min())) * np.eye(k)
    y0[:dim] = O.vecm(Xs)
    assert np.linalg.eigvalsh(O.mat(y0[:dim])).min() > 0 and y0.min() > 0
    b = np.zeros(m)
    G = np.zeros((p, n))
    for i in range(p):
        G[i, rng.choice(n, 10, replace=False)] = rng.standard_normal(10)
    d = G @ y0
    v0 = np.zeros(m)
    v0[:n] = rng.uniform(0.5, 1.5, n)
    Bz = rng.standard_normal((k, k)) / np.sqrt(k)
    v0[n:] = O.vecm(Bz @ Bz.T + 0.5 * np.eye(k))
    w0 = rng.standard_normal(p)
    c = G.T @ w0 - A.T @ v0                          # stationarity (src/ConicIP.jl:747,753): Qy + G'w - A'v = c
    return dict(name="C5-like", Q=np.zeros((n, n)), c=c, A=A, b=b, cone_dims=[("R", n), ("S", dim)], G=G, d=d)


if __name__ == "__main__":
    which = sys.argv[1:] or ["C1", "C2", "C3", "C5"]
    res = []
    if "C1" in which:
        res.append(run(P.config1(), True, O.pivot(O.kktsolver_2x2)))
    if "C3" in which:
        res.append(run(P.config3(), True))
    if "C2" in which:
        res.append(run(P.config2(), "C2o" in which))
    if "C5" in which:
        res.append(run(config5(), False))
    if "C5s" in which:
        res.append(run(config5(n=600, k=12, p=40), True, O.kktsolver_qr))
    json.dump(res, open("gpurun_out/fullsize_configs.json", "w"), indent=1)

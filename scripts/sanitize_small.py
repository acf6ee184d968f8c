"""Small run touching every kernel family, for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import conicip_b200 as cb
import oracle as O
from conicip_b200 import problems as P

rng = np.random.default_rng(0)
# R + Q cones + equality block (two Cholesky panels for H: n = 200 -> n_pad = 256)
prob = P.mixed(n=200, mr=150, ncones=4, k=9, p=6, seed=1)
s = cb.conicIP_native(prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"], prob["G"], prob["d"], optTol=1e-8)
print("mixed", s.status, s.Iter)
# S cone (order 6) + R + Q in one problem, Python driver (exercises cip_apply / maxstep / prod / div separately)
k = 6
dim = k * (k + 1) // 2
cones = [("R", 5), ("S", dim), ("Q", 4)]
m, n = 5 + dim + 4, 12
A = rng.standard_normal((m, n))
y0 = rng.standard_normal(n)
s0 = np.zeros(m)
s0[:5] = 1.0
B = rng.standard_normal((k, k))
s0[5:5 + dim] = O.vecm(B @ B.T + np.eye(k))
s0[5 + dim] = 2.0
b = A @ y0 - s0
sol = cb.conicIP(np.eye(n), rng.standard_normal(n), A, b, cones, optTol=1e-8)
print("r+s+q", sol.status, sol.Iter)
# sparse ingestion
import scipy.sparse as sp
e = cb.Engine(sp.identity(40, format="csc"), sp.random(60, 40, density=0.2, random_state=1, format="csc") + sp.vstack([sp.identity(40), sp.csc_matrix((20, 40))]), None, [("R", 60)])
e.factor_from_point(np.ones(60), np.ones(60) * 2)
print("csc", e.solve(np.ones(40), None, np.ones(60))[0][:2])
e.close()
# ---- round 2 additions
# three Cholesky panels (n = 300 -> n_pad = 384): chol_head_kernel on the panel chain, sweeps with several block rows,
# heterogeneous Q cones (32 lanes per cone class), cip_solve_multi with an odd count
cones = [("R", 7), ("Q", 2), ("Q", 65), ("Q", 3), ("R", 5), ("Q", 100)]
m, n = sum(k for _, k in cones), 300
A = rng.standard_normal((m, n)) / np.sqrt(n)
e = cb.Engine(np.eye(n) * 2.0, A, rng.standard_normal((3, n)), cones)
v, sv = np.zeros(m), np.zeros(m)
off = 0
for t, k in cones:
    for x in (v, sv):
        if t == "R":
            x[off:off + k] = rng.uniform(0.5, 2, k)
        else:
            u = rng.standard_normal(k - 1); x[off] = np.linalg.norm(u) + 0.5; x[off + 1:off + k] = u
    off += k
e.factor_from_point(v, sv)
DY, DW, DV = e.solve_multi(rng.standard_normal((n, 3)), rng.standard_normal((3, 3)), rng.standard_normal((m, 3)))
d = rng.standard_normal(m)
print("hetero", DY[0, 0], e.maxstep(v, d), e.maxstep(d), e.cone_div(d, v)[:1], e.apply(cb.OP_FINV, d)[:1])
e.close()
# one CTA per cone class (a Q cone above 1024 rows) next to small ones
cones = [("Q", 5), ("Q", 1030), ("R", 3)]
m, n = sum(k for _, k in cones), 20
e = cb.Engine(np.eye(n), rng.standard_normal((m, n)), None, cones)
v = np.zeros(m); v[0] = 3; v[5] = 40; v[1:5] = 0.1; v[6:1035] = 0.1; v[1035:] = 1.0
e.factor_from_point(v, v.copy())
print("cta-per-cone", e.solve(np.ones(n), None, np.ones(m))[0][:1], e.maxstep(v, rng.standard_normal(m)))
e.close()
# S cones through the DMMA panel kernel (order 17) and the global-workspace kernels (order 65), R rows beside them
for k in (17, 65):
    dim = k * (k + 1) // 2
    cones = [("R", 4), ("S", dim)]
    m, n = 4 + dim, 9
    A = rng.standard_normal((m, n))
    e = cb.Engine(np.eye(n), A, None, cones)
    v, sv = np.ones(m), np.ones(m)
    for x in (v, sv):
        B = rng.standard_normal((k, k)); x[4:] = O.vecm(B @ B.T + k * np.eye(k))
    e.factor_from_point(v, sv)
    print("s-cone order", k, e.solve(np.ones(n), None, np.ones(m))[0][:1], e.maxstep(v, 0.01 * rng.standard_normal(m)))
    e.close()
# scaling folded into the SYRK (R rows only)
e = cb.Engine(np.eye(150), rng.standard_normal((400, 150)), None, [("R", 400)], fold_scaling=1)
e.factor_from_point(rng.uniform(0.5, 2, 400), rng.uniform(0.5, 2, 400))
print("folded", e.solve(np.ones(150), None, np.ones(400))[0][:1])
e.close()

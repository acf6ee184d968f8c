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

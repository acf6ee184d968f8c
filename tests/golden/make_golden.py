"""Regenerates tests/golden/*.json from the oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The oracle itself is pinned to the reference's
recorded values in tests/test_oracle_goldens.py; these fixtures carry that pin to the GPU
tests as plain numbers (the GPU box has no /root/reference)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from conicip_b200 import problems as P  # noqa: E402

out = {"solves": {}, "cone_kernels": {}}
for name in ("sphere", "combined", "simplex", "soc_direct", "mixed"):
    prob = getattr(P, name)()
    tol = 1e-8 if name in ("simplex", "mixed") else prob.get("optTol", 1e-7)
    s = O.conicIP(prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"], prob["G"], prob["d"],
                  kktsolver=O.kktsolver_chol, optTol=tol)
    out["solves"][name] = dict(optTol=tol, status=s.status, Iter=s.Iter, Mu=s.Mu, y=s.y.tolist(),
                               w=s.w.tolist(), v=s.v.tolist(), prFeas=s.prFeas, duFeas=s.duFeas,
                               muFeas=s.muFeas, mu_trace=[t[1] for t in s.trace])

# cone-kernel known answers on a fixed Q^5 pair
z = np.array([2.0, 0.3, -0.5, 0.1, 0.7])
s_ = np.array([1.5, -0.2, 0.4, 0.6, -0.1])
W = O.nestod_soc(z, s_)
d = np.array([0.4, -1.0, 0.2, 0.9, -0.3])
out["cone_kernels"] = dict(
    z=z.tolist(), s=s_.tolist(), d=d.tolist(),
    nestod_soc_diag=W.Adiag.tolist(), nestod_soc_w=W.B.tolist(),
    lam=W.mul(z).tolist(), lam_alt=W.inv().mul(s_).tolist(),
    maxstep_soc=O.maxstep_soc(z, d), maxstep_soc_nothing=O.maxstep_soc(d, None),
    xsoc=O.xsoc(z, s_).tolist(), dsoc=O.dsoc(z, s_).tolist(),
    maxstep_rp=O.maxstep_rp(z, d), maxstep_rp_nothing=O.maxstep_rp(d, None))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path)

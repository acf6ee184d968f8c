"""Small, single-purpose workloads for `ncu --set full` captures (profiles/r02_*.md).  Each target launches the
kernel of interest a few times at a named shape and nothing else of note, so that `-k regex:... -c N` picks it:

    ncu --set full --clock-control none --import-source on -k regex:gemm_nt -c 2 -o X python scripts/ncu_targets.py syrk_fold
    targets: syrk_fold (C2 shape, W^-2 folded into the SYRK), syrk (C2 shape, Atil materialised), syrk_big (n = 16384, m = 65536),
             sweeps (n = 16384 triangular sweeps), chol (n = 8192 factorisation: potrf_diag / chol_head / updates),
             c5panel (config-5 shaped: S block of order 64 + R rows, scaled panel + cone kernels)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import torch

import conicip_b200 as cb

target = sys.argv[1]
dev = "cuda"
g = torch.Generator(device=dev)
g.manual_seed(1)


def r_problem(n, m, **kw):
    At = torch.randn((n, m), generator=g, dtype=torch.float64, device=dev) / n ** 0.5
    q = 1.0 + torch.rand(n, generator=g, dtype=torch.float64, device=dev)
    eng = cb.Engine(sp.diags(q.cpu().numpy()).tocsr(), At.t(), None, [("R", m)], **kw)
    eng._bind_stream()
    v = torch.rand(m, generator=g, dtype=torch.float64, device=dev) + 0.5
    s = torch.rand(m, generator=g, dtype=torch.float64, device=dev) + 0.5
    return eng, v, s


if target in ("syrk", "syrk_fold"):
    eng, v, s = r_problem(8192, 16384, fold_scaling=1 if target == "syrk_fold" else 2)
    eng.nt_scaling(v, s)
    for _ in range(2):
        eng.form_H()
elif target == "syrk_big":
    # n = 16384 with a quarter of config 4's rows: the tile grid, the panel sizes (far beyond L2) and the co-residency
    # pattern of the C4 SYRK at a quarter of its duration (DRAM traffic scales with m)
    eng, v, s = r_problem(16384, int(sys.argv[2]) if len(sys.argv) > 2 else 65536, fold_scaling=2)
    eng.nt_scaling(v, s)
    eng.form_H()
elif target == "sweeps":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    eng, v, s = r_problem(n, 2048)
    eng.factor_from_point(v, s)
    rhs = torch.randn(n, generator=g, dtype=torch.float64, device=dev)
    for _ in range(3):
        eng.solve_H(rhs)
elif target == "chol":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    eng, v, s = r_problem(n, 2048)
    os.environ.setdefault("CIP_CHOL_GRAPH", "0")        # plain stream launches: ncu sees every kernel by name
    eng.nt_scaling(v, s)
    for _ in range(2):
        eng.form_H()
        assert eng.factor_H() == 0
elif target == "c5panel":
    # LP rows plus one S block of order 64 (BASELINE config 5's cone mix), n columns
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    k = 64
    dim = k * (k + 1) // 2
    mr = 4096
    cones = [("R", mr), ("S", dim)]
    m = mr + dim
    At = torch.randn((n, m), generator=g, dtype=torch.float64, device=dev) / n ** 0.5
    eng = cb.Engine(sp.identity(n, format="csr"), At.t(), None, cones)
    eng._bind_stream()
    rng = np.random.default_rng(0)
    B1, B2 = rng.standard_normal((k, k)), rng.standard_normal((k, k))
    import oracle as O      # only to build vecm() of two PD matrices for the iterate
    v, s = np.ones(m), np.ones(m)
    v[mr:] = O.vecm(B1 @ B1.T + k * np.eye(k))
    s[mr:] = O.vecm(B2 @ B2.T + k * np.eye(k))
    v, s = torch.tensor(v, device=dev), torch.tensor(s, device=dev)
    d = torch.randn(m, generator=g, dtype=torch.float64, device=dev)
    for _ in range(2):
        lam = eng.nt_scaling(v, s)
        eng.form_H()
        eng.apply(cb.OP_F, d)
        eng.maxstep(v, 0.01 * d)
        eng.cone_prod(lam, d)
        eng.cone_div(d, lam)
else:
    raise SystemExit(f"unknown target {target}")
torch.cuda.synchronize()
print(target, "done", flush=True)
eng.close()

"""Exercise the HBM-bound kernels (K4-K8) at sizes far beyond L2 so that ncu's per-kernel duration and
DRAM byte counters give achieved GB/s (profiles/r01_hbm_kernels.md):
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:'mv_|scale_panel|nt_|apply_|maxstep_|prod_|div_' python scripts/hbm_kernels.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scipy.sparse as sp
import torch

import conicip_b200 as cb

dev = "cuda"
g = torch.Generator(device=dev)
g.manual_seed(0)


def run(name, n, cone_dims):
    m = sum(k for _, k in cone_dims)
    At = torch.randn((n, m), generator=g, dtype=torch.float64, device=dev)
    eng = cb.Engine(sp.identity(n, format="csr"), At.t(), None, cone_dims)
    del At
    torch.cuda.empty_cache()
    v = torch.rand(m, generator=g, dtype=torch.float64, device=dev) + 0.5
    s = torch.rand(m, generator=g, dtype=torch.float64, device=dev) + 0.5
    if cone_dims[0][0] == "Q":
        k = cone_dims[0][1]
        v.view(-1, k)[:, 0] += 10.0
        s.view(-1, k)[:, 0] += 10.0
    d = torch.randn(m, generator=g, dtype=torch.float64, device=dev)
    x = torch.randn(n, generator=g, dtype=torch.float64, device=dev)
    for _ in range(2):
        lam = eng.nt_scaling(v, s)
        eng.apply(cb.OP_F, d)
        eng.apply(cb.OP_FINVT, d)
        eng.maxstep(v, d)
        eng.maxstep(d, None)
        eng.cone_prod(lam, d)
        eng.cone_div(d, lam)
        eng.mul_A(x)
        eng.mul_A(d, trans=True)
        eng.form_H()
    print(name, "n", n, "m", m, "done", flush=True)
    eng.close()


which = sys.argv[1:] or ["R", "Q33"]
if "R" in which:
    run("R", 256, [("R", 1 << 24)])                      # 16.8 M rows: m-vectors 134 MB, A 34 GB
if "Q33" in which:
    run("Q33", 256, [("Q", 33)] * ((1 << 24) // 33))     # 508 k second-order cones of dimension 33

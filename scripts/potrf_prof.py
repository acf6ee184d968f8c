"""Phase breakdown of potrf_diag_kernel (clock64 marks; needs the -DCIP_POTRF_PROF build in csrc/build_prof)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conicip_b200._lib as L
L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), "csrc", "build_prof", "libprof.so")
import numpy as np, scipy.sparse as sp, torch
import conicip_b200 as cb
n, m = 1024, 2048
g = torch.Generator(device="cuda"); g.manual_seed(0)
A = torch.randn((n, m), generator=g, dtype=torch.float64, device="cuda").t()
eng = cb.Engine(sp.identity(n, format="csr"), A, None, [("R", m)])
v = torch.ones(m, dtype=torch.float64, device="cuda")
for _ in range(3):
    eng.factor_from_point(v, v)
torch.cuda.synchronize()
print(eng.stats()["ms_chol"])

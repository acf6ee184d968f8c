"""Three factorisations at n = 1024 (8 diagonal blocks each) for an ncu capture of potrf_diag_kernel:
   ncu --set full --clock-control none -k regex:potrf_diag -s 16 -c 1 --csv --page raw python scripts/potrf_ncu.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scipy.sparse as sp, torch
import conicip_b200 as cb
n, m = 1024, 2048
g = torch.Generator(device="cuda"); g.manual_seed(0)
A = torch.randn((n, m), generator=g, dtype=torch.float64, device="cuda").t()
eng = cb.Engine(sp.identity(n, format="csr"), A, None, [("R", m)])
v = torch.ones(m, dtype=torch.float64, device="cuda")
for _ in range(3):
    eng.factor_from_point(v, v)
torch.cuda.synchronize()
print("chol ms", eng.stats()["ms_chol"])

"""cip_solve_multi against k calls of cip_solve at a tall problem (the two products with A dominate a solve):
usage: python scripts/solve_multi_timing.py [n m]   (default 16384 131072: half of config 4's rows, 17 GB)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scipy.sparse as sp
import torch

import conicip_b200 as cb

n, m = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16384, 131072)
g = torch.Generator(device="cuda"); g.manual_seed(3)
At = torch.randn((n, m), generator=g, dtype=torch.float64, device="cuda") / n ** 0.5
q = 1.0 + torch.rand(n, generator=g, dtype=torch.float64, device="cuda")
eng = cb.Engine(sp.diags(q.cpu().numpy()).tocsr(), At.t(), None, [("R", m)])
del At
torch.cuda.empty_cache()
eng._bind_stream()
v = torch.rand(m, generator=g, dtype=torch.float64, device="cuda") + 0.5
s = torch.rand(m, generator=g, dtype=torch.float64, device="cuda") + 0.5
eng.factor_from_point(v, s)
k = 4
RY = torch.randn((n, k), generator=g, dtype=torch.float64, device="cuda")
RV = torch.randn((m, k), generator=g, dtype=torch.float64, device="cuda")
cols = [(RY[:, j].contiguous(), RV[:, j].contiguous()) for j in range(k)]


def timed(fn, reps=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_multi = timed(lambda: eng.solve_multi(RY, None, RV))
t_each = timed(lambda: [eng.solve(y, None, x) for y, x in cols])
DY, _, DV = eng.solve_multi(RY, None, RV)
err = max(float(torch.linalg.vector_norm(DY[:, j] - eng.solve(*[cols[j][0], None, cols[j][1]])[0]) /
                torch.linalg.vector_norm(DY[:, j])) for j in range(k))
gb = 2 * 8 * n * m / 1e9
print(f"n={n} m={m} k={k}: cip_solve_multi {t_multi:.2f} ms ({t_multi / k:.2f} per column), {k} x cip_solve {t_each:.2f} ms "
      f"({t_each / k:.2f} per column); A is {gb / 2:.1f} GB, streamed {k} times instead of {2 * k}; max rel diff {err:.1e}", flush=True)
eng.close()

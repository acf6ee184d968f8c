"""cip_imcols at the size of SURVEY config 5's equality block (p = 1000 rows, n = 20000, 100 redundant rows):
time and achieved HBM traffic (3 passes over the row-major working copy per step)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import conicip_b200 as cb

g = torch.Generator(device="cuda"); g.manual_seed(0)
for p, n, rank in [(1000, 20000, 900), (256, 4096, 256), (4096, 4096, 3000)]:
    base = torch.randn((n, rank), generator=g, dtype=torch.float64, device="cuda")
    mix = torch.randn((rank, p), generator=g, dtype=torch.float64, device="cuda")
    mix[:, :rank] = torch.eye(rank, dtype=torch.float64, device="cuda")
    At = (base @ mix).contiguous()                    # (n, p): A = At.t() is column-major p x n
    b = (At.t() @ torch.randn(n, generator=g, dtype=torch.float64, device="cuda")).cpu().numpy()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        R, ok = cb.imcols(At.t(), b)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    steps = min(p, n, rank + 1)
    print(f"p={p} n={n} rank={rank}: kept {len(R)} consistent {ok}  {dt*1e3:.1f} ms  "
          f"(~{24.0 * p * n * steps / dt / 1e9:.0f} GB/s over {steps} effective steps)", flush=True)

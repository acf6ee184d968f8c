"""Tall-skinny reduced systems (n << m): K1 time and TFLOP/s with the split-K path of gemm_nt
(profiles/r01_tall_skinny.md).  python scripts/tall_skinny.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scipy.sparse as sp
import torch

import conicip_b200 as cb

dev = "cuda"
g = torch.Generator(device=dev)
g.manual_seed(0)
for n, m in [(256, 1 << 22), (512, 1 << 21), (1000, 1 << 20), (1000, 2000), (2048, 1 << 19), (4096, 1 << 18)]:
    At = torch.randn((n, m), generator=g, dtype=torch.float64, device=dev)
    eng = cb.Engine(sp.identity(n, format="csr"), At.t(), None, [("R", m)])
    del At
    v = torch.rand(m, generator=g, dtype=torch.float64, device=dev) + 0.5
    s = torch.rand(m, generator=g, dtype=torch.float64, device=dev) + 0.5
    best = None
    for _ in range(4):
        eng.factor_from_point(v, s)
        st = eng.stats()
        best = st if best is None or st["ms_syrk"] < best["ms_syrk"] else best
    n_pad = (n + 127) // 128 * 128
    print(f"n={n} m={m}: scale {best['ms_scale']:.3f} ms  syrk {best['ms_syrk']:.3f} ms "
          f"({m * n * n / best['ms_syrk'] / 1e9:.2f} TF algorithmic, {m * n_pad * (n_pad + 128) / best['ms_syrk'] / 1e9:.2f} TF issued)  "
          f"chol {best['ms_chol']:.3f} ms", flush=True)
    eng.close()

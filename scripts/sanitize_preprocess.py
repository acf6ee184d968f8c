"""Small cip_imcols run for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import conicip_b200 as cb
rng = np.random.default_rng(0)
base = rng.standard_normal((20, 70))
A = np.vstack([base, base[:5] - base[5:10], np.zeros((1, 70))])
b = A @ rng.standard_normal(70)
print("imcols", cb.imcols(A, b))
b[22] += 5.0
print("imcols (inconsistent)", cb.imcols(A, b))

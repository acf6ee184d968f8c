"""Host mirror of the reference preprocessor (src/preprocessor.jl) on top of `cip_imcols`.

Same names and argument meaning as the reference: `imcols(A, b, eps)` and
`preprocess_conicIP(Q, c, A, b, cone_dims, G, d; options...)`.  The rank-revealing factorisation (the
O(p^2 n) part) runs on the device; the problem is then handed to `conicIP` / `conicIP_native`."""
import ctypes as C

import numpy as np

from ._lib import check, lib
from .driver import Solution, conicIP, conicIP_native

# the dual-side matrix [Q A' G'] has n rows of length n + m + p; beyond this many elements the row-pivoted
# Gram-Schmidt (n steps, three passes over the matrix each) is not worth running and the check is skipped
DUAL_CHECK_MAX_ELEMS = 1 << 28


def _dense(M):
    return np.asarray(M.todense() if hasattr(M, "todense") else M, dtype=np.float64)


def imcols(A, b, eps=1e-8, device=-1):
    """(R, consistent) as src/preprocessor.jl:10-28: R = sorted 0-based indices of a maximal linearly
    independent set of rows of A (1-based in the reference), consistent = whether A x = b has a solution.
    Like the reference, R is empty when the system is inconsistent."""
    import torch
    if isinstance(A, torch.Tensor):                   # device-resident, column-major: pass X.t() of a contiguous (n, p) tensor
        if A.dtype != torch.float64 or A.dim() != 2 or A.stride(0) != 1:
            raise ValueError("device A must be float64 and column-major: pass X.t() of a contiguous (n, p) tensor")
        p, n = A.shape
        lda = A.stride(1) if n > 1 else max(p, 1)
        a_ptr, keepalive = A.data_ptr(), A
    else:
        Ah = np.asfortranarray(_dense(A))
        if Ah.ndim != 2:
            raise ValueError("A must be a matrix")
        p, n = Ah.shape
        lda = max(p, 1)
        a_ptr, keepalive = Ah.ctypes.data, Ah
    bh = np.ascontiguousarray(b, dtype=np.float64)
    if bh.shape != (p,):
        raise ValueError(f"b has shape {bh.shape}, expected ({p},)")
    keep = np.zeros(max(p, 1), dtype=np.int32)
    nkeep, cons = C.c_int(0), C.c_int(1)
    check(lib().cip_imcols(device, a_ptr, lda, p, n, bh.ctypes.data, float(eps), keep.ctypes.data,
                           C.byref(nkeep), C.byref(cons)))
    del keepalive
    if not cons.value:
        return np.zeros(0, dtype=np.int64), False
    return np.flatnonzero(keep[:p]).astype(np.int64), True


def preprocess_conicIP(Q, c, A, b, cone_dims, G=None, d=None, *, verbose=False, native=True,
                       dual_check="auto", **options):
    """src/preprocessor.jl:40-96.  `native=True` solves with `cip_ipm_solve`, else with the Python host driver
    (which accepts `kktsolver=`).

    `dual_check`: the reference always runs the dual rank check on [Q A' G'] (:59) and augments Q + Z.  Above
    DUAL_CHECK_MAX_ELEMS matrix elements the dense working copy of that check does not fit comfortably, so
    "auto" skips it there WITH a `RuntimeWarning` (a rank-deficient dual can then end in Error where the
    reference reaches Optimal); True forces the check at any size, False never runs it."""
    c = np.asarray(c, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    n, m = len(c), len(b)
    Gd = np.zeros((0, n)) if G is None else _dense(G).reshape(-1, n)
    d = np.zeros(0) if d is None else np.asarray(d, dtype=np.float64)
    p = Gd.shape[0]
    IP, pcons = imcols(Gd, d)                                       # :58
    dcons, ID = True, np.arange(n)
    small = n * (n + m + len(IP)) <= DUAL_CHECK_MAX_ELEMS
    if dual_check is True or (dual_check == "auto" and small):
        ID, dcons = imcols(np.hstack([_dense(Q), _dense(A).T, Gd[IP, :].T]), c)     # :59
    elif dual_check == "auto":
        import warnings
        warnings.warn(f"preprocess_conicIP: dual rank check on the {n} x {n + m + len(IP)} matrix [Q A' G'] skipped "
                      "(too large); assuming full rank.  Pass dual_check=True to force it as the reference does "
                      "(src/preprocessor.jl:59).", RuntimeWarning, stacklevel=2)
    if not (pcons and dcons):                                       # :61-64
        return Solution(np.full(n, np.nan), np.full(p, np.nan), np.full(m, np.nan), status="Infeasible")
    if verbose:
        if len(IP) != p:
            print(f"   - Removing {p - len(IP)} redundant primal constraints ")
        if len(ID) != n:
            print(f"   - Augmenting {n - len(ID)} dual constraints")
        if len(ID) == n and len(IP) == p:
            print("   - No changes made")
    Qa = Q
    if len(ID) != n:
        z = np.ones(n)
        z[ID] = 0.0                                                 # :78
        Qa = _dense(Q) + np.diag(z)
    solve = conicIP_native if native else conicIP
    Gk = Gd[IP, :] if len(IP) else None
    sol = solve(Qa, c, A, b, cone_dims, Gk, d[IP] if len(IP) else None, verbose=verbose, **options)   # :82-84
    w = np.zeros(p)
    if len(IP):
        w[IP] = sol.w                                               # :91
    sol.w = w
    return sol

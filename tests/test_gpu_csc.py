"""Sparse LEVEL 1 (`cip_create_csc`): CSC ingestion must build exactly the engine the
dense path builds, for SciPy (0-based) and Julia-style (1-based Int64) index arrays."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O
from conicip_b200 import problems as P

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def sparse_problem(seed=0, n=150, mr=120, ncones=5, k=7, p=9):
    prob = P.mixed(n=n, mr=mr, ncones=ncones, k=k, p=p, seed=seed)
    rng = np.random.default_rng(seed)
    for key in ("A", "G"):
        M = prob[key].copy()
        M[rng.random(M.shape) < 0.8] = 0.0
        prob[key] = M
    prob["A"][np.arange(min(mr, n)), np.arange(min(mr, n))] += 1.0       # keep H well conditioned
    return prob


def test_csc_engine_equals_dense_engine():
    import conicip_b200 as cb
    prob = sparse_problem()
    Q, A, G, cd = prob["Q"], prob["A"], prob["G"], prob["cone_dims"]
    e_dense = cb.Engine(Q, A, G, cd)
    e_csc = cb.Engine(sp.csc_matrix(Q), sp.csr_matrix(A), sp.coo_matrix(G), cd)     # any scipy format
    rng = np.random.default_rng(1)
    m, n, p = A.shape[0], A.shape[1], G.shape[0]
    x, u, w = rng.standard_normal(n), rng.standard_normal(m), rng.standard_normal(p)
    for a, b in ((e_dense.mul_A(x), e_csc.mul_A(x)), (e_dense.mul_A(u, trans=True), e_csc.mul_A(u, trans=True)),
                 (e_dense.mul_Q(x), e_csc.mul_Q(x)), (e_dense.mul_G(x), e_csc.mul_G(x)),
                 (e_dense.mul_G(w, trans=True), e_csc.mul_G(w, trans=True))):
        assert np.array_equal(a, b)
    assert rel(e_csc.mul_A(x), A @ x) < 1e-13
    v, s = np.ones(m), np.ones(m)
    off = 0
    for t, kk in cd:
        if t == "Q":
            v[off + 1:off + kk] = 0.1
            s[off + 1:off + kk] = -0.05
        off += kk
    e_dense.factor_from_point(v, s)
    e_csc.factor_from_point(v, s)
    ry, rw, rv = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m)
    for a, b in zip(e_dense.solve(ry, rw, rv), e_csc.solve(ry, rw, rv)):
        assert np.array_equal(a, b)
    e_dense.close(); e_csc.close()


def test_csc_julia_style_one_based_int64():
    """Call cip_create_csc directly with 1-based Int64 colptr/rowval, the way the Julia shim does."""
    import conicip_b200 as cb
    from conicip_b200._lib import Csc, Options, lib
    prob = sparse_problem(seed=3, p=0)
    Q, A, cd = sp.csc_matrix(prob["Q"]), sp.csc_matrix(prob["A"]), prob["cone_dims"]
    keep = []

    def julia(M):
        cp, rv, nz = (M.indptr + 1).astype(np.int64), (M.indices + 1).astype(np.int64), M.data.astype(np.float64)
        keep.extend([cp, rv, nz])
        return Csc(M.shape[0], M.shape[1], cp.ctypes.data, rv.ctypes.data, nz.ctypes.data, 1)

    ct = np.array([cb.CONE_R if t == "R" else cb.CONE_Q for t, _ in cd], dtype=np.int32)
    cdim = np.array([k for _, k in cd], dtype=np.int32)
    o = Options(); o.struct_size = C.sizeof(Options); o.device = -1
    h = C.c_void_p()
    qs, as_ = julia(Q), julia(A)
    rc = lib().cip_create_csc(C.byref(h), A.shape[1], C.byref(qs), C.byref(as_), None, len(cd), ct.ctypes.data,
                              cdim.ctypes.data, C.byref(o))
    assert rc == 0, cb._lib.last_error()
    x = np.arange(A.shape[1], dtype=np.float64)
    y = np.zeros(A.shape[0])
    assert lib().cip_mul_A(h, 0, x.ctypes.data, y.ctypes.data) == 0
    assert rel(y, prob["A"] @ x) < 1e-13
    yq = np.zeros(A.shape[1])
    assert lib().cip_mul_Q(h, x.ctypes.data, yq.ctypes.data) == 0
    assert rel(yq, prob["Q"] @ x) < 1e-13
    lib().cip_destroy(h)


def test_readme_qp_with_sparse_inputs():
    """C1 exactly as the README writes it: sparse Q = S'S and A = sparse identity (README.md:59-68)."""
    import conicip_b200 as cb
    prob = P.config1()
    n = len(prob["c"])
    s = cb.conicIP(sp.csc_matrix(prob["Q"]), prob["c"], sp.identity(n, format="csc"), prob["b"], prob["cone_dims"],
                   optTol=1e-8)
    so = O.conicIP(prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"], optTol=1e-8,
                   kktsolver=O.pivot(O.kktsolver_2x2))
    assert s.status == so.status == "Optimal" and abs(s.Iter - so.Iter) <= 1
    assert rel(s.y, so.y) < 1e-6 and rel(s.v, so.v) < 1e-6

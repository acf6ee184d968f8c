"""Unit identities of the oracle's cone kernels, following the reference's own unit tests
(test/runtests.jl:27-88) plus defining properties of the NT scaling."""
import json
import math
import os

import numpy as np
import pytest

import oracle as O


def test_mat_vecm_roundtrip_and_inner_product():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((5, 5)); X = X + X.T
    Y = rng.standard_normal((5, 5)); Y = Y + Y.T
    assert np.allclose(O.mat(O.vecm(X)), X)
    assert np.isclose(O.vecm(X) @ O.vecm(Y), np.trace(X @ Y))      # src/ConicIP.jl:125-126
    assert np.allclose(O.vecm(np.array([[1, 2, 3], [2, 4, 5], [3, 5, 6.]])),
                       [1, 2 * math.sqrt(2), 3 * math.sqrt(2), 4, 5 * math.sqrt(2), 6])   # :131-132


def test_veccongurance_matches_dense():
    """runtests.jl:70-77."""
    rng = np.random.default_rng(1)
    Z = O.VecCongurance(rng.random((3, 3)))
    M = Z.dense()
    x = np.ones(6)
    assert np.allclose(Z.mul(x), M @ x)
    assert np.allclose(Z.inv().mul(x), np.linalg.solve(M, x))
    assert np.allclose(Z.adjoint().dense(), M.T)
    assert Z.size == 6


def test_maxstep_sdc_infinite():
    """runtests.jl:79-82."""
    assert O.maxstep_sdc(O.vecm(-np.eye(3)), O.vecm(np.eye(3))) == math.inf


def test_symwoodbury_inverse_and_dense():
    """runtests.jl:84-86 (sparse(sw) == Matrix(sw)) restated for the rank-1 form we use."""
    rng = np.random.default_rng(2)
    W = O.SymWoodbury(rng.random(50) + 0.5, rng.standard_normal(50), 1.0)
    M = W.dense()
    assert np.allclose(W.inv().dense(), np.linalg.inv(M))
    x = rng.standard_normal(50)
    assert np.allclose(W.mul(x), M @ x)


def test_block_algebra():
    """runtests.jl:27-66 restated: Block*x, Block'*x, inv against dense."""
    rng = np.random.default_rng(3)
    B = O.Block([O.Diag(rng.random(4) + 1), O.SymWoodbury(rng.random(3) + 1, rng.standard_normal(3), 1.0),
                 O.VecCongurance(rng.random((2, 2)) + np.eye(2))])
    M = B.dense()
    x = rng.standard_normal(B.size)
    assert B.size == 10
    assert np.allclose(B.mul(x), M @ x)
    assert np.allclose(B.tmul(x), M.T @ x)
    assert np.allclose(B.inv().mul(x), np.linalg.solve(M, x))
    assert np.allclose(B.inv_adjoint().mul(x), np.linalg.solve(M.T, x))


def test_nestod_soc_defining_property():
    """W z == W^-1 s (src/ConicIP.jl:167-169)."""
    rng = np.random.default_rng(4)
    for k in (2, 3, 33, 501):
        u = rng.standard_normal(k - 1); z = np.concatenate([[np.linalg.norm(u) + 0.3], u])
        u = rng.standard_normal(k - 1); s = np.concatenate([[np.linalg.norm(u) + 0.7], u])
        W = O.nestod_soc(z, s)
        assert np.allclose(W.mul(z), W.inv().mul(s), rtol=1e-9, atol=1e-11)


def test_nestod_sdc_defining_property():
    rng = np.random.default_rng(5)
    A = rng.standard_normal((4, 4)); Z = A @ A.T + np.eye(4)
    A = rng.standard_normal((4, 4)); S = A @ A.T + np.eye(4)
    W = O.nestod_sdc(O.vecm(Z), O.vecm(S))
    assert np.allclose(W.mul(O.vecm(Z)), W.inv().adjoint().mul(O.vecm(S)), atol=1e-9)


def test_maxstep_boundaries():
    x = np.array([1.0, 2.0, 3.0]); d = np.array([2.0, -1.0, 1.0])
    a = O.maxstep_rp(x, d)
    assert a == 0.5 and np.min(x - a * d) == 0.0
    assert O.maxstep_rp(x, -np.abs(d)) == math.inf
    assert O.maxstep_rp(x, None) == 0.0 and O.maxstep_rp(np.array([1.0, -2.0]), None) == -3.0
    xq = np.array([2.0, 0.5, 0.5]); dq = np.array([1.0, -1.0, 0.3])
    a = O.maxstep_soc(xq, dq)
    y = xq - a * dq
    assert abs(y[0] - np.linalg.norm(y[1:])) < 1e-12            # lands on the cone boundary
    assert O.maxstep_soc(xq, None) == 0.0
    assert O.maxstep_soc(np.array([0.0, 3.0, 4.0]), None) == -6.0


def test_jordan_division_inverts_product():
    rng = np.random.default_rng(6)
    u = rng.standard_normal(6); lam = np.concatenate([[np.linalg.norm(u) + 1], u])
    x = rng.standard_normal(7)
    assert np.allclose(O.dsoc(O.xsoc(lam, x), lam), x)
    X = rng.standard_normal((3, 3)); X = X @ X.T + np.eye(3)
    Y = rng.standard_normal((3, 3)); Y = Y + Y.T
    assert np.allclose(O.dsdc(O.xsdc(O.vecm(X), O.vecm(Y)), O.vecm(X)), O.vecm(Y))


def test_cone_kernel_fixture():
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_small.json")))["cone_kernels"]
    z, s, d = (np.array(g[k]) for k in ("z", "s", "d"))
    W = O.nestod_soc(z, s)
    assert np.allclose(W.B, g["nestod_soc_w"], rtol=1e-13) and np.allclose(W.Adiag, g["nestod_soc_diag"], rtol=1e-13)
    assert np.allclose(g["lam"], g["lam_alt"], rtol=1e-10)
    assert O.maxstep_soc(z, d) == pytest.approx(g["maxstep_soc"], rel=1e-13)


def test_slabbed_chol_kkt_equals_kktsolver_chol():
    """bench.py's CPU leg streams A in row slabs (C4 never fits the host twice); it must be the same
    solver as `kktsolver_chol` on the same data."""
    import oracle as O
    from oracle.kkt import SlabbedCholKKT
    rng = np.random.default_rng(11)
    n, m = 40, 90
    A = rng.standard_normal((m, n))
    q = rng.uniform(1, 2, n)
    f = rng.uniform(0.5, 2, m)
    ry, rv = rng.standard_normal(n), rng.standard_normal(m)
    F = O.Block([O.Diag(f)])
    solve = O.kktsolver_chol(np.diag(q), A, np.zeros((0, n)), [("R", m)])(F, F.inv())
    dy0, _, dv0 = solve(ry, np.zeros(0), rv)
    K = SlabbedCholKKT(n, q)
    slabs = [(0, 32), (32, 64), (64, 90)]
    for lo, hi in slabs:
        K.add_rows(A[lo:hi], f[lo:hi])
    K.factor()
    rhs, t1s = ry.copy(), []
    for lo, hi in slabs:
        t1, g = K.rhs_rows(A[lo:hi], f[lo:hi], rv[lo:hi])
        t1s.append(t1)
        rhs += g
    dy = K.solve(rhs)
    dv = np.concatenate([K.dv_rows(A[lo:hi], f[lo:hi], t1, dy) for (lo, hi), t1 in zip(slabs, t1s)])
    assert np.linalg.norm(dy - dy0) <= 1e-12 * np.linalg.norm(dy0)
    assert np.linalg.norm(dv - dv0) <= 1e-12 * np.linalg.norm(dv0)

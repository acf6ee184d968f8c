"""BASELINE.json full-size checks (C2: n=8192, m=16384) through size-independent properties:
sampled entries of H against a float64 host evaluation, the Cholesky identity on probe vectors,
the 3x3 KKT residual, linearity of the solve, and a full solve to 1e-8 verified by an
independent host evaluation of the optimality conditions."""
import numpy as np
import pytest

from conicip_b200 import problems as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    import conicip_b200 as cb
    prob = P.config2()
    eng = cb.Engine(prob["Q"], prob["A"], None, prob["cone_dims"])
    yield prob, eng
    eng.close()


def test_c2_form_factor_solve_properties(c2):
    prob, eng = c2
    Q, A = prob["Q"], prob["A"]
    m, n = A.shape
    rng = np.random.default_rng(0)
    v, s = rng.uniform(1e-4, 1e4, m), rng.uniform(1e-4, 1e4, m)       # wide NT scaling, as near convergence
    eng.nt_scaling(v, s)
    eng.form_H()
    H = eng.get_H()
    d = v / s                                                         # W^-2 = 1/f^2 = v/s
    idx = rng.integers(0, n, size=(200, 2))
    idx[:, 0], idx[:, 1] = np.maximum(idx[:, 0], idx[:, 1]), np.minimum(idx[:, 0], idx[:, 1])   # lower triangle
    for i, j in idx:
        want = Q[i, j] + np.dot(A[:, i] * d, A[:, j])
        assert abs(H[i, j] - want) <= 1e-11 * (abs(want) + np.linalg.norm(A[:, i] * d) * np.linalg.norm(A[:, j]) * 1e-3)
    Hl = np.tril(H) + np.tril(H, -1).T
    assert eng.factor_H() == 0
    L = np.tril(eng.get_H())
    x = rng.standard_normal((n, 3))
    assert np.linalg.norm(L @ (L.T @ x) - Hl @ x) <= 1e-12 * np.linalg.norm(Hl @ x) * 10
    # 3x3 system residual and linearity
    ry, rv = rng.standard_normal(n), rng.standard_normal(m)
    dy, _, dv = eng.solve(ry, None, rv)
    f2 = s / v
    r1 = Q @ dy - A.T @ dv - ry
    r3 = A @ dy + f2 * dv - rv
    assert np.linalg.norm(r1) <= 1e-9 * (np.linalg.norm(ry) + np.linalg.norm(A.T @ dv))
    assert np.linalg.norm(r3) <= 1e-9 * (np.linalg.norm(rv) + np.linalg.norm(f2 * dv))
    dy2, _, dv2 = eng.solve(2.0 * ry, None, 2.0 * rv)
    assert np.array_equal(dy2, 2.0 * dy) and np.array_equal(dv2, 2.0 * dv)      # exact: scaling by 2 is exact in FP64


def test_c2_full_solve_to_1e8():
    import conicip_b200 as cb
    prob = P.config2()
    s = cb.conicIP(prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"], optTol=1e-8)
    assert s.status == "Optimal" and max(s.prFeas, s.duFeas, s.muFeas) < 1e-8
    Q, A, b, c = prob["Q"], prob["A"], prob["b"], prob["c"]
    slack = A @ s.y - b
    # independent optimality check on the host (KKT of: min 1/2 y'Qy - c'y  s.t. Ay >= b)
    assert slack.min() > -1e-7 and s.v.min() > -1e-12
    assert np.linalg.norm(Q @ s.y - c - A.T @ s.v) <= 1e-7 * (1 + np.linalg.norm(c))
    assert abs(slack @ s.v) / len(b) < 1e-7

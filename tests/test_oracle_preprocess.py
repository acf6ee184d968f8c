"""Oracle restatement of the preprocessor against the reference's own preprocessor tests
(test/runtests.jl:653-676 "imcols correctness", :357-441 preprocessor tests).  The reference draws its data
from Julia's RNG, which is not reproducible here, so the assertions are the reference's (rank, consistency
flag, agreement of two formulations, status), on NumPy-generated data of the same shapes."""
import numpy as np

import oracle as O


def test_imcols_correctness_like_reference():
    rng = np.random.default_rng(42)
    A = rng.standard_normal((5, 10))                      # well-conditioned, full rank   (:657-661)
    b = rng.standard_normal(5)
    R, ok = O.imcols(A, b)
    assert len(R) == np.linalg.matrix_rank(A) == 5 and ok
    A2 = np.vstack([A, A[0:1] + A[1:2]])                  # redundant row                 (:664-668)
    b2 = np.r_[b, b[0] + b[1]]
    R2, ok2 = O.imcols(A2, b2)
    assert len(R2) == np.linalg.matrix_rank(A2) == 5 and ok2
    assert np.linalg.matrix_rank(A2[R2]) == 5
    A3 = np.vstack([A, A[0:1]])                           # inconsistent                  (:671-674)
    b3 = np.r_[b, b[0] + 100.0]
    R3, ok3 = O.imcols(A3, b3)
    assert not ok3 and len(R3) == 0
    R4, ok4 = O.imcols(np.zeros((0, 7)), np.zeros(0))     # empty matrix                  (src/preprocessor.jl:15)
    assert ok4 and len(R4) == 0


def test_preprocessor_redundant_equalities_like_reference():
    """test/runtests.jl:357-392: duplicated equality rows give the solution of the equivalent problem
    that states the equalities as pairs of inequalities."""
    rng = np.random.default_rng(0)
    n = 10
    h = rng.standard_normal(n)
    H = np.outer(h, h)
    c = np.arange(1.0, n + 1)
    A, b = np.eye(n), np.zeros(n)
    G0 = rng.random((6, n))
    G, d = np.vstack([G0, G0]), np.zeros(12)
    s1 = O.preprocess_conicIP(H, H @ c, A, b, [("R", n)], G, d, optTol=1e-8)
    s2 = O.preprocess_conicIP(H, H @ c, np.vstack([A, G, -G]), np.r_[b, d, -d], [("R", n + 24)], G, d, optTol=1e-8)
    assert s1.status == "Optimal" and s2.status == "Optimal"
    assert np.linalg.norm(s1.y - s2.y) < 1e-3             # the reference's `tol`
    assert len(s1.w) == 12


def test_preprocessor_bad_dual_constraints_like_reference():
    """test/runtests.jl:394-412: Q = 0 and A = [I I] leave ten variables without dual constraints."""
    n = 10
    Q = np.zeros((2 * n, 2 * n))
    c = -np.ones(2 * n)
    A = np.hstack([np.eye(n), np.eye(n)])
    sol = O.preprocess_conicIP(Q, c, A, np.zeros(n), [("R", n)], optTol=1e-8)
    assert np.linalg.norm(sol.y) < 1e-3


def test_preprocessor_infeasible_like_reference():
    """test/runtests.jl:414-441: x1 = 1 and x1 = -1."""
    rng = np.random.default_rng(0)
    n = 10
    h = rng.standard_normal(n)
    H = np.outer(h, h)
    c = np.arange(1.0, n + 1)
    G = np.zeros((2, n))
    G[:, 0] = 1.0
    sol = O.preprocess_conicIP(H, H @ c, np.eye(n), np.zeros(n), [("R", n)], G, np.array([1.0, -1.0]), optTol=1e-8)
    assert sol.status == "Infeasible"

"""`cip_imcols` (device rank repair) and the `preprocess_conicIP` mirror against the oracle
restatement of src/preprocessor.jl, on the reference's own preprocessor test cases (test/runtests.jl:357-441,
:653-676) and on larger seeded matrices."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def same_row_space(A, R1, R2, tol=1e-9):
    """two index sets span the same row space (the choice among dependent rows is not unique)"""
    if len(R1) != len(R2):
        return False
    if len(R1) == 0:
        return True
    r = np.linalg.matrix_rank(A[R1], tol=tol * np.linalg.norm(A))
    return r == len(R1) == np.linalg.matrix_rank(np.vstack([A[R1], A[R2]]), tol=tol * np.linalg.norm(A))


def test_imcols_reference_cases():
    import conicip_b200 as cb
    rng = np.random.default_rng(42)
    A = rng.standard_normal((5, 10))
    b = rng.standard_normal(5)
    R, ok = cb.imcols(A, b)
    assert ok and list(R) == [0, 1, 2, 3, 4]
    A2 = np.vstack([A, A[0:1] + A[1:2]])
    b2 = np.r_[b, b[0] + b[1]]
    R2, ok2 = cb.imcols(A2, b2)
    Ro, oko = O.imcols(A2, b2)
    assert ok2 and oko and len(R2) == len(Ro) == 5 and same_row_space(A2, R2, Ro)
    A3 = np.vstack([A, A[0:1]])
    b3 = np.r_[b, b[0] + 100.0]
    R3, ok3 = cb.imcols(A3, b3)
    assert not ok3 and len(R3) == 0
    R4, ok4 = cb.imcols(np.zeros((0, 7)), np.zeros(0))
    assert ok4 and len(R4) == 0


@pytest.mark.parametrize("p,n,rank,seed", [(40, 300, 25, 1), (130, 97, 60, 2), (257, 1000, 257, 3), (64, 33, 33, 4)])
def test_imcols_rank_and_consistency_against_oracle(p, n, rank, seed):
    """rows = random combinations of `rank` base rows, scaled over four orders of magnitude"""
    import conicip_b200 as cb
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((rank, n))
    mix = rng.standard_normal((p, rank))
    mix[:rank] = np.eye(rank)[rng.permutation(rank)] if p >= rank else mix[:rank]
    A = (mix @ base) * (10.0 ** rng.uniform(-2, 2, (p, 1)))
    x0 = rng.standard_normal(n)
    b = A @ x0
    R, ok = cb.imcols(A, b)
    Ro, oko = O.imcols(A, b)
    assert ok and oko
    assert len(R) == len(Ro) == min(rank, p, n)
    assert same_row_space(A, R, Ro)
    # an inconsistent right-hand side on one dependent row (if there is one)
    dep = np.setdiff1d(np.arange(p), R)
    if len(dep):
        b_bad = b.copy()
        b_bad[dep[0]] += 10.0 * np.linalg.norm(A)
        Rb, okb = cb.imcols(A, b_bad)
        assert not okb and len(Rb) == 0 and not O.imcols(A, b_bad)[1]


def test_imcols_device_resident_input_and_zero_rows():
    import torch
    import conicip_b200 as cb
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    At = torch.randn((50, 20), generator=g, dtype=torch.float64, device="cuda")      # (n, p) contiguous -> A = At.t()
    At[:, 7] = 0.0                                                                    # a zero row of A
    At[:, 9] = At[:, 3] - 2.0 * At[:, 4]
    b = torch.zeros(20, dtype=torch.float64)
    R, ok = cb.imcols(At.t(), b.numpy())
    assert ok and 7 not in R and len(R) == 18
    Ro, _ = O.imcols(At.t().cpu().numpy(), b.numpy())
    assert len(Ro) == 18


@pytest.mark.parametrize("native", [True, False])
def test_preprocess_redundant_equalities(native):
    """test/runtests.jl:357-392"""
    import conicip_b200 as cb
    rng = np.random.default_rng(0)
    n = 10
    h = rng.standard_normal(n)
    H = np.outer(h, h)
    c = np.arange(1.0, n + 1)
    A, b = np.eye(n), np.zeros(n)
    G0 = rng.random((6, n))
    G, d = np.vstack([G0, G0]), np.zeros(12)
    s = cb.preprocess_conicIP(H, H @ c, A, b, [("R", n)], G, d, optTol=1e-8, native=native)
    so = O.preprocess_conicIP(H, H @ c, A, b, [("R", n)], G, d, optTol=1e-8, kktsolver=O.pivot(O.kktsolver_2x2))
    assert s.status == so.status == "Optimal" and abs(s.Iter - so.Iter) <= 1
    assert np.linalg.norm(s.y - so.y) < 1e-6 * max(1.0, np.linalg.norm(so.y))
    assert rel(s.v, so.v) < 1e-6
    assert np.linalg.norm(G.T @ s.w - G.T @ so.w) < 1e-6 * max(1.0, np.linalg.norm(G.T @ so.w))   # w itself depends on the rows kept
    assert len(s.w) == 12 and np.count_nonzero(s.w) <= 6
    s2 = cb.preprocess_conicIP(H, H @ c, np.vstack([A, G, -G]), np.r_[b, d, -d], [("R", n + 24)], G, d, optTol=1e-8,
                               native=native)
    assert np.linalg.norm(s.y - s2.y) < 1e-3


def test_preprocess_bad_dual_constraints_and_infeasible():
    """test/runtests.jl:394-441"""
    import conicip_b200 as cb
    n = 10
    Q = np.zeros((2 * n, 2 * n))
    A = np.hstack([np.eye(n), np.eye(n)])
    sol = cb.preprocess_conicIP(Q, -np.ones(2 * n), A, np.zeros(n), [("R", n)], optTol=1e-8)
    so = O.preprocess_conicIP(Q, -np.ones(2 * n), A, np.zeros(n), [("R", n)], optTol=1e-8)
    assert sol.status == so.status and np.linalg.norm(sol.y) < 1e-3
    rng = np.random.default_rng(0)
    h = rng.standard_normal(n)
    H = np.outer(h, h)
    G = np.zeros((2, n))
    G[:, 0] = 1.0
    bad = cb.preprocess_conicIP(H, H @ np.arange(1.0, n + 1), np.eye(n), np.zeros(n), [("R", n)], G,
                                np.array([1.0, -1.0]), optTol=1e-8)
    assert bad.status == "Infeasible" and np.all(np.isnan(bad.y))

"""CPU restatement of the reference preprocessor (src/preprocessor.jl) -- TEST INFRASTRUCTURE ONLY.

`imcols` (src/preprocessor.jl:10-28) and `preprocess_conicIP` (:40-96).  The reference finds the
independent rows through SuiteSparse's sparse QR of A' (`F.R`, `F.pcol`); SuiteSparse is not available
here, so the restatement uses LAPACK's QR with column pivoting (`scipy.linalg.qr(A.T, pivoting=True)`),
which answers the same two questions -- how many rows are independent at the tolerance
`|R_ii| / ||A||_F > eps`, and is `A x = b` consistent -- but may pick a different (equally valid) subset
when rows are dependent.  Parity on the *choice* of rows is therefore unpinned; the reference's own tests
(test/runtests.jl:653-676, :357-441) only assert the rank, the consistency flag and the solution `y`."""
import numpy as np
import scipy.linalg as sla

from .conicip import Solution, conicIP


def imcols(A, b, eps=1e-8):
    """(R, consistent): sorted 0-based indices of a maximal independent set of rows of A; whether A x = b
    is consistent.  src/preprocessor.jl:10-28 (1-based there)."""
    A = np.asarray(A.todense() if hasattr(A, "todense") else A, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if A.size == 0:                                   # :15
        return np.zeros(0, dtype=np.int64), True
    nA = np.linalg.norm(A)                            # :13  (Frobenius)
    A = A / nA
    b = b / nA
    Qm, Rm, piv = sla.qr(A.T, mode="economic", pivoting=True)      # :17-21
    diag = np.abs(np.diag(Rm))
    r = int(np.count_nonzero(diag > eps))             # dgeqp3's diagonal is non-increasing in magnitude
    sel = piv[:r]
    R = np.sort(sel)                                  # :22
    if len(R) == 0:                                   # :24
        return np.zeros(0, dtype=np.int64), True
    # A[R,:] \ b[R] (:26): Julia solves the wide system through a QR factorisation (minimum-norm solution); the
    # same here from the factorisation at hand, A[sel,:]' = Qm[:, :r] Rm[:r, :r].  (An SVD-based lstsq leaves a
    # ten times larger residual, enough to flip the absolute 1e-8 test on badly scaled data: Miles problem 3.)
    y = sla.solve_triangular(Rm[:r, :r], b[sel], trans="T", lower=False)
    x = Qm[:, :r] @ y
    ok = bool(np.linalg.norm(A @ x - b, np.inf) < eps)
    return (R, True) if ok else (np.zeros(0, dtype=np.int64), False)


def preprocess_conicIP(Q, c, A, b, cone_dims, G=None, d=None, **options):
    """src/preprocessor.jl:40-96: drop redundant equality rows, augment Q on dual-deficient variables."""
    dense = lambda M: np.asarray(M.todense() if hasattr(M, "todense") else M, dtype=np.float64)
    Q, A = dense(Q), dense(A)
    c, b = np.asarray(c, dtype=np.float64), np.asarray(b, dtype=np.float64)
    n, m = len(c), A.shape[0]
    G = np.zeros((0, n)) if G is None else dense(G).reshape(-1, n)
    d = np.zeros(0) if d is None else np.asarray(d, dtype=np.float64)
    p = G.shape[0]
    IP, pcons = imcols(G, d)                                        # :58
    ID, dcons = imcols(np.hstack([Q, A.T, G[IP, :].T]), c)          # :59
    if not (pcons and dcons):                                       # :61-64
        return Solution(np.full(n, np.nan), np.full(p, np.nan), np.full(m, np.nan), status="Infeasible")
    z = np.ones(n)
    z[ID] = 0.0                                                     # :78
    sol = conicIP(Q + np.diag(z), c, A, b, cone_dims, G[IP, :], d[IP], **options)   # :82-84
    w = np.zeros(p)
    w[IP] = sol.w                                                   # :91
    sol.w = w
    return sol

"""Oracle: KKT solvers (test infrastructure).

Restates /root/reference/src/kktsolvers.jl:18-58 (kktsolver_qr),
:281-310 (kktsolver_2x2) and :316-349 (pivotgen / pivot) with dense
NumPy/SciPy LAPACK calls.  ``kktsolver_chol`` is NOT in the reference: it is
the CPU model of what the B200 engine computes (Cholesky of H + Schur
complement on G) through the same LAPACK routines, used as the timed CPU
baseline and as a second checker.

All three follow the reference's 3-level closure protocol
(docs/src/guides/kkt_solvers.md:84-109):
    kktsolver(Q, A, G, cone_dims) -> solve3x3gen(F, Finvt) -> solve3x3(y, w, v) -> (a, b, c)
"""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp


def _dense(M):
    return M.toarray() if sp.issparse(M) else np.asarray(M, dtype=np.float64)


def kktsolver_qr(Q, A, G, cone_dims):
    """src/kktsolvers.jl:18-58 -- CVXOPT double-QR."""
    Q = _dense(Q)
    A = _dense(A)
    G = _dense(G)
    n = Q.shape[0]
    p = G.shape[0]
    Q0, R1 = np.linalg.qr(G.T.reshape(n, p), mode="complete")      # :24-26
    R1 = R1[:p, :]
    Q1 = Q0[:, :p]
    Q2 = Q0[:, p:]

    def solve3x3gen(F, Finvt):
        Fi = F.inv()
        FinvT = Fi.dense().T                                        # :32
        Atil = FinvT @ A                                            # :33
        QpAtA = Q + Atil.T @ Atil                                   # :34
        Lq, Lr = np.linalg.qr(Q2.T @ QpAtA @ Q2)                    # :35

        def Lsolve(b):
            return sla.solve_triangular(Lr, Lq.T @ b) if Lr.size else np.zeros(0)

        def solve3x3(bx, by, bz):                                   # :37-52
            Q1tx = sla.solve_triangular(R1.T, by, lower=True) if p else np.zeros(0)
            rhs = bx + Atil.T @ (FinvT @ bz)
            Q2tx = Lsolve(Q2.T @ rhs - Q2.T @ (QpAtA @ (Q1 @ Q1tx)))
            if p:
                y = sla.solve_triangular(
                    R1, Q1.T @ rhs - Q1.T @ (QpAtA @ (Q1 @ Q1tx)) - Q1.T @ (QpAtA @ (Q2 @ Q2tx)))
            else:
                y = np.zeros(0)
            x = Q0 @ np.concatenate([Q1tx, Q2tx])                   # Q0'\[..] == Q0*[..]
            Fz = FinvT @ bz - Atil @ (Q1 @ Q1tx) - Atil @ (Q2 @ Q2tx)
            z = Fi.mul(Fz)
            return x, y, z

        return solve3x3

    return solve3x3gen


def kktsolver_2x2(Q, A, G, cone_dims):
    """src/kktsolvers.jl:281-310 -- form Q + A'F^-1F^-T A, LU of [H G'; G 0]."""
    Q = _dense(Q)
    A = _dense(A)
    G = _dense(G)
    n = Q.shape[0]
    p = G.shape[0]

    def solve2x2gen(F, Finvt):
        FA = Finvt.mul(A)                                           # :289-290
        H = A.T @ Finvt.tmul(FA)
        Z = np.zeros((n + p, n + p))
        Z[:n, :n] = Q + H
        Z[:n, n:] = G.T
        Z[n:, :n] = G
        lu = sla.lu_factor(Z)                                       # :295

        def solve2x2(dy, dw):                                       # :297-302
            z = sla.lu_solve(lu, np.concatenate([dy, dw]))
            return z[:n], z[n:]

        return solve2x2

    return solve2x2gen


def pivot(k2x2):
    """src/kktsolvers.jl:316-349."""

    def kktsolver(Q, A, G, cone_dims):
        solve2x2gen = k2x2(Q, A, G, cone_dims)
        Ad = A if sp.issparse(A) else np.asarray(A, dtype=np.float64)

        def solve3x3gen(F, Finvt):
            solve2x2 = solve2x2gen(F, Finvt)

            def solve3x3(y, w, v):
                t1 = Finvt.mul(Finvt.mul(v))                        # :326
                dy, dw = solve2x2(y + Ad.T @ t1, w)                 # :327
                t1 = t1 - Finvt.mul(Finvt.mul(Ad @ dy))             # :328
                return dy, dw, t1

            return solve3x3

        return solve3x3gen

    return kktsolver


def kktsolver_chol(Q, A, G, cone_dims):
    """CPU model of the B200 engine (not in the reference): same 3x3 contract as
    pivot(kktsolver_2x2) (src/kktsolvers.jl:324-332) but H = Q + Atil'Atil is
    formed as in kktsolver_qr (:33-34), factored by Cholesky (dpotrf), and the
    equality block is eliminated by the Schur complement S = G H^-1 G'."""
    Q = _dense(Q)
    A = _dense(A)
    G = _dense(G)
    n = Q.shape[0]
    p = G.shape[0]

    def solve3x3gen(F, Finvt):
        Atil = Finvt.mul(A)
        H = Q + Atil.T @ Atil
        L = sla.cholesky(H, lower=True, check_finite=False)
        if p:
            Y = sla.solve_triangular(L, G.T, lower=True, check_finite=False)
            S = Y.T @ Y
            LS = sla.cholesky(S, lower=True, check_finite=False)

        def hsolve(r):
            t = sla.solve_triangular(L, r, lower=True, check_finite=False)
            return sla.solve_triangular(L.T, t, lower=False, check_finite=False)

        Finv = F.inv()

        def solve3x3(y, w, v):
            t1 = Finv.mul(Finvt.mul(v))              # inv(F'F) v (== pivot's F^-T F^-T v for symmetric F)
            ry = y + A.T @ t1
            if p:
                u = hsolve(ry)
                rw = G @ u - w
                t = sla.solve_triangular(LS, rw, lower=True, check_finite=False)
                dw = sla.solve_triangular(LS.T, t, lower=False, check_finite=False)
                dy = hsolve(ry - G.T @ dw)
            else:
                dy = hsolve(ry)
                dw = np.zeros(0)
            dv = t1 - Finv.mul(Finvt.mul(A @ dy))
            return dy, dw, dv

        return solve3x3

    return solve3x3gen


class SlabbedCholKKT:
    """`kktsolver_chol` for K = R^m with A streamed in row slabs (the C4 shape, 34 GB, is never held on the
    host at once): the same arithmetic and the same BLAS/LAPACK routines Julia's LinearAlgebra dispatches to,
    slab by slab.  Used by bench.py's CPU legs (and pinned to `kktsolver_chol` by tests/test_oracle_units.py).

      add_rows   Atil = F^-T*A (src/kktsolvers.jl:33);  H += Atil'Atil -> BLAS dsyrk (:34; `syrk_wrapper!`)
      factor     cholesky(Q + H) -> LAPACK dpotrf      (stands in for qr(Q2'HQ2) :35 / lu([H G';G 0]) :295)
      rhs_rows   t1 = F^-T F^-T v;  y + A't1           (:326-327, slab share of the GEMV)
      solve      L \\ . , L' \\ .   -> dtrsv            (:299)
      dv_rows    t1 - F^-T F^-T (A dy)                 (:328, slab share of the GEMV)
    """

    def __init__(self, n, q_diag):
        self.n = n
        self.q_diag = np.asarray(q_diag, dtype=np.float64)
        self.H = np.zeros((n, n), order="F")
        self.L = None

    def add_rows(self, A_slab, f_slab):
        Atil = A_slab * (1.0 / f_slab)[:, None]                    # F^-T A, F = Diagonal(f)
        # Atil is C-ordered (rows x n): its transpose view is the Fortran-ordered n x rows matrix X, H += X X'
        self.H = sla.blas.dsyrk(1.0, Atil.T, beta=1.0, c=self.H, trans=0, lower=1, overwrite_c=1)

    def factor(self):
        Hq = self.H
        Hq[np.diag_indices(self.n)] += self.q_diag
        self.L = sla.cholesky(Hq, lower=True, check_finite=False, overwrite_a=True)
        self.H = None
        return self.L

    @staticmethod
    def rhs_rows(A_slab, f_slab, v_slab):
        t1 = v_slab / (f_slab * f_slab)
        return t1, A_slab.T @ t1

    def solve(self, rhs):
        t = sla.solve_triangular(self.L, rhs, lower=True, check_finite=False)
        return sla.solve_triangular(self.L, t, lower=True, trans="T", check_finite=False)

    @staticmethod
    def dv_rows(A_slab, f_slab, t1, dy):
        return t1 - (A_slab @ dy) / (f_slab * f_slab)

"""Oracle: the interior-point driver (test infrastructure).

Line-by-line NumPy restatement of /root/reference/src/ConicIP.jl:468-939
(`conicIP`): setup :513-565, cone closures :571-665, 4x4->3x3 reduction
:669-694, initial point :704-713, main loop :730-934.
"""
import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from . import cones as C
from .kkt import kktsolver_qr


@dataclass
class Solution:
    """src/ConicIP.jl:384-398."""
    y: np.ndarray
    w: np.ndarray
    v: np.ndarray
    status: str = "None"
    Iter: int = 0
    Mu: float = 0.0
    prFeas: float = math.inf
    duFeas: float = math.inf
    muFeas: float = math.inf
    pobj: float = math.inf
    dobj: float = -math.inf
    trace: list = field(default_factory=list)     # oracle-only: per-iteration (Iter, mu, rDu, rPr, rCp)
    solves: int = 0                                # oracle-only: number of LEVEL-3 calls
    factors: int = 0                               # oracle-only: number of LEVEL-2 calls


class V4:
    """src/ConicIP.jl:57-66 -- v4x1."""
    __slots__ = ("y", "w", "v", "s")

    def __init__(self, y, w, v, s):
        self.y, self.w, self.v, self.s = y, w, v, s

    def __sub__(a, b):
        return V4(a.y - b.y, a.w - b.w, a.v - b.v, a.s - b.s)

    def norm(a):
        return _nrm(a.y) + _normsafe(a.w) + _normsafe(a.v) + _normsafe(a.s)


def _nrm(x):
    return float(np.linalg.norm(x))


def _normsafe(x):
    """src/ConicIP.jl:51."""
    return 0.0 if len(x) == 0 else _nrm(x)


def _mul(M, x):
    return np.asarray(M @ x).ravel()


def conicIP(Q, c, A, b, cone_dims, G=None, d=None, *,
            kktsolver=kktsolver_qr, optTol=1e-6, DTB=0.01, verbose=False,
            maxRefinementSteps=3, maxIters=100, infeasTol=None,
            refinementThreshold=None):
    """src/ConicIP.jl:468-939."""
    c = np.asarray(c, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    n = len(c)
    if G is None:
        G = sp.csr_matrix((0, n))
    if d is None:
        d = np.zeros(0)
    d = np.asarray(d, dtype=np.float64)
    if infeasTol is None:
        infeasTol = optTol
    if refinementThreshold is None:
        refinementThreshold = optTol / 1e7

    At = A.T
    Gt = G.T
    m = A.shape[0]
    p = G.shape[0]

    block_types = [t for t, _ in cone_dims]
    block_sizes = [int(k) for _, k in cone_dims]
    offs = np.concatenate([[0], np.cumsum(block_sizes)]).astype(int)
    block_data = [(block_types[i], slice(offs[i], offs[i + 1]), i) for i in range(len(cone_dims))]

    normc = _nrm(c)
    normd = -math.inf if len(d) == 0 else _nrm(d)
    normb = _normsafe(b)

    # :536-542
    if Q.shape[0] != Q.shape[1]:
        raise ValueError("Q is not square")
    if b.shape[0] != m:
        raise ValueError("Inconsistency in inequalities")
    if A.shape[1] != n or Q.shape[0] != n:
        raise ValueError("Inconsistency in inequalities/objective")
    if d.shape[0] != p:
        raise ValueError("Inconsistency in equalities")
    if G.shape[1] != n:
        raise ValueError("Inconsistency in equalities/objective")
    if offs[-1] != m:
        raise ValueError("cone_dims do not cover the rows of A")

    # :547-565
    conedim = 0
    e = np.zeros(m)
    for btype, I, i in block_data:
        mi = I.stop - I.start
        if btype == "R":
            conedim += mi
            e[I] = 1.0
        elif btype == "Q":
            conedim += 1
            e[I.start] = 1.0
        elif btype == "S":
            k = C.ord_(mi)
            conedim += k
            e[I] = C.vecm(np.eye(k))

    def maxstep(x, dd):                                            # :571-587
        mn = math.inf
        for btype, I, i in block_data:
            xI = x[I]
            dI = None if dd is None else dd[I]
            if btype == "R":
                a = C.maxstep_rp(xI, dI)
            elif btype == "Q":
                a = C.maxstep_soc(xI, dI)
            else:
                a = C.maxstep_sdc(xI, dI)
            mn = min(a, mn)
        return mn

    def nt_scaling(x, y):                                          # :589-605
        B = []
        for btype, I, i in block_data:
            xI = x[I]
            yI = y[I]
            if btype == "R":
                B.append(C.Diag(np.sqrt(yI / xI)))
            elif btype == "Q":
                B.append(C.nestod_soc(xI, yI))
            else:
                B.append(C.nestod_sdc(xI, yI))
        return C.Block(B)

    def cone_div(x, y):                                            # :622-635  o = y^-1 o x
        o = np.zeros(len(x))
        for btype, I, i in block_data:
            if btype == "R":
                o[I] = C.drp(x[I], y[I])
            elif btype == "Q":
                o[I] = C.dsoc(x[I], y[I])
            else:
                o[I] = C.dsdc(x[I], y[I])
        return o

    def cone_prod(x, y):                                           # :652-665
        o = np.zeros(len(x))
        for btype, I, i in block_data:
            if btype == "R":
                o[I] = C.xrp(x[I], y[I])
            elif btype == "Q":
                o[I] = C.xsoc(x[I], y[I])
            else:
                o[I] = C.xsdc(x[I], y[I])
        return o

    counters = {"solves": 0, "factors": 0}
    solve3x3gen = kktsolver(Q, A, G, cone_dims)                    # :667

    def solve4x4gen(lam, F, Finvt):                                # :669-694
        counters["factors"] += 1
        solve3x3 = solve3x3gen(F, Finvt)

        def solve4x4(r):
            counters["solves"] += 1
            t1 = F.tmul(cone_div(r.s, lam))
            dy, dw, dv = solve3x3(r.y, r.w, r.v + t1)
            t1 = t1 - F.tmul(F.mul(dv))
            return V4(dy, dw, dv, t1)

        return solve4x4

    # ---- initial point :704-713
    I0 = C.Block([C.Diag(np.ones(k)) for k in block_sizes])
    r0 = V4(c, d, b, np.zeros(m))
    z = solve4x4gen(e, I0, I0)(r0)
    a_v = maxstep(z.v, None)
    a_s = maxstep(z.s, None)
    z.v = z.v - a_v * e
    z.s = z.s - a_s * e

    sol = Solution(z.y, z.w, z.v)                                  # aliases z (as :726)
    optBest = math.inf

    for Iter in range(1, maxIters + 1):                            # :730
        F = nt_scaling(z.v, z.s)
        Finvt = F.inv_adjoint()
        lam = F.mul(z.v)
        solve = solve4x4gen(lam, F, Finvt)

        Qy = _mul(Q, z.y)
        rleft = V4(Qy + _mul(Gt, z.w) - _mul(At, z.v),
                   _mul(G, z.y),
                   _mul(A, z.y) - z.s,
                   cone_prod(lam, lam))
        r0 = V4(rleft.y - c, rleft.w - d, rleft.v - b, rleft.s)    # :753

        mubar = float(np.dot(z.v, z.s))
        mu = mubar / conedim

        cty = float(np.dot(c, z.y))
        rDu = _nrm(r0.y) / (1 + normc)
        rPr = _normsafe(r0.v) / (1 + normb)
        rCp = _normsafe(r0.s) / (1 + abs(cty))
        sol.trace.append((Iter, mu, rDu, rPr, rCp))

        if max(rDu, rPr, rCp) < optBest:                           # :768-773 (y,w,v alias z)
            sol.Iter = Iter
            sol.Mu = mu
            sol.duFeas, sol.prFeas, sol.muFeas = rDu, rPr, rCp
            optBest = max(rDu, rPr, rCp)

        pobj = 0.5 * float(np.dot(z.y, Qy)) - cty
        dobj = pobj + float(np.dot(z.w, r0.w)) + float(np.dot(z.v, r0.v)) - mubar
        sol.pobj, sol.dobj = pobj, dobj

        if max(rDu, rPr, rCp) < optTol:                            # :786
            sol.status = "Optimal"

        if not (p == 0 and m == 0):                                # :790-852
            dty_btv = float(np.dot(d, z.w)) - float(np.dot(b, z.v))
            p_unscaled = _nrm(_mul(Gt, z.w) - _mul(At, z.v))
            with np.errstate(all="ignore"):
                p_cvx = p_unscaled / (_normsafe(z.y) + _normsafe(z.v)) if dty_btv < 0 else math.nan
                p_ecos = p_unscaled / (max(1, normc) * abs(dty_btv)) if dty_btv < 0 else math.nan
            p_infeas = float(np.maximum(p_cvx, p_ecos))
            if p_infeas < infeasTol:
                sol.y[:] = math.nan
                sol.w[:] = z.w / -dty_btv
                sol.v[:] = z.v / -dty_btv
                sol.status = "Infeasible"

            d1 = -math.inf if m == 0 else _nrm(_mul(A, z.y) - z.s)
            d2 = -math.inf if p == 0 else _nrm(_mul(G, z.y))
            d3 = _nrm(_mul(Q, z.y)) if np.all(np.isfinite(z.y)) else math.nan
            if cty > 0:
                d_cvx = max(d1 / max(1, normb), d2 / max(1, normd), d3 / max(1, normc)) / abs(cty)
                d_ecos = max(d1, d2, d3) / _nrm(z.y)
            else:
                d_cvx = d_ecos = math.nan
            d_infeas = abs(float(np.maximum(d_cvx, d_ecos)))
            if d_infeas < infeasTol:
                sol.y[:] = z.y / abs(cty)
                sol.v[:] = math.nan
                sol.w[:] = math.nan
                sol.status = "Unbounded"

        if verbose:
            print(f" {Iter:6d}  | {rDu:8.1e} {rPr:8.1e} {rCp:8.1e} | {pobj: 8.1e} {dobj: 8.1e} | mu {mu:8.1e}")

        if sol.status != "None":
            sol.solves, sol.factors = counters["solves"], counters["factors"]
            return sol

        if not all(math.isfinite(t) for t in (mu, rDu, rPr, rCp)):  # :870-873
            sol.status = "Error"
            sol.solves, sol.factors = counters["solves"], counters["factors"]
            return sol

        # ---- predictor :879-887
        d_aff = solve(r0)
        a_aff_v = min(maxstep(z.v, d_aff.v), 1)
        a_aff_s = min(maxstep(z.s, d_aff.s), 1)
        a_aff = min(a_aff_v, a_aff_s)
        rho = C.fts(z.v, a_aff, d_aff.v, z.s, a_aff, d_aff.s) / mubar
        sigma = max(0, min(1, rho)) ** 3

        # ---- corrector :893-901
        Fds = Finvt.mul(d_aff.s)
        Fdv = F.mul(d_aff.v)
        lc = cone_prod(Fds, Fdv)
        lc = -(lc - sigma * mu * e)
        r = V4(r0.y, r0.w, r0.v, rleft.s - lc)

        # ---- newton step + refinement :907-921
        dz = solve(r)
        for rStep in range(1, maxRefinementSteps + 1):
            pb1 = cone_prod(lam, F.mul(dz.v))
            pb2 = cone_prod(lam, Finvt.mul(dz.s))
            rkkt = V4(_mul(Q, dz.y) + _mul(Gt, dz.w) - _mul(At, dz.v),
                      _mul(G, dz.y),
                      _mul(A, dz.y) - dz.s,
                      pb1 + pb2)
            rIr = r - rkkt
            rnorm = rIr.norm() / (n + 2 * m)
            if rnorm < refinementThreshold:
                break
            dzr = solve(rIr)
            dz.y += dzr.y
            dz.w += dzr.w
            dz.v += dzr.v
            dz.s += dzr.s

        # ---- step :927-932
        a_v = min(maxstep(z.v, dz.v / (1 - DTB)), 1)
        a_s = min(maxstep(z.s, dz.s / (1 - DTB)), 1)
        alpha = min(a_v, a_s)
        z.y -= alpha * dz.y
        z.w -= alpha * dz.w
        z.v -= alpha * dz.v
        z.s -= alpha * dz.s

    sol.status = "Abandoned"
    sol.solves, sol.factors = counters["solves"], counters["factors"]
    return sol

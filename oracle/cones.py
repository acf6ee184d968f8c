"""Oracle: cone kernels and block-diagonal scaling matrices (test infrastructure).

Follows /root/reference/src/ConicIP.jl:35-40,69-151,160-360 and
/root/reference/src/blockmatrices.jl:35-43,107-131,173-200.  The SymWoodbury
algebra restates WoodburyMatrices.jl 0.5.6 (pinned in docs/Manifest.toml:507-511,
not vendored in the reference): ``A + B*D*B'`` with ``inv`` by the Woodbury
identity.
"""
import math

import numpy as np
import scipy.linalg as sla

SQRT2 = math.sqrt(2.0)


# ---------------------------------------------------------------- mat / vecm
def ord_(nvec):
    """src/ConicIP.jl:85 -- matrix order k from svec length k(k+1)/2."""
    return int(round((math.sqrt(1 + 8 * nvec) - 1) / 2))


def mat(x):
    """src/ConicIP.jl:93-119 -- svec (row-major upper triangle, off-diag*sqrt2) -> symmetric."""
    x = np.asarray(x, dtype=np.float64)
    n = ord_(len(x))
    Z = np.zeros((n, n))
    iu = np.triu_indices(n)          # row-major upper triangle == reference ordering
    Z[iu] = x
    off = iu[0] != iu[1]
    Z[iu[0][off], iu[1][off]] /= SQRT2
    Z = Z + np.triu(Z, 1).T
    return Z


def vecm(Z):
    """src/ConicIP.jl:128-151 -- symmetric -> svec, off-diagonals scaled by sqrt2."""
    Z = np.asarray(Z, dtype=np.float64)
    n = Z.shape[0]
    iu = np.triu_indices(n)
    x = Z[iu].copy()
    x[iu[0] != iu[1]] *= SQRT2
    return x


# ---------------------------------------------------------------- helpers
def QF(r):
    """src/ConicIP.jl:160 -- r'Jr."""
    return 2 * r[0] * r[0] - np.dot(r, r)


def Qf(x, y):
    """src/ConicIP.jl:161 -- x'Jy."""
    return 2 * x[0] * y[0] - np.dot(x, y)


def fts(x1, a1, y1, x2, a2, y2):
    """src/ConicIP.jl:162-163 -- (x1 - a1*y1)'(x2 - a2*y2)."""
    return (np.dot(x1, x2) - a2 * np.dot(x1, y2)
            - a1 * np.dot(y1, x2) + a1 * a2 * np.dot(y1, y2))


# ---------------------------------------------------------------- scaling blocks
class Diag:
    """Julia ``Diagonal`` block (R cones; also every block of the initial F=I)."""

    def __init__(self, diag):
        self.diag = np.asarray(diag, dtype=np.float64)

    @property
    def size(self):
        return len(self.diag)

    def mul(self, x):
        return self.diag * x if x.ndim == 1 else self.diag[:, None] * x

    tmul = mul

    def inv(self):
        return Diag(1.0 / self.diag)

    def adjoint(self):
        return self

    def dense(self):
        return np.diag(self.diag)


class SymWoodbury:
    """WoodburyMatrices.SymWoodbury(A::Diagonal, B::Vector, D::Real) = A + B*D*B'."""

    def __init__(self, Adiag, B, D):
        self.Adiag = np.asarray(Adiag, dtype=np.float64)
        self.B = np.asarray(B, dtype=np.float64)
        self.D = float(D)

    @property
    def size(self):
        return len(self.Adiag)

    def mul(self, x):
        if x.ndim == 1:
            return self.Adiag * x + self.B * (self.D * np.dot(self.B, x))
        return self.Adiag[:, None] * x + np.outer(self.B, self.D * (self.B @ x))

    tmul = mul  # symmetric

    def inv(self):
        # WoodburyMatrices 0.5.6: W=inv(A); X=W*B; Z=inv(-inv(D) - B'X); SymWoodbury(W,X,Z)
        W = 1.0 / self.Adiag
        X = W * self.B
        Z = 1.0 / (-(1.0 / self.D) - np.dot(self.B, X))
        return SymWoodbury(W, X, Z)

    def adjoint(self):
        return self

    def dense(self):
        return np.diag(self.Adiag) + self.D * np.outer(self.B, self.B)


class VecCongurance:
    """src/ConicIP.jl:35-40,69 -- W*x = vecm(R' mat(x) R)."""

    def __init__(self, R):
        self.R = np.asarray(R, dtype=np.float64)

    @property
    def size(self):
        k = self.R.shape[0]
        return k * (k + 1) // 2

    def _apply(self, R, x):
        if x.ndim == 1:
            return vecm(R.T @ mat(x) @ R)
        return np.stack([vecm(R.T @ mat(x[:, j]) @ R) for j in range(x.shape[1])], axis=1)

    def mul(self, x):
        return self._apply(self.R, x)

    def tmul(self, x):                      # adjoint(W) = VecCongurance(R')
        return self._apply(self.R.T, x)

    def inv(self):
        return VecCongurance(np.linalg.inv(self.R))

    def adjoint(self):
        return VecCongurance(self.R.T)

    def dense(self):
        """src/ConicIP.jl:71-79."""
        n = self.size
        return self.mul(np.eye(n))


class Block:
    """src/blockmatrices.jl:35-43 -- block-diagonal container; `*` per block (:107-131,:173-177)."""

    def __init__(self, blocks):
        self.blocks = list(blocks)
        sizes = [b.size for b in self.blocks]
        self.offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(int)

    @property
    def size(self):
        return int(self.offsets[-1])

    def _bc(self, fn, x):
        y = np.empty_like(x, dtype=np.float64)
        for i, b in enumerate(self.blocks):
            lo, hi = self.offsets[i], self.offsets[i + 1]
            y[lo:hi] = fn(b, x[lo:hi])
        return y

    def mul(self, x):
        return self._bc(lambda b, xi: b.mul(xi), np.asarray(x, dtype=np.float64))

    def tmul(self, x):
        return self._bc(lambda b, xi: b.tmul(xi), np.asarray(x, dtype=np.float64))

    def inv(self):
        return Block([b.inv() for b in self.blocks])

    def adjoint(self):
        return Block([b.adjoint() for b in self.blocks])

    def inv_adjoint(self):
        """src/blockmatrices.jl:193-198."""
        return Block([b.inv().adjoint() for b in self.blocks])

    def dense(self):
        """src/blockmatrices.jl:162-170."""
        O = np.zeros((self.size, self.size))
        for i, b in enumerate(self.blocks):
            lo, hi = self.offsets[i], self.offsets[i + 1]
            O[lo:hi, lo:hi] = b.dense()
        return O

    def __getitem__(self, i):
        return self.blocks[i]


# ---------------------------------------------------------------- NT scaling
def nestod_soc(z, s):
    """src/ConicIP.jl:165-194."""
    z = np.array(z, dtype=np.float64)
    s = np.array(s, dtype=np.float64)
    n = len(z)
    beta = (QF(s) / QF(z)) ** 0.25
    z = z / math.sqrt(QF(z))
    s = s / math.sqrt(QF(s))
    gamma = math.sqrt((1 + np.dot(z, s)) / 2)
    z = -z
    z[0] = -z[0]                                  # Jz
    w = (1.0 / (2.0 * gamma)) * (s + z)
    w[0] = w[0] + 1
    w = w * (math.sqrt(2 * beta) / math.sqrt(2 * w[0]))
    J = np.full(n, beta)
    J[0] = -beta
    return SymWoodbury(J, w, 1.0)


def nestod_sdc(z, s):
    """src/ConicIP.jl:196-210."""
    Ls = np.linalg.cholesky(mat(s))
    Lz = np.linalg.cholesky(mat(z))
    U, lam, _ = np.linalg.svd(Lz.T @ Ls)
    R = sla.solve_triangular(Lz, np.eye(Lz.shape[0]), lower=True).T @ U @ np.diag(np.sqrt(lam))
    return VecCongurance(R)


# ---------------------------------------------------------------- max step
def maxstep_rp(x, d):
    """src/ConicIP.jl:212-240."""
    if d is None:
        if np.all(x > 0):
            return 0.0
        return -1 + float(np.min(x))
    pos = d > 0
    if not np.any(pos):
        return math.inf
    return float(np.min(x[pos] / d[pos]))


def maxstep_soc(x, d):
    """src/ConicIP.jl:242-270."""
    if d is None:
        a = float(np.linalg.norm(x[1:]) - x[0])
        return 0.0 if a < 0 else -1 - a
    d = -d
    gamma = Qf(x, x)
    xbar = x / math.sqrt(gamma)
    beta = Qf(xbar, d)
    rho1 = beta / math.sqrt(gamma)
    mu = (beta + d[0]) / (xbar[0] + 1)
    rho2 = d[1:] - mu * xbar[1:]
    alpha = np.linalg.norm(rho2) / math.sqrt(gamma) - rho1
    if alpha < 0:
        return math.inf
    return 1.0 / alpha if alpha != 0 else math.inf


def maxstep_sdc(x, d):
    """src/ConicIP.jl:272-303."""
    X = mat(x)
    if d is None:
        lam = np.linalg.eigvals(X).real
        mn = float(np.min(lam))
        return 0.0 if mn > 0 else -1 + mn
    lamX, V = np.linalg.eigh(X)
    if np.any(lamX <= 0):
        return math.inf
    Xih = (V * lamX ** -0.5) @ V.T
    D = mat(d)
    XDX = Xih @ D @ Xih
    XDX = 0.5 * (XDX + XDX.T)
    lam = np.linalg.eigvalsh(XDX)
    neg = lam < 0
    if np.all(neg):
        return math.inf
    return 1.0 / float(np.max(lam[~neg]))


# ---------------------------------------------------------------- Jordan algebra
def drp(x, y):
    """src/ConicIP.jl:305-309 -- o = x ./ y."""
    return x / y


def xrp(x, y):
    """src/ConicIP.jl:311-315."""
    return x * y


def dsoc(y, x):
    """src/ConicIP.jl:317-338 -- o = arrow(x)^-1 y  (argument order as in the reference)."""
    y1 = x[0]
    yb = x[1:]
    alpha = y1 * y1 - np.dot(yb, yb)
    x1 = y[0]
    xb = y[1:]
    o = np.empty_like(y)
    ybxb = np.dot(yb, xb)
    o[0] = (y1 * x1 - ybxb) / alpha
    b1 = (-x1 / alpha) + ybxb / (y1 * alpha)
    b2 = 1 / y1
    o[1:] = yb * b1 + xb * b2
    return o


def xsoc(x, y):
    """src/ConicIP.jl:340-345."""
    o = np.empty_like(x)
    o[0] = np.dot(x, y)
    o[1:] = x[0] * y[1:] + y[0] * x[1:]
    return o


def dsdc(x, y):
    """src/ConicIP.jl:347-353 -- vecm(lyap(Y,-X)): solves Y*O + O*Y' = X."""
    X = mat(x)
    Y = mat(y)
    return vecm(sla.solve_continuous_lyapunov(Y, X))


def xsdc(x, y):
    """src/ConicIP.jl:355-360 -- vecm(XY + YX) (no 1/2)."""
    X = mat(x)
    Y = mat(y)
    return vecm(X @ Y + Y @ X)

"""Host-side data carriers for the scaling matrix F that crosses the kktsolver boundary.

They mirror the Julia block types a ConicIP `Block` holds
(/root/reference/src/blockmatrices.jl:13-43): `Diagonal` (`.diag`), WoodburyMatrices
`SymWoodbury` (`.A.diag`, `.B`, `.D`) and `VecCongurance` (`.R`, src/ConicIP.jl:35).
They carry data only -- all arithmetic happens on the device -- and know how to
flatten themselves into the (kind, fa, fb, fD, fR) arrays of `cip_factor`.
"""
import numpy as np

from ._lib import BLK_DIAG, BLK_VECCONG, BLK_WOODBURY


class Diagonal:
    kind = BLK_DIAG

    def __init__(self, diag):
        self.diag = np.ascontiguousarray(diag, dtype=np.float64)

    @property
    def size(self):
        return len(self.diag)


class SymWoodbury:
    """A + B*D*B' with diagonal A, vector B, scalar D (what nestod_soc builds,
    src/ConicIP.jl:189-192)."""
    kind = BLK_WOODBURY

    def __init__(self, A_diag, B, D):
        self.A_diag = np.ascontiguousarray(A_diag, dtype=np.float64)
        B = np.asarray(B, dtype=np.float64)
        D = np.asarray(D, dtype=np.float64)
        # WoodburyMatrices' general form (matrix B of rank > 1, matrix D: src/blockmatrices.jl:135-141,
        # src/kktsolvers.jl:73-90) is never produced by conicIP (nestod_soc builds rank 1, src/ConicIP.jl:192)
        # and the engine's flat format cannot carry it: reject it loudly instead of flattening it wrongly
        if (B.ndim == 2 and B.shape[1] != 1) or B.size != self.A_diag.size or D.size != 1:
            raise ValueError("SymWoodbury blocks of rank > 1 (matrix B / matrix D) are not supported by the "
                             "B200 engine; only the rank-1 form diag(A) + B*D*B' that nestod_soc builds is")
        self.B = np.ascontiguousarray(B).ravel()
        self.D = float(D.ravel()[0])

    @property
    def size(self):
        return len(self.A_diag)


class VecCongurance:
    kind = BLK_VECCONG

    def __init__(self, R):
        self.R = np.asfortranarray(R, dtype=np.float64)

    @property
    def size(self):
        k = self.R.shape[0]
        return k * (k + 1) // 2


class Block:
    """Block-diagonal container (src/blockmatrices.jl:35-43)."""

    def __init__(self, blocks):
        self.Blocks = list(blocks)

    def __getitem__(self, i):
        return self.Blocks[i]

    def __len__(self):
        return len(self.Blocks)

    @property
    def size(self):
        return sum(b.size for b in self.Blocks)

    @classmethod
    def from_flat(cls, cone_dims, kind, fa, fb, fD, Rs=None):
        """Inverse of `flatten`: rebuild the host Block from the arrays `cip_get_scaling` returns."""
        blocks, off, si = [], 0, 0
        for i, (t, k) in enumerate(cone_dims):
            k = int(k)
            if kind[i] == BLK_DIAG:
                blocks.append(Diagonal(fa[off:off + k]))
            elif kind[i] == BLK_WOODBURY:
                blocks.append(SymWoodbury(fa[off:off + k], fb[off:off + k], fD[i]))
            else:
                blocks.append(VecCongurance(Rs[si]))
            if t == "S":
                si += 1
            off += k
        return cls(blocks)

    def flatten(self):
        """-> (kind int32[nc], fa f64[m], fb f64[m], fD f64[nc], fR f64[sum k^2] or None)."""
        nc = len(self.Blocks)
        kind = np.zeros(nc, dtype=np.int32)
        fD = np.zeros(nc)
        fa, fb, fR = [], [], []
        for i, b in enumerate(self.Blocks):
            kind[i] = b.kind
            if b.kind == BLK_DIAG:
                fa.append(b.diag)
                fb.append(np.zeros(b.size))
            elif b.kind == BLK_WOODBURY:
                fa.append(b.A_diag)
                fb.append(b.B)
                fD[i] = b.D
            else:
                fa.append(np.zeros(b.size))
                fb.append(np.zeros(b.size))
                fR.append(b.R.ravel(order="F"))
        cat = lambda xs: np.ascontiguousarray(np.concatenate(xs)) if xs else np.zeros(0)
        return kind, cat(fa), cat(fb), fD, (cat(fR) if fR else None)


class DeviceBlock:
    """Token for "the scaling currently resident in the engine" (set by
    Engine.nt_scaling).  Passing it to solve3x3gen skips the host round trip."""

    def __init__(self, engine, inverse_adjoint=False):
        self.engine = engine
        self.inverse_adjoint = inverse_adjoint

    def to_host(self):
        """Materialise as a host `Block` (what a Julia caller would see)."""
        kind, fa, fb, fD, Rs = self.engine.get_scaling(with_R=True)
        off = self.engine.cone_off
        blocks, si = [], 0
        for i in range(len(kind)):
            lo, hi = off[i], off[i + 1]
            is_s = self.engine.cone_dims[i][0] == "S"
            if kind[i] == BLK_DIAG:
                blocks.append(Diagonal(fa[lo:hi]))
            elif kind[i] == BLK_WOODBURY:
                blocks.append(SymWoodbury(fa[lo:hi], fb[lo:hi], fD[i]))
            else:
                blocks.append(VecCongurance(Rs[si]))
            si += is_s
        return Block(blocks)

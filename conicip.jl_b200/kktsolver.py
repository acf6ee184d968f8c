"""`kktsolver_b200`: the drop-in for `kktsolver_qr` / `pivot(kktsolver_2x2)`.

Same three-level closure protocol as the reference
(/root/reference/docs/src/guides/kkt_solvers.md:84-109, src/ConicIP.jl:667,682,688):

    solve3x3gen = kktsolver_b200(Q, A, G, cone_dims)      # LEVEL 1: upload once (cip_create)
    solve3x3    = solve3x3gen(F, F_invT)                  # LEVEL 2: form H, factor (cip_factor)
    a, b, c     = solve3x3(y, w, v)                       # LEVEL 3: solve (cip_solve)

`F` is either a host `Block` of `Diagonal` / `SymWoodbury` blocks (flattened and sent to the
device, exactly what the Julia shim does) or a `DeviceBlock` token meaning "the scaling the
engine computed itself in nt_scaling" (no host round trip).  `F_invT` is accepted for
signature compatibility and ignored: the engine inverts F on the device.
"""
import numpy as np

from .blocks import Block, DeviceBlock
from .engine import Engine


def kktsolver_b200(Q, A, G, cone_dims, **engine_opts):
    eng = Engine(Q, A, G, cone_dims, **engine_opts)

    def solve3x3gen(F, F_invT=None):
        if isinstance(F, DeviceBlock):
            status = eng.factor_resident()
        elif isinstance(F, Block):
            status = eng.factor(F)
        else:
            raise TypeError("F must be a conicip_b200 Block or DeviceBlock")
        if status > 0:
            # the reference has no error channel: a failed pivot surfaces as
            # non-finite iterates -> status :Error at the next convergence check
            def solve3x3_failed(y, w, v):
                nan = lambda k: np.full(k, np.nan)
                return nan(eng.n), nan(eng.p), nan(eng.m)
            return solve3x3_failed

        def solve3x3(y, w, v):
            return eng.solve(y, w, v)

        return solve3x3

    solve3x3gen.engine = eng
    return solve3x3gen


def make_kktsolver(**engine_opts):
    """`kktsolver=make_kktsolver(reg_delta=1e-10)` -> a kktsolver with options bound."""
    def k(Q, A, G, cone_dims):
        return kktsolver_b200(Q, A, G, cone_dims, **engine_opts)
    return k

"""CPU oracle for the ConicIP.jl KKT hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy restatement of the reference algorithm
(MPF-Optimization-Laboratory/ConicIP.jl v0.2.0).  Julia is not installed in
the build image, so the reference itself cannot be executed; every function
here cites the reference ``file:line`` it follows.

Pinned against the reference's own recorded goldens (``test/runtests.jl``):
sphere ``Mu`` at iteration 5, combined R+Q ``Mu`` at iteration 10, simplex
``Mu`` at iteration 11 (see ``tests/test_oracle_goldens.py``).  Factor-level
parity (LAPACK qr/lu inside Julia) has no unit KAT in the reference and is
pinned only end-to-end through those goldens.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package, and
only as the checker / CPU baseline.  The product (``conicip.jl_b200``) never
imports it and has no CPU fallback.
"""

from .cones import (  # noqa: F401
    mat, vecm, ord_, QF, Qf, fts,
    nestod_soc, nestod_sdc,
    maxstep_rp, maxstep_soc, maxstep_sdc,
    drp, xrp, dsoc, xsoc, dsdc, xsdc,
    Diag, SymWoodbury, VecCongurance, Block,
)
from .kkt import kktsolver_qr, kktsolver_2x2, pivot, kktsolver_chol  # noqa: F401
from .conicip import conicIP, Solution  # noqa: F401
from .preprocess import imcols, preprocess_conicIP  # noqa: F401

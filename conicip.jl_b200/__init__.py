"""conicip.jl_b200 -- B200-native KKT engine behind ConicIP.jl's `kktsolver=` callback.

The directory name follows the reference repository (ConicIP.jl); because of the dot it is
imported through the top-level alias module `conicip_b200` (conicip_b200.py).

Public surface (mirrors the reference's names for this path):
    kktsolver_b200(Q, A, G, cone_dims)        drop-in for kktsolver_qr / pivot(kktsolver_2x2)
    conicIP(Q, c, A, b, cone_dims, G, d; ...) host driver with the reference's signature
    Engine                                    object wrapper over the C ABI handle
    Block, Diagonal, SymWoodbury, VecCongurance, DeviceBlock   data format of F at the boundary
"""
from ._lib import (BLK_DIAG, BLK_VECCONG, BLK_WOODBURY, CONE_Q, CONE_R, CONE_S,  # noqa: F401
                   OP_F, OP_FINV, OP_FINVT, OP_FT, CipError, LIB_PATH, SIGNATURES, lib)
from .blocks import Block, DeviceBlock, Diagonal, SymWoodbury, VecCongurance  # noqa: F401
from .engine import Engine, measure_fp64_peaks, nccl_unique_id, shard_plan  # noqa: F401
from .kktsolver import kktsolver_b200, make_kktsolver  # noqa: F401
from .driver import Solution, conicIP, conicIP_native  # noqa: F401
from .preprocess import imcols, preprocess_conicIP  # noqa: F401
from . import problems, dist  # noqa: F401

__version__ = "0.1.0"

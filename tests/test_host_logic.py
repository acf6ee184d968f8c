"""CPU tests of the host-side logic: the F data carriers, cone-aligned row sharding, and the
N>1 path (partial Gram matrices + all-reduce) on gloo with world_size 2."""
import os
import socket

import numpy as np
import pytest

import oracle as O
from conicip_b200 import blocks as B
from conicip_b200 import problems as P
from conicip_b200.dist import shard_cones


def test_block_flatten_layout():
    F = B.Block([B.Diagonal([1.0, 2.0]), B.SymWoodbury([-3.0, 3.0, 3.0], [0.1, 0.2, 0.3], 1.0), B.Diagonal([5.0])])
    kind, fa, fb, fD, fR = F.flatten()
    assert kind.tolist() == [0, 1, 0] and kind.dtype == np.int32
    assert fa.tolist() == [1, 2, -3, 3, 3, 5] and fb.tolist() == [0, 0, 0.1, 0.2, 0.3, 0]
    assert fD.tolist() == [0, 1, 0] and fR is None and F.size == 6
    assert fa.flags.c_contiguous and fa.dtype == np.float64


def test_block_flatten_matches_oracle_nt_scaling():
    rng = np.random.default_rng(0)
    u = rng.standard_normal(4); z = np.concatenate([[np.linalg.norm(u) + 1], u])
    u = rng.standard_normal(4); s = np.concatenate([[np.linalg.norm(u) + 1], u])
    W = O.nestod_soc(z, s)
    kind, fa, fb, fD, _ = B.Block([B.SymWoodbury(W.Adiag, W.B, W.D)]).flatten()
    x = rng.standard_normal(5)
    assert np.allclose(fa * x + fD[0] * fb * (fb @ x), W.mul(x))


@pytest.mark.parametrize("cones,nr", [([("R", 100)], 4), ([("Q", 33)] * 8, 4), ([("R", 10), ("Q", 5), ("Q", 5), ("R", 7)], 2),
                                      ([("R", 262144)], 8), ([("Q", 7)] * 3, 8), ([("R", 5)], 1)])
def test_shard_cones_partitions_rows_on_cone_boundaries(cones, nr):
    sh = shard_cones(cones, nr)
    m = sum(k for _, k in cones)
    assert len(sh) == nr and sh[0][0] == 0 and sh[-1][1] == m
    for (lo, hi, cd), nxt in zip(sh, sh[1:] + [None]):
        assert sum(k for _, k in cd) == hi - lo
        if nxt:
            assert nxt[0] == hi
    # Q cones are never cut
    assert sorted(k for _, _, cd in sh for t, k in cd if t == "Q") == sorted(k for t, k in cones if t == "Q")
    if all(t == "R" for t, _ in cones) and m >= nr:
        sizes = [hi - lo for lo, hi, _ in sh]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conicip_b200.dist import TorchReducer
    prob = P.mixed(n=24, mr=20, ncones=4, k=5, p=0, seed=11)
    A, cd = prob["A"], prob["cone_dims"]
    lo, hi, lcd = shard_cones(cd, world)[rank]
    rng = np.random.default_rng(5)
    m = A.shape[0]
    v = np.zeros(m); s = np.zeros(m); off = 0
    for t, k in cd:
        if t == "R":
            v[off:off + k] = rng.uniform(0.5, 2, k); s[off:off + k] = rng.uniform(0.5, 2, k)
        else:
            for x in (v, s):
                u = rng.standard_normal(k - 1); x[off] = np.linalg.norm(u) + 0.5; x[off + 1:off + k] = u
        off += k

    def scaling(cdims, vv, ss):
        bl, o = [], 0
        for t, k in cdims:
            bl.append(O.Diag(np.sqrt(ss[o:o + k] / vv[o:o + k])) if t == "R" else O.nestod_soc(vv[o:o + k], ss[o:o + k]))
            o += k
        return O.Block(bl)

    Fl = scaling(lcd, v[lo:hi], s[lo:hi]).inv_adjoint()
    At = Fl.mul(A[lo:hi])
    Hpart = torch.from_numpy(At.T @ At + (prob["Q"] if rank == 0 else 0 * prob["Q"]))
    dist.all_reduce(Hpart)                                    # the engine does this with NCCL
    Ff = scaling(cd, v, s).inv_adjoint()
    Atf = Ff.mul(A)
    Hfull = prob["Q"] + Atf.T @ Atf
    red = TorchReducer()
    tot = red.sum(float(v[lo:hi] @ s[lo:hi]))
    mn = red.min(float(np.min(v[lo:hi])))
    q.put((rank, float(np.abs(Hpart.numpy() - Hfull).max()), abs(tot - v @ s), abs(mn - v.min())))
    dist.destroy_process_group()


def test_sharded_gram_allreduce_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, dh, ds, dm in res:
        assert dh < 1e-12 and ds < 1e-12 and dm == 0.0

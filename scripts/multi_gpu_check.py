"""Row-sharded multi-GPU parity check (run under torchrun, one rank per GPU):
the sharded engine (partial SYRK + NCCL all-reduce + replicated Cholesky, A'v all-reduce) must
reproduce the single-GPU engine on the same global problem, and a full sharded solve must
match the single-GPU solve."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import conicip_b200 as cb
from conicip_b200 import problems as P
from conicip_b200.dist import TorchReducer, init_engine_comm, shard_cones


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    worst = 0.0
    cases = [(P.mixed(n=200, mr=400, ncones=12, k=33, p=7, seed=21), -1),
             (P.mixed(n=300, mr=2000, ncones=0, k=3, p=0, seed=22), -1),
             # several 512-column outer panels: exercises the block-cyclic distributed Cholesky
             (P.mixed(n=1700, mr=2600, ncones=4, k=17, p=9, seed=23), 1),
             (P.mixed(n=1700, mr=2600, ncones=4, k=17, p=9, seed=23), 0)]
    for prob, dist_chol in cases:
        Q, A, G, cd = prob["Q"], prob["A"], prob["G"], prob["cone_dims"]
        n, m, p = len(prob["c"]), A.shape[0], G.shape[0]
        lo, hi, lcd = shard_cones(cd, world)[rank]
        rng = np.random.default_rng(3)
        v, s = np.zeros(m), np.zeros(m)
        off = 0
        for t, k in cd:
            if t == "R":
                v[off:off + k] = rng.uniform(0.5, 2, k); s[off:off + k] = rng.uniform(0.5, 2, k)
            else:
                for x in (v, s):
                    u = rng.standard_normal(k - 1); x[off] = np.linalg.norm(u) + 0.5; x[off + 1:off + k] = u
            off += k
        ry, rw, rv = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m)
        # single-GPU reference on every rank
        e1 = cb.Engine(Q, A, G if p else None, cd)
        lam1 = e1.factor_from_point(v, s)
        dy1, dw1, dv1 = e1.solve(ry, rw, rv)
        H1 = np.tril(e1.get_H())
        # sharded
        es = cb.Engine(Q, np.ascontiguousarray(A[lo:hi]), G if p else None, lcd, dist_chol=dist_chol)
        init_engine_comm(es)
        lam = es.factor_from_point(v[lo:hi], s[lo:hi])
        dy, dw, dv = es.solve(ry, rw, rv[lo:hi])
        errs = [rel(lam, lam1[lo:hi]), rel(np.tril(es.get_H()), H1), rel(dy, dy1), rel(dv, dv1[lo:hi])]
        if p:
            errs.append(rel(dw, dw1))
        errs.append(rel(es.mul_A(rv[lo:hi], trans=True), A.T @ rv))
        worst = max(worst, max(errs))
        if rank == 0:
            print(prob["name"], "n", n, "dist_chol", dist_chol, "shard rows", hi - lo, "errs", ["%.1e" % e for e in errs], flush=True)
        # full sharded solve vs single GPU
        sol1 = cb.conicIP(Q, prob["c"], A, prob["b"], cd, G if p else None, prob["d"] if p else None, optTol=1e-8)
        kk = lambda Q_, A_, G_, cd_: _gen(es)
        sols = cb.conicIP(Q, prob["c"], A[lo:hi], prob["b"][lo:hi], lcd, G if p else None, prob["d"] if p else None,
                          optTol=1e-8, kktsolver=kk, reducer=TorchReducer())
        e = [rel(sols.y, sol1.y), rel(sols.v, sol1.v[lo:hi])]
        worst = max(worst, max(e))
        assert sols.status == sol1.status == "Optimal" and abs(sols.Iter - sol1.Iter) <= 1
        if rank == 0:
            print("  solve", sols.status, sols.Iter, sol1.Iter, ["%.1e" % x for x in e], flush=True)
        e1.close(); es.close()
    t = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("WORST", float(t), flush=True)
    dist.destroy_process_group()
    assert float(t) < 1e-8


def _gen(eng):
    def gen(F, Finvt=None):
        st = eng.factor_resident() if isinstance(F, cb.DeviceBlock) else eng.factor(F)
        assert st == 0
        return lambda y, w, v: eng.solve(y, w, v)
    gen.engine = eng
    return gen


if __name__ == "__main__":
    main()

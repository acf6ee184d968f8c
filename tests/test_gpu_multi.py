"""Multi-GPU (row-sharded) parity: needs >= 2 GPUs on the box, skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_engine_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731",
                        os.path.join(ROOT, "scripts", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0


def _interior_point(cd, rng):
    m = sum(k for _, k in cd)
    v, s = np.zeros(m), np.zeros(m)
    off = 0
    for t, k in cd:
        if t == "R":
            v[off:off + k] = rng.uniform(0.5, 2, k); s[off:off + k] = rng.uniform(0.5, 2, k)
        else:
            for x in (v, s):
                u = rng.standard_normal(k - 1); x[off] = np.linalg.norm(u) + 0.5; x[off + 1:off + k] = u
        off += k
    return v, s


def test_single_process_multi_gpu_handle_matches_one_device():
    """cip_options.ngpus (SURVEY 8b Threading / 8e): ONE process, ONE handle, the rows of A sliced over two
    devices behind the C ABI.  Every entry point takes and returns global vectors and must reproduce the
    one-device engine: LEVEL 2 from a host Block and from a point, LEVEL 3, the cone kernels, the resident
    mat-vecs and the whole native solve (`conicIP(...; kktsolver = kktsolver_b200(ngpus = 2))` in Julia)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import conicip_b200 as cb
    from conicip_b200 import problems as P
    rel = lambda a, b: float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))
    for prob, csc in [(P.mixed(n=200, mr=400, ncones=12, k=33, p=7, seed=21), False),
                      (P.mixed(n=1700, mr=2600, ncones=4, k=17, p=9, seed=23), False),     # distributed Cholesky
                      (P.mixed(n=300, mr=2000, ncones=0, k=3, p=0, seed=22), True)]:
        Q, A, G, cd = prob["Q"], prob["A"], prob["G"], prob["cone_dims"]
        n, m, p = len(prob["c"]), A.shape[0], G.shape[0]
        if csc:
            import scipy.sparse as sp
            Qa, Aa = sp.csc_matrix(Q), sp.csc_matrix(A)
        else:
            Qa, Aa = Q, A
        e1 = cb.Engine(Qa, Aa, G if p else None, cd)
        e2 = cb.Engine(Qa, Aa, G if p else None, cd, ngpus=2)
        plan = cb.shard_plan(cd, 2)
        assert plan[0][1] - plan[0][0] > 0 and plan[1][1] - plan[1][0] > 0
        rng = np.random.default_rng(5)
        v, s = _interior_point(cd, rng)
        ry, rw, rv = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m)
        lam1, lam2 = e1.factor_from_point(v, s), e2.factor_from_point(v, s)
        assert rel(lam2, lam1) < 1e-13
        assert rel(np.tril(e2.get_H()), np.tril(e1.get_H())) < 1e-11
        (dy1, dw1, dv1), (dy2, dw2, dv2) = e1.solve(ry, rw, rv), e2.solve(ry, rw, rv)
        assert rel(dy2, dy1) < 1e-9 and rel(dv2, dv1) < 1e-9 and (p == 0 or rel(dw2, dw1) < 1e-9)
        # several right-hand sides at once through the sharded handle (cip_solve_multi: global columns in and out)
        RY, RW, RV = rng.standard_normal((n, 3)), rng.standard_normal((p, 3)), rng.standard_normal((m, 3))
        MY, MW, MV = e2.solve_multi(RY, RW if p else None, RV)
        for j in range(3):
            sy, sw, sv = e1.solve(RY[:, j], RW[:, j], RV[:, j])
            assert rel(MY[:, j], sy) < 1e-9 and rel(MV[:, j], sv) < 1e-9 and (p == 0 or rel(MW[:, j], sw) < 1e-9)
        x = rng.standard_normal(m)
        for op in (cb.OP_F, cb.OP_FT, cb.OP_FINVT, cb.OP_FINV):
            assert rel(e2.apply(op, x), e1.apply(op, x)) < 1e-13
        assert rel(e2.cone_prod(v, s), e1.cone_prod(v, s)) < 1e-13
        assert rel(e2.cone_div(x, v), e1.cone_div(x, v)) < 1e-12
        assert abs(e2.maxstep(v, x) - e1.maxstep(v, x)) <= 1e-13 * abs(e1.maxstep(v, x))
        assert e2.maxstep(x) == e1.maxstep(x)
        assert rel(e2.mul_A(ry), A @ ry) < 1e-13 and rel(e2.mul_A(rv, trans=True), A.T @ rv) < 1e-13
        assert rel(e2.mul_Q(ry), Q @ ry) < 1e-13
        if p:
            assert rel(e2.mul_G(ry), G @ ry) < 1e-13 and rel(e2.mul_G(rw, trans=True), G.T @ rw) < 1e-13
        k1, k2 = e1.get_scaling(), e2.get_scaling()
        assert np.array_equal(k1[0], k2[0]) and all(rel(b_, a_) < 1e-13 for a_, b_ in zip(k1[1:], k2[1:]) if np.linalg.norm(a_))
        # LEVEL 2 from the host Block the reference hands to solve3x3gen (flattened scaling read back from e1)
        F = cb.Block.from_flat(cd, *k1) if hasattr(cb.Block, "from_flat") else None
        if F is not None:
            assert e2.factor(F) == 0
            dy3, dw3, dv3 = e2.solve(ry, rw, rv)
            assert rel(dy3, dy1) < 1e-9 and rel(dv3, dv1) < 1e-9
        # whole solve
        kw = dict(optTol=1e-8)
        s1 = cb.conicIP_native(Qa, prob["c"], Aa, prob["b"], cd, G if p else None, prob["d"] if p else None, engine=e1, **kw)
        s2 = cb.conicIP_native(Qa, prob["c"], Aa, prob["b"], cd, G if p else None, prob["d"] if p else None, engine=e2, **kw)
        assert s1.status == s2.status == "Optimal" and abs(s1.Iter - s2.Iter) <= 1
        assert rel(s2.y, s1.y) < 1e-7 and rel(s2.v, s1.v) < 1e-6
        assert max(s2.prFeas, s2.duFeas, s2.muFeas) < 1e-8
        st = e2.stats()
        assert st["m"] == m and st["factors"] >= 2
        e1.close(); e2.close()

"""S (PSD) cone kernels behind the C ABI against the oracle: VecCongurance apply, nestod_sdc,
maxstep_sdc, xsdc!/dsdc!, the scaled panel and the reference's SDP test (runtests.jl:527-552).
The NT scaling of an S cone is unique only up to an orthogonal column transform of R (sign /
order of the singular vectors), so it is compared through invariants: F'F, lambda's spectrum,
and the defining property F v = F^-T s."""
import math

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def rand_pd(rng, k, cond=10.0):
    A = rng.standard_normal((k, k))
    Qm, _ = np.linalg.qr(A)
    return (Qm * np.geomspace(1.0, cond, k)) @ Qm.T


@pytest.fixture(scope="module", params=[2, 6, 17, 64, 65, 100])   # 65, 100: above the shared-memory orders
def scase(request):
    import conicip_b200 as cb
    k = request.param
    dim = k * (k + 1) // 2
    rng = np.random.default_rng(100 + k)
    n = 7
    # one R block, one S block, one Q block: S cones must coexist with the other kernels
    cones = [("R", 5), ("S", dim), ("Q", 4)]
    m = 5 + dim + 4
    A = rng.standard_normal((m, n))
    eng = cb.Engine(np.eye(n), A, None, cones)
    v, s = np.zeros(m), np.zeros(m)
    v[:5], s[:5] = rng.uniform(0.5, 2, 5), rng.uniform(0.5, 2, 5)
    v[5:5 + dim], s[5:5 + dim] = O.vecm(rand_pd(rng, k, 50.0)), O.vecm(rand_pd(rng, k, 50.0))
    for x in (v, s):
        u = rng.standard_normal(3)
        x[5 + dim] = np.linalg.norm(u) + 0.5
        x[6 + dim:] = u
    yield eng, k, dim, rng, v, s, A, cones
    eng.close()


def oracle_block(cones, v, s):
    bl, off = [], 0
    for t, kk in cones:
        xv, xs = v[off:off + kk], s[off:off + kk]
        bl.append(O.Diag(np.sqrt(xs / xv)) if t == "R" else O.nestod_soc(xv, xs) if t == "Q" else O.nestod_sdc(xv, xs))
        off += kk
    return O.Block(bl)


def test_nt_scaling_invariants(scase):
    import conicip_b200 as cb
    eng, k, dim, rng, v, s, A, cones = scase
    lam = eng.nt_scaling(v, s)
    Fo = oracle_block(cones, v, s)
    sl = slice(5, 5 + dim)
    # lambda = F v = F^-T s  (src/ConicIP.jl:735) and its spectrum equals the oracle's
    assert rel(eng.apply(cb.OP_FINVT, s), lam) < 1e-9
    assert rel(np.sort(np.linalg.eigvalsh(O.mat(lam[sl]))), np.sort(np.linalg.eigvalsh(O.mat(Fo.mul(v)[sl])))) < 1e-10
    Lm = O.mat(lam[sl])
    assert np.abs(Lm - np.diag(np.diag(Lm))).max() < 1e-10 * np.abs(Lm).max()        # R'ZR is diagonal
    # F'F is invariant: compare its action with the oracle's
    x = rng.standard_normal(len(v))
    assert rel(eng.apply(cb.OP_FT, eng.apply(cb.OP_F, x)), Fo.tmul(Fo.mul(x))) < 1e-9
    assert rel(eng.apply(cb.OP_FINV, eng.apply(cb.OP_FINVT, x)), Fo.inv().mul(Fo.inv_adjoint().mul(x))) < 1e-8
    # inverse consistency
    assert rel(eng.apply(cb.OP_FINV, eng.apply(cb.OP_F, x)), x) < 1e-9
    assert rel(eng.apply(cb.OP_FINVT, eng.apply(cb.OP_FT, x)), x) < 1e-9
    kind, fa, fb, fD, Rs = eng.get_scaling(with_R=True)
    assert kind.tolist() == [cb.BLK_DIAG, cb.BLK_VECCONG, cb.BLK_WOODBURY] and Rs[0].shape == (k, k)


def test_apply_with_user_supplied_veccongurance(scase):
    """cip_set_scaling / cip_factor with a host Block holding a VecCongurance (kind 2)."""
    import conicip_b200 as cb
    eng, k, dim, rng, v, s, A, cones = scase
    R = rng.standard_normal((k, k)) + 3 * np.eye(k)
    w = rng.standard_normal(4)
    F = cb.Block([cb.Diagonal(rng.uniform(1, 2, 5)), cb.VecCongurance(R), cb.SymWoodbury([-2.0, 2, 2, 2], w, 1.0)])
    Fo = O.Block([O.Diag(F[0].diag), O.VecCongurance(R), O.SymWoodbury([-2.0, 2, 2, 2], w, 1.0)])
    eng.set_scaling(F)
    x = rng.standard_normal(len(v))
    assert rel(eng.apply(cb.OP_F, x), Fo.mul(x)) < 1e-12            # runtests.jl:73  Z*x == Matrix(Z)*x
    assert rel(eng.apply(cb.OP_FT, x), Fo.tmul(x)) < 1e-12
    assert rel(eng.apply(cb.OP_FINV, x), Fo.inv().mul(x)) < 1e-9     # runtests.jl:75
    assert rel(eng.apply(cb.OP_FINVT, x), Fo.inv_adjoint().mul(x)) < 1e-9
    # LEVEL 2 + 3 with this F against the oracle's QR solver (the only built-in that is right for S)
    assert eng.factor(F) == 0
    n = A.shape[1]
    ry, rv = rng.standard_normal(n), rng.standard_normal(len(v))
    dy, dw, dv = eng.solve(ry, None, rv)
    oy, ow, ov = O.kktsolver_qr(np.eye(n), A, np.zeros((0, n)), cones)(Fo, Fo.inv_adjoint())(ry, np.zeros(0), rv)
    assert rel(dy, oy) < 1e-8 and rel(dv, ov) < 1e-8
    Atil = Fo.inv_adjoint().mul(A)
    eng.form_H()
    assert rel(np.tril(eng.get_H()), np.tril(np.eye(n) + Atil.T @ Atil)) < 1e-10


def test_solve_with_diagonal_blocks_on_every_cone(scase):
    """The initial point of conicIP factors with F = I: a Diagonal block on EVERY cone, S and Q cones included
    (src/ConicIP.jl:704-706).  The S rows then take the elementwise path of the fused inv(F'F) kernel."""
    import conicip_b200 as cb
    eng, k, dim, rng, v, s, A, cones = scase
    diags = [rng.uniform(0.5, 2.0, kk) for _, kk in cones]
    F = cb.Block([cb.Diagonal(dg) for dg in diags])
    Fo = O.Block([O.Diag(dg) for dg in diags])
    assert eng.factor(F) == 0
    n = A.shape[1]
    ry, rv = rng.standard_normal(n), rng.standard_normal(len(v))
    dy, dw, dv = eng.solve(ry, None, rv)
    oy, ow, ov = O.kktsolver_qr(np.eye(n), A, np.zeros((0, n)), cones)(Fo, Fo.inv_adjoint())(ry, np.zeros(0), rv)
    assert rel(dy, oy) < 1e-9 and rel(dv, ov) < 1e-9
    # the 3x3 system itself: Q dy - A' dv = ry ; A dy + F'F dv = rv   (src/kktsolvers.jl:1-12)
    d2 = np.concatenate(diags) ** 2
    assert rel(dy - A.T @ dv, ry) < 1e-9 and rel(A @ dy + d2 * dv, rv) < 1e-9


def test_maxstep_sdc(scase):
    eng, k, dim, rng, v, s, A, cones = scase
    d = rng.standard_normal(len(v))
    d[5:5 + dim] = O.vecm((lambda B: B + B.T)(rng.standard_normal((k, k))))

    def want(x, dd):
        mn, off = math.inf, 0
        for t, kk in cones:
            xi, di = x[off:off + kk], None if dd is None else dd[off:off + kk]
            mn = min(mn, O.maxstep_rp(xi, di) if t == "R" else O.maxstep_soc(xi, di) if t == "Q" else O.maxstep_sdc(xi, di))
            off += kk
        return mn
    # isolate the S cone: make the other cones non-binding
    big = d.copy(); big[:5] = -1.0; big[5 + dim:] = 0.0; big[5 + dim] = -1.0
    a = eng.maxstep(v, big)
    assert a == pytest.approx(O.maxstep_sdc(v[5:5 + dim], big[5:5 + dim]), rel=1e-9)
    Xb = O.mat(v[5:5 + dim]) - a * O.mat(big[5:5 + dim])
    assert abs(np.linalg.eigvalsh(Xb).min()) < 1e-8 * np.abs(Xb).max()               # lands on the boundary
    assert eng.maxstep(v, d) == pytest.approx(want(v, d), rel=1e-9)
    assert eng.maxstep(d, None) == pytest.approx(want(d, None), rel=1e-9)
    assert eng.maxstep(v, None) == 0.0
    # runtests.jl:79-82: X = -I (not PD) -> Inf
    x = v.copy(); x[5:5 + dim] = O.vecm(-np.eye(k))
    dd = big.copy(); dd[5:5 + dim] = O.vecm(np.eye(k))
    assert math.isinf(eng.maxstep(x, dd))


def test_prod_div_sdc(scase):
    eng, k, dim, rng, v, s, A, cones = scase
    x = rng.standard_normal(len(v))
    sl = slice(5, 5 + dim)
    assert rel(eng.cone_prod(x, s)[sl], O.xsdc(x[sl], s[sl])) < 1e-12
    assert rel(eng.cone_div(x, v)[sl], O.dsdc(x[sl], v[sl])) < 1e-9
    assert rel(eng.cone_div(eng.cone_prod(v, x), v), x) < 1e-8


def test_reference_sdp_projection():
    """runtests.jl:527-552 -- projection onto the PSD cone, S cone of order 6; recorded Iter 6."""
    import conicip_b200 as cb
    c = O.vecm(np.diag([1.0, 1, 1, -1, -1, -1]))
    s = cb.conicIP(np.eye(21), c, np.eye(21), np.zeros(21), [("S", 21)], optTol=1e-7)
    so = O.conicIP(np.eye(21), c, np.eye(21), np.zeros(21), [("S", 21)], optTol=1e-7)
    assert s.status == so.status == "Optimal"
    assert abs(s.Iter - 6) <= 1 and abs(s.Iter - so.Iter) <= 1
    assert np.abs(O.mat(s.y) - np.diag([1.0, 1, 1, 0, 0, 0])).max() < 1e-3
    assert rel(s.y, so.y) < 1e-6


def test_mixed_r_q_s_problem_vs_oracle():
    """R + Q + S cones and an equality block in one solve (the C5 cone mix at test size)."""
    import conicip_b200 as cb
    rng = np.random.default_rng(42)
    n, k = 30, 5
    dim = k * (k + 1) // 2
    cones = [("R", 12), ("S", dim), ("Q", 6)]
    m = 12 + dim + 6
    A = rng.standard_normal((m, n)) / math.sqrt(n)
    y0 = rng.standard_normal(n)
    s0 = np.zeros(m)
    s0[:12] = rng.uniform(0.1, 1.1, 12)
    s0[12:12 + dim] = O.vecm(rand_pd(rng, k, 5.0))
    u = 0.1 * rng.standard_normal(5)
    s0[12 + dim] = 1 + np.linalg.norm(u); s0[13 + dim:] = u
    b = A @ y0 - s0
    G = rng.standard_normal((3, n)) / math.sqrt(n)
    d = G @ y0
    Q = np.eye(n) * 1.5
    c = rng.standard_normal(n)
    s = cb.conicIP(Q, c, A, b, cones, G, d, optTol=1e-8)
    so = O.conicIP(Q, c, A, b, cones, G, d, optTol=1e-8, kktsolver=O.kktsolver_qr)
    assert s.status == so.status == "Optimal" and abs(s.Iter - so.Iter) <= 1
    assert rel(s.y, so.y) < 1e-6 and rel(s.v, so.v) < 1e-6 and rel(s.w, so.w) < 1e-6
    assert max(s.prFeas, s.duFeas, s.muFeas) < 1e-8


def test_s_cone_order_limit_is_loud():
    import conicip_b200 as cb
    k = 513
    with pytest.raises(cb.CipError):
        cb.Engine(np.eye(2), np.zeros((k * (k + 1) // 2, 2)), None, [("S", k * (k + 1) // 2)])


def test_large_order_s_cone_solve_vs_oracle():
    """An S cone of order 80 (above the 64 that fit shared memory: the global-workspace path of sdp.cu) through a
    whole solve: projection of a symmetric matrix onto the PSD cone, as the reference's SDP test does at order 6
    (runtests.jl:527-552), plus R rows and an equality block."""
    import conicip_b200 as cb
    rng = np.random.default_rng(80)
    k = 80
    dim = k * (k + 1) // 2
    B = rng.standard_normal((k, k))
    target = O.vecm((B + B.T) / 2)
    n = dim
    cones = [("S", dim), ("R", 10)]
    A = np.vstack([np.eye(n), rng.standard_normal((10, n)) / math.sqrt(n)])
    b = np.concatenate([np.zeros(dim), -np.ones(10)])
    G = rng.standard_normal((4, n)) / math.sqrt(n)
    d = G @ O.vecm(np.eye(k))
    s = cb.conicIP_native(np.eye(n), target, A, b, cones, G, d, optTol=1e-8)
    so = O.conicIP(np.eye(n), target, A, b, cones, G, d, optTol=1e-8, kktsolver=O.kktsolver_qr)
    assert s.status == so.status == "Optimal" and abs(s.Iter - so.Iter) <= 1, (s.status, s.Iter, so.Iter)
    assert rel(s.y, so.y) < 1e-6 and rel(s.v, so.v) < 1e-5
    assert max(s.prFeas, s.duFeas, s.muFeas) < 1e-8
    assert np.linalg.eigvalsh(O.mat(s.y)).min() > -1e-7

"""GPU parity tests proper: every kernel family behind the C ABI against the oracle on the
same seeded inputs.  FP64 throughout; tolerances are relative 2-norm errors written per test
(north_star: final y,v,w within 1e-6; kernels themselves agree to ~1e-13)."""
import math

import numpy as np
import pytest

import oracle as O
from conicip_b200 import problems as P

pytestmark = pytest.mark.gpu
KTOL = 1e-12          # single kernel vs oracle (different summation order only)


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def interior_point(cone_dims, rng):
    m = sum(k for _, k in cone_dims)
    v, s, off = np.zeros(m), np.zeros(m), 0
    for t, k in cone_dims:
        if t == "R":
            v[off:off + k] = rng.uniform(0.5, 2, k)
            s[off:off + k] = rng.uniform(0.5, 2, k)
        else:
            for x in (v, s):
                u = rng.standard_normal(k - 1)
                x[off] = np.linalg.norm(u) + rng.uniform(0.1, 1)
                x[off + 1:off + k] = u
        off += k
    return v, s


def oracle_F(cone_dims, v, s):
    bl, off = [], 0
    for t, k in cone_dims:
        bl.append(O.Diag(np.sqrt(s[off:off + k] / v[off:off + k])) if t == "R"
                  else O.nestod_soc(v[off:off + k], s[off:off + k]))
        off += k
    return O.Block(bl)


def per_cone(cone_dims, fr, fq, *xs):
    m = sum(k for _, k in cone_dims)
    o, off = np.zeros(m), 0
    for t, k in cone_dims:
        o[off:off + k] = (fr if t == "R" else fq)(*[x[off:off + k] for x in xs])
        off += k
    return o


def o_maxstep(cone_dims, x, d):
    mn, off = math.inf, 0
    for t, k in cone_dims:
        xi = x[off:off + k]
        di = None if d is None else d[off:off + k]
        mn = min(mn, O.maxstep_rp(xi, di) if t == "R" else O.maxstep_soc(xi, di))
        off += k
    return mn


# shapes chosen to hit: ragged n/m (not multiples of 128/32/4), p = 0 and p > 0, p > one tile, one
# R row, many small Q cones, one large Q cone (CTA-per-cone path), cones straddling quads, m >> n (split-K)
CASES = {
    "mixed": dict(n=96, mr=80, ncones=6, k=9, p=5, seed=7),
    "ragged": dict(n=131, mr=77, ncones=3, k=33, p=0, seed=1),
    "r_only_p_tiles": dict(n=300, mr=1000, ncones=0, k=3, p=130, seed=4),
    "q_only_small": dict(n=40, mr=0, ncones=50, k=3, p=2, seed=5),
    "one_big_q": dict(n=64, mr=1, ncones=1, k=2500, p=0, seed=6),
    "tiny": dict(n=1, mr=1, ncones=0, k=2, p=0, seed=8),
    "n_multi_tile": dict(n=520, mr=640, ncones=8, k=17, p=1, seed=9),
    "tall_skinny": dict(n=200, mr=20000, ncones=40, k=33, p=3, seed=10),   # 3 C tiles, 667 k tiles: split-K SYRK
    # equality block of five row tiles on eleven column panels: the two-level Schur substitution (p > 384), with a
    # partial last outer panel
    "big_equality_block": dict(n=1300, mr=1500, ncones=3, k=9, p=520, seed=14),
}


# heterogeneous cone lists: the fused R+Q kernels pick the lanes-per-cone class (8 / 32 / 256) from the LARGEST Q cone
# of the problem, so small cones also run in the wider classes; dimensions sit on the class boundaries (64 | 65,
# 1024 | 1025), R blocks are interleaved with Q cones, and Q cones of dimension 2 (a single tail entry) are included
CASES["q_class8_edges"] = dict(cones=[("Q", 57), ("R", 1), ("Q", 64), ("Q", 2), ("Q", 8), ("Q", 9), ("R", 3), ("Q", 63)], n=70, p=2, seed=11)
CASES["q_class32_mixed"] = dict(cones=[("R", 7), ("Q", 2), ("Q", 64), ("Q", 65), ("Q", 3), ("R", 5), ("Q", 200), ("Q", 1024)], n=90, p=0, seed=12)
CASES["q_class256_mixed"] = dict(cones=[("Q", 5), ("Q", 1025), ("R", 3), ("Q", 33), ("Q", 1024), ("Q", 2)], n=50, p=3, seed=13)


def hetero_problem(cones, n, p, seed):
    """Like problems.mixed, for an arbitrary list of R / Q cones."""
    rng = np.random.default_rng(seed)
    m = sum(k for _, k in cones)
    A = rng.standard_normal((m, n)) / np.sqrt(n)
    y0 = rng.standard_normal(n)
    s0, off = np.zeros(m), 0
    for t, k in cones:
        if t == "R":
            s0[off:off + k] = rng.uniform(0.1, 1.1, k)
        else:
            u = 0.1 * rng.standard_normal(k - 1)
            s0[off] = 1.0 + np.linalg.norm(u)
            s0[off + 1:off + k] = u
        off += k
    G = rng.standard_normal((p, n)) / np.sqrt(n)
    return dict(name="hetero", Q=np.eye(n) * 1.5, c=rng.standard_normal(n), A=A, b=A @ y0 - s0, cone_dims=list(cones),
                G=G, d=G @ y0, optTol=1e-8)


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    import conicip_b200 as cb
    kw = CASES[request.param]
    prob = hetero_problem(**kw) if "cones" in kw else P.mixed(**kw)
    p = prob["G"].shape[0]
    eng = cb.Engine(prob["Q"], prob["A"], prob["G"] if p else None, prob["cone_dims"])
    rng = np.random.default_rng(123)
    v, s = interior_point(prob["cone_dims"], rng)
    yield prob, eng, rng, v, s
    eng.close()


def test_nt_scaling_and_apply(case):
    import conicip_b200 as cb
    prob, eng, rng, v, s = case
    cd = prob["cone_dims"]
    lam = eng.nt_scaling(v, s)
    Fo = oracle_F(cd, v, s)
    assert rel(lam, Fo.mul(v)) < KTOL
    assert rel(lam, Fo.inv_adjoint().mul(s)) < 1e-9            # lambda = F v = F^-T s
    x = rng.standard_normal(len(v))
    assert rel(eng.apply(cb.OP_F, x), Fo.mul(x)) < KTOL
    assert rel(eng.apply(cb.OP_FT, x), Fo.tmul(x)) < KTOL
    assert rel(eng.apply(cb.OP_FINVT, x), Fo.inv_adjoint().mul(x)) < 1e-11
    assert rel(eng.apply(cb.OP_FINV, x), Fo.inv().mul(x)) < 1e-11
    kind, fa, fb, fD = eng.get_scaling()
    off = 0
    for i, (t, k) in enumerate(cd):
        blk = Fo[i]
        if t == "R":
            assert kind[i] == cb.BLK_DIAG and rel(fa[off:off + k], blk.diag) < KTOL
        else:
            assert kind[i] == cb.BLK_WOODBURY and fD[i] == 1.0
            assert rel(fa[off:off + k], blk.Adiag) < KTOL and rel(fb[off:off + k], blk.B) < KTOL
        off += k


def test_maxstep(case):
    prob, eng, rng, v, s = case
    cd = prob["cone_dims"]
    d = rng.standard_normal(len(v))
    for x, dd in ((v, d), (s, -d), (v, None), (d, None), (v, -np.abs(v))):
        got, want = eng.maxstep(x, dd), o_maxstep(cd, x, dd)
        if math.isinf(want):
            assert math.isinf(got)
        else:
            assert got == pytest.approx(want, rel=1e-12, abs=1e-300)
    # the reference's DTB call form: maxstep(z, dz/(1-DTB))  (src/ConicIP.jl:927)
    assert eng.maxstep(v, d, 0.99) == pytest.approx(o_maxstep(cd, v, d / 0.99), rel=1e-12)


def test_cone_prod_div(case):
    prob, eng, rng, v, s = case
    cd = prob["cone_dims"]
    x = rng.standard_normal(len(v))
    assert rel(eng.cone_prod(x, s), per_cone(cd, O.xrp, O.xsoc, x, s)) < KTOL
    assert rel(eng.cone_div(x, v), per_cone(cd, O.drp, O.dsoc, x, v)) < 1e-11
    assert rel(eng.cone_div(eng.cone_prod(v, x), v), x) < 1e-10      # division inverts the product


def test_resident_matvecs(case):
    prob, eng, rng, v, s = case
    Q, A, G = prob["Q"], prob["A"], prob["G"]
    n, m, p = len(prob["c"]), A.shape[0], G.shape[0]
    x, u = rng.standard_normal(n), rng.standard_normal(m)
    assert rel(eng.mul_A(x), A @ x) < KTOL
    assert rel(eng.mul_A(u, trans=True), A.T @ u) < KTOL
    assert rel(eng.mul_Q(x), Q @ x) < KTOL
    if p:
        w = rng.standard_normal(p)
        assert rel(eng.mul_G(x), G @ x) < KTOL
        assert rel(eng.mul_G(w, trans=True), G.T @ w) < KTOL
    else:
        assert np.all(eng.mul_G(np.zeros(0), trans=True) == 0)


def test_form_H_cholesky_and_solve(case):
    """K1 (scaled SYRK), K2 (Cholesky + Schur), K3/K4 (solves) against the oracle."""
    prob, eng, rng, v, s = case
    Q, A, G, cd = prob["Q"], prob["A"], prob["G"], prob["cone_dims"]
    n, m, p = len(prob["c"]), A.shape[0], G.shape[0]
    eng.nt_scaling(v, s)
    Fo = oracle_F(cd, v, s)
    Fi = Fo.inv_adjoint()
    At = Fi.mul(A)
    Href = Q + At.T @ At                                       # src/kktsolvers.jl:33-34
    if p:
        Href = Href + G.T @ G                                  # default augmentation rho = 1 (opts.aug_rho)
    eng.form_H()
    assert rel(np.tril(eng.get_H()), np.tril(Href)) < KTOL
    assert eng.factor_H() == 0
    assert rel(np.tril(eng.get_H()), np.linalg.cholesky(Href)) < 1e-11
    ry, rw, rv = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m)
    dy, dw, dv = eng.solve(ry, rw, rv)
    for ks in (O.kktsolver_chol, O.pivot(O.kktsolver_2x2), O.kktsolver_qr):
        oy, ow, ov = ks(Q, A, G, cd)(Fo, Fi)(ry, rw, rv)
        assert rel(dy, oy) < 1e-9 and rel(dv, ov) < 1e-9
        if p:
            assert rel(dw, ow) < 1e-8
    # the defining 3x3 system (docs/src/guides/kkt_solvers.md:99-103)
    FtF = Fo.dense().T @ Fo.dense()
    scale = np.linalg.norm(np.concatenate([ry, rw, rv]))
    assert np.linalg.norm(Q @ dy + (G.T @ dw if p else 0) - A.T @ dv - ry) < 1e-10 * scale * (1 + np.linalg.norm(Href))
    assert np.linalg.norm(A @ dy + FtF @ dv - rv) < 1e-10 * scale * (1 + np.linalg.norm(FtF))
    if p:
        assert np.linalg.norm(G @ dy - rw) < 1e-10 * scale


def test_solve_multi_equals_column_by_column(case):
    """cip_solve_multi (north_star: the right-hand sides "together"): every column equals cip_solve on that column --
    bit for bit here, where the shared passes over A use the same split of the contraction -- for an odd count
    (pairs + a single), with host arrays and with device-resident tensors."""
    import torch
    prob, eng, rng, v, s = case
    n, m, p = len(prob["c"]), prob["A"].shape[0], prob["G"].shape[0]
    eng.nt_scaling(v, s)
    eng.factor_resident()
    k = 5
    RY, RW, RV = rng.standard_normal((n, k)), rng.standard_normal((p, k)), rng.standard_normal((m, k))
    DY, DW, DV = eng.solve_multi(RY, RW if p else None, RV)
    assert DY.shape == (n, k) and DV.shape == (m, k)
    for j in range(k):
        dy, dw, dv = eng.solve(RY[:, j], RW[:, j], RV[:, j])
        assert rel(DY[:, j], dy) < 1e-13 and rel(DV[:, j], dv) < 1e-13
        if p:
            assert rel(DW[:, j], dw) < 1e-12
    tY, tW, tV = (torch.as_tensor(X).cuda() for X in (RY, RW, RV))
    gY, gW, gV = eng.solve_multi(tY, tW if p else None, tV)
    assert rel(gY.cpu().numpy(), DY) < 1e-13 and rel(gV.cpu().numpy(), DV) < 1e-13
    # one column and zero columns are legal
    Y1, W1, V1 = eng.solve_multi(RY[:, :1], RW[:, :1] if p else None, RV[:, :1])
    assert rel(Y1[:, 0], DY[:, 0]) < 1e-13
    Y0, _, V0 = eng.solve_multi(RY[:, :0], RW[:, :0] if p else None, RV[:, :0])
    assert Y0.shape == (n, 0) and V0.shape == (m, 0)


def test_factor_with_host_block_equals_resident(case):
    """cip_factor(flattened Block) == nt_scaling on the device + factor; and the initial
    all-Diagonal(ones) call of src/ConicIP.jl:704 works for Q slots too."""
    import conicip_b200 as cb
    prob, eng, rng, v, s = case
    cd = prob["cone_dims"]
    n, m, p = len(prob["c"]), prob["A"].shape[0], prob["G"].shape[0]
    ry, rw, rv = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m)
    eng.nt_scaling(v, s)
    eng.factor_resident()
    a = eng.solve(ry, rw, rv)
    F_host = cb.DeviceBlock(eng).to_host()
    assert eng.factor(F_host) == 0
    b = eng.solve(ry, rw, rv)
    for x, y in zip(a, b):
        assert rel(x, y) < 1e-12 if len(y) else True
    I0 = cb.Block([cb.Diagonal(np.ones(k)) for _, k in cd])
    assert eng.factor(I0) == 0
    dy, dw, dv = eng.solve(ry, rw, rv)
    Io = O.Block([O.Diag(np.ones(k)) for _, k in cd])
    oy, ow, ov = O.pivot(O.kktsolver_2x2)(prob["Q"], prob["A"], prob["G"], cd)(Io, Io)(ry, rw, rv)
    assert rel(dy, oy) < 1e-9 and rel(dv, ov) < 1e-9


def test_device_pointer_api_matches_host_pointer_api(case):
    import torch
    prob, eng, rng, v, s = case
    n, m, p = len(prob["c"]), prob["A"].shape[0], prob["G"].shape[0]
    T = lambda x: torch.as_tensor(x).cuda()
    lam_h = eng.factor_from_point(v, s)
    ry, rw, rv = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m)
    h = eng.solve(ry, rw, rv)
    lam_d = eng.factor_from_point(T(v), T(s))
    d = eng.solve(T(ry), T(rw) if p else None, T(rv))
    assert torch.is_tensor(lam_d) and lam_d.is_cuda
    assert np.array_equal(lam_d.cpu().numpy(), lam_h)          # bit-identical: same kernels, same order
    for x, y in zip(h, d):
        assert np.array_equal(x, y.cpu().numpy())
    assert not np.shares_memory(h[0], ry)                      # outputs are fresh arrays (src/ConicIP.jl:920)


def test_solve_is_deterministic(case):
    prob, eng, rng, v, s = case
    n, m, p = len(prob["c"]), prob["A"].shape[0], prob["G"].shape[0]
    ry, rw, rv = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(m)
    eng.factor_from_point(v, s)
    a = eng.solve(ry, rw, rv)
    eng.factor_from_point(v, s)
    b = eng.solve(ry, rw, rv)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_cone_kernel_fixture_on_device():
    """tests/golden/oracle_small.json cone-kernel known answers through the C ABI."""
    import json
    import os
    import conicip_b200 as cb
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_small.json")))["cone_kernels"]
    z, s, d = (np.array(g[k]) for k in ("z", "s", "d"))
    eng = cb.Engine(np.eye(2), np.ones((5, 2)), None, [("Q", 5)])
    lam = eng.nt_scaling(z, s)
    kind, fa, fb, fD = eng.get_scaling()
    assert rel(fb, g["nestod_soc_w"]) < 1e-14 and rel(fa, g["nestod_soc_diag"]) < 1e-14
    assert rel(lam, g["lam"]) < 1e-14
    assert eng.maxstep(z, d) == pytest.approx(g["maxstep_soc"], rel=1e-13)
    assert eng.maxstep(d, None) == pytest.approx(g["maxstep_soc_nothing"], rel=1e-13)
    assert rel(eng.cone_prod(z, s), g["xsoc"]) < 1e-14
    assert rel(eng.cone_div(z, s), g["dsoc"]) < 1e-13
    eng.close()
    eng = cb.Engine(np.eye(2), np.ones((5, 2)), None, [("R", 5)])
    assert eng.maxstep(z, d) == pytest.approx(g["maxstep_rp"], rel=1e-15)
    assert eng.maxstep(d, None) == pytest.approx(g["maxstep_rp_nothing"], rel=1e-15)
    eng.close()


def test_error_behaviour():
    import conicip_b200 as cb
    with pytest.raises(ValueError):
        cb.Engine(np.eye(3), np.ones((4, 3)), None, [("R", 5)])            # cones do not cover rows
    with pytest.raises(ValueError):
        cb.Engine(np.eye(3), np.ones((4, 2)), None, [("R", 4)])            # A/Q mismatch (runtests.jl:507-523)
    # H not positive definite -> positive status with the failing column (SURVEY 8b "Errors")
    eng = cb.Engine(-np.eye(3), np.zeros((2, 3)), None, [("R", 2)])
    st = eng.factor(cb.Block([cb.Diagonal(np.ones(2))]))
    assert st == 1 and "pivot" in cb._lib.last_error()
    eng.close()


def test_singular_H_with_equalities_like_kktsolver_qr():
    """H singular on range(G') (LP with a free variable): kktsolver_qr handles it (src/kktsolvers.jl:35),
    a plain Cholesky of H cannot.  The default augmentation H + rho G'G gives the same solution."""
    import conicip_b200 as cb
    # variables (x1, x2, t): minimise t  s.t.  t - x1 - x2 = 0,  x1 >= 1, x2 >= 2   (t is free: H_tt = 0)
    Q = np.zeros((3, 3))
    c = np.array([0.0, 0.0, -1.0])                       # minimise -c'y = t
    A = np.array([[1.0, 0, 0], [0, 1.0, 0]])
    b = np.array([1.0, 2.0])
    G = np.array([[-1.0, -1.0, 1.0]])
    d = np.array([0.0])
    so = O.conicIP(Q, c, A, b, [("R", 2)], G, d, optTol=1e-8, kktsolver=O.kktsolver_qr)
    s = cb.conicIP(Q, c, A, b, [("R", 2)], G, d, optTol=1e-8)
    assert s.status == so.status == "Optimal" and abs(s.Iter - so.Iter) <= 1
    assert np.allclose(s.y, [1.0, 2.0, 3.0], atol=1e-6) and rel(s.y, so.y) < 1e-6 and rel(s.v, so.v) < 1e-6
    sn = cb.conicIP_native(Q, c, A, b, [("R", 2)], G, d, optTol=1e-8)
    assert sn.status == "Optimal" and rel(sn.y, so.y) < 1e-6
    # without the augmentation the factorisation reports the failing pivot (status > 0)
    eng = cb.Engine(Q, A, G, [("R", 2)], aug_rho=0.0)
    assert eng.factor(cb.Block([cb.Diagonal(np.ones(2))])) == 3
    eng.close()


def test_regularisation_option():
    import conicip_b200 as cb
    eng = cb.Engine(np.zeros((3, 3)), np.zeros((2, 3)), None, [("R", 2)], reg_delta=0.25)
    assert eng.factor(cb.Block([cb.Diagonal(np.ones(2))])) == 0
    assert np.allclose(np.diag(eng.get_H()), 0.5)                            # chol(0.25 I)
    eng.close()


def test_folded_scaling_matches_materialised_panel():
    """opts.fold_scaling: for K = R^m the SYRK applies W^-2 to its operand fragments and the handle keeps no
    Atil = F^-T A (src/kktsolvers.jl:33).  Same H (up to the order in which the two factors of w^2 meet the
    products), same solve, about half the resident bytes."""
    import conicip_b200 as cb
    rng = np.random.default_rng(77)
    for n, m in [(300, 2000), (128, 4096), (1000, 1000)]:
        A = rng.standard_normal((m, n)) / np.sqrt(n)
        Q = np.diag(rng.uniform(1, 2, n))
        v, s = rng.uniform(1e-3, 1e3, m), rng.uniform(1e-3, 1e3, m)
        ry, rv = rng.standard_normal(n), rng.standard_normal(m)
        e0 = cb.Engine(Q, A, None, [("R", m)], fold_scaling=2)
        e1 = cb.Engine(Q, A, None, [("R", m)], fold_scaling=1)
        assert e1.stats()["device_bytes"] < e0.stats()["device_bytes"] - 8 * m * n // 2
        for e in (e0, e1):
            e.nt_scaling(v, s)
            e.form_H()
        H0, H1 = np.tril(e0.get_H()), np.tril(e1.get_H())
        assert np.linalg.norm(H1 - H0) <= 1e-13 * np.linalg.norm(H0)
        want = Q + (A.T * (v / s)) @ A
        assert np.linalg.norm(H1 - np.tril(want)) <= 1e-12 * np.linalg.norm(want)
        assert e0.factor_H() == 0 and e1.factor_H() == 0
        (dy0, _, dv0), (dy1, _, dv1) = e0.solve(ry, None, rv), e1.solve(ry, None, rv)
        assert np.linalg.norm(dy1 - dy0) <= 1e-9 * np.linalg.norm(dy0)
        assert np.linalg.norm(dv1 - dv0) <= 1e-9 * np.linalg.norm(dv0)
        # the host Block path (cip_factor with DIAG kinds) folds as well
        F = cb.Block([cb.Diagonal(np.sqrt(s / v))])
        assert e1.factor(F) == 0
        dy2, _, dv2 = e1.solve(ry, None, rv)
        assert np.linalg.norm(dy2 - dy0) <= 1e-9 * np.linalg.norm(dy0)
        e0.close(); e1.close()


@pytest.mark.parametrize("name", ["q_class8_edges", "q_class32_mixed", "q_class256_mixed"])
def test_heterogeneous_q_cones_full_solve(name):
    """Whole solves (native loop and host driver) on the heterogeneous cone lists above against the oracle:
    north_star's gate -- iterations within +-1, y, v, w within 1e-6, residuals below 1e-8."""
    import conicip_b200 as cb
    kw = CASES[name]
    prob = hetero_problem(**kw)
    Q, c, A, b, cd, G, d = (prob[k] for k in ("Q", "c", "A", "b", "cone_dims", "G", "d"))
    p = G.shape[0]
    so = O.conicIP(Q, c, A, b, cd, G if p else None, d if p else None, optTol=1e-8, kktsolver=O.pivot(O.kktsolver_2x2))
    for solve in (cb.conicIP_native, cb.conicIP):
        s = solve(Q, c, A, b, cd, G if p else None, d if p else None, optTol=1e-8)
        assert s.status == so.status == "Optimal" and abs(s.Iter - so.Iter) <= 1, (solve.__name__, s.status, s.Iter, so.Iter)
        assert rel(s.y, so.y) < 1e-6 and rel(s.v, so.v) < 1e-6 and (p == 0 or rel(s.w, so.w) < 1e-6)
        assert max(s.prFeas, s.duFeas, s.muFeas) < 1e-8

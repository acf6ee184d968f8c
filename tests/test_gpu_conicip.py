"""End-to-end parity of the B200 path behind the reference's interface
(`conicIP(...; kktsolver=kktsolver_b200)`) against the oracle, the committed golden fixture and
the reference's own recorded goldens.  Gates from BASELINE.json: iteration count within +-1,
final y,v,w within 1e-6 relative, prFeas/duFeas/muFeas below the same tolerance."""
import json
import os

import numpy as np
import pytest

import oracle as O
from conicip_b200 import problems as P

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def run_b200(prob, **kw):
    import conicip_b200 as cb
    p = prob["G"].shape[0]
    opts = dict(optTol=prob.get("optTol", 1e-7))
    opts.update(kw)
    return cb.conicIP(prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"],
                      prob["G"] if p else None, prob["d"] if p else None, **opts)


def run_oracle(prob, solver=None, **kw):
    opts = dict(optTol=prob.get("optTol", 1e-7))
    opts.update(kw)
    return O.conicIP(prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"], prob["G"], prob["d"],
                     kktsolver=solver or O.pivot(O.kktsolver_2x2), **opts)


def assert_parity(s, so, tol):
    assert s.status == so.status
    assert abs(s.Iter - so.Iter) <= 1
    assert rel(s.y, so.y) < RTOL and rel(s.v, so.v) < RTOL
    if len(so.w):
        assert rel(s.w, so.w) < RTOL
    if s.status == "Optimal":
        assert max(s.prFeas, s.duFeas, s.muFeas) < tol


# reference goldens: test/runtests.jl:157-162, :197-202, :235-240
@pytest.mark.parametrize("gen,it,mu", [(P.sphere, 5, 2.866608128093695e-7), (P.combined, 10, 4.663886012743681e-7)])
def test_reference_mu_goldens(gen, it, mu):
    s = run_b200(gen(), DTB=0.01, maxRefinementSteps=3)
    assert s.status == "Optimal" and abs(s.Iter - it) <= 1
    mu_at = dict((t[0], t[1]) for t in s.trace)[it]
    assert abs(mu_at - mu) <= 1e-7 * mu


def test_reference_simplex_golden():
    s = run_b200(P.simplex(), optTol=1e-8)
    assert s.status == "Optimal" and s.Iter == 11
    assert abs(s.Mu - 2.7686402945528533e-9) <= 1e-7 * 2.7686402945528533e-9
    y = np.zeros(10); y[9] = 1
    assert np.linalg.norm(s.y - y) < 1e-3


@pytest.mark.parametrize("name", ["sphere", "combined", "simplex", "soc_direct", "mixed"])
def test_committed_golden_fixture(name):
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_small.json")))["solves"][name]
    s = run_b200(getattr(P, name)(), optTol=g["optTol"])
    assert s.status == g["status"] and abs(s.Iter - g["Iter"]) <= 1
    assert rel(s.y, g["y"]) < RTOL and rel(s.v, g["v"]) < RTOL
    if len(g["w"]):
        assert rel(s.w, g["w"]) < RTOL
    mu = [t[1] for t in s.trace]
    k = min(len(mu), len(g["mu_trace"]))
    assert np.allclose(mu[:k], g["mu_trace"][:k], rtol=1e-5)        # same trajectory, not just same end point


def test_box_qp_custom_kktsolver_problem():
    """runtests.jl:90-131 at its full size n=1000 (the reference's own plugin-boundary demo)."""
    prob = P.box_qp(1000)
    s = run_b200(prob, DTB=0.01, maxRefinementSteps=3)
    so = run_oracle(prob, DTB=0.01, maxRefinementSteps=3)
    assert_parity(s, so, 1e-7)
    c = np.arange(1.0, 1001)
    assert np.linalg.norm(s.y - np.clip(c, -1, 1)) / 1000 < 1e-3


def test_config1_readme_qp():
    """C1: README nonnegative QP n=1000 (BASELINE configs[0]) -- full solve vs the oracle."""
    prob = P.config1()
    s = run_b200(prob)
    so = run_oracle(prob)
    assert_parity(s, so, 1e-8)


def test_config3_socp_reduced():
    """C3 shape at a size the oracle finishes in seconds: 64 Q cones of dim 33, p = 32."""
    prob = P.config3(n=512, ncones=64, k=33, p=32, seed=3)
    s = run_b200(prob)
    so = run_oracle(prob, O.kktsolver_chol)
    assert_parity(s, so, 1e-8)


def test_statuses_abandoned_infeasible_unbounded():
    assert run_b200(P.simplex(), maxIters=2).status == "Abandoned"         # runtests.jl:246-269
    assert run_b200(P.infeasible()).status == "Infeasible"                  # runtests.jl:441-460
    s = run_b200(P.unbounded())                                             # runtests.jl:487-505
    assert s.status == "Unbounded" and np.all(np.isnan(s.v))


def test_bad_input_throws():
    import conicip_b200 as cb
    n = 10
    with pytest.raises(Exception):                                          # runtests.jl:507-523
        cb.conicIP(np.zeros((n, n)), np.arange(1.0, n + 1), np.eye(n + 2), np.zeros(n), [("R", n)])


def test_kktsolver_three_level_protocol():
    """The closure protocol of docs/src/guides/kkt_solvers.md:84-109 used directly."""
    import conicip_b200 as cb
    prob = P.mixed()
    cd = prob["cone_dims"]
    solve3x3gen = cb.kktsolver_b200(prob["Q"], prob["A"], prob["G"], cd)                   # LEVEL 1
    F = cb.Block([cb.Diagonal(np.full(k, 2.0)) for _, k in cd])
    solve3x3 = solve3x3gen(F, None)                                                         # LEVEL 2
    rng = np.random.default_rng(0)
    x, y, z = rng.standard_normal(96), rng.standard_normal(5), rng.standard_normal(prob["A"].shape[0])
    a, b, c = solve3x3(x, y, z)                                                             # LEVEL 3
    Q, A, G = prob["Q"], prob["A"], prob["G"]
    assert np.linalg.norm(Q @ a + G.T @ b - A.T @ c - x) < 1e-10 * np.linalg.norm(x) * 10
    assert np.linalg.norm(G @ a - y) < 1e-10
    assert np.linalg.norm(A @ a + 4.0 * c - z) < 1e-10 * np.linalg.norm(z) * 10


def _dense_h_problem(seed=0, n=10):
    rng = np.random.default_rng(seed)
    h = rng.standard_normal(n)
    H = np.outer(h, h)                                   # rank-1 "H = randn(n); H = H*H'" (runtests.jl:275-277)
    c = np.arange(1.0, n + 1)
    return H, H @ c, n, rng


def test_simplex_dense_rank1_H():
    """runtests.jl:271-303 (NumPy data): projection onto the simplex with a rank-one H.  H + A'W^-2A is
    positive definite because A = I."""
    import conicip_b200 as cb
    H, c, n, _ = _dense_h_problem()
    G, d = np.ones((1, n)), np.array([1.0])
    s = cb.conicIP(H, c, np.eye(n), np.zeros(n), [("R", n)], G, d, optTol=1e-7)
    so = O.conicIP(H, c, np.eye(n), np.zeros(n), [("R", n)], G, d, optTol=1e-7, kktsolver=O.kktsolver_qr)
    assert s.status == so.status == "Optimal" and abs(s.Iter - so.Iter) <= 1
    assert rel(s.y, so.y) < 1e-5 and abs(s.y.sum() - 1.0) < 1e-7 and s.y.min() > -1e-9


def test_linear_constraints_comparison():
    """runtests.jl:328-356: equalities through G must give the same y as the same equalities written as
    two inequality blocks [A; G; -G] >= [b; d; -d]."""
    import conicip_b200 as cb
    H, c, n, rng = _dense_h_problem(seed=1)
    H = H + 0.1 * np.eye(n)
    A, b = np.eye(n), np.zeros(n)
    G, d = rng.random((6, n)), np.zeros(6)
    y1 = cb.conicIP(H, c, A, b, [("R", n)], G, d, optTol=1e-7).y
    A2, b2 = np.vstack([A, G, -G]), np.concatenate([b, d, -d])
    y2 = cb.conicIP(H, c, A2, b2, [("R", n + 12)], optTol=1e-7).y
    assert np.linalg.norm(y1 - y2) < 1e-3                                 # the reference's tolerance


def test_infeasible_with_linear_constraints():
    """runtests.jl:462-485: x >= 0 together with x1 = -1."""
    import conicip_b200 as cb
    H, c, n, _ = _dense_h_problem(seed=2)
    G = np.zeros((1, n)); G[0, 0] = 1.0
    s = cb.conicIP(H, c, np.eye(n), np.zeros(n), [("R", n)], G, np.array([-1.0]), optTol=1e-7)
    so = O.conicIP(H, c, np.eye(n), np.zeros(n), [("R", n)], G, np.array([-1.0]), optTol=1e-7,
                   kktsolver=O.kktsolver_qr)
    assert s.status == so.status == "Infeasible"
    sn = cb.conicIP_native(H, c, np.eye(n), np.zeros(n), [("R", n)], G, np.array([-1.0]), optTol=1e-7)
    assert sn.status == "Infeasible" and np.all(np.isnan(sn.y))


def _b200_plugin_for_reference_loop(Q, A, G, cone_dims):
    """`kktsolver = kktsolver_b200` as the REFERENCE's host loop sees it: LEVEL 1 once, then every LEVEL-2 call hands
    over a host `Block` of Diagonal / SymWoodbury / VecCongurance blocks (here the oracle's classes, translated to the
    binding's -- in Julia they are ConicIP's own types and `flatten` reads them directly) and every LEVEL-3 call
    host vectors.  Nothing else of the loop touches the GPU: this is the drop-in of BASELINE's north star."""
    import conicip_b200 as cb
    gen = cb.kktsolver_b200(Q, A, G if G.shape[0] else None, cone_dims)

    def solve3x3gen(F, F_invT):
        blocks = []
        for B in F.blocks:
            if isinstance(B, O.Diag):
                blocks.append(cb.Diagonal(B.diag))
            elif isinstance(B, O.SymWoodbury):
                blocks.append(cb.SymWoodbury(B.Adiag, B.B, B.D))
            else:
                blocks.append(cb.VecCongurance(B.R))
        return gen(cb.Block(blocks), None)
    return solve3x3gen


@pytest.mark.parametrize("which", ["mixed", "sdp_mix"])
def test_reference_host_loop_with_b200_kktsolver_plugin(which):
    """The literal drop-in: the restated reference loop (oracle.conicIP, src/ConicIP.jl:468-939: host vectors, host
    cone arithmetic) with ONLY its `kktsolver` keyword replaced by the B200 engine through the three-level protocol
    (src/ConicIP.jl:667,682,688), against the same loop with the stock solvers."""
    if which == "mixed":
        prob = P.mixed()
        stock = O.pivot(O.kktsolver_2x2)
    else:
        rng = np.random.default_rng(3)
        k, n = 6, 25
        dim = k * (k + 1) // 2
        cones = [("R", 9), ("S", dim), ("Q", 5)]
        m = 9 + dim + 5
        A = rng.standard_normal((m, n)) / np.sqrt(n)
        s0 = np.zeros(m)
        s0[:9] = rng.uniform(0.2, 1.2, 9)
        B = rng.standard_normal((k, k))
        s0[9:9 + dim] = O.vecm(B @ B.T / k + 0.5 * np.eye(k))
        u = 0.1 * rng.standard_normal(4)
        s0[9 + dim] = 1 + np.linalg.norm(u); s0[10 + dim:] = u
        y0 = rng.standard_normal(n)
        G = rng.standard_normal((2, n)) / np.sqrt(n)
        prob = dict(Q=np.eye(n), c=rng.standard_normal(n), A=A, b=A @ y0 - s0, cone_dims=cones, G=G, d=G @ y0)
        stock = O.kktsolver_qr                     # the stock solver that is right for VecCongurance blocks
    args = (prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"], prob["G"], prob["d"])
    so = O.conicIP(*args, optTol=1e-8, kktsolver=stock)
    sb = O.conicIP(*args, optTol=1e-8, kktsolver=_b200_plugin_for_reference_loop)
    assert sb.status == so.status == "Optimal" and abs(sb.Iter - so.Iter) <= 1
    rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
    assert rel(sb.y, so.y) < 1e-6 and rel(sb.v, so.v) < 1e-6 and rel(sb.w, so.w) < 1e-6
    assert max(sb.prFeas, sb.duFeas, sb.muFeas) < 1e-8


@pytest.mark.parametrize("name,obj,x,xtol", [("moi_simple_lp", 1.0, [0.5, 0.5], 1e-2), ("moi_soc", np.sqrt(2.0), [1.0, 1.0, np.sqrt(2.0)], 1e-4),
                                             ("moi_max_sense", -2.0, [0.0, 1.0], 1e-2)])
def test_moi_wrapper_problems_on_the_device(name, obj, x, xtol):
    """test/runtests.jl:684-775 -- the LP, SOC and max-sense problems of the MOI wrapper tests (Q = 0) through both
    drivers: the reference's expected values, and the oracle's iterates."""
    import conicip_b200 as cb
    prob = getattr(P, name)()
    args = (prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"], prob["G"] if prob["G"].shape[0] else None,
            prob["d"] if prob["G"].shape[0] else None)
    so = O.conicIP(prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"], prob["G"], prob["d"], optTol=1e-6,
                   kktsolver=O.pivot(O.kktsolver_2x2))
    for solve in (cb.conicIP_native, cb.conicIP):
        s = solve(*args, optTol=1e-6)
        assert s.status == so.status == "Optimal" and abs(s.Iter - so.Iter) <= 1
        assert abs(-prob["c"] @ s.y - obj) < 1e-4 and np.abs(s.y - np.array(x)).max() < xtol
        assert np.linalg.norm(s.y - so.y) < 1e-6 * max(1.0, np.linalg.norm(so.y))

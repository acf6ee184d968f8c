"""Pin the oracle against the reference's own recorded goldens (test/runtests.jl).

The reference's `compare` (runtests.jl:15-21) only checks abs 1e-3; the recorded dict values
are far sharper known answers.  The recorded `:Iter` values predate the current stopping rule
(SURVEY.md section 4), so `Mu` is compared at the recorded iteration of the trajectory.
"""
import json
import os

import numpy as np
import pytest

import oracle as O
from conicip_b200 import problems as P

SOLVERS = {"qr": O.kktsolver_qr, "pivot2x2": O.pivot(O.kktsolver_2x2), "chol": O.kktsolver_chol}

# (problem, recorded Iter, recorded Mu, rtol)      -- test/runtests.jl:157-162, :197-202, :235-240
REF_GOLDENS = [
    (P.sphere, 5, 2.866608128093695e-7, 1e-7),
    (P.combined, 10, 4.663886012743681e-7, 1e-7),
]


def run(prob, solver, **kw):
    opts = dict(optTol=prob.get("optTol", 1e-7))
    opts.update(kw)
    return O.conicIP(prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"], prob["G"], prob["d"],
                     kktsolver=solver, **opts)


@pytest.mark.parametrize("sname", list(SOLVERS))
@pytest.mark.parametrize("gen,it,mu,rtol", REF_GOLDENS)
def test_reference_mu_trajectory(gen, it, mu, rtol, sname):
    s = run(gen(), SOLVERS[sname], DTB=0.01, maxRefinementSteps=3)
    assert s.status == "Optimal"
    mu_at = dict((t[0], t[1]) for t in s.trace)[it]
    assert abs(mu_at - mu) <= rtol * mu
    assert abs(s.Iter - it) <= 1


@pytest.mark.parametrize("sname", list(SOLVERS))
def test_simplex_golden_15_digits(sname):
    """runtests.jl:235-240: Iter 11, Mu = 2.7686402945528533e-9 (reproduced at optTol 1e-8)."""
    s = run(P.simplex(), SOLVERS[sname], optTol=1e-8)
    assert s.status == "Optimal" and s.Iter == 11
    assert abs(s.Mu - 2.7686402945528533e-9) < 1e-9 * 2.7686402945528533e-9 * 100
    y = np.zeros(10)
    y[9] = 1
    assert np.linalg.norm(s.y - y) < 1e-3                       # runtests.jl:230-233


def test_sphere_solution_and_residual_goldens():
    s = run(P.sphere(), O.kktsolver_qr)
    assert np.linalg.norm(s.y - np.ones(2) / np.sqrt(2)) < 1e-3     # runtests.jl:155
    tr = {t[0]: t for t in s.trace}
    assert abs(tr[5][4] - 1.621702501927476e-7) < 1e-12             # muFeas golden, runtests.jl:160
    assert tr[5][3] == 0.0                                           # prFeas golden 0.0


def test_combined_solution():
    s = run(P.combined(), O.kktsolver_qr)
    c = np.arange(1.0, 11)
    y = np.maximum(0, c)
    y /= np.linalg.norm(y)
    assert np.linalg.norm(s.y - y) < 1e-3                           # runtests.jl:192-195


def test_box_qp_pivot():
    """runtests.jl:90-131 (n reduced to 200 for CPU time): projection onto the box."""
    prob = P.box_qp(200)
    s = run(prob, O.pivot(O.kktsolver_2x2), DTB=0.01, maxRefinementSteps=3)
    assert s.status == "Optimal"
    c = np.arange(1.0, 201)
    assert np.linalg.norm(s.y - np.clip(c, -1, 1)) / 200 < 1e-3


def test_abandoned():
    """runtests.jl:246-269."""
    s = run(P.simplex(), O.kktsolver_qr, maxIters=2)
    assert s.status == "Abandoned"


@pytest.mark.parametrize("sname", ["qr", "pivot2x2"])
def test_infeasible_and_unbounded(sname):
    """runtests.jl:441-505 (NumPy data; qualitative status only)."""
    assert run(P.infeasible(), SOLVERS[sname]).status == "Infeasible"
    if sname == "qr":                                  # H = 0: only the QR solver tolerates singular H
        assert run(P.unbounded(), SOLVERS[sname]).status == "Unbounded"


def test_bad_input_throws():
    """runtests.jl:507-523."""
    n = 10
    with pytest.raises(Exception):
        O.conicIP(np.zeros((n, n)), np.arange(1.0, n + 1), np.eye(n + 2), np.zeros(n), [("R", n)])


def test_soc_direct():
    """runtests.jl:554-590."""
    s = run(P.soc_direct(), O.kktsolver_qr)
    assert s.status == "Optimal" and np.linalg.norm(s.y) < 1e-3


def test_sdp_projection():
    """runtests.jl:527-552: S cone of order 6; recorded Iter 6, solution diag(1,1,1,0,0,0)."""
    c = O.vecm(np.diag([1.0, 1, 1, -1, -1, -1]))
    s = O.conicIP(np.eye(21), c, np.eye(21), np.zeros(21), [("S", 21)], optTol=1e-7)
    assert s.status == "Optimal" and abs(s.Iter - 6) <= 1
    assert np.abs(O.mat(s.y) - np.diag([1.0, 1, 1, 0, 0, 0])).max() < 1e-3
    # the reference's own acceptance test for this record (`compare`, runtests.jl:15-21: absolute 1e-3 on every field)
    gold = dict(prFeas=4.2341217602756234e-16, Mu=3.4583513329836624e-10, muFeas=1.48267911727847e-9,
                duFeas=4.2341217602756234e-16)
    for k, g in gold.items():
        assert abs(getattr(s, k) - g) < 1e-3
    # Why `Mu` itself (3.458e-10 recorded, 6.745e-9 here) is not reproducible from the current source: on the central
    # path of this problem muFeas / Mu is a constant that depends only on the divisor of `mu = <v,s> / conedim`.
    # With conedim = ord = 6 (src/ConicIP.jl:551) it is sqrt(1.5) -- what the oracle gives -- while the recorded pair
    # has sqrt(1.5) * 21/6: the record was made by a version that divided by the vector length 21, i.e. it predates
    # the source under /root/reference (as the recorded Iter values do).
    assert abs(s.muFeas / s.Mu - np.sqrt(1.5)) < 1e-6
    assert abs((gold["muFeas"] / gold["Mu"]) / (s.muFeas / s.Mu) - 21 / 6) < 1e-3


def test_solvers_agree_on_mixed_problem():
    prob = P.mixed()
    sols = [run(prob, SOLVERS[k], optTol=1e-8) for k in SOLVERS]
    for s in sols[1:]:
        assert s.status == sols[0].status == "Optimal"
        assert abs(s.Iter - sols[0].Iter) <= 1
        assert np.linalg.norm(s.y - sols[0].y) <= 1e-6 * np.linalg.norm(sols[0].y)


def test_committed_golden_fixture_matches_oracle():
    """tests/golden/oracle_small.json is produced by tests/golden/make_golden.py; the oracle must
    keep reproducing it (and the GPU tests compare the engine to the same file)."""
    path = os.path.join(os.path.dirname(__file__), "golden", "oracle_small.json")
    gold = json.load(open(path))
    for name, g in gold["solves"].items():
        prob = getattr(P, name)()
        s = run(prob, O.kktsolver_chol, optTol=g["optTol"])
        assert s.status == g["status"] and s.Iter == g["Iter"]
        assert np.allclose(s.y, g["y"], rtol=1e-7, atol=1e-9)
        assert abs(s.Mu - g["Mu"]) <= 1e-6 * abs(g["Mu"])


@pytest.mark.parametrize("name,obj,x,xtol", [("moi_simple_lp", 1.0, [0.5, 0.5], 1e-2), ("moi_soc", np.sqrt(2.0), [1.0, 1.0, np.sqrt(2.0)], 1e-4),
                                             ("moi_max_sense", -2.0, [0.0, 1.0], 1e-2)])
def test_moi_wrapper_problems(name, obj, x, xtol):
    """test/runtests.jl:684-775 -- the three problems of the MOI wrapper tests in the form the wrapper hands to conicIP
    (Q = 0; src/MOI_wrapper.jl:142-285), with the reference's expected objective (atol 1e-4) and primal values."""
    prob = getattr(P, name)()
    for ks in SOLVERS.values():
        s = O.conicIP(prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"], prob["G"], prob["d"], optTol=1e-6, kktsolver=ks)
        assert s.status == "Optimal"
        assert abs(-prob["c"] @ s.y - obj) < 1e-4            # conicIP minimises 1/2 y'Qy - c'y
        assert np.abs(s.y - np.array(x)).max() < xtol

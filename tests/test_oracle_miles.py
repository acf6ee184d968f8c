"""The reference's "Miles's counterexamples" (test/runtests.jl:592-651, data test/testdata.jl:106-150): status-only
known-answer tests through `preprocess_conicIP`.  The oracle must reproduce every expected status; the data is the
reference's own fixture, extracted by tests/golden/make_miles.py."""
import os
import sys

import numpy as np
import pytest

import oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import miles  # noqa: E402

PROBLEMS = miles.load()


def solve(data, **kw):
    return O.preprocess_conicIP(data["Q"], data["c"], data["A"], data["b"], data["cone_dims"], data["G"], data["d"], **kw)


@pytest.mark.parametrize("name", ["miles_problem_1", "miles_problem_2"])
@pytest.mark.parametrize("ks", ["qr", "2x2"])
def test_miles_status(name, ks):
    p = PROBLEMS[name]
    c, A, b = miles.mpb_arrays(p)
    data = miles.mpb_to_conicip(c, A, b, p["con_cones"], p["var_cones"])
    kk = O.kktsolver_qr if ks == "qr" else O.pivot(O.kktsolver_2x2)
    assert solve(data, kktsolver=kk).status == p["expected_status"]          # :605, :615


def scaling_variants():
    """test/runtests.jl:621-648: kappa on everything, on (A, b) only, and the unscaled problem"""
    out = [("all", k) for k in (1e-8, 1e-6, 1e-4, 1, 1e4, 1e6, 1e8)]
    out += [("Ab", k) for k in (1e-4, 1, 1e4, 1e6)]
    return out


@pytest.mark.parametrize("which,kappa", scaling_variants())
def test_miles_problem_3_scaling(which, kappa):
    p = PROBLEMS["miles_problem_3"]
    c, A, b = miles.mpb_arrays(p)
    data = miles.mpb_to_conicip(kappa * c if which == "all" else c, kappa * A, kappa * b, p["con_cones"], p["var_cones"])
    assert solve(data).status == "Optimal"

"""The reference's "Miles's counterexamples" (test/runtests.jl:592-651) through the device path
(`cip_imcols` + `cip_ipm_solve`): the expected statuses, and agreement with the oracle where the solution is unique
enough to compare (objective value; these are degenerate LP/SOCPs with Q = 0)."""
import os
import sys

import numpy as np
import pytest

import oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import miles  # noqa: E402

pytestmark = pytest.mark.gpu
PROBLEMS = miles.load()


def both(data, **kw):
    import conicip_b200 as cb
    args = (data["Q"], data["c"], data["A"], data["b"], data["cone_dims"], data["G"], data["d"])
    return cb.preprocess_conicIP(*args, **kw), O.preprocess_conicIP(*args, **kw)


@pytest.mark.parametrize("name", ["miles_problem_1", "miles_problem_2"])
def test_miles_status_and_objective(name):
    p = PROBLEMS[name]
    c, A, b = miles.mpb_arrays(p)
    data = miles.mpb_to_conicip(c, A, b, p["con_cones"], p["var_cones"])
    s, so = both(data)
    assert s.status == so.status == p["expected_status"]
    if s.status == "Optimal":
        obj, objo = float(data["c"] @ s.y), float(data["c"] @ so.y)
        assert abs(obj - objo) <= 1e-5 * max(1.0, abs(objo))
        assert np.linalg.norm(data["G"] @ s.y - data["d"]) <= 1e-6 * max(1.0, np.linalg.norm(data["d"]))


@pytest.mark.parametrize("which,kappa", [("all", 1e-8), ("all", 1e-4), ("all", 1), ("all", 1e4), ("all", 1e8),
                                         ("Ab", 1e-4), ("Ab", 1e4), ("Ab", 1e6)])
def test_miles_problem_3_scaling(which, kappa):
    p = PROBLEMS["miles_problem_3"]
    c, A, b = miles.mpb_arrays(p)
    data = miles.mpb_to_conicip(kappa * c if which == "all" else c, kappa * A, kappa * b, p["con_cones"], p["var_cones"])
    s, so = both(data)
    assert s.status == so.status == "Optimal"

"""Extracts the reference's "Miles's counterexamples" (test/testdata.jl:106-150, used by test/runtests.jl:592-651 as
status-only known-answer tests) into tests/golden/miles_problems.npz (float64 / int32 arrays, compressed; the cone
lists as one small JSON string).  Runs only where /root/reference exists; the .npz travels to the GPU box.  The data stays in the MathProgBase form the reference stores it in; the conversion to the
solver's form is restated in tests/golden/miles.py."""
import json
import os
import re

import numpy as np

SRC = "/root/reference/test/testdata.jl"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "miles_problems.npz")
EXPECTED = {1: "Optimal", 2: "Infeasible", 3: "Optimal"}          # test/runtests.jl:605, :615, :627-648


def numbers(text):
    return [float(t) for t in text.split(",") if t.strip()]


def cones(text):
    return [[m.group(1), [int(t) for t in m.group(2).split(",") if t.strip()]]
            for m in re.finditer(r"\(:(\w+),\[([^\]]*)\]\)", text)]


def main():
    src = open(SRC).read()
    problems = []
    for k in (1, 2, 3):
        body = src[src.index(f"function miles_problem_{k}()"):]
        body = body[:body.index("\nend")]
        grab = lambda name: re.search(rf"^\s*{name} = \[(.*)\]\s*$", body, re.M).group(1)
        problems.append({
            "name": f"miles_problem_{k}", "expected_status": EXPECTED[k],
            "c": numbers(grab("c")), "b": numbers(grab("b")),
            "con_cones": cones(grab("con_cones")), "var_cones": cones(grab("var_cones")),
            "I": [int(v) for v in numbers(grab("I"))], "J": [int(v) for v in numbers(grab("J"))], "V": numbers(grab("V")),
        })
    arrays = {}
    meta = {"source": "test/testdata.jl:106-150 (1-based indices, MathProgBase conic form)", "problems": []}
    for p in problems:
        for key, dt in (("c", np.float64), ("b", np.float64), ("V", np.float64), ("I", np.int32), ("J", np.int32)):
            arrays[f"{p['name']}.{key}"] = np.asarray(p[key], dtype=dt)
        meta["problems"].append({k: p[k] for k in ("name", "expected_status", "con_cones", "var_cones")})
    np.savez_compressed(OUT, meta=np.array(json.dumps(meta)), **arrays)
    for p in problems:
        print(p["name"], "n", len(p["c"]), "rows", len(p["b"]), "nnz", len(p["V"]), p["con_cones"][0][0], p["expected_status"])


if __name__ == "__main__":
    main()

"""Identical-input dumps for a cross-run against the real reference.

Julia is not installed in the build image, so true-reference parity cannot be closed here; this script writes the
inputs of the BASELINE parity configurations exactly as the tests and the bench consume them -- raw little-endian
column-major float64, which is Julia's `Matrix{Float64}` memory layout -- plus this repo's results on them, and
`conicip.jl_b200/julia/crosscheck.jl` loads them, runs `ConicIP.conicIP` (stock `kktsolver_qr`, and
`ConicIPB200.kktsolver_b200` when the library is present) and prints iteration counts and relative differences.

    python scripts/dump_for_julia.py OUTDIR [C1 C2 C3 C5 ...] [--device]    # --device: also dump the B200 solution

Per configuration NAME: NAME.json (shapes, cone_dims, optTol, result summaries) and NAME_{Q,A,G}.f64 (column-major),
NAME_{c,b,d}.f64, NAME_oracle_{y,w,v}.f64, optionally NAME_b200_{y,w,v}.f64."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from conicip_b200 import problems as P


def colmajor(M):
    import scipy.sparse as sp
    M = M.toarray() if sp.issparse(M) else np.asarray(M, dtype=np.float64)
    return np.asfortranarray(M).ravel(order="F")


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    device = "--device" in sys.argv
    out = args[0]
    which = args[1:] or ["C1", "C3"]
    os.makedirs(out, exist_ok=True)
    gens = {"C1": P.config1, "C2": P.config2, "C3": P.config3, "C5": lambda: P.config5(n=2200, k=64, p=60),
            "C5full": P.config5}
    for name in which:
        prob = gens[name]()
        n, m, p = len(prob["c"]), prob["A"].shape[0], prob["G"].shape[0]
        meta = {"name": name, "n": n, "m": m, "p": p, "optTol": 1e-8,
                "cone_dims": [[t, int(k)] for t, k in prob["cone_dims"]], "layout": "column-major float64, little endian"}
        for key in ("Q", "A", "G"):
            colmajor(prob[key]).tofile(os.path.join(out, f"{name}_{key}.f64"))
        for key in ("c", "b", "d"):
            np.asarray(prob[key], dtype=np.float64).tofile(os.path.join(out, f"{name}_{key}.f64"))
        if name not in ("C2", "C5full"):                       # the oracle needs minutes beyond these sizes
            import oracle as O
            solver = O.kktsolver_qr if any(t == "S" for t, _ in prob["cone_dims"]) else O.kktsolver_chol
            so = O.conicIP(prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"], prob["G"], prob["d"],
                           optTol=1e-8, kktsolver=solver)
            for key in ("y", "w", "v"):
                np.asarray(getattr(so, key)).tofile(os.path.join(out, f"{name}_oracle_{key}.f64"))
            meta["oracle"] = {"status": so.status, "Iter": so.Iter, "Mu": so.Mu, "prFeas": so.prFeas,
                              "duFeas": so.duFeas, "muFeas": so.muFeas, "pobj": so.pobj}
        if device:
            import conicip_b200 as cb
            s = cb.conicIP_native(prob["Q"], prob["c"], prob["A"], prob["b"], prob["cone_dims"],
                                  prob["G"] if p else None, prob["d"] if p else None, optTol=1e-8)
            for key in ("y", "w", "v"):
                np.asarray(getattr(s, key)).tofile(os.path.join(out, f"{name}_b200_{key}.f64"))
            meta["b200"] = {"status": s.status, "Iter": s.Iter, "Mu": s.Mu, "prFeas": s.prFeas, "duFeas": s.duFeas,
                            "muFeas": s.muFeas, "pobj": s.pobj}
        json.dump(meta, open(os.path.join(out, f"{name}.json"), "w"), indent=1)
        print("wrote", name, meta.get("oracle"), meta.get("b200"), flush=True)


if __name__ == "__main__":
    main()

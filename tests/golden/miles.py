"""Loader for tests/golden/miles_problems.npz and a restatement of the reference's test-side conversion
`mpb_to_conicip` (test/testdata.jl:15-102): MathProgBase form  min c'x  s.t.  b - Ax in K_con, x in K_var
->  solver form  min 1/2 y'Qy - c'y  s.t.  Ay - b in K, Gy = d."""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    z = np.load(os.path.join(_HERE, "miles_problems.npz"))
    out = {}
    for p in json.loads(str(z["meta"]))["problems"]:
        for key in ("c", "b", "I", "J", "V"):
            p[key] = z[f"{p['name']}.{key}"]
        out[p["name"]] = p
    return out


def mpb_arrays(p):
    c = np.array(p["c"])
    b = np.array(p["b"])
    A = np.zeros((len(b), len(c)))
    np.add.at(A, (np.array(p["I"]) - 1, np.array(p["J"]) - 1), np.array(p["V"]))     # sparse(I, J, V) sums duplicates
    return c, A, b


def mpb_to_conicip(c, A, b, con_cones, var_cones):
    n = len(c)
    nA = np.linalg.norm(A)                                       # :23  scaling of the variable-cone rows
    eq, rowsA, rowsb, cone_dims = [], [], [], []
    for ctype, idx in con_cones:                                 # :29-46
        idx = np.array(idx) - 1
        if ctype == "Zero":
            eq.extend(idx)
        elif ctype == "NonPos":                                  # b - Ax <= 0  ->  Ax - b in R+
            rowsA.append(A[idx]); rowsb.append(b[idx]); cone_dims.append(("R", len(idx)))
        else:                                                    # NonNeg / SOC / SDP: negate A and b
            rowsA.append(-A[idx]); rowsb.append(-b[idx])
            cone_dims.append(({"NonNeg": "R", "SOC": "Q", "SDP": "S"}[ctype], len(idx)))
    G = A[eq] if eq else np.zeros((0, n))                        # :49-55
    d = b[eq] if eq else np.zeros(0)
    for vtype, idx in var_cones:                                 # :69-91
        if vtype == "Free":
            continue
        idx = np.array(idx) - 1
        blk = np.zeros((len(idx), n))
        blk[np.arange(len(idx)), idx] = -nA if vtype == "NonPos" else nA
        rowsA.append(blk); rowsb.append(np.zeros(len(idx)))
        cone_dims.append(({"NonNeg": "R", "NonPos": "R", "SOC": "Q", "SDP": "S"}[vtype], len(idx)))
    Ai = np.vstack(rowsA) if rowsA else np.zeros((0, n))         # :94-100
    bi = np.concatenate(rowsb) if rowsb else np.zeros(0)
    return dict(Q=np.zeros((n, n)), c=-c, A=Ai, b=bi, cone_dims=cone_dims, G=G, d=d)     # :102-106

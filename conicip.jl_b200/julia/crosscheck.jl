# Cross-run against the real reference on the inputs written by scripts/dump_for_julia.py:
#     julia --project=/path/to/ConicIP.jl conicip.jl_b200/julia/crosscheck.jl DUMPDIR [C1 C3 ...]
# For every configuration: conicIP with the stock kktsolver_qr (the reference itself), then -- if
# libconicip_b200.so can be loaded -- with ConicIPB200.kktsolver_b200, and the differences to the solutions this
# repository dumped (NAME_oracle_*.f64 from the NumPy oracle, NAME_b200_*.f64 from the device path).
using ConicIP, LinearAlgebra, SparseArrays
import JSON

dir = ARGS[1]
names = length(ARGS) > 1 ? ARGS[2:end] : ["C1", "C3"]
rd(name, key, dims...) = reshape(reinterpret(Float64, read(joinpath(dir, "$(name)_$(key).f64"))), dims...)
rel(a, b) = norm(a - b) / max(norm(b), floatmin())

have_b200 = try
    include(joinpath(@__DIR__, "ConicIPB200.jl")); true
catch err
    @warn "ConicIPB200 not loaded" err; false
end

for name in names
    meta = JSON.parsefile(joinpath(dir, "$(name).json"))
    n, m, p = meta["n"], meta["m"], meta["p"]
    Q = Matrix(rd(name, "Q", n, n)); A = Matrix(rd(name, "A", m, n)); G = Matrix(rd(name, "G", p, n))
    c = Vector(rd(name, "c", n)); b = Vector(rd(name, "b", m)); d = Vector(rd(name, "d", p))
    cone_dims = [(String(t), Int(k)) for (t, k) in meta["cone_dims"]]
    sol = conicIP(Q, c, A, b, cone_dims, G, d; optTol = meta["optTol"])            # the reference, stock solver
    println("$name reference: $(sol.status) Iter=$(sol.Iter) Mu=$(sol.Mu) prFeas=$(sol.prFeas) duFeas=$(sol.duFeas) muFeas=$(sol.muFeas)")
    for who in ("oracle", "b200")
        isfile(joinpath(dir, "$(name)_$(who)_y.f64")) || continue
        y = rd(name, "$(who)_y", n); v = rd(name, "$(who)_v", m); w = rd(name, "$(who)_w", p)
        println("  vs $who dump: dIter=$(sol.Iter - meta[who]["Iter"]) rel(y)=$(rel(y, sol.y)) rel(v)=$(rel(v, sol.v))",
                p > 0 ? " rel(w)=$(rel(w, sol.w))" : "")
    end
    if have_b200
        s2 = conicIP(Q, c, A, b, cone_dims, G, d; optTol = meta["optTol"], kktsolver = ConicIPB200.kktsolver_b200)
        println("  kktsolver_b200 behind the stock loop: $(s2.status) Iter=$(s2.Iter) rel(y)=$(rel(s2.y, sol.y)) rel(v)=$(rel(s2.v, sol.v))")
    end
end

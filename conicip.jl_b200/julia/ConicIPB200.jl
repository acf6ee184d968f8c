# ConicIPB200.jl -- the Julia-side binding a ConicIP.jl maintainer adds to use the B200 engine.
#
# It is a thin `ccall` shim over libconicip_b200.so (include/conicip_b200.h): no CUDA.jl, no kernel
# generation, no CPU fallback.  `kktsolver_b200` has exactly the signature of
# `ConicIP.kktsolver_qr` / `pivot(ConicIP.kktsolver_2x2)` (src/kktsolvers.jl:18,349), so
#
#     sol = conicIP(Q, c, A, b, cone_dims, G, d; kktsolver = ConicIPB200.kktsolver_b200)
#
# is the whole integration.  Julia is not installed in the build image, so this file is not
# executed by the test-suite; tests/ drive the same C symbols through ctypes (_lib.py) with the
# identical call sequence.
module ConicIPB200

using LinearAlgebra, SparseArrays
import ConicIP
using ConicIP: Block, VecCongurance
using ConicIP.WoodburyMatrices: SymWoodbury

const LIB = get(ENV, "CONICIP_B200_LIB", joinpath(@__DIR__, "..", "libconicip_b200.so"))

const CONE_CODE = Dict("R" => Cint(0), "Q" => Cint(1), "S" => Cint(2))
const BLK_DIAG, BLK_WOODBURY, BLK_VECCONG = Cint(0), Cint(1), Cint(2)

struct Options            # mirrors cip_options
    struct_size::Cint
    device::Cint
    reg_delta::Cdouble
    reg_eps_G::Cdouble
    q_kind::Cint
    verbose::Cint
    dist_chol::Cint
    aug_rho::Cdouble
end

lasterr() = unsafe_string(ccall((:cip_last_error, LIB), Cstring, ()))
function check(rc::Cint)
    rc < 0 && error("conicip_b200: ", lasterr())
    return rc
end

struct Csc                # mirrors cip_csc: a SparseMatrixCSC{Float64,Int64} passed field by field
    nrows::Cint
    ncols::Cint
    colptr::Ptr{Int64}
    rowval::Ptr{Int64}
    nzval::Ptr{Cdouble}
    index_base::Cint
end
Csc(M::SparseMatrixCSC{Float64,Int64}) = Csc(size(M, 1), size(M, 2), pointer(M.colptr), pointer(M.rowval), pointer(M.nzval), 1)

mutable struct Engine
    h::Ptr{Cvoid}
    n::Int; m::Int; p::Int
    function Engine(Q, A, G, cone_dims; reg_delta = 0.0)
        n = size(Q, 1); m = size(A, 1); p = size(G, 1)
        if A isa SparseMatrixCSC          # sparse LEVEL 1: no dense copy on the host (cip_create_csc)
            Qs = SparseMatrixCSC{Float64,Int64}(sparse(Q)); As = SparseMatrixCSC{Float64,Int64}(A)
            Gs = SparseMatrixCSC{Float64,Int64}(sparse(G))
            ct = Cint[CONE_CODE[t] for (t, _) in cone_dims]; cdim = Cint[k for (_, k) in cone_dims]
            opts = Ref(Options(Cint(sizeof(Options)), Cint(-1), reg_delta, 0.0, Cint(2), Cint(0), Cint(-1), -1.0))
            hp = Ref{Ptr{Cvoid}}(C_NULL)
            GC.@preserve Qs As Gs begin
                rc = ccall((:cip_create_csc, LIB), Cint,
                           (Ref{Ptr{Cvoid}}, Cint, Ref{Csc}, Ref{Csc}, Ref{Csc}, Cint, Ptr{Cint}, Ptr{Cint}, Ref{Options}),
                           hp, n, Ref(Csc(Qs)), Ref(Csc(As)), Ref(Csc(Gs)), length(ct), ct, cdim, opts)
            end
            rc != 0 && error("cip_create_csc: ", lasterr())
            e = new(hp[], n, m, p)
            finalizer(x -> ccall((:cip_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), e)
            return e
        end
        # Julia arrays are column-major: dense copies are passed as they are (SURVEY 8b "Argument types")
        Qd = Q isa Diagonal ? collect(Q.diag) : Matrix{Float64}(Q)
        qk = Q isa Diagonal ? Cint(1) : Cint(0)
        Ad = Matrix{Float64}(A)
        Gd = Matrix{Float64}(G)
        ct = Cint[CONE_CODE[t] for (t, _) in cone_dims]
        cdim = Cint[k for (_, k) in cone_dims]
        opts = Ref(Options(Cint(sizeof(Options)), Cint(-1), reg_delta, 0.0, qk, Cint(0), Cint(-1), -1.0))
        hp = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:cip_create, LIB), Cint,
                   (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Cint,
                    Ptr{Cdouble}, Cint, Cint, Ptr{Cint}, Ptr{Cint}, Ref{Options}),
                   hp, n, m, p, Qd, max(n, 1), Ad, max(m, 1), Gd, max(p, 1),
                   length(ct), ct, cdim, opts)
        rc != 0 && error("cip_create: ", lasterr())
        e = new(hp[], n, m, p)
        finalizer(x -> ccall((:cip_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), e)   # handle lifetime = closure lifetime
        return e
    end
end

# Flatten a ConicIP `Block` (src/blockmatrices.jl:35-43) into the arrays of cip_factor.
# The first call of a solve passes Diagonal(ones) for every cone, Q/S slots included
# (src/ConicIP.jl:704) -- the per-call `kind` array covers that.
function flatten(F::Block)
    nc = length(F.Blocks)
    kind = Vector{Cint}(undef, nc); fD = zeros(nc)
    fa = Float64[]; fb = Float64[]; fR = Float64[]
    for (i, B) in enumerate(F.Blocks)
        if B isa Diagonal
            kind[i] = BLK_DIAG; append!(fa, B.diag); append!(fb, zeros(length(B.diag)))
        elseif B isa SymWoodbury
            kind[i] = BLK_WOODBURY
            append!(fa, B.A.diag); append!(fb, vec(B.B)); fD[i] = B.D isa Number ? B.D : B.D[1, 1]
        elseif B isa VecCongurance
            kind[i] = BLK_VECCONG
            k = size(B, 1); append!(fa, zeros(k)); append!(fb, zeros(k)); append!(fR, vec(B.R))
        else
            error("unsupported scaling block type $(typeof(B))")
        end
    end
    return kind, fa, fb, fD, fR
end

"""
    kktsolver_b200(Q, A, G, cone_dims) -> solve3x3gen

Drop-in for `kktsolver_qr` / `pivot(kktsolver_2x2)`; three-level closure protocol of
docs/src/guides/kkt_solvers.md:84-109.
"""
function kktsolver_b200(Q, A, G, cone_dims)
    eng = Engine(Q, A, G, cone_dims)                                  # LEVEL 1: upload once

    function solve3x3gen(F, F⁻ᵀ)                                      # LEVEL 2: form H, factor
        kind, fa, fb, fD, fR = flatten(F)
        rc = ccall((:cip_factor, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                   eng.h, kind, fa, fb, fD, isempty(fR) ? C_NULL : pointer(fR))
        check(rc)                 # rc > 0: non-PD pivot -> NaN results -> status :Error (src/ConicIP.jl:870-873)
        failed = rc > 0

        function solve3x3(y, w, v)                                    # LEVEL 3: solve
            a = Vector{Float64}(undef, eng.n)                         # fresh outputs (mutated by axpy4!, :920)
            b = Vector{Float64}(undef, eng.p)
            c = Vector{Float64}(undef, eng.m)
            if failed
                fill!(a, NaN); fill!(b, NaN); fill!(c, NaN)
                return (a, b, c)
            end
            check(ccall((:cip_solve, LIB), Cint,
                        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                        eng.h, Vector{Float64}(y), Vector{Float64}(w), Vector{Float64}(v), a, b, c))
            return (a, b, c)
        end
        return solve3x3
    end
    return solve3x3gen
end

# ---- cone kernels (no callback exists for these in conicIP; a device-resident driver calls them)
nt_scaling!(eng::Engine, v, s, λ) = check(ccall((:cip_nt_scaling, LIB), Cint,
    (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), eng.h, v, s, λ))
function maxstep(eng::Engine, x, d = nothing; scale = 1.0)
    α = Ref{Cdouble}(0.0)
    check(ccall((:cip_maxstep, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ref{Cdouble}),
                eng.h, x, d === nothing ? C_NULL : pointer(d), scale, α))
    return α[]
end
apply!(eng::Engine, op::Integer, x, y) = check(ccall((:cip_apply, LIB), Cint,
    (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}), eng.h, op, x, y))
cone_prod!(eng::Engine, o, x, y) = check(ccall((:cip_cone_prod, LIB), Cint,
    (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), eng.h, x, y, o))
cone_div!(eng::Engine, o, x, y) = check(ccall((:cip_cone_div, LIB), Cint,
    (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), eng.h, x, y, o))

# ---- the whole solve behind one ccall (cip_ipm_solve; SURVEY 8f rank 1): same arguments, options
# and `Solution` as `ConicIP.conicIP`, with the loop of src/ConicIP.jl:468-939 running in the library
# on device-resident vectors.
struct IpmOptions
    struct_size::Cint; maxIters::Cint; maxRefinementSteps::Cint; verbose::Cint
    optTol::Cdouble; DTB::Cdouble; infeasTol::Cdouble; refinementThreshold::Cdouble
end
struct IpmResult
    status::Cint; Iter::Cint; factors::Cint; solves::Cint
    Mu::Cdouble; prFeas::Cdouble; duFeas::Cdouble; muFeas::Cdouble; pobj::Cdouble; dobj::Cdouble; seconds::Cdouble
end
const STATUS = (:None, :Optimal, :Infeasible, :Unbounded, :Abandoned, :Error)

function conicIP_b200(Q, c::AbstractVector, A, b::AbstractVector, cone_dims,
                      G = spzeros(0, length(c)), d = zeros(0);
                      optTol = 1e-6, DTB = 0.01, verbose = false, maxRefinementSteps = 3, maxIters = 100,
                      infeasTol = optTol, refinementThreshold = optTol / 1e7)
    eng = Engine(Q, A, G, cone_dims)
    y = Vector{Float64}(undef, eng.n); w = Vector{Float64}(undef, eng.p); v = Vector{Float64}(undef, eng.m)
    opts = Ref(IpmOptions(Cint(sizeof(IpmOptions)), maxIters, maxRefinementSteps, verbose, optTol, DTB, infeasTol,
                          refinementThreshold))
    res = Ref(IpmResult(0, 0, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0))
    check(ccall((:cip_ipm_solve, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{IpmOptions}, Ptr{Cdouble}, Ptr{Cdouble},
                 Ptr{Cdouble}, Ref{IpmResult}),
                eng.h, Vector{Float64}(c), Vector{Float64}(b), Vector{Float64}(d), opts, y, w, v, res))
    r = res[]
    return ConicIP.Solution(y, w, v, STATUS[r.status + 1], r.Iter, r.Mu, r.prFeas, r.duFeas, r.muFeas, r.pobj, r.dobj)
end

# ---- preprocessor (SURVEY 8f rank 4): drop-in for ConicIP.imcols (src/preprocessor.jl:10-28) on the device.
# Returns (R, consistent) with 1-based sorted row indices, R empty when the system is inconsistent.
function imcols_b200(A, b, ϵ = 1e-8; device = -1)
    Ad = Matrix{Float64}(A); p, n = size(Ad)
    keep = zeros(Cint, max(p, 1)); nkeep = Ref{Cint}(0); cons = Ref{Cint}(1)
    check(ccall((:cip_imcols, LIB), Cint,
                (Cint, Ptr{Cdouble}, Cint, Cint, Cint, Ptr{Cdouble}, Cdouble, Ptr{Cint}, Ref{Cint}, Ref{Cint}),
                device, Ad, max(p, 1), p, n, Vector{Float64}(b), ϵ, keep, nkeep, cons))
    cons[] == 0 && return (Int[], false)
    return (findall(!iszero, keep[1:p]), true)
end
# `preprocess_conicIP` itself needs no change beyond calling imcols_b200 at src/preprocessor.jl:58-59.

# ---- MOI: `ConicIP.Optimizer` has no kktsolver field (src/MOI_wrapper.jl:19-31) and optimize!
# forwards only verbose/optTol/maxIters (:278-282).  The one-field extension a maintainer adds:
#
#     mutable struct Optimizer ...; kktsolver::Function; end          # default ConicIP.kktsolver_qr
#     sol = preprocess_conicIP(Q, c, A, b, cone_dims, G, d; verbose, optTol, maxIters,
#                              kktsolver = optimizer.kktsolver)       # preprocessor forwards options... (:44,:82-84)
#
# after which `ConicIP.Optimizer(kktsolver = ConicIPB200.kktsolver_b200)` selects the engine.

end # module

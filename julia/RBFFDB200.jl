# RBFFDB200.jl -- drop-in Julia shim: the reference's exported hot-path functions forwarded to librbffd.so via ccall.
#
# NOT EXECUTED IN THIS REPO'S CI: the build image has no Julia toolchain (see DESIGN.md).  It is the binding a
# maintainer of RadialBasisFiniteDifferences.jl would add; every ccall matches a prototype in include/rbffd.h and the
# same symbols are exercised by the ctypes mirror (radialbasisfinitedifferences.jl_b200/api.py) in tests/.
#
# Usage inside the reference package (src/RadialBasisFiniteDifferences.jl):
#     include("RBFFDB200.jl"); using .RBFFDB200
#     generate_operator(X, Y, p, n, polydeg) = RBFFDB200.generate_operator(X, Y, p, n, polydeg)
# Signatures, argument order, 1-based indices and SparseMatrixCSC return types are those of
# src/generate_operator.jl:29,192, src/hyperviscosity_operator.jl:26,177 and src/calculateneighbors.jl:1.
module RBFFDB200

using SparseArrays, StaticArrays

const LIB = get(ENV, "RBFFD_LIB", joinpath(@__DIR__, "..", "radialbasisfinitedifferences.jl_b200", "librbffd.so"))
const MAX_OPS = 12

struct Options            # mirrors rbffd_options (include/rbffd.h)
    dim::Int32; p::Int32; polydeg::Int32; n::Int32; nops::Int32
    ops::NTuple{48,Int32}
    index_base::Int32; sort_columns::Int32; kernel::Int32; variant::Int32
    reserved::NTuple{4,Int32}
end

function make_options(dim, p, polydeg, n, ops::Vector{NTuple{4,Int}}; variant = 0)
    flat = zeros(Int32, 4 * MAX_OPS)
    for (i, o) in enumerate(ops), j in 1:4
        flat[4 * (i - 1) + j] = o[j]
    end
    Options(dim, p, polydeg, n, length(ops), Tuple(flat), 1, 0, 0, variant, (0, 0, 0, 0))   # index_base = 1: Julia
end

const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function context()
    if CTX[] == C_NULL
        rc = ccall((:rbffd_create, LIB), Cint, (Cint, Ptr{Ptr{Cvoid}}), 0, CTX)
        rc == 0 || error(unsafe_string(ccall((:rbffd_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    end
    CTX[]
end
check(rc) = rc == 0 || error("librbffd: " * unsafe_string(ccall((:rbffd_last_error, LIB), Cstring, (Ptr{Cvoid},), context())))

# UnitRange index sets of processmesh.jl:87,174,183 -> per-node group codes (0 interior, 1+2b boundary b, 2+2b ghosts b)
function groups(N, X_idx_in, X_idx_bc, X_idx_bc_g)
    g = zeros(Int32, N)
    for (b, r) in enumerate(X_idx_bc);   g[r] .= 1 + 2 * (b - 1); end
    for (b, r) in enumerate(X_idx_bc_g); g[r] .= 2 + 2 * (b - 1); end
    g
end

# Vector{SVector{2,Float64}} is already the interleaved layout the C ABI wants: pointer(X) is a Ptr{Float64} of length 2N
coords(X) = (Xc = convert(Vector{SVector{2,Float64}}, X); (Xc, Ptr{Float64}(pointer(Xc))))

function generate(X, Y, p, n, polydeg, ops, grp; variant = 0)
    Xc, px = coords(X); Yc, py = Y === nothing ? (Xc, Ptr{Float64}(C_NULL)) : coords(Y)
    N, M = length(Xc), length(Yc)
    opts = Ref(make_options(2, p, polydeg, n, ops; variant = variant))
    colind = Matrix{Int64}(undef, n, M)                 # row-major [M][n] in C == column-major (n, M) in Julia
    vals = Array{Float64,3}(undef, n, M, length(ops))
    GC.@preserve Xc Yc colind vals grp begin
        check(ccall((:rbffd_generate_operator_host, LIB), Cint,
                    (Ptr{Cvoid}, Ref{Options}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}),
                    context(), opts, px, N, py, M, grp === nothing ? C_NULL : pointer(grp), colind, vals))
    end
    rows = repeat((1:M)', n)                             # generate_operator.jl:171-182: sparse(I, J, V), zeros kept
    [sparse(vec(rows), vec(colind), vec(view(vals, :, :, o))) for o in 1:length(ops)]
end

const REF_OPS = [(0, 1, 0, 0), (0, 0, 1, 0), (0, 2, 0, 0), (0, 0, 2, 0), (0, 1, 1, 0)]   # Dx, Dy, Dxx, Dyy, Dxy

"generate_operator(X, Y, p, n, polydeg) -> E, Dx, Dy, Dxx, Dyy, Dxy        (src/generate_operator.jl:29)"
function generate_operator(X, Y, p, n, polydeg)
    m = generate(X, Y, p, n, polydeg, vcat([(0, 0, 0, 0)], REF_OPS), nothing)
    return m[1], m[2], m[3], m[4], m[5], m[6]
end
"boundary-aware method (src/generate_operator.jl:192); Y_idx_* are accepted and unused, as in the reference"
function generate_operator(X, Y, p, n, polydeg, X_idx_in, X_idx_bc, X_idx_bc_g, Y_idx_in, Y_idx_bc, Y_idx_bc_g)
    m = generate(X, Y, p, n, polydeg, vcat([(0, 0, 0, 0)], REF_OPS), groups(length(X), X_idx_in, X_idx_bc, X_idx_bc_g))
    return m[1], m[2], m[3], m[4], m[5], m[6]
end

"legacy collocated method generate_operator(X, p, n, polydeg)   (src/generate_operator.jl:354): variant = 1"
function generate_operator(X, p, n, polydeg)
    m = generate(X, nothing, p, n, polydeg, vcat([(0, 0, 0, 0)], REF_OPS), nothing; variant = 1)
    return m[1], m[2], m[3], m[4], m[5], m[6]
end
"legacy collocated method hyperviscosity_operator(k_deriv, X, p, n, polydeg)   (src/hyperviscosity_operator.jl:314)"
function hyperviscosity_operator(k_deriv, X, p, n, polydeg)
    m = generate(X, nothing, p, n, polydeg, [(0, k_deriv, 0, 0), (0, 0, k_deriv, 0)], nothing; variant = 1)
    return m[1], m[2]
end

"hyperviscosity_operator(k_deriv, X, Y, p, n, polydeg[, index sets]) -> Dxk, Dyk   (src/hyperviscosity_operator.jl:26,177)"
function hyperviscosity_operator(k_deriv, X, Y, p, n, polydeg, sets...)
    grp = isempty(sets) ? nothing : groups(length(X), sets[1], sets[2], sets[3])
    m = generate(X, Y, p, n, polydeg, [(0, k_deriv, 0, 0), (0, 0, k_deriv, 0)], grp)
    return m[1], m[2]
end

"calculateneighbors(X, Y, n, X_idx_in, X_idx_bc, X_idx_bc_g, Y_idx_in, Y_idx_bc, Y_idx_bc_g)   (src/calculateneighbors.jl:1)"
function calculateneighbors(X, Y, n, X_idx_in, X_idx_bc, X_idx_bc_g, Y_idx_in, Y_idx_bc, Y_idx_bc_g)
    Xc, px = coords(X); Yc, py = coords(Y)
    N, M = length(Xc), length(Yc)
    grp = groups(N, X_idx_in, X_idx_bc, X_idx_bc_g)
    idx = Matrix{Int64}(undef, n, N); idy = Vector{Int64}(undef, M)
    dx = Matrix{Float64}(undef, n, N); dy = Vector{Float64}(undef, M)
    GC.@preserve Xc Yc grp idx idy dx dy begin
        check(ccall((:rbffd_calculateneighbors_host, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int32, Int32, Ptr{Int32}, Int32,
                     Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
                    context(), px, N, py, M, 2, n, grp, 1, idx, idy, dx, dy))
    end
    idxs_x = [SVector{n}(view(idx, :, i)) for i in 1:N];  dists_x = [SVector{n}(view(dx, :, i)) for i in 1:N]
    return idxs_x, [SVector{1}(idy[i]) for i in 1:M], dists_x, [SVector{1}(dy[i]) for i in 1:M]
end

# ---- device-resident operators for the time loop (examples/adv_diff_test.jl:144-199): matrices never leave HBM ----
struct AdvDiffParams
    iE::Int32; iDx::Int32; iDy::Int32; iDxx::Int32; iDyy::Int32; iDxk::Int32; iDyk::Int32; reserved::Int32
    alpha::Float64; ux::Float64; uy::Float64; gamma::Float64
end
mutable struct DeviceOperator
    h::Ptr{Cvoid}
end
"upload SparseMatrixCSC operators that share one pattern (as returned above) once; returns a handle"
function DeviceOperator(mats::Vector{<:SparseMatrixCSC})
    M, N = size(mats[1]); At = [sparse(transpose(A)) for A in mats]      # CSC of A' == CSR of A
    n = At[1].colptr[2] - At[1].colptr[1]
    colind = At[1].rowval; vals = reduce(vcat, (a.nzval for a in At))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rbffd_operator_from_host, LIB), Cint,
                (Ptr{Cvoid}, Int64, Int64, Int32, Int32, Ptr{Int64}, Int32, Ptr{Float64}, Ptr{Ptr{Cvoid}}),
                context(), M, N, n, length(mats), colind, 1, vals, h))
    op = DeviceOperator(h[]); finalizer(o -> ccall((:rbffd_operator_destroy, LIB), Cint, (Ptr{Cvoid},), o.h), op); op
end
"du = E' * (α Dxx u + α Dyy u - ux Dx u - uy Dy u) - γ (Dxk + Dyk) u        (adv_diff_test.jl:151-152)"
function rhs_advdiff!(du::Vector{Float64}, op::DeviceOperator, prm::AdvDiffParams, u::Vector{Float64})
    check(ccall((:rbffd_rhs_advdiff_host, LIB), Cint, (Ptr{Cvoid}, Ref{AdvDiffParams}, Ptr{Float64}, Ptr{Float64}), op.h, Ref(prm), u, du))
    du
end

end # module

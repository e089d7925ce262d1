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
    index_width::Int32          # 0 / 64: Int64 indices (SparseMatrixCSC{Float64,Int64}); 32: Int32 indices, no widening pass
    reserved::NTuple{3,Int32}
end

function make_options(dim, p, polydeg, n, ops::Vector{NTuple{4,Int}}; variant = 0, index_width = 64)
    flat = zeros(Int32, 4 * MAX_OPS)
    for (i, o) in enumerate(ops), j in 1:4
        flat[4 * (i - 1) + j] = o[j]
    end
    Options(dim, p, polydeg, n, length(ops), Tuple(flat), 1, 0, 0, variant, index_width, (0, 0, 0))   # index_base = 1: Julia
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

# Ti = Int32 asks the library for Int32 indices (opts.index_width = 32): half the PCIe bytes of the pattern and no widening
# pass on the host; the results are SparseMatrixCSC{Float64,Int32}, which every downstream use in the reference accepts.
function generate(X, Y, p, n, polydeg, ops, grp; variant = 0, Ti::Type = Int64)
    Xc, px = coords(X); Yc, py = Y === nothing ? (Xc, Ptr{Float64}(C_NULL)) : coords(Y)
    N, M = length(Xc), length(Yc)
    opts = Ref(make_options(2, p, polydeg, n, ops; variant = variant, index_width = 8 * sizeof(Ti)))
    colind = Matrix{Ti}(undef, n, M)                    # row-major [M][n] in C == column-major (n, M) in Julia
    vals = Array{Float64,3}(undef, n, M, length(ops))
    GC.@preserve Xc Yc colind vals grp begin
        check(ccall((:rbffd_generate_operator_host, LIB), Cint,
                    (Ptr{Cvoid}, Ref{Options}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Cvoid}, Ptr{Float64}),
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
const ADVDIFF_COLLOCATED = Int32(1)   # rows == nodes: E = I, the whole line of cons_sys is ONE pass over the shared pattern
struct AdvDiffParams
    iE::Int32; iDx::Int32; iDy::Int32; iDxx::Int32; iDyy::Int32; iDxk::Int32; iDyk::Int32; flags::Int32
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


# ---- the whole time loop on the device, without CUDA.jl: operators generated straight into HBM, field vectors in library
#      buffers, cons_sys + ghost updates + SSP-RK stage combinations as library launches (examples/adv_diff_test.jl:144-199) ----
"generate the operators of `ops` on the device from host node sets; the matrices never cross PCIe"
function DeviceOperator(X, Y, p, n, polydeg, ops::Vector{NTuple{4,Int}}, grp = nothing)
    Xc, px = coords(X); Yc, py = Y === nothing ? (Xc, Ptr{Float64}(C_NULL)) : coords(Y)
    opts = Ref(make_options(2, p, polydeg, n, ops)); h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve Xc Yc grp begin
        check(ccall((:rbffd_operator_generate_host, LIB), Cint,
                    (Ptr{Cvoid}, Ref{Options}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Ptr{Cvoid}}),
                    context(), opts, px, length(Xc), py, length(Yc), grp === nothing ? C_NULL : pointer(grp), h))
    end
    op = DeviceOperator(h[]); finalizer(o -> ccall((:rbffd_operator_destroy, LIB), Cint, (Ptr{Cvoid},), o.h), op); op
end

mutable struct DeviceVector
    p::Ptr{Float64}; n::Int
end
function DeviceVector(v::Vector{Float64})
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rbffd_device_malloc, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}), context(), 8 * length(v), r))
    d = DeviceVector(Ptr{Float64}(r[]), length(v)); copyto!(d, v)
    finalizer(x -> ccall((:rbffd_device_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), context(), x.p), d); d
end
Base.copyto!(d::DeviceVector, v::Vector{Float64}) = (check(ccall((:rbffd_device_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64), context(), d.p, v, 8 * d.n)); d)
Base.Vector(d::DeviceVector) = (v = Vector{Float64}(undef, d.n); check(ccall((:rbffd_device_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64), context(), v, d.p, 8 * d.n)); v)

"du = interior line of cons_sys on device vectors (one launch with flags = ADVDIFF_COLLOCATED)"
rhs_advdiff!(du::DeviceVector, op::DeviceOperator, prm::AdvDiffParams, u::DeviceVector) =
    (check(ccall((:rbffd_rhs_advdiff_device, LIB), Cint, (Ptr{Cvoid}, Ref{AdvDiffParams}, Ptr{Float64}, Ptr{Float64}), op.h, Ref(prm), u.p, du.p)); du)
"out = a*u + b*(x + dt*du): Shu-Osher stage combination (out may alias u or x)"
stage_update!(out::DeviceVector, a, u::DeviceVector, b, x::DeviceVector, dt, du::DeviceVector) =
    (check(ccall((:rbffd_stage_update_device, LIB), Cint, (Ptr{Cvoid}, Int64, Float64, Ptr{Float64}, Float64, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Float64}),
                 context(), out.n, a, u.p, b, x.p, dt, du.p, out.p)); out)
"y = sum_i coef[i] * D[which[i]] * x in one pass over the shared pattern (which: 1-based matrix numbers)"
function spmv_multi!(y::DeviceVector, op::DeviceOperator, which::Vector{Int}, coef::Vector{Float64}, x::DeviceVector)
    w = Int32.(which .- 1)
    check(ccall((:rbffd_spmv_multi_device, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), op.h, length(w), w, coef, x.p, y.p)); y
end
"out = a*u + b*(x + dt * sum_i coef[i] D[which[i]] x): one SSP-RK stage of a linear semidiscretisation as ONE launch"
function spmv_stage!(out::DeviceVector, op::DeviceOperator, which::Vector{Int}, coef::Vector{Float64}, x::DeviceVector, a, u::DeviceVector, b, dt)
    w = Int32.(which .- 1)
    check(ccall((:rbffd_spmv_stage_device, LIB), Cint,
                (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Float64, Float64, Ptr{Float64}),
                op.h, length(w), w, coef, x.p, a, u.p, b, dt, out.p)); out
end

# ghost-node updates of cons_sys (adv_diff_test.jl:118-141 set-up, :162-176 per call)
mutable struct BoundaryConditions
    h::Ptr{Cvoid}
end
"kind[b] = 1: u[ghost_b] = -inv(D[bc_b, ghost_b]) (D[bc_b, rest] u) with D = matrix which[b] (1-based); kind[b] = 0: u[bc_b] = u[ghost_b] = value[b]"
function BoundaryConditions(op::DeviceOperator, kind::Vector{Int}, which::Vector{Int}, value::Vector{Float64}, bc::Vector{<:AbstractVector{Int}}, ghost::Vector{<:AbstractVector{Int}})
    ptr = Int64[0; cumsum(length.(bc))]; b = Int64.(reduce(vcat, bc)); g = Int64.(reduce(vcat, ghost)); h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rbffd_bc_create, LIB), Cint,
                (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{Ptr{Cvoid}}),
                op.h, length(kind), Int32.(kind), Int32.(which .- 1), value, ptr, b, g, 1, h))
    x = BoundaryConditions(h[]); finalizer(o -> ccall((:rbffd_bc_destroy, LIB), Cint, (Ptr{Cvoid},), o.h), x); x
end
apply!(bc::BoundaryConditions, u::DeviceVector) = (check(ccall((:rbffd_bc_apply_device, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), bc.h, u.p)); u)

"""
One SSP-RK3 step of `cons_sys` (adv_diff_test.jl:144-188) entirely on the device: per stage the interior line, the ghost
updates of the stage vector (after du is formed, as in the reference) and the stage combination -- nine launches, no host data.
"""
function ssprk3_step!(u::DeviceVector, u1::DeviceVector, u2::DeviceVector, du::DeviceVector, op, bc, prm, dt)
    rhs_advdiff!(du, op, prm, u);  apply!(bc, u);  stage_update!(u1, 0.0, u, 1.0, u, dt, du)
    rhs_advdiff!(du, op, prm, u1); apply!(bc, u1); stage_update!(u2, 0.75, u, 0.25, u1, dt, du)
    rhs_advdiff!(du, op, prm, u2); apply!(bc, u2); stage_update!(u, 1 / 3, u, 2 / 3, u2, dt, du)
    u
end

# ---- spatial-block shards over the GPUs of one box (one Julia process per GPU; BASELINE.json north_star, SURVEY.md §8e) ----
# The three collective steps of the wiring (ship every rank's halo requests to the owners; all-gather the 64-byte IPC handles and
# the offset tables; connect) are the host language's job: MPI.jl `Alltoallv` / `Allgather` here, torch.distributed in the Python
# mirror (radialbasisfinitedifferences.jl_b200/sharding.py: Shard.wire).
mutable struct Shard
    h::Ptr{Cvoid}; rank::Int; nparts::Int; n_owned::Int; n_interior::Int; n_halo::Int
end
"part[i] in 0:nparts-1: coordinate-quantile blocks (blocks = (b0, b1) fixes the block grid)"
function shard_plan(X, nparts; blocks = nothing)
    Xc, px = coords(X); part = Vector{Int32}(undef, length(Xc))
    b = blocks === nothing ? C_NULL : Int32[blocks..., 1][1:3]
    rc = ccall((:rbffd_shard_plan_host, LIB), Cint, (Ptr{Float64}, Int64, Int32, Int32, Ptr{Int32}, Ptr{Int32}), px, length(Xc), 2, nparts, b, part)
    rc == 0 || error("rbffd_shard_plan_host failed ($rc)"); part
end
function Shard(X, part::Vector{Int32}, nparts, rank, n)
    Xc, px = coords(X); h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rbffd_shard_create_host, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Ptr{Int32}, Int32, Int32, Int32, Ptr{Ptr{Cvoid}}),
                context(), px, length(Xc), 2, part, nparts, rank, n, h))
    a, b, c = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    check(ccall((:rbffd_shard_info, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), h[], a, b, c))
    s = Shard(h[], rank, nparts, a[], b[], c[]); finalizer(x -> ccall((:rbffd_shard_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), s); s
end
"local id (owned rows first, then halo) -> the caller's 1-based node number"
global_ids(s::Shard) = (g = Vector{Int64}(undef, s.n_owned + s.n_halo); check(ccall((:rbffd_shard_global_ids_host, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int64}), s.h, 1, g)); g)
function recv_ids(s::Shard, peer)
    c = Ref{Int64}(0); check(ccall((:rbffd_shard_recv_count, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int64}), s.h, peer, c))
    ids = Vector{Int64}(undef, c[]); check(ccall((:rbffd_shard_recv_ids_host, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Int64}), s.h, peer, 1, ids)); ids
end
set_send_ids!(s::Shard, peer, ids::Vector{Int64}) = check(ccall((:rbffd_shard_set_send_ids_host, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Int64}, Int64), s.h, peer, 1, ids, length(ids)))
"-> (64-byte IPC handle, offsets[2 nparts + 1]) to all-gather"
function finalize!(s::Shard)
    handle = Vector{UInt8}(undef, 64); off = Vector{Int64}(undef, 2 * s.nparts + 1)
    check(ccall((:rbffd_shard_finalize, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Int64}), s.h, handle, off)); handle, off
end
"peer_off = the offsets table the PEER returned from finalize!"
connect!(s::Shard, peer, handle::Vector{UInt8}, peer_off::Vector{Int64}) =
    check(ccall((:rbffd_shard_connect, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{UInt8}, Int64, Int64, Int64), s.h, peer, handle,
                peer_off[2 * s.rank + 1], peer_off[2 * s.rank + 2], peer_off[2 * s.nparts + 1]))
"operators of the shard's owned rows over its local pattern (no communication)"
function shard_operator(s::Shard, p, n, polydeg, ops::Vector{NTuple{4,Int}})
    X = Ref{Ptr{Float64}}(C_NULL); st = Ref{Ptr{Int32}}(C_NULL)
    check(ccall((:rbffd_shard_device_arrays, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Float64}}, Ptr{Ptr{Int32}}), s.h, X, st))
    nnz = s.n_owned * n; ci = Ref{Ptr{Cvoid}}(C_NULL); va = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rbffd_device_malloc, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}), context(), 4 * nnz, ci))
    check(ccall((:rbffd_device_malloc, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}), context(), 8 * nnz * length(ops), va))
    opts = Ref(make_options(2, p, polydeg, n, ops))
    check(ccall((:rbffd_weights_device, LIB), Cint,
                (Ptr{Cvoid}, Ref{Options}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Int32}, Int64, Ptr{Int32}, Ptr{Cvoid}, Ptr{Cvoid}),
                context(), opts, X[], s.n_owned + s.n_halo, X[], s.n_owned, st[], s.n_owned, C_NULL, ci[], va[]))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rbffd_operator_from_device, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
                context(), s.n_owned, s.n_owned + s.n_halo, n, length(ops), ci[], va[], h))
    DeviceOperator(h[])      # borrows ci / va: keep them alive as long as the operator (freed with rbffd_device_free)
end
"y[1:n_owned] = sum_i coef[i] D[which[i]] [x ; halo of x]: ONE launch, halo exchange fused (NVLink peer stores)"
function spmv!(y::DeviceVector, s::Shard, op::DeviceOperator, which::Vector{Int}, coef::Vector{Float64}, x::DeviceVector)
    w = Int32.(which .- 1)
    check(ccall((:rbffd_shard_spmv_device, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), s.h, op.h, length(w), w, coef, x.p, y.p)); y
end
"y[1:n_owned] = alpha D[which]' v + beta y with the reverse halo exchange (E' * v of adv_diff_test.jl:151)"
spmv_t!(y::DeviceVector, s::Shard, op::DeviceOperator, which::Int, v::DeviceVector; alpha = 1.0, beta = 0.0) =
    (check(ccall((:rbffd_shard_spmv_t_device, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64, Ptr{Float64}, Float64, Ptr{Float64}), s.h, op.h, which - 1, alpha, v.p, beta, y.p)); y)
"out = a*u + b*(x + dt * sum_i coef[i] D[which[i]] [x ; halo]): halo exchange + product + SSP-RK stage, ONE launch"
function spmv_stage!(out::DeviceVector, s::Shard, op::DeviceOperator, which::Vector{Int}, coef::Vector{Float64}, x::DeviceVector, a, u::DeviceVector, b, dt)
    w = Int32.(which .- 1)
    check(ccall((:rbffd_shard_spmv_stage_device, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Float64, Float64, Ptr{Float64}),
                s.h, op.h, length(w), w, coef, x.p, a, u.p, b, dt, out.p)); out
end

end # module

# EDCuda.jl -- ccall shim that puts libedcuda.so behind ExactDiagonalization.jl's own types.
#
# STATUS: written against include/edcuda.h, NOT EXECUTED: the build image has no Julia (and no network to
# install one).  The same C ABI is exercised end to end from Python (exactdiagonalization.jl_b200/edcuda, tests/).
#
# Usage (drop-in for the Hamiltonian-application path):
#
#     using ExactDiagonalization, EDCuda
#     hs, σ = ExactDiagonalization.Toolkit.spin_half_system(32)
#     H     = simplify(sum(σ(i, μ) * σ(mod1(i + 1, 32), μ) for i in 1:32, μ in (:x, :y, :z)))
#     hsr   = EDCuda.represent(HilbertSpaceSector(hs, 0))        # basis generated on the GPU
#     Hrep  = EDCuda.represent(hsr, H)                           # <: AbstractMatrix, works with Arpack/KrylovKit
#     y     = Hrep * x;  mul!(y, Hrep, x);  apply!(y, Hrep, x);  S = sparse(Hrep)
#
# Every method below replaces the reference method named in its comment (paths under ExactDiagonalization.jl/src).
module EDCuda

using LinearAlgebra, SparseArrays
import ExactDiagonalization
const ED = ExactDiagonalization

const libedcuda = get(ENV, "EDCUDA_LIB", joinpath(@__DIR__, "..", "libedcuda.so"))

const ED_F64, ED_C128 = Cint(0), Cint(1)
const SIDE_LEFT, SIDE_RIGHT = Cint(0), Cint(1)

last_error() = unsafe_string(ccall((:ed_last_error, libedcuda), Cstring, ()))

# status code -> the reference's exception types
function check(status::Cint)
    status == 0 && return nothing
    msg = last_error()
    status == 1 && throw(ArgumentError(msg))
    status == 2 && throw(DimensionMismatch(msg))
    status == 3 && throw(BoundsError())
    status == 4 && throw(KeyError(msg))
    error("edcuda error $status: $msg")
end

dtype_code(::Type{Float64}) = ED_F64
dtype_code(::Type{ComplexF64}) = ED_C128

# ---------------------------------------------------------------------------------------------------------------
# HilbertSpace -> ed_space   (HilbertSpace/hilbert_space.jl:25-41, site.jl:69-93)
mutable struct SpaceHandle
    ptr::Ptr{Cvoid}
    function SpaceHandle(hs::ED.HilbertSpace{QN}) where {QN}
        n_states = Int32[length(site.states) for site in hs.sites]
        n_qn = length(QN.parameters)
        qn = Int64[Int64(q) for site in hs.sites for state in site.states for q in state.quantum_number]
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ed_space_create, libedcuda), Cint, (Int32, Ptr{Int32}, Ptr{Int64}, Int32, Ref{Ptr{Cvoid}}),
                    length(hs.sites), n_states, qn, n_qn, out))
        h = new(out[])
        finalizer(x -> ccall((:ed_space_destroy, libedcuda), Cint, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
end

# ---------------------------------------------------------------------------------------------------------------
# HilbertSpaceRepresentation   (Representation/hilbert_space_representation.jl:16-73, 215-256)
mutable struct GpuHilbertSpaceRepresentation{HS<:ED.AbstractHilbertSpace, BR<:Unsigned} <: ED.AbstractHilbertSpaceRepresentation{Bool}
    hilbert_space::HS
    space::SpaceHandle
    ptr::Ptr{Cvoid}
end

ED.basespace(hsr::GpuHilbertSpaceRepresentation) = hsr.hilbert_space
ED.bintype(::Type{GpuHilbertSpaceRepresentation{HS, BR}}) where {HS, BR} = BR

function ED.dimension(hsr::GpuHilbertSpaceRepresentation)           # :90
    d = Ref{Int64}(0)
    check(ccall((:ed_basis_dim, libedcuda), Cint, (Ptr{Cvoid}, Ref{Int64}), hsr.ptr, d))
    return Int(d[])
end

function _wrap_basis(hs, space, ptr, ::Type{BR}) where {BR}
    h = GpuHilbertSpaceRepresentation{typeof(hs), BR}(hs, space, ptr)
    finalizer(x -> ccall((:ed_basis_destroy, libedcuda), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end

# represent(hs, BR) / represent(HilbertSpaceSector(hs, qn), BR)     (:215-230; basis :109-206, on device here)
function represent(hs::ED.HilbertSpace, ::Type{BR}=UInt) where {BR<:Unsigned}
    space = SpaceHandle(hs)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ed_basis_generate, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Int32, Ref{Ptr{Cvoid}}),
                space.ptr, C_NULL, -1, 8 * sizeof(BR), out))
    return _wrap_basis(hs, space, out[], BR)
end

function represent(hss::ED.HilbertSpaceSector{HS, QN}, ::Type{BR}=UInt) where {HS, QN, BR<:Unsigned}
    hs = ED.basespace(hss)
    space = SpaceHandle(hs)
    allowed = Int64[Int64(q) for qn in sort(collect(hss.allowed_quantum_numbers)) for q in qn]
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ed_basis_generate, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Int32, Ref{Ptr{Cvoid}}),
                space.ptr, allowed, length(hss.allowed_quantum_numbers), 8 * sizeof(BR), out))
    return _wrap_basis(hs, space, out[], BR)
end

# represent(hs, basis_list)    (:242-256: sorts if unsorted, duplicates -> ArgumentError)
function represent(hs::ED.AbstractHilbertSpace, basis_list::AbstractVector{BR}) where {BR<:Unsigned}
    base = ED.basespace(hs)
    space = SpaceHandle(base)
    words = UInt64.(basis_list)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ed_basis_from_list, libedcuda), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Int64, Int32, Ref{Ptr{Cvoid}}),
                space.ptr, words, length(words), 8 * sizeof(BR), out))
    return _wrap_basis(base, space, out[], BR)
end

# hsr.basis_list  (fetched lazily: the device owns it)
function basis_list(hsr::GpuHilbertSpaceRepresentation{HS, BR}) where {HS, BR}
    n = ED.dimension(hsr)
    words = Vector{UInt64}(undef, n)
    check(ccall((:ed_basis_download, libedcuda), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{UInt64}), hsr.ptr, 0, n, words))
    return BR.(words)
end

# get(hsr.basis_lookup, key, -1)    (frozensortedarray.jl:29-48)
function basis_lookup(hsr::GpuHilbertSpaceRepresentation, keys::AbstractVector{<:Unsigned})
    k = UInt64.(keys)
    idx = Vector{Int64}(undef, length(k))
    check(ccall((:ed_basis_lookup, libedcuda), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Int64, Ptr{Int64}), hsr.ptr, k, length(k), idx))
    return idx
end

# ---------------------------------------------------------------------------------------------------------------
# SumOperator -> ed_operator   (Operator/pure_operator.jl:25-52, sum_operator.jl:13-23)
mutable struct OperatorHandle
    ptr::Ptr{Cvoid}
    iscomplex::Bool
end

_terms(op::ED.PureOperator) = [op]
_terms(op::ED.SumOperator) = op.terms
_terms(::ED.NullOperator) = ED.PureOperator{Float64, UInt}[]

function OperatorHandle(op::ED.AbstractOperator)
    terms = _terms(op)
    S = valtype(op)
    cplx = S <: Complex
    mask = UInt64[t.bitmask for t in terms]
    row = UInt64[t.bitrow for t in terms]
    col = UInt64[t.bitcol for t in terms]
    amp = cplx ? reinterpret(Float64, ComplexF64[ComplexF64(t.amplitude) for t in terms]) : Float64[Float64(t.amplitude) for t in terms]
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ed_operator_create, libedcuda), Cint,
                (Int64, Ptr{UInt64}, Ptr{UInt64}, Ptr{UInt64}, Ptr{Float64}, Int32, Ref{Ptr{Cvoid}}),
                length(terms), mask, row, col, collect(amp), cplx ? 1 : 0, out))
    h = OperatorHandle(out[], cplx)
    finalizer(x -> ccall((:ed_operator_destroy, libedcuda), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end

# ---------------------------------------------------------------------------------------------------------------
# OperatorRepresentation / ReducedOperatorRepresentation
#   (Representation/operator_representation.jl:13-36, Symmetry/reduced_operator_representation.jl:16-42)
mutable struct GpuOperatorRepresentation{S<:Number, SP, O<:ED.AbstractOperator} <: ED.AbstractOperatorRepresentation{S}
    space::SP
    operator::O
    ophandle::OperatorHandle
    ptr::Ptr{Cvoid}
end

ED.get_space(opr::GpuOperatorRepresentation) = opr.space
Base.size(opr::GpuOperatorRepresentation) = (d = ED.dimension(opr.space); (d, d))

function represent(hsr::GpuHilbertSpaceRepresentation, op::ED.AbstractOperator)
    oh = OperatorHandle(op)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ed_oprep_create, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), hsr.ptr, oh.ptr, out))
    S = oh.iscomplex ? ComplexF64 : Float64
    h = GpuOperatorRepresentation{S, typeof(hsr), typeof(op)}(hsr, op, oh, out[])
    finalizer(x -> ccall((:ed_oprep_destroy, libedcuda), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end

# apply!(out, opr, state): out += opr * state       (abstract_operator_representation.jl:260-267, 296-316, 358-378)
function ED.apply!(out::Vector{T}, opr::GpuOperatorRepresentation, state::Vector{T}) where {T<:Union{Float64, ComplexF64}}
    check(ccall((:ed_apply, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int32, Int32, Int32),
                opr.ptr, out, length(out), state, length(state), dtype_code(T), SIDE_LEFT, 1))
    return out
end

# apply!(out, state, opr): out += state * opr       (:278-285, 327-347, 389-409)
function ED.apply!(out::Vector{T}, state::Vector{T}, opr::GpuOperatorRepresentation) where {T<:Union{Float64, ComplexF64}}
    check(ccall((:ed_apply, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int32, Int32, Int32),
                opr.ptr, out, length(out), state, length(state), dtype_code(T), SIDE_RIGHT, 1))
    return out
end

# mul!(out, opr, state): out = opr * state          (:110-118)
function LinearAlgebra.mul!(out::Vector{T}, opr::GpuOperatorRepresentation, state::Vector{T}) where {T<:Union{Float64, ComplexF64}}
    check(ccall((:ed_apply, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int32, Int32, Int32),
                opr.ptr, out, length(out), state, length(state), dtype_code(T), SIDE_LEFT, 0))
    return out
end

# opr * state, state * opr                           (operator_representation.jl:122-139)
function Base.:(*)(opr::GpuOperatorRepresentation{S}, state::AbstractVector{SV}) where {S, SV<:Number}
    T = promote_type(S, SV) <: Complex ? ComplexF64 : Float64
    out = zeros(T, ED.dimension(opr.space))
    return ED.apply!(out, opr, Vector{T}(state))
end
function Base.:(*)(state::AbstractVector{SV}, opr::GpuOperatorRepresentation{S}) where {S, SV<:Number}
    T = promote_type(S, SV) <: Complex ? ComplexF64 : Float64
    out = zeros(T, ED.dimension(opr.space))
    return ED.apply!(out, Vector{T}(state), opr)
end

# sparse(opr; tol)                                   (abstract_operator_representation.jl:136-204; CSC, 1-based Int)
function SparseArrays.sparse(opr::GpuOperatorRepresentation{S}; tol::Real=Base.rtoldefault(Float64)) where {S}
    nnz = Ref{Int64}(0)
    check(ccall((:ed_sparse_count, libedcuda), Cint, (Ptr{Cvoid}, Float64, Ref{Int64}), opr.ptr, Float64(tol), nnz))
    n = ED.dimension(opr.space)
    colptr = Vector{Int64}(undef, n + 1); rowval = Vector{Int64}(undef, nnz[]); nzval = Vector{S}(undef, nnz[])
    check(ccall((:ed_sparse_fetch, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid}), opr.ptr, colptr, rowval, nzval))
    return SparseMatrixCSC{S, Int}(n, n, colptr, rowval, nzval)
end

# Matrix(opr)                                        (:121-132)
function Base.Matrix(opr::GpuOperatorRepresentation{S}) where {S}
    n = ED.dimension(opr.space)
    out = zeros(S, n, n)
    check(ccall((:ed_dense, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), opr.ptr, out))
    return out
end

# get_row_iterator / get_column_iterator / get_element     (operator_representation.jl:66-119)
function _iterator(opr::GpuOperatorRepresentation{S}, i::Integer, side::Cint) where {S}
    cap = max(length(_terms(opr.operator)), 1)
    idx = Vector{Int64}(undef, cap); amp = Vector{Float64}(undef, 2cap); n = Ref{Int64}(0)
    check(ccall((:ed_oprep_row_iterator, libedcuda), Cint, (Ptr{Cvoid}, Int64, Int32, Int64, Ptr{Int64}, Ptr{Float64}, Ref{Int64}),
                opr.ptr, i, side, cap, idx, amp, n))
    vals = S <: Complex ? reinterpret(ComplexF64, amp)[1:n[]] : amp[1:n[]]
    return (Int(idx[k]) => S(vals[k]) for k in 1:n[])
end
ED.get_row_iterator(opr::GpuOperatorRepresentation, irow::Integer) = _iterator(opr, irow, SIDE_LEFT)
ED.get_column_iterator(opr::GpuOperatorRepresentation, icol::Integer) = _iterator(opr, icol, SIDE_RIGHT)
function ED.get_element(opr::GpuOperatorRepresentation{S}, irow::Integer, icol::Integer) where {S}
    v = Vector{Float64}(undef, 2)
    check(ccall((:ed_oprep_get_element, libedcuda), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}), opr.ptr, irow, icol, v))
    return S <: Complex ? ComplexF64(v[1], v[2]) : v[1]
end

# ---------------------------------------------------------------------------------------------------------------
# symmetry_reduce(hsr, symops_and_amplitudes; tol)   (Symmetry/symmetry_reduce_generic.jl:7-14, 22-255)
mutable struct GpuReducedHilbertSpaceRepresentation{HSR, BR, C} <: ED.AbstractHilbertSpaceRepresentation{C}
    parent::HSR
    symptr::Ptr{Cvoid}
    ptr::Ptr{Cvoid}
end

ED.basespace(r::GpuReducedHilbertSpaceRepresentation) = ED.basespace(r.parent)
function ED.dimension(r::GpuReducedHilbertSpaceRepresentation)
    d = Ref{Int64}(0)
    check(ccall((:ed_rbasis_dim, libedcuda), Cint, (Ptr{Cvoid}, Ref{Int64}), r.ptr, d))
    return Int(d[])
end

# any AbstractSymmetryOperation (SitePermutation, GlobalBitFlip, DirectProductOperation) -> (0-based site map, flip)
_flatten(p::ED.SitePermutation, n) = (Int32.(p.permutation.map .- 1), false)
_flatten(b::ED.GlobalBitFlip, n) = (Int32.(0:n-1), b.value)
function _flatten(d::ED.DirectProductOperation, n)
    perm, flip = Int32.(0:n-1), false
    for op in reverse(d.operations)                  # (ABC)(ψ) = A(B(C(ψ)))   symmetry_apply.jl:65-77
        p, f = _flatten(op, n)
        perm = Int32[p[perm[i] + 1] for i in 1:n]
        flip = xor(flip, f)
    end
    return (perm, flip)
end

function symmetry_reduce(hsr::GpuHilbertSpaceRepresentation{HS, BR}, symops_and_amplitudes::AbstractArray;
                         tol::Real=Base.rtoldefault(Float64)) where {HS, BR}
    n = length(ED.basespace(hsr).sites)
    flat = [_flatten(op, n) for (op, _) in symops_and_amplitudes]
    perm = reduce(vcat, first.(flat)); flip = UInt8[f[2] for f in flat]
    chi = reinterpret(Float64, ComplexF64[ComplexF64(a) for (_, a) in symops_and_amplitudes])
    sym = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ed_symmetry_create, libedcuda), Cint, (Int32, Int32, Ptr{Int32}, Ptr{UInt8}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                length(flat), n, perm, flip, collect(chi), sym))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ed_symmetry_reduce, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Ref{Ptr{Cvoid}}), hsr.ptr, sym[], Float64(tol), out))
    r = GpuReducedHilbertSpaceRepresentation{typeof(hsr), BR, ComplexF64}(hsr, sym[], out[])
    finalizer(r) do x
        ccall((:ed_rbasis_destroy, libedcuda), Cint, (Ptr{Cvoid},), x.ptr)
        ccall((:ed_symmetry_destroy, libedcuda), Cint, (Ptr{Cvoid},), x.symptr)
    end
    return r
end

# the reference's IrrepComponent entry points pass conj(irrep value) and store it un-conjugated
# (symmetry_reduce_translation.jl:26-27, 68): identical to handing get_irrep_iterator's pairs to the generic API.
symmetry_reduce(hsr::GpuHilbertSpaceRepresentation, ssic::ED.AbstractSymmetryIrrepComponent; kwargs...) =
    symmetry_reduce(hsr, collect(ED.get_irrep_iterator(ssic)); kwargs...)

function basis_list(r::GpuReducedHilbertSpaceRepresentation{HSR, BR, C}) where {HSR, BR, C}
    n = ED.dimension(r)
    words = Vector{UInt64}(undef, n)
    check(ccall((:ed_rbasis_download, libedcuda), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{UInt64}), r.ptr, 0, n, words))
    return BR.(words)
end

# rhsr.basis_mapping_index / rhsr.basis_mapping_amplitude, materialised on request (parent dimension sized)
function basis_mapping(r::GpuReducedHilbertSpaceRepresentation)
    n = ED.dimension(r.parent)
    idx = Vector{Int64}(undef, n); amp = Vector{ComplexF64}(undef, n)
    check(ccall((:ed_rbasis_mapping_rows, libedcuda), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Cvoid}), r.ptr, 0, n, idx, amp))
    return (Int.(idx), amp)
end

function represent(rhsr::GpuReducedHilbertSpaceRepresentation, op::ED.AbstractOperator)   # reduced_operator_representation.jl:40-42
    oh = OperatorHandle(op)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ed_oprep_create_reduced, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), rhsr.ptr, oh.ptr, out))
    h = GpuOperatorRepresentation{ComplexF64, typeof(rhsr), typeof(op)}(rhsr, op, oh, out[])
    finalizer(x -> ccall((:ed_oprep_destroy, libedcuda), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end

# symmetry_reduce(rhsr, large) / symmetry_unreduce(rhsr, small)     (Symmetry/symmetry_reduce.jl:29-56, 208-225)
function symmetry_reduce(r::GpuReducedHilbertSpaceRepresentation, large::Vector{T}) where {T<:Union{Float64, ComplexF64}}
    small = zeros(ComplexF64, ED.dimension(r))
    check(ccall((:ed_vector_reduce, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int32, Int32),
                r.ptr, small, length(small), large, length(large), dtype_code(T), 0))
    return small
end
function symmetry_unreduce(r::GpuReducedHilbertSpaceRepresentation, small::Vector{T}) where {T<:Union{Float64, ComplexF64}}
    large = zeros(ComplexF64, ED.dimension(r.parent))
    check(ccall((:ed_vector_unreduce, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int32),
                r.ptr, large, length(large), small, length(small), dtype_code(T)))
    return large
end

# ---------------------------------------------------------------------------------------------------------------
# Lanczos (not in the reference; replaces `eigs(Hrep; ...)` of docs/src/examples/spinhalf.md:26 for the lowest levels)
function lanczos(opr::GpuOperatorRepresentation{S}, nsteps::Integer; seed::Integer=0, nritz::Integer=4) where {S}
    alpha = zeros(nsteps); beta = zeros(nsteps); ritz = zeros(nritz); done = Ref{Int32}(0)
    check(ccall((:ed_lanczos, libedcuda), Cint,
                (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, UInt64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Ref{Int32}),
                opr.ptr, nsteps, C_NULL, dtype_code(S), seed, alpha, beta, ritz, nritz, done))
    return (alpha=alpha[1:done[]], beta=beta[1:done[]], ritz=ritz)
end

# ---------------------------------------------------------------------------------------------------------------
# Cached matrix (not in the reference): keep the assembled rows on device so that repeated `mul!`s (Arpack, KrylovKit)
# run as an SpMV instead of redoing the term walk / orbit searches; `side` = SIDE_LEFT for opr*x, SIDE_RIGHT for x*opr.
function cache_matrix!(opr, side::Cint=SIDE_LEFT)
    nnz = Ref{Int64}(0)
    check(ccall((:ed_oprep_cache_matrix, libedcuda), Cint, (Ptr{Cvoid}, Int32, Ref{Int64}), opr.ptr, side, nnz))
    return nnz[]
end
drop_cache!(opr) = (check(ccall((:ed_oprep_drop_cache, libedcuda), Cint, (Ptr{Cvoid},), opr.ptr)); opr)

# Row shard of a representation: apply!/mul! then read the full x and write rows lo+1:hi of out.
set_rows!(opr, lo::Integer, hi::Integer) =
    (check(ccall((:ed_oprep_set_rows, libedcuda), Cint, (Ptr{Cvoid}, Int64, Int64), opr.ptr, lo, hi)); opr)

# ---------------------------------------------------------------------------------------------------------------
# isinvariant(hs, symop, op) for every element at once (Symmetry/symmetry_apply.jl:110-135).  `represent(rhsr, op)`
# above already fails with ArgumentError for a non-invariant operator (ed_oprep_create_reduced runs this check on the
# host before any kernel is launched); this is the explicit form.
function isinvariant_all(rhsr::GpuReducedHilbertSpaceRepresentation, op::ED.AbstractOperator; tol::Real=Base.rtoldefault(Float64))
    oh = OperatorHandle(op)
    inv = Ref{Int32}(0); bad = Ref{Int32}(-1)
    check(ccall((:ed_operator_isinvariant, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Ref{Int32}, Ref{Int32}),
                rhsr.parent.space.ptr, rhsr.symptr, oh.ptr, Float64(tol), inv, bad))
    return inv[] != 0, Int(bad[]) + 1          # (invariant?, 1-based index of the first violating element or 0)
end

# ---------------------------------------------------------------------------------------------------------------
# Checkpoints (not in the reference): raw binary files bound to the Hilbert space / symmetry by hashes.
save(hsr::GpuHilbertSpaceRepresentation, path::AbstractString) =
    check(ccall((:ed_basis_save, libedcuda), Cint, (Ptr{Cvoid}, Cstring), hsr.ptr, path))
save(r::GpuReducedHilbertSpaceRepresentation, path::AbstractString) =
    check(ccall((:ed_rbasis_save, libedcuda), Cint, (Ptr{Cvoid}, Cstring), r.ptr, path))
function load_basis(hs::ED.AbstractHilbertSpace, path::AbstractString, ::Type{BR}=UInt) where {BR<:Unsigned}
    space = SpaceHandle(ED.basespace(hs)); out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ed_basis_load, libedcuda), Cint, (Ptr{Cvoid}, Cstring, Ref{Ptr{Cvoid}}), space.ptr, path, out))
    return _wrap_basis(hs, space, out[], BR)
end

# ---------------------------------------------------------------------------------------------------------------
# Multi-GPU (not in the reference, whose apply! spreads rows over Threads.@threads:
# Representation/abstract_operator_representation.jl:260-267, 358-378).  One Julia process drives all GPUs:
#
#     ctx  = EDCuda.Context([0, 1, 2, 3])                       # ed_ctx_create: ncclCommInitAll inside the library
#     sh   = EDCuda.ShardedOperator(ctx, dev -> EDCuda.represent(EDCuda.represent(sector), H))
#     x, y = EDCuda.DVector(sh), EDCuda.DVector(sh);  EDCuda.upload!(x, x_host)
#     mul!(y, sh, x);  EDCuda.download!(y_host, y);  res = EDCuda.lanczos(sh, 100)
#
# The row partition, the halo exchange over NVLink and the NCCL collectives all live in libedcuda.so.
mutable struct Context
    ptr::Ptr{Cvoid}; devices::Vector{Int32}
    function Context(devices::AbstractVector{<:Integer})
        out = Ref{Ptr{Cvoid}}(C_NULL); d = Int32.(devices)
        check(ccall((:ed_ctx_create, libedcuda), Cint, (Int32, Ptr{Int32}, Ref{Ptr{Cvoid}}), length(d), d, out))
        c = new(out[], d)
        finalizer(x -> ccall((:ed_ctx_destroy, libedcuda), Cint, (Ptr{Cvoid},), x.ptr), c)
        return c
    end
end

mutable struct ShardedOperator{S}
    ptr::Ptr{Cvoid}; ctx::Context; oprs::Vector{Any}; dim::Int
end
# make(dev) builds the representation on device `dev` (the library's current device is set before the call).
# exchange: 0 = automatic (tiled kernel: halo exchange by copy-engine pulls; everything else: NCCL all-gather of x),
#           1 = all-gather, 2 = pulls, 3 = owner SM pushes, 4 = owner copy-engine pushes, 5 = grouped ncclSend/ncclRecv;
# chunks:   launch chunks per matvec for the halo exchange (0 = the library's default, 8)
function ShardedOperator(ctx::Context, make::Function; exchange::Integer=0, chunks::Integer=0)
    oprs = Any[]
    for dev in ctx.devices
        check(ccall((:ed_set_device, libedcuda), Cint, (Cint,), dev))
        push!(oprs, make(dev))
    end
    S = eltype(oprs[1]); out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ed_sharded_create, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int32, Int32, Int32, Ref{Ptr{Cvoid}}),
                ctx.ptr, Ptr{Cvoid}[o.ptr for o in oprs], dtype_code(S), exchange, chunks, out))
    sh = ShardedOperator{S}(out[], ctx, oprs, size(oprs[1], 1))
    finalizer(x -> ccall((:ed_sharded_destroy, libedcuda), Cint, (Ptr{Cvoid},), x.ptr), sh)
    return sh
end

mutable struct DVector{S}
    ptr::Ptr{Cvoid}; sh::ShardedOperator{S}
    function DVector(sh::ShardedOperator{S}) where {S}
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ed_dvec_create, libedcuda), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), sh.ptr, out))
        v = new{S}(out[], sh)
        finalizer(x -> ccall((:ed_dvec_destroy, libedcuda), Cint, (Ptr{Cvoid},), x.ptr), v)
        return v
    end
end
upload!(v::DVector{S}, host::Vector{S}) where {S} =
    (length(host) == v.sh.dim || throw(DimensionMismatch("vector length differs from the dimension"));
     check(ccall((:ed_dvec_upload, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), v.ptr, host)); v)
download!(host::Vector{S}, v::DVector{S}) where {S} =
    (check(ccall((:ed_dvec_download, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), v.ptr, host)); host)

function LinearAlgebra.mul!(y::DVector{S}, sh::ShardedOperator{S}, x::DVector{S}) where {S}
    check(ccall((:ed_apply_sharded, libedcuda), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float64}), sh.ptr, y.ptr, x.ptr, 0, C_NULL))
    return y
end

function lanczos(sh::ShardedOperator{S}, nsteps::Integer; seed::Integer=0, nritz::Integer=4) where {S}
    alpha = zeros(nsteps); beta = zeros(nsteps); ritz = zeros(nritz); done = Ref{Int32}(0); ms = Ref{Float64}(0.0)
    check(ccall((:ed_lanczos_sharded, libedcuda), Cint,
                (Ptr{Cvoid}, Int32, UInt64, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Ref{Int32}, Ref{Float64}),
                sh.ptr, nsteps, seed, C_NULL, alpha, beta, ritz, nritz, done, ms))
    return (alpha=alpha[1:done[]], beta=beta[1:done[]], ritz=ritz, ms_per_step=ms[])
end

end # module

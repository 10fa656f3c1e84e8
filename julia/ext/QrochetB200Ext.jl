# QrochetB200Ext -- package extension for Qrochet.jl modelled on ext/QrochetAdaptExt.jl (:7-9 of the reference):
# moves Quantum / Product / Chain networks onto a B200 and dispatches the hot path to libqrochet_b200.so via ccall.
#
# NOT EXECUTED in this repository (no Julia toolchain in the build image).  It is the binding a maintainer adds, in the
# reference's own language; the executable mirror of every method below is qrochet.jl_b200/*.py (ctypes) and the same
# header is driven from plain C by tests/abi_smoke.c.  Header: include/qrochet_b200.h.
#
# Two levels, as in INTEGRATION.md:
#   1. Tenet level  -- `adapt(B200Array, ψ)` returns the SAME types the reference's extension returns (Quantum, Product,
#      Chain; QrochetAdaptExt.jl:7-9) with `B200Array` storage inside every `Tenet.Tensor`; the unmodified
#      src/Ansatz/Chain.jl then runs on the device through the methods of section 2 (contract, svd, qr, slice!, conj,
#      pinv(Diagonal), norm, ...), one ccall each.
#   2. Chain level  -- `B200Chain(ψ)` wraps a device-resident MPS in the library's private layout; canonize!,
#      mixed_canonize!, truncate!, evolve!, overlap, expect run as fused kernel chains (Chain.jl:460-752).
module QrochetB200Ext

using Qrochet
using Qrochet: Site, Dense, Open, id, sites, inputs, outputs, nsites, boundary, leftindex, rightindex
using Tenet
using Adapt
using LinearAlgebra

const lib = "libqrochet_b200"          # on LD_LIBRARY_PATH
# dtype enum of include/qrochet_b200.h.  ComplexF32 tensors are native float2 on the device (tcgen05 TF32-split
# contraction); Float32 Schmidt vectors are held widened to FP64 by the library.
const C128 = Int32(0); const C64 = Int32(1); const F64 = Int32(2); const F32 = Int32(3)
dtypecode(::Type{ComplexF64}) = C128; dtypecode(::Type{ComplexF32}) = C64
dtypecode(::Type{Float64}) = F64;     dtypecode(::Type{Float32}) = F32
const B200Elt = Union{ComplexF64,ComplexF32,Float64,Float32}
const PV = Ptr{Cvoid}

# ---- 0. context, errors ------------------------------------------------------------------------------------------
mutable struct Context
    h::PV
    function Context(device::Integer = 0)
        r = Ref{PV}(C_NULL)
        check(C_NULL, ccall((:qb200_create, lib), Int32, (Int32, Ref{PV}), device, r))
        finalizer(c -> ccall((:qb200_destroy, lib), Int32, (PV,), c.h), new(r[]))
    end
end
const CTX = Ref{Context}()
context() = isassigned(CTX) ? CTX[] : (CTX[] = Context())
ctxh() = context().h

function check(ctx, code::Int32)
    code == 0 && return nothing
    msg = unsafe_string(ccall((:qb200_last_error, lib), Cstring, (PV,), ctx))
    code == -1 && throw(ArgumentError(msg))                      # QB200_E_INVALID: Chain.jl:344,352,360,369,394,553-580
    code == -4 && throw(Qrochet.MissingSchmidtCoefficientsException(msg))   # QB200_E_NOSPECTRUM: Ansatz.jl:91-99
    error("libqrochet_b200 error $code: $msg")
end
check(code::Int32) = check(ctxh(), code)

# ---- 1. the device array type stored inside Tenet.Tensor{T,N,B200Array{T,N}} -----------------------------------------
mutable struct B200Array{T,N} <: AbstractArray{T,N}
    h::PV
    dims::NTuple{N,Int}
    function B200Array{T,N}(h::PV, dims::NTuple{N,Int}) where {T,N}
        # finalizers may run on any thread: qb200_tensor_free is stream-ordered (cudaFreeAsync) and never blocks
        finalizer(t -> ccall((:qb200_tensor_free, lib), Int32, (PV, PV), ctxh(), t.h), new{T,N}(h, dims))
    end
end
const B200Vector{T} = B200Array{T,1}
Base.size(a::B200Array) = a.dims
Base.IndexStyle(::Type{<:B200Array}) = IndexLinear()
# element access is a host read (truncate! reads the spectrum element-wise, Chain.jl:402,416): download once, index
Base.getindex(a::B200Array, i::Int) = Array(a)[i]

function B200Array{T,N}(::UndefInitializer, dims::NTuple{N,Int}) where {T<:B200Elt,N}
    r = Ref{PV}(C_NULL)
    check(ccall((:qb200_tensor_alloc, lib), Int32, (PV, Int32, Int32, Ptr{Int64}, Ref{PV}),
                ctxh(), dtypecode(T), N, collect(Int64, dims), r))
    B200Array{T,N}(r[], dims)
end
function B200Array(x::Array{T,N}) where {T<:B200Elt,N}
    a = B200Array{T,N}(undef, size(x))
    check(ccall((:qb200_tensor_upload, lib), Int32, (PV, PV, Ptr{T}), ctxh(), a.h, x))   # host pointer borrowed for the call
    a
end
function Base.Array(a::B200Array{T,N}) where {T,N}
    out = Array{T,N}(undef, a.dims)
    check(ccall((:qb200_tensor_download, lib), Int32, (PV, PV, Ptr{T}), ctxh(), a.h, out))   # synchronises
    out
end
Base.similar(a::B200Array{T}, ::Type{S}, dims::Dims{N}) where {T,S<:B200Elt,N} = B200Array{S,N}(undef, dims)
function Base.copy(a::B200Array{T,N}) where {T,N}                                          # Quantum.jl:87-90
    b = B200Array{T,N}(undef, a.dims)
    check(ccall((:qb200_tensor_copy, lib), Int32, (PV, PV, PV), ctxh(), a.h, b.h)); b
end
Base.zero(a::B200Array{T,N}) where {T,N} = B200Array(zeros(T, a.dims))
function Base.reshape(a::B200Array{T}, dims::Dims{N}) where {T,N}                           # Chain.jl:431,449
    b = copy(a)
    check(ccall((:qb200_tensor_reshape, lib), Int32, (PV, PV, Int32, Ptr{Int64}), ctxh(), b.h, N, collect(Int64, dims)))
    B200Array{T,N}(b.h, dims) |> x -> (b.h = C_NULL; x)      # the handle moves to the reshaped array
end
function Base.permutedims(a::B200Array{T,N}, perm) where {T,N}
    b = B200Array{T,N}(undef, ntuple(i -> a.dims[perm[i]], N))
    check(ccall((:qb200_permute, lib), Int32, (PV, PV, Ptr{Int32}, PV), ctxh(), a.h, Int32.(collect(perm) .- 1), b.h)); b
end
function Base.conj(a::B200Array{T,N}) where {T,N}                                           # Quantum.jl:105
    b = B200Array{T,N}(undef, a.dims)
    check(ccall((:qb200_conj, lib), Int32, (PV, PV, PV), ctxh(), a.h, b.h)); b
end
function LinearAlgebra.norm(a::B200Array)                                                   # Chain.jl:534,654
    r = Ref{Float64}(0)
    check(ccall((:qb200_norm2, lib), Int32, (PV, PV, Ref{Float64}), ctxh(), a.h, r)); r[]
end
function LinearAlgebra.rmul!(a::B200Array, s::Number)
    check(ccall((:qb200_scale, lib), Int32, (PV, PV, Ptr{Float64}), ctxh(), a.h, [real(s), imag(s)])); a
end
LinearAlgebra.normalize!(a::B200Array, p::Real = 2) = (p == 2 || throw(ArgumentError("only the 2-norm")); rmul!(a, 1 / norm(a)))
LinearAlgebra.normalize(a::B200Array, p::Real = 2) = normalize!(copy(a), p)
Base.isapprox(a::B200Array, b::AbstractArray; kw...) = isapprox(Array(a), b; kw...)           # Chain.jl:439,457
Base.isapprox(a::AbstractArray, b::B200Array; kw...) = isapprox(a, Array(b); kw...)
# slice!(tn, ind, 1:k) (Chain.jl:419) reaches selectdim / view; view(tn, ind => v) drops the index (distributed.jl:72,82)
function Base.selectdim(a::B200Array{T,N}, d::Integer, r::AbstractUnitRange) where {T,N}
    first(r) == 1 || throw(ArgumentError("device slices keep a prefix (truncate! keeps the largest Schmidt values)"))
    b = B200Array{T,N}(undef, ntuple(i -> i == d ? length(r) : a.dims[i], N))
    check(ccall((:qb200_slice_mode, lib), Int32, (PV, PV, Int32, Int64, PV), ctxh(), a.h, d - 1, length(r), b.h)); b
end
function Base.selectdim(a::B200Array{T,N}, d::Integer, i::Integer) where {T,N}
    b = B200Array{T,N - 1}(undef, ntuple(k -> a.dims[k < d ? k : k + 1], N - 1))
    check(ccall((:qb200_select_mode, lib), Int32, (PV, PV, Int32, Int64, PV), ctxh(), a.h, d - 1, i - 1, b.h)); b
end

# Adapt hooks: storage goes to the device and back; structure keeps the reference's types (QrochetAdaptExt.jl:7-9)
Adapt.adapt_storage(::Type{B200Array}, x::Array{<:B200Elt}) = B200Array(x)
Adapt.adapt_storage(::Type{Array}, x::B200Array) = Array(x)
Adapt.adapt_structure(to::Type{B200Array}, x::Quantum) = Quantum(adapt(to, TensorNetwork(x)), x.sites)
Adapt.adapt_structure(to::Type{B200Array}, x::Product) = Product(adapt(to, Quantum(x)))
Adapt.adapt_structure(to::Type{B200Array}, x::Chain) = Chain(adapt(to, Quantum(x)), boundary(x))
# the reference has no adapt method for gates (Dense): add it so that evolve! can take device gates
Adapt.adapt_structure(to::Type{B200Array}, x::Dense) = Dense(adapt(to, Quantum(x)))

# ---- 2. Tenet level: the arithmetic of Chain.jl on device tensors ---------------------------------------------------
const DevTensor{T,N} = Tensor{T,N,<:B200Array}
handle(t::Tensor) = parent(t).h
modeids!(table, is) = Int32[get!(table, i, Int32(length(table))) for i in is]

# contract(a, b; dims): sums dims ∩ inds(a) ∩ inds(b) (default: all shared); dims = () keeps shared indices
# (element-wise, the Λ absorb of Chain.jl:322,325,484,491,710,713).  Output order (inds(a) ∪ inds(b)) ∖ dims.
function Tenet.contract(a::DevTensor{T}, b::DevTensor{T}; dims = (∩(inds(a), inds(b)))) where {T}
    ia, ib = collect(inds(a)), collect(inds(b))
    summed = [i for i in dims if i ∈ ia && i ∈ ib]
    ic = [i for i in vcat(ia, [j for j in ib if j ∉ ia]) if i ∉ summed]
    c = B200Array{T,length(ic)}(undef, ntuple(k -> ic[k] ∈ ia ? size(a, ic[k]) : size(b, ic[k]), length(ic)))
    table = Dict{Symbol,Int32}()
    check(ccall((:qb200_contract, lib), Int32,
        (PV, PV, Ptr{Int32}, Int32, PV, Ptr{Int32}, Int32, PV, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
        ctxh(), handle(a), modeids!(table, ia), 0, handle(b), modeids!(table, ib), 0, c.h, modeids!(table, ic),
        [1.0, 0.0], [0.0, 0.0]))
    Tensor(c, ic)
end
# a real Schmidt vector against a complex site: the mode-scale kernel (K2/K3), HBM-bound, in one pass
function Tenet.contract(a::DevTensor{T}, λ::Tensor{<:Real,1,<:B200Array}; dims = ()) where {T<:Complex}
    isempty(dims) || return invoke(Tenet.contract, Tuple{Tensor,Tensor}, a, λ; dims)
    pos = findfirst(==(only(inds(λ))), collect(inds(a)))
    c = B200Array{T,ndims(a)}(undef, size(parent(a)))
    check(ccall((:qb200_scale_mode, lib), Int32, (PV, PV, Int32, PV, Int32, Float64, PV),
                ctxh(), handle(a), pos - 1, handle(λ), 0, 0.0, c.h))
    Tensor(c, inds(a))
end
Tenet.contract(λ::Tensor{<:Real,1,<:B200Array}, a::DevTensor{T}; dims = ()) where {T<:Complex} =
    isempty(dims) ? permutedims(Tenet.contract(a, λ; dims), vcat(only(inds(λ)), [i for i in inds(a) if i != only(inds(λ))])) :
                    invoke(Tenet.contract, Tuple{Tensor,Tensor}, λ, a; dims)
# diag(pinv(Diagonal(λ); atol)) (Chain.jl:491,710,713): the reciprocal with the reference's cut-off, on the host mirror
# (λ has at most 2χ entries and is read element-wise by truncate! anyway)
LinearAlgebra.Diagonal(v::B200Vector) = Diagonal(Array(v))
devpinv(λ::Tensor; atol) = Tensor(B200Array(diag(pinv(Diagonal(Array(parent(λ))); atol))), inds(λ))

function matricise(t::DevTensor, left_inds, right_inds)
    is = collect(inds(t))
    order = Int32[findfirst(==(i), is) - 1 for i in vcat(left_inds, right_inds)]
    rows = prod(size(t, i) for i in left_inds; init = 1); cols = prod(size(t, i) for i in right_inds; init = 1)
    order, rows, cols
end
# LinearAlgebra.svd(::Tensor; left_inds, right_inds, virtualind) (Chain.jl:365,645,705): U[left..., v], s[v] descending,
# Vt[right..., v] = conj(V) -- qb200_svd (QR-preconditioned one-sided block Jacobi); `maxdim` / `threshold` apply the
# truncate! rule on the spot (Chain.jl:404-417) and are optional extras of the device method.
function LinearAlgebra.svd(t::DevTensor{T}; left_inds, right_inds = [i for i in inds(t) if i ∉ left_inds],
                           virtualind = Symbol(Tenet.letter(Qrochet.nextindex())), maxdim = 0, threshold = -1.0) where {T}
    order, rows, cols = matricise(t, left_inds, right_inds)
    k = min(rows, cols)
    U = B200Array{T,length(left_inds) + 1}(undef, (map(i -> size(t, i), left_inds)..., k))
    Vt = B200Array{T,length(right_inds) + 1}(undef, (map(i -> size(t, i), right_inds)..., k))
    s = B200Array{real(T),1}(undef, (k,))
    kept = Ref{Int64}(0); dw = Ref{Float64}(0)
    check(ccall((:qb200_svd, lib), Int32,
        (PV, PV, Ptr{Int32}, Int32, Int64, Float64, PV, PV, PV, Ref{Int64}, Ref{Float64}),
        ctxh(), handle(t), order, length(left_inds), maxdim, threshold, U.h, s.h, Vt.h, kept, dw))
    if kept[] != k            # the library shrank the bond extent of its outputs
        U = B200Array{T,ndims(U)}(U.h, (U.dims[1:end-1]..., kept[])) |> x -> (U.h = C_NULL; x)
        Vt = B200Array{T,ndims(Vt)}(Vt.h, (Vt.dims[1:end-1]..., kept[])) |> x -> (Vt.h = C_NULL; x)
        s = B200Array{real(T),1}(s.h, (kept[],)) |> x -> (s.h = C_NULL; x)
    end
    Tensor(U, [left_inds..., virtualind]), Tensor(s, [virtualind]), Tensor(Vt, [right_inds..., virtualind])
end
# LinearAlgebra.qr(::Tensor; ...) (Chain.jl:367): Q[left..., v], R[v, right...] thin -- qb200_qr
function LinearAlgebra.qr(t::DevTensor{T}; left_inds, right_inds = [i for i in inds(t) if i ∉ left_inds],
                          virtualind = Symbol(Tenet.letter(Qrochet.nextindex()))) where {T}
    order, rows, cols = matricise(t, left_inds, right_inds)
    k = min(rows, cols)
    Q = B200Array{T,length(left_inds) + 1}(undef, (map(i -> size(t, i), left_inds)..., k))
    R = B200Array{T,length(right_inds) + 1}(undef, (k, map(i -> size(t, i), right_inds)...))
    check(ccall((:qb200_qr, lib), Int32, (PV, PV, Ptr{Int32}, Int32, PV, PV), ctxh(), handle(t), order, length(left_inds), Q.h, R.h))
    Tensor(Q, [left_inds..., virtualind]), Tensor(R, [virtualind, right_inds...])
end

# ---- 3. Chain level: the fused path ---------------------------------------------------------------------------------
# `Chain` is not parametric on the array type (Chain.jl:6-9), so the fused methods live on a thin wrapper that `adapt`
# does NOT create silently: B200Chain(ψ) uploads a host or device Chain into the library's private (l, o, r) layout,
# Chain(ψ::B200Chain) brings it back with its Schmidt vectors on hyperindices, exactly as canonize! leaves them.
mutable struct B200Chain <: Qrochet.Ansatz
    h::PV
    n::Int
    function B200Chain(h::PV, n::Int)
        finalizer(c -> ccall((:qb200_mps_free, lib), Int32, (PV, PV), ctxh(), c.h), new(h, n))
    end
end
Qrochet.nsites(ψ::B200Chain) = ψ.n
Qrochet.boundary(::B200Chain) = Open()

hostarray(t::Tensor) = parent(t) isa B200Array ? Array(parent(t)) : Array(parent(t))
function B200Chain(ψ::Chain)
    boundary(ψ) isa Open || throw(ArgumentError("the fused path covers open-boundary states"))
    n = nsites(ψ); r = Ref{PV}(C_NULL)
    check(ccall((:qb200_mps_create, lib), Int32, (PV, Int32, Ref{PV}), ctxh(), n, r))
    φ = B200Chain(r[], n)
    nλ = 0
    for i in 1:n
        t = tensors(ψ; at = Site(i))
        order = filter(!isnothing, [leftindex(ψ, Site(i)), inds(ψ; at = Site(i)), rightindex(ψ, Site(i))])
        a = ComplexF64.(permutedims(hostarray(t), [findfirst(==(j), collect(inds(t))) for j in order]))   # -> (l, o, r)
        χl = i == 1 ? 1 : size(a, 1); χr = i == n ? 1 : size(a, ndims(a))
        check(ccall((:qb200_mps_set_site, lib), Int32, (PV, PV, Int32, Int64, Int64, Int64, Ptr{ComplexF64}),
                    ctxh(), φ.h, i - 1, χl, length(a) ÷ (χl * χr), χr, a))
        if i < n                      # Schmidt vector on the bond (a hyperindex tensor; Ansatz.jl:78-89)
            λ = tensors(ψ; between = (Site(i), Site(i + 1)))
            if !isnothing(λ)
                v = Float64.(hostarray(λ))
                check(ccall((:qb200_mps_set_lambda, lib), Int32, (PV, PV, Int32, Int64, Ptr{Float64}), ctxh(), φ.h, i - 1, length(v), v))
                nλ += 1
            end
        end
    end
    # every bond carries its Schmidt vector <=> the chain came out of canonize! (Vidal form)
    check(ccall((:qb200_mps_set_form, lib), Int32, (PV, Int32), φ.h, nλ == n - 1 ? 1 : 0))
    φ
end
Adapt.adapt_structure(::Type{B200Chain}, ψ::Chain) = B200Chain(ψ)

function site_array(ψ::B200Chain, i::Int)
    d = zeros(Int64, 3)
    check(ccall((:qb200_mps_site_dims, lib), Int32, (PV, Int32, Ptr{Int64}), ψ.h, i - 1, d))
    a = Array{ComplexF64,3}(undef, d...)
    check(ccall((:qb200_mps_get_site, lib), Int32, (PV, PV, Int32, Ptr{ComplexF64}), ctxh(), ψ.h, i - 1, a)); a
end
function schmidt(ψ::B200Chain, b::Int)          # tensors(ψ; between = (Site(b), Site(b+1)))
    n = Ref{Int64}(0)
    code = ccall((:qb200_mps_get_lambda, lib), Int32, (PV, PV, Int32, Ptr{Float64}, Ref{Int64}), ctxh(), ψ.h, b - 1, C_NULL, n)
    code == -4 && return nothing
    check(code)
    v = zeros(n[])
    check(ccall((:qb200_mps_get_lambda, lib), Int32, (PV, PV, Int32, Ptr{Float64}, Ref{Int64}), ctxh(), ψ.h, b - 1, v, n)); v
end
# back to a host Chain: sites in the reference's default order (o, l, r) (Chain.jl:33), Λ pushed on its bond index
function Qrochet.Chain(ψ::B200Chain)
    n = ψ.n
    arrays = map(1:n) do i
        a = permutedims(site_array(ψ, i), (2, 1, 3))
        i == 1 ? a[:, 1, :] : (i == n ? a[:, :, 1] : a)
    end
    χ = Chain(Qrochet.State(), Open(), arrays)
    for b in 1:n-1
        λ = schmidt(ψ, b)
        isnothing(λ) || push!(TensorNetwork(χ), Tensor(λ, [inds(χ; bond = (Site(b), Site(b + 1)))]))
    end
    χ
end

Qrochet.canonize!(ψ::B200Chain) = (check(ccall((:qb200_mps_canonize, lib), Int32, (PV, PV), ctxh(), ψ.h)); ψ)
Qrochet.mixed_canonize!(ψ::B200Chain, c::Site) =
    (check(ccall((:qb200_mps_mixed_canonize, lib), Int32, (PV, PV, Int32), ctxh(), ψ.h, id(c) - 1)); ψ)
function LinearAlgebra.normalize!(ψ::B200Chain, root::Site)                                   # Chain.jl:532-536
    Qrochet.mixed_canonize!(ψ, root)
    λ = schmidt(ψ, id(root) - 1); λ ./= norm(λ)
    check(ccall((:qb200_mps_set_lambda, lib), Int32, (PV, PV, Int32, Int64, Ptr{Float64}), ctxh(), ψ.h, id(root) - 2, length(λ), λ)); ψ
end
function Qrochet.truncate!(ψ::B200Chain, bond; threshold = nothing, maxdim = nothing)           # Chain.jl:390-422
    kept = Ref{Int64}(0)
    check(ccall((:qb200_mps_truncate, lib), Int32, (PV, PV, Int32, Int64, Float64, Ref{Int64}),
                ctxh(), ψ.h, minimum(id, bond) - 1, something(maxdim, 0), something(threshold, -1.0), kept)); ψ
end

# gate array of a Dense operator with its dims ordered (o_lanes..., i_lanes...), lanes ascending: the index order of
# the gate tensor is whatever the caller built (Dense.jl:21-34 only fixes the site -> index map), so permute by that map
function gate_array(gate::Dense, lanes)
    t = only(tensors(gate))
    want = vcat([inds(gate; at = Site(l)) for l in lanes], [inds(gate; at = Site(l; dual = true)) for l in lanes])
    ComplexF64.(permutedims(hostarray(t), [findfirst(==(i), collect(inds(t))) for i in want]))
end
function check_gate(ψ::B200Chain, gate::Dense)                                                   # Chain.jl:553-580
    Qrochet.socket(gate) isa Qrochet.Operator || throw(ArgumentError("Gate must be an operator, but got $(Qrochet.socket(gate))"))
    issetequal(adjoint.(inputs(gate)), outputs(gate)) || throw(ArgumentError("Gate inputs ($(inputs(gate))) and outputs ($(outputs(gate))) must be the same"))
    lanes = sort!(collect(id.(outputs(gate))))
    all(l -> 1 <= l <= ψ.n, lanes) || throw(ArgumentError("Gate inputs ($(inputs(gate))) must be a subset of the TN sites"))
    length(lanes) <= 2 || throw(ArgumentError("Invalid number of lanes $(length(lanes)), maximum is 2"))
    length(lanes) == 1 || lanes[2] == lanes[1] + 1 || throw(ArgumentError("Gate lanes must be contiguous"))
    lanes
end
# evolve!(ψ, gate; threshold, maxdim, iscanonical = false, renormalize = false): same keywords and defaults as
# Chain.jl:543-550; the Vidal branch (contract_2sitewf! / unpack_2sitewf!) runs as one fused chain on the device
function Qrochet.evolve!(ψ::B200Chain, gate::Dense; threshold = nothing, maxdim = nothing, iscanonical = false, renormalize = false)
    lanes = check_gate(ψ, gate)
    g = gate_array(gate, lanes)
    if length(lanes) == 1
        check(ccall((:qb200_mps_evolve1, lib), Int32, (PV, PV, Int32, Ptr{ComplexF64}), ctxh(), ψ.h, lanes[1] - 1, g))
    else
        kept = Ref{Int64}(0); dw = Ref{Float64}(0)
        check(ccall((:qb200_mps_evolve2, lib), Int32,
            (PV, PV, Int32, Ptr{ComplexF64}, Int64, Float64, Int32, Int32, Ref{Int64}, Ref{Float64}),
            ctxh(), ψ.h, lanes[1] - 1, g, something(maxdim, 0), something(threshold, -1.0), renormalize, iscanonical, kept, dw))
    end
    ψ
end
# a gate list in program order (the loop `for g in gates evolve!(ψ, g; ...)`): one call, dependency-scheduled on the
# device; identical results to the loop (consecutive TEBD layers overlap)
function Qrochet.evolve!(ψ::B200Chain, gates::AbstractVector{<:Dense}; threshold = nothing, maxdim = nothing, iscanonical = false, renormalize = false)
    bonds = Int32[]; flat = ComplexF64[]
    for gate in gates
        lanes = check_gate(ψ, gate)
        length(lanes) == 2 || throw(ArgumentError("gate lists hold two-lane gates; apply one-lane gates with evolve!(ψ, gate)"))
        push!(bonds, lanes[1] - 1); append!(flat, vec(gate_array(gate, lanes)))
    end
    kept = zeros(Int64, length(bonds)); dw = zeros(Float64, length(bonds))
    check(ccall((:qb200_mps_evolve2_circuit, lib), Int32,
        (PV, PV, Int32, Ptr{Int32}, Ptr{ComplexF64}, Int64, Float64, Int32, Int32, Ptr{Int64}, Ptr{Float64}),
        ctxh(), ψ.h, length(bonds), bonds, flat, something(maxdim, 0), something(threshold, -1.0), renormalize, iscanonical, kept, dw))
    ψ
end
function Qrochet.overlap(a::B200Chain, b::B200Chain)                                           # Chain.jl:737-748: <b|a>
    a.n == b.n || throw(ArgumentError("Ansatzes must have the same sites"))
    r = zeros(2)
    check(ccall((:qb200_mps_overlap, lib), Int32, (PV, PV, PV, Ptr{Float64}), ctxh(), a.h, b.h, r))
    complex(r[1], r[2])
end
LinearAlgebra.norm(ψ::B200Chain) = sqrt(abs(Qrochet.overlap(ψ, ψ)))                            # Ansatz.jl:101-109
# expect(ψ, observables) (Chain.jl:724-735), any list of 1- and 2-lane observables: the reference's own composition
# (copy, evolve! each observable, contract with ψ') on the device
function Qrochet.expect(ψ::B200Chain, observables)
    nl = Int32[]; left = Int32[]; flat = ComplexF64[]
    for o in observables
        lanes = check_gate(ψ, o)
        push!(nl, length(lanes)); push!(left, lanes[1] - 1); append!(flat, vec(gate_array(o, lanes)))
    end
    r = zeros(2)
    check(ccall((:qb200_mps_expect, lib), Int32, (PV, PV, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{ComplexF64}, Ptr{Float64}),
                ctxh(), ψ.h, length(nl), nl, left, flat, r))
    complex(r[1], r[2])
end
# many independent single-site expectation values at once: all left / right environments are built once and shared
function expect_batch(ψ::B200Chain, ops::Vector{<:AbstractMatrix}, at::Vector{<:Site})
    flat = ComplexF64[]; foreach(o -> append!(flat, vec(ComplexF64.(o))), ops)
    r = zeros(2 * length(ops))
    check(ccall((:qb200_mps_expect1_batch, lib), Int32, (PV, PV, Int32, Ptr{Int32}, Ptr{ComplexF64}, Ptr{Float64}),
                ctxh(), ψ.h, length(ops), Int32.(id.(at) .- 1), flat, r))
    complex.(r[1:2:end], r[2:2:end])
end

# ---- 4. sliced contraction of a circuit network (examples/distributed.jl:29-101) --------------------------------------
# plan (ContractSimplification + path search + findslices, deterministic) and per-slice replay inside the library;
# ranks take slices rank, rank + nranks, ...; one NCCL sum replaces sum(fetch.(partial_results)) (:101).
function contract_sliced(tn::TensorNetwork; size = 2^24, rank = 0, nranks = 1)
    ts = tensors(tn); table = Dict{Symbol,Int32}()
    ranks = Int32[ndims(t) for t in ts]
    modes = reduce(vcat, [modeids!(table, inds(t)) for t in ts]; init = Int32[])
    exts = reduce(vcat, [collect(Int64, Base.size(parent(t))) for t in ts]; init = Int64[])
    plan = Ref{PV}(C_NULL)
    check(ccall((:qb200_tn_plan, lib), Int32, (PV, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}, Int64, Ref{PV}),
                ctxh(), length(ts), ranks, modes, exts, size, plan))
    leaves = [parent(t) isa B200Array ? parent(t) : B200Array(ComplexF64.(parent(t))) for t in ts]
    acc = zeros(2)
    GC.@preserve leaves begin
        check(ccall((:qb200_tn_contract_sliced, lib), Int32, (PV, PV, Ptr{PV}, Int64, Int64, Ptr{Float64}),
                    ctxh(), plan[], PV[l.h for l in leaves], rank, nranks, acc))
    end
    ccall((:qb200_tn_plan_free, lib), Int32, (PV, PV), ctxh(), plan[])
    nranks > 1 && check(ccall((:qb200_comm_allreduce_sum, lib), Int32, (PV, Ptr{Float64}, Int32), ctxh(), acc, 2))
    complex(acc[1], acc[2])
end
# one process per GPU: `id` = the 128 bytes of qb200_comm_unique_id broadcast by rank 0 (MPI / Distributed / a file)
comm_unique_id() = (buf = zeros(UInt8, 128); check(C_NULL, ccall((:qb200_comm_unique_id, lib), Int32, (Ptr{UInt8},), buf)); buf)
comm_init(nranks, rank, id::Vector{UInt8}) = check(ccall((:qb200_comm_init, lib), Int32, (PV, Int32, Int32, Ptr{UInt8}), ctxh(), nranks, rank, id))
# batched expectation values on several GPUs: replicate the state once over NVLink, deal the observables i mod W
function broadcast!(ψ::Union{B200Chain,Nothing}, root::Integer = 0)
    r = Ref{PV}(isnothing(ψ) ? C_NULL : ψ.h)
    check(ccall((:qb200_mps_broadcast, lib), Int32, (PV, Ref{PV}, Int32), ctxh(), r, root))
    isnothing(ψ) || return ψ
    d = zeros(Int64, 3); n = 0
    while ccall((:qb200_mps_site_dims, lib), Int32, (PV, Int32, Ptr{Int64}), r[], n, d) == 0
        n += 1
    end
    B200Chain(r[], n)
end

end

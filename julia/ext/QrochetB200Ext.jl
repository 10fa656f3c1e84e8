# QrochetB200Ext -- package extension for Qrochet.jl modelled on ext/QrochetAdaptExt.jl (:7-9 of the reference):
# moves a Chain onto a B200 and dispatches the hot path to libqrochet_b200.so through ccall.
#
# NOT EXECUTED in this repository (no Julia toolchain in the build image); it documents, in the reference's own
# language, the bindings a maintainer adds.  The executable mirror of this file is qrochet.jl_b200/*.py (ctypes).
module QrochetB200Ext

using Qrochet
using Tenet
using Adapt
using LinearAlgebra

const lib = "libqrochet_b200"          # on LD_LIBRARY_PATH
const C128 = Int32(0); const C64 = Int32(1); const F64 = Int32(2); const F32 = Int32(3)
# dtype enum of include/qrochet_b200.h.  ComplexF32 tensors are native float2 on the device (TF32-split tensor-core
# contraction); Float32 Schmidt vectors are held in FP64.
dtypecode(::Type{ComplexF64}) = C128; dtypecode(::Type{ComplexF32}) = C64
dtypecode(::Type{Float64}) = F64;     dtypecode(::Type{Float32}) = F32

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(C_NULL, ccall((:qb200_create, lib), Int32, (Int32, Ref{Ptr{Cvoid}}), device, r))
        ctx = new(r[])
        finalizer(c -> ccall((:qb200_destroy, lib), Int32, (Ptr{Cvoid},), c.h), ctx)
    end
end
const CTX = Ref{Context}()
context() = isassigned(CTX) ? CTX[] : (CTX[] = Context())

function check(ctx, code::Int32)
    code == 0 && return
    msg = unsafe_string(ccall((:qb200_last_error, lib), Cstring, (Ptr{Cvoid},), ctx))
    code == -1 && throw(ArgumentError(msg))                                     # Chain.jl:344,352,394,553-580
    code == -4 && throw(Qrochet.MissingSchmidtCoefficientsException((site"1", site"2")))  # Ansatz.jl:91-99
    error("libqrochet_b200 error $code: $msg")
end

# ---- the device array type stored inside Tenet.Tensor{T,N,B200Array{T,N}} -------------------------------------
mutable struct B200Array{T,N} <: AbstractArray{T,N}
    h::Ptr{Cvoid}
    dims::NTuple{N,Int}
end
Base.size(a::B200Array) = a.dims
const B200Elt = Union{ComplexF64,ComplexF32,Float64,Float32}
function B200Array(x::Array{T,N}) where {T<:B200Elt,N}
    ctx = context(); r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx.h, ccall((:qb200_tensor_alloc, lib), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int64}, Ref{Ptr{Cvoid}}),
                       ctx.h, dtypecode(T), N, collect(Int64, size(x)), r))
    check(ctx.h, ccall((:qb200_tensor_upload, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), ctx.h, r[], x))
    a = B200Array{T,N}(r[], size(x))
    finalizer(t -> ccall((:qb200_tensor_free, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), context().h, t.h), a)  # stream-ordered, never blocks
end
function Base.Array(a::B200Array{T,N}) where {T,N}
    out = Array{T,N}(undef, a.dims)
    check(context().h, ccall((:qb200_tensor_download, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), context().h, a.h, out))
    out
end
Adapt.adapt_storage(::Type{B200Array}, x::Array{<:B200Elt}) = B200Array(x)
Adapt.adapt_storage(::Type{Array}, x::B200Array) = Array(x)
# the reference has no adapt method for gates (Dense): add it so evolve! can take device gates
Adapt.adapt_structure(to, x::Qrochet.Dense) = Qrochet.Dense(adapt(to, Quantum(x)))

# ---- Tenet level: contract / svd / qr on device tensors (call sites Chain.jl:322-325, 365-372, 484-491, 705-713) ----
modeids(inds, table) = Int32[get!(table, i, Int32(length(table))) for i in inds]
function Tenet.contract(a::Tensor{T,N,<:B200Array}, b::Tensor{T,M,<:B200Array}; dims = (∩(inds(a), inds(b)))) where {T,N,M}
    table = Dict{Symbol,Int32}()
    ic = [i for i in vcat(collect(inds(a)), [j for j in inds(b) if j ∉ inds(a)]) if i ∉ dims]
    c = B200Array(Array{T}(undef, (i -> i ∈ inds(a) ? size(a, i) : size(b, i)).(ic)...))   # T = ComplexF64 or ComplexF32 (same in, same out)
    check(context().h, ccall((:qb200_contract, lib), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Int32, Ptr{Cvoid}, Ptr{Int32}, Int32, Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
        context().h, parent(a).h, modeids(inds(a), table), 0, parent(b).h, modeids(inds(b), table), 0, c.h, modeids(ic, table),
        [1.0, 0.0], [0.0, 0.0]))
    Tensor(c, ic)
end
# LinearAlgebra.svd(::Tensor{…B200Array}; left_inds, right_inds, virtualind) -> qb200_svd   (Chain.jl:365,645,705)
# LinearAlgebra.qr(::Tensor{…B200Array};  left_inds, right_inds, virtualind) -> qb200_qr    (Chain.jl:367)
# contract(x, Λ; dims=()) and pinv(Diagonal(λ); atol)                        -> qb200_scale_mode
# slice!(tn, ind, 1:k) / view(tn, ind => v) / conj / norm                    -> qb200_slice_mode / _select_mode / _conj / _norm2
# (same pattern as `contract` above: build the mode-position list, allocate outputs, one ccall.)

# ---- Chain level: the fused path.  Chain is not parametric on the array type (Chain.jl:6-9), so the fast methods
# ---- are reached through a thin wrapper created by adapt. -------------------------------------------------------
mutable struct B200Chain
    h::Ptr{Cvoid}
    sites::Dict{Site,Symbol}            # the Quantum site map stays on the Julia side (Quantum.jl:54-76)
end
function Adapt.adapt_structure(::Type{B200Array}, ψ::Chain)
    ctx = context(); n = nsites(ψ); r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx.h, ccall((:qb200_mps_create, lib), Int32, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), ctx.h, n, r))
    for i in 1:n
        t = tensors(ψ; at = Site(i))
        order = filter(!isnothing, [Qrochet.leftindex(ψ, Site(i)), inds(ψ; at = Site(i)), Qrochet.rightindex(ψ, Site(i))])
        a = ComplexF64.(permutedims(parent(t), [findfirst(==(j), inds(t)) for j in order]))   # -> (l, o, r)
        χl = i == 1 ? 1 : size(a, 1); χr = i == n ? 1 : size(a, ndims(a))
        check(ctx.h, ccall((:qb200_mps_set_site, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64, Int64, Ptr{Cvoid}),
                           ctx.h, r[], i - 1, χl, length(a) ÷ (χl * χr), χr, a))
    end
    B200Chain(r[], Quantum(ψ).sites)
end
Qrochet.canonize!(ψ::B200Chain) = (check(context().h, ccall((:qb200_mps_canonize, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), context().h, ψ.h)); ψ)
Qrochet.mixed_canonize!(ψ::B200Chain, c::Site) = (check(context().h, ccall((:qb200_mps_mixed_canonize, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32), context().h, ψ.h, id(c) - 1)); ψ)
function Qrochet.truncate!(ψ::B200Chain, bond; threshold = nothing, maxdim = nothing)
    kept = Ref{Int64}(0)
    check(context().h, ccall((:qb200_mps_truncate, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Float64, Ref{Int64}),
                             context().h, ψ.h, id(bond[1]) - 1, something(maxdim, 0), something(threshold, -1.0), kept)); ψ
end
function Qrochet.evolve!(ψ::B200Chain, gate::Qrochet.Dense; threshold = nothing, maxdim = nothing, iscanonical = true, renormalize = false)
    lanes = sort!(id.(outputs(gate)))
    g = ComplexF64.(Array(parent(only(tensors(gate)))))          # dims (o1,o2,i1,i2) for sites [l, l+1, l', (l+1)']
    if length(lanes) == 1
        check(context().h, ccall((:qb200_mps_evolve1, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}), context().h, ψ.h, lanes[1] - 1, g))
    else
        lanes[2] == lanes[1] + 1 || throw(ArgumentError("Gate lanes must be contiguous"))           # Chain.jl:574
        kept = Ref{Int64}(0); dw = Ref{Float64}(0)
        check(context().h, ccall((:qb200_mps_evolve2, lib), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int64, Float64, Int32, Ref{Int64}, Ref{Float64}),
            context().h, ψ.h, lanes[1] - 1, g, something(maxdim, 0), something(threshold, -1.0), renormalize, kept, dw))
    end
    ψ
end
# a gate list in program order (the loop `for g in gates evolve!(ψ, g; ...)`): one call, dependency-scheduled on the device
function Qrochet.evolve!(ψ::B200Chain, gates::AbstractVector{<:Qrochet.Dense}; threshold = nothing, maxdim = nothing, iscanonical = true, renormalize = false)
    bonds = Int32[]; flat = ComplexF64[]
    for gate in gates
        lanes = sort!(id.(outputs(gate)))
        length(lanes) == 2 && lanes[2] == lanes[1] + 1 || throw(ArgumentError("Gate lanes must be contiguous"))   # Chain.jl:574
        push!(bonds, lanes[1] - 1); append!(flat, vec(ComplexF64.(Array(parent(only(tensors(gate)))))))
    end
    kept = zeros(Int64, length(bonds)); dw = zeros(Float64, length(bonds))
    check(context().h, ccall((:qb200_mps_evolve2_circuit, lib), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Cvoid}, Int64, Float64, Int32, Ptr{Int64}, Ptr{Float64}),
        context().h, ψ.h, length(bonds), bonds, flat, something(maxdim, 0), something(threshold, -1.0), renormalize, kept, dw))
    ψ
end
function Qrochet.overlap(a::B200Chain, b::B200Chain)
    r = zeros(2)
    check(context().h, ccall((:qb200_mps_overlap, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}), context().h, a.h, b.h, r))
    complex(r[1], r[2])
end
LinearAlgebra.norm(ψ::B200Chain) = sqrt(abs(Qrochet.overlap(ψ, ψ)))
# expect(ψ, observables) -> qb200_mps_expect1_batch ; sliced contraction -> qb200_tn_plan / qb200_tn_contract_sliced
# + qb200_comm_init / qb200_comm_allreduce_sum replacing Distributed.@spawnat / fetch / sum (examples/distributed.jl:66-101)

end

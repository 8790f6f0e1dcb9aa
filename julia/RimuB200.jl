# RimuB200.jl -- Julia shim that puts librimu_b200.so under Rimu.jl (v0.14).
#
# It replaces exactly one thing in Rimu: the vector + working-memory types whose `apply_operator!` method performs the
# FCIQMC step (Interfaces/dictvectors.jl:90-140, DictVectors/pdworkingmemory.jl:297-309).  `ProjectorMonteCarloProblem`,
# `solve`, the shift / reporting / post-step strategies and StatsTools keep running unchanged, because the driver only
# talks to the vector through the `AbstractDVec` interface (Interfaces/dictvectors.jl:22-56).
#
# Every `ccall` below binds one entry point of include/rimu_b200.h; tests/test_julia_shim.py parses this file and checks
# each symbol and argument count against that header, the field counts of the three mirrored structs against the C structs,
# and that no helper is used without being defined.  Julia is not available in the build image, so the file is checked
# structurally there; the identical C ABI is exercised by the ctypes mirror rimu.jl_b200/_lib.py.
#
# Usage:
#     using Rimu, RimuB200
#     H   = HubbardMom1D(BoseFS((0,0,0,6,0,0,0,0)); u=4.0)
#     ctx = RimuB200.Context(RimuB200.words(typeof(starting_address(H))))
#     v   = RimuB200.GPUDVec(starting_address(H) => 10.0; style=IsDynamicSemistochastic(), ctx)
#     p   = ProjectorMonteCarloProblem(H; start_at=v, target_walkers=10_000)
#     solve(p)
module RimuB200

using Rimu
using Rimu.BitStringAddresses: BoseFS, FermiFS, CompositeFS, SingleComponentFockAddress, AbstractFockAddress,
    num_modes, num_particles, onr
using Rimu.StochasticStyles: IsDeterministic, IsStochasticInteger, IsDynamicSemistochastic, IsStochasticWithThreshold,
    ThresholdCompression, NoCompression
using Rimu.DictVectors: Initiator, SimpleInitiator, CoherentInitiator, NonInitiator, InitiatorRule, FrozenDVec, DVec
using Rimu.Hamiltonians: HubbardReal1D, HubbardReal1DEP, ExtendedHubbardReal1D, HubbardMom1D, HubbardMom1DEP, ExtendedHubbardMom1D, HubbardRealSpace,
    Transcorrelated1D, AbstractHamiltonian
using Rimu.Interfaces: AbstractDVec, StochasticStyle
import Rimu.Interfaces: apply_operator!, working_memory, localpart
import Rimu.DictVectors: walkernumber, walkernumber_and_length, freeze
import VectorInterface: zerovector, zerovector!, scale!, add!
import LinearAlgebra: dot, norm, mul!

const LIB = get(ENV, "RIMU_B200_LIB", joinpath(@__DIR__, "..", "rimu.jl_b200", "librimu_b200.so"))

# ---------------------------------------------------------------------------------------------------------------- status
const RIMU_OK, RIMU_ERR_TABLE_FULL, RIMU_ERR_VECTOR_FULL, RIMU_ERR_EXCHANGE_FULL, RIMU_ERR_WORKMEM = 0, 1, 2, 3, 4
const RIMU_ERR_INVALID = -1

struct RimuB200Error <: Exception
    status::Cint
    msg::String
end
Base.showerror(io::IO, e::RimuB200Error) = print(io, "RimuB200Error(status $(e.status)): $(e.msg)")

last_error() = unsafe_string(ccall((:rimu_last_error, LIB), Cstring, ()))
function check(status::Integer)
    status == RIMU_OK && return nothing
    msg = last_error()
    status == RIMU_ERR_INVALID && throw(ArgumentError(msg))   # ArgumentError in the reference (pdvec.jl:814-819)
    throw(RimuB200Error(Cint(status), msg))
end

# ---------------------------------------------------------------------------------------------------------------- addresses
# Device keys are W little-endian UInt64 words, word 0 least significant (include/rimu_b200.h).  The codec goes through
# the occupation-number representation, so it works for every storage type of the address (BitString of any chunk width,
# SortedParticleList): BoseFS -- mode 1 in the lowest bits, n ones then a 0 separator (bitstring.jl:464-472); FermiFS --
# bit m-1 <-> mode m (bitstring.jl:713-723); two-component FermiFS -- component c in bits [c*M, (c+1)*M).
words(::Type{<:BoseFS{N,M}}) where {N,M} = cld(N + M, 64)          # B + 1 bits: one spare bit marks empty table slots
words(::Type{<:FermiFS{N,M}}) where {N,M} = cld(M + 1, 64)
# CompositeFS: two FermiFS components of at most 32 modes are the one-word FermiFS2C layout (RIMU_ADDR_FERMI2C); every other
# CompositeFS (bosonic or mixed components, 3-4 components, wider fermions) is the general packed layout RIMU_ADDR_COMPOSITE:
# the components' bit strings side by side from the low bits (multicomponent.jl:10-34 keeps one BitString per component)
comp_bits(::Type{<:BoseFS{N,M}}) where {N,M} = N + M - 1
comp_bits(::Type{<:FermiFS{N,M}}) where {N,M} = M
comp_types(::Type{A}) where {A<:CompositeFS} = fieldtypes(fieldtype(A, :components))
is_fermi2c(::Type{A}) where {C,N,M,A<:CompositeFS{C,N,M}} = C == 2 && M <= 32 && all(T -> T <: FermiFS, comp_types(A))
words(::Type{A}) where {C,N,M,A<:CompositeFS{C,N,M}} =
    is_fermi2c(A) ? cld(2M, 64) : cld(sum(comp_bits, comp_types(A)) + 1, 64)
words(a::AbstractFockAddress) = words(typeof(a))

function set_bit!(key::Vector{UInt64}, pos::Int)
    key[pos >> 6 + 1] |= UInt64(1) << (pos & 63)
    return key
end
get_bit(key::AbstractVector{UInt64}, pos::Int) = (key[pos >> 6 + 1] >> (pos & 63)) & UInt64(1) == UInt64(1)

function to_key(a::BoseFS{N,M}) where {N,M}
    key = zeros(UInt64, words(typeof(a)))
    pos = 0
    for n in onr(a)
        for _ in 1:n
            set_bit!(key, pos)
            pos += 1
        end
        pos += 1                                                  # the 0 that closes the mode
    end
    return key
end
function to_key(a::FermiFS{N,M}) where {N,M}
    key = zeros(UInt64, words(typeof(a)))
    for (m, n) in enumerate(onr(a))
        n == 1 && set_bit!(key, m - 1)
    end
    return key
end
function to_key(a::CompositeFS{C,N,M}) where {C,N,M}
    key = zeros(UInt64, words(typeof(a)))
    base = 0
    for comp in a.components
        if comp isa FermiFS
            for (m, n) in enumerate(onr(comp))
                n == 1 && set_bit!(key, base + m - 1)
            end
        else
            pos = base
            for n in onr(comp)
                for _ in 1:n
                    set_bit!(key, pos)
                    pos += 1
                end
                pos += 1
            end
        end
        base += comp_bits(typeof(comp))
    end
    return key
end

function from_key(::Type{A}, key::AbstractVector{UInt64}) where {N,M,A<:BoseFS{N,M}}
    occ = zeros(Int, M)
    pos = 0
    for m in 1:M
        while pos < N + M - 1 && get_bit(key, pos)
            occ[m] += 1
            pos += 1
        end
        pos += 1
    end
    return A(Tuple(occ))
end
function from_key(::Type{A}, key::AbstractVector{UInt64}) where {N,M,A<:FermiFS{N,M}}
    return A(ntuple(m -> Int(get_bit(key, m - 1)), M))
end
function component_from_key(::Type{T}, key::AbstractVector{UInt64}, base::Int) where {N,M,T<:FermiFS{N,M}}
    return T(ntuple(m -> Int(get_bit(key, base + m - 1)), M))
end
function component_from_key(::Type{T}, key::AbstractVector{UInt64}, base::Int) where {N,M,T<:BoseFS{N,M}}
    occ = zeros(Int, M)
    pos = 0
    for m in 1:M
        while pos < N + M - 1 && get_bit(key, base + pos)
            occ[m] += 1
            pos += 1
        end
        pos += 1
    end
    return T(Tuple(occ))
end
function from_key(::Type{A}, key::AbstractVector{UInt64}) where {C,N,M,A<:CompositeFS{C,N,M}}
    base = 0
    comps = map(comp_types(A)) do T
        c = component_from_key(T, key, base)
        base += comp_bits(T)
        c
    end
    return A(comps)
end

# ---------------------------------------------------------------------------------------------------------------- context
# One GPU + stream + working memory (+ NCCL communicator) = PDWorkingMemory's buffers + the communicator
mutable struct Context
    ptr::Ptr{Cvoid}
    words::Int
end
function Context(W::Integer; device::Integer=0, table_slots::Integer=1 << 22)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rimu_ctx_create, LIB), Cint, (Cint, Cint, UInt64, Ptr{Ptr{Cvoid}}), device, W, table_slots, out))
    ctx = Context(out[], W)
    finalizer(c -> ccall((:rimu_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), c.ptr), ctx)
    return ctx
end
function table_slots(ctx::Context)
    out = Ref{UInt64}(0)
    check(ccall((:rimu_ctx_table_slots, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}), ctx.ptr, out))
    return out[]
end
resize_table!(ctx::Context, slots=4 * table_slots(ctx)) =
    check(ccall((:rimu_ctx_resize_table, LIB), Cint, (Ptr{Cvoid}, UInt64), ctx.ptr, slots))
synchronize(ctx::Context) = check(ccall((:rimu_ctx_synchronize, LIB), Cint, (Ptr{Cvoid},), ctx.ptr))
make_current(ctx::Context) = check(ccall((:rimu_ctx_make_current, LIB), Cint, (Ptr{Cvoid},), ctx.ptr))

# MPI -> NCCL (DictVectors/communicators.jl:546-606).  One Julia process per GPU; rank 0 creates the id, the launcher
# broadcasts its 128 bytes (MPI.Bcast!, Distributed, a file), every rank attaches.
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:rimu_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    return id
end
function comm_init!(ctx::Context, id::Vector{UInt8}, rank::Integer, nranks::Integer; records_per_peer::Integer=1 << 22)
    check(ccall((:rimu_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint, UInt64), ctx.ptr, id, rank, nranks, records_per_peer))
    return ctx
end
function comm_rank(ctx::Context)
    r, n = Ref{Cint}(0), Ref{Cint}(1)
    check(ccall((:rimu_comm_rank, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}), ctx.ptr, r, n))
    return Int(r[]), Int(n[])
end
comm_detach!(ctx::Context) = check(ccall((:rimu_comm_detach, LIB), Cint, (Ptr{Cvoid},), ctx.ptr))
function grow_exchange!(ctx::Context)
    cap, need = Ref{UInt64}(0), Ref{UInt64}(0)
    check(ccall((:rimu_comm_capacity, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt64}), ctx.ptr, cap, need))
    newcap = max(2 * cap[], need[] + need[] ÷ 2)
    check(ccall((:rimu_comm_reserve, LIB), Cint, (Ptr{Cvoid}, UInt64), ctx.ptr, newcap))
end
owner_rank(a::AbstractFockAddress, nranks::Integer) =
    Int(ccall((:rimu_addr_owner, LIB), Cint, (Ptr{UInt64}, Cint, Cint), to_key(a), words(a), nranks))

# ---------------------------------------------------------------------------------------------------------------- Hamiltonians
const MAX_MODES, MAX_TABLE_MODES, MAX_COMPONENTS = 128, 64, 4

struct HamDesc                        # == rimu_ham_desc (include/rimu_b200.h); asserted against rimu_sizeof_ham_desc()
    model::Int32
    addr_kind::Int32
    num_modes::Int32
    num_components::Int32
    num_particles::NTuple{2,Int32}
    ndim::Int32
    dims::NTuple{3,Int32}
    fold::NTuple{3,Int32}
    cutoff::Int32
    three_body_term::Int32
    has_potential::Int32
    boundary_condition::Int32
    u::Float64
    t::Float64
    v::Float64
    t_comp::NTuple{2,Float64}
    u_mat::NTuple{4,Float64}
    kes::NTuple{64,Float64}
    ws::NTuple{64,Float64}
    us::NTuple{64,Float64}
    potential::NTuple{256,Float64}
    comp_kind::NTuple{4,Int32}            # RIMU_ADDR_COMPOSITE: kind, particle number, t[c], u[i + C*j] per component
    comp_particles::NTuple{4,Int32}
    comp_t::NTuple{4,Float64}
    comp_u::NTuple{16,Float64}
end

pad(xs, n) = ntuple(i -> i <= length(xs) ? Float64(xs[i]) : 0.0, n)
pad3(xs, fill) = ntuple(i -> i <= length(xs) ? Int32(xs[i]) : Int32(fill), 3)

addr_kind(::BoseFS) = Int32(0)
addr_kind(::FermiFS) = Int32(1)
addr_kind(a::CompositeFS) = is_fermi2c(typeof(a)) ? Int32(2) : Int32(3)
particles(a::SingleComponentFockAddress) = (Int32(num_particles(a)), Int32(0))
particles(a::CompositeFS) = is_fermi2c(typeof(a)) ?
    (Int32(num_particles(a.components[1])), Int32(num_particles(a.components[2]))) : (Int32(0), Int32(0))
pad4(xs) = ntuple(i -> i <= length(xs) ? Int32(xs[i]) : Int32(0), 4)
comp_kinds(a::AbstractFockAddress) = pad4(())
comp_kinds(a::CompositeFS) = pad4([c isa BoseFS ? 0 : 1 for c in a.components])
comp_particles(a::AbstractFockAddress) = pad4(())
comp_particles(a::CompositeFS) = pad4([num_particles(c) for c in a.components])
components(a::SingleComponentFockAddress) = Int32(1)
components(a::CompositeFS{C}) where {C} = Int32(C)

function base_desc(model, a; ndim=0, dims=(), fold=(), cutoff=0, three_body=false, has_potential=false, bc=0,
                   u=0.0, t=0.0, v=0.0, t_comp=(0.0, 0.0), u_mat=(0.0, 0.0, 0.0, 0.0), kes=(), ws=(), us=(), potential=(),
                   comp_t=(), comp_u=())
    num_modes(a) <= MAX_MODES || throw(ArgumentError("at most $MAX_MODES modes"))
    return HamDesc(Int32(model), addr_kind(a), Int32(num_modes(a)), components(a), particles(a), Int32(ndim),
                   pad3(dims, 1), pad3(fold, 0), Int32(cutoff), Int32(three_body), Int32(has_potential), Int32(bc),
                   Float64(u), Float64(t), Float64(v), pad(t_comp, 2), pad(u_mat, 4),
                   pad(kes, MAX_TABLE_MODES), pad(ws, MAX_TABLE_MODES), pad(us, MAX_TABLE_MODES), pad(potential, 2 * MAX_MODES),
                   comp_kinds(a), comp_particles(a), pad(comp_t, MAX_COMPONENTS), pad(comp_u, MAX_COMPONENTS^2))
end

const BOUNDARY = Dict(:periodic => 0, :hard_wall => 1, :twisted => 2)

desc(h::HubbardReal1D) = base_desc(0, h.add; u=h.u, t=h.t)                                         # HubbardReal1D.jl:23-31
desc(h::HubbardReal1DEP) = base_desc(4, h.address; u=h.u, t=h.t, has_potential=true, potential=h.ep)  # HubbardReal1DEP.jl:47-59
function desc(h::ExtendedHubbardReal1D{<:Any,<:Any,U,V,T,BC}) where {U,V,T,BC}                      # ExtendedHubbardReal1D.jl:30-66
    BC isa Symbol || throw(ArgumentError("a complex twist angle has no device path"))
    return base_desc(5, h.address; u=U, v=V, t=T, bc=BOUNDARY[BC])
end
desc(h::HubbardMom1D) = base_desc(1, h.address; u=h.u, t=h.t, kes=h.kes)                            # HubbardMom1D.jl:43-65
function desc(h::ExtendedHubbardMom1D)                                                              # ExtendedHubbardMom1D.jl:37-57
    h.address isa BoseFS || throw(ArgumentError("ExtendedHubbardMom1D has a device path for BoseFS addresses only"))
    h.boundary_condition == 0 || throw(ArgumentError("a twisted boundary condition has no device path"))
    M = num_modes(h.address)
    # the cosines get_offdiagonal (:99-102) and extended_momentum_transfer_diagonal (excitations.jl:152) evaluate per element
    return base_desc(6, h.address; u=h.u, v=h.v, t=h.t, kes=h.kes,
                     ws=[cos(q * 2π / M) for q in 0:M-1], us=[cos(d * (2π / M)) for d in 0:M-1])
end
desc(h::HubbardMom1DEP) = base_desc(7, h.address; u=h.u, t=h.t, kes=h.kes, has_potential=true, potential=h.ep)  # HubbardMom1DEP.jl:68-96
function desc(h::HubbardRealSpace)                                                                  # HubbardRealSpace.jl:163-241
    g = h.geometry
    dims = size(g)
    C = Int(components(h.address))
    C <= MAX_COMPONENTS || throw(ArgumentError("at most $MAX_COMPONENTS components have a device layout"))
    pot = h.potential === nothing ? () : vec(h.potential)               # M x C column major = potential[c*M + site]
    common = (; ndim=length(dims), dims=dims, fold=Int.(collect(Rimu.Hamiltonians.fold(g))),
              has_potential=h.potential !== nothing, potential=pot)
    if h.address isa CompositeFS && !is_fermi2c(typeof(h.address))       # general CompositeFS: u[i + C*j], t[c]
        sum(comp_bits, comp_types(typeof(h.address))) <= 127 || throw(ArgumentError("addresses beyond 127 bits have no device layout"))
        umat = h.u === nothing ? () : vec(collect(h.u))
        return base_desc(2, h.address; common..., comp_t=Tuple(h.t), comp_u=umat)
    end
    umat = h.u === nothing ? (0.0, 0.0, 0.0, 0.0) : (C == 1 ? (h.u[1, 1], 0.0, 0.0, 0.0) : (h.u[1, 1], h.u[2, 1], h.u[1, 2], h.u[2, 2]))
    return base_desc(2, h.address; common..., t_comp=Tuple(h.t), u_mat=umat)
end
function desc(h::Transcorrelated1D)                                                                 # Transcorrelated1D.jl:55-89
    h.v_ho == 0 || throw(ArgumentError("Transcorrelated1D with v_ho != 0 has no device path"))
    return base_desc(3, h.address; t=h.t, v=h.v, cutoff=h.cutoff, three_body=h.three_body_term, kes=h.kes, ws=h.ws, us=h.us)
end
desc(h::AbstractHamiltonian) =
    throw(ArgumentError("$(typeof(h)) cannot run on the GPU: only the built-in lattice models have device code, and there is no CPU fallback"))

mutable struct GPUHam
    ptr::Ptr{Cvoid}
end
const HAM_CACHE = IdDict{Any,GPUHam}()                 # one device object per (Hamiltonian, context GPU)
function gpu_ham(h::AbstractHamiltonian, ctx::Context)
    return get!(HAM_CACHE, (h, ctx.ptr)) do
        make_current(ctx)                                # the tables go to the context's GPU
        out = Ref{Ptr{Cvoid}}(C_NULL)
        d = Ref(desc(h))
        check(ccall((:rimu_ham_create, LIB), Cint, (Ptr{HamDesc}, Ptr{Ptr{Cvoid}}), d, out))
        gh = GPUHam(out[])
        finalizer(x -> ccall((:rimu_ham_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), gh)
        gh
    end
end

# cross-check hooks: diagonal_element / num_offdiagonals / get_offdiagonal evaluated by the DEVICE code
function gpu_diagonal(ctx::Context, h::AbstractHamiltonian, addrs::AbstractVector)
    keys = reduce(vcat, to_key.(addrs))
    out = zeros(Float64, length(addrs))
    check(ccall((:rimu_ham_diagonal, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt64}, Int64, Ptr{Float64}),
                ctx.ptr, gpu_ham(h, ctx).ptr, keys, length(addrs), out))
    return out
end
function gpu_num_offdiagonals(ctx::Context, h::AbstractHamiltonian, addrs::AbstractVector)
    keys = reduce(vcat, to_key.(addrs))
    out = zeros(Int64, length(addrs))
    check(ccall((:rimu_ham_num_offdiagonals, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt64}, Int64, Ptr{Int64}),
                ctx.ptr, gpu_ham(h, ctx).ptr, keys, length(addrs), out))
    return out
end
function gpu_offdiagonals(ctx::Context, h::AbstractHamiltonian, addr::A, first::Integer=1, count::Integer=num_offdiagonals(h, addr)) where {A}
    W = words(A)
    keys_out = zeros(UInt64, W * count)
    vals_out = zeros(Float64, count)
    check(ccall((:rimu_ham_offdiagonals, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt64}, Int64, Int64, Ptr{UInt64}, Ptr{Float64}),
                ctx.ptr, gpu_ham(h, ctx).ptr, to_key(addr), first, count, keys_out, vals_out))
    return [(from_key(A, view(keys_out, (i - 1) * W + 1:i * W)), vals_out[i]) for i in 1:count]
end

# ---------------------------------------------------------------------------------------------------------------- the vector
const VAL_F64, VAL_I64 = 0, 1
val_type(::Type{Float64}) = VAL_F64
val_type(::Type{Int64}) = VAL_I64
val_type(::Type{V}) where {V} = throw(ArgumentError("device vectors hold Float64 or Int64 values, not $V"))

"""
    GPUDVec{K,V} <: AbstractDVec{K,V}

Dictionary-semantics vector living in HBM (replaces `DVec`, dvec.jl:44-47, and `PDVec`, pdvec.jl:156-163).
"""
mutable struct GPUDVec{K,V,S<:StochasticStyle{V},I<:InitiatorRule} <: AbstractDVec{K,V}
    ptr::Ptr{Cvoid}
    ctx::Context
    style::S
    initiator::I
end

function create_vec(ctx::Context, ::Type{V}, capacity::Integer) where {V}
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rimu_vec_create, LIB), Cint, (Ptr{Cvoid}, Cint, UInt64, Ptr{Ptr{Cvoid}}), ctx.ptr, val_type(V), capacity, out))
    return out[]
end
function GPUDVec{K,V}(; style::StochasticStyle{V}, ctx::Context, initiator::InitiatorRule=NonInitiator(), capacity::Integer=1 << 12) where {K,V}
    v = GPUDVec{K,V,typeof(style),typeof(initiator)}(create_vec(ctx, V, capacity), ctx, style, initiator)
    finalizer(x -> ccall((:rimu_vec_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), v)
    return v
end
# DVec(pairs...; style) (dvec.jl:62-100): duplicates are summed, zeros dropped, keys of other ranks dropped
function GPUDVec(pairs::Pair{K,V}...; style::StochasticStyle=default_style(V), ctx::Context, kw...) where {K,V}
    VV = eltype(style)
    v = GPUDVec{K,VV}(; style, ctx, capacity=max(length(pairs), 1 << 12), kw...)
    upload!(v, collect(first.(pairs)), VV.(collect(last.(pairs))))
    return v
end
default_style(::Type{<:Integer}) = IsStochasticInteger()
default_style(::Type{<:AbstractFloat}) = IsDeterministic()

function upload!(v::GPUDVec{K,V}, addrs::Vector{K}, vals::Vector{V}) where {K,V}
    keys = isempty(addrs) ? UInt64[] : reduce(vcat, to_key.(addrs))
    check(ccall((:rimu_vec_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{Cvoid}, Int64), v.ptr, keys, vals, length(vals)))
    return v
end
function download(v::GPUDVec{K,V}) where {K,V}
    n = length(v)
    W = v.ctx.words
    keys = zeros(UInt64, W * max(n, 1))
    vals = zeros(V, max(n, 1))
    got = Ref{Int64}(0)
    check(ccall((:rimu_vec_download, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{Cvoid}, Int64, Ptr{Int64}), v.ptr, keys, vals, n, got))
    return [from_key(K, view(keys, (i - 1) * W + 1:i * W)) => vals[i] for i in 1:got[]]
end

StochasticStyle(v::GPUDVec) = v.style
localpart(v::GPUDVec) = v
Base.eltype(::Type{<:GPUDVec{K,V}}) where {K,V} = Pair{K,V}
Base.keytype(::Type{<:GPUDVec{K}}) where {K} = K
Base.valtype(::Type{<:GPUDVec{K,V}}) where {K,V} = V
function Base.length(v::GPUDVec)                                              # length(localpart(v)), pdvec.jl:275-278
    n = Ref{Int64}(0)
    check(ccall((:rimu_vec_length, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), v.ptr, n))
    return Int(n[])
end
function Base.getindex(v::GPUDVec{K,V}, k::K) where {K,V}                      # missing -> zero (pdvec.jl:328-335)
    out = Ref{V}(zero(V))
    check(ccall((:rimu_vec_get, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{Cvoid}), v.ptr, to_key(k), out))
    return out[]
end
Base.pairs(v::GPUDVec) = download(v)                                          # convenience methods work on a downloaded copy
Base.keys(v::GPUDVec) = first.(download(v))
Base.values(v::GPUDVec) = last.(download(v))
Base.iterate(v::GPUDVec, st...) = iterate(download(v), st...)
Base.convert(::Type{DVec}, v::GPUDVec) = DVec(download(v)...; style=v.style)

function similar_empty(v::GPUDVec{K,V}; style=v.style) where {K,V}
    return GPUDVec{K,eltype(style)}(; style, ctx=v.ctx, initiator=v.initiator, capacity=max(length(v), 1 << 12))
end
Base.similar(v::GPUDVec) = similar_empty(v)
Base.empty(v::GPUDVec) = similar_empty(v)
zerovector(v::GPUDVec) = similar_empty(v)                                     # pmc_simulation.jl:142
function zerovector!(v::GPUDVec)
    check(ccall((:rimu_vec_clear, LIB), Cint, (Ptr{Cvoid},), v.ptr))
    return v
end
Base.empty!(v::GPUDVec) = zerovector!(v)
function Base.copy!(dst::GPUDVec, src::GPUDVec)                                # copy!/copyto!, eltype conversion on the device
    check(ccall((:rimu_vec_copy, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), dst.ptr, src.ptr))
    return dst
end
Base.copyto!(dst::GPUDVec, src::GPUDVec) = copy!(dst, src)
Base.copy(v::GPUDVec) = copy!(similar_empty(v), v)
Base.deepcopy(v::GPUDVec) = copy(v)
function scale!(v::GPUDVec, α::Number)                                         # pdvec.jl:714-729
    check(ccall((:rimu_vec_scale, LIB), Cint, (Ptr{Cvoid}, Float64), v.ptr, α))
    return v
end
function add!(y::GPUDVec, x::GPUDVec, α::Number=1, β::Number=1)                 # y = α x + β y  (pdvec.jl:731-758)
    check(ccall((:rimu_vec_axpby, LIB), Cint, (Float64, Ptr{Cvoid}, Float64, Ptr{Cvoid}, Ptr{Cvoid}), α, x.ptr, β, y.ptr, y.ptr))
    return y
end
function dot(x::GPUDVec, y::GPUDVec)                                            # pdvec.jl:760-796 (global with a communicator)
    out = Ref{Float64}(0.0)
    check(ccall((:rimu_vec_dot, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}), x.ptr, y.ptr, out))
    return out[]
end
function norm(v::GPUDVec, p::Real=2)                                            # abstractdvec.jl:200-256
    out = Ref{Float64}(0.0)
    check(ccall((:rimu_vec_norm, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), v.ptr, p == Inf ? 0 : Int(p), out))
    return out[]
end
walkernumber(v::GPUDVec) = norm(v, 1)                                           # abstractdvec.jl:258-260
walkernumber_and_length(v::GPUDVec) = (norm(v, 1), global_length(v))            # pdvec.jl:896-902
function global_length(v::GPUDVec)
    n = Float64[length(v)]
    check(ccall((:rimu_comm_allreduce_f64, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), v.ctx.ptr, n, 1))
    return Int(n[1])
end
freeze(v::GPUDVec) = FrozenDVec(download(v))                                    # projectors.jl:164: a host-side list of pairs
function dot(f::FrozenDVec, v::GPUDVec{K}) where {K}                            # pdvec.jl:773-779: per-key bucket-segment lookups
    prs = collect(pairs(f))
    keys = isempty(prs) ? UInt64[] : reduce(vcat, to_key.(first.(prs)))
    vals = Float64.(last.(prs))
    out = Ref{Float64}(0.0)
    check(ccall((:rimu_vec_dot_sparse, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{Float64}, Int64, Ptr{Float64}), v.ptr, keys, vals, length(vals), out))
    return out[]
end
dot(v::GPUDVec, f::FrozenDVec) = dot(f, v)
# three-argument dot (pdvec.jl:866-879; AdjointUnknown sweep abstractdvec.jl:313-324): an explicit H*v, then a dot
function dot(x::Union{GPUDVec,FrozenDVec}, op::AbstractHamiltonian, y::GPUDVec)
    yd = StochasticStyle(y) isa IsDeterministic ? y : copy!(similar_empty(y; style=IsDeterministic()), y)
    tmp = similar_empty(yd)
    mul!(tmp, op, yd)
    return dot(x, tmp)
end

# ---------------------------------------------------------------------------------------------------------------- the step
struct StepParams                     # == rimu_step_params; asserted against rimu_sizeof_step_params()
    style::Int32
    plain_h::Int32
    shift::Float64
    time_step::Float64
    boost::Float64
    proj_threshold::Float64
    rel_threshold::Float64
    abs_threshold::Float64
    compress_threshold::Float64
    seed::UInt64
    step::UInt64
    table_slots::UInt64
    initiator_rule::Int32
    ordered::Int32
    initiator_threshold::Float64
end
struct StepStats                      # == rimu_step_stats; asserted against rimu_sizeof_step_stats()
    exact_steps::Int64
    inexact_steps::Int64
    spawn_attempts::Int64
    len_before::Int64
    len::Int64
    spawns::Float64
    deaths::Float64
    clones::Float64
    zombies::Float64
    norm1::Float64
    ispawns::Int64
    ideaths::Int64
    iclones::Int64
    izombies::Int64
    inorm1::Int64
    local_len::Int64
    sent_records::Int64
    deposits::Int64
    ms_diag::Float32
    ms_spawn::Float32
    ms_exchange::Float32
    ms_compact::Float32
    ms_total::Float32
    ms_reduce::Float32
    buckets::Int64
    max_bucket_fill::Int64
end

function __init__()
    @assert sizeof(HamDesc) == ccall((:rimu_sizeof_ham_desc, LIB), Cint, ())
    @assert sizeof(StepParams) == ccall((:rimu_sizeof_step_params, LIB), Cint, ())
    @assert sizeof(StepStats) == ccall((:rimu_sizeof_step_stats, LIB), Cint, ())
    @assert sizeof(ShiftParams) == ccall((:rimu_sizeof_shift_params, LIB), Cint, ())
end

mutable struct GPUWorkingMemory{S,I}
    ctx::Context
    style::S
    initiator::I
    seed::UInt64
    counter::UInt64
    ordered::Bool          # order-deterministic Float64 summation (rimu_step_params.ordered): bit-reproducible steps
    last_stats::StepStats
end
# pmc_simulation.jl:125; PDWorkingMemory(v) pdworkingmemory.jl:104-108
working_memory(v::GPUDVec; seed=rand(UInt64), ordered=false) =
    GPUWorkingMemory(v.ctx, v.style, v.initiator, UInt64(seed), UInt64(0), ordered, StepStats(ntuple(_ -> 0, fieldcount(StepStats))...))

initiator_params(::NonInitiator) = (0, 0.0)                                       # initiators.jl:224-236
initiator_params(i::Initiator) = (1, Float64(i.threshold))                        # :132-160
initiator_params(i::SimpleInitiator) = (2, Float64(i.threshold))                  # :162-183
initiator_params(i::CoherentInitiator) = (3, Float64(i.threshold))                # :185-211

compression_threshold(c::ThresholdCompression) = Float64(c.threshold)
compression_threshold(::NoCompression) = 0.0
# (style id, projection threshold of the spawning strategy, rel / abs spawning thresholds, compression threshold)
style_params(::IsDeterministic) = (0, 0.0, 1.0, Inf, 0.0)                                        # styles.jl:76-105
style_params(::IsStochasticInteger) = (1, 0.0, 1.0, Inf, 0.0)                                    # :11-25
style_params(s::IsDynamicSemistochastic) = (2, Float64(s.spawning.strat.threshold), Float64(s.spawning.rel_threshold),
                                            Float64(s.spawning.abs_threshold), compression_threshold(s.compression))   # :175-214
style_params(s::IsStochasticWithThreshold) = (3, Float64(s.threshold), 1.0, Inf, 0.0)            # :117-130
style_params(s::StochasticStyle) = throw(ArgumentError("$(typeof(s)) has no device path"))

# names and order of step_stats (styles.jl:14-20, 94-96, 203-209; len_before from compression.jl:16)
function step_stats_tuple(::IsStochasticInteger, s::StepStats)
    return (:spawn_attempts, :spawns, :deaths, :clones, :zombies), (s.spawn_attempts, s.ispawns, s.ideaths, s.iclones, s.izombies)
end
step_stats_tuple(::IsDeterministic, s::StepStats) = (:exact_steps,), (s.exact_steps,)
function step_stats_tuple(::IsStochasticWithThreshold, s::StepStats)
    return (:spawn_attempts, :spawns, :deaths, :clones, :zombies), (s.spawn_attempts, s.spawns, s.deaths, s.clones, s.zombies)
end
function step_stats_tuple(st::IsDynamicSemistochastic, s::StepStats)
    names = (:exact_steps, :inexact_steps, :spawn_attempts, :spawns)
    values = (s.exact_steps, s.inexact_steps, s.spawn_attempts, s.spawns)
    st.compression isa ThresholdCompression || return names, values
    return (names..., :len_before), (values..., s.len_before)
end

"""
    apply_operator!(wm::GPUWorkingMemory, target, source, op, boost=1) -> (stat_names, stats, wm, target)

One FCIQMC step (op = FirstOrderTransitionOperator, fciqmc.jl:78-112) or one matrix-free `H*v` (op = H) on the GPU.
"""
function apply_operator!(wm::GPUWorkingMemory, target::GPUDVec, source::GPUDVec, op, boost=1)
    target === source && throw(ArgumentError("source and target must not alias"))   # Interfaces/dictvectors.jl:115-117
    ham, plain, shift, dτ = op isa Rimu.FirstOrderTransitionOperator ?
        (op.hamiltonian, 0, Float64(op.shift), Float64(op.time_step)) : (op, 1, 0.0, 0.0)
    sty, pt, rt, at, ct = style_params(wm.style)
    ir, it = initiator_params(wm.initiator)
    params = Ref(StepParams(sty, plain, shift, dτ, Float64(boost), pt, rt, at, ct, wm.seed, wm.counter, 0, ir, Int32(wm.ordered), it))
    stats = Ref(wm.last_stats)
    gh = gpu_ham(ham, wm.ctx)
    table_retries = 0
    while true
        st = ccall((:rimu_step, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{StepParams}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{StepStats}),
                   wm.ctx.ptr, gh.ptr, params, source.ptr, target.ptr, stats)
        if st == RIMU_ERR_TABLE_FULL && table_retries < 6         # table method only: recoverable, source untouched
            table_retries += 1
            resize_table!(wm.ctx)
            continue
        elseif st == RIMU_ERR_EXCHANGE_FULL                        # staged exchange: the same decision on every rank
            grow_exchange!(wm.ctx)
            continue
        end
        check(st)
        break
    end
    wm.counter += 1
    wm.last_stats = stats[]
    names, values = step_stats_tuple(wm.style, stats[])
    return names, values, wm, target
end

# ---- a batch of steps: the body of advance!(::FCIQMC) (fciqmc.jl:126-181) x nsteps with the shift update on the device
struct ShiftParams                    # == rimu_shift_params; asserted against rimu_sizeof_shift_params()
    strategy::Int32
    shift_mode::Int32
    target_walkers::Float64
    zeta::Float64
    xi::Float64
    shift::Float64
    pnorm::Float64
    max_length::Int64
end
# shiftstrategy.jl:77-215 -> (RIMU_SHIFT_* id, target_walkers, zeta, xi)
shift_strategy_params(s::Rimu.DontUpdate) = (Int32(0), Float64(s.target_walkers), 0.0, 0.0)
shift_strategy_params(s::Rimu.LogUpdate) = (Int32(1), 0.0, Float64(s.ζ), 0.0)
shift_strategy_params(s::Rimu.LogUpdateAfterTargetWalkers) = (Int32(2), Float64(s.target_walkers), Float64(s.ζ), 0.0)
shift_strategy_params(s::Rimu.DoubleLogUpdate) = (Int32(3), Float64(s.target_walkers), Float64(s.ζ), Float64(s.ξ))
shift_strategy_params(s::Rimu.DoubleLogUpdateAfterTargetWalkers) = (Int32(4), Float64(s.target_walkers), Float64(s.ζ), Float64(s.ξ))
shift_strategy_params(s) = throw(ArgumentError("$(typeof(s)) needs the vectors on the host every step: use the step-by-step loop"))

struct ProjectorArg                   # == rimu_projector: host pairs of a FrozenDVec
    keys::Ptr{UInt64}
    values::Ptr{Float64}
    n::Int64
end

"""
    advance_steps!(wm, v, pv, hamiltonian, shift_parameters, shift_strategy, nsteps; max_length=0, projectors=())
        -> (v, pv, stats::Vector{StepStats}, shifts::Vector{Float64}, dots::Matrix{Float64})

`nsteps` iterations of `apply_operator!`, swap and `update_shift_parameters!` in one call (`rimu_advance`); `shift_parameters`
(Rimu's `DefaultShiftParameters`) is updated in place.  `projectors` are `FrozenDVec`s (the frozen `vproj` / `hproj` of
`ProjectedEnergy`, poststepstrategy.jl:82-121): `dots[j, k]` = `dot(projectors[j], v)` after step `k`, evaluated on the device.
Fewer than `nsteps` entries come back when the run ended.
"""
function advance_steps!(wm::GPUWorkingMemory, v::GPUDVec, pv::GPUDVec, ham::AbstractHamiltonian, sp, strategy, nsteps::Integer;
                        max_length::Integer=0, projectors=())
    sty, pt, rt, at, ct = style_params(wm.style)
    ir, it = initiator_params(wm.initiator)
    params = Ref(StepParams(sty, 0, Float64(sp.shift), Float64(sp.time_step), 1.0, pt, rt, at, ct, wm.seed, wm.counter, 0, ir, Int32(wm.ordered), it))
    id, target, ζ, ξ = shift_strategy_params(strategy)
    shp = Ref(ShiftParams(id, Int32(sp.shift_mode), target, ζ, ξ, Float64(sp.shift), Float64(sp.pnorm), Int64(max_length)))
    stats = Vector{StepStats}(undef, nsteps)
    shifts = zeros(Float64, nsteps)
    done, in_w = Ref{Int64}(0), Ref{Int32}(0)
    nproj = length(projectors)
    pkeys = [isempty(pairs(f)) ? UInt64[] : reduce(vcat, to_key.(first.(collect(pairs(f))))) for f in projectors]
    pvals = [Float64.(last.(collect(pairs(f)))) for f in projectors]
    dots = zeros(Float64, max(nproj, 1), nsteps)                       # column k = step k (C: proj_out[k * nproj + j])
    GC.@preserve pkeys pvals begin
        pargs = [ProjectorArg(pointer(pkeys[j]), pointer(pvals[j]), length(pvals[j])) for j in 1:nproj]
        status = ccall((:rimu_advance, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{StepParams}, Ptr{ShiftParams}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{ProjectorArg}, Int32,
                     Ptr{StepStats}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int32}),
                    wm.ctx.ptr, gpu_ham(ham, wm.ctx).ptr, params, shp, v.ptr, pv.ptr, nsteps, pargs, nproj, stats, shifts, dots, done, in_w)
    end
    # the steps taken so far stay taken, also when a later one failed: update the host's bookkeeping before raising
    wm.counter += done[]
    sp.shift, sp.pnorm, sp.shift_mode = shp[].shift, shp[].pnorm, shp[].shift_mode != 0
    done[] > 0 && (wm.last_stats = stats[done[]])
    if status != 0 && in_w[] != 0
        v.ptr, pv.ptr = pv.ptr, v.ptr                                  # no return value on this path: `v` stays the current vector
    end
    check(status)
    in_w[] != 0 && ((v, pv) = (pv, v))
    return v, pv, stats[1:done[]], shifts[1:done[]], dots[1:nproj, 1:done[]]
end

function mul!(y::GPUDVec, op::AbstractHamiltonian, x::GPUDVec, wm=working_memory(x))   # pdvec.jl:810-822
    wm.style isa IsDeterministic ||
        throw(ArgumentError("Attempted to use `mul!` with non-deterministic working memory. Use `apply_operator!` instead."))
    apply_operator!(wm, y, x, op)
    return y
end
Base.:*(op::AbstractHamiltonian, x::GPUDVec) = mul!(similar_empty(x), op, x)

end # module

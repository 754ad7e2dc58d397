# NQCuda.jl -- NeuralQuantum.jl's hot path on libnqcuda (hand-written CUDA for sm_100a), bound with `ccall`.
#
# Drop this file next to NeuralQuantum.jl's `src/GPU/` (it replaces the GPUArrays specialisations there) and
# `include` it after the package: it adds methods to NeuralQuantum's OWN generic functions -- `cached`, `logψ!`,
# `logψ_and_∇logψ!`, `log_prob_ψ!`, the sampler cache / `init_sampler!` / `samplenext!`, `accumulate`-free local
# estimators, `setup_algorithm!`, `precondition!`, `workers_*` -- for the device-backed types defined here, so the
# driver loop of the examples (`sample!` -> `precondition!` -> `update!`) runs unchanged.
#
# STATUS: written against NeuralQuantum.jl v0.2.0 (file:line citations below are that checkout) and against
# include/nqcuda.h; NOT EXECUTED in this repository's build image (no `julia` binary there).  The executable host of
# the same ABI is the Python mirror `neuralquantum.jl_b200/nqcuda/` and the C test `tests/cabi_smoke.c`.
module NQCuda

using NeuralQuantum
using LinearAlgebra
import NeuralQuantum: cached, logψ!, logψ_and_∇logψ!, log_prob_ψ!, trainable, vec_data, out_type,
                      num_workers, worker_local_seed, workers_mean!, workers_sum!, workers_mean,
                      setup_algorithm!, precondition!, algorithm_cache, init_sampler!, samplenext!,
                      KLocalOperator, KLocalOperatorSum, KLocalOperatorTensor, KLocalLiouvillian,
                      MetropolisSampler, LocalRule, SR, HomogeneousSpin, HomogeneousFock, RBM, RBMSplit, NDM

const lib = get(ENV, "NQCUDA_LIB", joinpath(@__DIR__, "..", "neuralquantum.jl_b200", "libnqcuda.so"))

# ---- enums of include/nqcuda.h ---------------------------------------------------------------------------------
const NQ_F32, NQ_F64, NQ_C64, NQ_C128 = Cint(0), Cint(1), Cint(2), Cint(3)
const NQ_RBM, NQ_RBMSPLIT, NQ_NDM = Cint(0), Cint(1), Cint(2)
const NQ_SPIN, NQ_FOCK = Cint(0), Cint(1)
const NQ_SOFTPLUS, NQ_LOGCOSH = Cint(0), Cint(1)
const NQ_KET, NQ_SUPER = Cint(0), Cint(1)
const NQ_SOLVE_CHOLESKY, NQ_SOLVE_CG, NQ_SOLVE_MINRES, NQ_SOLVE_QLP, NQ_SOLVE_QLP_WARM = Cint.(0:4)
const NQ_ERR_NOT_POSDEF, NQ_ERR_NOT_CONVERGED = Cint(-5), Cint(-6)
const NQ_UNIQUE_ID_BYTES = 128

nqdtype(::Type{Float32}) = NQ_F32
nqdtype(::Type{Float64}) = NQ_F64
nqdtype(::Type{ComplexF32}) = NQ_C64
nqdtype(::Type{ComplexF64}) = NQ_C128
nqdtype(a::AbstractArray) = nqdtype(eltype(a))

struct NQError <: Exception
    status::Cint
    msg::String
end

# ---- context (one device + one stream; src/Parallel: one per worker) ----------------------------------------------
mutable struct Ctx
    h::Ptr{Cvoid}
end
function Ctx(device::Integer = 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:nq_ctx_create, lib), Cint, (Cint, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), device, C_NULL, h)
    st == 0 || throw(NQError(st, "nq_ctx_create"))
    c = Ctx(h[])
    finalizer(c -> (c.h != C_NULL && ccall((:nq_ctx_destroy, lib), Cint, (Ptr{Cvoid},), c.h); c.h = C_NULL), c)
    return c
end
function last_info(ctx::Ctx)
    v = Ref{Int64}(0)
    ccall((:nq_ctx_last_info, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}), ctx.h, v)
    return v[]
end
function check(ctx::Ctx, st::Cint)
    st == 0 && return nothing
    msg = unsafe_string(ccall((:nq_last_error, lib), Cstring, (Ptr{Cvoid},), ctx.h))
    st == NQ_ERR_NOT_POSDEF && throw(PosDefException(Int(last_info(ctx)) + 1))      # SRDirect.jl:79 (check = true)
    throw(NQError(st, msg))
end

# ---- machine plugin interface (base_batched_networks.jl:16-145, RBMBatched.jl, RBMSplitBatched.jl, NDMBatched.jl) ----
mutable struct CudaNet{N} <: NeuralQuantum.NNBatchedCache{N}
    ctx::Ctx
    h::Ptr{Cvoid}
    net::N
    P::Int
    batch_sz::Int
end
kind(::RBM) = NQ_RBM
kind(::RBMSplit) = NQ_RBMSPLIT
kind(::NDM) = NQ_NDM
activation(net) = hasproperty(net, :f) && net.f === NeuralQuantum.af_logcosh ? NQ_LOGCOSH : NQ_SOFTPLUS
hilbcode(::HomogeneousSpin) = NQ_SPIN
hilbcode(::HomogeneousFock) = NQ_FOCK

flat_params(net) = vcat((vec(x) for x in trainable(net))...)        # functor order (functor.jl:41, tuple_logic.jl:31-72)

nsites_of(net::RBM) = length(net.a)
nsites_of(net::RBMSplit) = length(net.ar)
nsites_of(net::NDM) = length(net.b_μ)
nhidden_of(net::Union{RBM,RBMSplit}) = length(net.b)
nhidden_of(net::NDM) = length(net.h_μ)
nancilla_of(net::NDM) = length(net.d_λ)
nancilla_of(net) = 0

"cached(ctx, net, hilb, batch_sz): the device twin of `cached(net, batch_sz)` (base_batched_networks.jl:20)."
function cached(ctx::Ctx, net::Union{RBM,RBMSplit,NDM}, hilb, batch_sz::Integer = 1)
    w = flat_params(net)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:nq_machine_create, lib), Cint,
          (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}),
          ctx.h, kind(net), hilbcode(hilb), nsites_of(net), nhidden_of(net), nancilla_of(net), activation(net),
          nqdtype(eltype(w)), h))
    c = CudaNet(ctx, h[], net, length(w), Int(batch_sz))
    finalizer(c -> (c.h != C_NULL && ccall((:nq_machine_destroy, lib), Cint, (Ptr{Cvoid},), c.h); c.h = C_NULL), c)
    sync_params!(c)
    return c
end
"Push the host network's parameters to the device (call after any host-side change; quirk Q9 has no cache here)."
function sync_params!(c::CudaNet)
    w = flat_params(c.net)
    GC.@preserve w check(c.ctx, ccall((:nq_machine_set_params, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                                       c.h, pointer(w), length(w)))
    return c
end
"Pull the device parameters back into the host network's arrays (field by field, functor order)."
function pull_params!(c::CudaNet)
    w = similar(flat_params(c.net))
    GC.@preserve w check(c.ctx, ccall((:nq_machine_get_params, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                                       c.h, pointer(w), length(w)))
    o = 0
    for x in trainable(c.net)
        copyto!(x, reshape(view(w, o+1:o+length(x)), size(x)))
        o += length(x)
    end
    return c.net
end
out_type(c::CudaNet) = out_type(c.net)

_ptrs(σ::AbstractArray) = (pointer(σ), Ptr{Cvoid}(C_NULL), eltype(σ), size(σ, 2))
_ptrs(σ::Tuple) = (pointer(σ[1]), pointer(σ[2]), eltype(σ[1]), size(σ[1], 2))

function logψ!(out::AbstractArray, c::CudaNet, σ)                    # RBMBatched.jl:37-56, NDMBatched.jl:94-175
    pr, pc, T, B = _ptrs(σ)
    GC.@preserve σ out check(c.ctx, ccall((:nq_logpsi, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}), c.h, pr, pc, nqdtype(T), B, pointer(out)))
    return out
end
function logψ_and_∇logψ!(∇, out::AbstractArray, c::CudaNet, σ)       # RBMBatched.jl:58-91, NDMBatched.jl:177-280
    G = vec_data(∇)[1]                                               # the ONE [P, B] buffer (tuple_logic.jl:82-118)
    pr, pc, T, B = _ptrs(σ)
    GC.@preserve σ out G check(c.ctx, ccall((:nq_logpsi_grad, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Int64),
        c.h, pr, pc, nqdtype(T), B, pointer(out), pointer(G), size(G, 1)))
    return out
end
function log_prob_ψ!(prob::AbstractArray, c::CudaNet, σ)             # base_batched_networks.jl:255-259
    pr, pc, T, B = _ptrs(σ)
    GC.@preserve σ prob check(c.ctx, ccall((:nq_log_prob, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}), c.h, pr, pc, nqdtype(T), B, pointer(prob)))
    return prob
end
"Optimisers.update!(Descent(η), net, Δw) on the device parameters (Optimisers/apply.jl:25-73, rules.jl:11-17)."
function update!(c::CudaNet, Δw::AbstractVector, η::Real)
    GC.@preserve Δw check(c.ctx, ccall((:nq_update, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble), c.h, pointer(Δw), η))
    return c
end

# ---- operators: flatten the per-local-row connection tables (KLocalOperator.jl:54-112) ----------------------------
struct FlatOp
    nsites::Vector{Int32}
    sites::Vector{Int32}
    rowptr::Vector{Int64}
    mel::Vector{ComplexF64}
    flip::Vector{UInt32}
    left::Vector{Int32}
    right::Vector{Int32}
end
FlatOp() = FlatOp(Int32[], Int32[], Int64[0], ComplexF64[], UInt32[], Int32[], Int32[])

"Append one k-local term as a part; returns its 0-based part index.  Entry order = op_conns order (diagonal first)."
function push_part!(f::FlatOp, t::KLocalOperator)
    push!(f.nsites, length(t.sites))
    append!(f.sites, Int32.(t.sites .- 1))
    for conns in t.op_conns                                          # one local row r = 1 .. d^k
        for i in 1:length(conns)
            m, cng = conns[i]
            mask = UInt32(0)
            for s in cng.to_change                                   # local dimension 2: a change list is a flip mask
                mask |= UInt32(1) << (findfirst(==(s), t.sites) - 1)
            end
            push!(f.mel, ComplexF64(m))
            push!(f.flip, mask)
        end
        push!(f.rowptr, length(f.mel))                               # row boundary
    end
    return Int32(length(f.nsites) - 1)
end
terms(op::KLocalOperatorSum) = NeuralQuantum.operators(op)
terms(op::KLocalOperator) = [op]

function flatten(op::Union{KLocalOperator,KLocalOperatorSum})        # ket operator: KLocalOperatorSum.jl:65-72
    f = FlatOp()
    for t in terms(op)
        push!(f.left, push_part!(f, t)); push!(f.right, Int32(-1))
    end
    return f, NQ_KET
end
function flatten(L::KLocalLiouvillian)                               # KLocalLiouvillian.jl:46-52: HnH_l, HnH_r, LLdag
    f = FlatOp()
    for grp in (L.HnH_l, L.HnH_r, L.LLdag), t in terms_tensor(grp)
        l = t.op_l === nothing ? Int32(-1) : push_part!(f, t.op_l)
        r = t.op_r === nothing ? Int32(-1) : push_part!(f, t.op_r)
        push!(f.left, l); push!(f.right, r)                          # (p,-1) | (-1,p) | (pl,pr): KLocalOperatorTensor.jl:129-157
    end
    return f, NQ_SUPER
end
terms_tensor(op::KLocalOperatorTensor) = [op]
terms_tensor(op) = NeuralQuantum.operators(op)                       # KLocalOperatorSum of tensors

mutable struct CudaOperator
    ctx::Ctx
    h::Ptr{Cvoid}
end
function CudaOperator(ctx::Ctx, op, N::Integer)
    f, space = flatten(op)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve f check(ctx, ccall((:nq_operator_create, lib), Cint,
        (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}, Ptr{Cdouble}, Ptr{UInt32}, Cint,
         Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Cvoid}}),
        ctx.h, space, N, length(f.nsites), f.nsites, f.sites, f.rowptr, Ptr{Cdouble}(pointer(f.mel)), f.flip,
        length(f.left), f.left, f.right, h))
    o = CudaOperator(ctx, h[])
    finalizer(o -> (o.h != C_NULL && ccall((:nq_operator_destroy, lib), Cint, (Ptr{Cvoid},), o.h); o.h = C_NULL), o)
    return o
end

"E_loc / L_loc of every configuration: replaces accumulate_connections! + AccumulatorObsScalar (:52-137)."
function local_scalar!(out::AbstractVector, c::CudaNet, op::CudaOperator, σ)
    pr, pc, T, B = _ptrs(σ)
    GC.@preserve σ out check(c.ctx, ccall((:nq_local_scalar, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
        c.h, op.h, pr, pc, nqdtype(T), B, C_NULL, pointer(out)))
    return out
end
"(L_loc, ∇L_loc) of every configuration: replaces the accumulator loop of BatchedGradSampler.jl:87-97."
function local_grad!(out::AbstractVector, ∇out::AbstractMatrix, c::CudaNet, op::CudaOperator, σ)
    pr, pc, T, B = _ptrs(σ)
    GC.@preserve σ out ∇out check(c.ctx, ccall((:nq_local_grad, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64),
        c.h, op.h, pr, pc, nqdtype(T), B, C_NULL, pointer(out), pointer(∇out), size(∇out, 1)))
    return out, ∇out
end

# ---- sampler: Metropolis + LocalRule (Metropolis.jl:54-167, LocalRule.jl:19-28) -----------------------------------
mutable struct CudaSamplerCache <: NeuralQuantum.SamplerCache{MetropolisSampler}
    ctx::Ctx
    h::Ptr{Cvoid}
    loc_chain_length::Int
    steps_done::Int
    doubled::Bool
end
"_sampler_cache(s, v, hilb, net, par_cache) (Metropolis.jl:74-93): chains are sharded over workers, Philox keyed by the global chain id."
function sampler_cache(s::MetropolisSampler{LocalRule}, c::CudaNet, batch_sz::Integer, par)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    offset = worker_rank(par) * batch_sz
    check(c.ctx, ccall((:nq_sampler_create, lib), Cint, (Ptr{Cvoid}, Int64, Cint, UInt64, Int64, Ptr{Ptr{Cvoid}}),
                       c.h, batch_sz, s.passes, UInt64(s.seed), offset, h))
    sc = CudaSamplerCache(c.ctx, h[], Int(ceil(s.chain_length / num_workers(par))), 0, !(c.net isa RBM))
    finalizer(x -> (x.h != C_NULL && ccall((:nq_sampler_destroy, lib), Cint, (Ptr{Cvoid},), x.h); x.h = C_NULL), sc)
    return sc
end
"init_sampler! (Metropolis.jl:101-115): fresh random chains; the burn-in runs inside the first sample_chain! call."
function init_sampler!(s::MetropolisSampler, net::CudaNet, σ, sc::CudaSamplerCache)
    check(sc.ctx, ccall((:nq_sampler_randomize, lib), Cint, (Ptr{Cvoid},), sc.h))
    sc.steps_done = -s.burn_length + 1
    return sc
end
"""
`_sample_state!` (BaseIterativeSampler.jl:5-17): burn + the whole loop over `samplenext!` in ONE launch, filling the
`[N, B, L]` sample array(s).  All L slices are written (quirk Q3 of the reference leaves the last one stale).
"""
function sample_chain!(samples, s::MetropolisSampler, sc::CudaSamplerCache)
    L = sc.loc_chain_length
    σr, σc = samples isa Tuple ? samples : (samples, nothing)
    GC.@preserve σr σc check(sc.ctx, ccall((:nq_sampler_sample, lib), Cint,
        (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
        sc.h, s.burn_length, L, C_NULL, C_NULL, pointer(σr), σc === nothing ? C_NULL : pointer(σc), nqdtype(σr)))
    sc.steps_done = L
    return samples
end
"propose_step! + accept/reject of ONE samplenext! with caller-supplied randomness (parity tests; Metropolis.jl:124-167)."
function samplenext_replay!(sc::CudaSamplerCache, sites::Matrix{Int32}, uniforms::Matrix, accepted::Matrix{UInt8})
    GC.@preserve sites uniforms accepted check(sc.ctx, ccall((:nq_sampler_replay, lib), Cint,
        (Ptr{Cvoid}, Ptr{Int32}, Ptr{Cvoid}, Ptr{UInt8}), sc.h, sites, pointer(uniforms), accepted))
    return accepted
end
function acceptance(sc::CudaSamplerCache)                            # passes_accepted / passes_done (Metropolis.jl:182-191)
    a, d = Ref{Int64}(0), Ref{Int64}(0)
    check(sc.ctx, ccall((:nq_sampler_counters, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), sc.h, a, d))
    return a[] / max(d[], 1)
end

# ---- parallel interface (Parallel/not_parallel.jl, Parallel/MPI/mpi.jl:21-74) --------------------------------------
struct NcclData
    ctx::Ctx
    world_sz::Int
    rank::Int
end
unique_id() = (id = zeros(UInt8, NQ_UNIQUE_ID_BYTES); ccall((:nq_comm_unique_id, lib), Cint, (Ptr{UInt8},), id) == 0 ||
               throw(NQError(Cint(-7), "nq_comm_unique_id")); id)
"One process per GPU: rank 0 makes the id, ships the 128 bytes (MPI.Bcast / any side channel), all call this."
function NcclData(ctx::Ctx, world_sz::Integer, rank::Integer, id::Vector{UInt8})
    check(ctx, ccall((:nq_comm_init, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx.h, world_sz, rank, id))
    return NcclData(ctx, world_sz, rank)
end
num_workers(p::NcclData) = p.world_sz
worker_rank(p::NcclData) = p.rank
worker_rank(::NeuralQuantum.NotParallel) = 0
worker_local_seed(seed, ::NcclData) = seed                           # streams are keyed by global chain id, not by rank
function workers_sum!(target::AbstractArray, p::NcclData)            # mpi.jl:43-46; `target` must be device-resident
    check(p.ctx, ccall((:nq_allreduce_sum, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint),
                       p.ctx.h, pointer(target), length(target), nqdtype(target)))
    return target
end
function workers_mean!(target::AbstractArray, p::NcclData)           # mpi.jl:29-34
    check(p.ctx, ccall((:nq_allreduce_mean, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint),
                       p.ctx.h, pointer(target), length(target), nqdtype(target)))
    return target
end
workers_mean!(target, data, p::NcclData) = workers_mean!(mean!(target, data), p)      # mpi.jl:21-27

# ---- algorithm interface (batched_algorithms.jl:1-11, SR/SRDirect.jl, SR/SRIterative.jl) --------------------------
mutable struct CudaSRCache{TS,TF}
    ctx::Ctx
    S::TS                   # [P, P] (explicit) or nothing (matrix-free)
    F::TF
    Δw::TF
    O                       # centred O kept for the matrix-free solvers
    Ns_total::Int
    real_params::Bool
    iters::Int
    converged::Bool
end
function algorithm_cache(algo::SR, ctx::Ctx, c::CudaNet)
    T = eltype(first(trainable(c.net)))
    explicit = algo.algorithm == NeuralQuantum.sr_cholesky || algo.use_fullmat
    S = explicit ? zeros(T, c.P, c.P) : nothing
    return CudaSRCache(ctx, S, zeros(T, c.P), zeros(T, c.P), nothing, 0, T <: Real, 0, true)
end
"_center_gradient! (BaseIterativeSampler.jl:19-26): <O> over all workers, O centred in place."
function center_gradient!(ctx::Ctx, O::AbstractMatrix, avg::AbstractVector)
    GC.@preserve O avg check(ctx, ccall((:nq_center, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Cint, Ptr{Cvoid}),
        ctx.h, pointer(O), size(O, 1), size(O, 1), size(O, 2), nqdtype(O), pointer(avg)))
    return avg
end
"_compute_gradient! for kets (BatchedValSampler.jl:97-115): F_k = <E_loc conj(O_k - <O_k>)>."
function force_ket!(∇C::AbstractVector, ctx::Ctx, Oc::AbstractMatrix, Eloc::AbstractVector)
    GC.@preserve ∇C Oc Eloc check(ctx, ccall((:nq_force_ket, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
        ctx.h, pointer(Oc), size(Oc, 1), size(Oc, 1), size(Oc, 2), nqdtype(Oc), pointer(Eloc), pointer(∇C)))
    return ∇C
end
"Liouvillian force (BatchedGradSampler.jl:108-118); returns the cost <|L_loc|^2>."
function force_liouvillian!(∇C::AbstractVector, ctx::Ctx, Lloc::AbstractVector, ∇Lloc::AbstractMatrix, avg::AbstractVector)
    cost = Ref{Cdouble}(0)
    GC.@preserve ∇C Lloc ∇Lloc avg check(ctx, ccall((:nq_force_liouvillian, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}),
        ctx.h, pointer(Lloc), pointer(∇Lloc), size(∇Lloc, 1), size(∇Lloc, 1), size(∇Lloc, 2), nqdtype(∇Lloc),
        pointer(avg), pointer(∇C), cost))
    return cost[]
end
"setup_algorithm!(cache, ∇C, O, par) (SRDirect.jl:26-49, SRIterative.jl:32-62): S and F from the centred O."
function setup_algorithm!(g::CudaSRCache, ∇C::AbstractVector, Oc::AbstractMatrix, par)
    P, Ns = size(Oc)
    g.Ns_total = Ns * num_workers(par)
    g.O = Oc
    if g.S === nothing                                                # matrix-free: only F (SR_notfull.jl:47-57)
        g.F .= g.real_params ? real.(∇C) : ∇C
        return g
    end
    GC.@preserve Oc ∇C g check(g.ctx, ccall((:nq_sr_setup, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
        g.ctx.h, pointer(Oc), P, P, Ns, g.Ns_total, nqdtype(Oc), pointer(∇C), g.real_params, pointer(g.S), pointer(g.F)))
    num_workers(par) > 1 && workers_sum!(g.S, par)                    # partial S normalised by the GLOBAL Ns (quirk Q5)
    return g
end
solver_code(algo::SR) = algo.algorithm == NeuralQuantum.sr_cholesky ? NQ_SOLVE_CHOLESKY :
                        algo.algorithm == NeuralQuantum.sr_cg ? NQ_SOLVE_CG :
                        algo.algorithm == NeuralQuantum.sr_minres ? NQ_SOLVE_MINRES : NQ_SOLVE_QLP
function _solve!(g::CudaSRCache, code::Cint, ϵ::Float64, tol::Float64)
    its = Ref{Int64}(0)
    P = length(g.F)
    st = if g.S !== nothing
        GC.@preserve g ccall((:nq_sr_solve, lib), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Cdouble, Cint, Cdouble, Int64, Ptr{Cvoid}, Ptr{Int64}),
            g.ctx.h, pointer(g.S), pointer(g.F), P, nqdtype(g.F), ϵ, code, tol, 0, pointer(g.Δw), its)
    else
        O = g.O
        GC.@preserve g O ccall((:nq_sr_solve_matfree_algo, lib), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Cint, Ptr{Cvoid}, Cint, Cdouble, Cint, Cdouble, Int64,
             Ptr{Cvoid}, Ptr{Int64}),
            g.ctx.h, pointer(O), P, P, size(O, 2), g.Ns_total, nqdtype(O), pointer(g.F), g.real_params, ϵ, code, tol, 0,
            pointer(g.Δw), its)
    end
    g.iters += its[]
    return st
end
"""
precondition!(cache, algo, iter) (SRDirect.jl:51-90, SRIterative.jl:71-153): (S + ϵ I) Δw = F, or the multiplicative
regulariser S + λ Diagonal(diag S).  Iterative solvers: on non-convergence up to 5 warm-started MINRES-QLP runs with the
default tolerance, then Δw = 0 (SRIterative.jl:133-150).
"""
function precondition!(g::CudaSRCache, algo::SR, iter_n)
    ϵ = Float64(algo.sr_diag_shift)                                   # stored as Float32 by SR() (quirk Q16)
    if algo.precondition_type == NeuralQuantum.sr_multiplicative
        λ = Float64(max(algo.λ0 * algo.b^iter_n, algo.λmin))
        GC.@preserve g check(g.ctx, ccall((:nq_sr_scale_diagonal, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Cdouble),
                                          g.ctx.h, pointer(g.S), length(g.F), nqdtype(g.F), λ))
        ϵ = 0.0
    elseif algo.precondition_type == NeuralQuantum.sr_none
        ϵ = 0.0
    end
    g.iters = 0
    st = _solve!(g, solver_code(algo), ϵ, Float64(algo.sr_precision))
    add_iters = 1
    while st == NQ_ERR_NOT_CONVERGED && add_iters <= 5
        println("minresqlp not conerged. Additional $(length(g.F)*10) iters for the $add_iters time.")
        st = _solve!(g, NQ_SOLVE_QLP_WARM, ϵ, sqrt(eps(real(eltype(g.F)))))
        add_iters += 1
    end
    g.converged = st == 0
    if st == NQ_ERR_NOT_CONVERGED
        g.Δw .= 0
    else
        check(g.ctx, st)
    end
    return g.Δw
end

# ---- transition rules other than LocalRule (MCMCRules/ExchangeRule.jl, Nagy.jl, OperatorRule.jl) --------------------
const NQ_RULE_LOCAL, NQ_RULE_EXCHANGE, NQ_RULE_NAGY, NQ_RULE_OPERATOR = Cint.(0:3)
couples(pairs) = Int32[x - 1 for p in pairs for x in p]                  # 1-based (i, j) tuples -> 0-based flat [n][2]
function set_rule!(sc::CudaSamplerCache, r::NeuralQuantum.ExchangeRule)
    cp = couples(r.distances)
    check(sc.net.ctx, ccall((:nq_sampler_set_rule, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Cvoid}),
                            sc.h, NQ_RULE_EXCHANGE, length(r.distances), cp, C_NULL))
end
function set_rule!(sc::CudaSamplerCache, r::NeuralQuantum.NagyRule)
    cp = couples(r.adjacency_list)
    check(sc.net.ctx, ccall((:nq_sampler_set_rule, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Cvoid}),
                            sc.h, NQ_RULE_NAGY, length(r.adjacency_list), cp, C_NULL))
end
function set_rule!(sc::CudaSamplerCache, r::NeuralQuantum.OperatorRule, op::CudaOperator)
    check(sc.net.ctx, ccall((:nq_sampler_set_rule, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Cvoid}),
                            sc.h, NQ_RULE_OPERATOR, 0, C_NULL, op.h))
end
# draws [4, B, passes] Int32: the integers the rule draws per proposal (see nq_sampler_replay_rule)
function samplenext_replay_rule!(sc::CudaSamplerCache, draws::Array{Int32,3}, uniforms::Matrix, accepted::Matrix{UInt8})
    check(sc.net.ctx, ccall((:nq_sampler_replay_rule, lib), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Cvoid}, Ptr{UInt8}),
                            sc.h, draws, uniforms, accepted))
end

# ---- full-space tools: ket / densitymatrix (utils/densitymatrix.jl:9-62), ExactSampler (Samplers/Exact.jl:135-181) ----
function fullspace_size(c::CudaNet)
    n = Ref{Int64}(0)
    check(c.ctx, ccall((:nq_fullspace_size, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}), c.h, n))
    return Int(n[])
end
function NeuralQuantum.ket(c::CudaNet{<:RBM}, hilb, norm = true)
    ψ = zeros(out_type(c.net), fullspace_size(c))
    check(c.ctx, ccall((:nq_fullspace_state, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), c.h, norm ? 1 : 0, ψ))
    return ψ
end
function NeuralQuantum.densitymatrix(c::CudaNet{<:Union{RBMSplit,NDM}}, hilb, norm = true)
    D = 1 << NeuralQuantum.nsites(NeuralQuantum.physical(hilb))
    ρ = zeros(out_type(c.net), D, D)
    check(c.ctx, ccall((:nq_fullspace_state, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), c.h, norm ? 1 : 0, ρ))
    return ρ
end
# init_sampler!(::ExactSampler): the cumulative table; samplenext!: basis numbers for L slots x B chains
function exact_table!(pdf::Vector{Float64}, c::CudaNet)
    check(c.ctx, ccall((:nq_exact_table, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.h, pdf))
    return pdf
end
function exact_sample!(indices::Matrix{Int64}, prow::Matrix{UInt64}, pcol, c::CudaNet, pdf::Vector{Float64}; seed, chain_offset = 0,
                       draw_base = 0, uniforms = nothing)
    B, L = size(indices)
    check(c.ctx, ccall((:nq_exact_sample, lib), Cint,
                       (Ptr{Cvoid}, Ptr{Float64}, UInt64, Int64, UInt64, Int64, Int64, Ptr{Float64}, Ptr{UInt64}, Ptr{UInt64}, Ptr{Int64}),
                       c.h, pdf, seed, chain_offset, draw_base, B, L, uniforms === nothing ? C_NULL : uniforms, prow,
                       pcol === nothing ? C_NULL : pcol, indices))
    return indices
end

# ---- NDMSymm (Networks/MixedDensityMatrix/NDMSymm.jl, NDMSymmBatched.jl) ---------------------------------------------
# The gather lists are the rows of the reference's own 0/1 matrices (net.∇b_mat ... net.∇u_mat) and set_bare_params!'s
# index map; they are built once from those matrices.
mutable struct CudaSymm
    ctx::Ctx
    h::Ptr{Cvoid}
    bare::CudaNet
    Ps::Int
end
function CudaSymm(bare::CudaNet, ptr::Vector{Int64}, idx::Vector{Int32}, scale::Vector{Float64}, src::Vector{Int32},
                  avg_ranges::Matrix{Int64})
    h = Ref{Ptr{Cvoid}}(C_NULL)
    Ps = length(scale)
    check(bare.ctx, ccall((:nq_symm_create, lib), Cint,
                          (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Cint, Ptr{Int64}, Ptr{Ptr{Cvoid}}),
                          bare.h, Ps, ptr, idx, scale, src, size(avg_ranges, 2), avg_ranges, h))
    g = CudaSymm(bare.ctx, h[], bare, Ps)
    finalizer(g -> (g.h != C_NULL && ccall((:nq_symm_destroy, lib), Cint, (Ptr{Cvoid},), g.h); g.h = C_NULL), g)
    return g
end
set_params!(g::CudaSymm, w::Vector) = check(g.ctx, ccall((:nq_symm_set_params, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), g.h, w, length(w)))
get_params!(w::Vector, g::CudaSymm) = (check(g.ctx, ccall((:nq_symm_get_params, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), g.h, w, length(w))); w)
update!(g::CudaSymm, Δw::AbstractVector, η::Real) =                     # NDMSymm.jl:27-30: step, then set_bare_params!
    check(g.ctx, ccall((:nq_symm_update, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble), g.h, Δw, η))
function symmetrize_∇logψ!(∇symm::AbstractMatrix, g::CudaSymm, ∇bare::AbstractMatrix)    # NDMSymmBatched.jl:22-36
    check(g.ctx, ccall((:nq_symm_gradient, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Cint, Ptr{Cvoid}, Int64),
                       g.h, ∇bare, size(∇bare, 1), size(∇bare, 2), nqdtype(∇bare), ∇symm, size(∇symm, 1)))
    return ∇symm
end

# ---- streaming S assembly (O never materialised for the whole batch) and the Nesterov step ---------------------------
function sr_accumulate!(ctx::Ctx, Sacc, sumO, Ochunk::AbstractMatrix, Ns_total::Integer, real_params::Bool, first::Bool)
    P, Nc = size(Ochunk)
    check(ctx, ccall((:nq_sr_accumulate, lib), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                     ctx.h, Ochunk, P, P, Nc, Ns_total, nqdtype(Ochunk), real_params ? 1 : 0, Sacc, sumO, first ? 1 : 0))
end
sr_finish!(ctx::Ctx, Sacc, sumO, P::Integer, Ns_total::Integer, T::Type, real_params::Bool) =
    check(ctx, ccall((:nq_sr_finish, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Cint, Cint),
                     ctx.h, Sacc, sumO, P, Ns_total, nqdtype(T), real_params ? 1 : 0))
# centring with the subtraction deferred to the S assembly (returns true when O was left uncentred), and its completion
function center_gradient_lazy!(ctx::Ctx, O::AbstractMatrix, avg::AbstractVector)
    flag = Ref{Cint}(0)
    check(ctx, ccall((:nq_center_lazy, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Cint, Ptr{Cvoid}, Ptr{Cint}),
                     ctx.h, O, size(O, 1), size(O, 1), size(O, 2), nqdtype(O), avg, flag))
    return flag[] != 0
end
center_finish!(ctx::Ctx, O::AbstractMatrix) =
    check(ctx, ccall((:nq_center_finish, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Cint),
                     ctx.h, O, size(O, 1), size(O, 1), size(O, 2), nqdtype(O)))
# structure of the gradient rows known to the caller (NDM: lambda rows real, mu rows imaginary): one-shot hint for the next setup
sr_hint_row_planes!(ctx::Ctx, planes::Vector{UInt8}) =
    check(ctx, ccall((:nq_sr_hint_row_planes, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64), ctx.h, planes, length(planes)))
# Optimisers.apply(o::Nesterov, x, Δ, state) on device vectors (rules.jl:44-55): returns -d in `delta`
nesterov!(ctx::Ctx, delta, velocity, Δ, n::Integer, T::Type, o) =
    check(ctx, ccall((:nq_nesterov, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Cdouble, Cdouble, Ptr{Cvoid}),
                     ctx.h, velocity, Δ, n, nqdtype(T), o.lr, o.μ, delta))

# One iteration step from HOST configurations (σ, σ′ as the reference's sampler holds them, [N, Ns] floats): copy, packing and the
# fused kernel software-pipelined inside the library.  prow / pcol / logρ / O / L_loc / ∇L_loc are device pointers (CuPtr or
# Ptr from the caller's allocator); replaces the two loops of sample!(is::BatchedGradSampler) (BatchedGradSampler.jl:80-97).
function logpsi_grad_local_host!(c::CudaNet, op::CudaOperator, σr::Matrix{T}, σc::Union{Matrix{T},Nothing},
                                 prow, pcol, logρ, O, ldO::Integer, Lloc, ∇Lloc, ld::Integer) where {T<:AbstractFloat}
    Ns = size(σr, 2)
    GC.@preserve σr σc check(c.ctx, ccall((:nq_logpsi_grad_local_host, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64,
         Ptr{Cvoid}, Ptr{Cvoid}, Int64),
        c.h, op.h, pointer(σr), σc === nothing ? C_NULL : pointer(σc), nqdtype(T), Ns, prow, pcol === nothing ? C_NULL : pcol,
        logρ, O, ldO, Lloc, ∇Lloc === nothing ? C_NULL : ∇Lloc, ld))
end

# ---- device-resident entry points and utilities (every remaining symbol of include/nqcuda.h) -------------------------
# The iteration keeps its arrays on the device between the calls: configurations as bit-packed words (what the sampler
# leaves behind), log ρ / O / L_loc / ∇L_loc in the caller's device buffers.  `DevPtr` is whatever the caller's allocator
# hands out (CUDA.jl: `pointer(::CuArray)` reinterpreted to Ptr{Cvoid}).
const DevPtr = Ptr{Cvoid}
version() = ccall((:nq_version, lib), Cint, ())
status_string(st::Integer) = unsafe_string(ccall((:nq_status_string, lib), Cstring, (Cint,), st))
sync(ctx::Ctx) = check(ctx, ccall((:nq_ctx_sync, lib), Cint, (Ptr{Cvoid},), ctx.h))
function launch_count(ctx::Ctx)
    n = Ref{UInt64}(0)
    check(ctx, ccall((:nq_ctx_launch_count, lib), Cint, (Ptr{Cvoid}, Ptr{UInt64}), ctx.h, n))
    return n[]
end
states_words(N::Integer) = ccall((:nq_states_words, lib), Cint, (Cint,), N)
"σ [N, B] floats (host or device) → packed words [W64, B] on the device (States.jl:12-32 ↔ HomogeneousSpin.jl:156-179)."
pack_states!(packed::DevPtr, ctx::Ctx, hilb, σ::AbstractMatrix{T}) where {T} =
    GC.@preserve σ check(ctx, ccall((:nq_pack_states, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                                    ctx.h, hilbcode(hilb), size(σ, 1), size(σ, 2), pointer(σ), nqdtype(T), packed))
unpack_states!(σ::AbstractMatrix{T}, ctx::Ctx, hilb, packed::DevPtr) where {T} =
    GC.@preserve σ check(ctx, ccall((:nq_unpack_states, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                                    ctx.h, hilbcode(hilb), size(σ, 1), size(σ, 2), packed, pointer(σ), nqdtype(T)))
function nparams(c::CudaNet)
    P = Ref{Int64}(0)
    check(c.ctx, ccall((:nq_machine_nparams, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}), c.h, P))
    return Int(P[])
end
function out_dtype(c::CudaNet)                                        # out_type(net)
    d = Ref{Cint}(0)
    check(c.ctx, ccall((:nq_machine_out_dtype, lib), Cint, (Ptr{Cvoid}, Ptr{Cint}), c.h, d))
    return d[]
end
function nparams(g::CudaSymm)
    P = Ref{Int64}(0)
    check(g.ctx, ccall((:nq_symm_nparams, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}), g.h, P))
    return Int(P[])
end
"logψ! / logψ_and_∇logψ! on packed device configurations (no conversion pass)."
logpsi_packed!(out::DevPtr, c::CudaNet, prow::DevPtr, pcol::DevPtr, B::Integer) =
    check(c.ctx, ccall((:nq_logpsi_packed, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}), c.h, prow, pcol, B, out))
logpsi_grad_packed!(out::DevPtr, O::DevPtr, ldO::Integer, c::CudaNet, prow::DevPtr, pcol::DevPtr, B::Integer) =
    check(c.ctx, ccall((:nq_logpsi_grad_packed, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                       c.h, prow, pcol, B, out, O, ldO))
"The fused iteration step on the samples the sampler left on the device (BatchedGradSampler.jl:83-97 / BatchedValSampler.jl:122-125)."
logpsi_grad_local_packed!(logψ::DevPtr, O::DevPtr, ldO::Integer, Lloc::DevPtr, ∇Lloc::DevPtr, ld::Integer, c::CudaNet,
                          op::CudaOperator, prow::DevPtr, pcol::DevPtr, B::Integer) =
    check(c.ctx, ccall((:nq_logpsi_grad_local_packed, lib), Cint,
                       (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                       c.h, op.h, prow, pcol, B, logψ, O, ldO, Lloc, ∇Lloc, ld))
local_scalar_packed!(Lloc::DevPtr, c::CudaNet, op::CudaOperator, prow::DevPtr, pcol::DevPtr, B::Integer) =
    check(c.ctx, ccall((:nq_local_scalar_packed, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}),
                       c.h, op.h, prow, pcol, B, Lloc))
local_grad_packed!(Lloc::DevPtr, ∇Lloc::DevPtr, ld::Integer, c::CudaNet, op::CudaOperator, prow::DevPtr, pcol::DevPtr, B::Integer) =
    check(c.ctx, ccall((:nq_local_grad_packed, lib), Cint,
                       (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                       c.h, op.h, prow, pcol, B, Lloc, ∇Lloc, ld))
function max_connections(op::CudaOperator)
    n = Ref{Int64}(0)
    check(op.ctx, ccall((:nq_operator_max_connections, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}), op.h, n))
    return Int(n[])
end
"""
row_valdiff! over a batch (BaseOperators.jl:25-36, KLocalOperator.jl:151-159): per configuration the ordered connection list
of the reference -- counts [B], matrix elements [max_conn, B], flip masks [W64, max_conn, B] (integer-parity artefacts).
"""
function connections(op::CudaOperator, hilb, σr::Matrix{T}, σc::Union{Matrix{T},Nothing} = nothing) where {T}
    N, B = size(σr); mc = max_connections(op); W = Int(states_words(N))
    counts = zeros(Int32, B); mels = zeros(ComplexF64, mc, B)
    fr = zeros(UInt64, W, mc, B); fc = σc === nothing ? nothing : zeros(UInt64, W, mc, B)
    GC.@preserve σr σc counts mels fr fc check(op.ctx, ccall((:nq_connections, lib), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Int32}, Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt64}),
        op.h, hilbcode(hilb), pointer(σr), σc === nothing ? C_NULL : pointer(σc), nqdtype(T), B, mc, counts, mels, fr,
        fc === nothing ? C_NULL : pointer(fc)))
    return counts, mels, fr, fc
end
"Chain state in and out (parity tests; Metropolis.jl:101-115 writes it with rand!)."
set_state!(sc::CudaSamplerCache, σr::Matrix{T}, σc = nothing) where {T} =
    GC.@preserve σr σc check(sc.ctx, ccall((:nq_sampler_set_state, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                                           sc.h, pointer(σr), σc === nothing ? C_NULL : pointer(σc), nqdtype(T)))
get_state!(σr::Matrix{T}, σc, sc::CudaSamplerCache) where {T} =
    GC.@preserve σr σc check(sc.ctx, ccall((:nq_sampler_get_state, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                                           sc.h, pointer(σr), σc === nothing ? C_NULL : pointer(σc), nqdtype(T)))
"diagonal = true: the chain of the observables sampler over ρ(σ, σ) (BatchedObsDMSampler.jl:59-105)."
set_mode!(sc::CudaSamplerCache, diagonal::Bool) =
    check(sc.ctx, ccall((:nq_sampler_set_mode, lib), Cint, (Ptr{Cvoid}, Cint), sc.h, diagonal ? 1 : 0))
"stat_analysis (utils/stats.jl:26-84) of [B, L] values on the device: (mean, error, variance, τ, R̂), global under sharding."
function stat_analysis(ctx::Ctx, vals::DevPtr, B::Integer, L::Integer, T::Type)
    out = zeros(Float64, 6)
    check(ctx, ccall((:nq_stat_analysis, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Cint, Ptr{Float64}), ctx.h, vals, B, L, nqdtype(T), out))
    return (mean = complex(out[1], out[2]), error = out[3], variance = out[4], tau = out[5], R = out[6])
end
"Matrix-free CG on the centred rows (SR_notfull.jl:47-148, SRIterative.jl:117-132): the plain-CG form of _solve! without an explicit S."
function solve_matfree_cg!(Δw::Vector, ctx::Ctx, Oc::AbstractMatrix, Ns_total::Integer, F::Vector, real_params::Bool,
                           ϵ::Float64, tol::Float64, maxiter::Integer = 0)
    its = Ref{Int64}(0)
    P = size(Oc, 1)
    GC.@preserve Δw Oc F check(ctx, ccall((:nq_sr_solve_matfree, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Cint, Ptr{Cvoid}, Cint, Cdouble, Cdouble, Int64, Ptr{Cvoid}, Ptr{Int64}),
        ctx.h, pointer(Oc), P, P, size(Oc, 2), Ns_total, nqdtype(Oc), pointer(F), real_params ? 1 : 0, ϵ, tol, maxiter,
        pointer(Δw), its))
    return Δw, Int(its[])
end
"abs2.(local_vals) (BatchedGradSampler.jl:99) on the device."
abs2!(out::DevPtr, ctx::Ctx, vals::DevPtr, n::Integer, T::Type) =
    check(ctx, ccall((:nq_abs2, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Ptr{Cvoid}), ctx.h, vals, n, nqdtype(T), out))
function comm_size(ctx::Ctx)                                          # num_workers / worker rank (mpi.jl:36-51)
    n, r = Ref{Cint}(1), Ref{Cint}(0)
    check(ctx, ccall((:nq_comm_size, lib), Cint, (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}), ctx.h, n, r))
    return Int(n[]), Int(r[])
end
comm_destroy!(ctx::Ctx) = check(ctx, ccall((:nq_comm_destroy, lib), Cint, (Ptr{Cvoid},), ctx.h))
"Uneven shards (Metropolis.jl:75 shards the chain length with ceil): the global sample count the normalisations use."
set_global_samples!(ctx::Ctx, ns_total::Integer) =
    check(ctx, ccall((:nq_comm_set_global_samples, lib), Cint, (Ptr{Cvoid}, Int64), ctx.h, ns_total))

end # module

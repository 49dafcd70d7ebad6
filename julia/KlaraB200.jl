# KlaraB200.jl -- thin Julia shim over libklara_b200.so (include/klara_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain.  It is kept deliberately
# thin (argument marshalling + ccall only) so that it can be reviewed against the header line by line.
# Written for Julia >= 1.6.  It mirrors the exported names of Klara.jl for the MCMC hot path:
#
#   BasicContMuvParameter(:p, logtarget=IsoGaussian())      src/variables/parameters/BasicContMuvParameter.jl:383
#   likelihood_model(p, false)                               src/models/generators.jl:18
#   MH(sigma::Vector), MALA(driftstep), HMC(leapstep, nleaps)   src/samplers/{MH,MALA,HMC}.jl
#   BasicMCRange(nsteps=, burnin=, thinning=)                src/ranges/BasicMCRange.jl:33
#   VanillaMCTuner(), AcceptanceRateMCTuner(rate), DualAveragingMCTuner(rate, nadapt)   src/tuners/
#   Hyperparameter(:λ), Data(:X) vertices whose v0 values reach the target (BayesLogit)   src/variables/variables.jl
#   BasicMCJob(model, sampler, mcrange, v0; tuner, outopts), run, reset, output   src/jobs/BasicMCJob.jl
#
# Batch extension: v0 = Dict(:p => Matrix{Float64}(d, nchains)) (one column per chain).  Stock Klara has no
# method for a Matrix initial value on a BasicContMuvParameter, so this hook does not conflict.
module KlaraB200

using LinearAlgebra: dot
import Statistics: mean      # Klara adds methods to mean (src/stats/mean.jl:7-11); so does the shim

export IsoGaussian, ShiftedIsoGaussian, Rosenbrock, DenseGaussian, BayesLogit, Hyperparameter, Data, DualAveragingMCTuner, run_host, GenericModel,
       SyntheticNormal, seek!, gathered, logistic_rate_score, erf_rate_score, ess, mcvar, mcse, iact, acceptance, BasicContMuvParameter, likelihood_model, MH, MALA, HMC, NUTS,
       BasicMCRange, VanillaMCTuner, AcceptanceRateMCTuner, BasicMCJob, run, reset, output

const LIB = get(ENV, "KLARA_B200_LIB", "libklara_b200.so")

# ---- target descriptors: also plain callables, so the same definition runs on stock Klara ----------------
abstract type Target end
struct IsoGaussian <: Target end
(::IsoGaussian)(z::Vector{Float64}) = -sum(abs2, z)
gradient(::IsoGaussian) = z -> -2z
struct ShiftedIsoGaussian <: Target; mu::Vector{Float64}; end
(t::ShiftedIsoGaussian)(z::Vector{Float64}) = -sum(abs2, z .- t.mu)
gradient(t::ShiftedIsoGaussian) = z -> -2 .* (z .- t.mu)
struct Rosenbrock <: Target; a::Float64; b::Float64; scale::Float64; end
Rosenbrock() = Rosenbrock(1.0, 100.0, 0.05)
# -z'Cz, -2Cz with a symmetric precision matrix (doc/examples/BivariateNormal/MALA/function/analytical.jl:8-9)
# DenseGaussian() is bound by BasicMCJob from the model's Hyperparameter(:C) vertex (v[1] of the reference's closures,
# doc/examples/BivariateNormal/MALA/function/analytical.jl:4-21)
mutable struct DenseGaussian <: Target; C::Matrix{Float64}; end
DenseGaussian() = DenseGaussian(zeros(0, 0))
(t::DenseGaussian)(z::Vector{Float64}) = -dot(z, t.C*z)
gradient(t::DenseGaussian) = z -> -2 .* (t.C*z)
# Bayesian logistic regression, N(0, λI) prior: the closures of doc/examples/swiss/HMC/noadaptation/analytical.jl:11-20.
# BayesLogit() is bound by BasicMCJob from v0[:λ], v0[:X], v0[:y] (the model's other vertices, in vertex order).
mutable struct BayesLogit <: Target; lambda::Float64; X::Matrix{Float64}; y::Vector{Float64}; end
BayesLogit() = BayesLogit(100.0, zeros(0, 0), zeros(0))
loglikelihood(t::BayesLogit) = p -> (Xp = t.X*p; dot(Xp, t.y) - sum(log.(1 .+ exp.(Xp))))
logprior(t::BayesLogit) = p -> -0.5*(dot(p, p)/t.lambda + length(p)*log(2*pi*t.lambda))
(t::BayesLogit)(p::Vector{Float64}) = loglikelihood(t)(p) + logprior(t)(p)
gradient(t::BayesLogit) = p -> t.X'*(t.y .- 1 ./ (1 .+ exp.(-t.X*p))) .- p ./ t.lambda
code(::IsoGaussian) = 0; code(::ShiftedIsoGaussian) = 1; code(::DenseGaussian) = 2; code(::Rosenbrock) = 3; code(::BayesLogit) = 4
struct Hyperparameter; key::Symbol; end
const Data = Hyperparameter

struct BasicContMuvParameter; key::Symbol; logtarget::Target; end
BasicContMuvParameter(key::Symbol; logtarget::Target, gradlogtarget=nothing, nkeys::Int=0) = BasicContMuvParameter(key, logtarget)
# GenericModel(vs; isindexed=false): vertices kept in the given order (src/models/GenericModel.jl:94-119); the job only
# uses the graph to find the parameter and the other vertices' values in vertex order
struct GenericModel; vertices::Vector{Any}; end
GenericModel(v::Vector; isindexed::Bool=true, isdirected::Bool=true) = GenericModel(Any[v...])
likelihood_model(p::BasicContMuvParameter, isindexed::Bool=true) = GenericModel(Any[p])
likelihood_model(v::Vector; isindexed::Bool=true, isdirected::Bool=true) = GenericModel(Any[v...])
# initial value generated on the device from the job's Philox streams (seed, global chain, transition 0)
struct SyntheticNormal; dim::Int; nchains::Int; end

struct MH; sigma::Vector{Float64}; end
struct MALA; driftstep::Float64; MALA(s=1.0) = (@assert s > 0 "Drift step is not positive"; new(s)); end
struct HMC
  leapstep::Float64; nleaps::Int
  function HMC(leapstep=0.1, nleaps=10)
    @assert leapstep > 0 "Leapfrog step is not positive"
    @assert nleaps > 0 "Number of leapfrog steps is not positive"
    new(leapstep, nleaps)
  end
end
# src/samplers/NUTS.jl:228-241 (the multivariate transition as the reference computes it; Vanilla or DualAveraging tuner)
struct NUTS
  leapstep::Float64; maxδ::Int; maxndoublings::Int
  function NUTS(leapstep=0.1; maxδ::Integer=1000, maxndoublings::Integer=5)
    @assert leapstep > 0 "Leapfrog step is not positive"
    @assert maxδ > 0 "maxδ is not positive"
    @assert maxndoublings > 0 "Maximum number of doublings is not positive"
    new(leapstep, maxδ, maxndoublings)
  end
end
struct BasicMCRange; burnin::Int; thinning::Int; nsteps::Int; npoststeps::Int; end
function BasicMCRange(; burnin::Int=0, thinning::Int=1, nsteps::Int=100)
  @assert burnin >= 0 "Number of burn-in iterations should be non-negative"
  @assert thinning >= 1 "Thinning should be >= 1"
  @assert nsteps > burnin "Total number of MCMC iterations should be greater than number of burn-in iterations"
  BasicMCRange(burnin, thinning, nsteps, length((burnin+1):thinning:nsteps))
end
struct VanillaMCTuner; period::Int; verbose::Bool; end
VanillaMCTuner(; period::Int=100, verbose::Bool=false) = VanillaMCTuner(period, verbose)
logistic_rate_score(x::Real, k::Real=7.) = 2/(1+exp(-k*x))          # src/tuners/AcceptanceRateMCTuner.jl:9
erf_rate_score(x::Real, k::Real=3.) = ccall((:erf, "libm"), Float64, (Float64,), k*x)+1   # src/tuners/AcceptanceRateMCTuner.jl:17
struct AcceptanceRateMCTuner; targetrate::Float64; score::Int32; k::Float64; period::Int; verbose::Bool; end
function AcceptanceRateMCTuner(rate; score::Function=logistic_rate_score, k=nothing, period::Int=100, verbose::Bool=false)
  score === logistic_rate_score || score === erf_rate_score || error("score must be logistic_rate_score or erf_rate_score")
  iserf = score === erf_rate_score
  AcceptanceRateMCTuner(rate, iserf ? 1 : 0, k === nothing ? (iserf ? 3.0 : 7.0) : k, period, verbose)
end
# src/tuners/DualAveragingMCTuner.jl:53-93 (HMC only; per-chain step and nleaps = max(1, round(λ/step)))
struct DualAveragingMCTuner
  targetrate::Float64; nadapt::Int; ε0bar::Float64; h0bar::Float64; γ::Float64; t0::Int; κ::Float64; period::Int; verbose::Bool
end
DualAveragingMCTuner(rate, nadapt; ε0bar=1.0, h0bar=0.0, γ=0.05, t0::Int=10, κ=0.75, period::Int=100, verbose::Bool=false) =
  DualAveragingMCTuner(rate, nadapt, ε0bar, h0bar, γ, t0, κ, period, verbose)
tunercode(::VanillaMCTuner) = 0; tunercode(::AcceptanceRateMCTuner) = 1; tunercode(::DualAveragingMCTuner) = 2

# ---- klb_config, field for field (include/klara_b200.h) ---------------------------------------------------
struct KlbConfig
  struct_size::UInt32; sampler::Int32; target::Int32; tuner::Int32; arith::Int32
  nchains::Int64; dim::Int64; nsteps::Int64; burnin::Int64; thinning::Int64
  step::Float64; nleaps::Int32
  target_rate::Float64; score_k::Float64; period::Int64
  verbose::Int32; monitor::UInt32; diagnostics::UInt32; destination::Int32
  seed::UInt64; chain_offset::Int64; device::Int32; score::Int32
  da_nadapt::Int64; da_t0::Int64; da_eps0bar::Float64; da_h0bar::Float64; da_gamma::Float64; da_kappa::Float64
  nuts_maxdelta::Int32; nuts_maxndoublings::Int32
end
struct KlbHostField; field::Int32; reserved::Int32; host_dst::Ptr{Cvoid}; nbytes::Int64; end

lasterror() = unsafe_string(ccall((:klb_last_error, LIB), Cstring, ()))
check(rc::Cint) = rc == 0 ? nothing : error("klara_b200 error $rc: $(lasterror())")

# handle: klb_job, or klb_multi when the chains are sharded over `ngpus` devices of this process (multi = true)
mutable struct BasicMCJob
  handle::Ptr{Cvoid}; nchains::Int; dim::Int; range::BasicMCRange; monitor::Vector{Symbol}; diagnostics::Vector{Symbol}
  multi::Bool
end
sym(job::BasicMCJob, name::String) = Symbol(job.multi ? "klb_multi_" : "klb_job_", name)

function BasicMCJob(model::GenericModel, sampler, range::BasicMCRange, v0::Dict;
                    tuner=VanillaMCTuner(), outopts::Dict=Dict{Symbol,Any}(), seed::Integer=0,
                    arith::Symbol=:reference, device::Integer=0, chain_offset::Integer=0, ngpus::Integer=1)
  pidx = findfirst(v -> v isa BasicContMuvParameter, model.vertices)
  p = model.vertices[pidx]::BasicContMuvParameter
  if length(model.vertices) > 1       # hyper-parameters / data reach the target in vertex order (BasicContMuvParameter.jl:497-501)
    vals = [v0[v.key] for (i, v) in enumerate(model.vertices) if i != pidx]
    t = p.logtarget
    if t isa BayesLogit
      t.lambda, t.X, t.y = Float64(vals[1]), Matrix{Float64}(vals[2]), Vector{Float64}(vals[3])
    elseif t isa DenseGaussian
      t.C = Matrix{Float64}(vals[1])                                    # v[1] = the Hyperparameter(:C) vertex
    else
      error("target $(typeof(t)) takes no hyper-parameters")
    end
  end
  x0 = v0[p.key]
  synthetic = x0 isa SyntheticNormal
  if !synthetic
    x0 = x0 isa Vector ? reshape(Float64.(x0), :, 1) : Matrix{Float64}(x0)   # d x nchains
  end
  d, n = synthetic ? (x0.dim, x0.nchains) : size(x0)
  monitor = get(outopts, :monitor, [:value]); diags = get(outopts, :diagnostics, Symbol[])
  dest = get(outopts, :destination, :nstate)
  mon = UInt32(sum(Dict(:value=>1, :logtarget=>2, :gradlogtarget=>4)[m] for m in monitor; init=0))
  smp = sampler isa MH ? 0 : sampler isa MALA ? 1 : sampler isa NUTS ? 3 : 2
  da = tuner isa DualAveragingMCTuner
  cfg = KlbConfig(sizeof(KlbConfig), smp, code(p.logtarget), tunercode(tuner),
                  arith == :fma ? 1 : 0, n, d, range.nsteps, range.burnin, range.thinning,
                  (sampler isa HMC || sampler isa NUTS) ? sampler.leapstep : sampler isa MALA ? sampler.driftstep : 1.0,
                  sampler isa HMC ? sampler.nleaps : 1,
                  (tuner isa AcceptanceRateMCTuner || da) ? tuner.targetrate : 0.5,
                  tuner isa AcceptanceRateMCTuner ? tuner.k : 7.0, tuner.period, tuner.verbose,
                  mon, ((:accept in diags) ? 1 : 0) | ((:ndoublings in diags) ? 2 : 0) | ((:a in diags) ? 4 : 0) | ((:na in diags) ? 8 : 0), dest == :none ? 1 : 0, seed, chain_offset, device,
                  tuner isa AcceptanceRateMCTuner ? tuner.score : 0,
                  da ? tuner.nadapt : 0, da ? tuner.t0 : 10, da ? tuner.ε0bar : 1.0, da ? tuner.h0bar : 0.0,
                  da ? tuner.γ : 0.05, da ? tuner.κ : 0.75,
                  sampler isa NUTS ? sampler.maxδ : 0, sampler isa NUTS ? sampler.maxndoublings : 0)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  multi = ngpus != 1                   # ngpus = 0: every visible device (klb_multi_create); chains in contiguous blocks
  if multi
    check(ccall((:klb_multi_create, LIB), Cint, (Ref{KlbConfig}, Int32, Ptr{Int32}, Ref{Ptr{Cvoid}}), cfg, ngpus, C_NULL, h))
  else
    check(ccall((:klb_job_create, LIB), Cint, (Ref{KlbConfig}, Ref{Ptr{Cvoid}}), cfg, h))
  end
  job = BasicMCJob(h[], n, d, range, monitor, diags, multi)
  finalizer(j -> j.multi ? ccall((:klb_multi_destroy, LIB), Cvoid, (Ptr{Cvoid},), j.handle) :
                           ccall((:klb_job_destroy, LIB), Cvoid, (Ptr{Cvoid},), j.handle), job)
  settarget(which, a) = check(ccall((sym(job, "set_target_f64"), LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64), job.handle, which, a, length(a)))
  t = p.logtarget
  t isa ShiftedIsoGaussian && settarget(0, t.mu)
  t isa Rosenbrock && settarget(3, [t.a, t.b, t.scale])
  t isa DenseGaussian && settarget(1, t.C)
  if t isa BayesLogit      # X travels row-major (row i = observation i): Julia's column-major X' is exactly that
    settarget(6, [t.lambda]); settarget(4, Matrix{Float64}(t.X')); settarget(5, t.y)
  end
  sampler isa MH && settarget(2, sampler.sigma)
  if synthetic                                                                                   # initialize!
    check(ccall((sym(job, "set_state_synthetic"), LIB), Cint, (Ptr{Cvoid},), job.handle))
  else
    GC.@preserve x0 check(ccall((sym(job, "set_state"), LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), job.handle, x0))
  end
  job
end

run(job::BasicMCJob) = (check(ccall((sym(job, "run"), LIB), Cint, (Ptr{Cvoid},), job.handle)); job)
# position the RNG streams: the next transition is number t + 1 (klb_job_seek / klb_multi_seek)
seek!(job::BasicMCJob, t::Integer) = check(ccall((sym(job, "seek"), LIB), Cint, (Ptr{Cvoid}, UInt64), job.handle, t))
# device g's copy of the closing all-gather of a sharded job: final states (d x nchains) of ALL chains
function gathered(job::BasicMCJob, g::Integer=0)
  a = Array{Float64}(undef, job.dim, job.nchains)
  GC.@preserve a check(ccall((:klb_multi_gathered_output, LIB), Cint, (Ptr{Cvoid}, Int32, Cint, Ptr{Cvoid}, Int64), job.handle, g, 4, a, sizeof(a)))
  a
end
run(jobs::Vector{BasicMCJob}) = map(run, jobs)
# reset(job, x0); run(job); output fields, in ONE pipelined call (klb_job_run_host): chain slices on their own streams,
# host->device copies, kernels and device->host copies overlap.  `outputs` maps KLB_OUT_* codes to preallocated Arrays.
function run_host(job::BasicMCJob, x0::Union{Matrix{Float64},Nothing}, outputs::Dict{Int,<:Array}; nslices::Integer=0)
  f = [KlbHostField(Int32(k), 0, pointer(a), sizeof(a)) for (k, a) in outputs]
  GC.@preserve x0 outputs f check(ccall((sym(job, "run_host"), LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{KlbHostField}, Int32, Int32),
                                        job.handle, x0 === nothing ? C_NULL : x0, f, length(f), nslices))
  job
end
reset(job::BasicMCJob) = check(ccall((sym(job, "reset"), LIB), Cint, (Ptr{Cvoid},), job.handle))
function reset(job::BasicMCJob, x::Matrix{Float64})
  GC.@preserve x check(ccall((sym(job, "set_state"), LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), job.handle, x))
end

function fetch!(job::BasicMCJob, field::Integer, a::Array)
  GC.@preserve a check(ccall((sym(job, "output"), LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64), job.handle, field, a, sizeof(a)))
  a
end

# output(job): value (d, npost, nchains), logtarget (npost, nchains), accept (npost, nchains) -- the NState layout
function output(job::BasicMCJob)
  P = job.range.npoststeps
  (value = :value in job.monitor ? fetch!(job, 0, Array{Float64}(undef, job.dim, P, job.nchains)) : nothing,
   logtarget = :logtarget in job.monitor ? fetch!(job, 1, Array{Float64}(undef, P, job.nchains)) : nothing,
   gradlogtarget = :gradlogtarget in job.monitor ? fetch!(job, 2, Array{Float64}(undef, job.dim, P, job.nchains)) : nothing,
   accept = :accept in job.diagnostics ? fetch!(job, 3, Array{UInt8}(undef, P, job.nchains)) .!= 0 : nothing,
   ndoublings = :ndoublings in job.diagnostics ? Int.(fetch!(job, 12, Array{UInt8}(undef, P, job.nchains))) : nothing,   # NUTS
   a = :a in job.diagnostics ? fetch!(job, 13, Array{Float64}(undef, P, job.nchains)) : nothing,          # NUTS + DualAveragingMCTuner
   na = :na in job.diagnostics ? Int.(fetch!(job, 14, Array{Int32}(undef, P, job.nchains))) : nothing)
end

# ess(output(job)): effective sample size (IMSE) per coordinate and chain, computed on the device
# (src/stats/convergence/ess.jl:3-14)
function ess(job::BasicMCJob)
  job.multi && error("statistics of a sharded job: ask every shard (klb_multi_job + klb_job_ess)")
  e = Array{Float64}(undef, job.dim, job.nchains)
  GC.@preserve e check(ccall((:klb_job_ess, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), job.handle, e))
  e
end

# The other post-hoc estimators of the stored output, same device pass (KLB_STAT_* of include/klara_b200.h):
#   mean (src/stats/mean.jl:7-11), mcvar(:iid | :imse) (src/stats/variance/mcvar.jl:5,75-105),
#   iact (src/stats/convergence/iact.jl:3-5), acceptance (src/stats/acceptance.jl:3-14,28-34)
function stat(job::BasicMCJob, code::Integer, perchain::Bool=false)
  job.multi && error("statistics of a sharded job: ask every shard (klb_multi_job + klb_job_stat)")
  r = perchain ? Array{Float64}(undef, job.nchains) : Array{Float64}(undef, job.dim, job.nchains)
  GC.@preserve r check(ccall((:klb_job_stat, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), job.handle, code, r))
  r
end
mean(job::BasicMCJob) = stat(job, 0)
mcvar(job::BasicMCJob, vtype::Symbol=:imse) =
  vtype == :iid ? stat(job, 1) : vtype == :imse ? stat(job, 2) : error("mcvar on the device supports :iid and :imse")
mcse(job::BasicMCJob, vtype::Symbol=:imse) = sqrt.(mcvar(job, vtype))
iact(job::BasicMCJob) = stat(job, 4)
acceptance(job::BasicMCJob; diagnostics::Bool=true) = stat(job, diagnostics ? 5 : 6, true)

end # module

# StochyB200.jl -- Julia host shim over libstochy_b200.so (include/stochy_b200.h).
#
# UNVERIFIED IN THIS ENVIRONMENT: no `julia` binary is installed in the build image and the
# reference itself is pinned to Julia 0.3 (REQUIRE:1), which no current Julia parses.  This
# file is written for Julia >= 1.6 against the C ABI that the Python mirror
# (stochy.jl_b200/api.py) exercises in the test-suite; the struct layouts below are the same
# ones `_native.py` declares with ctypes.
#
# It keeps the reference's entry points for the accelerated path
#     pmcmc(store, k, address, comp, numiterations, numparticles)      src/pmcmc.jl:78-99
#     smc(store, k, address, comp, n) = pmcmc(store,k,address,comp,1,n) src/pmcmc.jl:101-103
#     mh(store, k, address, comp, numsteps=10)                         src/mh.jl:47-93
# with `comp` a FIXED-STRUCTURE model descriptor instead of a CPS closure.  A closure (general
# CPS recursion, dp.jl-style variable dimension) is rejected with error(...): there is no CPU
# fallback.  Every entry ends with `k(store, Discrete(counts))`, exactly like the reference,
# and the native handle is destroyed in a `finally` block (the reference restores `ctx` there).
module StochyB200

export DiscreteHMM, LinearGaussian, BayesNet, OpsProgram, OpStep, pmcmc, smc, mh, enum, smc_logevidence, hmm2_example, rand_discrete, score_discrete

const LIB = get(ENV, "STOCHY_B200_LIB", "libstochy_b200")

# ---- status codes (include/stochy_b200.h)
const SB_OK = Int32(0)
const SB_FP32, SB_FP64 = Int32(0), Int32(1)
const SB_RESAMPLE_SYSTEMATIC, SB_RESAMPLE_MULTINOMIAL, SB_RESAMPLE_LITERAL = Int32(0), Int32(1), Int32(2)

# ---- POD structs, field for field
mutable struct SbOptions
    device::Int32
    precision::Int32
    resample_mode::Int32
    retained_pick::Int32
    store_logw::Int32
    use_graph::Int32
    seed::UInt64
end
SbOptions() = SbOptions(0, SB_FP32, SB_RESAMPLE_SYSTEMATIC, -1, 0, 0, 1)

mutable struct SbStats
    kernel_launches::Int64
    device_ms::Float64
    step_kernel_ms::Float64
    particle_steps::Int64
    logZ_first::Float64
    logZ_last::Float64
    accepts::Int64
end
SbStats() = SbStats(0, 0.0, 0.0, 0, 0.0, 0.0, 0)

const Handle = Ptr{Cvoid}

lasterror(h::Handle) = unsafe_string(ccall((:sb_last_error, LIB), Cstring, (Handle,), h))

# rc != 0  ->  Julia ErrorException with the reference's own message texts
# ("unexpected nan in rand" src/rand.jl:15, "Distribution has zero probability mass." src/erp.jl:48)
check(h::Handle, rc::Int32) = rc == SB_OK ? nothing : error(lasterror(h))

function create(; device=0, precision=SB_FP64, resample=SB_RESAMPLE_LITERAL, seed=1, retained_pick=-1)
    o = SbOptions()
    ccall((:sb_options_default, LIB), Cvoid, (Ref{SbOptions},), o)
    o.device = device; o.precision = precision; o.resample_mode = resample
    o.seed = seed; o.retained_pick = retained_pick
    h = Ref{Handle}(C_NULL)
    rc = ccall((:sb_create, LIB), Int32, (Ref{SbOptions}, Ref{Handle}), o, h)
    rc == SB_OK || error("sb_create failed: " * lasterror(Handle(C_NULL)))
    h[]
end
destroy(h::Handle) = ccall((:sb_destroy, LIB), Cvoid, (Handle,), h)

# ---- model descriptors (the fixed per-step structure the GPU path accepts)
abstract type Model end

# examples/hmm2.jl form: x_1 ~ Categorical(A[x0,:]) (or pi), x_t ~ Categorical(A[x_{t-1},:]),
# factor(logF[t, x_t]) after every step with has_factor[t].  States are 1-based on the Julia side
# (randominteger is 1-based, src/erp.jl:28-30) and 0-based across the ABI.
struct DiscreteHMM <: Model
    A::Matrix{Float64}             # K x K, row i = P(. | i)
    logF::Matrix{Float64}          # T x K
    x0::Int                        # fixed initial state (1-based) or 0 for x_1 ~ pi
    pi::Vector{Float64}
    has_factor::Vector{UInt8}
    value::Function                # maps the sampled state tuple to the program's return value
end
DiscreteHMM(A, logF; x0=0, pi=fill(1/size(A,1), size(A,1)), has_factor=ones(UInt8, size(logF,1)), value=identity) =
    DiscreteHMM(A, logF, x0, pi, has_factor, value)

struct LinearGaussian <: Model
    a::Float64; q::Float64; c::Float64; r::Float64; m0::Float64; P0::Float64
    y::Vector{Float64}
end

struct BayesNet <: Model
    site_parents::Vector{UInt32}
    site_cpt::Matrix{Float64}      # D x stride, P(site = true | packed parent bits)
    fac_parents::Vector{UInt32}
    fac_logf::Matrix{Float64}      # F x stride
    query::Vector{Int}             # 1-based site indices returned by the program
end

# One real latent per particle; per step a draw statement and an observe statement (include/stochy_b200.h: sb_op_step;
# src/pmcmc.jl:32-46, src/Stochy.jl:55-56 with the continuous ERPs of src/erp.jl:1-30).  isbits, C layout.
struct OpStep
    draw::Int32; score::Int32          # SB_DRAW_* / SB_SCORE_* codes
    dp::NTuple{4,Float64}; sp::NTuple{4,Float64}
    obs_off::Int32; obs_n::Int32
    par_off::Int32; par_n::Int32       # Dirichlet's alpha: obs[par_off+1 : par_off+par_n]
end
struct OpsProgram <: Model
    steps::Vector{OpStep}
    obs::Vector{Float64}
end
# e.g. p = ~Beta(2, 3); observe(Bernoulli(p), ys...):
#   OpsProgram([OpStep(2, 2, (2., 3., 0., 0.), (1., 0., 0., 0.), 0, length(ys), 0, 0)], Float64.(ys))
# smc_logevidence(prog, n) then returns the log-evidence estimate of one sweep.

# Row-major tables cross the ABI; Julia matrices are column-major.
rowmajor(M::Matrix{Float64}) = collect(vec(permutedims(M)))

function load!(h::Handle, m::DiscreteHMM)
    K, T = size(m.A, 1), size(m.logF, 1)
    check(h, ccall((:sb_model_discrete_hmm, LIB), Int32,
                   (Handle, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}),
                   h, K, T, m.x0 - 1, m.pi, rowmajor(m.A), rowmajor(m.logF), m.has_factor))
end
function load!(h::Handle, m::LinearGaussian)
    check(h, ccall((:sb_model_linear_gaussian, LIB), Int32,
                   (Handle, Int32, Float64, Float64, Float64, Float64, Float64, Float64, Ptr{Float64}),
                   h, length(m.y), m.a, m.q, m.c, m.r, m.m0, m.P0, m.y))
end
function load!(h::Handle, m::BayesNet)
    D, F = length(m.site_parents), length(m.fac_parents)
    stride = size(m.site_cpt, 2)
    qmask = UInt32(sum(1 << (q - 1) for q in m.query))
    check(h, ccall((:sb_model_bayesnet, LIB), Int32,
                   (Handle, Int32, Int32, Ptr{UInt32}, Ptr{Float64}, Ptr{UInt32}, Ptr{Float64}, Int32, UInt32),
                   h, D, F, m.site_parents, rowmajor(m.site_cpt), m.fac_parents, rowmajor(m.fac_logf), stride, qmask))
end
function load!(h::Handle, m::OpsProgram)
    check(h, ccall((:sb_model_ops, LIB), Int32, (Handle, Int32, Ptr{OpStep}, Ptr{Float64}, Int64),
                   h, length(m.steps), m.steps, m.obs, length(m.obs)))
end
load!(::Handle, comp) = error("model is not a fixed-structure descriptor: general CPS programs are not " *
                              "supported on the GPU path and there is no CPU fallback")

# ---- result type: the reference's own Discrete(counts) (src/erp.jl:39-56) is used when Stochy is
# loaded; otherwise a Dict of counts is passed to k.
discrete(counts::Dict) = isdefined(Main, :Stochy) ? Main.Stochy.Discrete(counts) : counts

function decode_joint(m::DiscreteHMM, joint::Vector{Int64})
    K, T = size(m.A, 1), size(m.logF, 1)
    counts = Dict{Any,Int64}()
    for code in findall(!iszero, joint)
        c = code - 1
        xs = ntuple(t -> (c ÷ K^(t - 1)) % K + 1, T)        # little-endian base-K digits, back to 1-based
        v = m.value(xs)
        counts[v] = get(counts, v, 0) + joint[code]
    end
    counts
end

# ---- pmcmc(store, k, address, comp, numiterations, numparticles)  -- ITERATIONS FIRST (src/pmcmc.jl:78)
# chains > 1: that many independent chains side by side on the GPU (numparticles <= 2047), counts pooled
function pmcmc(store, k::Function, address, comp, numiterations, numparticles; chains=1, kw...)
    comp isa DiscreteHMM || error("pmcmc: only the fixed-structure discrete model has a finite value encoding " *
                                  "on the GPU path (no CPU fallback)")
    h = create(; kw...)
    counts = try
        load!(h, comp)
        K, T = size(comp.A, 1), size(comp.logF, 1)
        joint = zeros(Int64, K^T)
        if chains > 1
            check(h, ccall((:sb_pmcmc_chains, LIB), Int32, (Handle, Int64, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}),
                           h, chains, numiterations, numparticles, 0, C_NULL, joint))
        else
            check(h, ccall((:sb_pmcmc_run, LIB), Int32, (Handle, Int64, Int64, Ptr{Int64}, Ptr{Int64}),
                           h, numiterations, numparticles, C_NULL, joint))
        end
        decode_joint(comp, joint)
    finally
        destroy(h)                       # `finally ctx = ctxold` (src/pmcmc.jl:95-97)
    end
    k(store, discrete(counts))
end

smc(store, k::Function, a, comp, n; kw...) = pmcmc(store, k, a, comp, 1, n; kw...)

# log-evidence of one unconditional sweep (north-star extension; no reference line)
function smc_logevidence(comp::Model, n; iter=0, precision=SB_FP32, resample=SB_RESAMPLE_SYSTEMATIC, seed=1)
    h = create(; precision=precision, resample=resample, seed=seed)
    try
        load!(h, comp)
        lz = Ref{Float64}(0.0)
        check(h, ccall((:sb_smc_sweep, LIB), Int32, (Handle, Int64, UInt32, Int32, Ref{Float64}), h, n, iter, 0, lz))
        lz[]
    finally
        destroy(h)
    end
end

# ---- enum(store, k, address, comp)   (src/enumerate.jl:57-85): exhaustive, exact.  returns[value] += exp(score)
function enum(store, k::Function, address, comp; kw...)
    comp isa DiscreteHMM || error("enum: only the fixed-structure discrete model is enumerated on the GPU path")
    h = create(; kw..., precision=SB_FP64)
    returns = try
        load!(h, comp)
        K, T = size(comp.A, 1), size(comp.logF, 1)
        joint = zeros(Float64, K^T)
        lz = Ref{Float64}(0.0)
        check(h, ccall((:sb_enum_hmm, LIB), Int32, (Handle, Ptr{Float64}, Ptr{Float64}, Ref{Float64}), h, joint, C_NULL, lz))
        out = Dict{Any,Float64}()
        for code in findall(!iszero, joint)
            c = code - 1
            v = comp.value(ntuple(t -> (c ÷ K^(t - 1)) % K + 1, T))
            out[v] = get(out, v, 0.0) + joint[code]
        end
        out
    finally
        destroy(h)
    end
    k(store, discrete(returns))
end

# ---- mh(store, k, address, comp, numsteps=10)   (src/mh.jl:47-93), batched over `chains`
function mh(store, k::Function, address, comp, numsteps=10; chains=1, kw...)
    comp isa BayesNet || error("mh: only fixed-structure binary Bayes nets run on the GPU path (no CPU fallback)")
    h = create(; kw...)
    counts = try
        load!(h, comp)
        nq = length(comp.query)
        raw = zeros(Int64, 1 << nq)
        check(h, ccall((:sb_mh_run, LIB), Int32, (Handle, Int64, Int64, Int64, Ptr{Int64}), h, chains, numsteps, 0, raw))
        out = Dict{Any,Int64}()
        for v in findall(!iszero, raw)
            bits = ntuple(i -> ((v - 1) >> (i - 1)) & 1 == 1, nq)
            out[nq == 1 ? bits[1] : bits] = raw[v]
        end
        out
    finally
        destroy(h)
    end
    k(store, discrete(counts))
end

# Discrete(x, p) as a model primitive (src/erp.jl:39-59) drawn / scored on the device: `dict` is the possibly
# un-normalised Dict the reference's constructor takes; draws come back as elements of keys(dict).
function rand_discrete(dict::Dict, n::Integer; seed=1, kw...)
    xs = collect(keys(dict)); ps = Vector{Float64}(collect(values(dict)))
    h = create(; kw...)
    try
        idx = Vector{Int32}(undef, n)
        check(h, ccall((:sb_erp_sample_discrete, LIB), Int32, (Handle, Ptr{Float64}, Int32, Int64, UInt64, Ptr{Int32}),
                       h, ps, length(ps), n, seed, idx))
        [xs[i] for i in idx]                       # rand(erp::Discrete) = erp.x[rand(erp.p)]  (src/erp.jl:52)
    finally
        destroy(h)
    end
end
function score_discrete(dict::Dict, values::Vector; kw...)
    xs = collect(keys(dict)); ps = Vector{Float64}(collect(values(dict)))
    pos = Dict(x => Int32(i) for (i, x) in enumerate(xs))
    idx = Int32[get(pos, v, Int32(0)) for v in values]          # 0 = outside the support: -Inf
    h = create(; kw...)
    try
        out = Vector{Float64}(undef, length(idx))
        check(h, ccall((:sb_erp_logpdf_discrete, LIB), Int32, (Handle, Ptr{Float64}, Int32, Ptr{Int32}, Int64, Ptr{Float64}),
                       h, ps, length(ps), idx, length(idx), out))
        out                                        # score(erp::Discrete, x) = log(erp.dict[x]/erp.z)  (src/erp.jl:56)
    finally
        destroy(h)
    end
end

# examples/hmm2.jl:5-24 as a descriptor: init false; state ~ flip(prev ? .9 : .1);
# factor(state == obs ? 0 : -1); returns the Bool list including the initial state.
function hmm2_example(observations=(true, true, true, true))
    A = [0.9 0.1; 0.1 0.9]
    logF = [((k == 2) == o) ? 0.0 : -1.0 for o in observations, k in 1:2]
    DiscreteHMM(A, Matrix{Float64}(logF); x0=1, value = xs -> (false, map(x -> x == 2, xs)...))
end

end # module

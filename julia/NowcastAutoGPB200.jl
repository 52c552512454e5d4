# NowcastAutoGPB200.jl — thin ccall shim over libnagp.so (include/nagp.h) for NowcastAutoGP.jl.
#
# NEVER EXECUTED IN THE BUILD IMAGE (no Julia there): kept deliberately thin — flatten the particles,
# one ccall per device, reshape. The same C ABI is tested end to end from Python (nowcastautogp_b200/engine.py)
# and from plain C (tests/c_abi/test_c_abi.c). The methods carry the reference's own names and signatures
# (`forecast_with_nowcasts`, `make_and_fit_model`): `using NowcastAutoGPB200` in place of `using NowcastAutoGP`.
# See INTEGRATION.md for where a maintainer hooks this into src/forecasting.jl.
module NowcastAutoGPB200

using AutoGP, Dates, LinearAlgebra, Random

import NowcastAutoGP
import NowcastAutoGP: TData

const libnagp = get(ENV, "NAGP_LIB", joinpath(@__DIR__, "..", "nowcastautogp_b200", "libnagp.so"))

available() = isfile(libnagp)

# ---- devices and contexts ---------------------------------------------------------------------------------------
# A nagp_ctx is bound to one CUDA device and is not re-entrant. The reference calls in from one task per scenario
# (`Threads.@spawn`, /root/reference/src/forecasting.jl:131-132) and tasks migrate between threads, so contexts are
# not keyed by `Threads.threadid()`: every device owns a lock-guarded pool, a call checks a context out for its
# duration (`with_ctx`) and returns it; the pool grows on demand.
const _devices = Ref{Vector{Int}}(haskey(ENV, "NAGP_DEVICES") ? parse.(Int, split(ENV["NAGP_DEVICES"], ",")) : [0])
const _pools = Dict{Int, Vector{Ptr{Cvoid}}}()
const _pool_lock = ReentrantLock()

"Devices the shim spreads (series, scenario) pairs over — e.g. `set_devices!(0:7)` on an 8 x B200 box."
set_devices!(devs) = (_devices[] = collect(Int, devs); nothing)
devices() = get(task_local_storage(), :nagp_devices, _devices[])::Vector{Int}   # a series dealt to one device overrides the list for its task

function _new_ctx(dev::Int)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:nagp_init, libnagp), Int32, (Int32, Ptr{Ptr{Cvoid}}), Int32(dev), out)
    rc == 0 || error("nagp_init(device $dev): " * unsafe_string(ccall((:nagp_last_error, libnagp), Cstring, (Ptr{Cvoid},), C_NULL)))
    return out[]
end

"Run `f(ctx)` with a context of device `dev` checked out of that device's pool."
function with_ctx(f, dev::Int = first(devices()))
    c = lock(_pool_lock) do
        pool = get!(() -> Ptr{Cvoid}[], _pools, dev)
        isempty(pool) ? C_NULL : pop!(pool)
    end
    c == C_NULL && (c = _new_ctx(dev))
    try
        return f(c)
    finally
        lock(_pool_lock) do
            push!(_pools[dev], c)
        end
    end
end

function check(c::Ptr{Cvoid}, rc::Int32)
    rc > 0 && throw(LinearAlgebra.PosDefException(rc))   # /root/reference/test/test_model_fitting.jl:97-98
    rc < 0 && error("libnagp: " * unsafe_string(ccall((:nagp_last_error, libnagp), Cstring, (Ptr{Cvoid},), c)))
    return nothing
end

"Contiguous split of `K` scenarios over `G` devices (series-first when the caller loops over series: one series per call)."
split_scenarios(K::Int, G::Int) = [((K * (r - 1)) ÷ G + 1):((K * r) ÷ G) for r in 1:G if (K * r) ÷ G > (K * (r - 1)) ÷ G]

# ---- wire format: AutoGP kernel tree -> post-order byte program + theta (docs/KERNEL_SPEC.md §1) ----
const GP = AutoGP.GP
emit!(prog, th, k::GP.Constant) = (push!(prog, 0x01); push!(th, k.value))
emit!(prog, th, k::GP.Linear) = (push!(prog, 0x02); push!(th, k.intercept, k.bias, k.amplitude))
emit!(prog, th, k::GP.SquaredExponential) = (push!(prog, 0x03); push!(th, k.lengthscale, k.amplitude))
emit!(prog, th, k::GP.GammaExponential) = (push!(prog, 0x04); push!(th, k.lengthscale, k.gamma, k.amplitude))
emit!(prog, th, k::GP.Periodic) = (push!(prog, 0x05); push!(th, k.lengthscale, k.period, k.amplitude))
emit!(prog, th, k::GP.Plus) = (emit!(prog, th, k.left); emit!(prog, th, k.right); push!(prog, 0x06))
emit!(prog, th, k::GP.Times) = (emit!(prog, th, k.left); emit!(prog, th, k.right); push!(prog, 0x07))
function emit!(prog, th, k::GP.ChangePoint)
    emit!(prog, th, k.left); emit!(prog, th, k.right); push!(prog, 0x08); push!(th, k.location, k.scale)
end

struct Flat
    P::Int
    prog::Vector{UInt8}; prog_off::Vector{Int64}
    theta::Vector{Float64}; theta_off::Vector{Int64}
    noise::Vector{Float64}; logw::Vector{Float64}
    y1::Vector{Float64}; ya::Float64; yb::Float64
    t_slope::Float64; t_intercept::Float64; t_train::Vector{Float64}
end

"Flatten a fitted model (the state Dict(model) serialises, /root/reference/src/forecasting.jl:128)."
function flatten_model(model::AutoGP.GPModel)
    kernels = AutoGP.covariance_kernels(model)
    noise = AutoGP.observation_noise_variances(model)
    logw = AutoGP.log_weights(model)    # unnormalised particle log-weights
    prog = UInt8[]; theta = Float64[]; po = Int64[0]; to = Int64[0]
    for k in kernels
        emit!(prog, theta, k); push!(po, length(prog)); push!(to, length(theta))
    end
    yt, dt = model.y_transform, model.ds_transform     # LinearTransform(slope, intercept)
    t_train = dt.slope .* AutoGP.Transforms.to_numeric.(model.ds) .+ dt.intercept
    y1 = yt.slope .* model.y .+ yt.intercept
    return Flat(length(kernels), prog, po, theta, to, noise, logw, y1, yt.slope, yt.intercept,
                dt.slope, dt.intercept, t_train)
end

"t[q] over [train | new dates] plus the lag-grid arguments (g, step) when the dates sit on a regular grid."
function time_arguments(fl::Flat, new_dates)
    num = vcat((fl.t_train .- fl.t_intercept) ./ fl.t_slope, AutoGP.Transforms.to_numeric.(new_dates))
    t = fl.t_slope .* num .+ fl.t_intercept
    d = round.(Int64, num .- num[1]); gcdv = reduce(gcd, d[2:end]; init = 0)
    if gcdv > 0 && all(num .- num[1] .== d)
        return t, Int32.(d .÷ gcdv), gcdv * fl.t_slope
    end
    return t, nothing, 0.0
end

# ---- z -> theta maps of the hyperparameter slots ------------------------------------------------------------------------
"""
    slot_spec(model) -> (kind, a, b, noise, z0, noise_z0, to_theta, to_noise)

What `nagp_hmc` needs to move a model's hyperparameters in unconstrained space (include/nagp.h): per theta slot the map
kind (0: exp(a + b z), 2: 2 logistic(a + b z), 3: z, 4: Phi(z), 5: constant a) with its (a, b) — AutoGP's LogNormal
(wildcard, period), scaled-logistic (gamma), identity (Linear intercept), uniform (ChangePoint location) priors and the
fixed ChangePoint scale, docs/KERNEL_SPEC.md §1 — the same triple for the noise, the current z of every slot
(`z0 [total]`, `noise_z0 [P]`: invert the maps on `flatten_model(model).theta / .noise`), and the two closures
that map `[total, K]` / `[P, K]` matrices of z back to constrained values. AutoGP keeps these transforms inside its
Gen model; the maintainer fills this in from `model.config.prior` (INTEGRATION.md §3).
"""
function slot_spec(model::AutoGP.GPModel)
    error("NowcastAutoGPB200.slot_spec: provide the z -> theta maps of AutoGP's hyperparameter slots (see the docstring) " *
          "to use n_hmc > 0 or forecast_n_hmc on the device; n_hmc == 0 needs none of this")
end

# ---- the reference's entry points, same names and signatures ---------------------------------------------------------

"One device's share of the default path (n_mcmc == n_hmc == 0, no forecast_n_hmc): ONE fused call."
function _fused_block(dev::Int, fl::Flat, t, g, step, y2::Matrix{Float64}, h::Int, D::Int, ess_threshold::Float64)
    k, K = size(y2); P = fl.P; n = length(fl.y1)
    comp = fill(Int32(-1), D, K); u = rand(D, K)
    u_res = ess_threshold > 0 ? rand(P, K) : nothing
    zeta = randn(h, D, K)
    x = Matrix{Float64}(undef, h, K * D); info = zeros(Int32, P)
    with_ctx(dev) do c
        GC.@preserve fl t g y2 comp u u_res zeta x info begin
            rc = ccall((:nagp_forecast_with_nowcasts, libnagp), Int32,
                (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{UInt8}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Float64,
                 Int64, Int64, Int64, Ptr{Float64}, Ptr{Int32}, Float64, Ptr{Float64}, Ptr{Float64}, Float64, Float64,
                 Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                c, K, P, D, fl.prog, fl.prog_off, fl.theta, fl.theta_off, fl.noise, -1.0,
                n, k, h, t, g === nothing ? C_NULL : g, step, fl.y1, y2, fl.ya, fl.yb,
                fl.logw, comp, u, u_res === nothing ? C_NULL : u_res, ess_threshold, zeta,
                x, C_NULL, C_NULL, info)
            check(c, rc)
        end
    end
    return x
end

"""
One device's share of the per-scenario-hyperparameter path (`n_mcmc == 0 && n_hmc > 0`, and/or `forecast_n_hmc`): all
K x P chains of the block advance together in `nagp_hmc` (`mcmc_parameters!` on every scenario's model copy at once,
/root/reference/src/forecasting.jl:148), then `nagp_forecast_instances` + `nagp_draw`. With `forecast_n_hmc` the
reference's per-draw loop (`:63-68`) becomes D rounds of (hmc, moments, one draw per scenario). `spec` carries the
z -> theta maps of the model's hyperparameter slots (`slot_spec(base_model)`).
"""
function _chain_block(dev::Int, fl::Flat, spec, t, g, step, y2::Matrix{Float64}, h::Int, D::Int, n_hmc::Int,
        forecast_n_hmc::Union{Int, Nothing})
    k, K = size(y2); P = fl.P; n = length(fl.y1); total = length(fl.theta); m = n + k
    z = repeat(spec.z0, 1, K); noise_z = repeat(spec.noise_z0, 1, K)
    tm = t[1:m]; gm = g === nothing ? nothing : g[1:m]
    n_hmc > 0 && hmc!(fl, spec.kind, spec.a, spec.b, spec.noise, z, noise_z, tm, gm, step, y2; n_steps = n_hmc, dev = dev)
    x = Matrix{Float64}(undef, h, K * D)
    draw_rounds = forecast_n_hmc === nothing ? [(1:D, D)] : [((i:i), 1) for i in 1:D]
    for (cols, Dn) in draw_rounds
        forecast_n_hmc === nothing || hmc!(fl, spec.kind, spec.a, spec.b, spec.noise, z, noise_z, tm, gm, step, y2;
                                           n_steps = forecast_n_hmc, dev = dev)
        theta_k = spec.to_theta(z); noise_k = spec.to_noise(noise_z)
        logw = Matrix{Float64}(undef, P, K); mu = Array{Float64}(undef, h, P, K); L = Array{Float64}(undef, h, h, P, K)
        info = Matrix{Int32}(undef, P, K)
        zeta = randn(h, Dn, K); u = rand(Dn, K); xb = Matrix{Float64}(undef, h, K * Dn)
        with_ctx(dev) do c
            GC.@preserve fl theta_k noise_k t g y2 logw mu L info zeta u xb begin
                rc = ccall((:nagp_forecast_instances, libnagp), Int32,
                    (Ptr{Cvoid}, Int64, Int64, Ptr{UInt8}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Int64, Ptr{Float64}, Int64,
                     Float64, Int64, Int64, Int64, Ptr{Float64}, Ptr{Int32}, Float64, Ptr{Float64}, Ptr{Float64},
                     Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
                    c, K, P, fl.prog, fl.prog_off, theta_k, fl.theta_off, total, noise_k, P, -1.0, n, k, h, t,
                    g === nothing ? C_NULL : g, step, fl.y1, y2, fl.ya, fl.yb, fl.logw, logw, mu, L, info, C_NULL, C_NULL)
                check(c, rc)
                # row-major (K, P, h, h) on the C side = column-major (h, h, P, K) with each factor transposed: nagp_draw
                # consumes exactly what nagp_forecast_instances wrote
                rc = ccall((:nagp_draw, libnagp), Int32,
                    (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64,
                     Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                    c, K, P, h, Dn, logw, mu, P * h, L, P * h * h, C_NULL, u, C_NULL, 0.0, zeta, xb, C_NULL, C_NULL)
                check(c, rc)
            end
        end
        if forecast_n_hmc === nothing
            x .= xb
        else
            for s in 1:K; x[:, (s - 1) * D + first(cols)] .= xb[:, s]; end
        end
    end
    return x
end

"""
    forecast_with_nowcasts(base_model, nowcasts, forecast_dates, forecast_draws_per_nowcast;
                           inv_transformation = y -> y, n_mcmc = 0, n_hmc = 0, ess_threshold = 0.0,
                           forecast_n_hmc = nothing, verbose = false)

Same name, positional and keyword arguments, assertions and return value as
`/root/reference/src/forecasting.jl:117-167`: a `Matrix{Float64}` of size
`(length(forecast_dates), length(nowcasts) * forecast_draws_per_nowcast)` in scenario-major column blocks; the base
model is not mutated. Where the reference spawns one task per scenario, this method issues one batched call per
device over that device's share of the scenarios (`devices()`, `set_devices!`). Kernel-structure moves stay in AutoGP:
with `n_mcmc > 0`, or scenarios that do not share their dates, the reference implementation itself runs.
`slot_spec(base_model)` must be provided by the maintainer for the `n_hmc > 0` / `forecast_n_hmc` paths (AutoGP's
parameter transforms are not part of its public API).
"""
function forecast_with_nowcasts(
        base_model::AutoGP.GPModel, nowcasts::AbstractVector{<:TData},
        forecast_dates, forecast_draws_per_nowcast::Int;
        inv_transformation = y -> y, n_mcmc = 0, n_hmc = 0, ess_threshold = 0.0,
        forecast_n_hmc::Union{Int, Nothing} = nothing, verbose::Bool = false
    )
    @assert !isempty(nowcasts) "nowcasts vector must not be empty"
    @assert !(n_mcmc > 0 && n_hmc == 0) "If n_mcmc > 0, n_hmc must also be > 0 for MCMC refinement"
    @assert 0.0 <= ess_threshold <= 1.0 "ess_threshold must be between 0 and 1"
    @assert forecast_n_hmc === nothing||forecast_n_hmc > 0 "forecast_n_hmc must be > 0 if specified"
    shared_ds = all(nc -> nc.ds == nowcasts[1].ds, nowcasts)
    if n_mcmc > 0 || !shared_ds || (ess_threshold > 0 && (n_hmc > 0 || forecast_n_hmc !== nothing))
        # structure moves (and whole-particle resampling before per-scenario rejuvenation) need AutoGP's own state
        return NowcastAutoGP.forecast_with_nowcasts(base_model, nowcasts, forecast_dates, forecast_draws_per_nowcast;
            inv_transformation, n_mcmc, n_hmc, ess_threshold, forecast_n_hmc, verbose)
    end
    D = forecast_draws_per_nowcast
    fl = flatten_model(base_model)                       # Dict(base_model) once, forecasting.jl:128
    dates = collect(forecast_dates)
    K, k, h = length(nowcasts), length(nowcasts[1].ds), length(dates)
    t, g, step = time_arguments(fl, vcat(nowcasts[1].ds, dates))
    y2 = Matrix{Float64}(undef, k, K)
    for (s, nc) in enumerate(nowcasts); y2[:, s] .= fl.ya .* nc.y .+ fl.yb; end
    devs = devices()
    shares = split_scenarios(K, length(devs))
    spec = (n_hmc > 0 || forecast_n_hmc !== nothing) ? slot_spec(base_model) : nothing
    tasks = map(enumerate(shares)) do (r, ks)
        Threads.@spawn begin
            if spec === nothing
                _fused_block(devs[r], fl, t, g, step, y2[:, ks], h, D, Float64(ess_threshold))
            else
                _chain_block(devs[r], fl, spec, t, g, step, y2[:, ks], h, D, Int(n_hmc), forecast_n_hmc)
            end
        end
    end
    x = hcat(map(fetch, tasks)...)                       # forecasting.jl:166
    verbose && @info "forecast_with_nowcasts: $(K) scenarios x $(fl.P) particles on $(length(shares)) device(s)"
    return inv_transformation.(x)
end

"""
`forecast_with_nowcasts` for several series at once (the per-jurisdiction loop of
`/root/reference/docs/vignettes/getting-started.jl:536-556`): whole series are dealt to the devices first (53 series
on 8 GPUs -> 7/7/7/7/7/6/6/6), so each particle is factored on exactly one device; returns one matrix per series.
"""
function forecast_with_nowcasts(base_models::AbstractVector{<:AutoGP.GPModel}, nowcasts::AbstractVector, forecast_dates,
        forecast_draws_per_nowcast::Int; kwargs...)
    devs = devices(); G = length(devs); S = length(base_models)
    owner(s) = S >= G ? findfirst(r -> s <= (S ÷ G) * r + min(r, S % G), 1:G) : 1
    all_devs = copy(devs)
    tasks = map(1:S) do s
        Threads.@spawn begin
            S >= G ? task_local_storage(:nagp_devices, [all_devs[owner(s)]]) : nothing
            forecast_with_nowcasts(base_models[s], nowcasts[s], forecast_dates, forecast_draws_per_nowcast; kwargs...)
        end
    end
    return map(fetch, tasks)
end

"""
    make_and_fit_model(data::TData; n_particles = 1, smc_data_proportion = 0.1, flat_threshold = 1.0e-3,
                       config = AutoGP.GP.GPConfig(), kwargs...)

Same signature and return type (`AutoGP.GPModel`) as `/root/reference/src/make_and_fit_model.jl:78-93`. Structure and
parameter proposals stay in AutoGP (BASELINE north_star), whose likelihood lives inside its Gen model: routing those
evaluations to `logml_batch` / `logml_grad` / `hmc!` below needs a hook inside AutoGP (INTEGRATION.md §3), not
something a wrapper can do from outside. Until that hook exists this method runs the reference's own steps — jitter
guard, `GPModel`, `linear_schedule`, `fit_smc!` — so callers can switch packages without touching their code, and it
audits the fitted particles' likelihood on the device (`logml_batch`) so a mismatch between AutoGP and the kernels
shows up at fit time instead of in the forecasts.
"""
function make_and_fit_model(
        data::TData; n_particles = 1, smc_data_proportion = 0.1,
        flat_threshold = 1.0e-3, config = AutoGP.GP.GPConfig(), kwargs...
    )
    model = NowcastAutoGP.make_and_fit_model(data; n_particles, smc_data_proportion, flat_threshold, config, kwargs...)
    if available()
        fl = flatten_model(model)
        t, g, step = time_arguments(fl, eltype(data.ds)[])
        lm = logml_batch(fl, t, g, step, fl.y1)
        all(isfinite, lm) || @warn "libnagp: a fitted particle's Gram is not positive definite on the device" lm
    end
    return model
end

"Batched log marginal likelihood of the model's particles over (t, y) — the fit_smc! primitive."
function logml_batch(fl::Flat, t::Vector{Float64}, g, step::Float64, y::Vector{Float64}; dev::Int = first(devices()))
    out = Vector{Float64}(undef, fl.P); info = Vector{Int32}(undef, fl.P)
    with_ctx(dev) do c
    GC.@preserve fl t g y out info begin
        rc = ccall((:nagp_logml_batch, libnagp), Int32,
            (Ptr{Cvoid}, Int64, Ptr{UInt8}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Int64,
             Ptr{Float64}, Ptr{Int32}, Float64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int32}),
            c, fl.P, fl.prog, fl.prog_off, fl.theta, fl.theta_off, fl.noise, length(t),
            t, g === nothing ? C_NULL : g, step, y, 0, out, info)
        rc < 0 && check(c, rc)   # > 0: a particle whose Gram is not PD has logml = NaN
    end
    end
    return out
end

"""
logML and its gradient w.r.t. every constrained hyperparameter and the noise, for K scenarios x P particles
(what mcmc_parameters! differentiates, /root/reference/src/forecasting.jl:148,65). `theta_k` is `[total, K]`
(K copies of the CSR theta vector), `noise_k` `[P, K]`, `y2` `[k, K]` scaled nowcast values (k may be 0).
Returns `(logml [P,K], grad_theta [total,K], grad_noise [P,K])`.
"""
function logml_grad(fl::Flat, theta_k::Matrix{Float64}, noise_k::Matrix{Float64}, t::Vector{Float64}, g, step::Float64,
        y2::Matrix{Float64}; dev::Int = first(devices()))
    K, P, total, k, n = size(theta_k, 2), fl.P, length(fl.theta), size(y2, 1), length(fl.y1)
    logml = Matrix{Float64}(undef, P, K); gth = Matrix{Float64}(undef, total, K); gnz = Matrix{Float64}(undef, P, K)
    info = Matrix{Int32}(undef, P, K)
    with_ctx(dev) do c
    GC.@preserve fl theta_k noise_k t g y2 logml gth gnz info begin
        rc = ccall((:nagp_logml_grad, libnagp), Int32,
            (Ptr{Cvoid}, Int64, Int64, Ptr{UInt8}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Int64, Ptr{Float64}, Int64,
             Int64, Int64, Ptr{Float64}, Ptr{Int32}, Float64, Ptr{Float64}, Int64, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
            c, K, P, fl.prog, fl.prog_off, theta_k, fl.theta_off, total, noise_k, P, n, k, t,
            g === nothing ? C_NULL : g, step, fl.y1, 0, k == 0 ? C_NULL : y2, logml, gth, gnz, info)
        rc < 0 && check(c, rc)   # > 0 only flags chains whose Gram is not PD: their logml is NaN, the sampler rejects
    end
    end
    return logml, gth, gnz
end

"""
`n_steps` HMC iterations (`n_leapfrog` stages of size `eps`) on the unconstrained hyperparameters of K x P chains,
integrator on the device (nagp_hmc): the drop-in for `AutoGP.mcmc_parameters!(model, n_hmc)` on every scenario's
copy at once (/root/reference/src/forecasting.jl:148,65). `z` `[total, K]` and `noise_z` `[P, K]` are updated in
place; `slot_kind/slot_a/slot_b` describe z -> theta per slot (include/nagp.h) — to be filled from AutoGP's
parameter transforms by the maintainer. Momenta and accept thresholds come from Julia's RNG, so a run is
reproducible under `Random.seed!` exactly like the reference.
"""
function hmc!(fl::Flat, slot_kind::Vector{Int32}, slot_a::Vector{Float64}, slot_b::Vector{Float64},
        noise_spec::Tuple{Int32, Float64, Float64}, z::Matrix{Float64}, noise_z::Matrix{Float64},
        t::Vector{Float64}, g, step::Float64, y2::Matrix{Float64}; n_steps::Int, n_leapfrog::Int = 10, eps::Float64 = 0.02,
        rng = Random.default_rng(), dev::Int = first(devices()))
    K, P, total, k, n = size(z, 2), fl.P, length(fl.theta), size(y2, 1), length(fl.y1)
    mom = randn(rng, total, K, n_steps); mnz = randn(rng, P, K, n_steps); logu = log.(rand(rng, P, K, n_steps))
    logml = Matrix{Float64}(undef, P, K); nacc = Matrix{Int32}(undef, P, K); info = Matrix{Int32}(undef, P, K)
    with_ctx(dev) do c
    GC.@preserve fl slot_kind slot_a slot_b z noise_z t g y2 mom mnz logu logml nacc info begin
        rc = ccall((:nagp_hmc, libnagp), Int32,
            (Ptr{Cvoid}, Int64, Int64, Ptr{UInt8}, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64},
             Int32, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Int32}, Float64,
             Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
            c, K, P, fl.prog, fl.prog_off, fl.theta_off, slot_kind, slot_a, slot_b,
            noise_spec[1], noise_spec[2], noise_spec[3], z, noise_z, n, k, t, g === nothing ? C_NULL : g, step,
            fl.y1, 0, k == 0 ? C_NULL : y2, n_steps, n_leapfrog, eps, mom, noise_spec[1] == 5 ? C_NULL : mnz, logu,
            logml, nacc, info)
        rc < 0 && check(c, rc)
    end
    end
    return logml, nacc
end

"""
The last host pass of forecast_with_nowcasts on the device: `inv_transformation.(x)` for the three built-in
transformations of `get_transformations` (kind 1 "positive", 2 "percentage", 3 "boxcox" with (λ, offset, max)),
in place, plus the per-date quantiles at `probs` (Julia's default definition). Returns the `(h, length(probs))` matrix.
"""
function forecast_summary!(x::Matrix{Float64}, kind::Integer, λ::Float64, offset::Float64, max_value::Float64,
        probs::Vector{Float64} = [0.25, 0.5, 0.75])
    h, N = size(x); nq = length(probs)
    q = Matrix{Float64}(undef, nq, h)           # row-major [h, nq] on the C side
    with_ctx() do c
    GC.@preserve x probs q begin
        check(c, ccall((:nagp_forecast_summary, libnagp), Int32,
            (Ptr{Cvoid}, Int32, Float64, Float64, Float64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Int64,
             Ptr{Float64}, Ptr{Float64}),
            c, Int32(kind), λ, offset, max_value, h, N, x, x, nq, probs, q))
    end
    end
    return permutedims(q)
end

end # module

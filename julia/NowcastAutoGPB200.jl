# NowcastAutoGPB200.jl — thin ccall shim over libnagp.so (include/nagp.h) for NowcastAutoGP.jl.
#
# NEVER EXECUTED IN THE BUILD IMAGE (no Julia there): kept deliberately thin — flatten the particles,
# one ccall, reshape. The same C ABI is tested end to end from Python (nowcastautogp_b200/engine.py).
# See INTEGRATION.md for where a maintainer hooks this into src/forecasting.jl.
module NowcastAutoGPB200

using AutoGP, Dates, LinearAlgebra, Random

const libnagp = get(ENV, "NAGP_LIB", joinpath(@__DIR__, "..", "nowcastautogp_b200", "libnagp.so"))
const _ctx = Dict{Int, Ptr{Cvoid}}()
const _lock = ReentrantLock()

available() = isfile(libnagp)

"One context per Julia thread: a nagp_ctx is not re-entrant."
function ctx()
    tid = Threads.threadid()
    lock(_lock) do
        get!(_ctx, tid) do
            out = Ref{Ptr{Cvoid}}(C_NULL)
            rc = ccall((:nagp_init, libnagp), Int32, (Int32, Ptr{Ptr{Cvoid}}), 0, out)
            rc == 0 || error("nagp_init: " * unsafe_string(ccall((:nagp_last_error, libnagp), Cstring, (Ptr{Cvoid},), C_NULL)))
            out[]
        end
    end
end

function check(rc::Int32)
    rc > 0 && throw(LinearAlgebra.PosDefException(rc))   # /root/reference/test/test_model_fitting.jl:97-98
    rc < 0 && error("libnagp: " * unsafe_string(ccall((:nagp_last_error, libnagp), Cstring, (Ptr{Cvoid},), ctx())))
    return nothing
end

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

"Drop-in for the body of forecast_with_nowcasts when n_mcmc == n_hmc == 0 and forecast_n_hmc === nothing."
function forecast_with_nowcasts_b200(base_model::AutoGP.GPModel, nowcasts::AbstractVector, forecast_dates, D::Int;
        inv_transformation = y -> y, ess_threshold = 0.0)
    @assert !isempty(nowcasts) "nowcasts vector must not be empty"
    @assert 0.0 <= ess_threshold <= 1.0 "ess_threshold must be between 0 and 1"
    fl = flatten_model(base_model)
    K, P = length(nowcasts), fl.P
    dates = collect(forecast_dates)
    k, h, n = length(nowcasts[1].ds), length(dates), length(fl.y1)
    t, g, step = time_arguments(fl, vcat(nowcasts[1].ds, dates))
    y2 = Matrix{Float64}(undef, k, K)
    for (s, nc) in enumerate(nowcasts); y2[:, s] .= fl.ya .* nc.y .+ fl.yb; end
    comp = fill(Int32(-1), D, K); u = rand(D, K)
    u_res = ess_threshold > 0 ? rand(P, K) : nothing
    zeta = randn(h, D, K)
    x = Matrix{Float64}(undef, h, K * D); info = zeros(Int32, P)
    GC.@preserve fl t g y2 comp u u_res zeta x info begin
        rc = ccall((:nagp_forecast_with_nowcasts, libnagp), Int32,
            (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{UInt8}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Float64,
             Int64, Int64, Int64, Ptr{Float64}, Ptr{Int32}, Float64, Ptr{Float64}, Ptr{Float64}, Float64, Float64,
             Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
            ctx(), K, P, D, fl.prog, fl.prog_off, fl.theta, fl.theta_off, fl.noise, -1.0,
            n, k, h, t, g === nothing ? C_NULL : g, step, fl.y1, y2, fl.ya, fl.yb,
            fl.logw, comp, u, u_res === nothing ? C_NULL : u_res, Float64(ess_threshold), zeta,
            x, C_NULL, C_NULL, info)
        check(rc)
    end
    return inv_transformation.(x)
end

"Batched log marginal likelihood of the model's particles over (t, y) — the fit_smc! primitive."
function logml_batch(fl::Flat, t::Vector{Float64}, g, step::Float64, y::Vector{Float64})
    out = Vector{Float64}(undef, fl.P); info = Vector{Int32}(undef, fl.P)
    GC.@preserve fl t g y out info begin
        rc = ccall((:nagp_logml_batch, libnagp), Int32,
            (Ptr{Cvoid}, Int64, Ptr{UInt8}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Int64,
             Ptr{Float64}, Ptr{Int32}, Float64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int32}),
            ctx(), fl.P, fl.prog, fl.prog_off, fl.theta, fl.theta_off, fl.noise, length(t),
            t, g === nothing ? C_NULL : g, step, y, 0, out, info)
        check(rc)
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
        y2::Matrix{Float64})
    K, P, total, k, n = size(theta_k, 2), fl.P, length(fl.theta), size(y2, 1), length(fl.y1)
    logml = Matrix{Float64}(undef, P, K); gth = Matrix{Float64}(undef, total, K); gnz = Matrix{Float64}(undef, P, K)
    info = Matrix{Int32}(undef, P, K)
    GC.@preserve fl theta_k noise_k t g y2 logml gth gnz info begin
        rc = ccall((:nagp_logml_grad, libnagp), Int32,
            (Ptr{Cvoid}, Int64, Int64, Ptr{UInt8}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Int64, Ptr{Float64}, Int64,
             Int64, Int64, Ptr{Float64}, Ptr{Int32}, Float64, Ptr{Float64}, Int64, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
            ctx(), K, P, fl.prog, fl.prog_off, theta_k, fl.theta_off, total, noise_k, P, n, k, t,
            g === nothing ? C_NULL : g, step, fl.y1, 0, k == 0 ? C_NULL : y2, logml, gth, gnz, info)
        rc < 0 && check(rc)      # > 0 only flags chains whose Gram is not PD: their logml is NaN, the sampler rejects
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
        rng = Random.default_rng())
    K, P, total, k, n = size(z, 2), fl.P, length(fl.theta), size(y2, 1), length(fl.y1)
    mom = randn(rng, total, K, n_steps); mnz = randn(rng, P, K, n_steps); logu = log.(rand(rng, P, K, n_steps))
    logml = Matrix{Float64}(undef, P, K); nacc = Matrix{Int32}(undef, P, K); info = Matrix{Int32}(undef, P, K)
    GC.@preserve fl slot_kind slot_a slot_b z noise_z t g y2 mom mnz logu logml nacc info begin
        rc = ccall((:nagp_hmc, libnagp), Int32,
            (Ptr{Cvoid}, Int64, Int64, Ptr{UInt8}, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64},
             Int32, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Int32}, Float64,
             Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
            ctx(), K, P, fl.prog, fl.prog_off, fl.theta_off, slot_kind, slot_a, slot_b,
            noise_spec[1], noise_spec[2], noise_spec[3], z, noise_z, n, k, t, g === nothing ? C_NULL : g, step,
            fl.y1, 0, k == 0 ? C_NULL : y2, n_steps, n_leapfrog, eps, mom, noise_spec[1] == 5 ? C_NULL : mnz, logu,
            logml, nacc, info)
        rc < 0 && check(rc)
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
    GC.@preserve x probs q begin
        check(ccall((:nagp_forecast_summary, libnagp), Int32,
            (Ptr{Cvoid}, Int32, Float64, Float64, Float64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Int64,
             Ptr{Float64}, Ptr{Float64}),
            ctx(), Int32(kind), λ, offset, max_value, h, N, x, x, nq, probs, q))
    end
    return permutedims(q)
end

end # module

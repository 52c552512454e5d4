# dump_golden.jl — emits TRUE reference golden vectors for the hot path (run off-box; needs Julia >= 1.11,
# AutoGP >= 0.1.13, JSON3). `julia --project julia/dump_golden.jl > tests/golden/autogp_golden.json`
#
# The build image has no Julia, so tests/golden/ currently holds mpmath / scikit-learn / closed-form
# vectors only (DESIGN.md §3, "parity unpinned"). Once this script's output is committed,
# tests/test_autogp_golden.py picks it up (test_autogp_golden for the CPU oracle, test_autogp_golden_gpu for the CUDA path)
# and both are pinned to AutoGP itself.
using AutoGP, Dates, Random, JSON3, Distributions
include(joinpath(@__DIR__, "NowcastAutoGPB200.jl"))
using .NowcastAutoGPB200: flatten_model, time_arguments

Random.seed!(20261018)
n, k, h = 60, 2, 4
ds = [Date(2022, 1, 1) + Week(i) for i in 0:(n + k + h - 1)]
tt = collect(0:(n + k + h - 1))
y = log(50) .+ sin.(2pi .* tt ./ 52) .+ 0.02 .* tt .+ 0.15 .* randn(length(tt))   # docs/vignettes/setting-priors.jl:96-98

cases = []
for P in (1, 4)
    model = AutoGP.GPModel(ds[1:n], y[1:n]; n_particles = P)
    AutoGP.fit_smc!(model; schedule = AutoGP.Schedule.linear_schedule(n, 0.25), n_mcmc = 5, n_hmc = 3, verbose = false)
    fl = flatten_model(model)
    t, g, step = time_arguments(fl, ds[(n + 1):end])
    logw_before = copy(AutoGP.log_weights(model))
    AutoGP.add_data!(model, ds[(n + 1):(n + k)], y[(n + 1):(n + k)])
    logw_after = copy(AutoGP.log_weights(model))
    dist = AutoGP.predict_mvn(model, ds[(n + k + 1):end])
    comps = Distributions.components(dist)
    Random.seed!(7); draws = rand(dist, 5)
    push!(cases, (; P, n, k, h, prog = fl.prog, prog_off = fl.prog_off, theta = fl.theta, theta_off = fl.theta_off,
        noise = fl.noise, t, g, step, y1 = fl.y1, ya = fl.ya, yb = fl.yb, y_new = y[(n + 1):(n + k)],
        logw_before, logw_after, mu = [mean(c) for c in comps], Sigma = [Matrix(cov(c)) for c in comps],
        weights = Distributions.probs(dist), draws_seed7 = draws))
end
JSON3.write(stdout, (; autogp_version = string(pkgversion(AutoGP)), cases))

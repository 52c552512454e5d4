/*
 * nagp.h — C ABI of the B200-native GP hot path behind NowcastAutoGP's public API.
 *
 * The reference (CDCgov/NowcastAutoGP, pure Julia) has no FFI of its own: its boundary to the
 * arithmetic is 13 plain Julia calls into AutoGP.jl. Each entry point below replaces the
 * arithmetic of one (or a fused run) of those call sites; the reference file:line is cited per
 * function. INTEGRATION.md shows the `ccall` stubs a maintainer adds on the Julia side.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; all sizes int64_t, reals double, status int32_t.
 *  - Every data pointer may be a HOST pointer (pageable or pinned) or a DEVICE pointer on the
 *    context's GPU; the library detects which (cudaPointerGetAttributes). Host inputs are copied
 *    in, host outputs are copied back and the call blocks until they have landed. If every output
 *    is a device pointer the call is asynchronous on the context's stream.
 *  - The caller owns every buffer it passes. The library owns device memory behind nagp_ctx /
 *    nagp_factor handles, released by nagp_destroy / nagp_factor_free.
 *  - Return: 0 ok; >0 LAPACK-style "leading minor i not positive definite" for at least one
 *    instance (per-instance codes in info[], the shim maps it to Julia's PosDefException —
 *    behaviour pinned by /root/reference/test/test_model_fitting.jl:97-110); <0 bad argument or
 *    CUDA failure, message via nagp_last_error(). No exception crosses the boundary.
 *  - A context is not re-entrant: one call at a time per nagp_ctx (use one ctx per host thread).
 *  - Kernel trees travel as post-order byte programs + theta (docs/KERNEL_SPEC.md §1), packed
 *    CSR-style: instance p's program is prog[prog_off[p] .. prog_off[p+1]).
 *  - Times: t[q] rescaled time points; g (nullable) int32 grid indices with `step` selects the
 *    lag-grid contract of KERNEL_SPEC §2 (Δ_ij = |g_i − g_j|·step).
 *  - Point order of a forecast problem: [n training | k nowcast | h forecast], m = n+k, q = m+h.
 *  - Sizes: q <= 232 runs entirely in shared memory; 232 < q <= 4096 keeps the factor in HBM
 *    (same entry points, same results; NAGP_E_SIZE beyond).
 */
#ifndef NAGP_H
#define NAGP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NAGP_OK 0
#define NAGP_E_ARG (-1)      /* bad argument                                  */
#define NAGP_E_CUDA (-2)     /* CUDA runtime failure                          */
#define NAGP_E_PROGRAM (-3)  /* malformed kernel program / stack or length cap */
#define NAGP_E_SIZE (-4)     /* problem does not fit the implemented paths     */

typedef struct nagp_ctx nagp_ctx;
typedef struct nagp_factor nagp_factor;

/* ---- lifetime ------------------------------------------------------------------------------ */
int32_t nagp_version(void);
/* Create a context on CUDA device `device`. */
int32_t nagp_init(int32_t device, nagp_ctx **out);
void nagp_destroy(nagp_ctx *ctx);
/* Last error text of this context (or of nagp_init when ctx == NULL). Never NULL. */
const char *nagp_last_error(const nagp_ctx *ctx);
/* Run subsequent calls on `cuda_stream` (a cudaStream_t; NULL = the context's own stream). */
int32_t nagp_set_stream(nagp_ctx *ctx, void *cuda_stream);
/* Diagonal jitter added to every Gram (AutoGP: 1e-5, KERNEL_SPEC §4). */
int32_t nagp_set_jitter(nagp_ctx *ctx, double jitter);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t nagp_launch_count(const nagp_ctx *ctx);
/* Pick the factorisation kernel for q <= 232: 0 = auto, 1 = shared-memory column kernel, 2 = tile kernel,
 * 3 = slot kernel (three matrices per SM) wherever it applies (lag-grid times, q <= 168), else as 2,
 * 4 = the factor-in-HBM kernel of the large path at every size (cross-check; 11.9 ms against 5.1 ms at the vignette size). */
int32_t nagp_set_variant(nagp_ctx *ctx, int32_t variant);
/* Which factorisation kernel the last fused launch of this context used: 1 column, 2 tile, 3 slot, 4 large; 0 none yet. */
int32_t nagp_last_kernel(const nagp_ctx *ctx);

/* ---- (a2) batched log marginal likelihood ---------------------------------------------------
 * Replaces the per-particle Gram -> dpotrf -> logdet/quad-form that AutoGP.fit_smc! /
 * mcmc_structure! evaluate (/root/reference/src/make_and_fit_model.jl:84-91,
 * /root/reference/src/forecasting.jl:146). B instances over the same n points; y is shared
 * (y_stride == 0) or per instance (y + b*y_stride). logml[B], info[B]. */
int32_t nagp_logml_batch(nagp_ctx *ctx, int64_t B,
                         const uint8_t *prog, const int64_t *prog_off,
                         const double *theta, const int64_t *theta_off, const double *noise,
                         int64_t n, const double *t, const int32_t *g, double step,
                         const double *y, int64_t y_stride,
                         double *logml, int32_t *info);

/* ---- general per-(scenario, particle) path ---------------------------------------------------
 * One fused Gram -> Cholesky -> solves per instance for K scenarios x P particles: replaces
 * GPModel(dict) + add_data! + predict_mvn + MvNormal's Cholesky for each of them
 * (/root/reference/src/forecasting.jl:133,135,46). Hyperparameters may differ per scenario
 * (theta_stride_k = total theta length, noise_stride_k = P; as after per-scenario HMC,
 * forecasting.jl:145-149) or be shared (strides 0). y1[n] and y2[K*k] are in AutoGP's scaled
 * space, (ya, yb) the scale map y_s = ya*y + yb used to un-scale predictions; noise_pred < 0
 * means "use each instance's own noise on the forecast block".
 * Outputs: logw[K*P] = logw0[p] + logML(m) - logML(n); mu[K*P*h]; L[K*P*h*h] (row-major lower
 * Cholesky factor of the predictive covariance, original units); info[K*P]; logml_n / logml_m
 * [K*P] (nullable) the two log marginal likelihoods themselves — what a Metropolis/HMC accept
 * step on the per-scenario hyperparameters needs (mcmc_parameters!, forecasting.jl:148,65). */
int32_t nagp_forecast_instances(nagp_ctx *ctx, int64_t K, int64_t P,
                                const uint8_t *prog, const int64_t *prog_off,
                                const double *theta, const int64_t *theta_off, int64_t theta_stride_k,
                                const double *noise, int64_t noise_stride_k, double noise_pred,
                                int64_t n, int64_t k, int64_t h,
                                const double *t, const int32_t *g, double step,
                                const double *y1, const double *y2, double ya, double yb,
                                const double *logw0,
                                double *logw, double *mu, double *L, int32_t *info,
                                double *logml_n, double *logml_m);

/* ---- (a3) factor store: Dict(model) / GPModel(dict) -------------------------------------------
 * Factor the P particles of a base model ONCE over [train | nowcast dates | forecast dates] and
 * keep the factors on the device (/root/reference/src/forecasting.jl:128,133 rebuild them K times).
 * logml_n[P] (nullable) receives each particle's training log marginal likelihood. */
int32_t nagp_factor_store(nagp_ctx *ctx, int64_t P,
                          const uint8_t *prog, const int64_t *prog_off,
                          const double *theta, const int64_t *theta_off,
                          const double *noise, double noise_pred,
                          int64_t n, int64_t k, int64_t h,
                          const double *t, const int32_t *g, double step,
                          const double *y1, double ya, double yb, const double *logw0,
                          nagp_factor **out, double *logml_n, int32_t *info);
void nagp_factor_free(nagp_factor *f);

/* ---- (f1) gradient of the log marginal likelihood: what HMC on the hyperparameters differentiates ----
 * AutoGP.mcmc_parameters! (/root/reference/src/forecasting.jl:148 and :65) and fit_smc!'s n_hmc steps
 * (/root/reference/src/make_and_fit_model.jl:91) run HMC on each particle's hyperparameters; the gradient of
 * the MVN log density is the expensive part. For K scenarios x P particles over the n + k observed points
 * (y = [y1 | y2[s]]; theta/noise shared or per scenario as in nagp_forecast_instances):
 *   logml[K*P]           log marginal likelihood over all n + k points
 *   grad_theta[K*total]  d logML / d theta for every theta slot, same CSR layout as theta (per scenario)
 *   grad_noise[K*P]      d logML / d noise
 * y1_stride: 0 = every instance shares y1[n]; n = instance b = s*P + p reads y1 + b*n (independent series fitted
 * in lockstep: one "particle" per (series, particle) pair, K = 1).
 * Constrained-space derivatives; the caller applies its own chain rule to the unconstrained parameters.
 * n + k <= 232 in this version (NAGP_E_SIZE beyond). */
int32_t nagp_logml_grad(nagp_ctx *ctx, int64_t K, int64_t P,
                        const uint8_t *prog, const int64_t *prog_off,
                        const double *theta, const int64_t *theta_off, int64_t theta_stride_k,
                        const double *noise, int64_t noise_stride_k,
                        int64_t n, int64_t k, const double *t, const int32_t *g, double step,
                        const double *y1, int64_t y1_stride, const double *y2,
                        double *logml, double *grad_theta, double *grad_noise, int32_t *info);

/* ---- (f1) mcmc_parameters!: HMC on the unconstrained hyperparameters, leapfrog on the device -------------------
 * /root/reference/src/forecasting.jl:148 and :65; the n_hmc steps of fit_smc! (/root/reference/src/make_and_fit_model.jl:91).
 * K scenarios x P particles = K*P chains advance together; nothing returns to the host inside a chain: each
 * leapfrog stage is [half kick, drift, z -> theta] -> logML + gradient kernels -> [chain rule, half kick], each
 * iteration ends with one accept/reject kernel, and the iteration is replayed as one CUDA graph.
 * Target: log p(y | theta(z)) + log N(z; 0, I). z -> theta per slot (slot_kind[total], parameters slot_a/slot_b):
 *   0 exp(a + b z)   2 2*logistic(a + b z)   3 z   4 Phi(z)   5 the constant a (not sampled);
 * the noise uses (noise_kind, noise_a, noise_b) the same way (kind 5: fixed noise = noise_a, noise_z untouched).
 *   z[K*total], noise_z[K*P]    host, in/out: the chains' unconstrained states (total = theta_off[P])
 *   momenta[n_steps*K*total], noise_momenta[n_steps*K*P] (NULL iff noise_kind == 5), log_u[n_steps*K*P]:
 *       the standard-normal momenta and log-uniform accept thresholds of every iteration, supplied by the caller
 *       (as the normals of the forecast draws are), so a host integrator fed the same numbers walks the same chain
 *   logml[K*P] log marginal likelihood of the final state, n_accept[K*P] accepted iterations, info[K*P] factorisation
 *       status of the final state (> 0: the INITIAL state was not positive definite and no proposal was accepted).
 * y1/y1_stride/y2, t/g/step as in nagp_logml_grad. n + k <= 232 (NAGP_E_SIZE beyond). */
int32_t nagp_hmc(nagp_ctx *ctx, int64_t K, int64_t P,
                 const uint8_t *prog, const int64_t *prog_off, const int64_t *theta_off,
                 const int32_t *slot_kind, const double *slot_a, const double *slot_b,
                 int32_t noise_kind, double noise_a, double noise_b,
                 double *z, double *noise_z,
                 int64_t n, int64_t k, const double *t, const int32_t *g, double step,
                 const double *y1, int64_t y1_stride, const double *y2,
                 int64_t n_steps, int64_t n_leapfrog, double eps,
                 const double *momenta, const double *noise_momenta, const double *log_u,
                 double *logml, int32_t *n_accept, int32_t *info);

/* ---- (a2/a4) appendable factor for long series: SMC data annealing, rank-append Cholesky ---------
 * AutoGP.fit_smc! walks a schedule of growing observation counts (/root/reference/src/make_and_fit_model.jl:
 * 89-91, linear_schedule) and re-scores every particle from scratch at each step; add_data! does the same
 * for one batch (/root/reference/src/forecasting.jl:135). nagp_factor_store_large factors the P particles
 * over the first n points (n <= capacity <= 4096) and keeps the WHOLE factor in device memory;
 * nagp_factor_append extends it in place by k_new points (streams the stored factor once: HBM-bound,
 * k/4 FLOP per byte) and returns dlogml[P] = logML(n + k_new) - logML(n) (the SMC log-weight increment)
 * and, if logml != NULL, the running logML[P]. Valid while the particles' (tree, theta) are unchanged;
 * after a rejuvenation move the caller stores a new factor. prog/offsets/t/g/y must be host arrays;
 * y in scaled space; g/g_new both NULL (pairwise times) or both non-NULL (lag grid, same origin). */
int32_t nagp_factor_store_large(nagp_ctx *ctx, int64_t P,
                                const uint8_t *prog, const int64_t *prog_off,
                                const double *theta, const int64_t *theta_off, const double *noise,
                                int64_t n, int64_t capacity,
                                const double *t, const int32_t *g, double step, const double *y,
                                nagp_factor **out, double *logml, int32_t *info);
int32_t nagp_factor_append(nagp_ctx *ctx, nagp_factor *f, int64_t k_new,
                           const double *t_new, const int32_t *g_new, const double *y_new,
                           double *dlogml, double *logml, int32_t *info);
/* Number of points currently held by a factor (-1 for NULL). */
int64_t nagp_factor_size(const nagp_factor *f);

/* ---- (a4) add_data! for K scenarios against a stored factor -----------------------------------
 * /root/reference/src/forecasting.jl:135. y2[K*k] scaled. logw[K*P] = logw0 + Δ logML;
 * mu[K*P*h] (nullable) = per-scenario predictive means in original units. */
int32_t nagp_append(nagp_ctx *ctx, const nagp_factor *f, int64_t K, const double *y2,
                    double *logw, double *mu);

/* ---- (a7) predict_mvn for the stored particles (no nowcast values: k must be 0) ---------------
 * /root/reference/src/forecasting.jl:46. mu[P*h], L[P*h*h] original units. With k > 0 use
 * nagp_append for the means; L is scenario-independent and is returned here in either case. */
int32_t nagp_predict(nagp_ctx *ctx, const nagp_factor *f, double *mu, double *L);

/* ---- (a5) maybe_resample!'s test: ESS per scenario --------------------------------------------
 * /root/reference/src/forecasting.jl:138-141. logw[K*P] -> ess[K], w[K*P] (nullable). */
int32_t nagp_ess(nagp_ctx *ctx, int64_t K, int64_t P, const double *logw, double *ess, double *w);

/* ---- (a8) rand(MixtureModel, D) ----------------------------------------------------------------
 * /root/reference/src/forecasting.jl:47. Per scenario s and draw d: component c = comp[s*D+d] if
 * comp != NULL and the entry is >= 0, else by inverse CDF of the normalised weights at u[s*D+d]
 * (after multinomial resampling with u_res[K*P] iff u_res != NULL and ESS < ess_thr*P);
 * x = mu_c + L_c * zeta in the fixed order of KERNEL_SPEC §7. mu/L are [K,P,h]/[K,P,h,h] with
 * scenario strides mu_stride_k / l_stride_k in doubles (0 = shared by all scenarios).
 * x[h, K*D] column-major (scenario-major column blocks); ess_out[K], comp_out[K*D] nullable. */
int32_t nagp_draw(nagp_ctx *ctx, int64_t K, int64_t P, int64_t h, int64_t D,
                  const double *logw, const double *mu, int64_t mu_stride_k,
                  const double *L, int64_t l_stride_k,
                  const int32_t *comp, const double *u, const double *u_res, double ess_thr,
                  const double *zeta, double *x, double *ess_out, int32_t *comp_out);

/* ---- fused forecast_with_nowcasts (n_mcmc = n_hmc = 0, forecast_n_hmc = nothing) ---------------
 * /root/reference/src/forecasting.jl:117-167 in one call: factor once per particle, append the K
 * scenarios, ESS/resample, draw. x[h, K*D] column-major; logw_out[K*P], ess_out[K] nullable. */
int32_t nagp_forecast_with_nowcasts(nagp_ctx *ctx, int64_t K, int64_t P, int64_t D,
                                    const uint8_t *prog, const int64_t *prog_off,
                                    const double *theta, const int64_t *theta_off,
                                    const double *noise, double noise_pred,
                                    int64_t n, int64_t k, int64_t h,
                                    const double *t, const int32_t *g, double step,
                                    const double *y1, const double *y2, double ya, double yb,
                                    const double *logw0,
                                    const int32_t *comp, const double *u, const double *u_res,
                                    double ess_thr, const double *zeta,
                                    double *x, double *logw_out, double *ess_out, int32_t *info);

/* The same body for PER-SCENARIO hyperparameters — what the particles look like after `mcmc_parameters!` has run on every
 * scenario's model copy (/root/reference/src/forecasting.jl:145-149 with n_hmc > 0, then :152-155): one fused Gram +
 * factorisation per (scenario, particle), log-weights, ESS/resample, draws, in ONE call (nagp_forecast_instances +
 * nagp_draw without the round trip of the moments). theta/noise strides as in nagp_forecast_instances; info[K*P]. */
int32_t nagp_forecast_with_nowcasts_theta(nagp_ctx *ctx, int64_t K, int64_t P, int64_t D,
                                          const uint8_t *prog, const int64_t *prog_off,
                                          const double *theta, const int64_t *theta_off, int64_t theta_stride_k,
                                          const double *noise, int64_t noise_stride_k, double noise_pred,
                                          int64_t n, int64_t k, int64_t h,
                                          const double *t, const int32_t *g, double step,
                                          const double *y1, const double *y2, double ya, double yb,
                                          const double *logw0,
                                          const int32_t *comp, const double *u, const double *u_res,
                                          double ess_thr, const double *zeta,
                                          double *x, double *logw_out, double *ess_out, int32_t *info);

/* ---- (f4) inverse transformation + per-date quantiles of a forecast matrix --------------------------------
 * /root/reference/src/forecasting.jl:50,73,166 apply `inv_transformation.(x)` to the (h, K*D) draws on the host,
 * and every vignette then takes row quantiles (/root/reference/docs/vignettes/getting-started.jl:432-435).
 * For the three built-in transformations of get_transformations (/root/reference/src/transformations.jl:134-170)
 * both happen here on the device. kind: 0 identity, 1 "positive" max(exp(y) - offset, 0), 2 "percentage"
 * max(logistic(y) * 100 - offset, 0), 3 "boxcox" (the inverse with the clamping rules of
 * /root/reference/src/transformations.jl:6-44; lambda, offset, max_value = maximum of the fitted values).
 * x[h, N] column-major (the layout nagp_draw / nagp_forecast_with_nowcasts write; host or device);
 * x_out (nullable; may alias x) receives the transformed matrix; q[h, nq] row-major (nullable iff nq == 0)
 * receives, per forecast date, Julia's default `quantile` (type 7) of the TRANSFORMED draws at probs[nq]. */
int32_t nagp_forecast_summary(nagp_ctx *ctx, int32_t kind, double lambda, double offset, double max_value,
                              int64_t h, int64_t N, const double *x, double *x_out,
                              int64_t nq, const double *probs, double *q);

#ifdef __cplusplus
}
#endif
#endif /* NAGP_H */

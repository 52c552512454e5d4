/*
 * test_c_abi.c — the drop-in boundary exercised from plain C: includes include/nagp.h, links libnagp.so, runs ONE
 * nagp_forecast_with_nowcasts call (the fused replacement of /root/reference/src/forecasting.jl:117-167 for
 * n_mcmc == n_hmc == 0) with host buffers, and checks the results against closed forms that need no oracle:
 *   - a single Constant(v) particle: K = v 11^T + s I, so logML(n) has the Sherman-Morrison closed form;
 *   - log-weight of a scenario = logw0 + logML(n+k) - logML(n);
 *   - every draw is finite and the output has the reference's (h, K*D) column-major layout.
 * Built and run by tests/test_c_abi_program.py (gcc ... -L nowcastautogp_b200 -lnagp); exits 0 on success.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "nagp.h"

static double logml_constant(int n, const double *y, double v, double s)
{
    /* K = v 11^T + s I: logdet = (n-1) log s + log(s + n v); y^T K^-1 y = (y.y - v (1.y)^2 / (s + n v)) / s */
    double sy = 0, syy = 0;
    for (int i = 0; i < n; ++i) { sy += y[i]; syy += y[i] * y[i]; }
    const double logdet = (n - 1) * log(s) + log(s + n * v);
    const double quad = (syy - v * sy * sy / (s + n * v)) / s;
    return -0.5 * (n * log(2.0 * M_PI) + logdet + quad);
}

int main(void)
{
    enum { n = 24, k = 2, h = 3, q = n + k + h, K = 5, P = 1, D = 4 };
    nagp_ctx *ctx = NULL;
    if (nagp_init(0, &ctx) != NAGP_OK) { fprintf(stderr, "nagp_init: %s\n", nagp_last_error(NULL)); return 2; }
    const uint8_t prog[1] = {1};                 /* Constant */
    const int64_t prog_off[2] = {0, 1}, theta_off[2] = {0, 1};
    const double v = 0.7, noise = 0.3, jitter = 1e-5;
    const double theta[1] = {v}, nz[1] = {noise}, logw0[1] = {-0.25};
    double t[q], y1[n], y2[K * k], zeta[K * D * h], u[K * D], x[h * K * D], logw[K * P];
    int32_t g[q], info[P];
    for (int i = 0; i < q; ++i) { t[i] = i / (double)(n - 1); g[i] = i; }
    for (int i = 0; i < n; ++i) y1[i] = sin(0.37 * i) + 0.1 * i / n;
    for (int i = 0; i < K * k; ++i) y2[i] = 0.2 * cos(1.3 * i);
    for (int i = 0; i < K * D * h; ++i) zeta[i] = sin(12.9898 * (i + 1)) * 1.7;
    for (int i = 0; i < K * D; ++i) u[i] = (i * 0.618033988749895) - floor(i * 0.618033988749895);
    const double ya = 1.0, yb = 0.0;
    int32_t rc = nagp_forecast_with_nowcasts(ctx, K, P, D, prog, prog_off, theta, theta_off, nz, -1.0, n, k, h, t, g, 1.0 / (n - 1),
                                             y1, y2, ya, yb, logw0, NULL, u, NULL, 0.0, zeta, x, logw, NULL, info);
    if (rc != NAGP_OK) { fprintf(stderr, "nagp_forecast_with_nowcasts rc=%d: %s\n", rc, nagp_last_error(ctx)); return 3; }
    int bad = 0;
    for (int s = 0; s < K; ++s) {
        double y[n + k];
        for (int i = 0; i < n; ++i) y[i] = y1[i];
        for (int i = 0; i < k; ++i) y[n + i] = y2[s * k + i];
        const double want = logw0[0] + logml_constant(n + k, y, v, noise + jitter) - logml_constant(n, y, v, noise + jitter);
        const double err = fabs(logw[s] - want) / fabs(want);
        if (!(err < 1e-9)) { fprintf(stderr, "scenario %d: logw %.15g, closed form %.15g (rel %.2e)\n", s, logw[s], want, err); bad = 1; }
    }
    for (int i = 0; i < h * K * D; ++i) if (!isfinite(x[i])) { fprintf(stderr, "draw %d not finite\n", i); bad = 1; }
    if (info[0] != 0) { fprintf(stderr, "info = %d\n", info[0]); bad = 1; }
    if (nagp_launch_count(ctx) < 3) { fprintf(stderr, "expected >= 3 kernel launches, got %lld\n", (long long)nagp_launch_count(ctx)); bad = 1; }
    nagp_destroy(ctx);
    if (!bad) printf("c-abi ok: %d scenarios, %d draws, log-weights match the closed form\n", (int)K, (int)(K * D));
    return bad;
}

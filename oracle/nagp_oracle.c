/*
 * nagp_oracle.c — CPU restatement of the reference's GP arithmetic for the hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product path (nowcastautogp_b200/) may link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker or the timed CPU baseline.
 *
 * PARITY UNPINNED: the arithmetic restated here lives in AutoGP.jl (uuid 6eb593e7-…, compat
 * 0.1.13, /root/reference/Project.toml:7,15), which is not vendored under /root/reference and
 * cannot run in this image (no Julia). The reference's own tests pin no numeric GP output
 * (SURVEY.md §4). Formulas follow docs/KERNEL_SPEC.md; the call order follows the reference's
 * call sites:
 *   - per-scenario schedule: /root/reference/src/forecasting.jl:131-155
 *   - predict + draw:        /root/reference/src/forecasting.jl:39-52
 *   - batched logML:         /root/reference/src/make_and_fit_model.jl:84-91 (fit_smc!)
 * What is pinned: kernel formulas against scikit-learn and mpmath, logML against closed forms
 * and a __float128 build of this same file (tests/test_oracle_pinning.py).
 *
 * Build: see oracle/Makefile (double build: libnagp_oracle.so; -DNAGP_QUAD: libnagp_oracle_q.so).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef NAGP_QUAD
#include <quadmath.h>
typedef __float128 real;
#define R_EXP expq
#define R_POW powq
#define R_SIN sinq
#define R_TANH tanhq
#define R_LOG logq
#define R_SQRT sqrtq
#define R_FABS fabsq
#define R_FMA fmaq
#define R_PI M_PIq
#define SYM(name) name##_q
#else
typedef double real;
#define R_EXP exp
#define R_POW pow
#define R_SIN sin
#define R_TANH tanh
#define R_LOG log
#define R_SQRT sqrt
#define R_FABS fabs
#define R_FMA fma
#define R_PI M_PI
#define SYM(name) name
#endif

enum {
    OP_CONSTANT = 1, OP_LINEAR = 2, OP_SQEXP = 3, OP_GAMMAEXP = 4,
    OP_PERIODIC = 5, OP_PLUS = 6, OP_TIMES = 7, OP_CHANGEPOINT = 8
};
#define NAGP_MAX_PROG 64
#define NAGP_MAX_STACK 16
#define NAGP_E_PROGRAM (-3)

static const int k_nparam[9] = {0, 1, 3, 2, 3, 3, 0, 0, 2};

/* Validate a post-order program; returns #theta slots consumed (>=0) or NAGP_E_PROGRAM. */
int64_t SYM(nagp_o_prog_check)(const uint8_t *prog, int64_t len)
{
    if (len <= 0 || len > NAGP_MAX_PROG) return NAGP_E_PROGRAM;
    int64_t sp = 0, nth = 0;
    for (int64_t i = 0; i < len; ++i) {
        int op = prog[i];
        if (op < 1 || op > 8) return NAGP_E_PROGRAM;
        nth += k_nparam[op];
        if (op <= OP_PERIODIC) {
            if (++sp > NAGP_MAX_STACK) return NAGP_E_PROGRAM;
        } else {
            if (sp < 2) return NAGP_E_PROGRAM;
            --sp;
        }
    }
    return sp == 1 ? nth : NAGP_E_PROGRAM;
}

/* k(ti,tj) for one pair; delta = |ti - tj| (pairwise) or |gi-gj|*step (lag grid).
 * docs/KERNEL_SPEC.md §3 — each line is the stated evaluation order. */
static real eval_pair(const uint8_t *prog, int64_t len, const double *theta,
                      real ti, real tj, real delta)
{
    real st[NAGP_MAX_STACK];
    int sp = 0;
    const double *th = theta;
    for (int64_t i = 0; i < len; ++i) {
        switch (prog[i]) {
        case OP_CONSTANT:
            st[sp++] = (real)th[0]; th += 1; break;
        case OP_LINEAR: {
            real u = ti - (real)th[0], w = tj - (real)th[0];
            st[sp++] = R_FMA((real)th[2], u * w, (real)th[1]); th += 3; break;
        }
        case OP_SQEXP: {
            real r = delta / (real)th[0];
            st[sp++] = (real)th[1] * R_EXP((real)-0.5 * (r * r)); th += 2; break;
        }
        case OP_GAMMAEXP: {
            real r = delta / (real)th[0];
            st[sp++] = (real)th[2] * R_EXP(-R_POW(r, (real)th[1])); th += 3; break;
        }
        case OP_PERIODIC: {
            real s = R_SIN(R_PI * (delta / (real)th[1]));
            real l = (real)th[0];
            st[sp++] = (real)th[2] * R_EXP((real)-2.0 * (s * s) / (l * l)); th += 3; break;
        }
        case OP_PLUS:
            st[sp - 2] = st[sp - 2] + st[sp - 1]; --sp; break;
        case OP_TIMES:
            st[sp - 2] = st[sp - 2] * st[sp - 1]; --sp; break;
        case OP_CHANGEPOINT: {
            real si = (real)0.5 * ((real)1.0 + R_TANH((ti - (real)th[0]) / (real)th[1]));
            real sj = (real)0.5 * ((real)1.0 + R_TANH((tj - (real)th[0]) / (real)th[1]));
            real kl = st[sp - 2], kr = st[sp - 1];
            st[sp - 2] = (((real)1.0 - si) * ((real)1.0 - sj)) * kl + (si * sj) * kr;
            --sp; th += 2; break;
        }
        default: break;
        }
    }
    return st[0];
}

double SYM(nagp_o_kernel_pair)(const uint8_t *prog, int64_t len, const double *theta,
                               double ti, double tj, double delta)
{
    return (double)eval_pair(prog, len, theta, (real)ti, (real)tj, (real)delta);
}

static inline real pair_delta(const double *t, const int32_t *g, double step, int64_t i, int64_t j)
{
    if (g) {
        int32_t d = g[i] - g[j];
        if (d < 0) d = -d;
#ifdef NAGP_QUAD
        return (real)d * (real)step;
#else
        return (double)d * step;
#endif
    }
    return R_FABS((real)t[i] - (real)t[j]);
}

/* Gram over q points into K (row-major, ld = q, full symmetric). diag_lo is added on i<m, diag_hi
 * on i>=m (KERNEL_SPEC §4). */
static void gram_real(const uint8_t *prog, int64_t len, const double *theta,
                      real diag_lo, real diag_hi, int64_t m, int64_t q,
                      const double *t, const int32_t *g, double step, real *K)
{
    for (int64_t i = 0; i < q; ++i)
        for (int64_t j = 0; j <= i; ++j) {
            real v = eval_pair(prog, len, theta, (real)t[i], (real)t[j], pair_delta(t, g, step, i, j));
            if (i == j) v += (i < m) ? diag_lo : diag_hi;
            K[i * q + j] = v;
            K[j * q + i] = v;
        }
}

int32_t SYM(nagp_o_gram)(const uint8_t *prog, int64_t len, const double *theta,
                         double diag_lo, double diag_hi, int64_t m, int64_t q,
                         const double *t, const int32_t *g, double step, double *K_out)
{
    if (SYM(nagp_o_prog_check)(prog, len) < 0) return NAGP_E_PROGRAM;
    real *K = (real *)malloc(sizeof(real) * q * q);
    gram_real(prog, len, theta, (real)diag_lo, (real)diag_hi, m, q, t, g, step, K);
    for (int64_t i = 0; i < q * q; ++i) K_out[i] = (double)K[i];
    free(K);
    return 0;
}

/* Lower Cholesky in place (row-major, ld), left-looking dot-product form. Returns 0 or the
 * 1-based index of the first non-positive pivot (LAPACK dpotrf convention). */
static int32_t potrf_real(int64_t n, real *A, int64_t ld)
{
    for (int64_t j = 0; j < n; ++j) {
        real d = A[j * ld + j];
        for (int64_t p = 0; p < j; ++p) d -= A[j * ld + p] * A[j * ld + p];
        if (!(d > 0)) return (int32_t)(j + 1);
        d = R_SQRT(d);
        A[j * ld + j] = d;
        for (int64_t i = j + 1; i < n; ++i) {
            real s = A[i * ld + j];
            for (int64_t p = 0; p < j; ++p) s -= A[i * ld + p] * A[j * ld + p];
            A[i * ld + j] = s / d;
        }
    }
    return 0;
}

static void trsv_lower_real(int64_t n, const real *L, int64_t ld, real *b)
{
    for (int64_t i = 0; i < n; ++i) {
        real s = b[i];
        for (int64_t p = 0; p < i; ++p) s -= L[i * ld + p] * b[p];
        b[i] = s / L[i * ld + i];
    }
}

static real logml_from(int64_t n, const real *L, int64_t ld, const real *z)
{
    real logdet = 0, quad = 0;
    for (int64_t i = 0; i < n; ++i) {
        logdet += R_LOG(L[i * ld + i]);
        quad += z[i] * z[i];
    }
    const real log2pi = R_LOG((real)2.0 * R_PI);
    return (real)-0.5 * ((real)n * log2pi + (real)2.0 * logdet + quad);
}

/* LU with partial pivoting, in place (row-major). Julia's `\` on a dense square Matrix (KERNEL_SPEC §5). */
static int32_t getrf_real(int64_t n, real *A, int64_t ld, int64_t *piv)
{
    for (int64_t j = 0; j < n; ++j) {
        int64_t p = j;
        real best = R_FABS(A[j * ld + j]);
        for (int64_t i = j + 1; i < n; ++i)
            if (R_FABS(A[i * ld + j]) > best) { best = R_FABS(A[i * ld + j]); p = i; }
        piv[j] = p;
        if (best == 0) return (int32_t)(j + 1);
        if (p != j)
            for (int64_t c = 0; c < n; ++c) {
                real tmp = A[j * ld + c]; A[j * ld + c] = A[p * ld + c]; A[p * ld + c] = tmp;
            }
        real inv = (real)1.0 / A[j * ld + j];
        for (int64_t i = j + 1; i < n; ++i) {
            real f = A[i * ld + j] * inv;
            A[i * ld + j] = f;
            for (int64_t c = j + 1; c < n; ++c) A[i * ld + c] -= f * A[j * ld + c];
        }
    }
    return 0;
}

static void getrs_real(int64_t n, const real *LU, int64_t ld, const int64_t *piv, real *b)
{
    for (int64_t j = 0; j < n; ++j)
        if (piv[j] != j) { real tmp = b[j]; b[j] = b[piv[j]]; b[piv[j]] = tmp; }
    for (int64_t i = 0; i < n; ++i) {
        real s = b[i];
        for (int64_t p = 0; p < i; ++p) s -= LU[i * ld + p] * b[p];
        b[i] = s;
    }
    for (int64_t i = n - 1; i >= 0; --i) {
        real s = b[i];
        for (int64_t p = i + 1; p < n; ++p) s -= LU[i * ld + p] * b[p];
        b[i] = s / LU[i * ld + i];
    }
}

/* logML of one (program, theta, noise) instance on the first n points — the fit_smc! primitive.
 * Returns info. */
int32_t SYM(nagp_o_logml)(const uint8_t *prog, int64_t len, const double *theta, double noise,
                          double jitter, int64_t n, const double *t, const int32_t *g, double step,
                          const double *y, double *logml_out)
{
    if (SYM(nagp_o_prog_check)(prog, len) < 0) return NAGP_E_PROGRAM;
    real *K = (real *)malloc(sizeof(real) * n * n);
    real *z = (real *)malloc(sizeof(real) * n);
    real d = (real)noise + (real)jitter;
    gram_real(prog, len, theta, d, d, n, n, t, g, step, K);
    int32_t info = potrf_real(n, K, n);
    if (info == 0) {
        for (int64_t i = 0; i < n; ++i) z[i] = (real)y[i];
        trsv_lower_real(n, K, n, z);
        *logml_out = (double)logml_from(n, K, n, z);
    } else {
        *logml_out = NAN;
    }
    free(K); free(z);
    return info;
}

/*
 * One (scenario, particle) instance following the REFERENCE schedule (KERNEL_SPEC §5):
 * rebuild(n) -> add_data!(m = n+k) -> predict_mvn(q = m+h, LU solves) -> MvNormal Cholesky(h).
 * y has m entries in scaled space; outputs mu[h] and Lsig[h*h] (row-major lower, zeros above) are
 * un-scaled with (ya, yb). noise_pred < 0 => forecast block uses the instance noise.
 */
int32_t SYM(nagp_o_instance_reference)(const uint8_t *prog, int64_t len, const double *theta,
                                       double noise, double jitter, double noise_pred,
                                       int64_t n, int64_t k, int64_t h,
                                       const double *t, const int32_t *g, double step,
                                       const double *y, double ya, double yb,
                                       double *logml_n, double *logml_m, double *mu, double *Lsig)
{
    if (SYM(nagp_o_prog_check)(prog, len) < 0) return NAGP_E_PROGRAM;
    const int64_t m = n + k, q = m + h;
    real d_lo = (real)noise + (real)jitter;
    real d_hi = (noise_pred >= 0 ? (real)noise_pred : (real)noise) + (real)jitter;
    int32_t info = 0;
    real *K = (real *)malloc(sizeof(real) * q * q);
    real *A = (real *)malloc(sizeof(real) * q * q);
    real *z = (real *)malloc(sizeof(real) * q);
    int64_t *piv = (int64_t *)malloc(sizeof(int64_t) * q);
    real lmn = 0, lmm = 0;

    /* (1) GPModel(dict): likelihood of the n training points, src/forecasting.jl:133 */
    gram_real(prog, len, theta, d_lo, d_lo, n, n, t, g, step, A);
    info = potrf_real(n, A, n);
    if (info) goto done;
    for (int64_t i = 0; i < n; ++i) z[i] = (real)y[i];
    trsv_lower_real(n, A, n, z);
    lmn = logml_from(n, A, n, z);

    /* (2) add_data!: likelihood of all m points from scratch, src/forecasting.jl:135 */
    gram_real(prog, len, theta, d_lo, d_lo, m, m, t, g, step, A);
    info = potrf_real(m, A, m);
    if (info) goto done;
    for (int64_t i = 0; i < m; ++i) z[i] = (real)y[i];
    trsv_lower_real(m, A, m, z);
    lmm = logml_from(m, A, m, z);

    /* (3) predict_mvn: joint Gram, conditional via LU solves, src/forecasting.jl:46 */
    if (h > 0) {
        gram_real(prog, len, theta, d_lo, d_hi, m, q, t, g, step, K);
        for (int64_t i = 0; i < m; ++i)
            for (int64_t j = 0; j < m; ++j) A[i * m + j] = K[i * q + j];
        info = getrf_real(m, A, m, piv);
        if (info) goto done;
        for (int64_t i = 0; i < m; ++i) z[i] = (real)y[i];
        getrs_real(m, A, m, piv, z);                       /* K11 \ y */
        real *X = (real *)malloc(sizeof(real) * m * h);    /* column c = K11 \ K12[:,c] */
        real *S = (real *)malloc(sizeof(real) * h * h);
        real *col = (real *)malloc(sizeof(real) * m);
        for (int64_t c = 0; c < h; ++c) {
            for (int64_t i = 0; i < m; ++i) col[i] = K[i * q + (m + c)];
            getrs_real(m, A, m, piv, col);
            for (int64_t i = 0; i < m; ++i) X[i * h + c] = col[i];
        }
        for (int64_t r = 0; r < h; ++r) {
            real acc = 0;
            for (int64_t i = 0; i < m; ++i) acc += K[(m + r) * q + i] * z[i];
            mu[r] = (double)((acc - (real)yb) / (real)ya);
            for (int64_t c = 0; c < h; ++c) {
                real s = 0;
                for (int64_t i = 0; i < m; ++i) s += K[(m + r) * q + i] * X[i * h + c];
                S[r * h + c] = K[(m + r) * q + (m + c)] - s;
            }
        }
        for (int64_t r = 0; r < h; ++r)                    /* 0.5 (S + S') then / a^2 */
            for (int64_t c = 0; c <= r; ++c) {
                real v = ((real)0.5 * S[r * h + c] + (real)0.5 * S[c * h + r]) / ((real)ya * (real)ya);
                S[r * h + c] = v; S[c * h + r] = v;
            }
        /* (4) MvNormal ctor: Cholesky of Sigma*, throws PosDefException on failure */
        info = potrf_real(h, S, h);
        if (!info)
            for (int64_t r = 0; r < h; ++r)
                for (int64_t c = 0; c < h; ++c) Lsig[r * h + c] = c <= r ? (double)S[r * h + c] : 0.0;
        free(X); free(S); free(col);
    }
done:
    *logml_n = info ? NAN : (double)lmn;
    *logml_m = info ? NAN : (double)lmm;
    free(K); free(A); free(z); free(piv);
    return info;
}

/*
 * Same instance via the joint factorisation the device uses (KERNEL_SPEC §6). ny = n (no scenario
 * values yet; z has n entries) or m. Ltail = rows n..q-1 of L, row-major (q-n) x q (scaled space).
 */
int32_t SYM(nagp_o_instance_joint)(const uint8_t *prog, int64_t len, const double *theta,
                                   double noise, double jitter, double noise_pred,
                                   int64_t n, int64_t k, int64_t h,
                                   const double *t, const int32_t *g, double step,
                                   const double *y, int64_t ny, double ya, double yb,
                                   double *logml_n, double *logml_m, double *z_out,
                                   double *Ltail, double *mu, double *Lsig)
{
    if (SYM(nagp_o_prog_check)(prog, len) < 0) return NAGP_E_PROGRAM;
    const int64_t m = n + k, q = m + h;
    real d_lo = (real)noise + (real)jitter;
    real d_hi = (noise_pred >= 0 ? (real)noise_pred : (real)noise) + (real)jitter;
    real *K = (real *)malloc(sizeof(real) * q * q);
    real *z = (real *)malloc(sizeof(real) * q);
    gram_real(prog, len, theta, d_lo, d_hi, m, q, t, g, step, K);
    int32_t info = potrf_real(q, K, q);
    if (info) {
        *logml_n = NAN; *logml_m = NAN;
        free(K); free(z);
        return info;
    }
    for (int64_t i = 0; i < ny; ++i) z[i] = (real)y[i];
    trsv_lower_real(ny, K, q, z);
    *logml_n = (double)logml_from(n, K, q, z);
    *logml_m = ny >= m ? (double)logml_from(m, K, q, z) : NAN;
    if (z_out) for (int64_t i = 0; i < ny; ++i) z_out[i] = (double)z[i];
    if (Ltail)
        for (int64_t r = n; r < q; ++r)
            for (int64_t c = 0; c < q; ++c) Ltail[(r - n) * q + c] = c <= r ? (double)K[r * q + c] : 0.0;
    if (mu && ny >= m)
        for (int64_t r = 0; r < h; ++r) {
            real acc = 0;
            for (int64_t i = 0; i < m; ++i) acc += K[(m + r) * q + i] * z[i];
            mu[r] = (double)((acc - (real)yb) / (real)ya);
        }
    if (Lsig)
        for (int64_t r = 0; r < h; ++r)
            for (int64_t c = 0; c < h; ++c)
                Lsig[r * h + c] = c <= r ? (double)(K[(m + r) * q + (m + c)] / (real)ya) : 0.0;
    free(K); free(z);
    return 0;
}

#ifndef NAGP_QUAD
/* ---------- weights / ESS / resampling / draws: KERNEL_SPEC §7 (double only, bit-exact spec) ---------- */

void nagp_o_normalize(int64_t P, const double *logw, double *w, double *ess)
{
    double mx = -INFINITY;
    for (int64_t p = 0; p < P; ++p) if (logw[p] > mx) mx = logw[p];
    double sum = 0;
    for (int64_t p = 0; p < P; ++p) { w[p] = exp(logw[p] - mx); sum += w[p]; }
    double s2 = 0;
    for (int64_t p = 0; p < P; ++p) { w[p] = w[p] / sum; s2 += w[p] * w[p]; }
    *ess = 1.0 / s2;
}

int32_t nagp_o_invcdf(int64_t P, const double *w, double u)
{
    double run = 0;
    for (int64_t p = 0; p < P; ++p) {
        run += w[p];
        if (run > u) return (int32_t)p;
    }
    return (int32_t)(P - 1);
}

/*
 * Draws for K scenarios. mu [K,P,h], L [K,P,h,h] (row-major lower), logw [K,P] -> x [h, K*D]
 * column-major. comp [K,D] (>=0 entries win) or u [K,D]; u_res [K,P] + ess_thr for resampling.
 * mu_stride_k / l_stride_k let scenario-shared factors be passed with stride 0.
 */
void nagp_o_draws(int64_t K, int64_t P, int64_t h, int64_t D,
                  const double *logw, const double *mu, int64_t mu_stride_k,
                  const double *L, int64_t l_stride_k,
                  const int32_t *comp, const double *u, const double *u_res, double ess_thr,
                  const double *zeta, double *x, double *ess_out, int32_t *comp_out)
{
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < K; ++s) {
        double *w = (double *)malloc(sizeof(double) * P);
        int32_t *parent = (int32_t *)malloc(sizeof(int32_t) * P);
        double ess;
        nagp_o_normalize(P, logw + s * P, w, &ess);
        if (ess_out) ess_out[s] = ess;
        int resampled = 0;
        for (int64_t p = 0; p < P; ++p) parent[p] = (int32_t)p;
        if (u_res && ess < ess_thr * (double)P) {
            for (int64_t p = 0; p < P; ++p) parent[p] = nagp_o_invcdf(P, w, u_res[s * P + p]);
            resampled = 1;
        }
        for (int64_t d = 0; d < D; ++d) {
            int32_t c;
            if (comp && comp[s * D + d] >= 0) c = comp[s * D + d];
            else if (resampled) {
                int64_t slot = (int64_t)(u[s * D + d] * (double)P);
                if (slot >= P) slot = P - 1;
                c = parent[slot];
            } else c = nagp_o_invcdf(P, w, u[s * D + d]);
            if (comp_out) comp_out[s * D + d] = c;
            const double *mc = mu + s * mu_stride_k + (int64_t)c * h;
            const double *Lc = L + s * l_stride_k + (int64_t)c * h * h;
            const double *zz = zeta + (s * D + d) * h;
            double *xo = x + (s * D + d) * h;
            for (int64_t i = 0; i < h; ++i) {
                double acc = mc[i];
                for (int64_t j = 0; j <= i; ++j) acc = fma(Lc[i * h + j], zz[j], acc);
                xo[i] = acc;
            }
        }
        free(w); free(parent);
    }
}

/*
 * forecast_with_nowcasts, reference schedule, OpenMP over (scenario, particle) like the reference's
 * task-per-scenario x thread-per-particle (src/forecasting.jl:131-132, :1). Shared tree per particle
 * (theta_stride_k = 0) or per-(scenario, particle) hyperparameters (theta_stride_k = total theta).
 * y1 [n] scaled; y2 [K,k] scaled. Outputs: logw [K,P], mu [K,P,h], Lsig [K,P,h,h], info [K,P].
 * use_joint = 0: three factorisations per instance as the reference does; 1: joint factorisation.
 */
int32_t nagp_o_forecast_instances(int64_t K, int64_t P,
                                  const uint8_t *prog, const int64_t *prog_off,
                                  const double *theta, const int64_t *theta_off, int64_t theta_stride_k,
                                  const double *noise, int64_t noise_stride_k,
                                  double jitter, double noise_pred,
                                  int64_t n, int64_t k, int64_t h,
                                  const double *t, const int32_t *g, double step,
                                  const double *y1, const double *y2, double ya, double yb,
                                  const double *logw0, int32_t use_joint,
                                  double *logw, double *mu, double *Lsig, int32_t *info)
{
    const int64_t m = n + k;
    int32_t worst = 0;
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
    for (int64_t s = 0; s < K; ++s)
        for (int64_t p = 0; p < P; ++p) {
            double *y = (double *)malloc(sizeof(double) * m);
            memcpy(y, y1, sizeof(double) * n);
            memcpy(y + n, y2 + s * k, sizeof(double) * k);
            const uint8_t *pr = prog + prog_off[p];
            int64_t len = prog_off[p + 1] - prog_off[p];
            const double *th = theta + s * theta_stride_k + theta_off[p];
            double nz = noise[s * noise_stride_k + p];
            double lmn, lmm;
            int32_t inf;
            double *mo = mu + (s * P + p) * h, *Lo = Lsig + (s * P + p) * h * h;
            if (use_joint)
                inf = nagp_o_instance_joint(pr, len, th, nz, jitter, noise_pred, n, k, h, t, g, step,
                                            y, m, ya, yb, &lmn, &lmm, NULL, NULL, mo, Lo);
            else
                inf = nagp_o_instance_reference(pr, len, th, nz, jitter, noise_pred, n, k, h, t, g,
                                                step, y, ya, yb, &lmn, &lmm, mo, Lo);
            info[s * P + p] = inf;
            logw[s * P + p] = logw0[p] + (lmm - lmn);
            if (inf) {
#pragma omp atomic write
                worst = inf;
            }
            free(y);
        }
    return worst;
}

/* Batched logML (fit_smc! primitive), OpenMP over instances. y shared [n]. */
int32_t nagp_o_logml_batch(int64_t B, const uint8_t *prog, const int64_t *prog_off,
                           const double *theta, const int64_t *theta_off, const double *noise,
                           double jitter, int64_t n, const double *t, const int32_t *g, double step,
                           const double *y, double *logml, int32_t *info)
{
    int32_t worst = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        info[b] = nagp_o_logml(prog + prog_off[b], prog_off[b + 1] - prog_off[b],
                               theta + theta_off[b], noise[b], jitter, n, t, g, step, y, &logml[b]);
        if (info[b]) {
#pragma omp atomic write
            worst = info[b];
        }
    }
    return worst;
}

/* bench.py's CPU arm: use every host core even when the launcher (torchrun) exported OMP_NUM_THREADS=1 */
void nagp_o_set_num_threads(int32_t n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int32_t nagp_o_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
#endif /* !NAGP_QUAD */

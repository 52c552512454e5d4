/*
 * nagp_cpu_blocked.c — the TIMED CPU baseline of bench.py (`cpu_baseline`, `--impl reference`).
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY (same rule as nagp_oracle.c: nothing under nowcastautogp_b200/ may link
 * or call it). The checker stays nagp_oracle.c: scalar, unblocked, -O2 -ffp-contract=off, evaluation order as
 * stated in docs/KERNEL_SPEC.md. A scalar unblocked factorisation is a soft baseline though (about 1.6 GFLOP/s per
 * core), and the reference itself runs LAPACK (OpenBLAS dpotrf / dgetrf through PDMats and `\`,
 * /root/reference/src/forecasting.jl:46,133,135). This file restates the SAME reference schedule
 *   rebuild(n) -> add_data!(m) -> predict_mvn(q, LU solves) -> MvNormal Cholesky(h)
 * (/root/reference/src/forecasting.jl:131-155, :39-52) with blocked, AVX2/FMA-vectorised kernels, built -O3:
 *   - Cholesky: left-looking by block column, the update as 4x2 register-blocked dot products over contiguous rows;
 *   - LU with partial pivoting: right-looking by panel, the trailing update as a 4x8 register-blocked GEMM;
 *   - triangular solves as vectorised dot products / axpys.
 * Its results are checked against nagp_oracle.c in tests/test_cpu_baseline.py (1e-9 relative), so the number
 * bench.py reports is the time of a correct computation.
 *
 * Build: oracle/Makefile -> libnagp_cpu_blocked.so (-O3 -march=x86-64-v3 -fopenmp).
 */
#define nagp_o_prog_check nagp_b_prog_check_
#define nagp_o_kernel_pair nagp_b_kernel_pair_
#define nagp_o_gram nagp_b_gram_
#define nagp_o_logml nagp_b_logml_
#define nagp_o_instance_reference nagp_b_instance_reference_
#define nagp_o_instance_joint nagp_b_instance_joint_
#define nagp_o_normalize nagp_b_normalize_
#define nagp_o_invcdf nagp_b_invcdf_
#define nagp_o_draws nagp_b_draws_
#define nagp_o_forecast_instances nagp_b_forecast_instances_
#define nagp_o_logml_batch nagp_b_logml_batch_
#define nagp_o_set_num_threads nagp_b_set_num_threads_
#define nagp_o_num_threads nagp_b_num_threads_
#include "nagp_oracle.c"   /* kernel-tree evaluation and Gram construction, compiled here at -O3 */

#include <immintrin.h>

#define NB 16

static inline double hsum4(__m256d v)
{
    __m128d lo = _mm256_castpd256_pd128(v), hi = _mm256_extractf128_pd(v, 1);
    lo = _mm_add_pd(lo, hi);
    return _mm_cvtsd_f64(_mm_add_sd(lo, _mm_unpackhi_pd(lo, lo)));
}

/* out[r][c] = sum_{k<K} A[r*lda + k] * B[c*ldb + k], r < 4, c < 2 (rows beyond the matrix are clamped by the caller) */
static inline void dot4x2(const double *A, int64_t lda, const double *B, int64_t ldb, int64_t K, double out[4][2])
{
    __m256d acc[4][2];
    for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = _mm256_setzero_pd();
    int64_t k = 0;
    for (; k + 4 <= K; k += 4) {
        const __m256d b0 = _mm256_loadu_pd(B + k), b1 = _mm256_loadu_pd(B + ldb + k);
        for (int r = 0; r < 4; ++r) {
            const __m256d a = _mm256_loadu_pd(A + r * lda + k);
            acc[r][0] = _mm256_fmadd_pd(a, b0, acc[r][0]);
            acc[r][1] = _mm256_fmadd_pd(a, b1, acc[r][1]);
        }
    }
    for (int r = 0; r < 4; ++r) {
        double s0 = hsum4(acc[r][0]), s1 = hsum4(acc[r][1]);
        for (int64_t kk = k; kk < K; ++kk) {
            s0 += A[r * lda + kk] * B[kk];
            s1 += A[r * lda + kk] * B[ldb + kk];
        }
        out[r][0] = s0; out[r][1] = s1;
    }
}

static inline double dotv(const double *a, const double *b, int64_t K)
{
    __m256d acc = _mm256_setzero_pd();
    int64_t k = 0;
    for (; k + 4 <= K; k += 4) acc = _mm256_fmadd_pd(_mm256_loadu_pd(a + k), _mm256_loadu_pd(b + k), acc);
    double s = hsum4(acc);
    for (; k < K; ++k) s += a[k] * b[k];
    return s;
}

/* Lower Cholesky in place, row-major, blocked left-looking. Returns 0 or the 1-based index of the first
 * non-positive pivot. */
static int32_t potrf_blocked(int64_t n, double *A, int64_t ld)
{
    for (int64_t J0 = 0; J0 < n; J0 += NB) {
        const int64_t J1 = J0 + NB < n ? J0 + NB : n;
        if (J0 > 0) {
            /* A[i][j] -= A[i][0:J0] . A[j][0:J0] for J0 <= j < J1, j <= i < n */
            for (int64_t i = J0; i < n; i += 4) {
                const int nr = (int)(n - i < 4 ? n - i : 4);
                const int64_t jmax = J1 < i + nr ? J1 : i + nr;
                for (int64_t j = J0; j < jmax; j += 2) {
                    const int nc = (int)(J1 - j < 2 ? J1 - j : 2);
                    if (nr == 4 && nc == 2) {
                        double out[4][2];
                        dot4x2(A + i * ld, ld, A + j * ld, ld, J0, out);
                        for (int r = 0; r < 4; ++r)
                            for (int c = 0; c < 2; ++c)
                                if (j + c <= i + r) A[(i + r) * ld + j + c] -= out[r][c];
                    } else {
                        for (int r = 0; r < nr; ++r)
                            for (int c = 0; c < nc; ++c)
                                if (j + c <= i + r) A[(i + r) * ld + j + c] -= dotv(A + (i + r) * ld, A + (j + c) * ld, J0);
                    }
                }
            }
        }
        /* diagonal block and the rows below it, terms of the block's own columns */
        for (int64_t j = J0; j < J1; ++j) {
            double d = A[j * ld + j] - dotv(A + j * ld + J0, A + j * ld + J0, j - J0);
            if (!(d > 0)) return (int32_t)(j + 1);
            d = sqrt(d);
            A[j * ld + j] = d;
            const double rd = 1.0 / d;
            for (int64_t i = j + 1; i < n; ++i)
                A[i * ld + j] = (A[i * ld + j] - dotv(A + i * ld + J0, A + j * ld + J0, j - J0)) * rd;
        }
    }
    return 0;
}

static void trsv_lower_fast(int64_t n, const double *L, int64_t ld, double *b)
{
    for (int64_t i = 0; i < n; ++i) b[i] = (b[i] - dotv(L + i * ld, b, i)) / L[i * ld + i];
}

/* LU with partial pivoting in place, row-major, right-looking by panels of NB columns. */
static int32_t getrf_blocked(int64_t n, double *A, int64_t ld, int64_t *piv)
{
    for (int64_t J0 = 0; J0 < n; J0 += NB) {
        const int64_t J1 = J0 + NB < n ? J0 + NB : n;
        /* panel: columns [J0, J1), unblocked; row swaps are applied to whole rows */
        for (int64_t j = J0; j < J1; ++j) {
            int64_t p = j;
            double best = fabs(A[j * ld + j]);
            for (int64_t i = j + 1; i < n; ++i)
                if (fabs(A[i * ld + j]) > best) { best = fabs(A[i * ld + j]); p = i; }
            piv[j] = p;
            if (best == 0) return (int32_t)(j + 1);
            if (p != j)
                for (int64_t c = 0; c < n; ++c) { double tmp = A[j * ld + c]; A[j * ld + c] = A[p * ld + c]; A[p * ld + c] = tmp; }
            const double inv = 1.0 / A[j * ld + j];
            for (int64_t i = j + 1; i < n; ++i) {
                const double f = A[i * ld + j] * inv;
                A[i * ld + j] = f;
                for (int64_t c = j + 1; c < J1; ++c) A[i * ld + c] -= f * A[j * ld + c];
            }
        }
        if (J1 >= n) break;
        /* U12 = L11^-1 A12: rows J0..J1 of the columns right of the panel */
        for (int64_t j = J0 + 1; j < J1; ++j)
            for (int64_t pp = J0; pp < j; ++pp) {
                const __m256d f = _mm256_set1_pd(A[j * ld + pp]);
                int64_t c = J1;
                for (; c + 4 <= n; c += 4)
                    _mm256_storeu_pd(A + j * ld + c, _mm256_fnmadd_pd(f, _mm256_loadu_pd(A + pp * ld + c), _mm256_loadu_pd(A + j * ld + c)));
                for (; c < n; ++c) A[j * ld + c] -= A[j * ld + pp] * A[pp * ld + c];
            }
        /* A22 -= L21 U12: 4 rows x 8 columns per register block, k over the panel */
        const int64_t kb = J1 - J0;
        for (int64_t i = J1; i < n; i += 4) {
            const int nr = (int)(n - i < 4 ? n - i : 4);
            int64_t c = J1;
            for (; c + 8 <= n; c += 8) {
                __m256d acc[4][2];
                for (int r = 0; r < nr; ++r) { acc[r][0] = _mm256_loadu_pd(A + (i + r) * ld + c); acc[r][1] = _mm256_loadu_pd(A + (i + r) * ld + c + 4); }
                for (int64_t kk = 0; kk < kb; ++kk) {
                    const __m256d u0 = _mm256_loadu_pd(A + (J0 + kk) * ld + c), u1 = _mm256_loadu_pd(A + (J0 + kk) * ld + c + 4);
                    for (int r = 0; r < nr; ++r) {
                        const __m256d l = _mm256_set1_pd(A[(i + r) * ld + J0 + kk]);
                        acc[r][0] = _mm256_fnmadd_pd(l, u0, acc[r][0]);
                        acc[r][1] = _mm256_fnmadd_pd(l, u1, acc[r][1]);
                    }
                }
                for (int r = 0; r < nr; ++r) { _mm256_storeu_pd(A + (i + r) * ld + c, acc[r][0]); _mm256_storeu_pd(A + (i + r) * ld + c + 4, acc[r][1]); }
            }
            for (; c < n; ++c)
                for (int r = 0; r < nr; ++r) {
                    double sacc = A[(i + r) * ld + c];
                    for (int64_t kk = 0; kk < kb; ++kk) sacc -= A[(i + r) * ld + J0 + kk] * A[(J0 + kk) * ld + c];
                    A[(i + r) * ld + c] = sacc;
                }
        }
    }
    return 0;
}

/* Solve with the factors of getrf_blocked for nrhs right-hand sides stored as ROWS of B (B[c*ldb + i]). */
static void getrs_rows(int64_t n, const double *LU, int64_t ld, const int64_t *piv, double *b)
{
    for (int64_t j = 0; j < n; ++j)
        if (piv[j] != j) { double tmp = b[j]; b[j] = b[piv[j]]; b[piv[j]] = tmp; }
    for (int64_t i = 0; i < n; ++i) b[i] -= dotv(LU + i * ld, b, i);
    for (int64_t i = n - 1; i >= 0; --i) b[i] = (b[i] - dotv(LU + i * ld + i + 1, b + i + 1, n - 1 - i)) / LU[i * ld + i];
}

/* One (scenario, particle) instance, reference schedule — the blocked twin of nagp_o_instance_reference. */
static int32_t instance_reference_blocked(const uint8_t *prog, int64_t len, const double *theta, double noise, double jitter,
                                          double noise_pred, int64_t n, int64_t k, int64_t h, const double *t,
                                          const int32_t *g, double step, const double *y, double ya, double yb,
                                          double *logml_n, double *logml_m, double *mu, double *Lsig, double *work)
{
    const int64_t m = n + k, q = m + h;
    const double d_lo = noise + jitter, d_hi = (noise_pred >= 0 ? noise_pred : noise) + jitter;
    double *K = work, *A = K + q * q, *z = A + q * q, *X = z + q, *S = X + h * m;
    int64_t *piv = (int64_t *)(S + h * h);
    int32_t info;
    double lmn = 0, lmm = 0;

    gram_real(prog, len, theta, d_lo, d_lo, n, n, t, g, step, A);              /* (1) GPModel(dict) */
    if ((info = potrf_blocked(n, A, n))) goto done;
    memcpy(z, y, sizeof(double) * n);
    trsv_lower_fast(n, A, n, z);
    lmn = logml_from(n, A, n, z);

    gram_real(prog, len, theta, d_lo, d_lo, m, m, t, g, step, A);              /* (2) add_data! */
    if ((info = potrf_blocked(m, A, m))) goto done;
    memcpy(z, y, sizeof(double) * m);
    trsv_lower_fast(m, A, m, z);
    lmm = logml_from(m, A, m, z);

    if (h > 0) {                                                                /* (3) predict_mvn */
        gram_real(prog, len, theta, d_lo, d_hi, m, q, t, g, step, K);
        for (int64_t i = 0; i < m; ++i) memcpy(A + i * m, K + i * q, sizeof(double) * m);
        if ((info = getrf_blocked(m, A, m, piv))) goto done;
        memcpy(z, y, sizeof(double) * m);
        getrs_rows(m, A, m, piv, z);                                            /* K11 \ y */
        for (int64_t c = 0; c < h; ++c) {                                       /* row c of X = (K11 \ K12[:, c])' */
            for (int64_t i = 0; i < m; ++i) X[c * m + i] = K[i * q + (m + c)];
            getrs_rows(m, A, m, piv, X + c * m);
        }
        for (int64_t r = 0; r < h; ++r) {
            mu[r] = (dotv(K + (m + r) * q, z, m) - yb) / ya;
            for (int64_t c = 0; c < h; ++c) S[r * h + c] = K[(m + r) * q + (m + c)] - dotv(K + (m + r) * q, X + c * m, m);
        }
        for (int64_t r = 0; r < h; ++r)
            for (int64_t c = 0; c <= r; ++c) {
                const double v = (0.5 * S[r * h + c] + 0.5 * S[c * h + r]) / (ya * ya);
                S[r * h + c] = v; S[c * h + r] = v;
            }
        info = potrf_blocked(h, S, h);                                          /* (4) MvNormal ctor */
        if (!info)
            for (int64_t r = 0; r < h; ++r)
                for (int64_t c = 0; c < h; ++c) Lsig[r * h + c] = c <= r ? S[r * h + c] : 0.0;
    }
done:
    *logml_n = info ? NAN : lmn;
    *logml_m = info ? NAN : lmm;
    return info;
}

static size_t work_doubles(int64_t n, int64_t k, int64_t h)
{
    const int64_t m = n + k, q = m + h;
    return (size_t)(2 * q * q + q + h * m + h * h + q + 8);
}

/* forecast_with_nowcasts, reference schedule, OpenMP over (scenario, particle): same arguments and outputs as
 * nagp_o_forecast_instances with use_joint = 0. */
int32_t nagp_b_forecast_instances(int64_t K, int64_t P, const uint8_t *prog, const int64_t *prog_off, const double *theta,
                                  const int64_t *theta_off, int64_t theta_stride_k, const double *noise,
                                  int64_t noise_stride_k, double jitter, double noise_pred, int64_t n, int64_t k, int64_t h,
                                  const double *t, const int32_t *g, double step, const double *y1, const double *y2,
                                  double ya, double yb, const double *logw0, double *logw, double *mu, double *Lsig,
                                  int32_t *info)
{
    const int64_t m = n + k;
    int32_t worst = 0;
#pragma omp parallel
    {
        double *work = (double *)malloc(sizeof(double) * work_doubles(n, k, h));
        double *y = (double *)malloc(sizeof(double) * (m + 1));
#pragma omp for collapse(2) schedule(dynamic, 1)
        for (int64_t s = 0; s < K; ++s)
            for (int64_t p = 0; p < P; ++p) {
                memcpy(y, y1, sizeof(double) * n);
                memcpy(y + n, y2 + s * k, sizeof(double) * k);
                double lmn, lmm;
                const int32_t inf = instance_reference_blocked(
                    prog + prog_off[p], prog_off[p + 1] - prog_off[p], theta + s * theta_stride_k + theta_off[p],
                    noise[s * noise_stride_k + p], jitter, noise_pred, n, k, h, t, g, step, y, ya, yb, &lmn, &lmm,
                    mu + (s * P + p) * h, Lsig + (s * P + p) * h * h, work);
                info[s * P + p] = inf;
                logw[s * P + p] = logw0[p] + (lmm - lmn);
                if (inf) {
#pragma omp atomic write
                    worst = inf;
                }
            }
        free(work); free(y);
    }
    return worst;
}

/* Batched logML (fit_smc! primitive; BASELINE configs[2]), OpenMP over instances. */
int32_t nagp_b_logml_batch(int64_t B, const uint8_t *prog, const int64_t *prog_off, const double *theta,
                           const int64_t *theta_off, const double *noise, double jitter, int64_t n, const double *t,
                           const int32_t *g, double step, const double *y, double *logml, int32_t *info)
{
    int32_t worst = 0;
#pragma omp parallel
    {
        double *A = (double *)malloc(sizeof(double) * (n * n + n));
        double *z = A + n * n;
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < B; ++b) {
            const double d = noise[b] + jitter;
            gram_real(prog + prog_off[b], prog_off[b + 1] - prog_off[b], theta + theta_off[b], d, d, n, n, t, g, step, A);
            info[b] = potrf_blocked(n, A, n);
            if (info[b] == 0) {
                memcpy(z, y, sizeof(double) * n);
                trsv_lower_fast(n, A, n, z);
                logml[b] = logml_from(n, A, n, z);
            } else {
                logml[b] = NAN;
#pragma omp atomic write
                worst = info[b];
            }
        }
        free(A);
    }
    return worst;
}

/* Factor B instances and keep the factors: L [B, n_cap, n_cap] row-major (leading dimension n_cap), the first n rows
 * filled; z [B, n_cap] = L^-1 y. The starting point of nagp_b_append (BASELINE configs[4]). */
int32_t nagp_b_factor_store(int64_t B, const uint8_t *prog, const int64_t *prog_off, const double *theta,
                            const int64_t *theta_off, const double *noise, double jitter, int64_t n, int64_t n_cap,
                            const double *t, const int32_t *g, double step, const double *y, double *L, double *z,
                            double *logml)
{
    int32_t worst = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        double *A = L + (size_t)b * n_cap * n_cap;
        const double d = noise[b] + jitter;
        const uint8_t *pr = prog + prog_off[b];
        const int64_t len = prog_off[b + 1] - prog_off[b];
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j <= i; ++j) {
                double v = eval_pair(pr, len, theta + theta_off[b], t[i], t[j], pair_delta(t, g, step, i, j));
                if (i == j) v += d;
                A[i * n_cap + j] = v;
            }
        const int32_t inf = potrf_blocked(n, A, n_cap);
        double *zb = z + (size_t)b * n_cap;
        if (!inf) {
            memcpy(zb, y, sizeof(double) * n);
            trsv_lower_fast(n, A, n_cap, zb);
            logml[b] = logml_from(n, A, n_cap, zb);
        } else {
            logml[b] = NAN;
#pragma omp atomic write
            worst = inf;
        }
    }
    return worst;
}

/* add_data! as a rank-append on stored factors: k new points extend every factor from n_old to n_old + k rows in
 * place (new row l = L^-1 k_new by forward substitution: the stored factor is read once), dlogml = logML(n_old + k) -
 * logML(n_old). t/g/y cover all n_old + k points. */
int32_t nagp_b_append(int64_t B, const uint8_t *prog, const int64_t *prog_off, const double *theta, const int64_t *theta_off,
                      const double *noise, double jitter, int64_t n_old, int64_t k, int64_t n_cap, const double *t,
                      const int32_t *g, double step, const double *y, double *L, double *z, double *dlogml)
{
    int32_t worst = 0;
    const double log2pi = log(2.0 * M_PI);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        double *A = L + (size_t)b * n_cap * n_cap;
        double *zb = z + (size_t)b * n_cap;
        const uint8_t *pr = prog + prog_off[b];
        const int64_t len = prog_off[b + 1] - prog_off[b];
        double acc = 0;
        int32_t inf = 0;
        for (int64_t r = n_old; r < n_old + k && !inf; ++r) {
            double *row = A + r * n_cap;
            for (int64_t j = 0; j <= r; ++j) row[j] = eval_pair(pr, len, theta + theta_off[b], t[r], t[j], pair_delta(t, g, step, r, j));
            row[r] += noise[b] + jitter;
            for (int64_t j = 0; j < r; ++j) row[j] = (row[j] - dotv(row, A + j * n_cap, j)) / A[j * n_cap + j];
            const double d = row[r] - dotv(row, row, r);
            if (!(d > 0)) { inf = (int32_t)(r + 1); break; }
            row[r] = sqrt(d);
            zb[r] = (y[r] - dotv(row, zb, r)) / row[r];
            acc += -0.5 * (log2pi + 2.0 * log(row[r]) + zb[r] * zb[r]);
        }
        dlogml[b] = inf ? NAN : acc;
        if (inf) {
#pragma omp atomic write
            worst = inf;
        }
    }
    return worst;
}

void nagp_b_set_num_threads(int32_t n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int32_t nagp_b_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Single-thread GFLOP/s of the blocked factorisations alone on a synthetic SPD matrix of order n (n^3/3 FLOP for
 * Cholesky, 2n^3/3 for LU), so bench.py can print what the linear algebra of the timed baseline sustains next to the
 * end-to-end figure (which is dominated by the per-entry kernel evaluation, as it is in the reference). */
void nagp_b_factor_rates(int64_t n, int64_t reps, double *potrf_gflops, double *getrf_gflops)
{
    double *A0 = (double *)malloc(sizeof(double) * n * n), *A = (double *)malloc(sizeof(double) * n * n);
    int64_t *piv = (int64_t *)malloc(sizeof(int64_t) * n);
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < n; ++j) {
            const double d = (double)(i > j ? i - j : j - i) / (double)n;
            A0[i * n + j] = exp(-4.0 * d * d) + (i == j ? 0.1 : 0.0);
        }
    double best_c = 1e30, best_l = 1e30;
    for (int64_t r = 0; r < reps; ++r) {
        memcpy(A, A0, sizeof(double) * n * n);
        double t0 = omp_get_wtime();
        potrf_blocked(n, A, n);
        double t1 = omp_get_wtime();
        if (t1 - t0 < best_c) best_c = t1 - t0;
        memcpy(A, A0, sizeof(double) * n * n);
        t0 = omp_get_wtime();
        getrf_blocked(n, A, n, piv);
        t1 = omp_get_wtime();
        if (t1 - t0 < best_l) best_l = t1 - t0;
    }
    *potrf_gflops = (double)n * n * n / 3.0 / best_c / 1e9;
    *getrf_gflops = 2.0 * (double)n * n * n / 3.0 / best_l / 1e9;
    free(A0); free(A); free(piv);
}

// nagp_summary.cu — the last host pass over the forecast matrix, on the device:
//   1. the built-in inverse transformations applied elementwise to the (h, K*D) draws, with the reference's
//      exact clamping rules (/root/reference/src/transformations.jl:6-44 Box-Cox, :145-146 percentage,
//      :149-150 positive) — replaces inv_transformation.(x) at /root/reference/src/forecasting.jl:50,73,166;
//   2. per-forecast-date quantiles of the transformed draws (Julia's default quantile definition, type 7),
//      what every vignette computes next (/root/reference/docs/vignettes/getting-started.jl:432-435).
//
// Both are HBM-bound byte work: the transform kernel streams 8 B in / 16 B out per element (the matrix in the
// caller's column-major layout plus a row-major copy so the selection reads are coalesced); the quantile kernel
// finds the two order statistics around each requested rank by an 8-pass most-significant-digit radix select
// over the row (16 B/element/pass from L2, no sort, no scratch beyond a 256-bin histogram in shared memory).
// Arithmetic contract: docs/KERNEL_SPEC.md §9.
#include "nagp_kernels.cuh"

namespace nagp {

namespace {

__device__ __forceinline__ double logistic_ref(double x)
{
    // LogExpFunctions.logistic for Float64: exp(x) / (1 + exp(x)), saturated outside (-744.44, 36.74)
    const double e = exp(x);
    return x < -744.4400719213812 ? 0.0 : (x > 36.7368005696771 ? 1.0 : e / (1.0 + e));
}

__device__ __forceinline__ double inverse_one(int kind, double lam, double offset, double max_value, double y)
{
    switch (kind) {
    case 1: return fmax(exp(y) - offset, 0.0);
    case 2: return fmax(logistic_ref(y) * 100.0 - offset, 0.0);
    case 3: {
        const double v = lam * y + 1.0;
        double res;
        if (lam > 0.0) res = pow(fmax(v, 1.0e-10), 1.0 / lam) - offset;
        else if (lam < 0.0) {
            if (v > 1.0e-10) res = pow(v, 1.0 / lam) - offset;
            else if (v <= 0.0) res = 0.0;
            else res = fmin(pow(v, 1.0 / lam), 1000.0 * max_value) - offset;
        } else res = exp(y) - offset;
        return fmax(res, 0.0);
    }
    default: return y;
    }
}

// x: [h, N] column-major (element (r, c) at x[r + h c]). out (nullable): same layout. rows (nullable): [h][N].
__global__ void __launch_bounds__(256) inverse_transform_kernel(int kind, double lam, double offset, double max_value,
                                                                int64_t h, int64_t N, const double *x, double *out,
                                                                double *rows)
{
    const int64_t total = h * N;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const double v = inverse_one(kind, lam, offset, max_value, x[e]);
        if (out) out[e] = v;
        if (rows) {
            const int64_t c = e / h, r = e - c * h;
            rows[r * N + c] = v;
        }
    }
}

__device__ __forceinline__ unsigned long long sort_key(double v)
{
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k)
{
    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

// rank-th smallest (0-based) of row[0..N): most-significant-digit radix select, 8 bits per pass
__device__ double select_rank(const double *row, int64_t N, int64_t rank, unsigned *hist, unsigned long long *s_state)
{
    unsigned long long prefix = 0, mask = 0;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        // latency-bound on the L2 reads of the row: four independent loads in flight per thread
        {
            const int64_t stride = blockDim.x;
            int64_t i = threadIdx.x;
            for (; i + 3 * stride < N; i += 4 * stride) {
                const unsigned long long k0 = sort_key(row[i]), k1 = sort_key(row[i + stride]);
                const unsigned long long k2 = sort_key(row[i + 2 * stride]), k3 = sort_key(row[i + 3 * stride]);
                if ((k0 & mask) == prefix) atomicAdd(&hist[(unsigned)(k0 >> shift) & 255u], 1u);
                if ((k1 & mask) == prefix) atomicAdd(&hist[(unsigned)(k1 >> shift) & 255u], 1u);
                if ((k2 & mask) == prefix) atomicAdd(&hist[(unsigned)(k2 >> shift) & 255u], 1u);
                if ((k3 & mask) == prefix) atomicAdd(&hist[(unsigned)(k3 >> shift) & 255u], 1u);
            }
            for (; i < N; i += stride) {
                const unsigned long long k0 = sort_key(row[i]);
                if ((k0 & mask) == prefix) atomicAdd(&hist[(unsigned)(k0 >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp 0: each lane sums 8 consecutive bins, exclusive scan over lanes, then the owning lane walks its 8
            const int lane = threadIdx.x;
            unsigned loc[8];
            unsigned long long tot = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { loc[j] = hist[lane * 8 + j]; tot += loc[j]; }
            unsigned long long incl = tot;
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            const unsigned long long excl = incl - tot;
            const bool mine = (unsigned long long)rank >= excl && (unsigned long long)rank < incl;
            const unsigned who = __ballot_sync(0xffffffffu, mine);
            // rank beyond the count can only come from an inconsistent histogram: clamp to the last non-empty lane
            const int owner = who ? __ffs(who) - 1 : 31;
            if (lane == owner) {
                unsigned long long cum = excl;
                int j = 0;
                for (; j < 7; ++j) {
                    if (cum + loc[j] > (unsigned long long)rank) break;
                    cum += loc[j];
                }
                s_state[0] = prefix | ((unsigned long long)(lane * 8 + j) << shift);
                s_state[1] = (unsigned long long)rank - cum;
            }
        }
        __syncthreads();
        prefix = s_state[0];
        rank = (int64_t)s_state[1];
        mask |= 0xffull << shift;
        __syncthreads();
    }
    return key_value(prefix);
}

// One CTA per (row, probability). Julia `quantile(v, p)` (Statistics, alpha = beta = 1):
//   aleph = n p + (1 - p); j = clamp(trunc(aleph), 1, n - 1); gamma = clamp(aleph - j, 0, 1);
//   q = v[j] + gamma (v[j+1] - v[j]) for finite neighbours, (1 - gamma) v[j] + gamma v[j+1] otherwise.
__global__ void __launch_bounds__(1024) row_quantile_kernel(int64_t h, int64_t N, const double *rows, int64_t nq,
                                                           const double *probs, double *q)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned long long s_state[2];
    const int64_t r = blockIdx.x / nq, jq = blockIdx.x % nq;
    const double *row = rows + r * N;
    const double p = probs[jq];
    double result;
    if (N == 1) {
        result = row[0];
    } else {
        const double nd = (double)N;
        const double aleph = nd * p + (1.0 - p);
        int64_t j = (int64_t)trunc(aleph);
        j = j < 1 ? 1 : (j > N - 1 ? N - 1 : j);
        double gamma = aleph - (double)j;
        gamma = gamma < 0.0 ? 0.0 : (gamma > 1.0 ? 1.0 : gamma);
        const double a = select_rank(row, N, j - 1, hist, s_state);
        const double b = select_rank(row, N, j, hist, s_state);
        result = (isfinite(a) && isfinite(b)) ? a + gamma * (b - a) : (1.0 - gamma) * a + gamma * b;
    }
    if (threadIdx.x == 0) q[r * nq + jq] = result;
}

}  // namespace

cudaError_t launch_inverse_transform(int kind, double lam, double offset, double max_value, int64_t h, int64_t N,
                                     const double *x, double *out, double *rows, int num_sms, cudaStream_t stream)
{
    const int64_t total = h * N;
    if (total == 0) return cudaSuccess;
    const int64_t want = (total + 255) / 256;
    const int grid = (int)(want < (int64_t)num_sms * 8 ? want : (int64_t)num_sms * 8);
    inverse_transform_kernel<<<grid, 256, 0, stream>>>(kind, lam, offset, max_value, h, N, x, out, rows);
    return cudaGetLastError();
}

cudaError_t launch_row_quantiles(int64_t h, int64_t N, const double *rows, int64_t nq, const double *probs, double *q,
                                 cudaStream_t stream)
{
    if (h * nq == 0) return cudaSuccess;
    row_quantile_kernel<<<(unsigned)(h * nq), 1024, 0, stream>>>(h, N, rows, nq, probs, q);
    return cudaGetLastError();
}

}  // namespace nagp

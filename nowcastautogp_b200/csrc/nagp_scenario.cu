// nagp_scenario.cu — per-scenario kernels: the O(k^2 + hk) add_data! tail against stored factors,
// weight normalisation / ESS / multinomial resampling, and mixture draws.
//
// Replaces AutoGP.add_data! + maybe_resample! + rand(MixtureModel, D) at
//   /root/reference/src/forecasting.jl:135, :138-141, :47
// Arithmetic contract: docs/KERNEL_SPEC.md §6-§7. The draw `x = mu_c + L_c zeta` uses the fixed
// fma order of §7, so it is bit-identical to the oracle given identical (mu, L, c, zeta).
#include "nagp_kernels.cuh"

namespace nagp {

namespace {

constexpr int kAppendThreads = 128;
constexpr int kMaxK = 16;     // nowcast points handled in registers per instance
constexpr int kDrawThreads = 128;

// One thread per (scenario, particle).
__global__ void __launch_bounds__(kAppendThreads) append_kernel(const AppendArgs a)
{
    const int64_t idx = (int64_t)blockIdx.x * kAppendThreads + threadIdx.x;
    if (idx >= a.K * a.P) return;
    const int64_t s = idx / a.P;
    const int p = (int)(idx % a.P);
    const int k = a.k, h = a.h, kh = k + h;
    const double *proj = a.proj + (int64_t)p * kh;
    const double *Lt = a.Ltail + (int64_t)p * kh * kh;
    const double *y2 = a.y2 + s * k;
    double z2[kMaxK];
    double logdet = 0.0, quad = 0.0;
#pragma unroll 1
    for (int r = 0; r < k; ++r) {
        double acc = y2[r] - proj[r];
        for (int c = 0; c < r; ++c) acc = fma(-Lt[r * kh + c], z2[c], acc);
        double d = Lt[r * kh + r];
        acc = acc / d;
        z2[r] = acc;
        logdet += log(d);
        quad = fma(acc, acc, quad);
    }
    const double log2pi = 1.8378770664093454835606594728112;
    const double dl = -0.5 * ((double)k * log2pi + 2.0 * logdet + quad);
    a.logw[idx] = (a.logw0 ? a.logw0[p] : 0.0) + dl;
    if (a.mu) {
        for (int i = 0; i < h; ++i) {
            double acc = proj[k + i];
            for (int c = 0; c < k; ++c) acc = fma(Lt[(k + i) * kh + c], z2[c], acc);
            a.mu[idx * h + i] = (acc - a.yb) / a.ya;
        }
    }
}

// One CTA per scenario. Dynamic shared memory: w[P], cw[P], parent[P] (int).
__global__ void __launch_bounds__(kDrawThreads) draw_kernel(const DrawArgs a)
{
    extern __shared__ double sm[];
    const int P = (int)a.P;
    double *w = sm;
    double *cw = w + P;
    int *parent = reinterpret_cast<int *>(cw + P);
    __shared__ double s_max, s_ess;
    __shared__ int s_resampled;

    const int64_t s = blockIdx.x;
    const int tid = threadIdx.x;
    const double *lw = a.logw + s * P;

    if (tid == 0) {
        double mx = -INFINITY;
        for (int p = 0; p < P; ++p) mx = lw[p] > mx ? lw[p] : mx;
        s_max = mx;
    }
    __syncthreads();
    for (int p = tid; p < P; p += kDrawThreads) w[p] = exp(lw[p] - s_max);
    __syncthreads();
    if (tid == 0) {
        // ascending-order sums: the order is part of the contract (KERNEL_SPEC §7)
        double sum = 0.0;
        for (int p = 0; p < P; ++p) sum += w[p];
        double s2 = 0.0, run = 0.0;
        for (int p = 0; p < P; ++p) {
            double v = w[p] / sum;
            w[p] = v;
            s2 += v * v;
            run += v;
            cw[p] = run;
        }
        s_ess = 1.0 / s2;
        s_resampled = (a.u_res != nullptr) && (s_ess < a.ess_thr * (double)P);
        if (a.ess_out) a.ess_out[s] = s_ess;
    }
    __syncthreads();
    if (a.w_out)
        for (int p = tid; p < P; p += kDrawThreads) a.w_out[s * P + p] = w[p];

    auto invcdf = [&](double u) {
        // smallest p with cw[p] > u (cw is a non-decreasing running sum), clamped to P-1
        int lo = 0, hi = P - 1;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (cw[mid] > u) hi = mid; else lo = mid + 1;
        }
        return lo;
    };

    if (s_resampled) {
        for (int p = tid; p < P; p += kDrawThreads) parent[p] = invcdf(a.u_res[s * P + p]);
        __syncthreads();
    }

    const int h = a.h;
    const int64_t D = a.D;
    for (int64_t e = tid; e < D * h; e += kDrawThreads) {
        const int64_t d = e / h;
        const int i = (int)(e - d * h);
        int c;
        if (a.comp && a.comp[s * D + d] >= 0) c = a.comp[s * D + d];
        else if (s_resampled) {
            int64_t slot = (int64_t)(a.u[s * D + d] * (double)P);
            if (slot >= P) slot = P - 1;
            c = parent[slot];
        } else c = invcdf(a.u[s * D + d]);
        if (i == 0 && a.comp_out) a.comp_out[s * D + d] = c;
        const double *mc = a.mu + s * a.mu_stride_k + (int64_t)c * h;
        const double *Lc = a.L + s * a.l_stride_k + (int64_t)c * h * h + (int64_t)i * h;
        const double *zz = a.zeta + (s * D + d) * h;
        double acc = mc[i];
        for (int j = 0; j <= i; ++j) acc = fma(Lc[j], zz[j], acc);
        a.x[(s * D + d) * h + i] = acc;
    }
}

}  // namespace

cudaError_t launch_append(const AppendArgs &a, cudaStream_t stream)
{
    if (a.k > kMaxK) return cudaErrorInvalidValue;
    int64_t total = a.K * a.P;
    unsigned grid = (unsigned)((total + kAppendThreads - 1) / kAppendThreads);
    append_kernel<<<grid, kAppendThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_draw(const DrawArgs &a, cudaStream_t stream)
{
    size_t smem = (size_t)a.P * (2 * sizeof(double) + sizeof(int));
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(draw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    draw_kernel<<<(unsigned)a.K, kDrawThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace nagp

// nagp_fused_v1.cu — fused Gram -> Cholesky -> forward solve -> logML / predictive moments,
// one CTA per (scenario, particle) instance, whole problem resident in shared memory.
//
// This is the straightforward column kernel (variant 1): packed lower-triangular storage, one
// thread per matrix row, left-looking column sweep; the observation vector rides along as an extra
// matrix row so the forward solve costs nothing. It covers q+1 <= 235 rows in 227 KB and is the
// on-device cross-check for the tile kernel (variant 2).
//
// Replaces, per instance, AutoGP's Gram + dpotrf + solves behind
//   /root/reference/src/forecasting.jl:133 (GPModel(dict)), :135 (add_data!), :46 (predict_mvn)
//   /root/reference/src/make_and_fit_model.jl:91 (fit_smc! likelihood evaluations)
// Arithmetic contract: docs/KERNEL_SPEC.md §3-§6.
#include "nagp_kernels.cuh"
#include "nagp_tree.cuh"

namespace nagp {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kThreads) fused_v1_kernel(const FusedArgs a)
{
    extern __shared__ double smem[];
    __shared__ TreeProgram tp;
    __shared__ double s_piv;
    __shared__ int s_info;
    __shared__ double s_red[4][kThreads / 32];

    const int tid = threadIdx.x;
    const int64_t b = blockIdx.x;
    const int64_t s = b / a.P;
    const int p = (int)(b % a.P);
    const int n = a.n, k = a.k, h = a.h, m = n + k, q = m + h;
    const bool have_y2 = (a.y2 != nullptr) || k == 0;
    const int ny = have_y2 ? m : n;           // rows of y available
    const int rows = q + 1;                   // row q carries y
    const int G = a.G;

    double *A = smem;                          // packed rows 0..q
    double *tt = A + tri(rows);
    double *th = tt + q;
    double *tab = th + MAX_THETA;
    double *sig = tab + a.ntab_cap * (G > 0 ? G : 0);
    int32_t *gg = reinterpret_cast<int32_t *>(sig + a.ncp_cap * q);

    const int64_t po = a.prog_off[p], plen = a.prog_off[p + 1] - po;
    const int64_t to = a.theta_off[p], ntheta = a.theta_off[p + 1] - to;
    const double *theta_g = a.theta + s * a.theta_stride_k + to;

    if (tid == 0) {
        s_info = 0;
        if (ntheta > MAX_THETA) tp.error = -3;
        else tree_compile(tp, a.prog + po, (int)plen, (int)ntheta, G > 0 ? a.ntab_cap : 0, a.ncp_cap);
    }
    for (int i = tid; i < q; i += kThreads) {
        tt[i] = a.t[i];
        if (a.g) gg[i] = a.g[i];
    }
    for (int i = tid; i < ntheta && i < MAX_THETA; i += kThreads) th[i] = theta_g[i];
    __syncthreads();
    if (tp.error) {
        if (tid == 0) {
            a.info[b] = tp.error;
            if (a.logml_n) a.logml_n[b] = nan("");
            if (a.logml_m) a.logml_m[b] = nan("");
            if (a.logw) a.logw[b] = nan("");
        }
        return;
    }
    const int ntab = tp.ntab, ncp = tp.ncp;

    // ---- lag tables and changepoint sigma tables ------------------------------------------------
    for (int e = tid; e < ntab * G; e += kThreads) {
        int id = e / G, lag = e - id * G;
        int s0 = tp.tab_src0[id], s1 = tp.tab_src1[id];
        tab[e] = tree_eval(tp.sop + s0, tp.sarg + s0, nullptr, s1 - s0, th, 0.0, 0.0,
                           (double)lag * a.step, 0, nullptr, 0, nullptr, 0, 0, 0);
    }
    for (int e = tid; e < ncp * q; e += kThreads) {
        int id = e / q, i = e - id * q;
        const double *cp = th + tp.cp_theta[id];
        sig[e] = 0.5 * (1.0 + tanh((tt[i] - cp[0]) / cp[1]));
    }
    __syncthreads();

    // ---- Gram: packed lower triangle + y row ----------------------------------------------------
    const double nz = a.noise[s * a.noise_stride_k + p];
    const double d_lo = nz + a.jitter;
    const double d_hi = (a.noise_pred >= 0.0 ? a.noise_pred : nz) + a.jitter;
    const int ntri = tri(q);
    for (int e = tid; e < ntri; e += kThreads) {
        int i = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
        while (tri(i + 1) <= e) ++i;
        while (tri(i) > e) --i;
        int j = e - tri(i);
        double delta;
        int lag = 0;
        if (a.g) {
            lag = gg[i] - gg[j];
            lag = lag < 0 ? -lag : lag;
            delta = (double)lag * a.step;
        } else {
            delta = fabs(tt[i] - tt[j]);
        }
        double v = tree_eval(tp.cop, tp.carg, tp.caux, tp.clen, th, tt[i], tt[j], delta, lag, tab, G,
                             sig, q, i, j);
        if (i == j) v += (i < m) ? d_lo : d_hi;
        A[e] = v;
    }
    {
        const double *y1 = a.y1 + b * a.y1_stride;
        double *yrow = A + tri(q);
        for (int j = tid; j < q; j += kThreads) {
            double v = 0.0;
            if (j < n) v = y1[j];
            else if (j < ny) v = a.y2 ? a.y2[s * k + (j - n)] : y1[j];
            yrow[j] = v;
        }
    }
    __syncthreads();

    // ---- left-looking column Cholesky; thread i owns row i (row q = y: columns < ny only) --------
    const int i = tid;
    const bool own = i < rows;
    const double *Ai = A + (own ? tri(i) : 0);
    for (int j = 0; j < q; ++j) {
        const double *Aj = A + tri(j);
        const bool active = own && i >= j && (i < q || j < ny);
        double sacc = 0.0;
        if (active) {
            sacc = Ai[j];
            for (int c = 0; c < j; ++c) sacc = fma(-Ai[c], Aj[c], sacc);
        }
        if (active && i == j) {
            if (!(sacc > 0.0)) { s_info = j + 1; sacc = nan(""); }
            else sacc = sqrt(sacc);
            s_piv = sacc;
            A[tri(i) + j] = sacc;
        }
        __syncthreads();   // pivot visible
        if (s_info) break;
        if (active && i > j) A[tri(i) + j] = sacc / s_piv;
        __syncthreads();   // column j final before row j+1 is read as the pivot row
    }
    __syncthreads();

    if (s_info) {
        if (tid == 0) {
            a.info[b] = s_info;
            if (a.logml_n) a.logml_n[b] = nan("");
            if (a.logml_m) a.logml_m[b] = nan("");
            if (a.logw) a.logw[b] = nan("");
        }
        return;
    }

    // ---- logML(n), logML(m) ----------------------------------------------------------------------
    const double *z = A + tri(q);
    double ld_n = 0, ld_m = 0, qd_n = 0, qd_m = 0;
    for (int r = tid; r < m; r += kThreads) {
        double l = log(A[tri(r) + r]);
        double zz = r < ny ? z[r] * z[r] : 0.0;
        ld_m += l; qd_m += zz;
        if (r < n) { ld_n += l; qd_n += zz; }
    }
    ld_n = warp_sum(ld_n); ld_m = warp_sum(ld_m); qd_n = warp_sum(qd_n); qd_m = warp_sum(qd_m);
    if ((tid & 31) == 0) {
        s_red[0][tid >> 5] = ld_n; s_red[1][tid >> 5] = ld_m;
        s_red[2][tid >> 5] = qd_n; s_red[3][tid >> 5] = qd_m;
    }
    __syncthreads();
    if (tid == 0) {
        double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
        for (int w = 0; w < kThreads / 32; ++w) { r0 += s_red[0][w]; r1 += s_red[1][w]; r2 += s_red[2][w]; r3 += s_red[3][w]; }
        const double log2pi = 1.8378770664093454835606594728112;
        double lmn = -0.5 * ((double)n * log2pi + 2.0 * r0 + r2);
        double lmm = have_y2 ? -0.5 * ((double)m * log2pi + 2.0 * r1 + r3) : nan("");
        if (a.logml_n) a.logml_n[b] = lmn;
        if (a.logml_m) a.logml_m[b] = lmm;
        if (a.logw) a.logw[b] = (a.logw0 ? a.logw0[p] : 0.0) + (lmm - lmn);
        a.info[b] = 0;
    }

    // ---- predictive moments / fast-path tail blocks ------------------------------------------------
    const int kh = k + h;
    if (a.mu && have_y2) {
        for (int r = tid; r < h; r += kThreads) {
            const double *row = A + tri(m + r);
            double acc = 0.0;
            for (int c = 0; c < m; ++c) acc = fma(row[c], z[c], acc);
            a.mu[b * h + r] = (acc - a.yb) / a.ya;
        }
    }
    if (a.L33) {
        for (int e = tid; e < h * h; e += kThreads) {
            int r = e / h, c = e - r * h;
            a.L33[b * h * h + e] = c <= r ? A[tri(m + r) + m + c] / a.ya : 0.0;
        }
    }
    if (a.proj) {
        for (int r = tid; r < kh; r += kThreads) {
            const double *row = A + tri(n + r);
            double acc = 0.0;
            for (int c = 0; c < n; ++c) acc = fma(row[c], z[c], acc);
            a.proj[b * kh + r] = acc;
        }
    }
    if (a.Ltail) {
        for (int e = tid; e < kh * kh; e += kThreads) {
            int r = e / kh, c = e - r * kh;
            a.Ltail[b * kh * kh + e] = c <= r ? A[tri(n + r) + n + c] : 0.0;
        }
    }
}

}  // namespace

size_t fused_smem_bytes_v1(int q, int G, int ntab_cap, int ncp_cap)
{
    size_t rows = (size_t)q + 1;
    size_t dbl = rows * (rows + 1) / 2 + (size_t)q + MAX_THETA + (size_t)ntab_cap * (G > 0 ? G : 0) +
                 (size_t)ncp_cap * q;
    return dbl * sizeof(double) + (size_t)q * sizeof(int32_t) + 16;
}

cudaError_t launch_fused_v1(const FusedArgs &a, cudaStream_t stream)
{
    const int q = a.n + a.k + a.h;
    if (q + 1 > kThreads) return cudaErrorInvalidValue;
    size_t smem = fused_smem_bytes_v1(q, a.G, a.ntab_cap, a.ncp_cap);
    cudaError_t e = cudaFuncSetAttribute(fused_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    fused_v1_kernel<<<(unsigned)a.B, kThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace nagp

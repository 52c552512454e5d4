// nagp_grad.cu — gradient of the log marginal likelihood with respect to every kernel hyperparameter and
// the observation noise, batched over (scenario, particle) instances.
//
// What it replaces: AutoGP's HMC on the hyperparameters differentiates the MVN log density of each particle
// through Gen on the CPU (mcmc_parameters!, /root/reference/src/forecasting.jl:148 and :65; fit_smc!'s
// n_hmc, /root/reference/src/make_and_fit_model.jl:91). Here one launch gives d logML / d theta for all B
// instances; the leapfrog integrator itself stays on the host with the proposals.
//
//   logML = -1/2 y^T K^-1 y - 1/2 log|K| - n/2 log 2 pi
//   d logML / d theta_j = sum_{a,b} W_ab dK_ab / d theta_j,   W = 1/2 (alpha alpha^T - K^-1),  alpha = K^-1 y
//
// One CTA per instance, consuming the factor L (tile-packed, operand layout) and z = L^-1 y that the tile
// kernel keeps when FusedArgs::Lkeep is set:
//   1. alpha = L^-T z by backward substitution (one warp, shuffle reductions);
//   2. S = K^-1 directly from L by the row recurrence  S_ij = (delta_ij / L_ii - sum_{k>i} L_ki S_kj) / L_ii,
//      i = n-1 .. 0, all j >= i in parallel (S symmetric, stored as a full n x n square in shared memory when it
//      fits: n <= 160, else in global scratch);
//   3. W in place over S;
//   4. reverse-mode differentiation of the kernel tree per matrix entry (forward pass over the post-order
//      program keeping node values, backward pass pushing the entry's weight to the leaves), per-thread
//      accumulators in local memory, block reduction at the end.
// First correct version of this row: plain FP64 FMA, no DMMA yet. Formulas: docs/KERNEL_SPEC.md §3, §8.
#include <algorithm>

#include "nagp_kernels.cuh"
#include "nagp_tree.cuh"

namespace nagp {

namespace {

constexpr int kGT = 256;

__device__ __forceinline__ int g_tri(int i) { return (i * (i + 1)) >> 1; }
__device__ __forceinline__ int g_op_idx(int r, int c) { return ((r * 4 + (c & 3)) << 1) + (c >> 2); }
__device__ __forceinline__ double g_warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct GradProgram {
    uint8_t op[MAX_PROG];
    int16_t arg[MAX_PROG];     // theta offset of the node's parameters
    int8_t left[MAX_PROG];     // index of the left child's root (binary nodes); the right child is i - 1
    int len, ntheta, error;
};

__device__ void grad_compile(GradProgram &gp, const uint8_t *prog, int len, int ntheta)
{
    gp.error = 0; gp.len = len; gp.ntheta = ntheta;
    if (len <= 0 || len > MAX_PROG || ntheta > MAX_THETA) { gp.error = -3; return; }
    int8_t size[MAX_PROG];
    int sp = 0, th = 0;
    int8_t stack[MAX_STACK];
    for (int i = 0; i < len; ++i) {
        const int op = prog[i];
        if (op < 1 || op > 8) { gp.error = -3; return; }
        gp.op[i] = (uint8_t)op; gp.arg[i] = (int16_t)th; gp.left[i] = -1;
        if (op <= OP_PERIODIC) {
            if (sp >= MAX_STACK) { gp.error = -3; return; }
            size[i] = 1; stack[sp++] = (int8_t)i;
        } else {
            if (sp < 2) { gp.error = -3; return; }
            const int r = stack[sp - 1], l = stack[sp - 2];
            gp.left[i] = (int8_t)l;
            size[i] = (int8_t)(size[r] + size[l] + 1);
            sp -= 2; stack[sp++] = (int8_t)i;
        }
        th += op_nparam(op);
    }
    if (sp != 1 || th != ntheta) gp.error = -3;
}

__global__ void __launch_bounds__(kGT, 1) grad_kernel(const GradArgs a, const int s_in_smem)
{
    extern __shared__ __align__(16) double gsm[];
    __shared__ GradProgram gp;
    __shared__ double s_theta[MAX_THETA];
    __shared__ double s_red[kGT / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, nt = (n + 7) / 8, Q = nt * 8;
    double *S = s_in_smem ? gsm : a.S + (size_t)blockIdx.x * n * n;
    double *alpha = s_in_smem ? gsm + (size_t)n * n : gsm;      // [Q]
    double *tt = alpha + Q;                                    // [Q]
    double *lcol = tt + Q;                                     // [Q] current column of L
    double *s_part = lcol + Q;                                 // [8][n] partial sums of a row step
    int *gg = reinterpret_cast<int *>(s_part + 8 * (size_t)n); // [Q]

    for (int i = tid; i < Q; i += kGT) {
        tt[i] = i < n ? a.t[i] : 0.0;
        gg[i] = (a.g && i < n) ? a.g[i] : 0;
    }

    for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x) {
        __syncthreads();
        const int64_t s = b / a.P;
        const int p = (int)(b % a.P);
        const int64_t po = a.prog_off[p], plen = a.prog_off[p + 1] - po;
        const int64_t to = a.theta_off[p], ntheta = a.theta_off[p + 1] - to;
        const int64_t ntot = a.theta_off[a.P];
        double *gout = a.grad_theta + s * ntot + to;
        if (a.info[b] != 0) {
            for (int j = tid; j < ntheta; j += kGT) gout[j] = nan("");
            if (tid == 0) a.grad_noise[b] = nan("");
            continue;
        }
        if (tid == 0) grad_compile(gp, a.prog + po, (int)plen, (int)ntheta);
        const double *theta_g = a.theta + s * a.theta_stride_k + to;
        for (int i = tid; i < ntheta && i < MAX_THETA; i += kGT) s_theta[i] = theta_g[i];
        const double *Lb = a.L + (size_t)b * ((size_t)g_tri(nt) * 64);
        auto Lel = [&](int i, int j) { return Lb[((size_t)g_tri(i >> 3) + (j >> 3)) * 64 + g_op_idx(i & 7, j & 7)]; };
        const double *zb = a.z + (size_t)b * Q;
        __syncthreads();
        if (gp.error) {
            for (int j = tid; j < ntheta; j += kGT) gout[j] = nan("");
            if (tid == 0) a.grad_noise[b] = nan("");
            continue;
        }

        // ---- 1. alpha = L^-T z -------------------------------------------------------------------------
        if (warp == 0) {
            for (int i = n - 1; i >= 0; --i) {
                double acc = 0.0;
                for (int k = i + 1 + lane; k < n; k += 32) acc = fma(Lel(k, i), alpha[k], acc);
                acc = g_warp_sum(acc);
                if (lane == 0) alpha[i] = (zb[i] - acc) / Lel(i, i);
                __syncwarp();
            }
        }
        // ---- 2. S = K^-1 by rows from the bottom ---------------------------------------------------------
        for (int i = n - 1; i >= 0; --i) {
            __syncthreads();
            for (int k = i + tid; k < n; k += kGT) lcol[k] = Lel(k, i);     // column i of L, read once per step
            __syncthreads();
            const double lii = lcol[i];
            // off-diagonal entries of row i first (they use rows > i only): `parts` threads share one j
            const int nj = n - 1 - i;
            int parts = 1;
            while (parts < 8 && nj * parts * 2 <= kGT) parts *= 2;
            const int jj = tid % (nj > 0 ? nj : 1), part = tid / (nj > 0 ? nj : 1);
            double acc = 0.0;
            if (nj > 0 && part < parts) {
                const int j = i + 1 + jj;
                for (int k = i + 1 + part; k < n; k += parts) acc = fma(lcol[k], S[(size_t)k * n + j], acc);
                if (parts > 1) s_part[part * n + jj] = acc;
            }
            if (parts > 1) {
                __syncthreads();
                if (nj > 0 && part == 0) {
                    acc = 0.0;
                    for (int q2 = 0; q2 < parts; ++q2) acc += s_part[q2 * n + jj];
                }
            }
            if (nj > 0 && part == 0) {
                const int j = i + 1 + jj;
                const double v = -acc / lii;
                S[(size_t)i * n + j] = v;
                S[(size_t)j * n + i] = v;
            }
            __syncthreads();
            // ... then the diagonal, which needs the column just written
            if (warp == 0) {
                double acc2 = 0.0;
                for (int k = i + 1 + lane; k < n; k += 32) acc2 = fma(lcol[k], S[(size_t)k * n + i], acc2);
                acc2 = g_warp_sum(acc2);
                if (lane == 0) S[(size_t)i * n + i] = (1.0 / lii - acc2) / lii;
            }
        }
        __syncthreads();
        // ---- 3 + 4. W = 1/2 (alpha alpha^T - S); reverse-mode through the tree per entry ------------------
        double gl[MAX_THETA];
        for (int j = 0; j < (int)ntheta; ++j) gl[j] = 0.0;
        double gnoise = 0.0;
        const int nent = g_tri(n - 1) + n;       // entries of the lower triangle
        for (int e = tid; e < nent; e += kGT) {
            int ia = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
            ia += (g_tri(ia + 1) <= e);
            ia -= (g_tri(ia) > e);
            const int ib = e - g_tri(ia);
            const double Wab = 0.5 * (alpha[ia] * alpha[ib] - S[(size_t)ia * n + ib]);
            const double w = ia == ib ? Wab : 2.0 * Wab;
            if (ia == ib) gnoise += Wab;
            const double ti = tt[ia], tj = tt[ib];
            double delta;
            if (a.g) { int lg = gg[ia] - gg[ib]; delta = (double)(lg < 0 ? -lg : lg) * a.step; }
            else delta = fabs(ti - tj);
            // forward
            double val[MAX_PROG], adj[MAX_PROG];
            const int len = gp.len;
            for (int i = 0; i < len; ++i) {
                const int op = gp.op[i];
                const double *th = s_theta + gp.arg[i];
                double v;
                switch (op) {
                case OP_CONSTANT: v = th[0]; break;
                case OP_LINEAR: v = fma(th[2], (ti - th[0]) * (tj - th[0]), th[1]); break;
                case OP_SQEXP: { double r = delta / th[0]; v = th[1] * exp(-0.5 * (r * r)); break; }
                case OP_GAMMAEXP: { double r = delta / th[0]; v = th[2] * exp(-pow(r, th[1])); break; }
                case OP_PERIODIC: {
                    double sn = sin(3.14159265358979323846 * (delta / th[1]));
                    v = th[2] * exp(-2.0 * (sn * sn) / (th[0] * th[0]));
                    break;
                }
                case OP_PLUS: v = val[gp.left[i]] + val[i - 1]; break;
                case OP_TIMES: v = val[gp.left[i]] * val[i - 1]; break;
                default: {   // OP_CHANGEPOINT
                    const double si = 0.5 * (1.0 + tanh((ti - th[0]) / th[1]));
                    const double sj = 0.5 * (1.0 + tanh((tj - th[0]) / th[1]));
                    v = ((1.0 - si) * (1.0 - sj)) * val[gp.left[i]] + (si * sj) * val[i - 1];
                    break;
                }
                }
                val[i] = v;
                adj[i] = 0.0;
            }
            // backward
            adj[len - 1] = w;
            for (int i = len - 1; i >= 0; --i) {
                const int op = gp.op[i];
                const double ad = adj[i];
                const double *th = s_theta + gp.arg[i];
                double *gth = gl + gp.arg[i];
                switch (op) {
                case OP_CONSTANT: gth[0] += ad; break;
                case OP_LINEAR: {
                    const double u = ti - th[0], v2 = tj - th[0];
                    gth[0] += ad * (-th[2] * (u + v2));
                    gth[1] += ad;
                    gth[2] += ad * (u * v2);
                    break;
                }
                case OP_SQEXP: {
                    const double r = delta / th[0];
                    gth[0] += ad * (val[i] * (r * r) / th[0]);
                    gth[1] += ad * (val[i] / th[1]);
                    break;
                }
                case OP_GAMMAEXP: {
                    const double r = delta / th[0];
                    const double rg = pow(r, th[1]);
                    gth[0] += ad * (val[i] * th[1] * rg / th[0]);
                    gth[1] += r > 0.0 ? ad * (-val[i] * rg * log(r)) : 0.0;
                    gth[2] += ad * (val[i] / th[2]);
                    break;
                }
                case OP_PERIODIC: {
                    const double ang = 3.14159265358979323846 * (delta / th[1]);
                    const double sn = sin(ang), cs = cos(ang);
                    const double l2 = th[0] * th[0];
                    gth[0] += ad * (val[i] * 4.0 * (sn * sn) / (l2 * th[0]));
                    gth[1] += ad * (val[i] * 4.0 * sn * cs * ang / (l2 * th[1]));
                    gth[2] += ad * (val[i] / th[2]);
                    break;
                }
                case OP_PLUS: adj[gp.left[i]] += ad; adj[i - 1] += ad; break;
                case OP_TIMES: adj[gp.left[i]] += ad * val[i - 1]; adj[i - 1] += ad * val[gp.left[i]]; break;
                default: {   // OP_CHANGEPOINT: (1-si)(1-sj) kL + si sj kR, si = sigma((ti - loc) / scale)
                    const double xi = (ti - th[0]) / th[1], xj = (tj - th[0]) / th[1];
                    const double si = 0.5 * (1.0 + tanh(xi)), sj = 0.5 * (1.0 + tanh(xj));
                    const double kl = val[gp.left[i]], kr = val[i - 1];
                    adj[gp.left[i]] += ad * ((1.0 - si) * (1.0 - sj));
                    adj[i - 1] += ad * (si * sj);
                    // d sigma / d x = 2 sigma (1 - sigma); dx / d loc = -1 / scale; dx / d scale = -x / scale
                    const double dsi = 2.0 * si * (1.0 - si), dsj = 2.0 * sj * (1.0 - sj);
                    const double dk_dsi = -(1.0 - sj) * kl + sj * kr;
                    const double dk_dsj = -(1.0 - si) * kl + si * kr;
                    gth[0] += ad * (dk_dsi * dsi + dk_dsj * dsj) * (-1.0 / th[1]);
                    gth[1] += ad * (dk_dsi * dsi * (-xi / th[1]) + dk_dsj * dsj * (-xj / th[1]));
                    break;
                }
                }
            }
        }
        // ---- 5. block reduction ----------------------------------------------------------------------------
        for (int j = 0; j <= (int)ntheta; ++j) {
            double v = g_warp_sum(j < (int)ntheta ? gl[j] : gnoise);
            __syncthreads();
            if (lane == 0) s_red[warp] = v;
            __syncthreads();
            if (tid == 0) {
                double r = 0.0;
                for (int w = 0; w < kGT / 32; ++w) r += s_red[w];
                if (j < (int)ntheta) gout[j] = r; else a.grad_noise[b] = r;
            }
        }
    }
}

}  // namespace

size_t grad_smem_bytes(int n, int smem_optin, bool *s_in_smem)
{
    const int Q = (n + 7) / 8 * 8;
    const size_t small = (size_t)Q * (8 + 8 + 8 + 4) + (size_t)8 * n * 8 + 64;
    const size_t full = (size_t)n * n * 8 + small;
    const size_t limit = (size_t)smem_optin - 8192;     // static shared memory (program, theta) + reservation
    *s_in_smem = full <= limit;
    return *s_in_smem ? full : small;
}

cudaError_t launch_grad(const GradArgs &a, int grid, size_t smem_bytes, cudaStream_t stream)
{
    const int s_in_smem = a.S == nullptr;
    cudaError_t e = cudaFuncSetAttribute(grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    grad_kernel<<<grid, kGT, smem_bytes, stream>>>(a, s_in_smem);
    return cudaGetLastError();
}

}  // namespace nagp

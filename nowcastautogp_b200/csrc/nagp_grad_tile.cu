// nagp_grad_tile.cu — gradient of the log marginal likelihood, tile (DMMA) version.
//
// What it replaces: AutoGP's HMC on the hyperparameters differentiates the MVN log density of each particle
// through Gen on the CPU (mcmc_parameters!, /root/reference/src/forecasting.jl:148 and :65; fit_smc!'s n_hmc,
// /root/reference/src/make_and_fit_model.jl:91). One launch gives d logML / d theta for all B instances.
//
//   d logML / d theta_j = sum_{a,b} W_ab dK_ab / d theta_j,   W = 1/2 (alpha alpha^T - K^-1),  alpha = K^-1 y
//
// One CTA per instance, two CTAs per SM at the vignette size. The CTA loads the factor L (8x8 tiles, operand
// layout, kept by the tile kernel together with the inverses of its diagonal tiles and z = L^-1 y) into shared
// memory and overwrites it IN PLACE with S = K^-1, tile column by tile column from the right:
//
//   T_J  = sum_{K>I} S_JK L_KI                (J > I; S symmetric, S_JK = S_KJ^T read transposed when K > J)
//   S_JI = -T_J L_II^-1
//   S_II = (L_II^-T - sum_{K>I} S_KI^T L_KI) L_II^-1
//   alpha_I^T = (z_I^T - sum_{K>I} alpha_K^T L_KI) L_II^-1      (alpha rides along as one more row)
//
// Column I of L is dead once column I of S exists, so no second matrix is needed. The factor is transposed tile by
// tile while it is loaded (tile (K, I) holds L_KI^T in operand layout), so every DMMA operand but the mirrored S_JK
// is one 16-byte LDS and no column has to be staged; two barriers per tile column (all reads of column I of L done /
// column I of S in place), the diagonal tile finished by the warp that needs it next. 2 q^3/3 FLOP, all on the FP64
// tensor pipe (mma.sync m8n8k4 f64: there is no tcgen05 kind for f64).
//
// The tree is then differentiated in reverse mode per matrix entry over the COMPILED program (stationary
// sub-trees folded into lag tables, ChangePoint sigmoids tabulated per point, as in the Gram pass of the tile
// kernel). A lane owns one lag d of a group of 32 and walks rows a with b = ginv[g_a - d]; the rows are dealt
// round-robin to the 8 warps. The adjoint that reaches a lag table is summed per (table, lag, warp) in a register
// — no atomics, fixed order, bit-reproducible — and only afterwards pushed through the stationary sub-tree once
// per (table, lag): G transcendental evaluations per table instead of n^2/2. Leaves that hang off the root through
// Plus nodes only (a sum of components is the typical AutoGP posterior structure) never reach the interpreter: their
// adjoint is the entry's weight w itself, so a lag table just receives the lag sums of w and a Linear / Constant leaf
// is differentiated from three weighted moments (sum w, sum w (ti + tj), sum w ti tj about the middle of the series).
// What is left after that, if at most nine compiled ops, is swept with every value and adjoint in a register
// (rev_entry_reg); longer programs go through the local-memory interpreter (rev_entry).
// Formulas: docs/KERNEL_SPEC.md §3, §8.
#include <algorithm>

#include "nagp_kernels.cuh"
#include "nagp_tree.cuh"
#include "nagp_tile.cuh"

#ifndef NAGP_GRAD_LOAD
#define NAGP_GRAD_LOAD 1        // 16-byte loads batched per thread while the factor streams in (measured: 1 17.3 ms, 2-4 18.0, 8 18.5
                                // per 32 000 chains - the other CTA of the SM hides this phase, and larger batches cost registers)
#endif
#ifndef NAGP_GRAD_LD
#define NAGP_GRAD_LD(p) (*(p))
#endif
#ifndef NAGP_GRAD_TRACE
#define NAGP_GRAD_TRACE 0   // n: clock64 stamps of the n-th instance of block 0 for tools/grad_timeline.py; 0 in product builds
#endif
#if NAGP_GRAD_TRACE
__device__ long long g_grad_trace[32 * 8 * 8 + 16];      // [tile column][warp][stamp], then per-phase stamps
extern "C" int nagp_debug_read_grad(long long *out, int count)
{
    return (int)cudaMemcpyFromSymbol(out, g_grad_trace, sizeof(long long) * count);
}
#define GTR(I, ph) do { if (tracing && lane == 0) g_grad_trace[((I) * 8 + warp) * 8 + (ph)] = clock64(); } while (0)
#define GTP(i) do { if (tracing && tid == 0) g_grad_trace[32 * 8 * 8 + (i)] = clock64(); } while (0)
#else
#define GTR(I, ph) do { } while (0)
#define GTP(i) do { } while (0)
#endif

namespace nagp {

namespace {

constexpr int kGW = 8;              // warps per CTA
constexpr int kGT2 = kGW * 32;
constexpr int kMaxRows = 4;         // work slots per warp and tile column (covers 32 tile rows; shared memory stops at 28)

struct RevProgram {
    int8_t cleft[MAX_PROG];    // compiled program: root index of the left child of a binary node
    int8_t sleft[MAX_PROG];    // source program: same
    int8_t slot[MAX_PROG];     // register-resident sweep: accumulator slot of a Linear / Constant leaf or a ChangePoint
    int reg_ok;                // the compiled program qualifies for the register-resident sweep
    // leaves that hang off the root through Plus nodes only (their adjoint is the entry's weight itself) are taken out
    // of the per-entry program: lag tables get the lag sums of the weights, Linear / Constant the weighted moments
    uint32_t peel_tab;                  // bit id: lag table id is such a leaf
    int npeel_lin, npeel_con;
    int16_t peel_lin[MAX_PROG / 2 + 1], peel_con[MAX_PROG / 2 + 1];   // theta offsets
};

#ifndef NAGP_GRAD_REG
#define NAGP_GRAD_REG 1
#endif
constexpr int kRegOps = 9;     // compiled ops the register-resident sweep holds (values and adjoints in registers)
constexpr int kRegLeaf = 4;    // Linear / Constant leaves with named accumulators
constexpr int kRegCp = 2;      // tabulated ChangePoints with named accumulators
constexpr int kRegTab = 5;     // lag tables (at most five leaves among nine ops)

// Thread-0 only: accumulator slots of the register-resident sweep, or reg_ok = 0 when the program does not fit it.
__device__ void reg_slots(const TreeProgram &tp, RevProgram &rp)
{
    int nl = 0, nc = 0;
    bool ok = tp.clen <= kRegOps && tp.ntab <= kRegTab;
    for (int i = 0; ok && i < tp.clen; ++i) {
        const int o = tp.cop[i];
        rp.slot[i] = 0;
        if (o == OP_LINEAR || o == OP_CONSTANT) { if (nl < kRegLeaf) rp.slot[i] = (int8_t)nl++; else ok = false; }
        else if (o == OP_CHANGEPOINT_TAB) { if (nc < kRegCp) rp.slot[i] = (int8_t)nc++; else ok = false; }
        else if (o != OP_TABLE && o != OP_PLUS && o != OP_TIMES) ok = false;
    }
    rp.reg_ok = ok && NAGP_GRAD_REG;
}

// Reverse-mode sweep of a short compiled program for one pair with every value and adjoint in a register: the ops
// loop is unrolled so positions are compile-time, and the only run-time indices (left child, accumulator slot, table
// id) are warp-uniform and resolved by select chains. Same arithmetic and order as rev_entry.
struct RegAcc {
    double la[kRegLeaf][3];
    double cp[kRegCp][2];
};
__device__ __forceinline__ void rev_entry_reg(const TreeProgram &tp, const RevProgram &rp, const double *th, double ti,
                                              double tj, int lag, int pi, int pj, const double *tab, int G,
                                              const double *sig, int Q, double w, RegAcc &acc, double (&hacc)[kRegTab])
{
    const int clen = tp.clen;
    double val[kRegOps], adj[kRegOps];
#pragma unroll
    for (int i = 0; i < kRegOps; ++i) {
        double v = 0.0;
        if (i < clen) {
            const uint32_t cw = tp.cword[i];
            const int o = cw & 0xff, arg = (cw >> 8) & 0xffff;
            if (o == OP_TABLE) v = tab[arg * G + lag];
            else if (o == OP_LINEAR) v = fma(th[arg + 2], (ti - th[arg]) * (tj - th[arg]), th[arg + 1]);
            else if (o == OP_CONSTANT) v = th[arg];
            else if (i >= 2) {
                const int l = rp.cleft[i];
                double vl = val[0];
#pragma unroll
                for (int j = 1; j < i - 1; ++j) vl = (l == j) ? val[j] : vl;
                const double vr = val[i - 1];
                if (o == OP_PLUS) v = vl + vr;
                else if (o == OP_TIMES) v = vl * vr;
                else {
                    const int ax = cw >> 24;
                    const double si = sig[ax * Q + pi], sj = sig[ax * Q + pj];
                    v = ((1.0 - si) * (1.0 - sj)) * vl + (si * sj) * vr;
                }
            }
        }
        val[i] = v;
        adj[i] = (i == clen - 1) ? w : 0.0;
    }
#pragma unroll
    for (int i = kRegOps - 1; i >= 0; --i) {
        if (i < clen) {
            const uint32_t cw = tp.cword[i];
            const int o = cw & 0xff, arg = (cw >> 8) & 0xffff;
            const double ad = adj[i];
            if (o == OP_TABLE) {
#pragma unroll
                for (int j = 0; j < kRegTab; ++j) hacc[j] += (arg == j) ? ad : 0.0;
            } else if (o == OP_LINEAR || o == OP_CONSTANT) {
                double g0, g1 = 0.0, g2 = 0.0;
                if (o == OP_LINEAR) {
                    const double u = ti - th[arg], v2 = tj - th[arg];
                    g0 = ad * (-th[arg + 2] * (u + v2));
                    g1 = ad;
                    g2 = ad * (u * v2);
                } else g0 = ad;
                const int sl = rp.slot[i];
#pragma unroll
                for (int q = 0; q < kRegLeaf; ++q)
                    if (sl == q) { acc.la[q][0] += g0; acc.la[q][1] += g1; acc.la[q][2] += g2; }
            } else if (i >= 2) {
                const int l = rp.cleft[i];
                double kl = val[0];
#pragma unroll
                for (int j = 1; j < i - 1; ++j) kl = (l == j) ? val[j] : kl;
                const double kr = val[i - 1];
                double al, ar;
                if (o == OP_PLUS) { al = ad; ar = ad; }
                else if (o == OP_TIMES) { al = ad * kr; ar = ad * kl; }
                else {
                    const int ax = cw >> 24;
                    const double si = sig[ax * Q + pi], sj = sig[ax * Q + pj];
                    const double xi = (ti - th[arg]) / th[arg + 1], xj = (tj - th[arg]) / th[arg + 1];
                    al = ad * ((1.0 - si) * (1.0 - sj));
                    ar = ad * (si * sj);
                    const double dsi = 2.0 * si * (1.0 - si), dsj = 2.0 * sj * (1.0 - sj);
                    const double dk_dsi = -(1.0 - sj) * kl + sj * kr;
                    const double dk_dsj = -(1.0 - si) * kl + si * kr;
                    const double c0 = ad * (dk_dsi * dsi + dk_dsj * dsj) * (-1.0 / th[arg + 1]);
                    const double c1 = ad * (dk_dsi * dsi * (-xi / th[arg + 1]) + dk_dsj * dsj * (-xj / th[arg + 1]));
                    const int sl = rp.slot[i];
#pragma unroll
                    for (int q = 0; q < kRegCp; ++q)
                        if (sl == q) { acc.cp[q][0] += c0; acc.cp[q][1] += c1; }
                }
                adj[i - 1] += ar;
#pragma unroll
                for (int j = 0; j < i - 1; ++j) adj[j] += (l == j) ? al : 0.0;
            }
        }
    }
}

// left-child roots of a post-order program (leaf: opcode <= OP_PERIODIC or OP_TABLE)
__device__ void left_roots(const uint8_t *op, int len, int8_t *left)
{
    int8_t stack[MAX_STACK];
    int sp = 0;
    for (int i = 0; i < len; ++i) {
        const int o = op[i];
        left[i] = -1;
        if (o <= OP_PERIODIC || o == OP_TABLE) { if (sp < MAX_STACK) stack[sp++] = (int8_t)i; }
        else if (sp >= 2) { left[i] = stack[sp - 2]; sp -= 2; stack[sp++] = (int8_t)i; }
    }
}

// Thread-0 only. Splits the compiled program at its root: the program is a sum of terms (the operands of the Plus
// nodes reachable from the root through Plus nodes); a term that is a single lag table, Linear or Constant leaf is
// recorded in rp.peel_* and removed, the other terms are re-joined by Plus nodes into the program the per-entry
// sweep interprets (tp.clen == 0 when nothing is left). Then left-child roots and register slots of what is left.
__device__ void peel_root(TreeProgram &tp, RevProgram &rp)
{
    const int len = tp.clen;
    uint8_t op[MAX_PROG]; int16_t arg[MAX_PROG]; int8_t aux[MAX_PROG], size[MAX_PROG], st[MAX_PROG];
    left_roots(tp.cop, len, rp.cleft);
    for (int i = 0; i < len; ++i) {
        op[i] = tp.cop[i]; arg[i] = tp.carg[i]; aux[i] = tp.caux[i];
        size[i] = rp.cleft[i] < 0 ? 1 : (int8_t)(1 + size[i - 1] + size[rp.cleft[i]]);
    }
    rp.peel_tab = 0; rp.npeel_lin = 0; rp.npeel_con = 0;
    unsigned long long roots = 0;
    int sp = 0;
    if (len > 0) st[sp++] = (int8_t)(len - 1);
    while (sp > 0) {
        const int i = st[--sp], o = op[i];
        if (o == OP_PLUS) { st[sp++] = rp.cleft[i]; st[sp++] = (int8_t)(i - 1); }
        else if (o == OP_TABLE) rp.peel_tab |= 1u << arg[i];
        else if (o == OP_LINEAR) rp.peel_lin[rp.npeel_lin++] = arg[i];
        else if (o == OP_CONSTANT) rp.peel_con[rp.npeel_con++] = arg[i];
        else roots |= 1ull << i;
    }
    int nout = 0, nterms = 0;
    auto emit = [&](int o, int ar, int ax) {
        tp.cop[nout] = (uint8_t)o; tp.carg[nout] = (int16_t)ar; tp.caux[nout] = (int8_t)ax;
        tp.cword[nout] = (uint32_t)o | ((uint32_t)(uint16_t)ar << 8) | ((uint32_t)(uint8_t)ax << 24);
        ++nout;
    };
    for (int i = 0; i < len; ++i) {
        if (!((roots >> i) & 1ull)) continue;
        for (int j = i - size[i] + 1; j <= i; ++j) emit(op[j], arg[j], aux[j]);
        if (++nterms > 1) emit(OP_PLUS, 0, 0);
    }
    tp.clen = nout;
    left_roots(tp.cop, nout, rp.cleft);
    reg_slots(tp, rp);
}

// Reverse-mode sweep of ops [i0, i1) for one pair: forward values, then the adjoint `w` of the root pushed to
// the leaves. Parameter gradients go to gl[theta slot]; the adjoint reaching lag table `id` goes to hacc[id].
__device__ __forceinline__ void rev_entry(const uint8_t *op, const int16_t *arg, const int8_t *aux, const int8_t *left,
                                          int i0, int i1, const double *th, double ti, double tj, double delta,
                                          int lag, int pi, int pj, const double *tab, int G, const double *sig, int Q,
                                          double w, double *gl, double *hacc)
{
    double val[MAX_PROG], adj[MAX_PROG];
    for (int i = i0; i < i1; ++i) {
        const int o = op[i];
        const double *p = th + arg[i];
        double v;
        switch (o) {
        case OP_CONSTANT: v = p[0]; break;
        case OP_LINEAR: v = fma(p[2], (ti - p[0]) * (tj - p[0]), p[1]); break;
        case OP_SQEXP: { double r = delta / p[0]; v = p[1] * exp(-0.5 * (r * r)); break; }
        case OP_GAMMAEXP: { double r = delta / p[0]; v = p[2] * exp(-pow(r, p[1])); break; }
        case OP_PERIODIC: {
            double sn = sin(3.14159265358979323846 * (delta / p[1]));
            v = p[2] * exp(-2.0 * (sn * sn) / (p[0] * p[0]));
            break;
        }
        case OP_TABLE: v = tab[arg[i] * G + lag]; break;
        case OP_PLUS: v = val[left[i]] + val[i - 1]; break;
        case OP_TIMES: v = val[left[i]] * val[i - 1]; break;
        case OP_CHANGEPOINT_TAB: {
            const double si = sig[aux[i] * Q + pi], sj = sig[aux[i] * Q + pj];
            v = ((1.0 - si) * (1.0 - sj)) * val[left[i]] + (si * sj) * val[i - 1];
            break;
        }
        default: {   // OP_CHANGEPOINT
            const double si = 0.5 * (1.0 + tanh((ti - p[0]) / p[1]));
            const double sj = 0.5 * (1.0 + tanh((tj - p[0]) / p[1]));
            v = ((1.0 - si) * (1.0 - sj)) * val[left[i]] + (si * sj) * val[i - 1];
            break;
        }
        }
        val[i] = v;
        adj[i] = 0.0;
    }
    adj[i1 - 1] = w;
    for (int i = i1 - 1; i >= i0; --i) {
        const int o = op[i];
        const double ad = adj[i];
        const double *p = th + arg[i];
        double *gp = gl + arg[i];
        switch (o) {
        case OP_CONSTANT: gp[0] += ad; break;
        case OP_LINEAR: {
            const double u = ti - p[0], v2 = tj - p[0];
            gp[0] += ad * (-p[2] * (u + v2));
            gp[1] += ad;
            gp[2] += ad * (u * v2);
            break;
        }
        case OP_SQEXP: {
            const double r = delta / p[0];
            gp[0] += ad * (val[i] * (r * r) / p[0]);
            gp[1] += ad * (val[i] / p[1]);
            break;
        }
        case OP_GAMMAEXP: {
            const double r = delta / p[0];
            const double rg = pow(r, p[1]);
            gp[0] += ad * (val[i] * p[1] * rg / p[0]);
            gp[1] += r > 0.0 ? ad * (-val[i] * rg * log(r)) : 0.0;
            gp[2] += ad * (val[i] / p[2]);
            break;
        }
        case OP_PERIODIC: {
            const double ang = 3.14159265358979323846 * (delta / p[1]);
            const double sn = sin(ang), cs = cos(ang);
            const double l2 = p[0] * p[0];
            gp[0] += ad * (val[i] * 4.0 * (sn * sn) / (l2 * p[0]));
            gp[1] += ad * (val[i] * 4.0 * sn * cs * ang / (l2 * p[1]));
            gp[2] += ad * (val[i] / p[2]);
            break;
        }
        case OP_TABLE: hacc[arg[i]] += ad; break;
        case OP_PLUS: adj[left[i]] += ad; adj[i - 1] += ad; break;
        case OP_TIMES: adj[left[i]] += ad * val[i - 1]; adj[i - 1] += ad * val[left[i]]; break;
        default: {   // ChangePoint (direct or tabulated): (1-si)(1-sj) kL + si sj kR, si = sigma((ti - loc) / scale)
            const double xi = (ti - p[0]) / p[1], xj = (tj - p[0]) / p[1];
            double si, sj;
            if (o == OP_CHANGEPOINT_TAB) { si = sig[aux[i] * Q + pi]; sj = sig[aux[i] * Q + pj]; }
            else { si = 0.5 * (1.0 + tanh(xi)); sj = 0.5 * (1.0 + tanh(xj)); }
            const double kl = val[left[i]], kr = val[i - 1];
            adj[left[i]] += ad * ((1.0 - si) * (1.0 - sj));
            adj[i - 1] += ad * (si * sj);
            // d sigma / d x = 2 sigma (1 - sigma); dx / d loc = -1 / scale; dx / d scale = -x / scale
            const double dsi = 2.0 * si * (1.0 - si), dsj = 2.0 * sj * (1.0 - sj);
            const double dk_dsi = -(1.0 - sj) * kl + sj * kr;
            const double dk_dsj = -(1.0 - si) * kl + si * kr;
            gp[0] += ad * (dk_dsi * dsi + dk_dsj * dsj) * (-1.0 / p[1]);
            gp[1] += ad * (dk_dsi * dsi * (-xi / p[1]) + dk_dsj * dsj * (-xj / p[1]));
            break;
        }
        }
    }
}

struct GradTileLayout {
    int nt;
    int region_bytes;         // union region: {part} during the sweep, {tt, theta, gg, ginv} afterwards
    int Gd;                   // lags handled (lag-grid extent, or n when times are pairwise)
    int scratch_stride;       // bytes of global scratch per CTA (lag tables, sigma tables, per-warp table adjoints)
    char *scratch;
    unsigned long long *work_counter;
};

__global__ void __launch_bounds__(kGT2, 2) grad_tile_kernel(const GradArgs a, const GradTileLayout lay)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ TreeProgram tp;
    __shared__ RevProgram rp;
    __shared__ long long s_next;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, nt = lay.nt, Q = nt * 8, ntiles = tri(nt);
    const int G = a.G, Gd = lay.Gd;

    double *tiles = smem;
    double *alpha = tiles + (ntiles + nt) * 64; // [Q] z during the sweep (tile row nt of `tiles` is alpha's row), alpha after it
    double *invs = alpha + Q;                   // [2][64] inverses of the diagonal tiles of L, current and next column
    double *region = invs + 128;
    // sweep view of the region
    double *part = region;                      // [kGW][64] partial sums of the diagonal tile, accumulator layout
    // differentiation view of the region
    double *tt = region;                        // [Q]
    double *th = tt + Q;                        // [MAX_THETA]
    int *gg = reinterpret_cast<int *>(th + MAX_THETA);           // [Q]
    short *ginv = reinterpret_cast<short *>(gg + Q);             // [Gd]
    double *tab = reinterpret_cast<double *>(lay.scratch + (size_t)blockIdx.x * lay.scratch_stride);   // [ntab_cap][G]
    double *sig = tab + a.ntab_cap * (G > 0 ? G : 0);                                                    // [ncp_cap][Q]
    double *hpart = sig + a.ncp_cap * Q;         // [kGW][ntab_cap][32 ceil(Gd/32)] per-warp partial table adjoints

    const int lr = lane >> 2, lj = lane & 3;
    const int oi0 = op_idx(lr, 2 * lj), oi1 = op_idx(lr, 2 * lj + 1);
    const int tr0 = op_idx(lj, lr);                       // transposed fragment: element (lj, lr); chunk 1 is +32
    const int cv0 = (lane & ~3) + (lj >> 1), cv1 = cv0 + 2;
    const int ts0 = lj * 4 + (lr >> 1);                   // accumulator layout -> A operand of the transposed tile
    const bool odd = lane & 1;
    const uint32_t tiles_a = smem_addr(tiles), invs_a = smem_addr(invs);

    // C (accumulator layout) times invL, result in accumulator layout: X * invL = X * (invL^T)^T
    auto times_inv = [&](uint32_t inv_a, double c0, double c1, double &x0, double &x1) {
        const double ibx = lds64(inv_a + tr0 * 8), iby = lds64(inv_a + tr0 * 8 + 256);
        const double v00 = shfl(c0, cv0), v01 = shfl(c1, cv0);
        const double v10 = shfl(c0, cv1), v11 = shfl(c1, cv1);
        x0 = 0.0; x1 = 0.0;
        dmma(x0, x1, odd ? v01 : v00, ibx);
        dmma(x0, x1, odd ? v11 : v10, iby);
    };

#if NAGP_GRAD_TRACE
    int n_inst = 0;
#endif
    for (;;) {
        __syncthreads();
        if (tid == 0) s_next = (long long)atomicAdd(lay.work_counter, 1ull);
        __syncthreads();
        const int64_t b = s_next;
        if (b >= a.B) break;
#if NAGP_GRAD_TRACE
        const bool tracing = (blockIdx.x == 0 && ++n_inst == NAGP_GRAD_TRACE);
        if (tracing && tid == 0) g_grad_trace[32 * 8 * 8 + 15] = b % a.P;
#endif
        GTP(0);
        const int64_t s = b / a.P;
        const int p = (int)(b % a.P);
        const int64_t po = a.prog_off[p], plen = a.prog_off[p + 1] - po;
        const int64_t to = a.theta_off[p], ntheta = a.theta_off[p + 1] - to;
        const int64_t ntot = a.theta_off[a.P];
        double *gout = a.grad_theta + s * ntot + to;
        const double *theta_g = a.theta + s * a.theta_stride_k + to;
        if (a.info[b] != 0) {
            for (int j = tid; j < ntheta; j += kGT2) gout[j] = nan("");
            if (tid == 0) a.grad_noise[b] = nan("");
            continue;
        }
        // ---- 0. program, factor -----------------------------------------------------------------------------
        if (a.compiled) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(a.compiled + p);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&tp);
            for (int i = tid; i < (int)(sizeof(TreeProgram) / 4); i += kGT2) dst[i] = src[i];
        } else if (tid == 0) {
            if (ntheta > MAX_THETA) tp.error = -3;
            else tree_compile(tp, a.prog + po, (int)plen, (int)ntheta, G > 0 ? a.ntab_cap : 0, a.ncp_cap);
        }
        {   // factor into shared memory, every tile transposed: tile (K, I) then holds L_KI^T in operand layout, which is
            // the B operand of every product of the sweep (one 16-byte LDS), and no column has to be staged
            const double2 *Lg = reinterpret_cast<const double2 *>(a.L + (size_t)b * ((size_t)ntiles * 64));
            // several 16-byte loads in flight per thread: the factor is ~100 KB straight from HBM
            const int nld = ntiles * 32;
            for (int i0 = tid; i0 < nld; i0 += kGT2 * NAGP_GRAD_LOAD) {
                double2 v[NAGP_GRAD_LOAD];
#pragma unroll
                for (int q = 0; q < NAGP_GRAD_LOAD; ++q)
                    if (i0 + q * kGT2 < nld) v[q] = NAGP_GRAD_LD(Lg + i0 + q * kGT2);
#pragma unroll
                for (int q = 0; q < NAGP_GRAD_LOAD; ++q) {
                    const int i = i0 + q * kGT2;
                    if (i < nld) {
                        const int ln = i & 31, r = ln >> 2, c = ln & 3;
                        double *dst = tiles + (i >> 5) * 64 + op_idx(c, r);
                        dst[0] = v[q].x;         // element (r, c) -> (c, r)
                        dst[32] = v[q].y;        // element (r, c + 4) -> (c + 4, r)
                    }
                }
            }
            // alpha rides along as tile row nt (its first row; the other seven stay zero), z waits in shared memory
            for (int i = tid; i < nt * 64; i += kGT2) tiles[ntiles * 64 + i] = 0.0;
            const double *zg = a.z + (size_t)b * Q;
            for (int i = tid; i < Q; i += kGT2) alpha[i] = zg[i];
        }
        __syncthreads();
        if (tp.error) {
            for (int j = tid; j < ntheta; j += kGT2) gout[j] = nan("");
            if (tid == 0) a.grad_noise[b] = nan("");
            continue;
        }
        if (tid == 0) peel_root(tp, rp);
        if (tid == 32) left_roots(tp.sop, tp.slen, rp.sleft);
        const double *Wb = a.Winv + (size_t)b * ((size_t)nt * 64);

        // ---- 1. S = K^-1 in place, alpha = K^-1 y ------------------------------------------------------------
        // Two barriers per tile column. A: every product that reads column I of L is done (results wait in registers),
        // so the column may be overwritten by S. B: column I of S and the partial sums of its diagonal tile are in place.
        // The diagonal tile S_II is finished after B by the warp that owns row I, which needs it only for that row (its
        // last of the next column); the other warps go straight on with rows that do not touch it.
        GTP(1);
        if (tid < 64) invs[((nt - 1) & 1) * 64 + tid] = Wb[(nt - 1) * 64 + tid];
        auto finish_diag = [&](int D) {
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int w = 0; w < kGW; ++w) {
                const double2 v = *reinterpret_cast<const double2 *>(part + w * 64 + lane * 2);
                t0 += v.x; t1 += v.y;
            }
            // invL^T in accumulator layout: element (lr, c) = invL[c][lr]
            const double *iv = invs + (D & 1) * 64;
            const double r0 = iv[op_idx(2 * lj, lr)] - t0, r1 = iv[op_idx(2 * lj + 1, lr)] - t1;
            double x0, x1;
            times_inv(invs_a + (uint32_t)((D & 1) * 512), r0, r1, x0, x1);
            const uint32_t dt = tiles_a + (uint32_t)((tri(D) + D) * 512);
            sts64(dt + oi0 * 8, x0);
            sts64(dt + oi1 * 8, x1);
            __syncwarp();
        };
        for (int I = nt - 1; I >= 0; --I) {
            GTR(I, 0);
            __syncthreads();                                   // B of column I+1 (first pass: the factor is loaded)
            GTR(I, 1);
            const uint32_t inv_cur = invs_a + (uint32_t)((I & 1) * 512);
            double wnext = 0.0;
            if (tid < 64 && I > 0) wnext = Wb[(I - 1) * 64 + tid];
            // Work of column I, dealt afresh: the warp that finishes the diagonal tile of column I+1 takes row I+1 (the
            // only row that needs that tile) as its last slot; alpha and the rows below are dealt round-robin to the
            // other seven warps first, so the serial piece sits on the least loaded warp.
            const int r = nt - 1 - I, w0 = (I + 1) % kGW, pos = (warp - w0 - 1) & (kGW - 1);
            auto row_of = [&](int u) -> int {
                if (u == kMaxRows - 1 && warp == w0) return r >= 1 ? I + 1 : -1;
                const int e = pos + kGW * u;                    // e = 0 is alpha's row nt, e >= 1 is row nt - e
                return (e == 0 || e <= r - 1) ? nt - e : -1;
            };
            if (I + 1 < nt && warp == w0) finish_diag(I + 1);
            GTR(I, 2);
            double pd[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            double xs[kMaxRows][2];
#pragma unroll
            for (int u = 0; u < kMaxRows; ++u) {
                const int J = row_of(u);
                if (J >= 0) {
                    // four independent accumulator pairs (two tile products in flight): the dependent DMMA chain of a
                    // row is what a warp waits on, and the FP64 tensor pipe is rarely full
                    double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
                    const bool is_alpha = (J == nt);              // alpha_I^T = -(sum_K alpha_K^T L_KI - z_I^T) L_II^-1
                    const int Jd = is_alpha ? nt - 1 : J;         // last directly stored tile of the row
                    if (is_alpha && lr == 0) { acc[0][0] = -alpha[I * 8 + 2 * lj]; acc[0][1] = -alpha[I * 8 + 2 * lj + 1]; }
                    const uint32_t rowa = tiles_a + (uint32_t)(tri(J) * 512 + lane * 16);
                    uint32_t cb = tiles_a + (uint32_t)((tri(I + 1) + I) * 512 + lane * 16);   // tile (K, I), K = I + 1
                    int K = I + 1;
                    for (; K + 1 <= Jd; K += 2) {
                        const double2 af0 = lds128(rowa + (uint32_t)K * 512u);
                        const double2 bf0 = lds128(cb);
                        const double2 af1 = lds128(rowa + (uint32_t)K * 512u + 512u);
                        const double2 bf1 = lds128(cb + (uint32_t)(K + 1) * 512u);
                        cb += (uint32_t)(2 * K + 3) * 512u;
                        dmma(acc[0][0], acc[0][1], af0.x, bf0.x);
                        dmma(acc[1][0], acc[1][1], af0.y, bf0.y);
                        dmma(acc[2][0], acc[2][1], af1.x, bf1.x);
                        dmma(acc[3][0], acc[3][1], af1.y, bf1.y);
                    }
                    if (K <= Jd) {
                        const double2 af = lds128(rowa + (uint32_t)K * 512u);
                        const double2 bf = lds128(cb);
                        cb += (uint32_t)(K + 1) * 512u;
                        dmma(acc[0][0], acc[0][1], af.x, bf.x);
                        dmma(acc[1][0], acc[1][1], af.y, bf.y);
                        ++K;
                    }
                    // K = J + 1: mirrored tiles S_JK = S_KJ^T, read transposed
                    uint32_t ta = tiles_a + (uint32_t)((tri(K) + J) * 512 + tr0 * 8);
                    for (; K + 1 < nt; K += 2) {
                        const double ax0 = lds64(ta), ay0 = lds64(ta + 256);
                        const double2 bf0 = lds128(cb);
                        const uint32_t ta1 = ta + (uint32_t)(K + 1) * 512u;
                        const double ax1 = lds64(ta1), ay1 = lds64(ta1 + 256);
                        const double2 bf1 = lds128(cb + (uint32_t)(K + 1) * 512u);
                        ta += (uint32_t)(2 * K + 3) * 512u;
                        cb += (uint32_t)(2 * K + 3) * 512u;
                        dmma(acc[0][0], acc[0][1], ax0, bf0.x);
                        dmma(acc[1][0], acc[1][1], ay0, bf0.y);
                        dmma(acc[2][0], acc[2][1], ax1, bf1.x);
                        dmma(acc[3][0], acc[3][1], ay1, bf1.y);
                    }
                    if (K < nt) {
                        const double ax = lds64(ta), ay = lds64(ta + 256);
                        const double2 bf = lds128(cb);
                        dmma(acc[0][0], acc[0][1], ax, bf.x);
                        dmma(acc[1][0], acc[1][1], ay, bf.y);
                    }
                    acc[0][0] = (acc[0][0] + acc[2][0]) + (acc[1][0] + acc[3][0]);
                    acc[0][1] = (acc[0][1] + acc[2][1]) + (acc[1][1] + acc[3][1]);
                    const uint32_t cola = tiles_a + (uint32_t)(I * 512 + lane * 16);      // + tri(K) * 512: tile (K, I)
                    double x0, x1;
                    times_inv(inv_cur, -acc[0][0], -acc[0][1], x0, x1);
                    xs[u][0] = x0; xs[u][1] = x1;
                    if (!is_alpha) {
                        // partial sum of the diagonal tile, S_JI^T L_JI: S_JI^T as an A operand straight from the registers
                        const double s00 = shfl(x0, ts0), s01 = shfl(x1, ts0);
                        const double s10 = shfl(x0, ts0 + 16), s11 = shfl(x1, ts0 + 16);
                        const double2 bf = lds128(cola + (uint32_t)tri(J) * 512u);
                        dmma(pd[0][0], pd[0][1], (lr & 1) ? s01 : s00, bf.x);
                        dmma(pd[1][0], pd[1][1], (lr & 1) ? s11 : s10, bf.y);
                    }
                }
            }
            GTR(I, 3);
            __syncthreads();                                   // A: column I of L is dead
            GTR(I, 4);
#pragma unroll
            for (int u = 0; u < kMaxRows; ++u) {
                const int J = row_of(u);
                if (J >= 0) {
                    const uint32_t dt = tiles_a + (uint32_t)((tri(J) + I) * 512);
                    sts64(dt + oi0 * 8, xs[u][0]);
                    sts64(dt + oi1 * 8, xs[u][1]);
                }
            }
            *reinterpret_cast<double2 *>(part + warp * 64 + lane * 2) = make_double2(pd[0][0] + pd[1][0], pd[0][1] + pd[1][1]);
            if (tid < 64 && I > 0) invs[((I - 1) & 1) * 64 + tid] = wnext;
        }
        __syncthreads();
        if (warp == 0) finish_diag(0);
        for (int i = tid; i < Q; i += kGT2) alpha[i] = tiles[(ntiles + (i >> 3)) * 64 + op_idx(0, i & 7)];
        __syncthreads();
        GTP(2);

        // ---- 2. tables for the forward values of the compiled program ------------------------------------------
        for (int i = tid; i < Q; i += kGT2) {
            tt[i] = i < n ? a.t[i] : 0.0;
            gg[i] = (a.g && i < n) ? a.g[i] : i;
        }
        for (int i = tid; i < Gd; i += kGT2) ginv[i] = -1;
        for (int i = tid; i < ntheta && i < MAX_THETA; i += kGT2) th[i] = theta_g[i];
        __syncthreads();
        const int g0 = gg[0];
        for (int i = tid; i < n; i += kGT2) ginv[gg[i] - g0] = (short)i;
        const int ntab = tp.ntab, ncp = tp.ncp;
        for (int e = tid; e < ntab * G; e += kGT2) {
            const int id = e / G, lg = e - id * G;
            const int s0 = tp.tab_src0[id], s1 = tp.tab_src1[id];
            tab[e] = tree_eval(tp.sop + s0, tp.sarg + s0, nullptr, s1 - s0, th, 0.0, 0.0, (double)lg * a.step, 0,
                               nullptr, 0, nullptr, 0, 0, 0);
        }
        for (int e = tid; e < ncp * Q; e += kGT2) {
            const int id = e / Q, i = e - id * Q;
            const double *cp = th + tp.cp_theta[id];
            sig[e] = 0.5 * (1.0 + tanh((tt[i] - cp[0]) / cp[1]));
        }
        __syncthreads();

        GTP(3);
        GTR(31, 0);
        // ---- 3. reverse mode per entry: lane <-> lag within a group of 32 lags, rows dealt round-robin to the warps --
        // Every warp walks every lag group over its own rows (a = warp, warp + 8, ...), so the triangle's uneven
        // diagonals cost all warps the same; a lag's adjoint sum is then split over the 8 warps and recombined in a
        // fixed order below (bit-reproducible, no atomics).
        const bool grid = a.g != nullptr;
        double gl[MAX_THETA];
        for (int j = 0; j < (int)ntheta; ++j) gl[j] = 0.0;
        double gnoise = 0.0;
        // what peel_root took out of the program: per entry these leaves cost one add (lag tables) or five FLOP (moments)
        const bool resid = tp.clen > 0;
        const uint32_t peel_tab = rp.peel_tab;
        const bool peel_mom = (rp.npeel_lin | rp.npeel_con) != 0;
        const bool regular = grid && gg[n - 1] - g0 == n - 1;  // no gaps in the time grid
        const double tc = 0.5 * (tt[0] + tt[n - 1]);            // moments are taken about the middle of the series
        double m0 = 0.0, m1 = 0.0, m2 = 0.0;
        // Short compiled programs — one leaf, or two leaves under one Plus / Times / tabulated ChangePoint, leaves being
        // lag tables, Linear or Constant (most prior-sampled trees once the stationary sub-trees are folded) — are
        // differentiated in registers without the interpreter: adjoints and parameter sums live in named registers.
        auto short_leaf = [](int o) { return o == OP_TABLE || o == OP_LINEAR || o == OP_CONSTANT; };
        const int k0 = tp.cop[0], k1 = tp.clen == 3 ? tp.cop[1] : 0, kb = tp.clen == 3 ? tp.cop[2] : 0;
        const bool short_prog = resid && short_leaf(k0) &&
                                (tp.clen == 1 || (tp.clen == 3 && short_leaf(k1) &&
                                                  (kb == OP_PLUS || kb == OP_TIMES || kb == OP_CHANGEPOINT_TAB)));
        const int a0 = tp.carg[0], a1 = tp.clen == 3 ? tp.carg[1] : 0, ab = tp.clen == 3 ? tp.carg[2] : 0;
        const double *sgb = sig + (tp.clen == 3 ? tp.caux[2] : 0) * Q;
        double la0[3] = {0.0, 0.0, 0.0}, la1[3] = {0.0, 0.0, 0.0}, cpa[2] = {0.0, 0.0};
        const bool reg_prog = resid && !short_prog && rp.reg_ok;
        RegAcc racc;
#pragma unroll
        for (int q = 0; q < kRegLeaf; ++q) { racc.la[q][0] = 0.0; racc.la[q][1] = 0.0; racc.la[q][2] = 0.0; }
#pragma unroll
        for (int q = 0; q < kRegCp; ++q) { racc.cp[q][0] = 0.0; racc.cp[q][1] = 0.0; }
        const int ngroups = (Gd + 31) >> 5, Gp = ngroups * 32;
        for (int lg = 0; lg < ngroups; ++lg) {
            const int d = lg * 32 + lane;
            double hacc[MAX_TABLES];
#pragma unroll
            for (int j = 0; j < MAX_TABLES; ++j) hacc[j] = 0.0;
            double ha0 = 0.0, ha1 = 0.0;
            double wsum = 0.0, q1 = 0.0, q2 = 0.0;             // this lag's sums of w, w u_a, w u_a^2 (u = t - tc)
            const double fw = d == 0 ? 0.5 : 1.0;              // the entry's weight: W_ab on the diagonal, 2 W_ab below it
            double rh[kRegTab];
#pragma unroll
            for (int j = 0; j < kRegTab; ++j) rh[j] = 0.0;
            if (!resid && regular) {
                // Nothing left to interpret and a complete grid (most particles of a fitted ensemble): the entry is five to
                // ten FP64 operations behind three shared-memory loads. Four rows per trip on two accumulator chains, so
                // that a trip is not one dependent chain through an FP64 pipe the other resident CTA fills with DMMAs.
                double ws1 = 0.0, q1b = 0.0, q2b = 0.0;
                for (int ia = lg * 32 + warp; ia < n; ia += 4 * kGW) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int iau = ia + u * kGW, ibu = iau - d;
                        const bool ok = iau < n && ibu >= 0;
                        const int ic = ok ? iau : 0, jc = ok ? ibu : 0;
                        const double Sab = tiles[(tri(ic >> 3) + (jc >> 3)) * 64 + op_idx(ic & 7, jc & 7)];
                        const double w = ok ? fw * (alpha[ic] * alpha[jc] - Sab) : 0.0;
                        if (u & 1) ws1 += w; else wsum += w;
                        if (peel_mom) {
                            const double ua = tt[ic] - tc, wu = w * ua;
                            if (u & 1) { q1b += wu; q2b += wu * ua; } else { q1 += wu; q2 += wu * ua; }
                        }
                    }
                }
                wsum += ws1; q1 += q1b; q2 += q2b;
            } else
            for (int ia = warp; ia < n; ia += kGW) {
                int ib;
                if (regular) {                                 // complete grid: the partner of row ia at lag d is row ia - d
                    if (ia < lg * 32) continue;
                    ib = ia - d;
                    if (ib < 0) continue;
                } else {
                    const int ga = gg[ia] - g0;
                    if (ga < lg * 32) continue;                // no lag of this group reaches back from row ia
                    const int gb = ga - d;
                    ib = (gb >= 0 && d < Gd) ? ginv[gb] : -1;
                    if (ib < 0) continue;
                }
                const double Sab = tiles[(tri(ia >> 3) + (ib >> 3)) * 64 + op_idx(ia & 7, ib & 7)];
                const double w = fw * (alpha[ia] * alpha[ib] - Sab);
                wsum += w;            // lag 0: d logML / d noise; a lag table under the root: the weight IS the adjoint
                if (!peel_mom && !resid) continue;
                const double ti = tt[ia];
                if (peel_mom && grid) {                        // t_b = t_a - d step: moments of the row time only
                    const double wu = w * (ti - tc);
                    q1 += wu; q2 += wu * (ti - tc);
                    if (!resid) continue;
                }
                const double tj = tt[ib];
                if (peel_mom && !grid) {
                    const double ui = ti - tc, uj = tj - tc;
                    m1 += w * (ui + uj); m2 += w * (ui * uj);
                }
                if (!resid) continue;
                if (short_prog) {
                    auto leaf_val = [&](int o, int arg) {
                        if (o == OP_TABLE) return tab[arg * G + d];
                        if (o == OP_LINEAR) return fma(th[arg + 2], (ti - th[arg]) * (tj - th[arg]), th[arg + 1]);
                        return th[arg];
                    };
                    auto leaf_back = [&](int o, int arg, double ad, double &ha, double (&la)[3]) {
                        if (o == OP_TABLE) ha += ad;
                        else if (o == OP_LINEAR) {
                            const double u = ti - th[arg], v2 = tj - th[arg];
                            la[0] += ad * (-th[arg + 2] * (u + v2));
                            la[1] += ad;
                            la[2] += ad * (u * v2);
                        } else la[0] += ad;
                    };
                    if (kb == 0) {
                        leaf_back(k0, a0, w, ha0, la0);
                    } else {
                        const double v0 = leaf_val(k0, a0), v1 = leaf_val(k1, a1);
                        double w0, w1;
                        if (kb == OP_PLUS) { w0 = w; w1 = w; }
                        else if (kb == OP_TIMES) { w0 = w * v1; w1 = w * v0; }
                        else {
                            const double si = sgb[ia], sj = sgb[ib];
                            const double xi = (ti - th[ab]) / th[ab + 1], xj = (tj - th[ab]) / th[ab + 1];
                            w0 = w * ((1.0 - si) * (1.0 - sj));
                            w1 = w * (si * sj);
                            const double dsi = 2.0 * si * (1.0 - si), dsj = 2.0 * sj * (1.0 - sj);
                            const double dk_dsi = -(1.0 - sj) * v0 + sj * v1;
                            const double dk_dsj = -(1.0 - si) * v0 + si * v1;
                            cpa[0] += w * (dk_dsi * dsi + dk_dsj * dsj) * (-1.0 / th[ab + 1]);
                            cpa[1] += w * (dk_dsi * dsi * (-xi / th[ab + 1]) + dk_dsj * dsj * (-xj / th[ab + 1]));
                        }
                        leaf_back(k0, a0, w0, ha0, la0);
                        leaf_back(k1, a1, w1, ha1, la1);
                    }
                    continue;
                }
                if (reg_prog) {
                    rev_entry_reg(tp, rp, th, ti, tj, d, ia, ib, tab, G, sig, Q, w, racc, rh);
                    continue;
                }
                const double delta = grid ? (double)d * a.step : fabs(ti - tj);
                rev_entry(tp.cop, tp.carg, tp.caux, rp.cleft, 0, tp.clen, th, ti, tj, delta, d, ia, ib, tab, G, sig, Q,
                          w, gl, hacc);
            }
            if (d == 0) gnoise += wsum;
            if (peel_mom) {
                m0 += wsum;
                if (grid) {               // sum w (ua + ub) and sum w ua ub with ub = ua - d step
                    const double dd = (double)d * a.step;
                    m1 += 2.0 * q1 - dd * wsum;
                    m2 += q2 - dd * q1;
                }
            }
            if (reg_prog) {
#pragma unroll
                for (int j = 0; j < kRegTab; ++j) hacc[j] += rh[j];
            }
#pragma unroll
            for (int j = 0; j < MAX_TABLES; ++j)
                if ((peel_tab >> j) & 1u) hacc[j] += wsum;
            if (short_prog) {
                // named accumulators back to their tables (two leaves may share nothing: table ids are distinct)
                if (k0 == OP_TABLE) hacc[a0] += ha0;
                if (k1 == OP_TABLE) hacc[a1] += ha1;
            }
#pragma unroll
            for (int j = 0; j < MAX_TABLES; ++j)
                if (j < ntab) hpart[(warp * a.ntab_cap + j) * Gp + d] = hacc[j];
        }
        for (int q = 0; q < rp.npeel_lin; ++q) {
            // sum w (ti - p0)(tj - p0) and its derivatives from the moments about tc: ti - p0 = (ti - tc) + dl
            const int arg = rp.peel_lin[q];
            const double dl = tc - th[arg];
            gl[arg] += -th[arg + 2] * (m1 + 2.0 * dl * m0);
            gl[arg + 1] += m0;
            gl[arg + 2] += m2 + dl * m1 + (dl * dl) * m0;
        }
        for (int q = 0; q < rp.npeel_con; ++q) gl[rp.peel_con[q]] += m0;
        if (short_prog) {
            if (k0 == OP_LINEAR) { gl[a0] += la0[0]; gl[a0 + 1] += la0[1]; gl[a0 + 2] += la0[2]; }
            else if (k0 == OP_CONSTANT) gl[a0] += la0[0];
            if (k1 == OP_LINEAR) { gl[a1] += la1[0]; gl[a1 + 1] += la1[1]; gl[a1 + 2] += la1[2]; }
            else if (k1 == OP_CONSTANT) gl[a1] += la1[0];
            if (kb == OP_CHANGEPOINT_TAB) { gl[ab] += cpa[0]; gl[ab + 1] += cpa[1]; }
        }
        if (reg_prog) {
            for (int i = 0; i < tp.clen; ++i) {
                const int o = tp.cop[i], arg = tp.carg[i], sl = rp.slot[i];
                if (o == OP_LINEAR || o == OP_CONSTANT) {
                    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
                    for (int q = 0; q < kRegLeaf; ++q)
                        if (sl == q) { g0 = racc.la[q][0]; g1 = racc.la[q][1]; g2 = racc.la[q][2]; }
                    gl[arg] += g0;
                    if (o == OP_LINEAR) { gl[arg + 1] += g1; gl[arg + 2] += g2; }
                } else if (o == OP_CHANGEPOINT_TAB) {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int q = 0; q < kRegCp; ++q)
                        if (sl == q) { c0 = racc.cp[q][0]; c1 = racc.cp[q][1]; }
                    gl[arg] += c0; gl[arg + 1] += c1;
                }
            }
        }
        GTR(31, 1);
        __syncthreads();
        GTP(4);
        // ---- 4. table adjoints through the stationary sub-trees, once per (table, lag) -------------------------
        for (int d = tid; d < G; d += kGT2) {
            for (int j = 0; j < ntab; ++j) {
                double A = 0.0;
#pragma unroll
                for (int w = 0; w < kGW; ++w) A += hpart[(w * a.ntab_cap + j) * Gp + d];
                if (A != 0.0) {
                    double dummy[1];
                    rev_entry(tp.sop, tp.sarg, nullptr, rp.sleft, tp.tab_src0[j], tp.tab_src1[j], th, 0.0, 0.0,
                              (double)d * a.step, 0, 0, 0, nullptr, 0, nullptr, 0, A, gl, dummy);
                }
            }
        }
        // ---- 5. block reduction: warp sums side by side in the (now free) region, one barrier pair per 64 slots ----
        GTR(31, 2);
        __syncthreads();
        double *red = region;                       // [kGW][64]
        for (int j0 = 0; j0 <= (int)ntheta; j0 += 64) {
            const int cnt = min(64, (int)ntheta + 1 - j0);
            for (int jj = 0; jj < cnt; ++jj) {
                const int j = j0 + jj;
                const double v = warp_sum(j < (int)ntheta ? gl[j] : gnoise);
                if (lane == 0) red[warp * 64 + jj] = v;
            }
            __syncthreads();
            if (tid < cnt) {
                double r = 0.0;
                for (int w = 0; w < kGW; ++w) r += red[w * 64 + tid];
                const int j = j0 + tid;
                if (j < (int)ntheta) gout[j] = r; else a.grad_noise[b] = r;
            }
            __syncthreads();
        }
        GTR(31, 3);
        GTP(5);
    }
}

size_t region_bytes_for(int nt, int Gd)
{
    const int Q = nt * 8;
    const size_t sweep = (size_t)kGW * 64 * 8;
    size_t diff = ((size_t)Q + MAX_THETA) * 8 + (size_t)Q * 4 + (size_t)Gd * 2;
    diff = (diff + 15) & ~size_t(15);
    return std::max(sweep, diff);
}

}  // namespace

GradTilePlan plan_grad_tile(int n, int G, int ntab_cap, int ncp_cap, int smem_optin, int smem_per_sm)
{
    GradTilePlan pl{};
    const int nt = (n + 7) / 8, Q = nt * 8;
    pl.nt = nt;
    pl.Gd = G > 0 ? G : n;
    if (pl.Gd > 16384 || nt > 8 * kMaxRows) { pl.ok = 0; return pl; }   // ginv holds 16-bit row indices; row slots per warp
    cudaFuncAttributes fa{};
    size_t static_smem = 4096;
    if (cudaFuncGetAttributes(&fa, grad_tile_kernel) == cudaSuccess) static_smem = fa.sharedSizeBytes + 1024;
    else cudaGetLastError();
    const size_t base = ((size_t)(nt * (nt + 1) / 2 + nt) * 64 + Q + 128) * 8;
    pl.nsec = 0;
    pl.region_bytes = (int)region_bytes_for(nt, pl.Gd);
    pl.smem_bytes = base + pl.region_bytes;
    const size_t Gp = (size_t)((pl.Gd + 31) / 32) * 32;
    pl.scratch_stride = (int)((((size_t)ntab_cap * (G > 0 ? G : 0) + (size_t)ncp_cap * Q + (size_t)kGW * ntab_cap * Gp) * 8 + 255) & ~size_t(255));
    if (pl.scratch_stride == 0) pl.scratch_stride = 256;
    pl.ok = pl.smem_bytes + static_smem - 1024 <= (size_t)smem_optin;
    (void)smem_per_sm;
    return pl;
}

int grad_tile_grid(const GradTilePlan &pl, int64_t B, int num_sms)
{
    int per_sm = 0;
    cudaFuncSetAttribute(grad_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, grad_tile_kernel, kGT2, pl.smem_bytes) != cudaSuccess ||
        per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    return (int)std::min<int64_t>((int64_t)per_sm * num_sms, B);
}

cudaError_t launch_grad_tile(const GradArgs &a, const GradTilePlan &pl, char *scratch, unsigned long long *work_counter,
                             int grid, cudaStream_t stream)
{
    GradTileLayout lay{};
    lay.nt = pl.nt; lay.region_bytes = pl.region_bytes; lay.Gd = pl.Gd;
    lay.scratch_stride = pl.scratch_stride; lay.scratch = scratch; lay.work_counter = work_counter;
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(grad_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) return e;
    grad_tile_kernel<<<grid, kGT2, pl.smem_bytes, stream>>>(a, lay);
    return cudaGetLastError();
}

}  // namespace nagp

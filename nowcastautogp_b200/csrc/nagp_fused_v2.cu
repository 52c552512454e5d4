// nagp_fused_v2.cu — tile kernel (variant 2): fused Gram -> blocked Cholesky -> forward solve ->
// logML / predictive moments, one persistent CTA stream of (scenario, particle) instances.
//
// Layout: the lower triangle of the joint q x q Gram lives in shared memory as 8x8 FP64 tiles
// (tile-packed, 512 B each). The factorisation is left-looking by tile column on DMMA (mma.sync m8n8k4 f64,
// accumulators in registers, operands fetched as one 16-byte LDS per lane from the row-major tile). Panel
// schedule (default): one chain warp per CTA, on hardware scheduler 0 (shared with the other resident CTA's chain
// warp), factors every diagonal tile in registers with shuffles while building its inverse by the same row
// operations; seven row-owning warps accumulate sum_P L_IP L_JP^T with one column of lookahead, solve their tiles of the column with one DMMA
// pair against that inverse, and add the last term of the next column with the solved tiles still in registers.
//
// One tile layout serves as A operand, B operand and accumulator: a sum over k may run in any order, so
// the two k-chunks of a DMMA pair are taken as the even columns (k-slot t <-> column 2t) and the odd
// columns (k-slot t <-> column 2t+1) of the tile instead of columns 0-3 and 4-7. Lane (g, t) then needs
// elements (g, 2t) and (g, 2t+1) of a tile as operand — the very pair it holds as accumulator — so tiles
// are plain row-major, every fragment is one 16-byte access at lane * 16, and no value ever moves
// between lanes on its way from accumulator to operand. The observation vector is carried by one row owner
// (two FMAs per term on the B fragment it has loaded anyway), so z = L^-1 y needs no separate solve.
//
// FP64 has no tcgen05/UMMA kind on sm_100a, so DMMA is the tensor path for this workload; measured
// DMMA peak on this pool's B200: 37.1 TFLOP/s (profiles/r01_fp64_peak.json).
//
// Replaces, per instance, AutoGP's Gram + dpotrf + solves behind
//   /root/reference/src/forecasting.jl:133 (GPModel(dict)), :135 (add_data!), :46 (predict_mvn)
//   /root/reference/src/make_and_fit_model.jl:91 (fit_smc! likelihood evaluations)
// Arithmetic contract: docs/KERNEL_SPEC.md §3-§6.
#include <algorithm>


#include "nagp_kernels.cuh"
#include "nagp_tree.cuh"
#include "nagp_tile.cuh"

#ifndef NAGP_V2_TRACE
#define NAGP_V2_TRACE 0     // n: clock64 stamps of the n-th instance of block 0 for tools/v2_timeline.py; 0 in product builds
#endif
#if NAGP_V2_TRACE
// [column][published, handed over], [warp][phase arrival], misc, [column][warp] arrival at the wait for the inverse
__device__ long long g_v2_trace[32 * 2 + 8 * 8 + 8 + 32 * 8];
extern "C" int nagp_debug_read_v2(long long *out, int count)
{
    return (int)cudaMemcpyFromSymbol(out, g_v2_trace, sizeof(long long) * count);
}
#define VTC(J, i) do { if (tracing && lane == 0) g_v2_trace[(J) * 2 + (i)] = clock64(); } while (0)
#define VTW(ph) do { if (tracing && lane == 0) g_v2_trace[64 + warp * 8 + (ph)] = clock64(); } while (0)
#define VTO(J) do { if (tracing && lane == 0) g_v2_trace[136 + (J) * 8 + warp] = clock64(); } while (0)
#else
#define VTC(J, i) do { } while (0)
#define VTW(ph) do { } while (0)
#define VTO(J) do { } while (0)
#endif

namespace nagp {

namespace {

#ifndef NAGP_V2_WARPS
#define NAGP_V2_WARPS 8
#endif
constexpr int kW2 = NAGP_V2_WARPS;      // warps per CTA of the tile kernel
constexpr int kT2 = kW2 * 32;
// Panel schedule: a dependent FP64 chain shares its scheduler's issue slots and FP64 pipe with whatever else
// runs there (tools/chol8_bench2.cu: the 8x8 factorisation takes 1.2 k cycles next to idle warps or to busy
// warps on the OTHER three schedulers, 1.8 k next to one DMMA-issuing warp on its own scheduler, 4.4 k next to
// three), so the chain is kept away from the row owners as far as possible: one warp of scheduler 0 factors every
// diagonal tile and does nothing else, the six warps of schedulers 1-3 own tile rows, and so does the chain's scheduler
// mate (NAGP_V2_PANEL_ROWS=6 leaves that warp idle: same speed at equal register budget, but seven row owners need
// only three register slots each at 21 tile rows). Roles are dealt by hardware warp slot at kernel start, so the chain
// warps of both resident CTAs sit on the same scheduler.
static_assert(kW2 == 8, "the panel schedule assumes 8 warps: two per scheduler");
#ifndef NAGP_V2_PANEL_ROWS
#define NAGP_V2_PANEL_ROWS 7            // 6: the second warp of the chain's scheduler idles; 7: it owns rows too
#endif
constexpr int kNB = NAGP_V2_PANEL_ROWS; // row-owning warps
constexpr int kTB = kNB * 32;
constexpr int kMaxTilesPerWarp = (29 + kNB - 1) / kNB;   // ceil(nt / kNB), nt <= 29
// The kernel is instantiated per KM = register slots (tile rows) per row-owning warp, 3, 4 and the full size: the
// accumulator / C / solved-tile arrays of a slot take 16 registers, and at 128 registers that is the difference
// between ptxas having room to hoist loads and not (six row owners: 5 slots 5.59 ms, 4 slots 5.43 ms; seven row
// owners with 3 slots — 21 tile rows, q <= 168: the weekly-series workloads — 116 registers, 5.34 ms).
constexpr int kSlots3 = kMaxTilesPerWarp < 3 ? kMaxTilesPerWarp : 3;
constexpr int kSlots4 = kMaxTilesPerWarp < 4 ? kMaxTilesPerWarp : 4;

// DMMA inner loop of the left-looking update for NA tile rows of one warp: per P one 16-byte LDS for the shared B
// fragment (tile (Jc, P)), and per row one 16-byte LDS + two DMMAs (one per k-chunk, on separate accumulator chains).
// With WY the warp that carries the observation vector also adds L_{Jc,P} z_P to its per-lane partial sums (two FMAs on the B fragment it has loaded anyway; the four
// lanes of a row are summed once per column).
template <int KM, int NA, bool WY>
__device__ __forceinline__ void kloop_p(double (&acc)[KM][2][2], double &ys0, double &ys1, uint32_t bp,
                                        uint32_t yp, const uint32_t (&rowa)[KM], int P0, int P1)
{
#pragma unroll 2
    for (int P = P0; P < P1; ++P) {
        const uint32_t off = (uint32_t)P * 512u;
        const double2 bf = lds128(bp + off);
        if (WY) {
            const double2 zf = lds128(yp + (uint32_t)P * 64u);
            ys0 = fma(bf.x, zf.x, ys0);
            ys1 = fma(bf.y, zf.y, ys1);
        }
#pragma unroll
        for (int u = 0; u < NA; ++u) {
            const double2 af = lds128(rowa[u] + off);
            dmma(acc[u][0][0], acc[u][0][1], af.x, bf.x);
            dmma(acc[u][1][0], acc[u][1][1], af.y, bf.y);
        }
    }
}

// Row owner, column J, its tile (J+1, J) in slot U: solve it, store it, add its square to the partial sum of
// diagonal tile J+1 (operands straight from the accumulator registers) and put C_{J+1,J+1} in that tile's place
// for the chain warp.
template <int KM, int U>
__device__ __forceinline__ void hand_over(const double (&c)[KM][2], double (&acc)[KM][2][2],
                                          const uint32_t (&rowa)[KM], const double2 ib, int J)
{
    const double2 g2 = lds128(rowa[U] + (uint32_t)(J + 1) * 512u);
    double x0 = 0.0, x1 = 0.0;
    dmma(x0, x1, c[U][0], ib.x);
    dmma(x0, x1, c[U][1], ib.y);
    sts128(rowa[U] + (uint32_t)J * 512u, x0, x1);
    dmma(acc[U][0][0], acc[U][0][1], x0, x0);
    dmma(acc[U][1][0], acc[U][1][1], x1, x1);
    sts128(rowa[U] + (uint32_t)(J + 1) * 512u, g2.x - (acc[U][0][0] + acc[U][1][0]), g2.y - (acc[U][0][1] + acc[U][1][1]));
    acc[U][0][0] = acc[U][0][1] = acc[U][1][0] = acc[U][1][1] = 0.0;
}

struct V2Layout {
    int nt;            // tile rows/cols of the matrix (rows padded to Q = 8 nt)
    int aux_off[5];    // byte offsets of th, gg, tt, sig, tab inside their home
    int aux_smem[5];   // 1: shared memory (offset from aux base), 0: per-CTA global scratch
    int scratch_stride;   // bytes of global scratch per CTA
    char *scratch;
    unsigned long long *work_counter;   // dynamic instance scheduler (zeroed before the launch)
};

// KEEP: also write the factor, z and the inverse diagonal tiles to HBM for the gradient kernel (a separate
// instantiation, so the forecast path does not carry that code: it measured 4 % slower with it inline).
template <bool KEEP, int KM>
__global__ void __launch_bounds__(kT2, 2) fused_v2_kernel(const FusedArgs a, const V2Layout lay)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ TreeProgram tp;
    __shared__ int s_info;
    __shared__ long long s_next;
    __shared__ double s_red[4][kW2];
    __shared__ unsigned char s_ti[436];         // packed tile index -> tile row I (J = index - tri(I)), nt <= 29

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, k = a.k, h = a.h, m = n + k, q = m + h;
    const int nt = lay.nt, Q = nt * 8, ntiles = tri(nt);
    const bool have_y2 = (a.y2 != nullptr) || k == 0;
    const int ny = have_y2 ? m : n;
    const int G = a.G;

    double *tiles = smem;
    double *yv = tiles + ntiles * 64;
    double *invL = yv + Q;
    char *aux_s = reinterpret_cast<char *>(invL + 128);   // two inverse-tile buffers
    char *aux_g = lay.scratch + (size_t)blockIdx.x * lay.scratch_stride;
    auto aux = [&](int i) { return (lay.aux_smem[i] ? aux_s : aux_g) + lay.aux_off[i]; };
    double *th = reinterpret_cast<double *>(aux(0));
    int *gg = reinterpret_cast<int *>(aux(1));
    double *tt = reinterpret_cast<double *>(aux(2));
    double *sig = reinterpret_cast<double *>(aux(3));
    double *tab = reinterpret_cast<double *>(aux(4));

    // Roles by hardware scheduler, not by warp index: the hardware warp slot (%warpid; slot mod 4 = scheduler) is
    // some permutation of the CTA's warps that differs between co-resident CTAs, and the point of the panel
    // schedule is that the chain warps of BOTH resident CTAs share one scheduler with as little row work as possible.
    // Chain = first warp on scheduler 0; the other warp there is the last row owner (the one with the fewest tile
    // rows) or idles (NAGP_V2_PANEL_ROWS=6); the remaining six own rows in order. (Only speed depends on this: any
    // assignment of one chain warp and kNB row owners is correct.)
    __shared__ int s_sched[kW2];
    __shared__ int s_role[kW2];                 // -1 chain, -2 idle, else row-owner index
    if (lane == 0) {
        unsigned wid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        s_sched[warp] = (int)(wid & 3u);
    }
    __syncthreads();
    if (tid == 0) {
        int chain = -1, idle = -1;
        for (int w = 0; w < kW2; ++w)
            if (s_sched[w] == 0) { if (chain < 0) chain = w; else if (idle < 0) idle = w; }
        if (chain < 0) chain = 0;
        if (idle < 0) idle = (chain == kW2 - 1) ? kW2 - 2 : kW2 - 1;
        int nb = 0;
        // seven row owners: the chain's scheduler mate takes the last index, the one with the fewest tile rows
        for (int w = 0; w < kW2; ++w)
            s_role[w] = (w == chain) ? -1 : (w == idle) ? (kNB == kW2 - 1 ? kNB - 1 : -2) : nb++;
    }
    __syncthreads();
    const int role = s_role[warp];
    for (int I = tid; I < nt; I += kT2)
        for (int J = 0; J <= I; ++J) s_ti[tri(I) + J] = (unsigned char)I;
    // times are common to every instance of the launch
    for (int i = tid; i < Q; i += kT2) {
        tt[i] = i < q ? a.t[i] : 0.0;
        gg[i] = (a.g && i < q) ? a.g[i] : 0;
    }

#if NAGP_V2_TRACE
    int n_inst = 0;
#endif
    // instances differ a lot in cost (kernel-tree size): hand them out dynamically
    for (;;) {
        __syncthreads();   // previous instance fully consumed (shared memory and s_next)
        if (tid == 0) s_next = (long long)atomicAdd(lay.work_counter, 1ull);
        __syncthreads();
        const int64_t b = s_next;
        if (b >= a.B) break;
#if NAGP_V2_TRACE
        const bool tracing = (blockIdx.x == 0 && ++n_inst == NAGP_V2_TRACE);
        if (tracing && tid == 0) g_v2_trace[64 + 64] = b % a.P;
#endif
        VTW(0);
        const int64_t s = b / a.P;
        const int p = (int)(b % a.P);
        const int64_t po = a.prog_off[p], plen = a.prog_off[p + 1] - po;
        const int64_t to = a.theta_off[p], ntheta = a.theta_off[p + 1] - to;
        const double *theta_g = a.theta + s * a.theta_stride_k + to;

        if (tid == 0) s_info = 0;
        if (a.compiled) {
            // compiled once per particle on the host: every (scenario, particle) instance just copies it
            const uint32_t *src = reinterpret_cast<const uint32_t *>(a.compiled + p);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&tp);
            for (int i = tid; i < (int)(sizeof(TreeProgram) / 4); i += kT2) dst[i] = src[i];
        } else if (tid == 0) {
            if (ntheta > MAX_THETA) tp.error = -3;
            else tree_compile(tp, a.prog + po, (int)plen, (int)ntheta, G > 0 ? a.ntab_cap : 0, a.ncp_cap);
        }
        for (int i = tid; i < ntheta && i < MAX_THETA; i += kT2) th[i] = theta_g[i];
        __syncthreads();
        VTW(1);
        if (tp.error) {
            if (tid == 0) {
                a.info[b] = tp.error;
                if (a.logml_n) a.logml_n[b] = nan("");
                if (a.logml_m) a.logml_m[b] = nan("");
                if (a.logw) a.logw[b] = nan("");
            }
            continue;
        }
        const int ntab = tp.ntab, ncp = tp.ncp;

        // ---- lag tables and changepoint sigma tables ------------------------------------------------
        for (int e = tid; e < ntab * G; e += kT2) {
            int id = e / G, lg = e - id * G;
            int s0 = tp.tab_src0[id], s1 = tp.tab_src1[id];
            tab[e] = tree_eval(tp.sop + s0, tp.sarg + s0, nullptr, s1 - s0, th, 0.0, 0.0,
                               (double)lg * a.step, 0, nullptr, 0, nullptr, 0, 0, 0);
        }
        for (int e = tid; e < ncp * Q; e += kT2) {
            int id = e / Q, i = e - id * Q;
            const double *cp = th + tp.cp_theta[id];
            sig[e] = 0.5 * (1.0 + tanh((tt[i] - cp[0]) / cp[1]));
        }
        VTW(2);
        __syncthreads();

        // ---- Gram into row-major tiles: a warp writes two whole tiles per iteration; lane (r, c) owns
        //      the accumulator-layout pair (r, 2c), (r, 2c+1) of each, so stores are contiguous 16 B per lane
        const double nz = a.noise[s * a.noise_stride_k + p];
        const double d_lo = nz + a.jitter;
        const double d_hi = (a.noise_pred >= 0.0 ? a.noise_pred : nz) + a.jitter;
        const bool single_table = (tp.clen == 1 && tp.cop[0] == OP_TABLE);
        EvalCtx cx;
        cx.th = th; cx.tt = tt; cx.tab = tab; cx.sig = sig;
        cx.step = a.step; cx.G = G; cx.Q = Q; cx.grid = a.g != nullptr;
        if (single_table) {
            // One stationary leaf (most particles of a fitted ensemble): an entry is a table look-up by lag. One tile per
            // warp and iteration, four iterations in flight: the loop is a chain of dependent shared-memory loads
            // (tile index -> grid indices -> table) and nothing else.
            const int gr = lane >> 2, gc = (lane & 3) * 2;
#pragma unroll 4
            for (int tix = warp; tix < ntiles; tix += kW2) {
                const int I = s_ti[tix], J = tix - tri(I);
                const int i = I * 8 + gr, j0 = J * 8 + gc;
                const int gi = gg[i];
                const int2 gj = *reinterpret_cast<const int2 *>(gg + j0);
                const int l0 = gi - gj.x, l1 = gi - gj.y;
                double v0 = tab[l0 < 0 ? -l0 : l0], v1 = tab[l1 < 0 ? -l1 : l1];
                if (!(I > J && I * 8 + 7 < q)) {        // diagonal tile or padding
                    const bool ri = i < q;
                    if (!(ri && j0 < q)) v0 = (i == j0) ? 1.0 : 0.0;
                    else if (i == j0) v0 += (i < m) ? d_lo : d_hi;
                    if (!(ri && j0 + 1 < q)) v1 = (i == j0 + 1) ? 1.0 : 0.0;
                    else if (i == j0 + 1) v1 += (i < m) ? d_lo : d_hi;
                }
                *reinterpret_cast<double2 *>(tiles + tix * 64 + lane * 2) = make_double2(v0, v1);
            }
        } else {
            // general trees: a warp interprets the compiled program for three tiles (6 entries per lane) at a time
            // (two tiles: 5.27 ms, three: 5.21 ms, four: spills and a doubled value stack in local memory, slower)
            constexpr int GT = 3, GW = 2 * GT;
            const int gr = lane >> 2, gc = (lane & 3) * 2;
            for (int t0 = warp * GT; t0 < ntiles; t0 += kW2 * GT) {
                int ii[GW], jj[GW], lag[GW], tixs[GT];
                bool interior = true;      // all tiles strictly below the diagonal and free of padding
#pragma unroll
                for (int h2 = 0; h2 < GT; ++h2) {
                    const int tix = min(t0 + h2, ntiles - 1);
                    tixs[h2] = tix;
                    const int I = s_ti[tix], J = tix - tri(I);
                    interior = interior && I > J && I * 8 + 7 < q;
                    const int gi = gg[I * 8 + gr];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int x = h2 * 2 + e;
                        ii[x] = I * 8 + gr;
                        jj[x] = J * 8 + gc + e;
                        const int lg = gi - gg[jj[x]];
                        lag[x] = lg < 0 ? -lg : lg;
                    }
                }
                double out[GW];
                tree_evalw<GW>(tp, cx, ii, jj, lag, out);
                if (!interior) {
#pragma unroll
                    for (int x = 0; x < GW; ++x) {
                        if (!(ii[x] < q && jj[x] < q)) out[x] = (ii[x] == jj[x]) ? 1.0 : 0.0;
                        else if (ii[x] == jj[x]) out[x] += (ii[x] < m) ? d_lo : d_hi;
                    }
                }
#pragma unroll
                for (int h2 = 0; h2 < GT; ++h2)
                    if (t0 + h2 < ntiles)
                        *reinterpret_cast<double2 *>(tiles + tixs[h2] * 64 + lane * 2) = make_double2(out[2 * h2], out[2 * h2 + 1]);
            }
        }
        {
            const double *y1 = a.y1 + b * a.y1_stride;
            for (int jx = tid; jx < Q; jx += kT2) {
                double v = 0.0;
                if (jx < n) v = y1[jx];
                else if (jx < ny) v = a.y2 ? a.y2[s * k + (jx - n)] : y1[jx];
                yv[jx] = v;
            }
        }
        VTW(3);
        __syncthreads();

        // ---- left-looking tile-column Cholesky with one column of lookahead ---------------------------
        // A warp owns the tile rows I == warp (mod 8). Slot u counts them from the bottom (u = 0 is the
        // last row below nt), so the rows still active in column J are always the prefix u < NA and the
        // DMMA loop is instantiated per NA without predicates. The observation vector is one more row
        // (index nt) owned by warp nt mod 8. While the owner of diagonal tile J factors it, the other
        // warps already accumulate column J+1 over P < J (every term that does not need column J).
        const int lr = lane >> 2, lj = lane & 3;
        const int oi0 = op_idx(lr, 2 * lj), oi1 = op_idx(lr, 2 * lj + 1);
        const bool bulk = role >= 0;
        const int bi = bulk ? role : kNB;   // row-owner index 0..5
        const int nreg = bi < nt ? (nt - 1 - bi) / kNB + 1 : 0;   // regular rows of this warp
        const int Ilast = bi + (nreg - 1) * kNB;
        const bool has_y = (bi == nt % kNB);
        const uint32_t tiles_a = smem_addr(tiles), yv_a = smem_addr(yv), invL_a = smem_addr(invL);
        uint32_t rowa[KM];        // shared address of this lane's fragment in tile (I_u, 0)
#pragma unroll
        for (int u = 0; u < KM; ++u) {
            const int I = Ilast - u * kNB;
            rowa[u] = tiles_a + (uint32_t)(tri(I > 0 ? I : 0) * 512 + lane * 16);
        }
        double accn[KM][2][2];   // [slot][k-chunk chain][acc regs] partial sums of the current column
#pragma unroll
        for (int u = 0; u < KM; ++u) { accn[u][0][0] = accn[u][0][1] = accn[u][1][0] = accn[u][1][1] = 0.0; }
        // ---- panel schedule -----------------------------------------------------------------------------
        // Named barriers: 1 = "inverse of diagonal tile J published" (chain warp arrives, row owners wait),
        // 2 = "C_JJ is in its tile" (the owner of row J arrives, the chain warp waits), 3 = "tile (J+1, J) is
        // written" and 4 = "tile (J+2, J) is written" among the row owners (the producer only arrives, the other
        // five wait: nobody waits for a whole column of somebody else's rows; everything else written in column J
        // is first read after barrier 1 of column J+1, which all row owners reach after their stores). A producer
        // meets its barrier next as a waiting member after barrier 1 of the next column, by which time the phase
        // it arrived in is complete. Two inverse buffers: the chain may publish column J+1 while slow rows still
        // solve column J.
        if (role == -1) {
            for (int J = 0; J < nt; ++J) {
                if (J > 0) asm volatile("bar.sync 2, 64;" ::: "memory");
                const uint32_t dt = tiles_a + (uint32_t)((tri(J) + J) * 512 + lane * 16);
                const double2 cj = lds128(dt);
                double d0 = cj.x, d1 = cj.y, w0, w1, piv[8];
                const int bad = chol8_inv(d0, d1, w0, w1, lane, q - J * 8, piv);
                sts128(dt, d0, d1);
                sts128(invL_a + (uint32_t)((J & 1) * 512 + lane * 16), w0, w1);
                if (KEEP && a.Wkeep) {
                    double *wk = a.Wkeep + ((size_t)b * nt + J) * 64;
                    wk[oi0] = w0; wk[oi1] = w1;
                }
                if (bad && lane == 0) s_info = J * 8 + bad;
                __syncwarp();
                VTC(J, 0);
                asm volatile("bar.arrive 1, %0;" ::"n"(kTB + 32) : "memory");
                if (bad) break;
            }
        } else if (bulk) {
            // Entering column J a row owner holds, in registers, C = A_IJ - sum for its rows below diagonal J
            // (c[u], and cyv for the observation row) and the partial sums of column J+1 over P < J (accn, ys).
            // Per column: wait for the inverse, solve (the tile that heads the next column first: its owner folds
            // it into the next diagonal tile and hands that to the chain warp), then with the solved tiles still in
            // registers as A fragments add the last term of column J+1, form its C, and run the lookahead of column
            // J+2 while the chain warp factors diagonal tile J+1.
            const uint32_t yp = yv_a + lj * 16;
            double c[KM][2], cyv = 0.0, ys0 = 0.0, ys1 = 0.0;
            {
                const int ns0 = (bi == 0) ? nreg - 1 : nreg;
#pragma unroll
                for (int u = 0; u < KM; ++u) {
                    c[u][0] = c[u][1] = 0.0;
                    if (u < ns0) { const double2 g2 = lds128(rowa[u]); c[u][0] = g2.x; c[u][1] = g2.y; }
                }
                if (has_y) cyv = lds64(yv_a + lr * 8);
            }
            auto lookahead = [&](int Jc, int P1, int NA) {     // NA: this warp's rows I >= Jc
                if (P1 <= 0) return;
                const uint32_t bp = tiles_a + (uint32_t)(tri(Jc) * 512 + lane * 16);
                if (has_y) {
                    switch (NA) {
                    case 0: kloop_p<KM, 0, true>(accn, ys0, ys1, bp, yp, rowa, 0, P1); break;
                    case 1: kloop_p<KM, 1, true>(accn, ys0, ys1, bp, yp, rowa, 0, P1); break;
                    case 2: kloop_p<KM, (KM >= 2 ? 2 : 1), true>(accn, ys0, ys1, bp, yp, rowa, 0, P1); break;
                    case 3: kloop_p<KM, (KM >= 3 ? 3 : 1), true>(accn, ys0, ys1, bp, yp, rowa, 0, P1); break;
                    case 4: kloop_p<KM, (KM >= 4 ? 4 : 1), true>(accn, ys0, ys1, bp, yp, rowa, 0, P1); break;
                    case 5: kloop_p<KM, (KM >= 5 ? 5 : 1), true>(accn, ys0, ys1, bp, yp, rowa, 0, P1); break;
                    default: break;
                    }
                } else {
                    switch (NA) {
                    case 1: kloop_p<KM, 1, false>(accn, ys0, ys1, bp, yp, rowa, 0, P1); break;
                    case 2: kloop_p<KM, (KM >= 2 ? 2 : 1), false>(accn, ys0, ys1, bp, yp, rowa, 0, P1); break;
                    case 3: kloop_p<KM, (KM >= 3 ? 3 : 1), false>(accn, ys0, ys1, bp, yp, rowa, 0, P1); break;
                    case 4: kloop_p<KM, (KM >= 4 ? 4 : 1), false>(accn, ys0, ys1, bp, yp, rowa, 0, P1); break;
                    case 5: kloop_p<KM, (KM >= 5 ? 5 : 1), false>(accn, ys0, ys1, bp, yp, rowa, 0, P1); break;
                    default: break;
                    }
                }
            };
            // row bookkeeping without divisions: na = this warp's rows I >= J, ph = (bi - J) mod kNB (0: row J is mine,
            // 1: row J+1, 2: row J+2)
            int na = nreg, ph = bi;
            for (int J = 0; J < nt; ++J, ph = (ph == 0 ? kNB - 1 : ph - 1)) {
                const int nsolve = (ph == 0) ? na - 1 : na;              // rows strictly below the diagonal
                if (ph == 0) --na;
                const uint32_t joff = (uint32_t)J * 512u;
                const bool more = J + 1 < nt;
                const bool owns_next = more && ph == 1;
                const int n2 = owns_next ? nsolve - 1 : nsolve;          // rows below diagonal J+1
                VTO(J);
                asm volatile("bar.sync 1, %0;" ::"n"(kTB + 32) : "memory");
                // The chain warp can already be factoring tile J+1 (it only needs the hand-over of one row owner): a
                // failure it flags there must not make a late row owner leave one column before the others.
                {
                    const int fl = s_info;
                    if (fl && ((fl - 1) >> 3) <= J) break;
                }
                // (1) triangular solve of the column: X = C * invL^T
                const double2 ib = lds128(invL_a + (uint32_t)((J & 1) * 512 + lane * 16));
                if (owns_next) {
                    switch (nsolve) {
                    case 1: hand_over<KM, 0>(c, accn, rowa, ib, J); break;
                    case 2: hand_over<KM, (KM >= 2 ? 1 : 0)>(c, accn, rowa, ib, J); break;
                    case 3: hand_over<KM, (KM >= 3 ? 2 : 0)>(c, accn, rowa, ib, J); break;
                    case 4: hand_over<KM, (KM >= 4 ? 3 : 0)>(c, accn, rowa, ib, J); break;
                    case 5: hand_over<KM, (KM >= 5 ? 4 : 0)>(c, accn, rowa, ib, J); break;
                    default: break;
                    }
                    __syncwarp();
                    VTC(J, 1);
                    asm volatile("bar.arrive 2, 64;" ::: "memory");
                    asm volatile("bar.arrive 3, %0;" ::"n"(kTB) : "memory");    // tile (J+1, J) is written
                }
                // topmost row first: the owner of row J+2 releases the others' lookahead as soon as tile (J+2, J) is stored
                const bool owns_next2 = (J + 2 < nt) && ph == 2;
                double x[KM][2];
#pragma unroll
                for (int u = KM - 1; u >= 0; --u) {
                    x[u][0] = x[u][1] = 0.0;
                    if (u < n2) {
                        // the accumulator pair (g, 2t), (g, 2t+1) is this lane's A fragment of both k-chunks
                        dmma(x[u][0], x[u][1], c[u][0], ib.x);
                        dmma(x[u][0], x[u][1], c[u][1], ib.y);
                        sts128(rowa[u] + joff, x[u][0], x[u][1]);
                        if (owns_next2 && u == n2 - 1) {
                            __syncwarp();
                            asm volatile("bar.arrive 4, %0;" ::"n"(kTB) : "memory");
                        }
                    }
                }
                if (has_y) {
                    // z_J = invL * cy: lane (g, t) has row g of the inverse at columns 2t, 2t+1
                    const double ca = shfl(cyv, (2 * lj) * 4), cb = shfl(cyv, (2 * lj + 1) * 4);
                    double part = fma(ib.y, cb, ib.x * ca);
                    part += __shfl_xor_sync(kFull, part, 1);
                    part += __shfl_xor_sync(kFull, part, 2);
                    if (lj == 0) sts64(yv_a + (J * 8 + lr) * 8, part);
                }
                if (!more) break;
                if (!owns_next) asm volatile("bar.sync 3, %0;" ::"n"(kTB) : "memory");   // tile (J+1, J) is written (its owner only arrives)
                else __syncwarp();
                // (2) last term of column J+1 (A fragments = the solved tiles, still in registers) and its C
                const double2 bf = lds128(tiles_a + (uint32_t)((tri(J + 1) + J) * 512 + lane * 16));
                const uint32_t j1off = joff + 512u;
#pragma unroll
                for (int u = 0; u < KM; ++u) {
                    c[u][0] = c[u][1] = 0.0;
                    if (u < n2) {
                        const double2 g2 = lds128(rowa[u] + j1off);
                        dmma(accn[u][0][0], accn[u][0][1], x[u][0], bf.x);
                        dmma(accn[u][1][0], accn[u][1][1], x[u][1], bf.y);
                        c[u][0] = g2.x - (accn[u][0][0] + accn[u][1][0]);
                        c[u][1] = g2.y - (accn[u][0][1] + accn[u][1][1]);
                        accn[u][0][0] = accn[u][0][1] = accn[u][1][0] = accn[u][1][1] = 0.0;
                    }
                }
                if (has_y) {
                    const double2 zf = lds128(yp + (uint32_t)J * 64u);
                    double sy = fma(bf.y, zf.y, ys1) + fma(bf.x, zf.x, ys0);
                    sy += __shfl_xor_sync(kFull, sy, 1);
                    sy += __shfl_xor_sync(kFull, sy, 2);
                    cyv = lds64(yv_a + ((J + 1) * 8 + lr) * 8) - sy;
                    ys0 = ys1 = 0.0;
                }
                // (3) lookahead: column J+2 over P <= J, in the shadow of the factorisation of diagonal tile J+1
                if (J + 2 < nt) {
                    if (!owns_next2) asm volatile("bar.sync 4, %0;" ::"n"(kTB) : "memory");   // tile (J+2, J) is written
                    lookahead(J + 2, J + 1, n2);
                }
            }
        }
        VTW(4);
        __syncthreads();

        if (KEEP && a.Lkeep && !s_info) {
            // keep the factor for the gradient kernel in its fragment-major tile layout (op_idx: columns c and
            // c+4 of a row adjacent), z padded to Q
            double *Lo = a.Lkeep + (size_t)b * ((size_t)ntiles * 64);
            for (int i = tid; i < ntiles * 32; i += kT2) {
                const int tile = i >> 5, e = i & 31, r = e >> 2, cq = e & 3;
                const double *src = tiles + tile * 64 + r * 8 + cq;
                reinterpret_cast<double2 *>(Lo)[i] = make_double2(src[0], src[4]);
            }
            double *zo = a.zkeep + (size_t)b * Q;
            for (int i = tid; i < Q; i += kT2) zo[i] = yv[i];
        }

        if (s_info) {
            if (tid == 0) {
                a.info[b] = s_info;
                if (a.logml_n) a.logml_n[b] = nan("");
                if (a.logml_m) a.logml_m[b] = nan("");
                if (a.logw) a.logw[b] = nan("");
            }
            continue;
        }

        // element (i, j), i >= j, of the factor
        auto Lel = [&](int i, int j) {
            return tiles[(tri(i >> 3) + (j >> 3)) * 64 + (i & 7) * 8 + (j & 7)];
        };

        // ---- logML(n), logML(m) ----------------------------------------------------------------------
        const double *z = yv;
        double ld_n = 0, ld_m = 0, qd_n = 0, qd_m = 0;
        for (int r = tid; r < m; r += kT2) {
            double l = log(Lel(r, r));
            double zz = r < ny ? z[r] * z[r] : 0.0;
            ld_m += l; qd_m += zz;
            if (r < n) { ld_n += l; qd_n += zz; }
        }
        ld_n = warp_sum(ld_n); ld_m = warp_sum(ld_m); qd_n = warp_sum(qd_n); qd_m = warp_sum(qd_m);
        if (lane == 0) { s_red[0][warp] = ld_n; s_red[1][warp] = ld_m; s_red[2][warp] = qd_n; s_red[3][warp] = qd_m; }
        __syncthreads();
        if (tid == 0) {
            double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
            for (int w = 0; w < kW2; ++w) { r0 += s_red[0][w]; r1 += s_red[1][w]; r2 += s_red[2][w]; r3 += s_red[3][w]; }
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
            // one warp per forecast row: lanes stride the m columns, then reduce
            for (int r = warp; r < h; r += kW2) {
                double accv = 0.0;
                for (int cix = lane; cix < m; cix += 32) accv = fma(Lel(m + r, cix), z[cix], accv);
                accv = warp_sum(accv);
                if (lane == 0) a.mu[b * h + r] = (accv - a.yb) / a.ya;
            }
        }
        if (a.L33) {
            for (int e = tid; e < h * h; e += kT2) {
                int r = e / h, cix = e - r * h;
                a.L33[b * h * h + e] = cix <= r ? Lel(m + r, m + cix) / a.ya : 0.0;
            }
        }
        if (a.proj) {
            for (int r = warp; r < kh; r += kW2) {
                double accv = 0.0;
                for (int cix = lane; cix < n; cix += 32) accv = fma(Lel(n + r, cix), z[cix], accv);
                accv = warp_sum(accv);
                if (lane == 0) a.proj[b * kh + r] = accv;
            }
        }
        if (a.Ltail) {
            for (int e = tid; e < kh * kh; e += kT2) {
                int r = e / kh, cix = e - r * kh;
                a.Ltail[b * kh * kh + e] = cix <= r ? Lel(n + r, n + cix) : 0.0;
            }
        }
        VTW(5);
    }
}

// aux arrays: th, gg, tt, sig, tab (bytes)
void aux_sizes(int Q, int G, int ntheta_cap, int ntab_cap, int ncp_cap, size_t (&sz)[5])
{
    auto up = [](size_t x) { return (x + 15) & ~size_t(15); };
    sz[0] = up((size_t)std::max(ntheta_cap, 1) * sizeof(double));
    sz[1] = up((size_t)Q * sizeof(int));
    sz[2] = up((size_t)Q * sizeof(double));
    sz[3] = up((size_t)ncp_cap * Q * sizeof(double));
    sz[4] = up((size_t)ntab_cap * (G > 0 ? G : 0) * sizeof(double));
}

}  // namespace

int fused_v2_max_q() { return 8 * (kMaxTilesPerWarp * kNB - 1); }

// Plans shared memory for the tile kernel: the tiles, yv and invL are mandatory; the aux arrays go
// to shared memory in priority order while the CTA stays within `budget` bytes, else to global
// scratch (L1-resident: a few KB per CTA).
V2Plan plan_fused_v2(int q, int G, int ntheta_cap, int ntab_cap, int ncp_cap, int smem_optin, int smem_per_sm)
{
    V2Plan pl{};
    const int nt = (q + 7) / 8, Q = nt * 8;
    pl.nt = nt;
    size_t base = ((size_t)(nt * (nt + 1) / 2) * 64 + Q + 128) * sizeof(double);
    size_t sz[5];
    aux_sizes(Q, G, ntheta_cap, ntab_cap, ncp_cap, sz);
    size_t total_aux = sz[0] + sz[1] + sz[2] + sz[3] + sz[4];
    cudaFuncAttributes fa{};
    size_t static_smem = 2560 + 1024;
    if (cudaFuncGetAttributes(&fa, fused_v2_kernel<true, kMaxTilesPerWarp>) == cudaSuccess) static_smem = fa.sharedSizeBytes + 1024;   // + per-CTA reservation
    else cudaGetLastError();
    // budget: 2 CTAs/SM if the mandatory part allows it, else everything the opt-in limit gives
    size_t two = (size_t)smem_per_sm / 2;
    size_t budget = (base + static_smem <= two) ? two - static_smem : (size_t)smem_optin - 1536;
    if (base > (size_t)smem_optin - 1536) { pl.ok = 0; return pl; }
    if (base + total_aux <= budget) budget = base + total_aux;   // everything fits
    size_t s_off = 0, g_off = 0;
    const int order[5] = {0, 1, 4, 3, 2};   // th, gg, tab, sig, tt
    for (int oi = 0; oi < 5; ++oi) {
        int i = order[oi];
        if (base + s_off + sz[i] <= budget) { pl.aux_smem[i] = 1; pl.aux_off[i] = (int)s_off; s_off += sz[i]; }
        else { pl.aux_smem[i] = 0; pl.aux_off[i] = (int)g_off; g_off += sz[i]; }
    }
    pl.smem_bytes = base + s_off;
    pl.scratch_stride = (int)((g_off + 255) & ~size_t(255));
    pl.ok = 1;
    return pl;
}

int fused_v2_grid(const V2Plan &pl, int64_t B, int num_sms)
{
    int per_sm = 0;
    cudaFuncSetAttribute(fused_v2_kernel<true, kMaxTilesPerWarp>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_v2_kernel<true, kMaxTilesPerWarp>, kT2, pl.smem_bytes) != cudaSuccess ||
        per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    int64_t g = (int64_t)per_sm * num_sms;
    return (int)std::min<int64_t>(g, B);
}

cudaError_t launch_fused_v2(const FusedArgs &a, const V2Plan &pl, char *scratch, unsigned long long *work_counter,
                            int grid, cudaStream_t stream)
{
    V2Layout lay{};
    lay.work_counter = work_counter;
    cudaError_t e0 = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), stream);
    if (e0 != cudaSuccess) return e0;
    lay.nt = pl.nt;
    for (int i = 0; i < 5; ++i) { lay.aux_off[i] = pl.aux_off[i]; lay.aux_smem[i] = pl.aux_smem[i]; }
    lay.scratch_stride = pl.scratch_stride;
    lay.scratch = scratch;
    const bool keep = a.Lkeep != nullptr;
    const int need = (pl.nt + kNB - 1) / kNB;                  // tile rows per row-owning warp
    auto kern = keep ? (need <= kSlots3 ? fused_v2_kernel<true, kSlots3>
                        : need <= kSlots4 ? fused_v2_kernel<true, kSlots4> : fused_v2_kernel<true, kMaxTilesPerWarp>)
                     : (need <= kSlots3 ? fused_v2_kernel<false, kSlots3>
                        : need <= kSlots4 ? fused_v2_kernel<false, kSlots4> : fused_v2_kernel<false, kMaxTilesPerWarp>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<grid, kT2, pl.smem_bytes, stream>>>(a, lay);
    return cudaGetLastError();
}

}  // namespace nagp

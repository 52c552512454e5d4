// nagp_fused_v3.cu — slot kernel (variant 3): the fused Gram -> blocked Cholesky -> forward solve -> logML /
// predictive moments path with THREE matrices in flight per SM inside one role-specialised CTA.
//
// Why: the tile kernel (nagp_fused_v2.cu) holds the whole lower triangle of an instance in shared memory
// (107 KB at q = 160), so two matrices fit an SM and the latency of the tile-column chain (factor the diagonal
// tile -> solve the tile below -> update the next diagonal tile) is hidden by one other matrix only: 24 % of the
// DMMA peak, issue slots 41 % busy, barrier/wait stalls on top (profiles/r01_v2i_final_summary.csv). Here
//   * a left-looking factorisation only ever reads the rows at or below the current column, so the tiles of a
//     finished row are dead: the factor lives in a POOL of recycled tile slots under a static map computed on
//     the host (113 slots instead of 210 tiles at 20 tile rows), and the Gram is produced on demand, one tile
//     column at a time, straight into the slots its column will be factored in;
//   * one CTA per SM runs three matrix SLOTS of five warps each — one chain warp (factors + inverts every
//     diagonal tile in registers), one Gram producer, three row owners (tile rows dealt mod 3: DMMA accumulation
//     with one column of lookahead, column solve against the published inverse, last term of the next column) —
//     dealt by hardware warp scheduler: the three chain warps share the scheduler that holds three warps, every
//     other scheduler gets one row owner of every slot plus one Gram producer;
//   * the slots are independent instance streams; inside a slot the roles are coupled by monotone counters in
//     shared memory (release stores / acquire loads), never by CTA-wide barriers.
// Same arithmetic as the tile kernel (docs/KERNEL_SPEC.md §3-§6): one 8x8 tile layout as A operand, B operand and
// accumulator, chol8_inv on the diagonal, the observation vector carried by one row owner.
//
// FP64 has no tcgen05/UMMA kind on sm_100a: mma.sync m8n8k4 f64 (DMMA) is the tensor path for this workload.
//
// Replaces, per instance, AutoGP's Gram + dpotrf + solves behind
//   /root/reference/src/forecasting.jl:133 (GPModel(dict)), :135 (add_data!), :46 (predict_mvn)
//   /root/reference/src/make_and_fit_model.jl:91 (fit_smc! likelihood evaluations)
#include <algorithm>
#include <cstring>
#include <vector>

#include "nagp_kernels.cuh"
#include "nagp_tree.cuh"
#include "nagp_tile.cuh"

#ifndef NAGP_V3_TRACE
#define NAGP_V3_TRACE 0   // 1: per-step clock64 stamps of one instance (block 0, slot 0) for tools/v3_timeline.py; 0 in product builds
#endif
#if NAGP_V3_TRACE
__device__ long long g_v3_trace[32 * 5 * 8 + 16];
extern "C" int nagp_debug_read(long long *out, int count)
{
    return (int)cudaMemcpyFromSymbol(out, g_v3_trace, sizeof(long long) * count);
}
#define TR(J, ph) do { if (tracing && lane == 0 && role < 5) g_v3_trace[((J) * 5 + role) * 8 + (ph)] = clock64(); } while (0)
#define TRG(i) do { if (tracing && stid == 0) g_v3_trace[32 * 5 * 8 + (i)] = clock64(); } while (0)
#else
#define TR(J, ph) do { } while (0)
#define TRG(i) do { } while (0)
#endif

namespace nagp {

namespace {

#ifndef NAGP_V3_CHAIN_UNROLLED
#define NAGP_V3_CHAIN_UNROLLED 0   // 1: the chain warp runs the fully unrolled 8x8 factorisation of the tile kernel (700 instructions)
#endif
#ifndef NAGP_V3_RO
#define NAGP_V3_RO 3      // row owners per slot (3: one per row-owner scheduler; 6: two — measured slower)
#endif
constexpr int kM3 = 3;                  // matrix slots per CTA
constexpr int kRO = NAGP_V3_RO;         // row owners per slot: tile rows dealt mod kRO
constexpr int kWS = 2 + kRO;            // warps per slot: chain, Gram producer, row owners
// With six row owners per slot the CTA carries three spare warps that exit at once: 27 warps fall on the four schedulers
// as 7/7/7/6, the scheduler with six holds the three chain warps and the three spares (so the chains stay alone on it), the
// others six row owners and one Gram producer each.
constexpr int kSpare = kRO == 6 ? 3 : 0;
constexpr int kW3 = kM3 * kWS + kSpare; // warps per CTA
constexpr int kT3 = kW3 * 32;
constexpr int kST = kWS * 32;           // threads per slot
constexpr int kMaxNt3 = 21;             // the slot map has one byte per tile of a 21 x 21 triangle
constexpr int kInvBufs = 4;             // inverse-tile buffers (power of two; the chain checks the row owners' progress before reuse)
constexpr int kBig = 1 << 20;           // counter value that releases every waiter (abort)
static_assert(kRO == 3 || kRO == 6, "row owners per slot");

// per-slot control block (ints)
enum : int { C_INV = 0, C_DIAG = 1, C_GRAM = 2, C_INFO = 3, C_TOP = 4, C_DONE = 4 + kRO, C_NEXT = 4 + 2 * kRO /* 2 ints, 8-byte aligned */,
             C_RED = 6 + 2 * kRO };
static_assert((C_NEXT % 2) == 0 && C_RED * 4 + 4 * kWS * 8 <= (kRO <= 3 ? 256 : 384), "control block layout");
constexpr int kCtrlBytes = kRO <= 3 ? 256 : 384;

// Counters in shared memory couple the roles of a slot. The writer's lanes store their data, __syncwarp(), then
// lane 0 stores the counter; the reader polls the counter and then loads the data. NAGP_V3_FENCE=1 makes the
// counter store a release (MEMBAR.ALL.CTA + STS) and the poll an acquire; 0 relies on the in-order shared-memory
// pipeline of an SM (a warp's shared-memory instructions are performed in issue order, and both sides are
// `asm volatile` with memory clobbers, so the compiler keeps the order too).
#ifndef NAGP_V3_FENCE
#define NAGP_V3_FENCE 0
#endif
__device__ __forceinline__ void st_release(uint32_t addr, int v)
{
#if NAGP_V3_FENCE
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
#else
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// spin until the counter at `addr` reaches `want`; `out` = the value seen (>= kBig: abort). A macro, so that the
// profiler attributes every wait to its own source line.
#if NAGP_V3_FENCE
#define NAGP_LD_CNT "ld.acquire.cta.shared.s32"
#else
#define NAGP_LD_CNT "ld.volatile.shared.s32"
#endif
#ifndef NAGP_V3_SLEEP
#define NAGP_V3_SLEEP 32     // ns of nanosleep between polls (0: spin flat out; a tight poll loop steals issue slots and
#endif                       // shared-memory bandwidth from the warps it is waiting for)
#define WAIT_GE(out, addr, want)                                                                   \
    do {                                                                                           \
        asm volatile(NAGP_LD_CNT " %0, [%1];" : "=r"(out) : "r"(addr) : "memory");                 \
        if ((out) >= (want)) break;                                                                \
        if (NAGP_V3_SLEEP) __nanosleep(NAGP_V3_SLEEP);                                             \
    } while (true)

// Code size is a first-order concern here: the L0 instruction cache of a warp scheduler holds about 6 KB (384
// instructions) and the SM's L1.5 32 KB; three slots in different phases plus a Gram producer per scheduler thrash
// anything larger (the first version of this kernel, with the row loops unrolled over register-resident accumulators
// as in the tile kernel, was 15.5 k instructions and spent 50-78 % of its warp samples waiting for instructions).
// So every hot loop below is rolled, rows are addressed at run time in chunks of four, and no accumulator lives in
// registers across a column step: partial sums are parked in the tile's own slot (over the Gram tile).


// W-wide interpreter for programs whose stationary leaves and changepoints are all tabulated (the only programs the
// dispatcher sends to this kernel): OP_TABLE / OP_LINEAR / OP_CONSTANT leaves, Plus / Times / OP_CHANGEPOINT_TAB.
// Same formulas and evaluation order as tree_evalw (nagp_tile.cuh) without its direct transcendental paths.
template <int W>
__device__ __forceinline__ void tree_evalw_tab(const TreeProgram &tp, const EvalCtx &cx, const int (&ii)[W],
                                               const int (&jj)[W], const int (&lag)[W], double (&top)[W])
{
    double st[MAX_STACK][W];
    double sec[W];
    int sp = 0;
    const int len = tp.clen;
#pragma unroll
    for (int e = 0; e < W; ++e) { top[e] = 0.0; sec[e] = 0.0; }
#pragma unroll 1
    for (int o = 0; o < len; ++o) {
        const uint32_t wd = tp.cword[o];
        const int op = wd & 0xff;
        const double *p = cx.th + ((wd >> 8) & 0xffff);
        if (op <= OP_PERIODIC || op == OP_TABLE) {
            if (sp >= 2) {
#pragma unroll
                for (int e = 0; e < W; ++e) st[sp - 2][e] = sec[e];
            }
#pragma unroll
            for (int e = 0; e < W; ++e) sec[e] = top[e];
            ++sp;
            if (op == OP_TABLE) {
                const double *tb = cx.tab + ((wd >> 8) & 0xffff) * cx.G;
#pragma unroll
                for (int e = 0; e < W; ++e) top[e] = tb[lag[e]];
            } else if (op == OP_LINEAR) {
                const double c0 = p[0], b0 = p[1], a0 = p[2];
#pragma unroll
                for (int e = 0; e < W; ++e) top[e] = fma(a0, (cx.tt[ii[e]] - c0) * (cx.tt[jj[e]] - c0), b0);
            } else {
#pragma unroll
                for (int e = 0; e < W; ++e) top[e] = p[0];
            }
        } else {
            --sp;   // left operand is sec, right operand is top
            if (op == OP_PLUS) {
#pragma unroll
                for (int e = 0; e < W; ++e) top[e] = sec[e] + top[e];
            } else if (op == OP_TIMES) {
#pragma unroll
                for (int e = 0; e < W; ++e) top[e] = sec[e] * top[e];
            } else {   // OP_CHANGEPOINT_TAB
                const double *sg = cx.sig + (wd >> 24) * cx.Q;
#pragma unroll
                for (int e = 0; e < W; ++e) {
                    const double si = sg[ii[e]], sj = sg[jj[e]];
                    top[e] = ((1.0 - si) * (1.0 - sj)) * sec[e] + (si * sj) * top[e];
                }
            }
            if (sp >= 2) {
#pragma unroll
                for (int e = 0; e < W; ++e) sec[e] = st[sp - 2][e];
            }
        }
    }
}

// Row-owner building blocks over CH owned rows I0, I0 + 3, ... (all valid: the callers split a row set into chunks of
// 4, 2 and 1, so no operand is ever loaded for a row that does not exist — shared-memory bandwidth, 128 B per cycle per
// SM for nine row owners, three chain warps and three producers, is the scarcest resource of this kernel).
struct RowCtx {
    uint32_t pool_lane, smap_a;
};
__device__ __forceinline__ uint32_t tile_addr(const RowCtx &rc, int I, int P)
{
    return rc.pool_lane + (lds_u8(rc.smap_a + (uint32_t)(tri(I) + P)) << 9);
}
// (a) X = C inv^T for tiles (I, J)
template <int CH>
__device__ __forceinline__ void solve_chunk(const RowCtx &rc, int I0, int J, const double2 ib)
{
    uint32_t sa[CH];
    double2 cc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) sa[c] = tile_addr(rc, I0 + kRO * c, J);
#pragma unroll
    for (int c = 0; c < CH; ++c) cc[c] = lds128(sa[c]);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        double x0 = 0.0, x1 = 0.0;
        dmma(x0, x1, cc[c].x, ib.x);
        dmma(x0, x1, cc[c].y, ib.y);
        sts128(sa[c], x0, x1);
    }
}
// (c) tile (I, J+1) -= X_I L_{J+1,J}^T
template <int CH>
__device__ __forceinline__ void lastterm_chunk(const RowCtx &rc, int I0, int J, const double2 bf)
{
    uint32_t ca[CH];
    double2 xa[CH], pk[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const uint32_t mp = rc.smap_a + (uint32_t)(tri(I0 + kRO * c) + J);
        xa[c] = lds128(rc.pool_lane + (lds_u8(mp) << 9));
        ca[c] = rc.pool_lane + (lds_u8(mp + 1) << 9);
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) pk[c] = lds128(ca[c]);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        double e0 = 0.0, e1 = 0.0, o0 = 0.0, o1 = 0.0;
        dmma(e0, e1, xa[c].x, bf.x);
        dmma(o0, o1, xa[c].y, bf.y);
        sts128(ca[c], pk[c].x - (e0 + o0), pk[c].y - (e1 + o1));
    }
}
// (d) tile (I, Jc) -= sum_{P<P1} L_IP L_{Jc,P}^T. (A hand-pipelined version of this loop — slot bytes two terms ahead,
// operand tiles one term ahead — measured slower: 5.80 ms against 5.49 at three row owners; the registers it needs
// cost more than the overlap gains.)
template <int CH>
__device__ __forceinline__ void lookahead_chunk(const RowCtx &rc, int I0, int Jc, int P1)
{
    const uint32_t mapB = rc.smap_a + (uint32_t)tri(Jc);
    uint32_t mapA[CH];
    double acc[CH][2][2];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        mapA[c] = rc.smap_a + (uint32_t)tri(I0 + kRO * c);
        acc[c][0][0] = acc[c][0][1] = acc[c][1][0] = acc[c][1][1] = 0.0;
    }
#pragma unroll 1
    for (int P = 0; P < P1; ++P) {
        const double2 bfp = lds128(rc.pool_lane + (lds_u8(mapB + P) << 9));
        double2 af[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) af[c] = lds128(rc.pool_lane + (lds_u8(mapA[c] + P) << 9));
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            dmma(acc[c][0][0], acc[c][0][1], af[c].x, bfp.x);
            dmma(acc[c][1][0], acc[c][1][1], af[c].y, bfp.y);
        }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const uint32_t ta = rc.pool_lane + (lds_u8(mapA[c] + Jc) << 9);
        const double2 g2 = lds128(ta);
        sts128(ta, g2.x - (acc[c][0][0] + acc[c][1][0]), g2.y - (acc[c][0][1] + acc[c][1][1]));
    }
}
// rows I0, I0 + 3, ... < nt in chunks of 4, then one chunk of the 3, 2 or 1 rows that remain
#define NAGP_ROW_CHUNKS(FN, I0_, ...)                                              \
    do {                                                                           \
        int i0__ = (I0_), nrem__ = (nt - i0__ + kRO - 1) / kRO;                    \
        for (; nrem__ >= 4; nrem__ -= 4, i0__ += 4 * kRO) FN<4>(rc, i0__, __VA_ARGS__); \
        if (nrem__ == 3) FN<3>(rc, i0__, __VA_ARGS__);                             \
        else if (nrem__ == 2) FN<2>(rc, i0__, __VA_ARGS__);                        \
        else if (nrem__ == 1) FN<1>(rc, i0__, __VA_ARGS__);                        \
    } while (0)

__global__ void __launch_bounds__(kT3, 1) fused_v3_kernel(const FusedArgs a, const V3Plan lay, unsigned long long *work_counter)
{
    extern __shared__ __align__(16) unsigned char smem3[];
    __shared__ int s_sched[kW3];
    __shared__ signed char s_slot[kW3], s_role[kW3];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, k = a.k, h = a.h, m = n + k, q = m + h;
    const int nt = lay.nt, Q = nt * 8;
    const bool have_y2 = (a.y2 != nullptr) || k == 0;
    const int ny = have_y2 ? m : n;
    const int G = a.G;

    // ---- roles by hardware warp scheduler (%warpid mod 4): only speed depends on the assignment ----------------
    if (lane == 0) {
        unsigned wid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        s_sched[warp] = (int)(wid & 3u);
    }
    double *tt = reinterpret_cast<double *>(smem3 + lay.off_tt);
    int *gg = reinterpret_cast<int *>(smem3 + lay.off_gg);
    unsigned char *smap = smem3 + lay.off_map;
    unsigned char *need = smem3 + lay.off_need;
    for (int i = tid; i < Q; i += kT3) {
        tt[i] = i < q ? a.t[i] : 0.0;
        gg[i] = (a.g && i < q) ? a.g[i] : 0;
    }
    for (int i = tid; i < (int)sizeof(lay.smap); i += kT3) smap[i] = lay.smap[i];
    for (int i = tid; i < (int)sizeof(lay.need); i += kT3) need[i] = lay.need[i];
    __syncthreads();
    if (tid == 0) {
        int cnt[4] = {0, 0, 0, 0}, so[4] = {0, 1, 2, 3};
#pragma unroll 1
        for (int w = 0; w < kW3; ++w) cnt[s_sched[w]]++;
        for (int i = 0; i < 4; ++i)
            for (int j = i + 1; j < 4; ++j)
                if (cnt[so[j]] < cnt[so[i]]) { int t = so[i]; so[i] = so[j]; so[j] = t; }
        // (thread 0 only, once per kernel: rolled loops over shared arrays, so this costs no registers in the role code)
        auto pick = [&](int sched) {
#pragma unroll 1
            for (int w = 0; w < kW3; ++w) if (s_role[w] < 0 && s_sched[w] == sched) return w;
#pragma unroll 1
            for (int w = 0; w < kW3; ++w) if (s_role[w] < 0) return w;
            return 0;
        };
#pragma unroll 1
        for (int w = 0; w < kW3; ++w) s_role[w] = -1;
        // chains on the least populated scheduler (with the spare warps, if any); row owner r of every slot on scheduler
        // 1 + r mod 3 of the remaining three, one Gram producer on each of those
        for (int mm = 0; mm < kM3; ++mm) { int w = pick(so[0]); s_slot[w] = (signed char)mm; s_role[w] = 0; }
        for (int i = 0; i < kSpare; ++i) { int w = pick(so[0]); s_slot[w] = 0; s_role[w] = 127; }      // spare: exits below
        for (int mm = 0; mm < kM3; ++mm)
            for (int r = 0; r < kRO; ++r) { int w = pick(so[1 + r % 3]); s_slot[w] = (signed char)mm; s_role[w] = (signed char)(2 + r); }
        for (int mm = 0; mm < kM3; ++mm) { int w = pick(so[1 + mm % 3]); s_slot[w] = (signed char)mm; s_role[w] = 1; }
    }
    __syncthreads();
    const int slot = s_slot[warp], role = s_role[warp];
    if (role == 127) return;                  // spare warp (keeps the chains' scheduler free of row work)
    const int stid = role * 32 + lane;        // thread index inside the slot

    unsigned char *sb = smem3 + lay.off_slots + (size_t)slot * lay.slot_stride;
    int *ctrl = reinterpret_cast<int *>(sb + lay.o_ctrl);
    TreeProgram &tp = *reinterpret_cast<TreeProgram *>(sb + lay.o_tp);
    double *yv = reinterpret_cast<double *>(sb + lay.o_yv);
    double *diagv = reinterpret_cast<double *>(sb + lay.o_diag);
    double *th = reinterpret_cast<double *>(sb + lay.o_th);
    double *tab = reinterpret_cast<double *>(sb + lay.o_tab);
    double *sig = reinterpret_cast<double *>(sb + lay.o_sig);
    double *pool = reinterpret_cast<double *>(sb + lay.o_pool);
    double *s_red = reinterpret_cast<double *>(ctrl + C_RED);
    const uint32_t ctrl_a = smem_addr(ctrl), pool_a = smem_addr(pool), yv_a = smem_addr(yv), inv_a = smem_addr(sb + lay.o_inv),
                   diag_a = smem_addr(diagv), smap_a = smem_addr(smap);
    const uint32_t pool_lane = pool_a + lane * 16;
    const int barid = 1 + slot;
    auto slotbar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(barid), "n"(kST) : "memory"); };
    // shared address of this lane's fragment of tile (I, P)
    auto tile_at = [&](int I, int P) { return pool_lane + (lds_u8(smap_a + (uint32_t)(tri(I) + P)) << 9); };

#if NAGP_V3_TRACE
    int n_inst = 0;
#endif
    for (;;) {
        slotbar();   // previous instance fully consumed
        if (stid == 0) {
            *reinterpret_cast<long long *>(ctrl + C_NEXT) = (long long)atomicAdd(work_counter, 1ull);
            for (int i = 0; i < C_NEXT; ++i) ctrl[i] = 0;
        }
        slotbar();
        const int64_t b = *reinterpret_cast<long long *>(ctrl + C_NEXT);
        if (b >= a.B) break;
        const int64_t s = b / a.P;
        const int p = (int)(b % a.P);
#if NAGP_V3_TRACE
        ++n_inst;
        const bool tracing = (blockIdx.x == 0 && slot == 0 && n_inst == NAGP_V3_TRACE);
        TRG(0);
        if (tracing && stid == 0) g_v3_trace[32 * 5 * 8 + 8] = p;
#endif
        const int64_t to = a.theta_off[p], ntheta = a.theta_off[p + 1] - to;
        const double *theta_g = a.theta + s * a.theta_stride_k + to;
        {
            // programs are compiled once per particle on the host (the dispatcher guarantees it for this kernel)
            const uint32_t *src = reinterpret_cast<const uint32_t *>(a.compiled + p);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&tp);
            for (int i = stid; i < (int)(sizeof(TreeProgram) / 4); i += kST) dst[i] = src[i];
        }
        for (int i = stid; i < ntheta && i < MAX_THETA; i += kST) th[i] = theta_g[i];
        {
            const double *y1 = a.y1 + b * a.y1_stride;
            for (int jx = stid; jx < Q; jx += kST) {
                double v = 0.0;
                if (jx < n) v = y1[jx];
                else if (jx < ny) v = a.y2 ? a.y2[s * k + (jx - n)] : y1[jx];
                yv[jx] = v;
            }
        }
        slotbar();
        if (tp.error) {
            if (stid == 0) {
                a.info[b] = tp.error;
                if (a.logml_n) a.logml_n[b] = nan("");
                if (a.logml_m) a.logml_m[b] = nan("");
                if (a.logw) a.logw[b] = nan("");
            }
            continue;
        }
        const int ntab = tp.ntab, ncp = tp.ncp;
        // ---- lag tables and changepoint sigma tables (all five warps of the slot) ------------------------------
        for (int e = stid; e < ntab * G; e += kST) {
            int id = e / G, lg = e - id * G;
            int s0 = tp.tab_src0[id], s1 = tp.tab_src1[id];
            tab[e] = tree_eval(tp.sop + s0, tp.sarg + s0, nullptr, s1 - s0, th, 0.0, 0.0,
                               (double)lg * a.step, 0, nullptr, 0, nullptr, 0, 0, 0);
        }
        for (int e = stid; e < ncp * Q; e += kST) {
            int id = e / Q, i = e - id * Q;
            const double *cp = th + tp.cp_theta[id];
            sig[e] = 0.5 * (1.0 + tanh((tt[i] - cp[0]) / cp[1]));
        }
        slotbar();

        TRG(1);
        const int lr = lane >> 2, lj = lane & 3;
        if (role == 0) {
            // ================= chain warp: every diagonal tile, in registers ====================================
#pragma unroll 1
            for (int J = 0; J < nt; ++J) {
                int seen;
                if (J == 0) WAIT_GE(seen, ctrl_a + C_GRAM * 4, 1);
                else WAIT_GE(seen, ctrl_a + C_DIAG * 4, J + 1);
                if (J >= kInvBufs) {
                    // the inverse buffer of step J was last read in step J - kInvBufs: wait until every row owner is past it
                    for (int r = 0; r < kRO; ++r) WAIT_GE(seen, ctrl_a + (C_DONE + r) * 4, J - kInvBufs + 1);
                }
                TR(J, 0);
                const uint32_t dt = tile_at(J, J);
                const double2 cj = lds128(dt);
                double d0 = cj.x, d1 = cj.y, w0, w1;
#if NAGP_V3_CHAIN_UNROLLED
                double piv[8];
                const int bad = chol8_inv(d0, d1, w0, w1, lane, q - J * 8, piv);
#else
                const int bad = chol8_inv_rolled(d0, d1, w0, w1, lane, q - J * 8);
#endif
                sts128(dt, d0, d1);
                sts128(inv_a + (uint32_t)((J & (kInvBufs - 1)) * 512 + lane * 16), w0, w1);
                if (lj == (lr >> 1)) sts64(diag_a + (uint32_t)(J * 8 + lr) * 8, (lr & 1) ? d1 : d0);
                if (bad && lane == 0) ctrl[C_INFO] = J * 8 + bad;
                __syncwarp();
                if (lane == 0) st_release(ctrl_a + C_INV * 4, bad ? kBig : J + 1);
                TR(J, 1);
                if (bad) break;
            }
        } else if (role == 1) {
            // ================= Gram producer: one tile column at a time into the recycled slots =================
            const double nz = a.noise[s * a.noise_stride_k + p];
            const double d_lo = nz + a.jitter;
            const double d_hi = (a.noise_pred >= 0.0 ? a.noise_pred : nz) + a.jitter;
            const bool single_table = (tp.clen == 1 && tp.cop[0] == OP_TABLE);
            const bool single_linear = (tp.clen == 1 && tp.cop[0] == OP_LINEAR);
            const double lin_c = single_linear ? th[tp.carg[0]] : 0.0, lin_b = single_linear ? th[tp.carg[0] + 1] : 0.0,
                         lin_a = single_linear ? th[tp.carg[0] + 2] : 0.0;
            EvalCtx cx;
            cx.th = th; cx.tt = tt; cx.tab = tab; cx.sig = sig;
            cx.step = a.step; cx.G = G; cx.Q = Q; cx.grid = a.g != nullptr;
            const int gr = lane >> 2, gc = (lane & 3) * 2;
            bool aborted = false;
#pragma unroll 1
            for (int C = 0; C < nt; ++C) {
                const int nd = need[C];
                if (nd > 0) {
                    // the slots of this column are free once every row owner has finished step nd - 1
                    int seen;
                    for (int r = 0; r < kRO; ++r) WAIT_GE(seen, ctrl_a + (C_DONE + r) * 4, nd);
                    if (ctrl[C_INFO]) { aborted = true; break; }
                }
                TR(C, 0);
                const uint32_t mapC = smap_a + C;    // + tri(I): slot of tile (I, C)
                if (single_linear) {
                    // one Linear leaf: fma(a, (t_i - c)(t_j - c), b), the column factors hoisted
                    const int j0 = C * 8 + gc;
                    const double uj0 = tt[j0] - lin_c, uj1 = tt[j0 + 1] - lin_c;
#pragma unroll 4
                    for (int I = C; I < nt; ++I) {
                        const int i = I * 8 + gr;
                        const double ui = tt[i] - lin_c;
                        double v0 = fma(lin_a, ui * uj0, lin_b), v1 = fma(lin_a, ui * uj1, lin_b);
                        if (!(I > C && I * 8 + 7 < q)) {        // diagonal tile or padding
                            const bool ri = i < q;
                            if (!(ri && j0 < q)) v0 = (i == j0) ? 1.0 : 0.0;
                            else if (i == j0) v0 += (i < m) ? d_lo : d_hi;
                            if (!(ri && j0 + 1 < q)) v1 = (i == j0 + 1) ? 1.0 : 0.0;
                            else if (i == j0 + 1) v1 += (i < m) ? d_lo : d_hi;
                        }
                        sts128(pool_lane + (lds_u8(mapC + tri(I)) << 9), v0, v1);
                    }
                } else if (single_table) {
                    const int j0 = C * 8 + gc;
                    const int2 gj = *reinterpret_cast<const int2 *>(gg + j0);
#pragma unroll 4
                    for (int I = C; I < nt; ++I) {
                        const int i = I * 8 + gr;
                        const int gi = gg[i];
                        const int l0 = gi - gj.x, l1 = gi - gj.y;
                        double v0 = tab[l0 < 0 ? -l0 : l0], v1 = tab[l1 < 0 ? -l1 : l1];
                        if (!(I > C && I * 8 + 7 < q)) {        // diagonal tile or padding
                            const bool ri = i < q;
                            if (!(ri && j0 < q)) v0 = (i == j0) ? 1.0 : 0.0;
                            else if (i == j0) v0 += (i < m) ? d_lo : d_hi;
                            if (!(ri && j0 + 1 < q)) v1 = (i == j0 + 1) ? 1.0 : 0.0;
                            else if (i == j0 + 1) v1 += (i < m) ? d_lo : d_hi;
                        }
                        sts128(pool_lane + (lds_u8(mapC + tri(I)) << 9), v0, v1);
                    }
                } else {
                    constexpr int GT = 3, GW = 2 * GT;     // three tiles per interpreter pass (two: 5.96 ms against 5.49)
#pragma unroll 1
                    for (int I0 = C; I0 < nt; I0 += GT) {
                        int ii[GW], jj[GW], lag[GW];
                        uint32_t dst[GT];
                        bool interior = true;
#pragma unroll
                        for (int h2 = 0; h2 < GT; ++h2) {
                            const int I = min(I0 + h2, nt - 1);
                            dst[h2] = pool_lane + (lds_u8(mapC + tri(I)) << 9);
                            interior = interior && I > C && I * 8 + 7 < q;
                            const int gi = gg[I * 8 + gr];
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int x = h2 * 2 + e;
                                ii[x] = I * 8 + gr;
                                jj[x] = C * 8 + gc + e;
                                const int lg = gi - gg[jj[x]];
                                lag[x] = lg < 0 ? -lg : lg;
                            }
                        }
                        double out[GW];
                        tree_evalw_tab<GW>(tp, cx, ii, jj, lag, out);
                        if (!interior) {
#pragma unroll
                            for (int x = 0; x < GW; ++x) {
                                if (!(ii[x] < q && jj[x] < q)) out[x] = (ii[x] == jj[x]) ? 1.0 : 0.0;
                                else if (ii[x] == jj[x]) out[x] += (ii[x] < m) ? d_lo : d_hi;
                            }
                        }
#pragma unroll
                        for (int h2 = 0; h2 < GT; ++h2)
                            if (I0 + h2 < nt) sts128(dst[h2], out[2 * h2], out[2 * h2 + 1]);
                    }
                }
                __syncwarp();
                if (lane == 0) st_release(ctrl_a + C_GRAM * 4, C + 1);
                TR(C, 1);
            }
            if (aborted) { __syncwarp(); if (lane == 0) st_release(ctrl_a + C_GRAM * 4, kBig); }
        } else {
            // ================= row owner r: tile rows I == r (mod 3) =============================================
            // Entering column step J, tile (I, J) of every owned row I >= J holds C_IJ = G_IJ - sum_{P<J} L_IP L_JP^T and
            // tile (I, J+1) holds G - sum_{P<J} (the lookahead of step J-1 parked it there). Per step: wait for the
            // inverse of diagonal tile J; (a) solve the owned tiles of column J, topmost row first (row J+1 heads the
            // next column: its owner folds it into diagonal tile J+1 and hands that to the chain warp; row J+2 is the
            // B operand of everybody's lookahead); (c) subtract the last term from column J+1; (d) lookahead: subtract
            // sum_{P<=J} from column J+2 while the chain warp factors diagonal tile J+1.
            const int r = role - 2;
            const bool has_y = (r == nt % kRO);
            const uint32_t yp = yv_a + lj * 16;
            RowCtx rc;
            rc.pool_lane = pool_lane; rc.smap_a = smap_a;
            double cyv = has_y ? lds64(yv_a + lr * 8) : 0.0, ys0 = 0.0, ys1 = 0.0;
            bool aborted = false;
#pragma unroll 1
            for (int J = 0; J < nt; ++J) {
                const bool more = J + 1 < nt;
                int It = J + 1 + (r + kRO * nt - (J + 1)) % kRO;    // first owned row strictly below the diagonal
                // the Gram tiles of column J+1 (and of column 0 at the start) are in their slots
                int seen;
                WAIT_GE(seen, ctrl_a + C_GRAM * 4, more ? J + 2 : J + 1);
                WAIT_GE(seen, ctrl_a + C_INV * 4, J + 1);
                if (seen >= kBig) { aborted = true; break; }
                TR(J, 0);
                const double2 ib = lds128(inv_a + (uint32_t)((J & (kInvBufs - 1)) * 512 + lane * 16));
                // (a) topmost owned row first
                int Irest = It;
                if (It < nt) {
                    const uint32_t sa = tile_at(It, J);
                    const double2 cc = lds128(sa);
                    double x0 = 0.0, x1 = 0.0;
                    dmma(x0, x1, cc.x, ib.x);
                    dmma(x0, x1, cc.y, ib.y);
                    sts128(sa, x0, x1);
                    if (It == J + 1) {
                        // heads the next column: diagonal tile J+1 -= X X^T (operands straight from the accumulator
                        // registers), then it is the chain warp's
                        const uint32_t da = tile_at(It, It);
                        const double2 g2 = lds128(da);
                        double s00 = 0.0, s01 = 0.0, s10 = 0.0, s11 = 0.0;
                        dmma(s00, s01, x0, x0);
                        dmma(s10, s11, x1, x1);
                        sts128(da, g2.x - (s00 + s10), g2.y - (s01 + s11));
                        __syncwarp();
                        if (lane == 0) st_release(ctrl_a + C_DIAG * 4, J + 2);
                    } else {
                        __syncwarp();
                    }
                    Irest = It + kRO;
                }
                if (lane == 0) st_release(ctrl_a + (C_TOP + r) * 4, J + 1);
                // the other owned rows
                NAGP_ROW_CHUNKS(solve_chunk, Irest, J, ib);
                if (has_y) {
                    // z_J = invL * cy: lane (g, t) has row g of the inverse at columns 2t, 2t+1
                    const double ca = shfl(cyv, (2 * lj) * 4), cb = shfl(cyv, (2 * lj + 1) * 4);
                    double part = fma(ib.y, cb, ib.x * ca);
                    part += __shfl_xor_sync(kFull, part, 1);
                    part += __shfl_xor_sync(kFull, part, 2);
                    if (lj == 0) sts64(yv_a + (J * 8 + lr) * 8, part);
                }
                __syncwarp();
                TR(J, 1);
                if (!more) break;
                // (b) tile (J+1, J) is written (its owner published it)
                if (It != J + 1) {
                    WAIT_GE(seen, ctrl_a + (C_TOP + (J + 1) % kRO) * 4, J + 1);
                    if (seen >= kBig) { aborted = true; break; }
                }
                TR(J, 2);
                const int I2 = (It == J + 1) ? It + kRO : It;           // first owned row below diagonal J+1
                // (c) last term of column J+1: tile (I, J+1) -= X_I L_{J+1,J}^T  (now C_{I,J+1})
                const double2 bf = lds128(tile_at(J + 1, J));
                NAGP_ROW_CHUNKS(lastterm_chunk, I2, J, bf);
                if (has_y) {
                    // observation row of column J+1: cy = y_{J+1} - sum_{P<=J} L_{J+1,P} z_P; the terms P < J were
                    // summed in the lookahead of the previous step (the slots of row J+1 are recycled after it)
                    const double2 zf = lds128(yp + (uint32_t)J * 64u);
                    double sy = fma(bf.y, zf.y, ys1) + fma(bf.x, zf.x, ys0);
                    sy += __shfl_xor_sync(kFull, sy, 1);
                    sy += __shfl_xor_sync(kFull, sy, 2);
                    cyv = lds64(yv_a + ((J + 1) * 8 + lr) * 8) - sy;
                    ys0 = ys1 = 0.0;
                }
                TR(J, 3);
                // (d) lookahead: column J+2 over P <= J, in the shadow of the factorisation of diagonal tile J+1
                if (J + 2 < nt) {
                    WAIT_GE(seen, ctrl_a + C_GRAM * 4, J + 3);     // the partial sums are parked over the Gram tiles of column J+2
                    if ((J + 2) % kRO != r) {
                        WAIT_GE(seen, ctrl_a + (C_TOP + (J + 2) % kRO) * 4, J + 1);
                        if (seen >= kBig) { aborted = true; break; }
                    }
                    TR(J, 4);
                    const uint32_t mapB = smap_a + (uint32_t)tri(J + 2);
                    if (has_y) {
                        // observation row of column J+2 over the same terms: sum_{P<=J} L_{J+2,P} z_P
#pragma unroll 2
                        for (int P = 0; P <= J; ++P) {
                            const double2 lf = lds128(pool_lane + (lds_u8(mapB + P) << 9));
                            const double2 zf = lds128(yp + (uint32_t)P * 64u);
                            ys0 = fma(lf.x, zf.x, ys0);
                            ys1 = fma(lf.y, zf.y, ys1);
                        }
                    }
                    NAGP_ROW_CHUNKS(lookahead_chunk, I2, J + 2, J + 1);
                }
                __syncwarp();
                if (lane == 0) st_release(ctrl_a + (C_DONE + r) * 4, J + 1);
                TR(J, 5);
            }
            __syncwarp();
            if (lane == 0) {
                if (aborted) st_release(ctrl_a + (C_TOP + r) * 4, kBig);
                st_release(ctrl_a + (C_DONE + r) * 4, kBig);   // nothing of this instance is read by the rows any more
            }
        }
        slotbar();
        TRG(2);

        const int s_info = ctrl[C_INFO];
        if (s_info) {
            if (stid == 0) {
                a.info[b] = s_info;
                if (a.logml_n) a.logml_n[b] = nan("");
                if (a.logml_m) a.logml_m[b] = nan("");
                if (a.logw) a.logw[b] = nan("");
            }
            continue;
        }

        // element (i, j), i >= j, of the factor: rows at or below lay.iep are kept to the end
        auto Lel = [&](int i, int j) {
            return pool[(size_t)smap[tri(i >> 3) + (j >> 3)] * 64 + (i & 7) * 8 + (j & 7)];
        };

        // ---- logML(n), logML(m) ----------------------------------------------------------------------
        const double *z = yv;
        double ld_n = 0, ld_m = 0, qd_n = 0, qd_m = 0;
        for (int rr = stid; rr < m; rr += kST) {
            double l = log(diagv[rr]);
            double zz = rr < ny ? z[rr] * z[rr] : 0.0;
            ld_m += l; qd_m += zz;
            if (rr < n) { ld_n += l; qd_n += zz; }
        }
        ld_n = warp_sum(ld_n); ld_m = warp_sum(ld_m); qd_n = warp_sum(qd_n); qd_m = warp_sum(qd_m);
        if (lane == 0) { s_red[0 * kWS + role] = ld_n; s_red[1 * kWS + role] = ld_m; s_red[2 * kWS + role] = qd_n; s_red[3 * kWS + role] = qd_m; }
        slotbar();
        if (stid == 0) {
            double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
            for (int w = 0; w < kWS; ++w) { r0 += s_red[w]; r1 += s_red[kWS + w]; r2 += s_red[2 * kWS + w]; r3 += s_red[3 * kWS + w]; }
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
            for (int rr = role; rr < h; rr += kWS) {
                double accv = 0.0;
                for (int cix = lane; cix < m; cix += 32) accv = fma(Lel(m + rr, cix), z[cix], accv);
                accv = warp_sum(accv);
                if (lane == 0) a.mu[b * h + rr] = (accv - a.yb) / a.ya;
            }
        }
        if (a.L33) {
            for (int e = stid; e < h * h; e += kST) {
                int rr = e / h, cix = e - rr * h;
                a.L33[b * h * h + e] = cix <= rr ? Lel(m + rr, m + cix) / a.ya : 0.0;
            }
        }
        if (a.proj) {
            for (int rr = role; rr < kh; rr += kWS) {
                double accv = 0.0;
                for (int cix = lane; cix < n; cix += 32) accv = fma(Lel(n + rr, cix), z[cix], accv);
                accv = warp_sum(accv);
                if (lane == 0) a.proj[b * kh + rr] = accv;
            }
        }
        if (a.Ltail) {
            for (int e = stid; e < kh * kh; e += kST) {
                int rr = e / kh, cix = e - rr * kh;
                a.Ltail[b * kh * kh + e] = cix <= rr ? Lel(n + rr, n + cix) : 0.0;
            }
        }
    }
}

}  // namespace

// ---- host side: slot map, shared-memory plan, launch ------------------------------------------------------------

// Static slot map of the recycled tile pool and the producer's schedule. Tile (I, P), P <= I, is born when the Gram
// producer writes column P and dies when the row owners have finished step max(I - 2, P): step I - 2 is the last one
// that reads row I's tiles P <= I - 2 (lookahead of column I), step I - 1 reads tile (I, I - 1), step I the diagonal
// tile. Rows >= iep feed the epilogue and never die. Given a pool of `nslots`, column C is scheduled as early as the
// pool allows: need[C] = the smallest number of completed row-owner steps after which nt - C slots are free (columns
// are produced in order, so need is non-decreasing). The lookahead of step C - 2 parks its partial sums over the Gram
// tiles of column C, so the schedule is only valid if need[C] <= C - 2 (the producer at least two columns ahead).
// Returns false when the pool is too small for that.
static bool build_slot_map(int nt, int nslots, int iep, unsigned char *smap, unsigned char *need)
{
    if (nslots > 256) nslots = 256;
    std::vector<int> live_slot, live_death, freel;
    for (int sl = nslots - 1; sl >= 0; --sl) freel.push_back(sl);
    int nd = 0;
    for (int C = 0; C < nt; ++C) {
        for (;;) {
            for (size_t i = 0; i < live_slot.size();) {
                if (live_death[i] < nd) {
                    freel.push_back(live_slot[i]);
                    live_slot[i] = live_slot.back(); live_slot.pop_back();
                    live_death[i] = live_death.back(); live_death.pop_back();
                } else ++i;
            }
            if ((int)freel.size() >= nt - C) break;
            if (++nd > std::max(0, C - 2)) return false;
        }
        need[C] = (unsigned char)nd;
        std::sort(freel.begin(), freel.end(), [](int x, int y) { return x > y; });
        for (int I = C; I < nt; ++I) {
            const int sl = freel.back();
            freel.pop_back();
            smap[I * (I + 1) / 2 + C] = (unsigned char)sl;
            live_slot.push_back(sl);
            live_death.push_back(I >= iep ? (1 << 30) : std::max(I - 2, C));
        }
    }
    return true;
}

int fused_v3_max_q() { return 8 * kMaxNt3; }

V3Plan plan_fused_v3(int n, int k, int h, bool tail_rows_from_n, int G, int ntheta_cap, int ntab_cap, int ncp_cap, int smem_optin)
{
    V3Plan pl;
    std::memset(&pl, 0, sizeof(pl));
    const int q = n + k + h;
    const int nt = (q + 7) / 8, Q = nt * 8;
    if (nt < 3 || nt > kMaxNt3) return pl;
    pl.nt = nt;
    pl.iep = (tail_rows_from_n ? n : n + k) / 8;
    auto up = [](size_t x) { return (x + 15) & ~size_t(15); };
    size_t off = 0;
    pl.off_tt = (int)off; off += up((size_t)Q * sizeof(double));
    pl.off_gg = (int)off; off += up((size_t)Q * sizeof(int));
    pl.off_map = (int)off; off += up(sizeof(pl.smap));
    pl.off_need = (int)off; off += up(sizeof(pl.need));
    pl.off_slots = (int)off;
    size_t so = 0;
    pl.o_ctrl = (int)so; so += kCtrlBytes;
    pl.o_tp = (int)so; so += up(sizeof(TreeProgram));
    pl.o_yv = (int)so; so += up((size_t)Q * sizeof(double));
    pl.o_diag = (int)so; so += up((size_t)Q * sizeof(double));
    pl.o_inv = (int)so; so += (size_t)kInvBufs * 512;
    pl.o_th = (int)so; so += up((size_t)std::max(ntheta_cap, 1) * sizeof(double));
    pl.o_tab = (int)so; so += up((size_t)ntab_cap * (G > 0 ? G : 0) * sizeof(double));
    pl.o_sig = (int)so; so += up((size_t)ncp_cap * Q * sizeof(double));
    pl.o_pool = (int)so;
    cudaFuncAttributes fa{};
    size_t static_smem = 256;
    if (cudaFuncGetAttributes(&fa, fused_v3_kernel) == cudaSuccess) static_smem = fa.sharedSizeBytes;
    else cudaGetLastError();
    const size_t budget = (size_t)smem_optin - static_smem;
    // every tile slot the opt-in shared memory leaves room for goes to the pools (more slots = the producer further ahead)
    if (off + (size_t)kM3 * so >= budget) return pl;
    int ns = (int)((budget - off) / kM3 - so) / 512;
    ns = std::min(ns, nt * (nt + 1) / 2);
    if (ns < 1 || !build_slot_map(nt, ns, pl.iep, pl.smap, pl.need)) return pl;
    pl.ns = ns;
    pl.lead = nt;
    for (int C = 2; C < nt; ++C) pl.lead = std::min(pl.lead, C - (int)pl.need[C]);     // the tightest column
    pl.slot_stride = (int)(so + (size_t)ns * 512);
    pl.smem_bytes = off + (size_t)kM3 * pl.slot_stride;
    pl.ok = 1;
    return pl;
}

int fused_v3_grid(int64_t B, int num_sms) { return (int)std::min<int64_t>(num_sms, B); }

cudaError_t launch_fused_v3(const FusedArgs &a, const V3Plan &pl, unsigned long long *work_counter, int grid,
                            cudaStream_t stream)
{
    cudaError_t e0 = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), stream);
    if (e0 != cudaSuccess) return e0;
    cudaError_t e = cudaFuncSetAttribute(fused_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) return e;
    fused_v3_kernel<<<grid, kT3, pl.smem_bytes, stream>>>(a, pl, work_counter);
    return cudaGetLastError();
}

}  // namespace nagp

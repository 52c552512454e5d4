// nagp_large.cu — factorisation beyond shared-memory size (q up to 4096) and the in-place rank-append.
//
//  chol_large_kernel   one CTA per instance, persistent grid with a dynamic instance queue. Blocked
//                      left-looking Cholesky by block columns of kCB = 4 tiles (32 columns): the factor lives
//                      in HBM/L2 as 8x8 FP64 tiles in DMMA operand order (tile-packed lower triangle,
//                      plus one tile row for z = L^-1 y), Gram tiles are evaluated on the fly by the
//                      kernel-tree interpreter and never stored. Per block column the warps pull rows from a
//                      shared queue: a row (two rows at a time from n = 768 up) accumulates sum_P L_IP L_JP^T
//                      for the block's column tiles in registers (DMMA; one 16-byte fragment load per operand
//                      tile), the rows of the diagonal block go to shared memory where one warp factors the
//                      32x32 block (8x8 in-register Cholesky + inverse per tile) while the others run ahead,
//                      every other row finishes with an in-register right-looking triangular solve against
//                      that block and writes its tiles.
//  rank_append_kernel  one CTA per particle: extends a stored factor by new rows in place (up-looking):
//                      streams the existing L once in storage order, split-K over the warps.
//
// Replaces, for long series, AutoGP's Gram + dpotrf per likelihood evaluation behind
//   /root/reference/src/make_and_fit_model.jl:84-91 (GPModel + fit_smc! data annealing: add a batch of
//   observations, re-score every particle) and /root/reference/src/forecasting.jl:133,135,46.
// Arithmetic contract: docs/KERNEL_SPEC.md §3-§6.
#include <algorithm>

#include "nagp_kernels.cuh"
#include "nagp_tree.cuh"
#include "nagp_tile.cuh"

#ifndef NAGP_LARGE_TRACE
#define NAGP_LARGE_TRACE 0  // n: clock64 stamps of the n-th instance of block 0 for tools/large_timeline.py; 0 in product builds
#endif
#if NAGP_LARGE_TRACE
// [block column][warp][arrival at the end, cycles waited for the diagonal block, row groups solved, -], then
// [block column][start, diagonal block begins, ends, -]
__device__ long long g_large_trace[128 * 8 * 4 + 128 * 4];
extern "C" int nagp_debug_read_large(long long *out, int count)
{
    return (int)cudaMemcpyFromSymbol(out, g_large_trace, sizeof(long long) * count);
}
#endif

namespace nagp {

namespace {

constexpr int kBlk = 8;   // tile rows per storage block / rank-append group
#ifndef NAGP_LARGE_CB
#define NAGP_LARGE_CB 4     // measured (1024 x n = 512 / 256 x n = 2048, one row per warp): 8 -> 4.38 / 49.2 ms, 4 -> 4.20 / 46.1 ms
#endif
constexpr int kCB = NAGP_LARGE_CB;                // tile columns per block column of chol_large_kernel
constexpr int kCBT = kCB * (kCB + 1) / 2;
// Rows below the diagonal block that a warp sums at a time against the same panel tiles (one panel load per kRows DMMA
// pairs; the L1 hit rate of those loads was 69 % and the L2 hit rate 53 %): a template parameter of the kernel. Two
// rows pay from n = 768 up (512 x n = 768 / 1024 / 1536: 6.19 / 13.7 / 41.8 ms with one row, 5.79 / 12.3 / 37.1 ms with
// two; 256 x n = 2048: 46.1 -> 43.3 ms) and cost at n = 512 (4.20 -> 4.40 ms); four rows, or two at kCB = 8, spill.
#ifndef NAGP_LARGE_DIAG_UNROLL
#define NAGP_LARGE_DIAG_UNROLL 4
#endif
constexpr int kDiagUnroll = NAGP_LARGE_DIAG_UNROLL;
// unroll depth of the accumulate loops (terms whose operand loads are issued together): measured 1024 x n = 512 /
// 512 x n = 1024 / 256 x n = 2048 at depth 2: 3.79 / 11.7 / 37.0 ms, 3: 3.73 / 12.2 / 45.1, 4: 3.71 / 11.5 / 40.0
#ifndef NAGP_LARGE_ROWS_FROM
#define NAGP_LARGE_ROWS_FROM 96     // tile rows (n >= 768); final build, one / two rows: n = 512 3.59 / 3.74 ms, 640: 3.54 / 3.67, 768: 5.42 / 5.40, 1024: 12.7 / 11.6, 2048: 44.4 / 39.0
#endif
constexpr int kRowsLargeFrom = NAGP_LARGE_ROWS_FROM;                // tile rows (n >= 640): two rows per warp
static_assert(kBlk % kCB == 0, "block columns must tile the storage blocks");
static_assert(kWarps == 2 * kCB, "the rows of a diagonal block are summed by two warps each");

__device__ __forceinline__ double2 ldg128(const double *p)
{
    return *reinterpret_cast<const double2 *>(p);
}
__device__ __forceinline__ double2 ldg128_stream(const double *p)
{
    return __ldcg(reinterpret_cast<const double2 *>(p));
}

// Rank-append streaming ring: stages of 8 tiles (4 KB) per warp that land in shared memory (cp.async, L2 only) instead
// of registers, so a warp keeps (stages - 1) x 4 KB of the stored row in flight while it consumes one stage.
// 0: the register-landed loop (4 KB in flight per warp, and only between its compute phases).
#ifndef NAGP_APPEND_STAGES
#define NAGP_APPEND_STAGES 3
#endif
constexpr int kApStages = NAGP_APPEND_STAGES;
__device__ __forceinline__ void cp_async16(uint32_t dst, const double *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// accumulator layout (lane (r, j) holds (r, 2j), (r, 2j+1)) -> A-operand fragment (lane (r, kk) holds
// (r, kk) and (r, kk + 4))
__device__ __forceinline__ double2 acc_to_frag(double c0, double c1, int lane)
{
    const int lj = lane & 3;
    const int cv0 = (lane & ~3) + (lj >> 1), cv1 = cv0 + 2;
    const bool odd = lane & 1;
    const double v00 = shfl(c0, cv0), v01 = shfl(c1, cv0);
    const double v10 = shfl(c0, cv1), v11 = shfl(c1, cv1);
    return make_double2(odd ? v01 : v00, odd ? v11 : v10);
}

// store an accumulator-layout tile in operand layout
__device__ __forceinline__ void store_op(double *tile, double c0, double c1, int lane)
{
    const int lr = lane >> 2, lj = lane & 3;
    tile[op_idx(lr, 2 * lj)] = c0;
    tile[op_idx(lr, 2 * lj + 1)] = c1;
}

struct LargeLayout {
    int nt;              // tile rows holding real points: ceil(q / 8)
    int ntp;             // nt rounded up to a multiple of kBlk
    int yrow;            // tile row index of the z = L^-1 y row (>= ntp)
    int aux_off[5];      // th, gg, tt, sig, tab
    int aux_smem[5];
    int scratch_stride;
    char *scratch;
    double *L;           // factor storage
    size_t L_stride;     // doubles per slot
    int keep;            // 1: slot = instance b (factor kept), 0: slot = blockIdx.x (workspace)
    double *W;           // [slot][ntp_cap] inverse diagonal tiles (operand layout), nullable
    size_t W_stride;
    unsigned long long *work_counter;
    // block-append mode of chol_large_kernel: tile rows >= i0 are (re)computed against the stored rows above them, the
    // diagonal blocks above are loaded instead of factored; the increment of logML over the points n_old .. n-1 goes to
    // dlogml and is added to logml_acc
    int i0, n_old;
    double *dlogml, *logml_acc;
};

struct Setup {
    double *th; int *gg; double *tt; double *sig; double *tab;
};

__device__ __forceinline__ Setup aux_pointers(const LargeLayout &lay, char *aux_s)
{
    char *aux_g = lay.scratch + (size_t)blockIdx.x * lay.scratch_stride;
    Setup s;
    auto aux = [&](int i) { return (lay.aux_smem[i] ? aux_s : aux_g) + lay.aux_off[i]; };
    s.th = reinterpret_cast<double *>(aux(0));
    s.gg = reinterpret_cast<int *>(aux(1));
    s.tt = reinterpret_cast<double *>(aux(2));
    s.sig = reinterpret_cast<double *>(aux(3));
    s.tab = reinterpret_cast<double *>(aux(4));
    return s;
}

// Program compile + theta + lag / sigma tables for one instance (all threads; ends with a barrier).
__device__ __forceinline__ void instance_setup(TreeProgram &tp, const FusedArgs &a, const Setup &su, int64_t s, int p,
                                               int Q, int tid)
{
    const int64_t po = a.prog_off[p], plen = a.prog_off[p + 1] - po;
    const int64_t to = a.theta_off[p], ntheta = a.theta_off[p + 1] - to;
    const double *theta_g = a.theta + s * a.theta_stride_k + to;
    if (tid == 0) {
        if (ntheta > MAX_THETA) tp.error = -3;
        else tree_compile(tp, a.prog + po, (int)plen, (int)ntheta, a.G > 0 ? a.ntab_cap : 0, a.ncp_cap);
    }
    for (int i = tid; i < ntheta && i < MAX_THETA; i += kThreads) su.th[i] = theta_g[i];
    __syncthreads();
    if (tp.error) return;
    const int G = a.G;
    for (int e = tid; e < tp.ntab * G; e += kThreads) {
        int id = e / G, lg = e - id * G;
        int s0 = tp.tab_src0[id], s1 = tp.tab_src1[id];
        su.tab[e] = tree_eval(tp.sop + s0, tp.sarg + s0, nullptr, s1 - s0, su.th, 0.0, 0.0,
                              (double)lg * a.step, 0, nullptr, 0, nullptr, 0, 0, 0);
    }
    for (int e = tid; e < tp.ncp * Q; e += kThreads) {
        int id = e / Q, i = e - id * Q;
        const double *cp = su.th + tp.cp_theta[id];
        su.sig[e] = 0.5 * (1.0 + tanh((su.tt[i] - cp[0]) / cp[1]));
    }
    __syncthreads();
}

struct GramCtx {
    EvalCtx cx;
    const int *gg;
    int q, m;
    double d_lo, d_hi;
    bool single_table;
};

// Gram tiles (I, Ja) and (I, Jb) in accumulator layout: out[0..1] of the first, out[2..3] of the second.
// Rows/columns >= q are identity padding.
__device__ __forceinline__ void gram_pair(const TreeProgram &tp, const GramCtx &gc, int I, int Ja, int Jb, int lane,
                                          double (&out)[4])
{
    const int gr = lane >> 2, gcn = (lane & 3) * 2;
    int ii[4], jj[4], lag[4];
    bool real[4];
    const int gi = gc.gg[I * 8 + gr];
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
        const int J = h2 ? Jb : Ja;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int x = h2 * 2 + e;
            ii[x] = I * 8 + gr;
            jj[x] = J * 8 + gcn + e;
            const int lg = gi - gc.gg[jj[x]];
            lag[x] = lg < 0 ? -lg : lg;
            real[x] = ii[x] < gc.q && jj[x] < gc.q;
        }
    }
    if (gc.single_table) {
#pragma unroll
        for (int x = 0; x < 4; ++x) out[x] = gc.cx.tab[lag[x]];
    } else {
        tree_eval4(tp, gc.cx, ii, jj, lag, out);
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        if (!real[x]) out[x] = (ii[x] == jj[x]) ? 1.0 : 0.0;
        else if (ii[x] == jj[x]) out[x] += (ii[x] < gc.m) ? gc.d_lo : gc.d_hi;
    }
}

// observation values of columns [J*8, J*8+8) as the row-0 entries of an accumulator-layout tile
__device__ __forceinline__ void y_tile(const FusedArgs &a, int64_t b, int64_t s, int J, int ny, int lane, double &c0, double &c1)
{
    c0 = 0.0; c1 = 0.0;
    if ((lane >> 2) != 0) return;
    const double *y1 = a.y1 + b * a.y1_stride;
    const int n = a.n, k = a.k;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int jx = J * 8 + (lane & 3) * 2 + e;
        double v = 0.0;
        if (jx < n) v = y1[jx];
        else if (jx < ny) v = a.y2 ? a.y2[s * k + (jx - n)] : y1[jx];
        if (e) c1 = v; else c0 = v;
    }
}

// Factor the 64x64 diagonal block held in s_C (lower tiles, accumulator/row-major layout) by ONE warp:
// right-looking over its 8 tile columns. Writes L tiles (operand layout) to s_L and to the factor, inverse
// diagonal tiles to s_W (and the factor's W store). Returns 0 or the 1-based index of the first bad pivot.
__device__ __forceinline__ int factor_diag_block(double *s_C, double *s_L, double *s_W, double *Lb, double *Wb, int c0,
                                                 int q, int lane)
{
    int info = 0;
    for (int j = 0; j < kCB; ++j) {
        const int Jg = c0 + j;
        double2 cj = *reinterpret_cast<double2 *>(s_C + (tri(j) + j) * 64 + lane * 2);
        double d0 = cj.x, d1 = cj.y, w0, w1, piv[8];
        const int bad = chol8_inv(d0, d1, w0, w1, lane, q - Jg * 8, piv);
        if (bad && !info) info = Jg * 8 + bad;
        double *ltile = s_L + (tri(j) + j) * 64;
        store_op(ltile, d0, d1, lane);
        store_op(Lb + ((size_t)tri(Jg) + Jg) * 64, d0, d1, lane);
        store_op(s_W + j * 64, w0, w1, lane);
        if (Wb) store_op(Wb + (size_t)Jg * 64, w0, w1, lane);
        __syncwarp();
        const double2 ib = *reinterpret_cast<double2 *>(s_W + j * 64 + lane * 2);
        // column trsm: X_aj = C_aj W_jj^T
        for (int a = j + 1; a < kCB; ++a) {
            const double2 c = *reinterpret_cast<double2 *>(s_C + (tri(a) + j) * 64 + lane * 2);
            const double2 fr = acc_to_frag(c.x, c.y, lane);
            double x0 = 0.0, x1 = 0.0;
            dmma(x0, x1, fr.x, ib.x);
            dmma(x0, x1, fr.y, ib.y);
            store_op(s_L + (tri(a) + j) * 64, x0, x1, lane);
            store_op(Lb + ((size_t)tri(c0 + a) + Jg) * 64, x0, x1, lane);
        }
        __syncwarp();
        // trailing update inside the block: C_ab -= X_aj X_bj^T
        for (int a = j + 1; a < kCB; ++a) {
            double2 af = *reinterpret_cast<double2 *>(s_L + (tri(a) + j) * 64 + lane * 2);
            af.x = -af.x; af.y = -af.y;
            for (int b = j + 1; b <= a; ++b) {
                const double2 bf = *reinterpret_cast<double2 *>(s_L + (tri(b) + j) * 64 + lane * 2);
                double2 *cp = reinterpret_cast<double2 *>(s_C + (tri(a) + b) * 64 + lane * 2);
                double2 c = *cp;
                dmma(c.x, c.y, af.x, bf.x);
                dmma(c.x, c.y, af.y, bf.y);
                *cp = c;
            }
        }
        __syncwarp();
    }
    return info;
}

template <int kRows>
__global__ void __launch_bounds__(kThreads, 2) chol_large_kernel(const FusedArgs a, const LargeLayout lay)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ TreeProgram tp;
    __shared__ int s_info;
    __shared__ long long s_next;
    __shared__ int s_queue, s_diagdone;
    __shared__ volatile int s_flag;
    __shared__ double s_red[4][kWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, k = a.k, h = a.h, m = n + k, q = m + h;
    const int ntp = lay.ntp, Q = ntp * 8, yrow = lay.yrow;
    const bool have_y2 = (a.y2 != nullptr) || k == 0;
    const int ny = have_y2 ? m : n;

    double *s_C = smem;                    // tri(kCB) tiles, accumulator layout
    double *s_L = s_C + kCBT * 64;         // tri(kCB) tiles, operand layout
    double *s_W = s_L + kCBT * 64;         // kCB tiles, operand layout
    double *s_P = s_W + kCB * 64;          // tri(kCB) tiles: partial sums of the diagonal block's rows (second half of the terms)
    char *aux_s = reinterpret_cast<char *>(s_P + kCBT * 64);
    const Setup su = aux_pointers(lay, aux_s);

    for (int i = tid; i < Q; i += kThreads) {
        su.tt[i] = i < q ? a.t[i] : 0.0;
        su.gg[i] = (a.g && i < q) ? a.g[i] : 0;
    }

#if NAGP_LARGE_TRACE
    int n_inst = 0;
#endif
    for (;;) {
        __syncthreads();
        if (tid == 0) s_next = (long long)atomicAdd(lay.work_counter, 1ull);
        __syncthreads();
        const int64_t b = s_next;
        if (b >= a.B) break;
        const int64_t s = b / a.P;
        const int p = (int)(b % a.P);
        const size_t slot = lay.keep ? (size_t)b : (size_t)blockIdx.x;
        double *Lb = lay.L + slot * lay.L_stride;
        double *Wb = lay.W ? lay.W + slot * lay.W_stride : nullptr;

        if (tid == 0) { s_info = 0; s_flag = 0; s_queue = 0; s_diagdone = 0; }
        instance_setup(tp, a, su, s, p, Q, tid);
        if (tp.error) {
            if (tid == 0) {
                a.info[b] = tp.error;
                if (a.logml_n) a.logml_n[b] = nan("");
                if (a.logml_m) a.logml_m[b] = nan("");
                if (a.logw) a.logw[b] = nan("");
            }
            continue;
        }
        GramCtx gc;
        gc.cx.th = su.th; gc.cx.tt = su.tt; gc.cx.tab = su.tab; gc.cx.sig = su.sig;
        gc.cx.step = a.step; gc.cx.G = a.G; gc.cx.Q = Q; gc.cx.grid = a.g != nullptr;
        gc.gg = su.gg; gc.q = q; gc.m = m;
        {
            const double nz = a.noise[s * a.noise_stride_k + p];
            gc.d_lo = nz + a.jitter;
            gc.d_hi = (a.noise_pred >= 0.0 ? a.noise_pred : nz) + a.jitter;
        }
        gc.single_table = (tp.clen == 1 && tp.cop[0] == OP_TABLE);

        const int nbc = ntp / kCB;
#if NAGP_LARGE_TRACE
        const bool tracing = (blockIdx.x == 0 && ++n_inst == NAGP_LARGE_TRACE && nbc <= 128);
#endif
        for (int Jb = 0; Jb < nbc; ++Jb) {
            const int c0 = Jb * kCB;
#if NAGP_LARGE_TRACE
            long long waited = 0; int items_done = 0;
            if (tracing && tid == 0) g_large_trace[128 * 8 * 4 + Jb * 4] = clock64();
#endif
            // items: the kCB rows of the diagonal block, the rows below in groups of kRows (a warp sums kRows rows against
            // the same panel tiles: one operand load per kRows DMMA pairs), the y row
            const int ngroups = (ntp - c0 - kCB) / kRows;
            // append mode: a block column above the first new tile row has its diagonal block in the store (loaded, not
            // factored), no z tiles to compute, and only the row groups that reach the new rows to sum
            const bool old_block = c0 + kCB <= lay.i0;
            const int gfirst = old_block ? (lay.i0 - c0 - kCB) / kRows : 0;
            const int nitems = kCB + ngroups + (old_block ? 0 : 1);
            // The rows of the diagonal block come first and on all eight warps (each row's sum split in two): everything
            // else ends up waiting for them and for the factorisation that follows, and warps that stream rows below at the
            // same time take three quarters of the FP64 tensor pipe away from them (trace: 52 terms at 780 cycles each).
            bool diag_phase = !old_block;
            if (old_block) {
                for (int i = tid; i < kCBT * 64; i += kThreads) {
                    const int tl = i >> 6, e = i & 63;
                    int ra = 0;
                    while (tri(ra + 1) <= tl) ++ra;                 // tile tl of the packed block is (ra, tl - tri(ra))
                    s_L[i] = Lb[((size_t)tri(c0 + ra) + c0 + (tl - tri(ra))) * 64 + e];
                }
                for (int i = tid; i < kCB * 64; i += kThreads) s_W[i] = Wb[(size_t)c0 * 64 + i];
                if (tid == 0) { s_queue = gfirst; s_flag = Jb + 1; }
                __syncthreads();
            }
            for (;;) {
                int item = kCB - 1 - (warp % kCB), Plo = 0, Phi = c0;
                const int half = diag_phase ? warp / kCB : 0;
                if (diag_phase) {
                    const int mid = c0 >> 1;
                    Plo = half ? mid : 0; Phi = half ? c0 : mid;
                } else {
                    if (lane == 0) item = kCB + atomicAdd(&s_queue, 1);
                    item = __shfl_sync(kFull, item, 0);
                    if (item >= nitems) break;
                }
                const bool diag = item < kCB;
                const int arow = kCB - 1 - item;                       // longest diagonal rows first
                const bool is_y = !old_block && (item == nitems - 1);
                const int I = diag ? c0 + arow : (is_y ? yrow : c0 + kCB + (item - kCB) * kRows);
                const int NC = diag ? arow + 1 : kCB;
                const int NR = (diag || is_y) ? 1 : kRows;              // rows I .. I + NR - 1

                // ---- sum_{P < c0} L_IP L_{c0+b,P}^T for the NC column tiles ------------------------
                double acc[kRows][kCB][2];
#pragma unroll
                for (int r = 0; r < kRows; ++r)
#pragma unroll
                    for (int bb = 0; bb < kCB; ++bb) { acc[r][bb][0] = 0.0; acc[r][bb][1] = 0.0; }
                {
                    const double *arowp = Lb + (size_t)tri(I) * 64 + lane * 2;
                    const double *browp = Lb + (size_t)tri(c0) * 64 + lane * 2;
                    if (NR == kRows && kRows > 1) {
#pragma unroll 2
                        for (int P = Plo; P < Phi; ++P) {
                            double2 af[kRows];
#pragma unroll
                            for (int r = 0; r < kRows; ++r)      // row I + r starts r * I + tri(r) tiles after row I
                                af[r] = ldg128_stream(arowp + ((size_t)(r * I + ((r * (r + 1)) >> 1)) + P) * 64);
#pragma unroll
                            for (int bb = 0; bb < kCB; ++bb) {
                                // tile (c0+bb, P) sits tri(c0+bb) - tri(c0) = bb*c0 + tri(bb) tiles after (c0, P)
                                const double2 bf = ldg128(browp + ((size_t)(bb * c0 + ((bb * (bb + 1)) >> 1)) + P) * 64);
#pragma unroll
                                for (int r = 0; r < kRows; ++r) {
                                    dmma(acc[r][bb][0], acc[r][bb][1], af[r].x, bf.x);
                                    dmma(acc[r][bb][0], acc[r][bb][1], af[r].y, bf.y);
                                }
                            }
                        }
                    } else if (NC == kCB) {
#pragma unroll 4
                        for (int P = Plo; P < Phi; ++P) {
                            const double2 af = ldg128_stream(arowp + (size_t)P * 64);
#pragma unroll
                            for (int bb = 0; bb < kCB; ++bb) {
                                const double2 bf = ldg128(browp + ((size_t)(bb * c0 + ((bb * (bb + 1)) >> 1)) + P) * 64);
                                dmma(acc[0][bb][0], acc[0][bb][1], af.x, bf.x);
                                dmma(acc[0][bb][0], acc[0][bb][1], af.y, bf.y);
                            }
                        }
                    } else {
                        // the rows of the diagonal block: everything else ends up waiting for them (and for the
                        // factorisation that follows), so their operand loads are issued four terms ahead
#pragma unroll kDiagUnroll
                        for (int P = Plo; P < Phi; ++P) {
                            const double2 af = ldg128_stream(arowp + (size_t)P * 64);
#pragma unroll
                            for (int bb = 0; bb < kCB; ++bb) {
                                if (bb < NC) {
                                    const double2 bf = ldg128(browp + ((size_t)(bb * c0 + ((bb * (bb + 1)) >> 1)) + P) * 64);
                                    dmma(acc[0][bb][0], acc[0][bb][1], af.x, bf.x);
                                    dmma(acc[0][bb][0], acc[0][bb][1], af.y, bf.y);
                                }
                            }
                        }
                    }
                }
                // ---- C_b = A(I, c0+b) - acc_b --------------------------------------------------------
                if (is_y) {
#pragma unroll
                    for (int bb = 0; bb < kCB; ++bb) {
                        double y0, y1v;
                        y_tile(a, b, s, c0 + bb, ny, lane, y0, y1v);
                        acc[0][bb][0] = y0 - acc[0][bb][0];
                        acc[0][bb][1] = y1v - acc[0][bb][1];
                    }
                } else if (diag && half) {
#pragma unroll
                    for (int bb = 0; bb < kCB; ++bb) { acc[0][bb][0] = -acc[0][bb][0]; acc[0][bb][1] = -acc[0][bb][1]; }
                } else {
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        if (r < NR) {
#pragma unroll
                            for (int bb = 0; bb < kCB; bb += 2) {
                                if (bb < NC) {
                                    double out[4];
                                    gram_pair(tp, gc, I + r, c0 + bb, c0 + bb + 1, lane, out);
                                    acc[r][bb][0] = out[0] - acc[r][bb][0];
                                    acc[r][bb][1] = out[1] - acc[r][bb][1];
                                    acc[r][bb + 1][0] = out[2] - acc[r][bb + 1][0];
                                    acc[r][bb + 1][1] = out[3] - acc[r][bb + 1][1];
                                }
                            }
                        }
                    }
                }
                if (diag) {
                    // acc holds A - (first half of the terms) on the warps with half == 0 and -(second half) on the others
                    if (half) {
#pragma unroll
                        for (int bb = 0; bb < kCB; ++bb)
                            if (bb < NC)
                                *reinterpret_cast<double2 *>(s_P + (tri(arow) + bb) * 64 + lane * 2) = make_double2(acc[0][bb][0], acc[0][bb][1]);
                    }
                    __syncthreads();
                    diag_phase = false;
                    if (half) continue;
#pragma unroll
                    for (int bb = 0; bb < kCB; ++bb) {
                        if (bb < NC) {
                            const double2 pp = *reinterpret_cast<const double2 *>(s_P + (tri(arow) + bb) * 64 + lane * 2);
                            acc[0][bb][0] += pp.x;
                            acc[0][bb][1] += pp.y;
                        }
                    }
#pragma unroll
                    for (int bb = 0; bb < kCB; ++bb)
                        if (bb < NC)
                            *reinterpret_cast<double2 *>(s_C + (tri(arow) + bb) * 64 + lane * 2) = make_double2(acc[0][bb][0], acc[0][bb][1]);
                    __threadfence_block();
                    __syncwarp();
                    int done = 0;
                    if (lane == 0) done = atomicAdd(&s_diagdone, 1);
                    done = __shfl_sync(kFull, done, 0);
                    if (done == kCB - 1) {
                        // last diagonal row in: this warp factors the block while the others run ahead
                        __threadfence_block();
#if NAGP_LARGE_TRACE
                        if (tracing && lane == 0) g_large_trace[128 * 8 * 4 + Jb * 4 + 1] = clock64();
#endif
                        const int bad = factor_diag_block(s_C, s_L, s_W, Lb, Wb, c0, q, lane);
#if NAGP_LARGE_TRACE
                        if (tracing && lane == 0) g_large_trace[128 * 8 * 4 + Jb * 4 + 2] = clock64();
#endif
                        if (lane == 0 && bad && !s_info) s_info = bad;
                        __threadfence_block();
                        __syncwarp();
                        if (lane == 0) s_flag = Jb + 1;
                    }
                    continue;
                }
                // ---- wait for the diagonal block, then the in-register triangular solve -------------------
#if NAGP_LARGE_TRACE
                const long long tw0 = clock64();
#endif
                while (s_flag < Jb + 1) __nanosleep(64);
#if NAGP_LARGE_TRACE
                waited += clock64() - tw0; ++items_done;
#endif
                __threadfence_block();
                __syncwarp();
                // right-looking: a solved tile goes into every column tile still open at once (independent DMMA chains, same
                // order of terms per tile as a left-looking sum); the rows of a group are independent chains too
#pragma unroll
                for (int aa = 0; aa < kCB; ++aa) {
                    const double2 ib = *reinterpret_cast<const double2 *>(s_W + aa * 64 + lane * 2);
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        if (r < NR) {
                            const double2 fr = acc_to_frag(acc[r][aa][0], acc[r][aa][1], lane);
                            double x0 = 0.0, x1 = 0.0;
                            dmma(x0, x1, fr.x, ib.x);
                            dmma(x0, x1, fr.y, ib.y);
                            if (I + r >= lay.i0) store_op(Lb + ((size_t)tri(I + r) + c0 + aa) * 64, x0, x1, lane);   // (append mode: a stored row that shares a group with a new one stays as it is)
                            if (aa + 1 < kCB) {
                                const double2 xf = acc_to_frag(x0, x1, lane);
#pragma unroll
                                for (int bb = aa + 1; bb < kCB; ++bb) {
                                    const double2 bf = *reinterpret_cast<const double2 *>(s_L + (tri(bb) + aa) * 64 + lane * 2);
                                    dmma(acc[r][bb][0], acc[r][bb][1], -xf.x, bf.x);
                                    dmma(acc[r][bb][0], acc[r][bb][1], -xf.y, bf.y);
                                }
                            }
                        }
                    }
                }
            }
#if NAGP_LARGE_TRACE
            if (tracing && lane == 0) {
                long long *tr = g_large_trace + (Jb * 8 + warp) * 4;
                tr[0] = clock64(); tr[1] = waited; tr[2] = items_done;
            }
#endif
            __syncthreads();
            if (tid == 0) { s_queue = 0; s_diagdone = 0; }
            __syncthreads();
            if (s_info) break;
        }

        if (s_info) {
            if (tid == 0) {
                a.info[b] = s_info;
                if (lay.dlogml) { lay.dlogml[b] = nan(""); if (lay.logml_acc) lay.logml_acc[b] = nan(""); }
                if (a.logml_n) a.logml_n[b] = nan("");
                if (a.logml_m) a.logml_m[b] = nan("");
                if (a.logw) a.logw[b] = nan("");
            }
            continue;
        }

        auto Lel = [&](int i, int j) {
            return Lb[((size_t)tri(i >> 3) + (j >> 3)) * 64 + op_idx(i & 7, j & 7)];
        };
        auto zel = [&](int r) { return Lb[((size_t)tri(yrow) + (r >> 3)) * 64 + op_idx(0, r & 7)]; };

        // ---- logML(n), logML(m) ----------------------------------------------------------------------
        // (the logs of the diagonal are taken here by all threads: inside the one-warp factorisation of the diagonal block
        // the other warps were waiting for them)
        double qd_n = 0, qd_m = 0, ld_n = 0, ld_m = 0;
        for (int r = tid; r < m; r += kThreads) {
            const double l = log(Lel(r, r));
            ld_m += l;
            if (r < n) ld_n += l;
            if (r < ny) {
                const double zz = zel(r);
                qd_m = fma(zz, zz, qd_m);
                if (r < n) qd_n = fma(zz, zz, qd_n);
            }
        }
        qd_n = warp_sum(qd_n); qd_m = warp_sum(qd_m); ld_n = warp_sum(ld_n); ld_m = warp_sum(ld_m);
        if (lane == 0) { s_red[0][warp] = qd_n; s_red[1][warp] = qd_m; s_red[2][warp] = ld_n; s_red[3][warp] = ld_m; }
        __syncthreads();
        if (lay.dlogml) {
            // block-append mode: the increment over the new points only
            __syncthreads();
            double ld = 0.0, qd = 0.0;
            for (int r = lay.n_old + tid; r < n; r += kThreads) {
                ld += log(Lel(r, r));
                const double zz = zel(r);
                qd = fma(zz, zz, qd);
            }
            ld = warp_sum(ld); qd = warp_sum(qd);
            if (lane == 0) { s_red[0][warp] = ld; s_red[1][warp] = qd; }
            __syncthreads();
            if (tid == 0) {
                double l0 = 0, q0 = 0;
                for (int w = 0; w < kWarps; ++w) { l0 += s_red[0][w]; q0 += s_red[1][w]; }
                const double log2pi = 1.8378770664093454835606594728112;
                const double dl = -0.5 * ((double)(n - lay.n_old) * log2pi + 2.0 * l0 + q0);
                lay.dlogml[b] = dl;
                if (lay.logml_acc) lay.logml_acc[b] += dl;
                a.info[b] = 0;
            }
            continue;
        }
        if (tid == 0) {
            double r2 = 0, r3 = 0, l0 = 0, l1 = 0;
            for (int w = 0; w < kWarps; ++w) { r2 += s_red[0][w]; r3 += s_red[1][w]; l0 += s_red[2][w]; l1 += s_red[3][w]; }
            const double log2pi = 1.8378770664093454835606594728112;
            double lmn = -0.5 * ((double)n * log2pi + 2.0 * l0 + r2);
            double lmm = have_y2 ? -0.5 * ((double)m * log2pi + 2.0 * l1 + r3) : nan("");
            if (a.logml_n) a.logml_n[b] = lmn;
            if (a.logml_m) a.logml_m[b] = lmm;
            if (a.logw) a.logw[b] = (a.logw0 ? a.logw0[p] : 0.0) + (lmm - lmn);
            a.info[b] = 0;
        }

        // ---- predictive moments / fast-path tail blocks ------------------------------------------------
        const int kh = k + h;
        if (a.mu && have_y2) {
            for (int r = warp; r < h; r += kWarps) {
                double accv = 0.0;
                for (int cix = lane; cix < m; cix += 32) accv = fma(Lel(m + r, cix), zel(cix), accv);
                accv = warp_sum(accv);
                if (lane == 0) a.mu[b * h + r] = (accv - a.yb) / a.ya;
            }
        }
        if (a.L33) {
            for (int e = tid; e < h * h; e += kThreads) {
                int r = e / h, cix = e - r * h;
                a.L33[b * h * h + e] = cix <= r ? Lel(m + r, m + cix) / a.ya : 0.0;
            }
        }
        if (a.proj) {
            for (int r = warp; r < kh; r += kWarps) {
                double accv = 0.0;
                for (int cix = lane; cix < n; cix += 32) accv = fma(Lel(n + r, cix), zel(cix), accv);
                accv = warp_sum(accv);
                if (lane == 0) a.proj[b * kh + r] = accv;
            }
        }
        if (a.Ltail) {
            for (int e = tid; e < kh * kh; e += kThreads) {
                int r = e / kh, cix = e - r * kh;
                a.Ltail[b * kh * kh + e] = cix <= r ? Lel(n + r, n + cix) : 0.0;
            }
        }
    }
}

// ---- rank-append ---------------------------------------------------------------------------------------
// Extends each particle's stored factor from n_old to n_new = a.n points in place. Tile rows
// I0 = floor(n_old / 8) .. nt-1 are (re)computed up-looking in groups of <= 8 tile rows:
//   X_J = (A(I, J) - sum_{P<J} X_P L_JP^T) W_J^T   for J < I,   chol8 of the remainder for J == I,
// and the same recurrence gives the new entries of z = L^-1 y. The stored rows J of L are swept in blocks
// of 8: warp w streams row J0+w (contiguous in storage, 4 KB in flight per warp, read exactly once per
// group) and sums over the columns P < J0 on its own; the 8x8-tile triangle of the block is then finished
// row by row, one barrier per row.
struct AppendLayout {
    LargeLayout base;
    int n_old;
    int ring;            // 1: the launch has the streaming ring in shared memory (appends inside one tile row)
    double *logml;       // [P] in/out: running log marginal likelihood of the stored factor
    double *dlogml;      // [P] out
};

// RING: the single-tile-row sweep streams through the shared-memory ring (its own instantiation: with both stream loops in
// one kernel the block-append path spills twice as much at the 128-register cap and a 19-append schedule ran 10 % slower).
template <bool RING>
__global__ void __launch_bounds__(kThreads, 2) rank_append_kernel(const FusedArgs a, const AppendLayout al)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ TreeProgram tp;
    __shared__ int s_info;
    __shared__ double s_acc[2];   // sum log L_ii, sum z_i^2 over the new rows

    const LargeLayout &lay = al.base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, q = n;                 // append works on observed points only (k = h = 0)
    const int nt = lay.nt, Q = lay.ntp * 8, yrow = lay.yrow;
    const int n_old = al.n_old;
    const int I0 = n_old >> 3;

    double *s_x = smem;                                   // [8] X tiles of the current block (operand layout)
    double *s_ring = s_x + kWarps * 64;                   // RING: [kWarps][kApStages][8 tiles] streaming ring (R == 1)
    double *s_T = s_ring;                                 // !RING: [kBlk][kWarps] C tiles of a block, by new row (block appends)
    constexpr bool use_ring = RING && kApStages > 0;
    char *aux_s = reinterpret_cast<char *>(s_ring + (use_ring ? kWarps * kApStages * 8 * 64 : kBlk * kWarps * 64));
    const Setup su = aux_pointers(lay, aux_s);

    for (int i = tid; i < Q; i += kThreads) {
        su.tt[i] = i < q ? a.t[i] : 0.0;
        su.gg[i] = (a.g && i < q) ? a.g[i] : 0;
    }
    __syncthreads();

    for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x) {
        __syncthreads();
        const int p = (int)b;
        double *Lb = lay.L + (size_t)b * lay.L_stride;
        double *Wb = lay.W + (size_t)b * lay.W_stride;
        if (tid == 0) { s_info = 0; s_acc[0] = 0.0; s_acc[1] = 0.0; }
        instance_setup(tp, a, su, 0, p, Q, tid);
        if (tp.error) {
            if (tid == 0) { a.info[b] = tp.error; if (al.dlogml) al.dlogml[b] = nan(""); }
            continue;
        }
        GramCtx gc;
        gc.cx.th = su.th; gc.cx.tt = su.tt; gc.cx.tab = su.tab; gc.cx.sig = su.sig;
        gc.cx.step = a.step; gc.cx.G = a.G; gc.cx.Q = Q; gc.cx.grid = a.g != nullptr;
        gc.gg = su.gg; gc.q = q; gc.m = q;
        gc.d_lo = gc.d_hi = a.noise[p] + a.jitter;
        gc.single_table = (tp.clen == 1 && tp.cop[0] == OP_TABLE);

        for (int g0 = I0; g0 < nt && !s_info; g0 += kBlk) {
            const int R = min(kBlk, nt - g0);           // new tile rows g0 .. g0+R-1 (+ the y row)
            const int Jend = g0 + R;                    // rows J = 0 .. Jend-1 of L are swept
            if (R == 1) {
                // ---- one new tile row (appends of <= 8 points): the HBM-bound case ------------------------
                // Row J of the stored factor contributes X_J = (A(g0,J) - sum_{P<J} X_P L_JP^T) W_J^T. Warp w
                // streams row J0+w (8 tiles = 4 KB in flight per warp, next to the matching tiles of the new
                // row from L1/L2) for the columns P < J0; the last <= 7 terms use the X tiles of this block,
                // handed from warp to warp through shared memory.
                const double *xrow = Lb + (size_t)tri(g0) * 64 + lane * 2;      // the new row (A operand)
                const double *zrow = Lb + (size_t)tri(yrow) * 64 + lane * 2;    // z = L^-1 y
                const uint32_t ring_a = smem_addr(s_ring) + (uint32_t)(warp * (kApStages * 4096) + lane * 16);
#pragma unroll
                for (int st = 0; st < kApStages - 1; ++st) cp_async_commit();   // block J0 = 0 streams nothing
                for (int J0 = 0; J0 <= g0; J0 += kWarps) {
                    const int J = J0 + warp;
                    const bool rowv = J <= g0;
                    const bool last = J == g0;                                   // the new row itself
                    const double *lrow = Lb + (size_t)tri(rowv ? J : 0) * 64 + lane * 2;
                    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;               // two chains of the row sum
                    double y0 = 0.0, y1 = 0.0;                                   // z_g0 sum (last row only)
                    if constexpr (!use_ring) {
                      if (rowv) {
                        int P = 0;
                        for (; P + 8 <= J0; P += 8) {      // 4 KB of the stored row in flight per warp
                            double2 bf[8], af[4];
#pragma unroll
                            for (int i = 0; i < 8; ++i) bf[i] = ldg128_stream(lrow + (size_t)(P + i) * 64);
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) af[i] = ldg128(xrow + (size_t)(P + hh * 4 + i) * 64);
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    dmma(a0, a1, af[i].x, bf[hh * 4 + i].x);
                                    dmma(b0, b1, af[i].y, bf[hh * 4 + i].y);
                                }
                            }
                            if (last) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const double2 zf = ldg128(zrow + (size_t)(P + i) * 64);
                                    dmma(y0, y1, zf.x, bf[i].x);
                                    dmma(y0, y1, zf.y, bf[i].y);
                                }
                            }
                        }
                      }
                    } else {
                        // The first stages of this row were issued before the previous block's triangle (below): every
                        // lane copies exactly the 16 bytes per tile it consumes, so the ring needs no barrier at all.
                        const int nch = rowv ? (J0 >> 3) : 0;          // chunks of 8 tiles: P = 8 c
                        int slot = 0, islot = (kApStages - 1) % (kApStages > 0 ? kApStages : 1);
                        if (last) {
                            // the new row reads its own X tiles: they exist only now, after the previous block's triangle
#pragma unroll
                            for (int st = 0; st < kApStages - 1; ++st) {
                                if (st < nch) {
                                    const uint32_t dst = ring_a + (uint32_t)(st * 4096);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) cp_async16(dst + i * 512, lrow + (size_t)(st * 8 + i) * 64);
                                }
                                cp_async_commit();
                            }
                        }
                        for (int c = 0; c < nch; ++c) {
                            if (c + kApStages - 1 < nch) {
                                const double *src = lrow + (size_t)(c + kApStages - 1) * 8 * 64;
                                const uint32_t dst = ring_a + (uint32_t)(islot * 4096);
#pragma unroll
                                for (int i = 0; i < 8; ++i) cp_async16(dst + i * 512, src + (size_t)i * 64);
                            }
                            cp_async_commit();
                            cp_async_wait<(kApStages > 0 ? kApStages - 1 : 0)>();
                            const uint32_t sa = ring_a + (uint32_t)(slot * 4096);
                            const int P = c * 8;
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                double2 af[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) af[i] = ldg128(xrow + (size_t)(P + hh * 4 + i) * 64);
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const double2 bf = lds128(sa + (uint32_t)((hh * 4 + i) * 512));
                                    dmma(a0, a1, af[i].x, bf.x);
                                    dmma(b0, b1, af[i].y, bf.y);
                                }
                            }
                            if (last) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const double2 zf = ldg128(zrow + (size_t)(P + i) * 64);
                                    const double2 bf = lds128(sa + (uint32_t)(i * 512));
                                    dmma(y0, y1, zf.x, bf.x);
                                    dmma(y0, y1, zf.y, bf.y);
                                }
                            }
                            slot = slot + 1 == kApStages ? 0 : slot + 1;
                            islot = islot + 1 == kApStages ? 0 : islot + 1;
                        }
                        cp_async_wait<0>();
                        // first stages of the next block's row, in flight under this block's triangle
                        const int Jn = J0 + kWarps + warp;
                        if (J0 + kWarps <= g0) {
                            const int nchn = Jn < g0 ? ((J0 + kWarps) >> 3) : 0;      // not the new row: see above
                            const double *lrown = Lb + (size_t)tri(Jn < g0 ? Jn : 0) * 64 + lane * 2;
#pragma unroll
                            for (int st = 0; st < kApStages - 1; ++st) {
                                if (st < nchn) {
                                    const uint32_t dst = ring_a + (uint32_t)(st * 4096);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) cp_async16(dst + i * 512, lrown + (size_t)(st * 8 + i) * 64);
                                }
                                cp_async_commit();
                            }
                        }
                    }
                    // operands of the in-block terms that do not depend on this block: L_JP (old rows), W_J, A(g0,J)
                    double2 bfb[kWarps - 1];
                    double2 ib = make_double2(0.0, 0.0);
                    double gt0 = 0.0, gt1 = 0.0;
                    if (rowv) {
#pragma unroll
                        for (int i = 0; i < kWarps - 1; ++i)
                            bfb[i] = (!last && i < warp) ? ldg128_stream(lrow + (size_t)(J0 + i) * 64) : make_double2(0.0, 0.0);
                        if (!last) ib = ldg128(Wb + (size_t)J * 64 + lane * 2);
                        double out[4];
                        gram_pair(tp, gc, g0, J, J, lane, out);
                        gt0 = out[0]; gt1 = out[1];
                    }
                    for (int j = 0; j < kWarps; ++j) {
                        // every row still open takes the X tile of step j-1 as soon as it exists (same order of terms
                        // as a row-by-row sum, but a step's critical path is one term, not up to seven)
                        if (j > 0 && warp >= j && rowv) {
                            const double2 af = *reinterpret_cast<const double2 *>(s_x + (j - 1) * 64 + lane * 2);
                            double2 bf = af;
                            if (!last) {
#pragma unroll
                                for (int i = 0; i < kWarps - 1; ++i) bf = (i == j - 1) ? bfb[i] : bf;
                            }
                            dmma(a0, a1, af.x, bf.x);
                            dmma(b0, b1, af.y, bf.y);
                            if (last) {
                                const double2 zf = ldg128(zrow + (size_t)(J0 + j - 1) * 64);
                                dmma(y0, y1, zf.x, bf.x);
                                dmma(y0, y1, zf.y, bf.y);
                            }
                        }
                        if (warp == j && rowv) {
                            const double c0v = gt0 - (a0 + b0), c1v = gt1 - (a1 + b1);
                            if (!last) {
                                const double2 fr = acc_to_frag(c0v, c1v, lane);
                                double x0 = 0.0, x1 = 0.0;
                                dmma(x0, x1, fr.x, ib.x);
                                dmma(x0, x1, fr.y, ib.y);
                                store_op(s_x + j * 64, x0, x1, lane);
                                store_op(Lb + ((size_t)tri(g0) + J) * 64, x0, x1, lane);
                            } else {
                                double d0 = c0v, d1 = c1v, w0, w1, piv[8];
                                const int bad = chol8_inv(d0, d1, w0, w1, lane, q - J * 8, piv);
                                if (bad && lane == 0 && !s_info) s_info = J * 8 + bad;
                                double ld = 0.0;
#pragma unroll
                                for (int pp = 0; pp < 8; ++pp) {
                                    const int row = J * 8 + pp;
                                    if (row >= n_old && row < n) ld += 0.5 * log(piv[pp]);
                                }
                                store_op(Lb + ((size_t)tri(J) + J) * 64, d0, d1, lane);
                                store_op(Wb + (size_t)J * 64, w0, w1, lane);
                                store_op(s_x + j * 64, w0, w1, lane);      // W_g0 as a B operand for the z tile
                                __syncwarp();
                                const double2 iw = *reinterpret_cast<const double2 *>(s_x + j * 64 + lane * 2);
                                double yg0, yg1;
                                y_tile(a, 0, 0, J, n, lane, yg0, yg1);
                                const double2 fr = acc_to_frag(yg0 - y0, yg1 - y1, lane);
                                double x0 = 0.0, x1 = 0.0;
                                dmma(x0, x1, fr.x, iw.x);
                                dmma(x0, x1, fr.y, iw.y);
                                store_op(Lb + ((size_t)tri(yrow) + J) * 64, x0, x1, lane);
                                double zz = 0.0;
                                if ((lane >> 2) == 0) {
                                    const int c = J * 8 + (lane & 3) * 2;
                                    if (c >= n_old && c < n) zz = fma(x0, x0, zz);
                                    if (c + 1 >= n_old && c + 1 < n) zz = fma(x1, x1, zz);
                                }
                                zz = warp_sum(zz);
                                if (lane == 0) { s_acc[0] += ld; s_acc[1] += zz; }
                            }
                        }
                        if (J0 + j + 1 <= g0) __syncthreads();
                        else break;
                    }
                    __syncthreads();
                }
                continue;
            }
            for (int J0 = 0; J0 < Jend; J0 += kWarps) {
                // ---- streaming part: warp w owns row J = J0 + w of L and sums over the columns P < J0 ------
                const int J = J0 + warp;
                const bool rowv = J < Jend;
                const int r_lo = J > g0 ? J - g0 : 0;   // new rows r >= r_lo still need column J
                const bool y_on = J >= g0;              // z of this group's rows
                double acc[kBlk + 1][2];
#pragma unroll
                for (int r = 0; r <= kBlk; ++r) { acc[r][0] = 0.0; acc[r][1] = 0.0; }
                const double *lrow = Lb + (size_t)tri(rowv ? J : 0) * 64 + lane * 2;
                auto terms = [&](int P, const double2 bf) {
#pragma unroll
                    for (int r = 0; r < kBlk; ++r) {
                        if (r >= r_lo && r < R) {
                            const double2 af = ldg128(Lb + ((size_t)tri(g0 + r) + P) * 64 + lane * 2);
                            dmma(acc[r][0], acc[r][1], af.x, bf.x);
                            dmma(acc[r][0], acc[r][1], af.y, bf.y);
                        }
                    }
                    if (y_on) {
                        const double2 af = ldg128(Lb + ((size_t)tri(yrow) + P) * 64 + lane * 2);
                        dmma(acc[kBlk][0], acc[kBlk][1], af.x, bf.x);
                        dmma(acc[kBlk][0], acc[kBlk][1], af.y, bf.y);
                    }
                };
                if (rowv) {
                    int P = 0;
                    for (; P + 8 <= J0; P += 8) {      // 4 KB of the row in flight per warp
                        double2 bf[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) bf[i] = ldg128_stream(lrow + (size_t)(P + i) * 64);
#pragma unroll
                        for (int i = 0; i < 8; ++i) terms(P + i, bf[i]);
                    }
                    for (; P < J0; ++P) terms(P, ldg128_stream(lrow + (size_t)P * 64));
                }
                // ---- everything of the block that depends on nobody else, in parallel on the eight warps: the Gram tiles
                // (and y) of every open row against column J, turned into C = A - sum at once. (They used to be evaluated
                // inside the row's turn of the triangle below, i.e. by one warp at a time with seven waiting: up to eight
                // interpreter passes per turn.)
                const int r_d = J - g0;                 // J is itself a new row: its diagonal tile is slot r_d
                if (rowv) {
#pragma unroll
                    for (int r = 0; r <= kBlk; ++r) {
                        const bool isy = (r == kBlk);
                        const bool open = isy ? y_on : (r < R && g0 + r >= J);      // rows below J, and J's own diagonal tile
                        if (!open) continue;
                        double g0v, g1v;
                        if (isy) {
                            y_tile(a, 0, 0, J, n, lane, g0v, g1v);
                        } else {
                            double out[4];
                            gram_pair(tp, gc, g0 + r, J, J, lane, out);
                            g0v = out[0]; g1v = out[1];
                        }
                        acc[r][0] = g0v - acc[r][0];
                        acc[r][1] = g1v - acc[r][1];
                    }
                }
                if (!RING && J0 + kWarps <= g0) {
                    // ---- block of stored rows only: the new rows do not depend on each other here, so the block's
                    // triangle is solved per NEW row — one exchange through shared memory, then every warp runs an
                    // in-register right-looking solve of its new row against the (stored) 8 x 8-tile diagonal block,
                    // with no turn-taking at all
                    // (z = L^-1 y does not change at stored columns: no y row here)
#pragma unroll
                    for (int r = 0; r < kBlk; ++r)
                        if (r < R)
                            *reinterpret_cast<double2 *>(s_T + (r * kWarps + warp) * 64 + lane * 2) = make_double2(acc[r][0], acc[r][1]);
                    __syncthreads();
                    for (int u = warp; u < R; u += kWarps) {
                        const int ru = u;
                        const int I = g0 + u;
                        double c[kWarps][2];
#pragma unroll
                        for (int cc = 0; cc < kWarps; ++cc) {
                            const double2 v = *reinterpret_cast<const double2 *>(s_T + (ru * kWarps + cc) * 64 + lane * 2);
                            c[cc][0] = v.x; c[cc][1] = v.y;
                        }
#pragma unroll
                        for (int cc = 0; cc < kWarps; ++cc) {
                            const int Jc = J0 + cc;
                            const double2 ib = ldg128(Wb + (size_t)Jc * 64 + lane * 2);
                            const double2 fr = acc_to_frag(c[cc][0], c[cc][1], lane);
                            double x0 = 0.0, x1 = 0.0;
                            dmma(x0, x1, fr.x, ib.x);
                            dmma(x0, x1, fr.y, ib.y);
                            store_op(Lb + ((size_t)tri(I) + Jc) * 64, x0, x1, lane);
                            if (cc + 1 < kWarps) {
                                const double2 xf = acc_to_frag(x0, x1, lane);
#pragma unroll
                                for (int c2 = cc + 1; c2 < kWarps; ++c2) {
                                    const double2 bf = ldg128(Lb + ((size_t)tri(J0 + c2) + Jc) * 64 + lane * 2);
                                    dmma(c[c2][0], c[c2][1], -xf.x, bf.x);
                                    dmma(c[c2][0], c[c2][1], -xf.y, bf.y);
                                }
                            }
                        }
                    }
                    __threadfence_block();
                    __syncthreads();
                    continue;
                }
                // ---- the 8 x 8-tile triangle of this block: row J0+j needs the tiles finished by rows < J0+j; every row
                // still open takes the tiles of step j-1 as soon as they exist, so a turn is one term and the solves
                for (int j = 0; j < kWarps; ++j) {
                    if (j > 0 && warp >= j && rowv) {
                        const double2 bf = ldg128(lrow + (size_t)(J0 + j - 1) * 64);
                        terms(J0 + j - 1, make_double2(-bf.x, -bf.y));
                    }
                    if (warp == j && rowv) {
                        double2 ib;
                        if (J >= g0) {
                            // J is a new row: its diagonal tile (new row r = J - g0)
                            double d0 = 0.0, d1 = 0.0, w0, w1, piv[8];
#pragma unroll
                            for (int r = 0; r < kBlk; ++r) if (r == r_d) { d0 = acc[r][0]; d1 = acc[r][1]; }
                            const int bad = chol8_inv(d0, d1, w0, w1, lane, q - J * 8, piv);
                            if (bad && lane == 0 && !s_info) s_info = J * 8 + bad;
                            double ld = 0.0;
#pragma unroll
                            for (int pp = 0; pp < 8; ++pp) {
                                const int row = J * 8 + pp;
                                if (row >= n_old && row < n) ld += 0.5 * log(piv[pp]);
                            }
                            if (lane == 0) s_acc[0] += ld;
                            store_op(Lb + ((size_t)tri(J) + J) * 64, d0, d1, lane);
                            store_op(Wb + (size_t)J * 64, w0, w1, lane);
                            __syncwarp();
                        }
                        ib = ldg128(Wb + (size_t)J * 64 + lane * 2);
#pragma unroll
                        for (int r = 0; r <= kBlk; ++r) {
                            const bool isy = (r == kBlk);
                            if (isy ? !y_on : !(r < R && g0 + r > J)) continue;
                            const double2 fr = acc_to_frag(acc[r][0], acc[r][1], lane);
                            double x0 = 0.0, x1 = 0.0;
                            dmma(x0, x1, fr.x, ib.x);
                            dmma(x0, x1, fr.y, ib.y);
                            const int I = isy ? yrow : g0 + r;
                            store_op(Lb + ((size_t)tri(I) + J) * 64, x0, x1, lane);
                            if (isy && (lane >> 2) == 0) {
                                double zz = 0.0;
                                const int c = J * 8 + (lane & 3) * 2;
                                if (c >= n_old && c < n) zz = fma(x0, x0, zz);
                                if (c + 1 >= n_old && c + 1 < n) zz = fma(x1, x1, zz);
                                zz += __shfl_xor_sync(0x0000000fu, zz, 1);
                                zz += __shfl_xor_sync(0x0000000fu, zz, 2);
                                if (lane == 0) s_acc[1] += zz;
                            }
                        }
                    }
                    if (J0 + j + 1 < Jend) { __threadfence_block(); __syncthreads(); }
                    else break;
                }
                __syncthreads();
            }
        }
        if (tid == 0) {
            const double log2pi = 1.8378770664093454835606594728112;
            if (s_info) {
                a.info[b] = s_info;
                if (al.dlogml) al.dlogml[b] = nan("");
                if (al.logml) al.logml[b] = nan("");
            } else {
                const double d = -0.5 * ((double)(n - n_old) * log2pi + 2.0 * s_acc[0] + s_acc[1]);
                a.info[b] = 0;
                if (al.dlogml) al.dlogml[b] = d;
                if (al.logml) al.logml[b] += d;
            }
        }
    }
}

void large_aux_sizes(int Q, int G, int ntheta_cap, int ntab_cap, int ncp_cap, size_t (&sz)[5])
{
    auto up = [](size_t x) { return (x + 15) & ~size_t(15); };
    sz[0] = up((size_t)std::max(ntheta_cap, 1) * sizeof(double));
    sz[1] = up((size_t)Q * sizeof(int));
    sz[2] = up((size_t)Q * sizeof(double));
    sz[3] = up((size_t)ncp_cap * Q * sizeof(double));
    sz[4] = up((size_t)ntab_cap * (G > 0 ? G : 0) * sizeof(double));
}

}  // namespace

int large_max_q() { return 4096; }

LargePlan plan_large(int q, int q_cap, int G, int ntheta_cap, int ntab_cap, int ncp_cap, int smem_per_sm, bool append,
                     bool ring)
{
    LargePlan pl{};
    pl.nt = (q + 7) / 8;
    pl.ntp = (pl.nt + kBlk - 1) / kBlk * kBlk;
    const int nt_cap = (std::max(q, q_cap) + 7) / 8;
    pl.ntp_cap = (nt_cap + kBlk - 1) / kBlk * kBlk;
    pl.yrow = pl.ntp_cap;
    pl.L_stride = ((size_t)pl.ntp_cap * (pl.ntp_cap + 1) / 2 + pl.ntp_cap) * 64;
    const int Q = pl.ntp * 8;
    pl.ring = append && ring && kApStages > 0;
    size_t base = append ? (size_t)(kWarps * 64 + (pl.ring ? kWarps * kApStages * 8 * 64 : kBlk * kWarps * 64)) * sizeof(double)
                         : (size_t)(3 * kCBT + kCB) * 64 * sizeof(double);
    size_t sz[5];
    large_aux_sizes(Q, G, ntheta_cap, ntab_cap, ncp_cap, sz);
    const size_t static_smem = 2048 + 1024;
    const size_t budget = (size_t)smem_per_sm / 2 - static_smem;   // keep two CTAs per SM
    size_t s_off = 0, g_off = 0;
    const int order[5] = {0, 1, 4, 3, 2};
    for (int oi = 0; oi < 5; ++oi) {
        int i = order[oi];
        if (base + s_off + sz[i] <= budget) { pl.aux_smem[i] = 1; pl.aux_off[i] = (int)s_off; s_off += sz[i]; }
        else { pl.aux_smem[i] = 0; pl.aux_off[i] = (int)g_off; g_off += sz[i]; }
    }
    pl.smem_bytes = base + s_off;
    pl.scratch_stride = (int)((g_off + 255) & ~size_t(255));
    pl.ok = q <= large_max_q();
    return pl;
}

static LargeLayout make_layout(const LargePlan &pl, char *scratch, double *L, int keep, double *W,
                               unsigned long long *work_counter)
{
    LargeLayout lay{};
    lay.nt = pl.nt; lay.ntp = pl.ntp; lay.yrow = pl.yrow;
    for (int i = 0; i < 5; ++i) { lay.aux_off[i] = pl.aux_off[i]; lay.aux_smem[i] = pl.aux_smem[i]; }
    lay.scratch_stride = pl.scratch_stride;
    lay.scratch = scratch;
    lay.L = L; lay.L_stride = pl.L_stride; lay.keep = keep;
    lay.W = W; lay.W_stride = (size_t)pl.ntp_cap * 64;
    lay.work_counter = work_counter;
    return lay;
}

int large_grid(const LargePlan &pl, int64_t B, int num_sms, bool append)
{
    int per_sm = 0;
    const void *fn = append ? (pl.ring ? (const void *)rank_append_kernel<true> : (const void *)rank_append_kernel<false>)
                            : (pl.nt >= kRowsLargeFrom ? (const void *)chol_large_kernel<2> : (const void *)chol_large_kernel<1>);
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, pl.smem_bytes) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    return (int)std::min<int64_t>((int64_t)per_sm * num_sms, B);
}

cudaError_t launch_chol_large(const FusedArgs &a, const LargePlan &pl, char *scratch, double *L, int keep, double *W,
                              unsigned long long *work_counter, int grid, cudaStream_t stream, int n_old, double *dlogml,
                              double *logml_acc)
{
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    LargeLayout lay = make_layout(pl, scratch, L, keep, W, work_counter);
    lay.i0 = dlogml ? n_old >> 3 : 0; lay.n_old = n_old; lay.dlogml = dlogml; lay.logml_acc = logml_acc;
    auto kern = pl.nt >= kRowsLargeFrom ? chol_large_kernel<2> : chol_large_kernel<1>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<grid, kThreads, pl.smem_bytes, stream>>>(a, lay);
    return cudaGetLastError();
}

cudaError_t launch_rank_append(const FusedArgs &a, const LargePlan &pl, char *scratch, double *L, double *W,
                               int n_old, double *logml, double *dlogml, int grid, cudaStream_t stream)
{
    AppendLayout al{};
    al.base = make_layout(pl, scratch, L, 1, W, nullptr);
    al.n_old = n_old; al.ring = pl.ring; al.logml = logml; al.dlogml = dlogml;
    auto kern = pl.ring ? rank_append_kernel<true> : rank_append_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<grid, kThreads, pl.smem_bytes, stream>>>(a, al);
    return cudaGetLastError();
}

}  // namespace nagp

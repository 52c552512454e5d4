// nagp_tile.cuh — device helpers shared by the tile kernels (nagp_fused_v2.cu, nagp_large.cu):
// DMMA m8n8k4 wrapper, explicit shared-memory accesses, the 8x8 operand-layout index, the 4-wide
// kernel-tree interpreter and the in-register 8x8 Cholesky + inverse. Arithmetic contract:
// docs/KERNEL_SPEC.md §3-§4.
#pragma once
#include "nagp_tree.cuh"

namespace nagp {
namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(kFull, v, src); }

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// explicit 32-bit shared-memory accesses: keeps the hot loops free of generic->shared address math
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 lds128(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, double v0, double v1)
{
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(v0), "d"(v1) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// element (r, c) of an operand-layout tile: both k-chunks of a fragment lane are adjacent
__device__ __forceinline__ int op_idx(int r, int c) { return ((r * 4 + (c & 3)) << 1) + (c >> 2); }

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// W-wide interpreter over arbitrary entries (ii[e], jj[e]): one lane evaluates its two accumulator-
// layout entries of W/2 tiles at once (op decoding is paid once per W entries; the W independent look-up and
// arithmetic chains per op hide the shared-memory latency). The two top stack levels live in registers; deeper levels
// (expression depth >= 3) go to local memory. Same formulas and evaluation order as tree_eval
// (docs/KERNEL_SPEC.md §3).
struct EvalCtx {
    const double *th, *tt, *tab, *sig;
    double step;
    int G, Q;
    bool grid;
};

template <int W>
__device__ __forceinline__ void tree_evalw(const TreeProgram &tp, const EvalCtx &cx, const int (&ii)[W],
                                           const int (&jj)[W], const int (&lag)[W], double (&top)[W])
{
    double st[MAX_STACK][W];
    double sec[W];
    int sp = 0;
    const int len = tp.clen;
#pragma unroll
    for (int e = 0; e < W; ++e) { top[e] = 0.0; sec[e] = 0.0; }
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
            } else if (op == OP_CONSTANT) {
#pragma unroll
                for (int e = 0; e < W; ++e) top[e] = p[0];
            } else {
                // stationary leaf evaluated directly (pairwise times, or no table slot left)
#pragma unroll
                for (int e = 0; e < W; ++e) {
                    const double delta = cx.grid ? (double)lag[e] * cx.step : fabs(cx.tt[ii[e]] - cx.tt[jj[e]]);
                    double v;
                    if (op == OP_SQEXP) { double r = delta / p[0]; v = p[1] * exp(-0.5 * (r * r)); }
                    else if (op == OP_GAMMAEXP) { double r = delta / p[0]; v = p[2] * exp(-pow(r, p[1])); }
                    else {
                        double sn = sin(3.14159265358979323846 * (delta / p[1]));
                        v = p[2] * exp(-2.0 * (sn * sn) / (p[0] * p[0]));
                    }
                    top[e] = v;
                }
            }
        } else {
            --sp;   // left operand is sec, right operand is top
            if (op == OP_PLUS) {
#pragma unroll
                for (int e = 0; e < W; ++e) top[e] = sec[e] + top[e];
            } else if (op == OP_TIMES) {
#pragma unroll
                for (int e = 0; e < W; ++e) top[e] = sec[e] * top[e];
            } else if (op == OP_CHANGEPOINT_TAB) {
                const double *sg = cx.sig + (wd >> 24) * cx.Q;
#pragma unroll
                for (int e = 0; e < W; ++e) {
                    const double si = sg[ii[e]], sj = sg[jj[e]];
                    top[e] = ((1.0 - si) * (1.0 - sj)) * sec[e] + (si * sj) * top[e];
                }
            } else {   // OP_CHANGEPOINT evaluated directly
#pragma unroll
                for (int e = 0; e < W; ++e) {
                    const double si = 0.5 * (1.0 + tanh((cx.tt[ii[e]] - p[0]) / p[1]));
                    const double sj = 0.5 * (1.0 + tanh((cx.tt[jj[e]] - p[0]) / p[1]));
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


__device__ __forceinline__ void tree_eval4(const TreeProgram &tp, const EvalCtx &cx, const int (&ii)[4],
                                           const int (&jj)[4], const int (&lag)[4], double (&top)[4])
{
    tree_evalw<4>(tp, cx, ii, jj, lag, top);
}

#ifndef NAGP_CHOL8_OLD
#define NAGP_CHOL8_OLD 0   // 1: the first (rsqrt-on-the-chain) version, kept for A/B timing builds
#endif

// MUFU seeds (about 20 good bits: the instruction reads only the upper word of its operand) refined by one
// third-order step each: enough for a result within an ulp or two, and the shortest dependent chain.
__device__ __forceinline__ double rcp_seeded(double d)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    const double e = fma(-d, x, 1.0);
    const double q = fma(e, e, e);
    return fma(x, q, x);
}
__device__ __forceinline__ double rsqrt_seeded(double d)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-(y * y), d, 1.0);
    const double t = fma(e, 0.375, 0.5);
    return fma(t, y * e, y);
}

// In-register Cholesky of an 8x8 tile held in DMMA accumulator layout (lane (r = l>>2, j = l&3)
// holds columns 2j, 2j+1 of row r), together with its inverse built by the same row operations
// (L^-1 A = L^T, so the operations that reduce A to L^T turn I into L^-1).
// Returns 0 or 1 + index of the first non-positive pivot among the first `nreal` rows.
//
// This is the serial chain of the tile-column factorisation (one warp, everybody else waits for it), so it is
// written for dependent depth: per pivot the chain is shuffle(d) -> MUFU -> 3 FMA (1/d) -> 1 FMA (Schur
// update). The products that do not need 1/d are formed while it is being computed, rows and columns that
// must not change are masked through zero factors instead of selects on the result, the scaling by
// 1/sqrt(d) (its own MUFU chain) happens off the chain, and the inverse is carried unscaled (row p of
// L^-1 is scaled once at the end).
#if !NAGP_CHOL8_OLD
__device__ __forceinline__ int chol8_inv(double &c0, double &c1, double &w0, double &w1, int lane,
                                         int nreal, double (&piv)[8])
{
    const int r = lane >> 2, j = lane & 3;
    w0 = (r == 2 * j) ? 1.0 : 0.0;
    w1 = (r == 2 * j + 1) ? 1.0 : 0.0;
    double my_rinv = 0.0;
    int bad = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const int pl = p >> 1;
        const double colv = (p & 1) ? c1 : c0;            // column p lives in the lanes j == pl
        const double d = shfl(colv, p * 4 + pl);
        piv[p] = d;
        if (!(d > 0.0) && p < nreal && bad == 0) bad = p + 1;
        const double rinv = rsqrt_seeded(d);
        if (p < 7) {
            const double arp = shfl(colv, r * 4 + pl);        // (r, p)
            const double ac0 = shfl(colv, (2 * j) * 4 + pl);  // (2j, p)
            const double ac1 = shfl(colv, (2 * j + 1) * 4 + pl);
            const double wp0 = shfl(w0, p * 4 + j);           // unscaled row p of the inverse
            const double wp1 = shfl(w1, p * 4 + j);
            const double x = rcp_seeded(d);
            const double am = r > p ? arp : 0.0;              // rows <= p do not change
            const double u0 = am * (2 * j > p ? ac0 : 0.0);   // columns <= p do not change
            const double u1 = am * (2 * j + 1 > p ? ac1 : 0.0);
            const double v0 = am * wp0, v1 = am * wp1;
            const double fin = r >= p ? colv * rinv : 0.0;
            c0 = fma(-u0, x, c0);
            c1 = fma(-u1, x, c1);
            w0 = fma(-v0, x, w0);
            w1 = fma(-v1, x, w1);
            if (j == pl) { if (p & 1) c1 = fin; else c0 = fin; }
        } else {
            if (j == pl) c1 = r >= p ? colv * rinv : 0.0;
        }
        if (r == p) my_rinv = rinv;
    }
    w0 *= my_rinv;
    w1 *= my_rinv;
    return bad;
}
#else
__device__ __forceinline__ int chol8_inv(double &c0, double &c1, double &w0, double &w1, int lane,
                                         int nreal, double (&piv)[8])
{
    const int r = lane >> 2, j = lane & 3;
    w0 = (r == 2 * j) ? 1.0 : 0.0;
    w1 = (r == 2 * j + 1) ? 1.0 : 0.0;
    int bad = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const int pl = p >> 1;
        const double colv = (p & 1) ? c1 : c0;
        const double d = shfl(colv, p * 4 + pl);
        piv[p] = d;
        if (!(d > 0.0) && p < nreal && bad == 0) bad = p + 1;
        const double rinv = rsqrt(d);
        const double lrp = shfl(colv, r * 4 + pl) * rinv;
        const double lc0 = shfl(colv, (2 * j) * 4 + pl) * rinv;
        const double lc1 = shfl(colv, (2 * j + 1) * 4 + pl) * rinv;
        const double wp0 = shfl(w0, p * 4 + j) * rinv;
        const double wp1 = shfl(w1, p * 4 + j) * rinv;
        if (r == p) { w0 = wp0; w1 = wp1; }
        else if (r > p) { w0 = fma(-lrp, wp0, w0); w1 = fma(-lrp, wp1, w1); }
        if (r > p) {
            if (2 * j > p) c0 = fma(-lrp, lc0, c0);
            if (2 * j + 1 > p) c1 = fma(-lrp, lc1, c1);
        }
        if (j == pl) {
            const double fin = r >= p ? lrp : 0.0;
            if (p & 1) c1 = fin; else c0 = fin;
        }
    }
    return bad;
}
#endif

// 8x8 Cholesky + inverse of a tile in accumulator layout, same operations in the same order as chol8_inv
// (nagp_tile.cuh) with the pivot loop rolled.
__device__ __forceinline__ int chol8_inv_rolled(double &c0, double &c1, double &w0, double &w1, int lane, int nreal)
{
    const int r = lane >> 2, j = lane & 3;
    w0 = (r == 2 * j) ? 1.0 : 0.0;
    w1 = (r == 2 * j + 1) ? 1.0 : 0.0;
    double my_rinv = 0.0;
    int bad = 0;
#pragma unroll 1
    for (int p = 0; p < 8; ++p) {
        const int pl = p >> 1;
        const bool odd = p & 1;
        const double colv = odd ? c1 : c0;                // column p lives in the lanes j == pl
        const double d = shfl(colv, p * 4 + pl);
        if (!(d > 0.0) && p < nreal && bad == 0) bad = p + 1;
        const double rinv = rsqrt_seeded(d);
        const double fin = r >= p ? colv * rinv : 0.0;
        if (p < 7) {
            const double arp = shfl(colv, r * 4 + pl);        // (r, p)
            const double ac0 = shfl(colv, (2 * j) * 4 + pl);  // (2j, p)
            const double ac1 = shfl(colv, (2 * j + 1) * 4 + pl);
            const double wp0 = shfl(w0, p * 4 + j);           // unscaled row p of the inverse
            const double wp1 = shfl(w1, p * 4 + j);
            const double x = rcp_seeded(d);
            const double am = r > p ? arp : 0.0;              // rows <= p do not change
            const double u0 = am * (2 * j > p ? ac0 : 0.0);   // columns <= p do not change
            const double u1 = am * (2 * j + 1 > p ? ac1 : 0.0);
            const double v0 = am * wp0, v1 = am * wp1;
            c0 = fma(-u0, x, c0);
            c1 = fma(-u1, x, c1);
            w0 = fma(-v0, x, w0);
            w1 = fma(-v1, x, w1);
        }
        if (j == pl) { if (odd) c1 = fin; else c0 = fin; }
        if (r == p) my_rinv = rinv;
    }
    w0 *= my_rinv;
    w1 *= my_rinv;
    return bad;
}


}  // namespace
}  // namespace nagp

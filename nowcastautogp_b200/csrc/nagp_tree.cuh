// nagp_tree.cuh — device-side kernel-tree program: validation, stationary-subtree folding and the
// per-entry interpreter. Formulas and evaluation order: docs/KERNEL_SPEC.md §3 (AutoGP.GP node set,
// codes from /root/reference/docs/src/vignettes/setting-priors.md:229-236).
#pragma once
#include <stdint.h>

namespace nagp {

enum : int {
    OP_CONSTANT = 1, OP_LINEAR = 2, OP_SQEXP = 3, OP_GAMMAEXP = 4, OP_PERIODIC = 5,
    OP_PLUS = 6, OP_TIMES = 7, OP_CHANGEPOINT = 8,
    // compiled-only opcodes
    OP_TABLE = 9,       // stationary subtree folded into a lag table; arg = table id
    OP_CHANGEPOINT_TAB = 10  // ChangePoint whose sigma(t_i) is tabulated per point; arg = theta offset, aux = sigma table id
};
constexpr int MAX_PROG = 64;
constexpr int MAX_STACK = 16;
constexpr int MAX_THETA = 3 * MAX_PROG;
constexpr int MAX_TABLES = 8;   // lag tables per instance (maximal stationary subtrees)
constexpr int MAX_CPTAB = 8;    // tabulated ChangePoint nodes per instance

__host__ __device__ inline int op_nparam(int op)
{
    switch (op) {
    case OP_CONSTANT: return 1;
    case OP_LINEAR: return 3;
    case OP_SQEXP: return 2;
    case OP_GAMMAEXP: return 3;
    case OP_PERIODIC: return 3;
    case OP_CHANGEPOINT: return 2;
    default: return 0;
    }
}

// Per-instance program state kept in shared memory.
struct TreeProgram {
    // source program (as received) with per-op theta offsets
    uint8_t sop[MAX_PROG];
    int16_t sarg[MAX_PROG];
    int slen;
    // compiled program evaluated per matrix entry
    uint8_t cop[MAX_PROG];
    int16_t carg[MAX_PROG];
    int8_t caux[MAX_PROG];
    uint32_t cword[MAX_PROG];   // cop | carg << 8 | caux << 24: one load per op in the hot interpreter
    int clen;
    // lag tables: slice [tab_src0, tab_src1) of the source program
    int ntab;
    int16_t tab_src0[MAX_TABLES], tab_src1[MAX_TABLES];
    // tabulated changepoints: theta offset of (location, scale)
    int ncp;
    int16_t cp_theta[MAX_CPTAB];
    int error;  // 0 or NAGP_E_PROGRAM
};

// Thread-0 only. Validates `prog`, records theta offsets, and (while table slots remain) folds every maximal
// stationary subtree (no Linear / ChangePoint inside) into an OP_TABLE node. A subtree is a
// contiguous slice of a post-order program, so folding is a slice replacement.
__host__ __device__ inline void tree_compile(TreeProgram &tp, const uint8_t *prog, int len, int ntheta,
                                    int tab_cap, int cp_cap)
{
    const bool use_tables = tab_cap > 0, use_cptab = cp_cap > 0;
    if (tab_cap > MAX_TABLES) tab_cap = MAX_TABLES;
    if (cp_cap > MAX_CPTAB) cp_cap = MAX_CPTAB;
    tp.error = 0; tp.ntab = 0; tp.ncp = 0; tp.clen = 0; tp.slen = len;
    if (len <= 0 || len > MAX_PROG) { tp.error = -3; return; }
    struct Ent { int16_t out_pos, src0; bool stat; };
    Ent st[MAX_STACK];
    int sp = 0, th = 0, nout = 0;

    auto seal = [&](Ent &e, int src1) {
        // replace compiled ops [e.out_pos, end-of-e) by one OP_TABLE; caller fixes nout
        if (!use_tables || !e.stat || tp.ntab >= tab_cap) return false;
        int id = tp.ntab++;
        tp.tab_src0[id] = e.src0; tp.tab_src1[id] = (int16_t)src1;
        tp.cop[e.out_pos] = OP_TABLE; tp.carg[e.out_pos] = (int16_t)id; tp.caux[e.out_pos] = 0;
        return true;
    };

    for (int i = 0; i < len; ++i) {
        int op = prog[i];
        if (op < 1 || op > 8) { tp.error = -3; return; }
        tp.sop[i] = (uint8_t)op; tp.sarg[i] = (int16_t)th;
        if (op <= OP_PERIODIC) {
            if (sp >= MAX_STACK) { tp.error = -3; return; }
            st[sp].out_pos = (int16_t)nout; st[sp].src0 = (int16_t)i; st[sp].stat = (op != OP_LINEAR);
            ++sp;
            tp.cop[nout] = (uint8_t)op; tp.carg[nout] = (int16_t)th; tp.caux[nout] = 0; ++nout;
        } else {
            if (sp < 2) { tp.error = -3; return; }
            Ent &L = st[sp - 2], &R = st[sp - 1];
            bool both = L.stat && R.stat && op != OP_CHANGEPOINT;
            if (!both) {
                // seal the right child first (it is the tail of the compiled buffer) ...
                if (seal(R, i)) nout = R.out_pos + 1;
                // ... then the left child: its slice ends where the right child's source begins
                if (L.stat && use_tables && tp.ntab < tab_cap) {
                    int l_end = R.out_pos;  // compiled end of L
                    seal(L, R.src0);
                    int shift = l_end - (L.out_pos + 1);
                    if (shift > 0) {
                        for (int c = l_end; c < nout; ++c) {
                            tp.cop[c - shift] = tp.cop[c]; tp.carg[c - shift] = tp.carg[c];
                            tp.caux[c - shift] = tp.caux[c];
                        }
                        nout -= shift;
                    }
                }
            }
            int cop = op, aux = 0;
            if (op == OP_CHANGEPOINT && use_cptab && tp.ncp < cp_cap) {
                aux = tp.ncp; tp.cp_theta[tp.ncp++] = (int16_t)th; cop = OP_CHANGEPOINT_TAB;
            }
            tp.cop[nout] = (uint8_t)cop; tp.carg[nout] = (int16_t)th; tp.caux[nout] = (int8_t)aux; ++nout;
            L.stat = both;
            --sp;
        }
        th += op_nparam(op);
    }
    if (sp != 1 || th != ntheta || th > MAX_THETA) { tp.error = -3; return; }
    if (st[0].stat && seal(st[0], len)) nout = 1;
    tp.clen = nout;
    for (int c = 0; c < nout; ++c)
        tp.cword[c] = (uint32_t)tp.cop[c] | ((uint32_t)(uint16_t)tp.carg[c] << 8) | ((uint32_t)(uint8_t)tp.caux[c] << 24);
}

// Evaluate ops[0..len) for one pair. `lag` indexes the lag tables (ignored when there are none).
// tab: [ntab][G] lag tables; sig: [ncp][q] tabulated changepoint sigmas.
__device__ __forceinline__ double tree_eval(const uint8_t *ops, const int16_t *args, const int8_t *aux,
                                            int len, const double *theta, double ti, double tj,
                                            double delta, int lag, const double *tab, int G,
                                            const double *sig, int q, int pi, int pj)
{
    double st[MAX_STACK];
    int sp = 0;
    for (int i = 0; i < len; ++i) {
        const int op = ops[i];
        const double *th = theta + args[i];
        switch (op) {
        case OP_CONSTANT: st[sp++] = th[0]; break;
        case OP_LINEAR: {
            double u = ti - th[0], w = tj - th[0];
            st[sp++] = fma(th[2], u * w, th[1]);
            break;
        }
        case OP_SQEXP: {
            double r = delta / th[0];
            st[sp++] = th[1] * exp(-0.5 * (r * r));
            break;
        }
        case OP_GAMMAEXP: {
            double r = delta / th[0];
            st[sp++] = th[2] * exp(-pow(r, th[1]));
            break;
        }
        case OP_PERIODIC: {
            double s = sin(3.14159265358979323846 * (delta / th[1]));
            double l = th[0];
            st[sp++] = th[2] * exp(-2.0 * (s * s) / (l * l));
            break;
        }
        case OP_PLUS: st[sp - 2] = st[sp - 2] + st[sp - 1]; --sp; break;
        case OP_TIMES: st[sp - 2] = st[sp - 2] * st[sp - 1]; --sp; break;
        case OP_CHANGEPOINT: {
            double si = 0.5 * (1.0 + tanh((ti - th[0]) / th[1]));
            double sj = 0.5 * (1.0 + tanh((tj - th[0]) / th[1]));
            st[sp - 2] = ((1.0 - si) * (1.0 - sj)) * st[sp - 2] + (si * sj) * st[sp - 1];
            --sp;
            break;
        }
        case OP_TABLE: st[sp++] = tab[args[i] * G + lag]; break;
        case OP_CHANGEPOINT_TAB: {
            double si = sig[aux[i] * q + pi], sj = sig[aux[i] * q + pj];
            st[sp - 2] = ((1.0 - si) * (1.0 - sj)) * st[sp - 2] + (si * sj) * st[sp - 1];
            --sp;
            break;
        }
        default: break;
        }
    }
    return st[0];
}

}  // namespace nagp

// nagp_fused_sq.cu — tile kernel with recycled shared-memory tiles: three resident matrices per SM.
//
// Same algorithm and arithmetic as nagp_fused_v2.cu (left-looking tile-column Cholesky, DMMA m8n8k4
// accumulation with one column of lookahead, in-register 8x8 diagonal factor + inverse, column solve as a
// DMMA pair against the inverse, observation vector as a virtual tile row). What changes is where tiles
// live. In a left-looking factorisation tile (I, P) is needed only while P < J <= I, so at most
// J (nt - J) <= (nt/2)^2 tiles are alive at once — half the triangle — and the Gram tile (I, J) is needed
// only when column J is formed. Hence:
//   * the factor is stored in an H x H square of tiles, H = ceil(nt / 2), by a static map under which a
//     tile's slot is reused exactly when its predecessor dies:
//         (I, P), I < H           -> square (row P,     col I)        top-left triangle, transposed
//         (I, P), I >= H, P < H   -> square (row I - H, col P)        the rectangle below it
//         (I, P), P >= H          -> square (row P - H, col I - H)    bottom-right triangle, in the rows the
//                                                                     rectangle frees one by one
//     (tiles whose row AND column lie in the forecast tail rows >= Tt are kept apart: the outputs read them);
//     for a fixed row the address is affine in P inside each of the three regimes, so the DMMA loop keeps its
//     base + P * stride form and needs no slot table;
//   * the Gram is produced one tile column ahead into a buffer of nt tiles by the warps that own the rows,
//     in the shadow of the diagonal-tile factorisation — the separate Gram pass is gone;
//   * diagonal tiles are consumed in registers (log-determinant accumulated on the fly), only W = L_JJ^-1 is kept.
// C2 (q = 160): 64 KB instead of 107 KB per instance => 3 CTAs of 6 warps per SM instead of 2 CTAs of 8.
//
// Replaces the same reference call sites as nagp_fused_v2.cu (/root/reference/src/forecasting.jl:133,135,46;
// /root/reference/src/make_and_fit_model.jl:91). Arithmetic contract: docs/KERNEL_SPEC.md §3-§6.
#include <algorithm>

#include "nagp_kernels.cuh"
#include "nagp_tree.cuh"
#include "nagp_tile.cuh"

namespace nagp {

namespace {

constexpr int kWs = 6;                 // warps per CTA
constexpr int kTs = kWs * 32;
constexpr int kSlots = 5;              // tile rows per warp: ceil(nt / kWs), nt <= 29

struct SqLayout {
    int nt, H, Tt, ntail;              // tile rows, square edge, first tail tile row, tail tile rows
    int off_tail, off_gbuf, off_yv, off_invL, off_aux;   // byte offsets in dynamic shared memory
    int aux_off[5], aux_smem[5];
    int scratch_stride;
    char *scratch;
    unsigned long long *work_counter;
};

struct Seg { uint32_t base, stride; };   // lane address of tile (I, P) = base + P * stride (mod 2^32)

template <int NA>
__device__ __forceinline__ void kloop_sq(double (&acc)[kSlots][2][2], const Seg b, const Seg (&r)[kSlots], int P0, int P1)
{
#pragma unroll 2
    for (int P = P0; P < P1; ++P) {
        const double2 bf = lds128(b.base + (uint32_t)P * b.stride);
#pragma unroll
        for (int u = 0; u < NA; ++u) {
            const double2 af = lds128(r[u].base + (uint32_t)P * r[u].stride);
            dmma(acc[u][0][0], acc[u][0][1], af.x, bf.x);
            dmma(acc[u][1][0], acc[u][1][1], af.y, bf.y);
        }
    }
}

__global__ void __launch_bounds__(kTs, 3) fused_sq_kernel(const FusedArgs a, const SqLayout lay)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ TreeProgram tp;
    __shared__ int s_info;
    __shared__ long long s_next;
    __shared__ double s_red[4][kWs];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, k = a.k, h = a.h, m = n + k, q = m + h;
    const int nt = lay.nt, Q = nt * 8, H = lay.H, Tt = lay.Tt;
    const bool have_y2 = (a.y2 != nullptr) || k == 0;
    const int ny = have_y2 ? m : n;
    const int G = a.G;

    char *sm = reinterpret_cast<char *>(smem);
    double *yv = reinterpret_cast<double *>(sm + lay.off_yv);
    char *aux_s = sm + lay.off_aux;
    char *aux_g = lay.scratch + (size_t)blockIdx.x * lay.scratch_stride;
    auto aux = [&](int i) { return (lay.aux_smem[i] ? aux_s : aux_g) + lay.aux_off[i]; };
    double *th = reinterpret_cast<double *>(aux(0));
    int *gg = reinterpret_cast<int *>(aux(1));
    double *tt = reinterpret_cast<double *>(aux(2));
    double *sig = reinterpret_cast<double *>(aux(3));
    double *tab = reinterpret_cast<double *>(aux(4));
    const uint32_t sq_a = smem_addr(sm), tail_a = smem_addr(sm + lay.off_tail), gbuf_a = smem_addr(sm + lay.off_gbuf);
    const uint32_t yv_a = smem_addr(yv), invL_a = smem_addr(sm + lay.off_invL);
    const uint32_t HS = (uint32_t)H * 512u;
    const int bA = H < Tt ? H : Tt;                  // regime A: P < bA

    // lane address of tile (I, P) per regime (see the header)
    auto segA = [&](int I) {
        Seg s;
        if (I < H) { s.base = sq_a + (uint32_t)I * 512u + lane * 16; s.stride = HS; }
        else { s.base = sq_a + (uint32_t)(I - H) * HS + lane * 16; s.stride = 512u; }
        return s;
    };
    auto segB = [&](int I) {
        Seg s;
        s.base = sq_a + (uint32_t)((I - H) - H * H) * 512u + lane * 16;
        s.stride = HS;
        return s;
    };
    auto segT = [&](int I) {
        Seg s;
        s.base = tail_a + (uint32_t)(tri(I - Tt) - Tt) * 512u + lane * 16;
        s.stride = 512u;
        return s;
    };
    // byte address of tile (I, P), I >= P (without the lane offset)
    auto tile_addr = [&](int I, int P) {
        if (P >= Tt) return tail_a + (uint32_t)(tri(I - Tt) + (P - Tt)) * 512u;
        if (I < H) return sq_a + (uint32_t)(P * H + I) * 512u;
        if (P < H) return sq_a + (uint32_t)((I - H) * H + P) * 512u;
        return sq_a + (uint32_t)((P - H) * H + (I - H)) * 512u;
    };

    for (int i = tid; i < Q; i += kTs) {
        tt[i] = i < q ? a.t[i] : 0.0;
        gg[i] = (a.g && i < q) ? a.g[i] : 0;
    }

    for (;;) {
        __syncthreads();
        if (tid == 0) s_next = (long long)atomicAdd(lay.work_counter, 1ull);
        __syncthreads();
        const int64_t b = s_next;
        if (b >= a.B) break;
        const int64_t s = b / a.P;
        const int p = (int)(b % a.P);
        const int64_t po = a.prog_off[p], plen = a.prog_off[p + 1] - po;
        const int64_t to = a.theta_off[p], ntheta = a.theta_off[p + 1] - to;
        const double *theta_g = a.theta + s * a.theta_stride_k + to;

        if (tid == 0) {
            s_info = 0;
            if (ntheta > MAX_THETA) tp.error = -3;
            else tree_compile(tp, a.prog + po, (int)plen, (int)ntheta, G > 0 ? a.ntab_cap : 0, a.ncp_cap);
        }
        for (int i = tid; i < ntheta && i < MAX_THETA; i += kTs) th[i] = theta_g[i];
        __syncthreads();
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
        for (int e = tid; e < ntab * G; e += kTs) {
            int id = e / G, lg = e - id * G;
            int s0 = tp.tab_src0[id], s1 = tp.tab_src1[id];
            tab[e] = tree_eval(tp.sop + s0, tp.sarg + s0, nullptr, s1 - s0, th, 0.0, 0.0,
                               (double)lg * a.step, 0, nullptr, 0, nullptr, 0, 0, 0);
        }
        for (int e = tid; e < ncp * Q; e += kTs) {
            int id = e / Q, i = e - id * Q;
            const double *cp = th + tp.cp_theta[id];
            sig[e] = 0.5 * (1.0 + tanh((tt[i] - cp[0]) / cp[1]));
        }
        {
            const double *y1 = a.y1 + b * a.y1_stride;
            for (int jx = tid; jx < Q; jx += kTs) {
                double v = 0.0;
                if (jx < n) v = y1[jx];
                else if (jx < ny) v = a.y2 ? a.y2[s * k + (jx - n)] : y1[jx];
                yv[jx] = v;
            }
        }
        __syncthreads();

        const double nz = a.noise[s * a.noise_stride_k + p];
        const double d_lo = nz + a.jitter;
        const double d_hi = (a.noise_pred >= 0.0 ? a.noise_pred : nz) + a.jitter;
        const bool single_table = (tp.clen == 1 && tp.cop[0] == OP_TABLE);
        EvalCtx cx;
        cx.th = th; cx.tt = tt; cx.tab = tab; cx.sig = sig;
        cx.step = a.step; cx.G = G; cx.Q = Q; cx.grid = a.g != nullptr;

        // Gram tiles (Ia, J) and (Ib, J) into the column buffer (accumulator layout)
        auto gram2 = [&](int Ia, int Ib, int J) {
            const int gr = lane >> 2, gc = (lane & 3) * 2;
            int ii[4], jj[4], lag[4];
            bool real[4];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int I = h2 ? Ib : Ia;
                const int gi = gg[I * 8 + gr];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int x = h2 * 2 + e;
                    ii[x] = I * 8 + gr;
                    jj[x] = J * 8 + gc + e;
                    const int lg = gi - gg[jj[x]];
                    lag[x] = lg < 0 ? -lg : lg;
                    real[x] = ii[x] < q && jj[x] < q;
                }
            }
            double out[4];
            if (single_table) {
#pragma unroll
                for (int x = 0; x < 4; ++x) out[x] = tab[lag[x]];
            } else {
                tree_eval4(tp, cx, ii, jj, lag, out);
            }
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                if (!real[x]) out[x] = (ii[x] == jj[x]) ? 1.0 : 0.0;
                else if (ii[x] == jj[x]) out[x] += (ii[x] < m) ? d_lo : d_hi;
            }
            sts128(gbuf_a + (uint32_t)Ia * 512u + lane * 16, out[0], out[1]);
            if (Ib != Ia) sts128(gbuf_a + (uint32_t)Ib * 512u + lane * 16, out[2], out[3]);
        };
        // column Jn of the Gram for the tile rows I >= Jn owned by warp `w`
        auto gram_rows = [&](int w, int Jn) {
            int I = Jn + ((w - Jn) % kWs + kWs) % kWs;          // first row >= Jn with I == w (mod kWs)
            for (; I < nt; I += 2 * kWs) {
                const int Ib = I + kWs < nt ? I + kWs : I;
                gram2(I, Ib, Jn);
            }
        };
        gram_rows(warp, 0);
        __syncthreads();

        // ---- left-looking tile-column Cholesky with one column of lookahead ---------------------------
        const int lr = lane >> 2, lj = lane & 3;
        const int oi0 = op_idx(lr, 2 * lj), oi1 = op_idx(lr, 2 * lj + 1);
        const int cv0 = (lane & ~3) + (lj >> 1), cv1 = cv0 + 2;
        const bool odd = lane & 1;
        const int nreg = warp < nt ? (nt - 1 - warp) / kWs + 1 : 0;
        const int Ilast = warp + (nreg - 1) * kWs;
        const bool has_y = (warp == nt % kWs);
        double accn[kSlots][2][2];
        double yacc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int u = 0; u < kSlots; ++u) { accn[u][0][0] = accn[u][0][1] = accn[u][1][0] = accn[u][1][1] = 0.0; }
        int pre_done = 0;
        double ld_n = 0.0, ld_m = 0.0;         // log-determinant terms of the diagonal tiles this warp factored

        // sum_{P0 <= P < P1} L_IP L_{Jc,P}^T for the owned rows I >= Jc (and the observation row)
        auto accumulate = [&](int Jc, int P0, int P1) {
            if (P0 >= P1) return;
            const int NAc = Ilast >= Jc ? (Ilast - Jc) / kWs + 1 : 0;
#pragma unroll 1
            for (int reg = 0; reg < 3; ++reg) {
                const int lo = reg == 0 ? 0 : (reg == 1 ? H : Tt);
                const int hi = reg == 0 ? bA : (reg == 1 ? Tt : P1);
                const int p0 = P0 > lo ? P0 : lo, p1 = P1 < hi ? P1 : hi;
                if (p0 >= p1) continue;
                Seg r[kSlots];
#pragma unroll
                for (int u = 0; u < kSlots; ++u) {
                    const int I = Ilast - u * kWs;
                    r[u] = reg == 0 ? segA(I > 0 ? I : 0) : (reg == 1 ? segB(I) : segT(I));
                }
                const Seg bs = reg == 0 ? segA(Jc) : (reg == 1 ? segB(Jc) : segT(Jc));
                switch (NAc) {
                case 1: kloop_sq<1>(accn, bs, r, p0, p1); break;
                case 2: kloop_sq<2>(accn, bs, r, p0, p1); break;
                case 3: kloop_sq<3>(accn, bs, r, p0, p1); break;
                case 4: kloop_sq<4>(accn, bs, r, p0, p1); break;
                case 5: kloop_sq<5>(accn, bs, r, p0, p1); break;
                default: break;
                }
                if (has_y) {
                    const uint32_t yp = yv_a + lj * 8;
#pragma unroll 2
                    for (int P = p0; P < p1; ++P) {
                        const double2 bf = lds128(bs.base + (uint32_t)P * bs.stride);
                        const double a0 = lr == 0 ? lds64(yp + P * 64) : 0.0;
                        const double a1 = lr == 0 ? lds64(yp + P * 64 + 32) : 0.0;
                        dmma(yacc[0][0], yacc[0][1], a0, bf.x);
                        dmma(yacc[1][0], yacc[1][1], a1, bf.y);
                    }
                }
            }
        };

        for (int J = 0; J < nt; ++J) {
            const int ow = J % kWs;
            const bool owner = (warp == ow);
            const int NA = Ilast >= J ? (Ilast - J) / kWs + 1 : 0;   // active regular rows (I >= J)
            accumulate(J, pre_done, J);
            double c[kSlots][2], cy[2] = {0.0, 0.0}, d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int u = 0; u < kSlots; ++u) {
                c[u][0] = 0.0; c[u][1] = 0.0;
                if (u < NA) {
                    const double2 g2 = lds128(gbuf_a + (uint32_t)(Ilast - u * kWs) * 512u + lane * 16);
                    c[u][0] = g2.x - (accn[u][0][0] + accn[u][1][0]);
                    c[u][1] = g2.y - (accn[u][0][1] + accn[u][1][1]);
                    d0 = c[u][0]; d1 = c[u][1];   // ends up holding slot NA-1: the diagonal tile of its owner
                }
                accn[u][0][0] = accn[u][0][1] = accn[u][1][0] = accn[u][1][1] = 0.0;
            }
            if (has_y) {
                const double y0 = lr == 0 ? lds64(yv_a + (J * 8 + 2 * lj) * 8) : 0.0;
                const double y1v = lr == 0 ? lds64(yv_a + (J * 8 + 2 * lj + 1) * 8) : 0.0;
                cy[0] = y0 - (yacc[0][0] + yacc[1][0]);
                cy[1] = y1v - (yacc[0][1] + yacc[1][1]);
                yacc[0][0] = yacc[0][1] = yacc[1][0] = yacc[1][1] = 0.0;
            }
            if (owner) {
                // this warp's Gram rows of column J are consumed: the helper may overwrite them
                __syncwarp();
                asm volatile("bar.arrive 2, 64;" ::: "memory");
                double w0, w1, piv[8];
                const int bad = chol8_inv(d0, d1, w0, w1, lane, q - J * 8, piv);
#pragma unroll
                for (int pp = 0; pp < 8; ++pp) {
                    const int row = J * 8 + pp;
                    if (row < m) {
                        const double l = 0.5 * log(piv[pp]);
                        ld_m += l;
                        if (row < n) ld_n += l;
                    }
                }
                if (J >= Tt) {   // diagonal tiles of the tail rows are read by the outputs
                    const uint32_t dt = tile_addr(J, J);
                    sts64(dt + oi0 * 8, d0);
                    sts64(dt + oi1 * 8, d1);
                }
                sts64(invL_a + oi0 * 8, w0);
                sts64(invL_a + oi1 * 8, w1);
                if (bad && lane == 0) s_info = J * 8 + bad;
                __syncwarp();
                asm volatile("bar.arrive 1, %0;" ::"n"(kTs) : "memory");
                pre_done = 0;
            } else {
                // next Gram column for the own rows; the next owner also produces the rows of the busy owner
                if (J + 1 < nt) gram_rows(warp, J + 1);
                if (warp == (J + 1) % kWs) {
                    asm volatile("bar.sync 2, 64;" ::: "memory");
                    if (J + 1 < nt) gram_rows(ow, J + 1);
                }
                if ((warp & 3) == (ow & 3)) {
                    pre_done = 0;          // shares the owner's scheduler: leave the FP64 pipe to the diagonal tile
                } else {
                    if (J + 1 < nt) accumulate(J + 1, 0, J);
                    pre_done = J;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kTs) : "memory");
            }
            // column solve: X = C * W_J^T, stored in operand layout at the tile's slot
            if (!s_info) {
                const double2 ib = lds128(invL_a + lane * 16);
                const int nsolve = owner ? NA - 1 : NA;   // rows strictly below the diagonal
#pragma unroll
                for (int u = 0; u < kSlots; ++u) {
                    if (u < nsolve) {
                        const double v00 = shfl(c[u][0], cv0), v01 = shfl(c[u][1], cv0);
                        const double v10 = shfl(c[u][0], cv1), v11 = shfl(c[u][1], cv1);
                        double x0 = 0.0, x1 = 0.0;
                        dmma(x0, x1, odd ? v01 : v00, ib.x);
                        dmma(x0, x1, odd ? v11 : v10, ib.y);
                        const uint32_t dt = tile_addr(Ilast - u * kWs, J);
                        sts64(dt + oi0 * 8, x0);
                        sts64(dt + oi1 * 8, x1);
                    }
                }
                if (has_y) {
                    const double v00 = shfl(cy[0], cv0), v01 = shfl(cy[1], cv0);
                    const double v10 = shfl(cy[0], cv1), v11 = shfl(cy[1], cv1);
                    double x0 = 0.0, x1 = 0.0;
                    dmma(x0, x1, odd ? v01 : v00, ib.x);
                    dmma(x0, x1, odd ? v11 : v10, ib.y);
                    if (lr == 0) {
                        sts64(yv_a + (J * 8 + 2 * lj) * 8, x0);
                        sts64(yv_a + (J * 8 + 2 * lj + 1) * 8, x1);
                    }
                }
            }
            __syncthreads();
            if (s_info) break;
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

        // element (i, j), i >= j, of the factor: only rows >= 8 Tt are guaranteed to be intact here
        auto Lel = [&](int i, int j) {
            const double *t = reinterpret_cast<const double *>(sm + (tile_addr(i >> 3, j >> 3) - sq_a));
            return t[op_idx(i & 7, j & 7)];
        };

        // ---- logML(n), logML(m) ----------------------------------------------------------------------
        const double *z = yv;
        double qd_n = 0, qd_m = 0;
        for (int r = tid; r < ny; r += kTs) {
            const double zz = z[r] * z[r];
            qd_m += zz;
            if (r < n) qd_n += zz;
        }
        qd_n = warp_sum(qd_n); qd_m = warp_sum(qd_m);
        if (lane == 0) { s_red[0][warp] = ld_n; s_red[1][warp] = ld_m; s_red[2][warp] = qd_n; s_red[3][warp] = qd_m; }
        __syncthreads();
        if (tid == 0) {
            double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
            for (int w = 0; w < kWs; ++w) { r0 += s_red[0][w]; r1 += s_red[1][w]; r2 += s_red[2][w]; r3 += s_red[3][w]; }
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
            for (int r = warp; r < h; r += kWs) {
                double accv = 0.0;
                for (int cix = lane; cix < m; cix += 32) accv = fma(Lel(m + r, cix), z[cix], accv);
                accv = warp_sum(accv);
                if (lane == 0) a.mu[b * h + r] = (accv - a.yb) / a.ya;
            }
        }
        if (a.L33) {
            for (int e = tid; e < h * h; e += kTs) {
                int r = e / h, cix = e - r * h;
                a.L33[b * h * h + e] = cix <= r ? Lel(m + r, m + cix) / a.ya : 0.0;
            }
        }
        if (a.proj) {
            for (int r = warp; r < kh; r += kWs) {
                double accv = 0.0;
                for (int cix = lane; cix < n; cix += 32) accv = fma(Lel(n + r, cix), z[cix], accv);
                accv = warp_sum(accv);
                if (lane == 0) a.proj[b * kh + r] = accv;
            }
        }
        if (a.Ltail) {
            for (int e = tid; e < kh * kh; e += kTs) {
                int r = e / kh, cix = e - r * kh;
                a.Ltail[b * kh * kh + e] = cix <= r ? Lel(n + r, n + cix) : 0.0;
            }
        }
    }
}

}  // namespace

// Shared-memory plan of the recycled-tile kernel. ok = 0 unless three CTAs fit on an SM (otherwise the
// plain tile kernel with two CTAs of eight warps is the better configuration).
SqPlan plan_fused_sq(int q, int n, bool need_tail, int G, int ntheta_cap, int ntab_cap, int ncp_cap, int smem_per_sm)
{
    SqPlan pl{};
    const int nt = (q + 7) / 8, Q = nt * 8;
    pl.nt = nt;
    pl.H = (nt + 1) / 2;
    pl.Tt = need_tail ? std::min(n >> 3, nt) : nt;
    pl.ntail = nt - pl.Tt;
    if (nt > kSlots * kWs - 1 || (need_tail && pl.Tt < pl.H)) { pl.ok = 0; return pl; }   // tail rows must lie below the fold
    auto up = [](size_t x) { return (x + 15) & ~size_t(15); };
    size_t off = (size_t)pl.H * pl.H * 512;
    pl.off_tail = (int)off; off += (size_t)(pl.ntail * (pl.ntail + 1) / 2) * 512;
    pl.off_gbuf = (int)off; off += (size_t)nt * 512;
    pl.off_yv = (int)off; off += up((size_t)Q * sizeof(double));
    pl.off_invL = (int)off; off += 512;
    pl.off_aux = (int)off;
    size_t sz[5];
    sz[0] = up((size_t)std::max(ntheta_cap, 1) * sizeof(double));
    sz[1] = up((size_t)Q * sizeof(int));
    sz[2] = up((size_t)Q * sizeof(double));
    sz[3] = up((size_t)ncp_cap * Q * sizeof(double));
    sz[4] = up((size_t)ntab_cap * (G > 0 ? G : 0) * sizeof(double));
    const size_t static_smem = 1280 + 1024;
    const size_t budget = (size_t)smem_per_sm / 3 - static_smem;
    if (off > budget) { pl.ok = 0; return pl; }
    size_t s_off = 0, g_off = 0;
    const int order[5] = {0, 1, 4, 3, 2};   // th, gg, tab, sig, tt
    for (int oi = 0; oi < 5; ++oi) {
        int i = order[oi];
        if (off + s_off + sz[i] <= budget) { pl.aux_smem[i] = 1; pl.aux_off[i] = (int)s_off; s_off += sz[i]; }
        else { pl.aux_smem[i] = 0; pl.aux_off[i] = (int)g_off; g_off += sz[i]; }
    }
    pl.smem_bytes = off + s_off;
    pl.scratch_stride = (int)((g_off + 255) & ~size_t(255));
    pl.ok = 1;
    return pl;
}

int fused_sq_grid(const SqPlan &pl, int64_t B, int num_sms)
{
    int per_sm = 0;
    cudaFuncSetAttribute(fused_sq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_sq_kernel, kTs, pl.smem_bytes) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    return (int)std::min<int64_t>((int64_t)per_sm * num_sms, B);
}

cudaError_t launch_fused_sq(const FusedArgs &a, const SqPlan &pl, char *scratch, unsigned long long *work_counter,
                            int grid, cudaStream_t stream)
{
    SqLayout lay{};
    lay.nt = pl.nt; lay.H = pl.H; lay.Tt = pl.Tt; lay.ntail = pl.ntail;
    lay.off_tail = pl.off_tail; lay.off_gbuf = pl.off_gbuf; lay.off_yv = pl.off_yv; lay.off_invL = pl.off_invL;
    lay.off_aux = pl.off_aux;
    for (int i = 0; i < 5; ++i) { lay.aux_off[i] = pl.aux_off[i]; lay.aux_smem[i] = pl.aux_smem[i]; }
    lay.scratch_stride = pl.scratch_stride;
    lay.scratch = scratch;
    lay.work_counter = work_counter;
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fused_sq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) return e;
    fused_sq_kernel<<<grid, kTs, pl.smem_bytes, stream>>>(a, lay);
    return cudaGetLastError();
}

}  // namespace nagp

// chol8_bench2.cu — latency per call and accuracy of chol8_inv (csrc/nagp_tile.cuh) on one warp, alone on the
// GPU. Build twice: -DNAGP_CHOL8_OLD=1 (first version) and =0 (short-chain version). Not part of the product.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#ifndef NAGP_CHOL8_BLOCKED
#define NAGP_CHOL8_BLOCKED 0
#endif
#include "../nowcastautogp_b200/csrc/nagp_tile.cuh"
using namespace nagp;

// Experiment kept with its benchmark (round 2): 2x2 block pivots. Correct (|LL^T-A|/|A| 3.7e-16) but NOT faster alone on
// the GPU (1242 cycles against 1117 for chol8_inv): the rounds are halved, the instruction count is not (26 SHFL.32 and
// about 30 FP64 instructions per round), and a single warp issues them in order. Build with -DNAGP_CHOL8_BLOCKED=1.
namespace nagp { namespace {
// 8x8 Cholesky + inverse with 2x2 BLOCK pivots: three dependent elimination rounds instead of seven. Block b is
// rows/columns (2b, 2b+1) — exactly the column pair the lanes j == b hold. A round eliminates both columns at once
// through the inverse of the 2x2 pivot block P = [a bb; bb cc], written as adj(P) / det(P): the dependent chain of a
// round is shuffle (pivot block) -> det -> 1/det (MUFU + cubic step) -> one FMA per element, where the scalar version
// runs shuffle -> 1/d -> FMA twice. A warp issues in order, so nothing with a latency chain of its own may sit between
// the rounds: adj(P) u_r and its products are formed while 1/det is computed, and every square root is deferred to the
// end, where the four rsqrt chains a lane needs (its column block's and its row block's) run side by side.
// The eliminations leave A = Lb blockdiag(P_b) Lb^T with Lb block-unit-lower and turn I into Lb^-1; L = Lb
// blockdiag(chol P_b), so L^-1 = blockdiag((chol P_b)^-1) Lb^-1: one exchange between the two rows of a block at the
// end. Same pivots as the scalar factorisation (d_2b = a, d_2b+1 = det / a), so `bad` reports the same leading minor;
// results agree with chol8_inv to rounding (|L L^T - A| / |A| ~ 4e-16, tools/chol8_bench2.cu).
__device__ __forceinline__ int chol8_inv_blocked(double &c0, double &c1, double &w0, double &w1, int lane, int nreal)
{
    const int r = lane >> 2, j = lane & 3;
    w0 = (r == 2 * j) ? 1.0 : 0.0;
    w1 = (r == 2 * j + 1) ? 1.0 : 0.0;
    int bad = 0;
    double ca = 1.0, cb = 0.0, cd = 1.0;       // pivot block (a, bb, det) of this lane's COLUMN block j
    double ra_ = 1.0, rb_ = 0.0, rd_ = 1.0;    // ... and of its ROW block r >> 1
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double a = shfl(c0, (2 * b) * 4 + b);           // A[2b][2b]
        const double bb = shfl(c0, (2 * b + 1) * 4 + b);      // A[2b+1][2b]
        const double cc = shfl(c1, (2 * b + 1) * 4 + b);      // A[2b+1][2b+1]
        const double det = fma(a, cc, -(bb * bb));
        if (bad == 0) {
            if (!(a > 0.0) && 2 * b < nreal) bad = 2 * b + 1;
            else if (!(det > 0.0) && 2 * b + 1 < nreal) bad = 2 * b + 2;
        }
        if (j == b) { ca = a; cb = bb; cd = det; }
        if ((r >> 1) == b) { ra_ = a; rb_ = bb; rd_ = det; }
        if (b < 3) {
            const double ur0 = shfl(c0, r * 4 + b), ur1 = shfl(c1, r * 4 + b);                     // A[r][2b], A[r][2b+1]
            const double ua0 = shfl(c0, (2 * j) * 4 + b), ua1 = shfl(c1, (2 * j) * 4 + b);         // A[2j][2b], A[2j][2b+1]
            const double ub0 = shfl(c0, (2 * j + 1) * 4 + b), ub1 = shfl(c1, (2 * j + 1) * 4 + b); // A[2j+1][...]
            const double wa0 = shfl(w0, (2 * b) * 4 + j), wa1 = shfl(w1, (2 * b) * 4 + j);         // rows 2b, 2b+1 of the inverse
            const double wb0 = shfl(w0, (2 * b + 1) * 4 + j), wb1 = shfl(w1, (2 * b + 1) * 4 + j);
            const double x = rcp_seeded(det);
            const bool live = r > 2 * b + 1;                  // rows of the pivot block and above do not change
            const double m0 = live ? fma(ur0, cc, -(ur1 * bb)) : 0.0;     // adj(P) u_r
            const double m1 = live ? fma(ur1, a, -(ur0 * bb)) : 0.0;
            const double g0 = j > b ? fma(m0, ua0, m1 * ua1) : 0.0;       // columns of the pivot block and left of it do not change
            const double g1 = j > b ? fma(m0, ub0, m1 * ub1) : 0.0;
            const double h0 = fma(m0, wa0, m1 * wb0), h1 = fma(m0, wa1, m1 * wb1);
            c0 = fma(-g0, x, c0);
            c1 = fma(-g1, x, c1);
            w0 = fma(-h0, x, w0);
            w1 = fma(-h1, x, w1);
        }
    }
    // square roots, all at once: sqrt(det / a) = sqrt(det) / sqrt(a), so the four rsqrt chains are independent
    const double cra = rsqrt_seeded(ca), crd = rsqrt_seeded(cd), rra = rsqrt_seeded(ra_), rrd = rsqrt_seeded(rd_);
    {
        const double l21 = cb * cra, r1 = crd * (ca * cra);               // 1 / l22 = sqrt(a) / sqrt(det)
        const double f0 = r >= 2 * j ? c0 * cra : 0.0;                    // L[r][2j], L[r][2j+1]; zeros above the diagonal
        c1 = r >= 2 * j + 1 ? fma(-f0, l21, c1) * r1 : 0.0;
        c0 = f0;
    }
    const double i22 = rrd * (ra_ * rra), mm = -((rb_ * rra) * rra) * i22;   // (chol P)^-1 = [1/l11 0; -l21/(l11 l22) 1/l22]
    const double p0 = __shfl_xor_sync(kFull, w0, 4), p1 = __shfl_xor_sync(kFull, w1, 4);   // the other row of the block
    if (r & 1) { w0 = fma(mm, p0, i22 * w0); w1 = fma(mm, p1, i22 * w1); }
    else { w0 *= rra; w1 *= rra; }
    return bad;
}

} }

__global__ void bench(const double *A, double *L, double *W, long long *cyc, int reps)
{
    const int lane = threadIdx.x, r = lane >> 2, j = lane & 3;
    double a0 = A[r * 8 + 2 * j], a1 = A[r * 8 + 2 * j + 1];
    double c0 = a0, c1 = a1, w0 = 0, w1 = 0, piv[8];
    int bad = 0;
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
        c0 = a0 + c0 * 1e-300; c1 = a1 + c1 * 1e-300;   // serialise calls
#if NAGP_CHOL8_BLOCKED
        bad += chol8_inv_blocked(c0, c1, w0, w1, lane, 8); (void)piv;
#else
        bad += chol8_inv(c0, c1, w0, w1, lane, 8, piv);
#endif
    }
    long long t1 = clock64();
    L[r * 8 + 2 * j] = c0; L[r * 8 + 2 * j + 1] = c1;
    W[r * 8 + 2 * j] = w0; W[r * 8 + 2 * j + 1] = w1;
    if (lane == 0) { cyc[0] = (t1 - t0) / reps; cyc[1] = bad; }
}

// chol8_inv on warp 0 while the other warps of the CTA run a background load: mode 1 = DMMA only (independent
// accumulator chains), 2 = LDS.128 only, 3 = LDS.128 + DMMA as in the tile kernel's inner loop. `mask` selects
// which warps run the background (bit w); the rest exit at once.
__global__ void bench_bg(const double *A, long long *cyc, int reps, int mode, unsigned mask, double *sink)
{
    __shared__ double sm[16 * 64 * 4];
    __shared__ volatile int stop;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, r = lane >> 2, j = lane & 3;
    for (int i = threadIdx.x; i < 16 * 64 * 4; i += blockDim.x) sm[i] = 1e-3 * (i & 7);
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    if (warp == 0) {
        double a0 = A[r * 8 + 2 * j], a1 = A[r * 8 + 2 * j + 1];
        double c0 = a0, c1 = a1, w0 = 0, w1 = 0, piv[8];
        int bad = 0;
        for (int i = 0; i < 20; ++i) { c0 = a0 + c0 * 1e-300; c1 = a1 + c1 * 1e-300; bad += chol8_inv(c0, c1, w0, w1, lane, 8, piv); }
        long long t0 = clock64();
        for (int i = 0; i < reps; ++i) {
            c0 = a0 + c0 * 1e-300; c1 = a1 + c1 * 1e-300;
            bad += chol8_inv(c0, c1, w0, w1, lane, 8, piv);
        }
        long long t1 = clock64();
        __syncwarp();
        if (lane == 0) { cyc[0] = (t1 - t0) / reps; cyc[1] = bad; stop = 1; }
        sink[lane] = c0 + w0;
    } else if ((mask >> warp) & 1u) {
        double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        const uint32_t base = smem_addr(sm) + warp * 2048 + lane * 16;
        double2 bf = make_double2(1e-3, 2e-3), af = make_double2(3e-3, 1e-3);
        long long n = 0;
        while (!stop) {
#pragma unroll 2
            for (int P = 0; P < 8; ++P) {
                if (mode >= 2) { bf = lds128(base + (P & 3) * 512); af = lds128(base + ((P + 1) & 3) * 512); }
                if (mode == 1 || mode == 3) {
                    dmma(acc[0][0], acc[0][1], af.x, bf.x); dmma(acc[1][0], acc[1][1], af.y, bf.y);
                    dmma(acc[2][0], acc[2][1], af.y, bf.x); dmma(acc[3][0], acc[3][1], af.x, bf.y);
                } else { acc[0][0] += af.x + bf.y; }
            }
            ++n;
        }
        sink[32 + threadIdx.x] = acc[0][0] + acc[1][1] + acc[2][0] + acc[3][1] + (double)n;
    }
}

int main()
{
    double hA[64], hL[64], hW[64];
    double worst1 = 0, worst2 = 0; long long cycles = 0, bad = 0;
    double *A, *L, *W; long long *cyc, hc[2];
    cudaMalloc(&A, 512); cudaMalloc(&L, 512); cudaMalloc(&W, 512); cudaMalloc(&cyc, 16);
    for (int trial = 0; trial < 20; ++trial) {
        // SPD test matrices of growing condition number
        unsigned s = 12345u + trial;
        double B[8][8];
        for (int i = 0; i < 8; ++i) for (int k = 0; k < 8; ++k) { s = s * 1664525u + 1013904223u; B[i][k] = (double)(s >> 8) / 16777216.0 - 0.5; }
        const double ridge = pow(10.0, -0.5 * trial);
        for (int i = 0; i < 8; ++i) for (int k = 0; k < 8; ++k) {
            double v = 0; for (int m = 0; m < 8; ++m) v += B[i][m] * B[k][m];
            hA[i * 8 + k] = v + (i == k ? ridge : 0.0);
        }
        cudaMemcpy(A, hA, 512, cudaMemcpyHostToDevice);
        bench<<<1, 32>>>(A, L, W, cyc, 200);
        cudaMemcpy(hL, L, 512, cudaMemcpyDeviceToHost); cudaMemcpy(hW, W, 512, cudaMemcpyDeviceToHost);
        cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
        double e1 = 0, e2 = 0, nrm = 0;
        for (int i = 0; i < 64; ++i) nrm = fmax(nrm, fabs(hA[i]));
        for (int i = 0; i < 8; ++i) for (int k = 0; k < 8; ++k) {
            double sacc = 0, t = 0;
            for (int m = 0; m < 8; ++m) { sacc += hL[i * 8 + m] * hL[k * 8 + m]; t += hW[i * 8 + m] * hL[m * 8 + k]; }
            e1 = fmax(e1, fabs(sacc - hA[i * 8 + k]) / nrm); e2 = fmax(e2, fabs(t - (i == k)));
            if (k > i && (hL[i * 8 + k] != 0.0 || hW[i * 8 + k] != 0.0)) e1 = 1.0;   // upper triangles must be exact zeros
        }
        worst1 = fmax(worst1, e1); if (trial < 12) worst2 = fmax(worst2, e2);
        cycles = hc[0]; bad += hc[1];
    }
    printf("chol8_inv (NAGP_CHOL8_OLD=%d, BLOCKED=" "%d" "): %lld cycles/call, bad=%lld, max |LL^T-A|/|A| = %.2e, max |WL-I| (cond <= 1e11) = %.2e, %s\n",
           NAGP_CHOL8_OLD, NAGP_CHOL8_BLOCKED, cycles, bad, worst1, worst2, cudaGetErrorString(cudaGetLastError()));
    {
        double *sink; cudaMalloc(&sink, 8192);
        const char *mn[4] = {"idle", "DMMA", "LDS.128", "LDS.128+DMMA"};
        struct { int threads; unsigned mask; const char *what; } cfg[] = {
            {256, 0xfeu, "7 warps (all SMSPs)"}, {256, 0xeeu, "6 warps, none on the chain's SMSP"},
            {256, 0x10u, "1 warp on the chain's SMSP"}, {512, 0xfffeu, "15 warps"}, {512, 0x1110u, "3 warps on the chain's SMSP"},
            {512, 0xeeeeu, "12 warps, none on the chain's SMSP"}, {1024, 0xeeeeeeeeu, "24 warps, none on the chain's SMSP"}};
        for (int mode = 0; mode < 4; ++mode)
            for (auto &c : cfg) {
                if (mode == 0 && c.mask != 0xfeu) continue;
                bench_bg<<<1, c.threads>>>(A, cyc, 200, mode, mode == 0 ? 0u : c.mask, sink);
                cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
                printf("  background %-13s on %-36s: %lld cycles/call (%s)\n", mn[mode], mode == 0 ? "-" : c.what, hc[0], cudaGetErrorString(cudaGetLastError()));
            }
    }
    return 0;
}

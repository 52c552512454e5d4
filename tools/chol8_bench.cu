// chol8_bench.cu — isolates the in-register 8x8 diagonal-tile factorisation (+inverse) of the tile kernel
// to measure its latency per call on one warp and try variants. Not part of the product.
#include <cstdio>
#include <cuda_runtime.h>
constexpr unsigned kFull = 0xffffffffu;
__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(kFull, v, src); }

// variant 0: as in nagp_fused_v2.cu
__device__ __forceinline__ int chol8_v0(double &c0, double &c1, double &w0, double &w1, int lane, int nreal)
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

// fast reciprocal square root without the library's special-case call: MUFU seed + the same
// third-order correction (inputs are positive normal pivots; failures are flagged separately)
__device__ __forceinline__ double rsqrt_fast(double a)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-(y * y), a, 1.0);          // 1 - a y^2
    double t = fma(e, 0.375, 0.5);             // 1/2 + 3/8 e
    double ye = y * e;
    y = fma(t, ye, y);
    // one more Newton step for full double accuracy
    e = fma(-(y * y), a, 1.0);
    y = fma(0.5 * y, e, y);
    return y;
}

// variant 1: keeps the trailing block fully symmetric so every cross-lane value comes from row p
// (lanes 4p..4p+3), masks hoisted, fast rsqrt
__device__ __forceinline__ int chol8_v1(double &c0, double &c1, double &w0, double &w1, int lane, int nreal)
{
    const int r = lane >> 2, j = lane & 3;
    w0 = (r == 2 * j) ? 1.0 : 0.0;
    w1 = (r == 2 * j + 1) ? 1.0 : 0.0;
    int bad = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const int pl = p >> 1;
        const double d = shfl((p & 1) ? c1 : c0, p * 4 + pl);       // a_pp
        const double arp = shfl((r & 1) ? c1 : c0, p * 4 + (r >> 1));   // a_pr = a_rp (symmetric)
        const double ap0 = shfl(c0, p * 4 + j);                     // a_{p,2j}
        const double ap1 = shfl(c1, p * 4 + j);                     // a_{p,2j+1}
        const double wq0 = shfl(w0, p * 4 + j);
        const double wq1 = shfl(w1, p * 4 + j);
        if (!(d > 0.0) && p < nreal && bad == 0) bad = p + 1;
        const double rinv = rsqrt_fast(d);
        const double lrp = arp * rinv;
        const double lc0 = ap0 * rinv, lc1 = ap1 * rinv, wp0 = wq0 * rinv, wp1 = wq1 * rinv;
        if (r == p) { w0 = wp0; w1 = wp1; c0 = lc0; c1 = lc1; }       // row p becomes row p of L^T (finalised below)
        else if (r > p) {
            w0 = fma(-lrp, wp0, w0); w1 = fma(-lrp, wp1, w1);
            c0 = fma(-lrp, lc0, c0); c1 = fma(-lrp, lc1, c1);
        }
        if (j == pl && r > p) { if (p & 1) c1 = lrp; else c0 = lrp; }   // column p of L
    }
    // zero the strict upper triangle (rows kept L^T values there)
    if (2 * j > r) c0 = 0.0;
    if (2 * j + 1 > r) c1 = 0.0;
    return bad;
}


__device__ __forceinline__ double rcp_fast(double a)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    double e = fma(-a, x, 1.0);
    x = fma(x, e, x);
    e = fma(-a, x, 1.0);
    x = fma(x, e, x);
    return x;
}
__device__ __forceinline__ int op_idx(int r, int c) { return ((r * 4 + (c & 3)) << 1) + (c >> 2); }

// variant 2: every lane factors the whole tile redundantly in registers (no shuffles on the chain);
// the reciprocal pivot for the Schur update comes from its own MUFU+Newton chain so the rsqrt is off
// the critical path; lane c solves column c of the inverse. In/out through shared memory.
__device__ __forceinline__ int chol8_v2(double *sm_in, double *sm_L, double *sm_W, int lane, int nreal)
{
    double s[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) s[i][j] = sm_in[i * 8 + j];
    double rinv[8];
    int bad = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const double d = s[p][p];
        if (!(d > 0.0) && p < nreal && bad == 0) bad = p + 1;
        const double r = rcp_fast(d);
        const double ri = rsqrt_fast(d);
        rinv[p] = ri;
        double t[8];
#pragma unroll
        for (int i = p + 1; i < 8; ++i) t[i] = s[i][p] * r;
#pragma unroll
        for (int i = p + 1; i < 8; ++i)
#pragma unroll
            for (int j = p + 1; j <= i; ++j) s[i][j] = fma(-t[i], s[j][p], s[i][j]);
#pragma unroll
        for (int i = p + 1; i < 8; ++i) s[i][p] *= ri;
        s[p][p] = d * ri;
    }
    // column c of W = L^-1 (forward substitution on e_c)
    const int c = lane & 7;
    double w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double acc = (i == c) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) acc = fma(-s[i][k], w[k], acc);
        w[i] = acc * rinv[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) sm_L[op_idx(i, j)] = s[i][j];
    if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sm_W[op_idx(i, c)] = w[i];
    }
    return bad;
}

__global__ void bench2(const double *A, double *L, double *W, long long *cyc, int reps)
{
    __shared__ double sm_in[64], sm_L[64], sm_W[64];
    const int lane = threadIdx.x, r = lane >> 2, j = lane & 3;
    double a0 = A[r * 8 + 2 * j], a1 = A[r * 8 + 2 * j + 1];
    double c0 = a0, c1 = a1;
    int bad = 0;
    for (int i = lane; i < 64; i += 32) { sm_L[i] = 0; sm_W[i] = 0; }
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
        c0 = a0 + c0 * 1e-300; c1 = a1 + c1 * 1e-300;   // serialise calls
        *reinterpret_cast<double2 *>(sm_in + lane * 2) = make_double2(c0, c1);
        __syncwarp();
        bad += chol8_v2(sm_in, sm_L, sm_W, lane, 8);
        __syncwarp();
        double2 v = *reinterpret_cast<double2 *>(sm_W + lane * 2);   // consumer-side fragment load
        c0 = v.x; c1 = v.y;
    }
    long long t1 = clock64();
    __syncwarp();
    for (int e = lane; e < 64; e += 32) {
        int rr = e >> 3, cc = e & 7;
        L[e] = sm_L[op_idx(rr, cc)]; W[e] = sm_W[op_idx(rr, cc)];
    }
    if (lane == 0) { cyc[0] = (t1 - t0) / reps; cyc[1] = bad; }
}

template <int V>
__global__ void bench(const double *A, double *L, double *W, long long *cyc, int reps)
{
    const int lane = threadIdx.x, r = lane >> 2, j = lane & 3;
    double a0 = A[r * 8 + 2 * j], a1 = A[r * 8 + 2 * j + 1];
    double c0 = a0, c1 = a1, w0 = 0, w1 = 0;
    int bad = 0;
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
        c0 = a0 + c0 * 1e-300; c1 = a1 + c1 * 1e-300;   // serialise calls
        bad += V == 0 ? chol8_v0(c0, c1, w0, w1, lane, 8) : chol8_v1(c0, c1, w0, w1, lane, 8);
    }
    long long t1 = clock64();
    L[r * 8 + 2 * j] = c0; L[r * 8 + 2 * j + 1] = c1;
    W[r * 8 + 2 * j] = w0; W[r * 8 + 2 * j + 1] = w1;
    if (lane == 0) { cyc[0] = (t1 - t0) / reps; cyc[1] = bad; }
}

int main()
{
    double hA[64], hL[64], hW[64];
    for (int i = 0; i < 8; ++i) for (int k = 0; k < 8; ++k) hA[i * 8 + k] = 1.0 / (1 + abs(i - k)) + (i == k ? 2.0 : 0.0);
    double *A, *L, *W; long long *cyc, hc[2];
    cudaMalloc(&A, 512); cudaMalloc(&L, 512); cudaMalloc(&W, 512); cudaMalloc(&cyc, 16);
    cudaMemcpy(A, hA, 512, cudaMemcpyHostToDevice);
    for (int v = 0; v < 3; ++v) {
        if (v == 0) bench<0><<<1, 32>>>(A, L, W, cyc, 200); else if (v == 1) bench<1><<<1, 32>>>(A, L, W, cyc, 200); else bench2<<<1, 32>>>(A, L, W, cyc, 200);
        cudaMemcpy(hL, L, 512, cudaMemcpyDeviceToHost); cudaMemcpy(hW, W, 512, cudaMemcpyDeviceToHost);
        cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
        // check L L^T = A and W L = I
        double e1 = 0, e2 = 0;
        for (int i = 0; i < 8; ++i) for (int k = 0; k < 8; ++k) {
            double s = 0, t = 0;
            for (int m = 0; m < 8; ++m) { s += hL[i * 8 + m] * hL[k * 8 + m]; t += hW[i * 8 + m] * hL[m * 8 + k]; }
            e1 = fmax(e1, fabs(s - hA[i * 8 + k])); e2 = fmax(e2, fabs(t - (i == k)));
        }
        printf("variant %d: %lld cycles/call, bad=%lld, |LL^T-A|=%.2e |WL-I|=%.2e\n", v, hc[0], hc[1], e1, e2);
    }
    return 0;
}

// chol8_bench2.cu — latency per call and accuracy of chol8_inv (csrc/nagp_tile.cuh) on one warp, alone on the
// GPU. Build twice: -DNAGP_CHOL8_OLD=1 (first version) and =0 (short-chain version). Not part of the product.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../nowcastautogp_b200/csrc/nagp_tile.cuh"
using namespace nagp;

__global__ void bench(const double *A, double *L, double *W, long long *cyc, int reps)
{
    const int lane = threadIdx.x, r = lane >> 2, j = lane & 3;
    double a0 = A[r * 8 + 2 * j], a1 = A[r * 8 + 2 * j + 1];
    double c0 = a0, c1 = a1, w0 = 0, w1 = 0, piv[8];
    int bad = 0;
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
        c0 = a0 + c0 * 1e-300; c1 = a1 + c1 * 1e-300;   // serialise calls
        bad += chol8_inv(c0, c1, w0, w1, lane, 8, piv);
    }
    long long t1 = clock64();
    L[r * 8 + 2 * j] = c0; L[r * 8 + 2 * j + 1] = c1;
    W[r * 8 + 2 * j] = w0; W[r * 8 + 2 * j + 1] = w1;
    if (lane == 0) { cyc[0] = (t1 - t0) / reps; cyc[1] = bad; }
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
    printf("chol8_inv (NAGP_CHOL8_OLD=%d): %lld cycles/call, bad=%lld, max |LL^T-A|/|A| = %.2e, max |WL-I| (cond <= 1e11) = %.2e, %s\n",
           NAGP_CHOL8_OLD, cycles, bad, worst1, worst2, cudaGetErrorString(cudaGetLastError()));
    return 0;
}

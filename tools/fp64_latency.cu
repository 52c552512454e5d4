// fp64_latency.cu — dependent-chain latencies (cycles) of the FP64 building blocks on this GPU:
// DFMA, DADD, DMMA m8n8k4, SHFL(double), rsqrt(double), LDS.128. One warp, clock64 around N dependent ops.
#include <cstdio>
#include <cuda_runtime.h>

#define N 256
__global__ void lat_kernel(long long *out, double seed, double *sink)
{
    __shared__ double2 sm[64];
    const int lane = threadIdx.x;
    sm[lane] = make_double2(seed, seed); sm[lane + 32] = make_double2(seed, seed);
    __syncthreads();
    double x = seed + lane * 1e-9, y = 1.0000001;
    long long t0, t1;
    // DFMA
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
    t1 = clock64(); if (lane == 0) out[0] = t1 - t0;
    // DADD
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = x + y;
    t1 = clock64(); if (lane == 0) out[1] = t1 - t0;
    // DMMA dependent through accumulator
    double c0 = x, c1 = y;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(y), "d"(y));
    t1 = clock64(); if (lane == 0) out[2] = t1 - t0;
    // DMMA dependent through A operand (result feeds next A)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double d0 = 0, d1 = 0;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(c0), "d"(y));
        c0 = d0;
    }
    t1 = clock64(); if (lane == 0) out[3] = t1 - t0;
    // SHFL double
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
    t1 = clock64(); if (lane == 0) out[4] = t1 - t0;
    // rsqrt(double)
    x = fabs(x) + 1.0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.0;
    t1 = clock64(); if (lane == 0) out[5] = t1 - t0;
    // LDS.128 dependent (address from loaded value)
    int idx = lane;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) { double2 v = sm[idx]; idx = (idx + (int)v.x) & 63; }
    t1 = clock64(); if (lane == 0) out[6] = t1 - t0;
    // sqrt + div
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = sqrt(x) + 2.0;
    t1 = clock64(); if (lane == 0) out[7] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = 3.0 / x + 1.0;
    t1 = clock64(); if (lane == 0) out[8] = t1 - t0;
    // FFMA for reference
    float f = (float)x;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) f = fmaf(f, 1.0001f, 0.5f);
    t1 = clock64(); if (lane == 0) out[9] = t1 - t0;
    sink[lane] = x + c0 + c1 + idx + f;
}

int main()
{
    long long *out, h[10]; double *sink;
    cudaMalloc(&out, sizeof(h)); cudaMalloc(&sink, 32 * sizeof(double));
    for (int r = 0; r < 2; ++r) lat_kernel<<<1, 32>>>(out, 0.0, sink);
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const char *nm[10] = {"dfma", "dadd", "dmma_acc_chain", "dmma_operand_chain", "shfl_f64", "rsqrt_f64_plus_dadd",
                          "lds128_dependent", "sqrt_f64_plus_dadd", "div_f64_plus_dadd", "ffma"};
    printf("{");
    for (int i = 0; i < 10; ++i) printf("\"%s\": %.1f%s", nm[i], (double)h[i] / N, i < 9 ? ", " : "");
    printf("}\n");
    return 0;
}

// fp64_peak.cu — measures the FP64 roofline denominators MEASURED_PEAKS.json lacks: DFMA (vector)
// and DMMA (mma.sync f64) throughput on this GPU. Prints one JSON object.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double *out, int iters, double a, double b)
{
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void dmma884_kernel(double *out, int iters, double a, double b)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k8 f64 (sm_90+ shape): A 4 regs, B 2 regs, C/D 4 regs per thread
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

__global__ void dmma1688_kernel(double *out, int iters, double av, double bv)
{
    double c[6][4];
    double a[4] = {av, av, av, av}, b[2] = {bv, bv};
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + i + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 6; ++i) dmma1688(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F launch, int reps)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const int threads = 256, blocks = sms * 8, iters = 4096;
    double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    float t_fma = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 10);
    float t_884 = time_ms([&] { dmma884_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 10);
    float t_1688 = time_ms([&] { dmma1688_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 10);
    cudaError_t err = cudaDeviceSynchronize();
    double n_thr = (double)blocks * threads;
    double fl_fma = n_thr * iters * 8 * 2;
    double fl_884 = (n_thr / 32) * iters * 8 * (8 * 8 * 4 * 2);
    double fl_1688 = (n_thr / 32) * iters * 6 * (16 * 8 * 8 * 2);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.3f, \"dmma_m8n8k4_tflops\": %.3f, "
           "\"dmma_m16n8k8_tflops\": %.3f, \"status\": \"%s\"}\n",
           prop.name, sms, fl_fma / t_fma * 1e-9, fl_884 / t_884 * 1e-9, fl_1688 / t_1688 * 1e-9,
           cudaGetErrorString(err));
    return 0;
}

// Microbenchmark: FP64 DADD / DMUL / DFMA throughput and F2F conversions on sm_100a (what bounds the exact powf of
// CS_YCBCR).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe fp64_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double seed)
{
    double a[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = seed + i + threadIdx.x;
        f[i] = (float)(seed * i) + threadIdx.x;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0)
                a[i] = __dadd_rn(a[i], seed);
            else if (MODE == 1)
                a[i] = __dmul_rn(a[i], 1.0000001);
            else if (MODE == 2)
                a[i] = __fma_rn(a[i], 1.0000001, seed);
            else if (MODE == 3) // f32 -> f64 -> f32 round trip through an add
                f[i] = (float)__dadd_rn((double)f[i], seed);
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        r += a[i] + f[i];
    if (r == 123.456)
        out[0] = r;
}

template <int MODE>
void run(const char *name, double *d, int ops)
{
    const int iters = 4000, blocks = 148 * 4;
    k<MODE><<<blocks, 256>>>(d, 10, 1.5);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, iters, 1.5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double lane_ops = (double)blocks * 256 * iters * ops;
    printf("%-34s %8.3f ms  %6.2f lane-ops/clk/SM  (%5.2f SMSP-cycles per warp instruction)\n", name, ms,
           lane_ops / (ms * 1e-3) / 148 / 1.965e9, 32.0 / (lane_ops / (ms * 1e-3) / 148 / 1.965e9 / 4));
}

int main()
{
    double *d;
    cudaMalloc(&d, 8);
    run<0>("8 x DADD", d, 8);
    run<1>("8 x DMUL", d, 8);
    run<2>("8 x DFMA", d, 8);
    run<3>("8 x (F2F.64.32, DADD, F2F.32.64)", d, 8);
    return 0;
}

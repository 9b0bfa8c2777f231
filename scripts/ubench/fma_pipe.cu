// Microbenchmark: issue/pipe throughput of scalar FFMA vs packed FFMA2 / FMUL2 / FADD2 on sm_100a,
// and of a 50/50 mix (does the scalar stream use a pipe the packed one does not?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_pipe fma_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float seed)
{
    float2 a[8];
    float s[16];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        a[i] = make_float2(seed + i + threadIdx.x, seed - i);
#pragma unroll
    for (int i = 0; i < 16; ++i)
        s[i] = seed * i + threadIdx.x;
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(seed, -seed);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) { // 16 scalar FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i)
                s[i] = __fmaf_rn(s[i], 1.0001f, seed);
        } else if (MODE == 1) { // 8 FFMA2 (= 16 FMAs per lane)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                a[i] = __ffma2_rn(a[i], m, c);
        } else if (MODE == 2) { // 8 FMUL2
#pragma unroll
            for (int i = 0; i < 8; ++i)
                a[i] = __fmul2_rn(a[i], m);
        } else if (MODE == 3) { // 8 FADD2
#pragma unroll
            for (int i = 0; i < 8; ++i)
                a[i] = __fadd2_rn(a[i], c);
        } else if (MODE == 4) { // 4 FFMA2 + 8 FFMA (= 16 FMAs per lane)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                a[i] = __ffma2_rn(a[i], m, c);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                s[i] = __fmaf_rn(s[i], 1.0001f, seed);
        } else if (MODE == 5) { // 16 scalar FMUL
#pragma unroll
            for (int i = 0; i < 16; ++i)
                s[i] = __fmul_rn(s[i], 1.0001f);
        } else if (MODE == 6) { // 16 scalar FADD
#pragma unroll
            for (int i = 0; i < 16; ++i)
                s[i] = __fadd_rn(s[i], seed);
        } else if (MODE == 7) { // 16 FMNMX
#pragma unroll
            for (int i = 0; i < 16; ++i)
                s[i] = fmaxf(fminf(s[i], 1e8f + i), 1e-4f * (it + 1));
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        r += a[i].x + a[i].y;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        r += s[i];
    if (r == 123.456f)
        out[0] = r;
}

template <int MODE>
void run(const char *name, float *d, int fmas_per_iter)
{
    const int iters = 20000, blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(d, 100, 1.5f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, iters, 1.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // per SMSP: warps = blocks*8/ (148*4); lane-ops per clk
    const double lane_ops = (double)blocks * 256 * iters * fmas_per_iter;
    const double per_clk_sm = lane_ops / (ms * 1e-3) / 148 / 1.965e9;
    printf("%-28s %8.3f ms  %7.1f lane-ops/clk/SM (at 1.965 GHz)\n", name, ms, per_clk_sm);
}

int main()
{
    float *d;
    cudaMalloc(&d, 4);
    run<0>("16 x FFMA", d, 16);
    run<1>("8 x FFMA2", d, 16);
    run<2>("8 x FMUL2", d, 16);
    run<3>("8 x FADD2", d, 16);
    run<4>("4 x FFMA2 + 8 x FFMA", d, 16);
    run<5>("16 x FMUL", d, 16);
    run<6>("16 x FADD", d, 16);
    run<7>("16 x (FMNMX,FMNMX)", d, 32);
    return 0;
}

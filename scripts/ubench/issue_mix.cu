// Microbenchmark: does a packed FFMA2 take one issue slot (pipe busy 2 cycles) or two?  Mixes packed FP32 with
// independent integer / min-max / MUFU work and reports cycles per loop iteration per SMSP (8 warps resident).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float seed, unsigned iseed)
{
    float2 a[8];
    unsigned u[8];
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = make_float2(seed + i + threadIdx.x, seed - i);
        u[i] = iseed * (i + 1) + threadIdx.x;
        s[i] = seed * (i + 2) + threadIdx.x;
    }
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(seed, -seed);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 4 || MODE == 6) { // 8 FFMA2
#pragma unroll
            for (int i = 0; i < 8; ++i)
                a[i] = __ffma2_rn(a[i], m, c);
        }
        if (MODE == 1 || MODE == 2) { // 8 LOP3 (xor-and mix)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                u[i] = (u[i] ^ iseed) & (u[(i + 1) & 7] | 0x55u);
        }
        if (MODE == 3 || MODE == 5) { // 8 FMNMX pairs (clamp)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                s[i] = fmaxf(fminf(s[i], 1e8f + i), 1e-4f * (it + 1));
        }
        if (MODE == 4 || MODE == 7) { // 4 MUFU.RCP
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(s[i]));
        }
        if (MODE == 6 || MODE == 8) { // 16 scalar FFMA
#pragma unroll
            for (int i = 0; i < 8; ++i)
                s[i] = __fmaf_rn(s[i], 1.0001f, seed);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                s[i] = __fmaf_rn(s[i], 0.9999f, seed);
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        r += a[i].x + a[i].y + s[i] + (float)u[i];
    if (r == 123.456f)
        out[0] = r;
}

template <int MODE>
void run(const char *name, float *d)
{
    const int iters = 20000, blocks = 148 * 4; // 4 blocks x 8 warps per SM = 8 warps per SMSP
    k<MODE><<<blocks, 256>>>(d, 100, 1.5f, 12345u);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, iters, 1.5f, 12345u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9 / iters / 8.0; // cycles per warp-iteration per SMSP
    printf("%-40s %8.3f ms  %6.2f SMSP-cycles per warp-iteration\n", name, ms, cyc);
}

int main()
{
    float *d;
    cudaMalloc(&d, 4);
    run<0>("8 FFMA2", d);
    run<1>("8 LOP3", d);
    run<2>("8 FFMA2 + 8 LOP3", d);
    run<5>("8 x (FMNMX,FMNMX)", d);
    run<3>("8 FFMA2 + 8 x (FMNMX,FMNMX)", d);
    run<7>("4 MUFU.RCP", d);
    run<4>("8 FFMA2 + 4 MUFU.RCP", d);
    run<8>("16 FFMA", d);
    run<6>("8 FFMA2 + 16 FFMA", d);
    return 0;
}

/*
 * powf_reach.c -- can the transform reach the one input where an UNFUSED replay of glibc's powf differs from the host libm
 * for y = 1/0.1593f (x = 0x1.7b1e06p-11, scripts/powchk.c)?  PQ decode calls powf(q, 1/0.1593f) with
 * q = max(0, Vp - c1) / (c2 - c3 Vp), Vp = powf(v, 1/78.8438f), and v is always clamped to [0, 1]
 * (src/luma_quantizer.cpp:453-459, :340): enumerate every float v in (0, 1] and count those whose q has that bit pattern.
 *
 *     gcc -O2 -ffp-contract=off -fopenmp scripts/powf_reach.c -lm -o /tmp/powf_reach && /tmp/powf_reach
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
static inline uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
int main(void)
{
    const float m = 78.8438f, c1 = 0.8359f, c2 = 18.8516f, c3 = 18.6875f;
    const float inv_m = 1.0f / m;
    uint64_t hits = 0; 
#pragma omp parallel for reduction(+:hits) schedule(static)
    for (int64_t u = 1; u <= (int64_t)f2u(1.0f); u++) {
        float v = u2f((uint32_t)u);
        float Vp = powf(v, inv_m);
        float num = fmaxf(0.0f, Vp - c1);
        float den = c2 - c3 * Vp;
        float x = num / den;
        if (f2u(x) == 0x3a3d8f03u) {
            hits++;
#pragma omp critical
            if (hits < 5) printf("v = %a (0x%08x) -> Vp %a -> x %a\n", v, (unsigned)u, Vp, x);
        }
    }
    printf("floats v in (0,1] whose PQ-decode inner quotient is 0x1.7b1e06p-11: %llu\n", (unsigned long long)hits);
    return 0;
}

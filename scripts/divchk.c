/*
 * divchk.c -- exhaustive check of the 3-instruction constant division used by the kernels
 * (lumahdrv_b200/csrc/luma_device.cuh div_const_int<D>, luma_fast.cuh div_const2<D>):
 *
 *     rc = RN(1/d);  q = RN(x*rc);  r = fma(-q, d, x);  q' = fma(r, rc, q)
 *
 * against IEEE x / d for EVERY float bit pattern x (NaN results compared as "both NaN").
 *
 *     gcc -O2 -ffp-contract=off -fopenmp scripts/divchk.c -lm -o /tmp/divchk && /tmp/divchk 255
 *
 * Prints the number of mismatching inputs per divisor, split into inputs whose quotient is a normal
 * number (must be 0: the kernels only use the sequence there -- operands are O(1) chromaticities
 * times 410 or 1640) and the rest (subnormal quotients / overflow, where a plain FMA sequence
 * legitimately differs and the kernels never go).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline float u2f(uint32_t u)
{
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline uint32_t f2u(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

int main(int argc, char **argv)
{
    int rc_all = 0;
    for (int a = 1; a < argc; a++) {
        const float d = (float)atof(argv[a]);
        const float rc = 1.0f / d;
        uint64_t bad_normal = 0, bad_other = 0;
#pragma omp parallel for reduction(+ : bad_normal, bad_other) schedule(static)
        for (int64_t i = 0; i < ((int64_t)1 << 32); i++) {
            const float x = u2f((uint32_t)i);
            const float want = x / d;
            const float q = x * rc;
            const float r = fmaf(-q, d, x);
            const float got = fmaf(r, rc, q);
            if (f2u(want) == f2u(got) || (want != want && got != got))
                continue;
            if (isnormal(want) && isfinite(x))
                bad_normal++;
            else
                bad_other++;
        }
        printf("d = %g: %llu mismatches with a normal quotient, %llu elsewhere (subnormal quotient / inf)\n", (double)d,
               (unsigned long long)bad_normal, (unsigned long long)bad_other);
        if (bad_normal)
            rc_all = 1;
    }
    return rc_all;
}

/*
 * thrchk.c -- exhaustive proof that the decision thresholds the kernels search (lumacu_derive_thresholds, the host half of
 * lumacu_set_quantizer) reproduce LumaQuantizer::quantize's luma branch (src/luma_quantizer.cpp:219-235: bisect the LUT,
 * then pick the nearer neighbour with fp32 differences) for EVERY non-NaN float, not just the sampled ones of the tests.
 *
 * For each LUT: code_ref(x) = the reference's loop (restated below, the same lines as oracle/luma_oracle.c lo_quantize);
 * code_thr(x) = number of thresholds whose ordered key is <= key(x).  All 2^32 bit patterns except NaNs (the kernels
 * special-case NaN -> max code, as the reference's comparisons do).
 *
 *     gcc -O2 -ffp-contract=off -fopenmp -Iinclude scripts/thrchk.c -Llumahdrv_b200 -llumacu -Wl,-rpath,$PWD/lumahdrv_b200 -lm -o /tmp/thrchk
 *     /tmp/thrchk            (every shipped transfer function at its usual depths; ~30 s per LUT on 8 cores; no GPU needed)
 *
 * Output of this container: profiles/r02_thrchk.log.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lumacu.h"

static inline float u2f(uint32_t u)
{
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline uint32_t okey(uint32_t bits) { return (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u); }

static int check(const char *name, int ptf, unsigned bits, float lmax, float lmin)
{
    const uint32_t n = 1u << bits;
    float *lut = malloc(n * sizeof(float));
    uint32_t *thr = malloc(n * sizeof(uint32_t));
    if (lumacu_build_lut(ptf, bits, lmax, lmin, lut, n) != 0) {
        printf("%s-%u: no table (PTF tables not compiled in)\n", name, bits);
        return 0;
    }
    if (!lumacu_derive_thresholds(lut, n, thr)) {
        printf("%s-%u: LUT not strictly increasing: the kernels replay the reference loop literally (nothing to prove)\n", name, bits);
        return 0;
    }
    const int max_val = (int)n - 1;
    uint64_t bad = 0, total = 0;
#pragma omp parallel for reduction(+ : bad, total) schedule(static)
    for (int64_t i = 0; i < ((int64_t)1 << 32); i++) {
        const uint32_t b = (uint32_t)i;
        if ((b & 0x7fffffffu) > 0x7f800000u)
            continue; /* NaN */
        const float val = u2f(b);
        int lo = 0, hi = max_val;
        while (lo + 1 < hi) {
            const int mid = (lo + hi) / 2;
            if (val < lut[mid])
                hi = mid;
            else
                lo = mid;
        }
        const int ref = (val - lut[lo] < lut[hi] - val) ? lo : hi;
        /* number of thresholds <= key */
        const uint32_t k = okey(b);
        int a = 0, z = max_val; /* thr[0 .. max_val-1] ascending */
        while (a < z) {
            const int m = (a + z) / 2;
            if (thr[m] <= k)
                a = m + 1;
            else
                z = m;
        }
        total++;
        bad += (a != ref);
    }
    printf("%s-%u (Lmax %g, Lmin %g): %llu floats, %llu where the threshold count differs from the reference's search\n", name, bits,
           (double)lmax, (double)lmin, (unsigned long long)total, (unsigned long long)bad);
    free(lut);
    free(thr);
    return bad != 0;
}

int main(void)
{
    int rc = 0;
    /* lumacu_ptf: PSI 0, PQ 1, LOG 2, JND_HDRVDP 3, LINEAR 4 (the reference's enum order) */
    rc |= check("PQ", 1, 11, 1e4f, 0.005f);
    rc |= check("PQ", 1, 10, 1e4f, 0.005f);
    rc |= check("PQ", 1, 10, 1000.0f, 0.01f);
    rc |= check("PQ", 1, 12, 1e4f, 0.005f);
    rc |= check("PQ", 1, 8, 1e4f, 0.005f);
    rc |= check("PQ", 1, 16, 1e4f, 0.005f);
    rc |= check("LOG", 2, 12, 1e4f, 0.005f);
    rc |= check("LOG", 2, 11, 1e4f, 0.005f);
    rc |= check("PSI", 0, 11, 1e4f, 0.005f);
    rc |= check("JND_HDRVDP", 3, 12, 1e4f, 0.005f);
    rc |= check("LINEAR", 4, 11, 1e4f, 0.005f);
    return rc;
}

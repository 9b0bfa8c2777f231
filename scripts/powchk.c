/*
 * powchk.c -- exhaustive comparison of the device powf restatement (lumahdrv_b200/csrc/powf_glibc.cuh) with the host libm.
 *
 * The device function replays glibc's powf (sysdeps/ieee754/flt-32/e_powf.c) in IEEE double.  glibc ships two builds of
 * that file and picks one at load time (ifunc): the plain one and, on CPUs with FMA + AVX2, one compiled with -mfma,
 * where a*b+c becomes a fused multiply-add.  This program evaluates both contraction choices on the CPU -- replay_nofma
 * (every multiply and add rounded separately) and replay_fma (every a*b+c fused; the sequence the device executes, DFMA
 * being the same IEEE operation as fma()) -- for EVERY float in [lo, hi] and one exponent, and counts the inputs whose
 * result differs from powf(x, y) of the libm it is linked with.
 *
 *     gcc -O2 -ffp-contract=off -fopenmp scripts/powchk.c -lm -o /tmp/powchk
 *     for e in 0.1593 inv:78.8438 inv:0.1593 78.8438; do /tmp/powchk $e 1.1754944e-38 3.4028234e38; done
 *
 * ("inv:X" = 1.0f / Xf, evaluated in float like the reference's 1.0f/n.)  About 15 s per exponent on 8 cores.
 * Output of this container (glibc 2.39-0ubuntu8.5, Xeon with FMA + AVX2): profiles/r02_powf_exhaustive.log.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
static inline uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline double u2d(uint64_t u){double f;memcpy(&f,&u,8);return f;}
static inline uint64_t d2u(double f){uint64_t u;memcpy(&u,&f,8);return u;}
static const double LT[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2}, {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2},
    {0x1.49539f0f010bp+0, -0x1.7418b0a1fb77bp-2},  {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2},
    {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2}, {0x1.25e227b0b8eap+0, -0x1.97c1d1b3b7afp-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3}, {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4},
    {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4},  {0x1.ca4b31f026aap-1, 0x1.476a9543891bap-3},
    {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2},
    {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2},  {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2}};
static const uint64_t ET[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
static const double A[5] = {0x1.27616c9496e0bp-2, -0x1.71969a075c67ap-2, 0x1.ec70a6ca7baddp-2, -0x1.7154748bef6c8p-1, 0x1.71547652ab82bp+0};
static const double Cc[3] = {0x1.c6af84b912394p-5, 0x1.ebfce50fac4f3p-3, 0x1.62e42ff0c52d6p-1};
#define SHIFT 0x1.8p+47
static double g_ylogx;
#pragma omp threadprivate(g_ylogx)
// no-FMA replay (what powf_glibc.cuh does); returns the double before rounding
static inline double replay_nofma(uint32_t ix, double y)
{
    uint32_t tmp = ix - 0x3f330000u; int i = (tmp >> 19) & 15; uint32_t top = tmp & 0xff800000u; uint32_t iz = ix - top; int k = (int32_t)top >> 23;
    double invc = LT[i][0], logc = LT[i][1], z = (double)u2f(iz);
    volatile double t;
    double r = z * invc; t = r; r = t - 1.0;
    double y0 = logc + (double)k;
    double r2 = r * r;
    double yy = A[0] * r; t = yy; yy = t + A[1];
    double p = A[2] * r; t = p; p = t + A[3];
    double r4 = r2 * r2;
    double q = A[4] * r; t = q; q = t + y0;
    double pq = p * r2; t = pq; q = t + q;
    double yr = yy * r4; t = yr; yy = t + q;
    double ylogx = y * yy; g_ylogx = ylogx;
    double kd = ylogx + SHIFT; uint64_t ki = d2u(kd); kd -= SHIFT; double rr = ylogx - kd;
    uint64_t tt = ET[ki & 31]; tt += ki << 47; double s = u2d(tt);
    double zz = Cc[0] * rr; t = zz; zz = t + Cc[1];
    double rr2 = rr * rr;
    double o = Cc[2] * rr; t = o; o = t + 1.0;
    double zr = zz * rr2; t = zr; o = t + o;
    return o * s;
}
// FMA everywhere a*b+c appears (lean device candidate)
static inline double replay_fma(uint32_t ix, double y)
{
    uint32_t tmp = ix - 0x3f330000u; int i = (tmp >> 19) & 15; uint32_t top = tmp & 0xff800000u; uint32_t iz = ix - top; int k = (int32_t)top >> 23;
    double invc = LT[i][0], logc = LT[i][1], z = (double)u2f(iz);
    double r = fma(z, invc, -1.0);
    double y0 = logc + (double)k;
    double r2 = r * r;
    double yy = fma(A[0], r, A[1]);
    double p = fma(A[2], r, A[3]);
    double r4 = r2 * r2;
    double q = fma(A[4], r, y0);
    q = fma(p, r2, q);
    yy = fma(yy, r4, q);
    double ylogx = y * yy; g_ylogx = ylogx;
    double kd = ylogx + SHIFT; uint64_t ki = d2u(kd); kd -= SHIFT; double rr = ylogx - kd;
    uint64_t tt = ET[ki & 31]; tt += ki << 47; double s = u2d(tt);
    double zz = fma(Cc[0], rr, Cc[1]);
    double rr2 = rr * rr;
    double o = fma(Cc[2], rr, 1.0);
    o = fma(zz, rr2, o);
    return o * s;
}
static inline float finish(double o, double ylogx)
{
    float res = (float)o;
    uint32_t hi = (uint32_t)(d2u(ylogx) >> 32) & 0x7fff8000u;
    if (hi >= 0x405f8000u) {
        if (ylogx > 0x1.fffffffd1d571p+6) res = u2f(0x7f800000u);
        else if (ylogx <= -150.0) res = 0.0f;
        else if (ylogx < -149.0) res = u2f(1u);
    }
    return res;
}
int main(int argc, char **argv)
{
    float y = strncmp(argv[1], "inv:", 4) == 0 ? 1.0f / (float)atof(argv[1] + 4) : (float)atof(argv[1]);
    uint32_t lo = f2u((float)atof(argv[2])), hi = f2u((float)atof(argv[3]));
    uint64_t n = 0, bad_nofma = 0, bad_fma = 0;
    (void)argc;
#pragma omp parallel for reduction(+:n,bad_nofma,bad_fma) schedule(static)
    for (int64_t u = lo; u <= (int64_t)hi; u++) {
        float x = u2f((uint32_t)u);
        float want = powf(x, y);
        double a = replay_nofma((uint32_t)u, (double)y); double yla = g_ylogx;
        double b = replay_fma((uint32_t)u, (double)y); double ylb = g_ylogx;
        n++;
        if (f2u(finish(a, yla)) != f2u(want)) {
            bad_nofma++;
#pragma omp critical
            printf("    no-FMA replay differs: x = %a (0x%08x)  libm %a  replay %a\n", x, (unsigned)u, want, finish(a, yla));
        }
        if (f2u(finish(b, ylb)) != f2u(want)) {
            bad_fma++;
#pragma omp critical
            printf("    FMA replay differs: x = %a (0x%08x)  libm %a  replay %a\n", x, (unsigned)u, want, finish(b, ylb));
        }
    }
    printf("y=%.9g x in [%g, %g]: %llu inputs; no-FMA replay != libm: %llu; FMA replay != libm: %llu\n", y, u2f(lo), u2f(hi),
           (unsigned long long)n, (unsigned long long)bad_nofma, (unsigned long long)bad_fma);
    return 0;
}

/*
 * powf_glibc.cuh -- device restatement of the host libm's powf.
 *
 * The reference evaluates PQ per pixel for CS_YCBCR through libm powf
 * (/root/reference/src/luma_quantizer.cpp:493-499).  The reference neither
 * vendors nor pins a libm: results are "whatever glibc returns".  glibc's powf
 * (2.28 ... 2.39, sysdeps/ieee754/flt-32/e_powf.c, derived from ARM's
 * optimized-routines) is < 1 ulp but not correctly rounded, so neither CUDA's
 * powf nor a double-precision pow reproduces its bits.  This header restates
 * the published algorithm: log2(x) from a 16-entry {1/c, log2 c} table plus a
 * degree-5 polynomial, times y, then exp2 through a 32-entry 2^(i/32) table
 * and a cubic, all in IEEE double with every a*b+c FUSED, which is what the
 * host executes (glibc's ifunc picks its FMA build on every CPU with FMA and
 * AVX2).  scripts/powchk.c proves it by exhaustion: equal to the host powf
 * (glibc 2.39-0ubuntu8.5) for every positive normal float and each of the four
 * PQ exponents, 8.5e9 evaluations, 0 mismatches
 * (profiles/r02_powf_exhaustive.log).  tests/test_gpu_parity.py::
 * test_ycbcr_powf_dense checks the device function against the host libm
 * through the YCbCr transform on ~3e7 evaluations spanning 1e-12 .. 1e6 plus
 * the special values.
 *
 * Restricted to what the PQ call sites need: y is a finite, positive,
 * non-integer constant (0.1593f, 78.8438f and their fp32 reciprocals); x is
 * any float.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace lumacu {

/* __powf_log2_data (POWF_LOG2_TABLE_BITS = 4, POWF_SCALE_BITS = 0) */
static __device__ const double2 k_powf_log2_tab[16] = {
    {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2}, {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2},
    {0x1.49539f0f010bp+0, -0x1.7418b0a1fb77bp-2},  {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2},
    {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2}, {0x1.25e227b0b8eap+0, -0x1.97c1d1b3b7afp-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3}, {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4},
    {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4},  {0x1.ca4b31f026aap-1, 0x1.476a9543891bap-3},
    {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2},
    {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2},  {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2}};

/* __exp2f_data.tab (EXP2F_TABLE_BITS = 5): bits of 2^(i/32) with the exponent
 * contribution i << 47 subtracted */
static __device__ const unsigned long long k_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

/* Per-block shared-memory copies of the two tables: the lookups are data dependent (every lane its own entry), so
 * constant memory would serialise them (measured: 2.5x slower) and global loads cost a 64-bit address, a descriptor
 * and an L1 round trip each (~10 instructions of a ~75-instruction powf); LDS with a uniform base is one instruction.
 * Every kernel that can reach powf_glibc stages them once with powf_tables_stage(). */
static __shared__ double2 s_powf_log2_tab[16];
static __shared__ unsigned long long s_exp2f_tab[32];

__device__ __forceinline__ void powf_tables_stage() /* whole block; blockDim.x >= 32 */
{
    if (threadIdx.x < 16)
        s_powf_log2_tab[threadIdx.x] = k_powf_log2_tab[threadIdx.x];
    if (threadIdx.x < 32)
        s_exp2f_tab[threadIdx.x] = k_exp2f_tab[threadIdx.x];
    __syncthreads();
}

/* Polynomial coefficients live in constant memory so that DMUL / DADD take them as constant-bank operands; as
 * literals the compiler rebuilt each 64-bit value with two UMOVs at every use (15 % of the instruction stream).
 * [0..4] = __powf_log2_data.poly (A), [5..7] = __exp2f_data.poly_scaled (C). */
static __constant__ double k_powf_poly[8] = {0x1.27616c9496e0bp-2, -0x1.71969a075c67ap-2, 0x1.ec70a6ca7baddp-2,
                                             -0x1.7154748bef6c8p-1, 0x1.71547652ab82bp+0,
                                             0x1.c6af84b912394p-5, 0x1.ebfce50fac4f3p-3,  0x1.62e42ff0c52d6p-1};

/* Everything e_powf.c does before / instead of the main path: x zero, inf, nan, negative or subnormal.  Out of
 * line and rare: in the PQ call sites x is a clamped positive number except for 0 (LUT entry 0) and NaN pixels. */
static __device__ __noinline__ float powf_glibc_rare(float x, float y);

/* Straight-line main path: rare inputs are detected up front and recomputed out of line at the end, results that
 * overflow or underflow only override the value, so that the compiler can interleave the independent powf chains
 * of a pixel (three PQ curves at a time) instead of serialising them behind branches.
 *
 * Every a*b+c of the published algorithm is ONE fused multiply-add here.  That is what the host executes: glibc selects
 * its FMA build of this function (__powf_fma, sysdeps/x86_64/fpu/multiarch/e_powf.c) on every CPU with FMA + AVX2 --
 * every B200 host -- and scripts/powchk.c compares both contraction choices with the host powf for EVERY positive normal
 * float and the four PQ exponents (4 x 2 130 706 432 inputs, profiles/r02_powf_exhaustive.log): the fused sequence
 * equals libm on all of them; the unfused one (rounds 1 and early 2 of this file, 27 operations instead of 18) differs
 * for one input each of y = 1/0.1593f (x = 0x1.7b1e06p-11; no v in (0,1] leads PQ decode there) and y = 78.8438f (x > 1, not
 * reachable).
 * RANGE_CHECK = false: for callers whose exponent cannot overflow or underflow a positive normal x (|y| < 0.85:
 * 0.1593f and 1/78.8438f), the |y log2 x| >= 126 test is dropped. */
template <bool RANGE_CHECK = true>
__device__ __forceinline__ float powf_glibc_main(uint32_t ix, float y)
{
    /* log2_inline: x = 2^k z, z in [OFF, 2 OFF), c near the centre of z's subinterval */
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> (23 - 4)) & 15u);
    const uint32_t top = tmp & 0xff800000u;
    const uint32_t iz = ix - top;
    const int k = (int32_t)top >> 23;
    const double2 tc = s_powf_log2_tab[i];
    const double z = (double)__uint_as_float(iz);

    const double A0 = k_powf_poly[0], A1 = k_powf_poly[1], A2 = k_powf_poly[2], A3 = k_powf_poly[3], A4 = k_powf_poly[4];
    const double r = __fma_rn(z, tc.x, -1.0);
    const double y0 = __dadd_rn(tc.y, (double)k);
    const double r2 = __dmul_rn(r, r);
    double yy = __fma_rn(A0, r, A1);
    const double p = __fma_rn(A2, r, A3);
    const double r4 = __dmul_rn(r2, r2);
    double q = __fma_rn(A4, r, y0);
    q = __fma_rn(p, r2, q);
    yy = __fma_rn(yy, r4, q);

    const double ylogx = __dmul_rn((double)y, yy);

    /* exp2_inline: x = k/N + r, |r| <= 1/(2N), N = 32 */
    const double SHIFT = 0x1.8p+47; /* 0x1.8p52 / N */
    const double C0 = k_powf_poly[5], C1 = k_powf_poly[6], C2 = k_powf_poly[7];
    double kd = __dadd_rn(ylogx, SHIFT);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dadd_rn(kd, -SHIFT);
    const double rr = __dadd_rn(ylogx, -kd);
    unsigned long long t = s_exp2f_tab[ki & 31u];
    t += ki << (52 - 5);
    const double s = __longlong_as_double((long long)t);
    const double zz = __fma_rn(C0, rr, C1);
    const double rr2 = __dmul_rn(rr, rr);
    double o = __fma_rn(C2, rr, 1.0);
    o = __fma_rn(zz, rr2, o);
    o = __dmul_rn(o, s);
    float res = __double2float_rn(o);

    if (RANGE_CHECK) {
        const uint32_t hi = (uint32_t)__double2hiint(ylogx) & 0x7fff8000u;
        if (__builtin_expect(hi >= 0x405f8000u, 0)) { /* |y log2 x| >= 126 */
            if (ylogx > 0x1.fffffffd1d571p+6)
                res = __int_as_float(0x7f800000); /* overflow */
            else if (ylogx <= -150.0)
                res = 0.0f; /* underflow */
            else if (ylogx < -149.0)
                res = __int_as_float(0x00000001); /* __math_may_uflowf: 0x1.4p-75f squared */
        }
    }
    return res;
}

/* SMALL_EXPONENT: y is one of the two PQ exponents below 1 in magnitude (see RANGE_CHECK above).  Subnormal x still goes
 * through the checked path (out of line). */
template <bool SMALL_EXPONENT = false>
__device__ __forceinline__ float powf_glibc(float x, float y)
{
    const uint32_t ix = __float_as_uint(x);
    const bool rare = ix - 0x00800000u >= 0x7f800000u - 0x00800000u; /* not a positive normal number */
    float res = powf_glibc_main<!SMALL_EXPONENT>(rare ? 0x3f800000u : ix, y);
    if (__builtin_expect(rare, 0))
        res = powf_glibc_rare(x, y);
    return res;
}

static __device__ __noinline__ float powf_glibc_rare(float x, float y)
{
    uint32_t ix = __float_as_uint(x);
    /* x < 0x1p-126, or inf, or nan (e_powf.c "zeroinfnan (ix)" and below) */
    if (2u * ix - 1u >= 2u * 0x7f800000u - 1u)
        return __fmul_rn(x, x); /* +-0 -> +0, +-inf -> +inf, nan -> nan (y > 0, not an odd integer) */
    if (ix & 0x80000000u)
        return __int_as_float(0x7fffffff); /* finite x < 0, non-integer y: invalid */
    /* normalise a subnormal x so that the exponent goes negative */
    ix = __float_as_uint(__fmul_rn(x, 0x1p23f));
    ix &= 0x7fffffffu;
    ix -= 23u << 23;
    return powf_glibc_main<true>(ix, y);
}

} // namespace lumacu

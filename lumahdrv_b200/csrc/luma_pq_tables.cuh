/*
 * luma_pq_tables.cuh -- exhaustive, L2-resident tables of the two PQ curves for CS_YCBCR.
 *
 * CS_YCBCR evaluates PQ per pixel through libm powf (reference src/luma_quantizer.cpp:331-337,447-459,485-501): eight
 * calls per pixel each way.  powf_glibc.cuh reproduces the host libm bit for bit, but one call is ~45 instructions
 * (18 of them FP64, at a quarter of the fp32 rate).  Two of the call sites have a SMALL domain once you look at what feeds them, so their
 * results can be tabulated for EVERY possible input, with the exact device powf, once per quantizer:
 *
 *  (1) PQ decode, L * powf(max(0, Vp - c1) / (c2 - c3 Vp), 1/n) with Vp = powf(v, 1/m), is only ever applied to
 *      v = clamp01(.) (inverse transform, :453-459) or to v = (219 y' + 16)/255 (forward, :340).  The exponent 1/m =
 *      0.0127 makes Vp a staircase: ~40-79 consecutive floats v share one Vp, hence one result.  For every float v in
 *      [2^-8, 1] the table holds, per bucket of 32 consecutive floats, the result below the step, the result above it
 *      and the position of the step (pqd: 2.1 M buckets x 16 B = 33.5 MB).  The builder evaluates Vp for ALL 2^26 floats
 *      of the range and marks a bucket irregular (lookup falls back to the exact evaluation) unless its 32 values are
 *      "one value, then a second one" -- so glibc's rare non-monotonic roundings at a step edge cannot be misrepresented.
 *  (2) the outer power of PQ encode, powf(b, m) with b = (c1 + c2 Lp) / (1 + c3 Lp): whatever Lp is, b lies in
 *      [c1, c2/c3] = [0.8359, 1.00878] -- 2.85 M floats.  pqe holds powf(b, m) for each of them (11.4 MB).
 *
 *  (4) pqh: the reference's frames come from OpenEXR half-float pixels (src/exr_interface.cpp:73-143), i.e. every input
 *      sample is one of 65 536 values.  pqh holds the complete PQ encode R' = PQenc(max(c * preScaling, 1e-10)) for every
 *      half bit pattern (256 KB, L1/L2 resident, rebuilt when preScaling or Lmax change: one small launch).  A sample
 *      that is a normal half-float value or zero is looked up; any other float is evaluated as before.  For
 *      EXR-sourced content the forward path then contains no powf at all.
 *
 * Both tables fit in the 126 MB L2 many times over; a lookup is one 32-byte sector from L2 (or L1, for neighbouring
 * pixels of natural images) instead of 110-220 instructions.  The inner power of PQ encode, Lp = powf(x / Lmax, n),
 * has an unbounded domain and stays an exact evaluation.  Every table entry is produced by powf_glibc itself, so a
 * table lookup returns exactly what the per-pixel evaluation returns -- tests run both (tuned kernels: tables;
 * generic kernels: per-pixel powf) against the oracle.
 */
#pragma once

#include <cuda_fp16.h>

#include "luma_device.cuh"

namespace lumacu {

constexpr uint32_t kPqdKey0 = 0x3B800000u;                /* 2^-8 */
constexpr uint32_t kPqdKeys = 0x3F800000u - kPqdKey0 + 1u; /* ... 1.0f inclusive */
constexpr uint32_t kPqdBuckets = (kPqdKeys + 31u) / 32u;
constexpr uint32_t kPqeKey0 = 0x3F55C28Fu;                /* 0.835f */
constexpr uint32_t kPqeKeys = 0x3F8147AEu - kPqeKey0 + 1u; /* ... 1.01f inclusive */

/* (3) forward path, luma: code = quantize(PQdec(v), 0) with v = (219 y' + 16)/255 is a monotone step function of the float
 *     v with one step per code (1023 of them for HDR10).  vdtab is a direct search table keyed on the bits of v, in the
 *     format of the luminance-keyed one (luma_fast.cuh DirectSearch): buckets of 2^13 consecutive floats over
 *     [2^-5, 1], at most one step per bucket, 20 KB of shared memory -- the luma code of a pixel costs a clamp, a shift, a
 *     shared-memory read and an add instead of an L2 lookup of PQdec(v) plus the luminance search.  Built on the device
 *     from the exact pieces (pqd / powf_glibc + the quantizer's own search), each step located by bisection and its
 *     neighbourhood checked for strict monotonicity; if anything is irregular (two steps in a bucket: more than ~10
 *     luma bits; a non-monotonic neighbourhood) the table is not used. */
constexpr uint32_t kVdShift = 13u;
constexpr uint32_t kVdLoBucket = 0x3D000000u >> kVdShift; /* 2^-5 */
constexpr uint32_t kVdHiBucket = 0x3F800000u >> kVdShift; /* 1.0: the last bucket starts exactly there */
constexpr uint32_t kVdBuckets = kVdHiBucket - kVdLoBucket + 1u;
constexpr uint32_t kVdEntries = (kVdBuckets + 3u) & ~3u; /* staged 16 bytes at a time */

/* the tail of transformPQ(decode) once Vp is known (src/luma_quantizer.cpp:498-499) */
__device__ __forceinline__ float pq_decode_from_vp(float Vp, float l_max)
{
    const float n = 0.1593f, c1 = 0.8359f, c2 = 18.8516f, c3 = 18.6875f;
    const float inv_n = 1.0f / n;
    const float num = fmaxf(0.0f, __fsub_rn(Vp, c1));
    const float den = __fsub_rn(c2, __fmul_rn(c3, Vp));
    return __fmul_rn(l_max, powf_glibc(__fdiv_rn(num, den), inv_n));
}

#ifdef LUMA_PQ_TABLE_BUILDERS /* one translation unit only (luma_kern_tu.cu, generic CS 2) */
/* one warp per bucket of 32 consecutive floats */
__global__ void __launch_bounds__(256) build_pqd_kernel(uint4 *tab, float l_max)
{
    powf_tables_stage();
    const float m = 78.8438f;
    const float inv_m = 1.0f / m;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < kPqdBuckets; b += warps) {
        const float v = __uint_as_float(kPqdKey0 + b * 32u + lane);
        const float Vp = powf_glibc(v, inv_m);
        const uint32_t vb = __float_as_uint(Vp);
        const uint32_t first = __shfl_sync(0xffffffffu, vb, 0), last = __shfl_sync(0xffffffffu, vb, 31);
        const uint32_t m_lo = __ballot_sync(0xffffffffu, vb == first), m_hi = __ballot_sync(0xffffffffu, vb == last);
        const uint32_t n_lo = __popc(m_lo);
        /* regular: all equal, or a run of `first` followed by a run of `last` and nothing else */
        const bool regular = (first == last) ? (m_lo == 0xffffffffu)
                                             : ((m_lo | m_hi) == 0xffffffffu && m_lo == (0xffffffffu >> (32u - n_lo)));
        float val = 0.0f;
        if (lane == 0 || lane == 31)
            val = pq_decode_from_vp(Vp, l_max);
        const uint32_t v_lo = __shfl_sync(0xffffffffu, __float_as_uint(val), 0);
        const uint32_t v_hi = __shfl_sync(0xffffffffu, __float_as_uint(val), 31);
        if (lane == 0)
            tab[b] = make_uint4(v_lo, v_hi, regular ? n_lo : 0xffffffffu, 0u);
    }
}

__global__ void __launch_bounds__(256) build_pqe_kernel(float *tab)
{
    powf_tables_stage();
    const float m = 78.8438f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < kPqeKeys; i += gridDim.x * blockDim.x)
        tab[i] = powf_glibc(__uint_as_float(kPqeKey0 + i), m);
}
#endif /* LUMA_PQ_TABLE_BUILDERS */

/* ---- lookups (tuned kernels) ---------------------------------------------------------------------------- */
/* transformPQ(v, decode); identical to pq_decode(v, l_max) for every v */
__device__ __forceinline__ float pq_decode_tab(const QuantDev &q, float v, float l_max)
{
    const uint32_t idx = __float_as_uint(v) - kPqdKey0;
    if (q.pqd && idx < kPqdKeys) {
        const uint4 e = __ldg(q.pqd + (idx >> 5));
        if (e.z <= 32u)
            return __uint_as_float((idx & 31u) < e.z ? e.x : e.y);
    }
    if (v == 0.0f) /* black and super-black (the [0,1] clamp): powf(0, .) = 0 all the way through */
        return __fmul_rn(l_max, 0.0f);
    return pq_decode(v, l_max); /* below 2^-8, NaN, or an irregular bucket: the exact evaluation */
}

/* val / d with a precomputed rc = RN(1/d): q = RN(val rc), r = val - q d (exact in an FMA), RN(q + r rc).  Three
 * instructions instead of the ~10 of an IEEE division; correctly rounded for "most" divisors and operands -- which
 * ones is established by exhaustion, never assumed (below, and scripts/divchk.c for the compile-time divisor 255). */
__device__ __forceinline__ float div_by_rc(float val, float d, float rc)
{
    const float qq = __fmul_rn(val, rc);
    const float r = __fmaf_rn(-qq, d, val);
    return __fmaf_rn(r, rc, qq);
}
#ifdef LUMA_PQ_TABLE_BUILDERS
/* Every float val in [1e-10, FLT_MAX] (what max(c * sc, 1e-10f) can hand to PQ encode; +inf and NaN end as NaN either
 * way, src/luma_quantizer.cpp:493-494): does div_by_rc give the bits of val / l_max?  1.35e9 inputs, under a
 * millisecond. */
__global__ void __launch_bounds__(256) check_lmax_division_kernel(float l_max, float rc, uint32_t *bad)
{
    const uint32_t lo = __float_as_uint(1e-10f), hi = 0x7f7fffffu;
    bool differs = false;
    for (uint64_t u = (uint64_t)lo + blockIdx.x * blockDim.x + threadIdx.x; u <= hi; u += (uint64_t)gridDim.x * blockDim.x) {
        const float val = __uint_as_float((uint32_t)u);
        differs |= __float_as_uint(div_by_rc(val, l_max, rc)) != __float_as_uint(__fdiv_rn(val, l_max));
    }
    if (differs)
        atomicOr(bad, 1u);
}
#endif

/* transformPQ(val, encode); identical to pq_encode(val, l_max) for every val */
__device__ __forceinline__ float pq_encode_tab(const QuantDev &q, float val, float l_max)
{
    const float m = 78.8438f, n = 0.1593f, c1 = 0.8359f, c2 = 18.8516f, c3 = 18.6875f;
    const float Lp = powf_glibc<true>(q.lmax_rc != 0.0f ? div_by_rc(val, l_max, q.lmax_rc) : __fdiv_rn(val, l_max), n);
    const float num = __fadd_rn(c1, __fmul_rn(c2, Lp));
    const float den = __fadd_rn(1.0f, __fmul_rn(c3, Lp));
    const float b = __fdiv_rn(num, den);
    const uint32_t idx = __float_as_uint(b) - kPqeKey0;
    if (q.pqe && idx < kPqeKeys)
        return __ldg(q.pqe + idx);
    return powf_glibc(b, m);
}

/* Index of c in pqh when c is a (normal or zero) half-float value, else 0xFFFFFFFF.  Integer test on the float bits: the
 * 13 low mantissa bits clear and the exponent in the half range -- two instructions for a sample that is not (float ->
 * half -> float conversions would cost a quarter-rate pipe slot each, for every sample).  Half subnormals, infinities
 * and NaN are evaluated. */
__device__ __forceinline__ uint32_t half_index(float c)
{
    const uint32_t b = __float_as_uint(c);
    if ((b & 0x1FFFu) != 0u)
        return 0xFFFFFFFFu;
    const uint32_t e = (b >> 23) & 0xFFu;
    if (e - 113u <= 29u) /* 2^-14 <= |c| <= 65504 */
        return ((b >> 16) & 0x8000u) | ((e - 112u) << 10) | ((b >> 13) & 0x3FFu);
    return (b << 1) == 0u ? (b >> 16) : 0xFFFFFFFFu; /* +-0 */
}
/* PQenc of one input sample c (NOT yet multiplied by preScaling), evaluated */
__device__ __forceinline__ float pq_encode_sample(const QuantDev &q, float c, float sc, bool prescale, float l_max)
{
    const float v = prescale ? __fmul_rn(c, sc) : c;
    return pq_encode_tab(q, max_nan(v, 1e-10f), l_max);
}
#ifdef LUMA_PQ_TABLE_BUILDERS
__global__ void __launch_bounds__(256) build_pqh_kernel(const QuantDev q, float *tab, float sc, int prescale, float l_max)
{
    powf_tables_stage();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < 65536u; i += gridDim.x * blockDim.x) {
        const float c = __half2float(__ushort_as_half((unsigned short)i));
        const float v = prescale ? __fmul_rn(c, sc) : c;
        tab[i] = pq_encode_tab(q, max_nan(v, 1e-10f), l_max);
    }
}
#endif

/* BT.2020 Y'CbCr forward / inverse of one pixel with the tables (same expressions as ycbcr_forward_px /
 * ycbcr_inverse_px in luma_device.cuh) */
static __device__ __noinline__ float3 ycbcr_forward_px_tab(const QuantDev &q, float R, float G, float B, float l_max)
{
    const float Rp = pq_encode_tab(q, max_nan(R, 1e-10f), l_max);
    const float Gp = pq_encode_tab(q, max_nan(G, 1e-10f), l_max);
    const float Bp = pq_encode_tab(q, max_nan(B, 1e-10f), l_max);
    const float y = dot3(0.2627f, 0.6780f, 0.0593f, Rp, Gp, Bp);
    float3 c;
    c.x = pq_decode_tab(q, __fdiv_rn(__fadd_rn(__fmul_rn(219.0f, y), 16.0f), 255.0f), l_max);
    c.y = __fdiv_rn(__fadd_rn(__fmul_rn(224.0f, __fdiv_rn(__fsub_rn(Bp, y), 1.8814f)), 128.0f), 255.0f);
    c.z = __fdiv_rn(__fadd_rn(__fmul_rn(224.0f, __fdiv_rn(__fsub_rn(Rp, y), 1.4746f)), 128.0f), 255.0f);
    return c;
}

static __device__ __noinline__ float3 ycbcr_inverse_px_tab(const QuantDev &q, float y, float ca, float cb, float l_max)
{
    float blue = __fadd_rn(y, ca);
    float red = __fadd_rn(y, cb);
    float green = __fdiv_rn(__fsub_rn(__fsub_rn(y, __fmul_rn(0.2627f, red)), __fmul_rn(0.0593f, blue)), 0.6780f);
    red = clamp01_std(red);
    green = clamp01_std(green);
    blue = clamp01_std(blue);
    float3 o;
    o.x = pq_decode_tab(q, red, l_max);
    o.y = pq_decode_tab(q, green, l_max);
    o.z = pq_decode_tab(q, blue, l_max);
    return o;
}

/* Pixel PAIRS: the two pixels' table lookups (3 + 1 each on the forward path, 3 each on the inverse path) are all in
 * flight together, which is what hides the L2 latency of the gathers; same expressions, evaluated per pixel. */
struct Float3x2 {
    float3 a, b;
};
/* Forward transform of a pixel pair.  LUMA_V: return v = (219 y' + 16)/255 itself in .x (the caller searches it in vdtab)
 * instead of PQdec(v).  HALF: all six input samples are known to be half-float values (the caller checked its whole tile):
 * their PQ encodes come out of pqh (preScaling folded in); otherwise they are evaluated (p0, p1 arrive WITHOUT preScaling
 * either way). */
template <bool LUMA_V, bool HALF>
static __device__ __noinline__ Float3x2 ycbcr_forward_px2_tab(const QuantDev &q, const float *pqh, float3 p0, float3 p1, float sc,
                                                                bool prescale, float l_max)
{
    float Rp0, Gp0, Bp0, Rp1, Gp1, Bp1;
    if (HALF) {
        Rp0 = __ldg(pqh + half_index(p0.x)), Gp0 = __ldg(pqh + half_index(p0.y)), Bp0 = __ldg(pqh + half_index(p0.z));
        Rp1 = __ldg(pqh + half_index(p1.x)), Gp1 = __ldg(pqh + half_index(p1.y)), Bp1 = __ldg(pqh + half_index(p1.z));
    } else {
        Rp0 = pq_encode_sample(q, p0.x, sc, prescale, l_max), Rp1 = pq_encode_sample(q, p1.x, sc, prescale, l_max);
        Gp0 = pq_encode_sample(q, p0.y, sc, prescale, l_max), Gp1 = pq_encode_sample(q, p1.y, sc, prescale, l_max);
        Bp0 = pq_encode_sample(q, p0.z, sc, prescale, l_max), Bp1 = pq_encode_sample(q, p1.z, sc, prescale, l_max);
    }
    const float y0 = dot3(0.2627f, 0.6780f, 0.0593f, Rp0, Gp0, Bp0), y1 = dot3(0.2627f, 0.6780f, 0.0593f, Rp1, Gp1, Bp1);
    /* x / 255 by div_const_int<255>: exact whenever the quotient is a normal number or zero (scripts/divchk.c, all 2^32
     * operands); the numerators here are 219 y' + 16 and 224 t + 128 with y', t = O(1): sums of floats of magnitude
     * >= 16 resp. 128 ulps-of-128 apart, i.e. zero or >= 2^-17 in magnitude, never subnormal; NaN stays NaN.  (1.8814
     * and 1.4746 do NOT qualify -- millions of mismatches -- and keep the IEEE division.) */
    const float v0 = div_const_int<255>(__fadd_rn(__fmul_rn(219.0f, y0), 16.0f));
    const float v1 = div_const_int<255>(__fadd_rn(__fmul_rn(219.0f, y1), 16.0f));
    Float3x2 c;
    c.a.x = LUMA_V ? v0 : pq_decode_tab(q, v0, l_max);
    c.b.x = LUMA_V ? v1 : pq_decode_tab(q, v1, l_max);
    c.a.y = div_const_int<255>(__fadd_rn(__fmul_rn(224.0f, __fdiv_rn(__fsub_rn(Bp0, y0), 1.8814f)), 128.0f));
    c.a.z = div_const_int<255>(__fadd_rn(__fmul_rn(224.0f, __fdiv_rn(__fsub_rn(Rp0, y0), 1.4746f)), 128.0f));
    c.b.y = div_const_int<255>(__fadd_rn(__fmul_rn(224.0f, __fdiv_rn(__fsub_rn(Bp1, y1), 1.8814f)), 128.0f));
    c.b.z = div_const_int<255>(__fadd_rn(__fmul_rn(224.0f, __fdiv_rn(__fsub_rn(Rp1, y1), 1.4746f)), 128.0f));
    return c;
}

static __device__ __noinline__ Float3x2 ycbcr_inverse_px2_tab(const QuantDev &q, float2 y, float2 ca, float2 cb, float l_max)
{
    float v[6];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float yy = i ? y.y : y.x;
        const float blue = __fadd_rn(yy, i ? ca.y : ca.x);
        const float red = __fadd_rn(yy, i ? cb.y : cb.x);
        const float green = __fdiv_rn(__fsub_rn(__fsub_rn(yy, __fmul_rn(0.2627f, red)), __fmul_rn(0.0593f, blue)), 0.6780f);
        v[3 * i + 0] = clamp01_std(red);
        v[3 * i + 1] = clamp01_std(green);
        v[3 * i + 2] = clamp01_std(blue);
    }
    /* A lookup costs one 32-byte L2 sector; on content without locality (noise) three of them per pixel saturate the
     * L2 (~235 G sectors/s on B200: 79 kMpx/s) while the SMs idle.  So the green of every other pixel is EVALUATED (two
     * exact powf, ~220 instructions) instead of looked up: both resources busy.  Measured on 4K noise, decode:
     * 1.19 TB/s-equivalent all looked up, 1.26 with every green evaluated, 1.29 with every other one.  Re-measured with
     * the fused powf (scripts/ycbcr_green_probe.py, decoder tuning 2000 = every green): 771 vs 767 us per 8 frames on
     * noise, 589 vs 637 us on the reference's test pattern -- every other one stays. */
    float o[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        if ((i == 1 || (i == 4 && (q.tune_flags & 1u))) && q.pqd) /* green of the first pixel of the pair (tuning: of both) */
            o[i] = (v[i] == 0.0f) ? __fmul_rn(l_max, 0.0f) : pq_decode(v[i], l_max);
        else
            o[i] = pq_decode_tab(q, v[i], l_max);
    }
    Float3x2 r;
    r.a = make_float3(o[0], o[1], o[2]);
    r.b = make_float3(o[3], o[4], o[5]);
    return r;
}

#ifdef LUMA_PQ_TABLE_BUILDERS
/* vdtab: one thread per bucket of 2^13 floats */
__device__ __forceinline__ uint32_t vd_code(const QuantDev &q, const SearchCtx &s, uint32_t key, float l_max)
{
    return search_code<false>(s, pq_decode_tab(q, __uint_as_float(key), l_max));
}
__global__ void __launch_bounds__(256) build_vdtab_kernel(const QuantDev q, uint32_t *tab, uint32_t *bad, float l_max)
{
    powf_tables_stage();
    SearchCtx s;
    s.thr = q.thr, s.bucket = q.bucket, s.lut = q.lut;
    s.max_val = q.max_val, s.shift = q.shift, s.base = q.base, s.nbm1 = q.nbm1, s.walk = q.walk, s.mode = q.search_mode;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < kVdEntries; b += gridDim.x * blockDim.x) {
        if (b >= kVdBuckets) {
            tab[b] = 0u;
            continue;
        }
        const uint32_t kb = kVdLoBucket + b, k0 = kb << kVdShift, k1 = k0 + (1u << kVdShift) - 1u;
        const uint32_t c_first = vd_code(q, s, k0, l_max), c_last = vd_code(q, s, k1, l_max);
        uint32_t thr_low = 1u << kVdShift;
        bool ok = c_last == c_first || c_last == c_first + 1u;
        if (b + 1u < kVdBuckets && vd_code(q, s, k1 + 1u, l_max) < c_last)
            ok = false; /* codes must not decrease across the bucket boundary */
        if (ok && c_last != c_first) {
            uint32_t lo = k0, hi = k1; /* code(lo) < c_last <= code(hi) */
            while (hi - lo > 1u) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                if (vd_code(q, s, mid, l_max) >= c_last)
                    hi = mid;
                else
                    lo = mid;
            }
            /* glibc's powf may round non-monotonically within a few floats of a step of Vp: the step must be clean */
            for (uint32_t d = 1; d <= 96u && ok; ++d)
                ok = vd_code(q, s, hi - d, l_max) < c_last && vd_code(q, s, hi + d - 1u, l_max) >= c_last;
            thr_low = hi - k0;
        }
        if (!ok)
            atomicOr(bad, 1u);
        tab[b] = (c_first << 16) + 0x10000u - thr_low - (kb << kVdShift);
    }
}
#endif /* LUMA_PQ_TABLE_BUILDERS */

} // namespace lumacu

/*
 * luma_device.cuh -- per-pixel device math of the B200 HDR<->integer transform.
 *
 * Everything here must produce the same bits as the reference CPU path
 * (/root/reference/src/luma_quantizer.cpp, compiled without FMA contraction):
 *   - every reference multiply/add/divide is an explicit round-to-nearest
 *     intrinsic (__fmul_rn/__fadd_rn/__fdiv_rn), evaluated in the reference's
 *     association order; the translation unit is also built with -fmad=false;
 *   - std::min/std::max on the XYZ clamps propagate NaN (min.NaN/max.NaN),
 *     the chroma/RGB' clamps do not (fminf/fmaxf), exactly like the
 *     compare-select semantics of the reference expressions;
 *   - FMAs appear only inside exactly-rounded division sequences.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "luma_kernels_decl.cuh"
#include "powf_glibc.cuh"

namespace lumacu {

/* ---- NaN-aware compare-selects ------------------------------------------------ */
__device__ __forceinline__ float min_nan(float a, float b)
{
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float max_nan(float a, float b)
{
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
/* std::max(0.0f, std::min(1.0f, v)) of the YCbCr inverse (src/luma_quantizer.cpp:453-455).  The C++
 * library functions are compare-selects -- (v < 1) ? v : 1, then (0 < t) ? t : 0 -- so NaN -> 1.  Both
 * fmaxf(0, fminf(1, v)) and the literal compare-selects are recognised by the compiler as a saturate
 * (FADD.SAT), which sends NaN to 0; hence the saturate is written out and NaN is patched explicitly. */
__device__ __forceinline__ float clamp01_std(float v)
{
    const float t = __saturatef(v);
    return (v != v) ? 1.0f : t;
}

/* std::max(std::min(v, 1e8f), 1e-4f) (src/luma_quantizer.cpp:284-286,303-305,412-414):
 * NaN in -> NaN out. */
__device__ __forceinline__ float clamp_xyz(float v) { return max_nan(min_nan(v, 100000000.0f), 0.0001f); }

/* ((m0*a)+(m1*b))+(m2*c), no contraction */
__device__ __forceinline__ float dot3(float m0, float m1, float m2, float a, float b, float c)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(m0, a), __fmul_rn(m1, b)), __fmul_rn(m2, c));
}

/* x / d for a compile-time constant d, correctly rounded, 3 instructions.
 * q = RN(x*rc), r = x - q*d (exact in an FMA), q' = RN(q + r*rc).  scripts/divchk.c checks the
 * sequence against IEEE division for all 2^32 inputs: for d = 255 (the only divisor used) it is
 * exact whenever the quotient is a normal number (3 mismatches in total, all with subnormal
 * quotients).  Callers only use it on O(1) chromaticities times 410 / 1640; everything else
 * goes through __fdiv_rn. */
template <int D>
__device__ __forceinline__ float div_const_int(float x)
{
    const float d = (float)D;
    const float rc = 1.0f / d;
    float q = __fmul_rn(x, rc);
    float r = __fmaf_rn(-q, d, x);
    return __fmaf_rn(r, rc, q);
}

/* ---- colour matrices (include/luma/luma_quantizer.h:79-87) --------------------- */
#define LUMA_M00 0.412424f
#define LUMA_M01 0.357579f
#define LUMA_M02 0.180464f
#define LUMA_M10 0.212656f
#define LUMA_M11 0.715158f
#define LUMA_M12 0.072186f
#define LUMA_M20 0.019332f
#define LUMA_M21 0.119193f
#define LUMA_M22 0.950444f

#define LUMA_I00 3.240708f
#define LUMA_I01 -1.537259f
#define LUMA_I02 -0.498570f
#define LUMA_I10 -0.969257f
#define LUMA_I11 1.875995f
#define LUMA_I12 0.041555f
#define LUMA_I20 0.055636f
#define LUMA_I21 -0.203996f
#define LUMA_I22 1.057069f

/* ---- PQ (src/luma_quantizer.cpp:485-501), used per pixel only by CS_YCBCR ------ */
__device__ __forceinline__ float pq_encode(float val, float l_max)
{
    const float m = 78.8438f, n = 0.1593f, c1 = 0.8359f, c2 = 18.8516f, c3 = 18.6875f;
    float Lp = powf_glibc<true>(__fdiv_rn(val, l_max), n);
    float num = __fadd_rn(c1, __fmul_rn(c2, Lp));
    float den = __fadd_rn(1.0f, __fmul_rn(c3, Lp));
    return powf_glibc(__fdiv_rn(num, den), m);
}
__device__ __forceinline__ float pq_decode(float val, float l_max)
{
    const float m = 78.8438f, n = 0.1593f, c1 = 0.8359f, c2 = 18.8516f, c3 = 18.6875f;
    const float inv_m = 1.0f / m; /* evaluated in fp32 like the reference's 1.0f/m */
    const float inv_n = 1.0f / n;
    float Vp = powf_glibc<true>(val, inv_m);
    float num = fmaxf(0.0f, __fsub_rn(Vp, c1)); /* std::max(0.0f, x): NaN -> 0 */
    float den = __fsub_rn(c2, __fmul_rn(c3, Vp));
    return __fmul_rn(l_max, powf_glibc(__fdiv_rn(num, den), inv_n));
}

/* BT.2020 Y'CbCr of one pixel (src/luma_quantizer.cpp:317-354): eight exact powf evaluations.  Deliberately NOT
 * inlined: a tile holds 8 pixels, and 64 inlined copies of the powf sequence (~100 instructions each) made the
 * kernels instruction-fetch bound (ncu: "no_instruction" was the top stall).  One copy, called per pixel, with
 * the three PQ curves of a pixel interleaved inside it. */
static __device__ __noinline__ float3 ycbcr_forward_px(float R, float G, float B, float l_max)
{
    /* std::max(v, 1e-10f) with v first: NaN stays NaN */
    const float Rp = pq_encode(max_nan(R, 1e-10f), l_max);
    const float Gp = pq_encode(max_nan(G, 1e-10f), l_max);
    const float Bp = pq_encode(max_nan(B, 1e-10f), l_max);
    const float y = dot3(0.2627f, 0.6780f, 0.0593f, Rp, Gp, Bp);
    float3 c;
    c.x = pq_decode(__fdiv_rn(__fadd_rn(__fmul_rn(219.0f, y), 16.0f), 255.0f), l_max);
    c.y = __fdiv_rn(__fadd_rn(__fmul_rn(224.0f, __fdiv_rn(__fsub_rn(Bp, y), 1.8814f)), 128.0f), 255.0f);
    c.z = __fdiv_rn(__fadd_rn(__fmul_rn(224.0f, __fdiv_rn(__fsub_rn(Rp, y), 1.4746f)), 128.0f), 255.0f);
    return c;
}

/* ---- forward colour transform of one pixel (src/luma_quantizer.cpp:269-373) ---- */
template <int CS>
__device__ __forceinline__ void color_forward(float R, float G, float B, float l_max, float &c0, float &c1,
                                              float &c2)
{
    if (CS == CS_XYZ) {
        c0 = clamp_xyz(dot3(LUMA_M00, LUMA_M01, LUMA_M02, R, G, B));
        c1 = clamp_xyz(dot3(LUMA_M10, LUMA_M11, LUMA_M12, R, G, B));
        c2 = clamp_xyz(dot3(LUMA_M20, LUMA_M21, LUMA_M22, R, G, B));
    } else if (CS == CS_LUV) {
        float X = clamp_xyz(dot3(LUMA_M00, LUMA_M01, LUMA_M02, R, G, B));
        float Y = clamp_xyz(dot3(LUMA_M10, LUMA_M11, LUMA_M12, R, G, B));
        float Z = clamp_xyz(dot3(LUMA_M20, LUMA_M21, LUMA_M22, R, G, B));
        float sum = __fadd_rn(__fadd_rn(X, Y), Z);
        float x = __fdiv_rn(X, sum);
        float y = __fdiv_rn(Y, sum);
        /* ((-2x) + (12y)) + 3 ; -2x is exact so folding it into an FMA keeps the bits */
        float den = __fadd_rn(__fmaf_rn(-2.0f, x, __fmul_rn(12.0f, y)), 3.0f);
        c0 = Y;
        /* (((4x)/den)*410)/255 ; 4x is exact */
        c1 = div_const_int<255>(__fmul_rn(__fdiv_rn(__fmul_rn(4.0f, x), den), 410.0f));
        c2 = div_const_int<255>(__fmul_rn(__fdiv_rn(__fmul_rn(9.0f, y), den), 410.0f));
    } else if (CS == CS_YCBCR) {
        const float3 c = ycbcr_forward_px(R, G, B, l_max);
        c0 = c.x;
        c1 = c.y;
        c2 = c.z;
    } else { /* CS_RGB */
        c0 = R;
        c1 = G;
        c2 = B;
    }
}

/* ---- inverse colour transform (src/luma_quantizer.cpp:374-479) ------------------
 * Split in a chroma part (identical for the 4 pixels of a 4:2:0 block) and a
 * per-pixel part.  For LUV the chroma part yields x/y and (1-x-y)/y. */
struct ChromaInv {
    float a, b;
};

template <int CS>
__device__ __forceinline__ ChromaInv chroma_inverse(float c1, float c2)
{
    ChromaInv r;
    if (CS == CS_LUV) {
        float u = __fdiv_rn(__fmul_rn(c1, 255.0f), 410.0f);
        float v = __fdiv_rn(__fmul_rn(c2, 255.0f), 410.0f);
        /* ((6u) - (16v)) + 12 ; 16v is exact */
        float den = __fadd_rn(__fmaf_rn(-16.0f, v, __fmul_rn(6.0f, u)), 12.0f);
        float x = __fdiv_rn(__fmul_rn(9.0f, u), den);
        float y = __fdiv_rn(__fmul_rn(4.0f, v), den);
        r.a = __fdiv_rn(x, y);
        r.b = __fdiv_rn(__fsub_rn(__fsub_rn(1.0f, x), y), y);
    } else if (CS == CS_YCBCR) {
        /* 1.8814*(255*c1 - 128)/224 and 1.4746*(255*c2 - 128)/224 */
        r.a = __fdiv_rn(__fmul_rn(1.8814f, __fsub_rn(__fmul_rn(255.0f, c1), 128.0f)), 224.0f);
        r.b = __fdiv_rn(__fmul_rn(1.4746f, __fsub_rn(__fmul_rn(255.0f, c2), 128.0f)), 224.0f);
    } else {
        r.a = c1;
        r.b = c2;
    }
    return r;
}

/* BT.2020 Y'CbCr -> linear RGB once y = ((255 PQenc(L)) - 16) / 219 is known (src/luma_quantizer.cpp:449-459).  The
 * tuned decode kernel reads y from a host-built per-code table instead of evaluating PQenc per pixel. */
static __device__ __noinline__ float ycbcr_luma_term(float c0, float l_max) /* ((255 PQenc(L)) - 16) / 219 */
{
    const float y = pq_encode(c0, l_max);
    return __fdiv_rn(__fsub_rn(__fmul_rn(255.0f, y), 16.0f), 219.0f);
}
static __device__ __noinline__ float3 ycbcr_inverse_px(float y, float ca, float cb, float l_max) /* one copy, see ycbcr_forward_px */
{
    float blue = __fadd_rn(y, ca);
    float red = __fadd_rn(y, cb);
    float green = __fdiv_rn(__fsub_rn(__fsub_rn(y, __fmul_rn(0.2627f, red)), __fmul_rn(0.0593f, blue)), 0.6780f);
    /* std::max(0.0f, std::min(1.0f, v)): NaN -> 1 */
    red = clamp01_std(red);
    green = clamp01_std(green);
    blue = clamp01_std(blue);
    float3 o;
    o.x = pq_decode(red, l_max);
    o.y = pq_decode(green, l_max);
    o.z = pq_decode(blue, l_max);
    return o;
}
__device__ __forceinline__ void ycbcr_inverse_from_y(float y, ChromaInv ch, float l_max, float &R, float &G, float &B)
{
    const float3 o = ycbcr_inverse_px(y, ch.a, ch.b, l_max);
    R = o.x;
    G = o.y;
    B = o.z;
}

template <int CS>
__device__ __forceinline__ void color_inverse(float c0, ChromaInv ch, float l_max, float &R, float &G, float &B)
{
    if (CS == CS_LUV) {
        float Y = clamp_xyz(c0);
        float X = clamp_xyz(__fmul_rn(ch.a, c0));
        float Z = clamp_xyz(__fmul_rn(ch.b, c0));
        R = dot3(LUMA_I00, LUMA_I01, LUMA_I02, X, Y, Z);
        G = dot3(LUMA_I10, LUMA_I11, LUMA_I12, X, Y, Z);
        B = dot3(LUMA_I20, LUMA_I21, LUMA_I22, X, Y, Z);
    } else if (CS == CS_XYZ) {
        R = dot3(LUMA_I00, LUMA_I01, LUMA_I02, c0, ch.a, ch.b);
        G = dot3(LUMA_I10, LUMA_I11, LUMA_I12, c0, ch.a, ch.b);
        B = dot3(LUMA_I20, LUMA_I21, LUMA_I22, c0, ch.a, ch.b);
    } else if (CS == CS_YCBCR) {
        ycbcr_inverse_from_y(ycbcr_luma_term(c0, l_max), ch, l_max, R, G, B);
    } else {
        R = c0;
        G = ch.a;
        B = ch.b;
    }
}

/* ---- luma search ------------------------------------------------------------------
 * The reference (src/luma_quantizer.cpp:219-235) bisects the LUT for the
 * bracketing codes (l, l+1) and picks the nearer one using fp32 differences.
 * For a strictly increasing LUT that function is monotone in `val`, so it is
 * fully described by max_val thresholds T_k = smallest float whose code is >= k
 * (derived on the host with the reference's own expression).  The code is then
 * the number of thresholds <= val, compared as order-preserving integer keys. */
template <bool POSITIVE>
__device__ __forceinline__ uint32_t ordered_key(float v)
{
    (void)POSITIVE; /* kept for the call sites; the general transform is always used (see encode_kernel) */
    const uint32_t b = __float_as_uint(v);
    const uint32_t k = b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
    return (v != v) ? 0xFFFFFFFFu : k; /* the reference maps every NaN to max_val */
}

struct SearchCtx {
    const uint32_t *thr;
    const uint16_t *bucket;
    const float *lut;
    uint32_t max_val, shift, base, nbm1, walk;
    int mode;
};

template <bool POSITIVE>
__device__ __forceinline__ uint32_t search_code(const SearchCtx &s, float val)
{
    if (s.mode == SEARCH_BUCKET) {
        const uint32_t key = ordered_key<POSITIVE>(val);
        uint32_t b = min(max(key >> s.shift, s.base), s.base + s.nbm1) - s.base;
        const uint32_t c0 = s.bucket[b];
        uint32_t c = c0 + (s.thr[c0] <= key) + (s.thr[c0 + 1] <= key);
        for (uint32_t j = 2; j < s.walk; ++j)
            c += (s.thr[c0 + j] <= key);
        return min(c, s.max_val);
    } else if (s.mode == SEARCH_BINARY) {
        const uint32_t key = ordered_key<POSITIVE>(val);
        uint32_t lo = 0, n = s.max_val;
        while (n > 0) {
            uint32_t half = n >> 1;
            if (s.thr[lo + half] <= key) {
                lo += half + 1;
                n -= half + 1;
            } else {
                n = half;
            }
        }
        return lo;
    } else {
        /* literal replica for LUTs that are not strictly increasing */
        int l = 0, r = (int)s.max_val;
        while (l + 1 < r) {
            int m = (l + r) / 2;
            if (val < s.lut[m])
                r = m;
            else
                l = m;
        }
        return (__fsub_rn(val, s.lut[l]) < __fsub_rn(s.lut[r], val)) ? (uint32_t)l : (uint32_t)r;
    }
}

/* chroma branch of quantize (src/luma_quantizer.cpp:239-240) */
__device__ __forceinline__ uint32_t quantize_chroma(float val, float max_c)
{
    float res = floorf(__fadd_rn(__fmul_rn(max_c, val), 0.5f));
    res = fmaxf(0.0f, fminf(max_c, res)); /* NaN -> max_c */
    return (uint32_t)res;
}

/* chroma branch of dequantize (src/luma_quantizer.cpp:261) */
__device__ __forceinline__ float dequantize_chroma(float code, float max_c)
{
    /* std::max(val/maxC, 1e-10f): NaN first operand stays NaN */
    return max_nan(__fdiv_rn(code, max_c), 1e-10f);
}

} // namespace lumacu

/*
 * luma_kern_tu.cu -- one kernel translation unit; built eight times:
 *   -DLUMA_TU_CS=<0..3>   colour space (CS_LUV, CS_RGB, CS_YCBCR, CS_XYZ)
 *   -DLUMA_TU_FAST=<0|1>  0: generic kernels (luma_kernels.cuh), 1: tuned kernels (luma_fast.cuh)
 * so that `make -j` compiles the instantiations in parallel.  Each unit only exports host-side
 * getters (luma_dispatch.h); lumacu.cu launches what they return.
 */
#ifndef LUMA_TU_CS
#error "build with -DLUMA_TU_CS=<0..3> -DLUMA_TU_FAST=<0|1>"
#endif

#include "luma_dispatch.h"

#if !LUMA_TU_FAST && LUMA_TU_CS == 1
#define LUMA_TU_ELEMENTWISE 1 /* the element-wise API kernels live in this unit only */
#endif

#if LUMA_TU_FAST
#include "luma_fast.cuh"
#else
#include "luma_kernels.cuh"
#if LUMA_TU_CS == 2
#define LUMA_PQ_TABLE_BUILDERS 1
#include "luma_pq_tables.cuh"
#endif
#endif

#define LUMA_CAT2(a, b) a##b
#define LUMA_CAT(a, b) LUMA_CAT2(a, b)

namespace lumacu {

constexpr int kCS = LUMA_TU_CS;

#if !LUMA_TU_FAST

enc_fn LUMA_CAT(get_encode_generic_cs, LUMA_TU_CS)(bool sub, int bytes, bool vec)
{
    if (sub) {
        if (bytes == 2)
            return vec ? encode_kernel<kCS, true, 2, true> : encode_kernel<kCS, true, 2, false>;
        return vec ? encode_kernel<kCS, true, 1, true> : encode_kernel<kCS, true, 1, false>;
    }
    if (bytes == 2)
        return vec ? encode_kernel<kCS, false, 2, true> : encode_kernel<kCS, false, 2, false>;
    return vec ? encode_kernel<kCS, false, 1, true> : encode_kernel<kCS, false, 1, false>;
}

dec_fn LUMA_CAT(get_decode_generic_cs, LUMA_TU_CS)(bool sub, int bytes, bool vec)
{
    if (sub) {
        if (bytes == 2)
            return vec ? decode_kernel<kCS, true, 2, true> : decode_kernel<kCS, true, 2, false>;
        return vec ? decode_kernel<kCS, true, 1, true> : decode_kernel<kCS, true, 1, false>;
    }
    if (bytes == 2)
        return vec ? decode_kernel<kCS, false, 2, true> : decode_kernel<kCS, false, 2, false>;
    return vec ? decode_kernel<kCS, false, 1, true> : decode_kernel<kCS, false, 1, false>;
}

#if LUMA_TU_CS == 2
/* builders of the CS_YCBCR PQ tables (luma_pq_tables.cuh) live in this unit */
void launch_build_pq_tables(unsigned blocks, cudaStream_t st, void *pqd, float *pqe, float l_max)
{
    build_pqd_kernel<<<blocks, 256, 0, st>>>((uint4 *)pqd, l_max);
    build_pqe_kernel<<<blocks, 256, 0, st>>>(pqe);
}
void launch_check_lmax_division(unsigned blocks, cudaStream_t st, float l_max, float rc, uint32_t *bad)
{
    check_lmax_division_kernel<<<blocks, 256, 0, st>>>(l_max, rc, bad);
}
void launch_build_pqh(cudaStream_t st, const QuantDev &q, float *tab, float sc, int prescale, float l_max)
{
    build_pqh_kernel<<<64, 256, 0, st>>>(q, tab, sc, prescale, l_max);
}
void launch_build_vdtab(cudaStream_t st, const QuantDev &q, uint32_t *tab, uint32_t *bad, float l_max)
{
    build_vdtab_kernel<<<(kVdEntries + 255u) / 256u, 256, 0, st>>>(q, tab, bad, l_max);
}
#endif

#if LUMA_TU_CS == 1
/* the element-wise API kernels live in one unit only */
void launch_transform(int cs, bool fwd, unsigned blocks, cudaStream_t st, float *c0, float *c1, float *c2, size_t n,
                      float sc, float l_max)
{
#define LAUNCH_T(CSV)                                                                   \
    if (fwd)                                                                            \
        transform_kernel<CSV, true><<<blocks, kThreads, 0, st>>>(c0, c1, c2, n, sc, l_max); \
    else                                                                                \
        transform_kernel<CSV, false><<<blocks, kThreads, 0, st>>>(c0, c1, c2, n, sc, l_max);
    switch (cs) {
    case CS_LUV: LAUNCH_T(CS_LUV) break;
    case CS_RGB: LAUNCH_T(CS_RGB) break;
    case CS_YCBCR: LAUNCH_T(CS_YCBCR) break;
    default: LAUNCH_T(CS_XYZ) break;
    }
#undef LAUNCH_T
}

void launch_quantize(unsigned blocks, size_t smem, cudaStream_t st, const QuantDev &q, const float *in, float *out,
                     size_t n, int use_lut)
{
    quantize_kernel<<<blocks, kThreads, smem, st>>>(q, in, out, n, use_lut);
}

void launch_dequantize(unsigned blocks, cudaStream_t st, const QuantDev &q, const float *in, float *out, size_t n,
                       int use_lut)
{
    dequantize_kernel<<<blocks, kThreads, 0, st>>>(q, in, out, n, use_lut);
}

const void *quantize_kernel_ptr() { return (const void *)quantize_kernel; }

void launch_test_frame(unsigned blocks, cudaStream_t st, float *rgb, uint32_t w, uint32_t h)
{
    test_frame_kernel<<<blocks, kThreads, 0, st>>>(rgb, w, h);
}

void launch_half_rgba_to_frame(unsigned blocks, cudaStream_t st, const void *rgba, float *rgb, size_t n, size_t plane_stride, int mode)
{
    half_rgba_to_frame_kernel<<<blocks, kThreads, 0, st>>>((const uint2 *)rgba, rgb, n, plane_stride, mode);
}
void launch_frame_to_half_rgba(unsigned blocks, cudaStream_t st, const float *rgb, void *rgba, size_t n, size_t plane_stride)
{
    frame_to_half_rgba_kernel<<<blocks, kThreads, 0, st>>>(rgb, (uint2 *)rgba, n, plane_stride);
}

void launch_pfs_channels(bool to_rgb, unsigned blocks, cudaStream_t st, const float *a0, const float *a1, const float *a2,
                         float *o0, float *o1, float *o2, size_t n)
{
    if (to_rgb)
        pfs_channels_kernel<true><<<blocks, kThreads, 0, st>>>(a0, a1, a2, o0, o1, o2, n);
    else
        pfs_channels_kernel<false><<<blocks, kThreads, 0, st>>>(a0, a1, a2, o0, o1, o2, n);
}

void launch_display_linear(unsigned blocks, unsigned n_frames, cudaStream_t st, const DecArgs &a, int cs, int sub, int bytes)
{
    display_linear_kernel<<<dim3(blocks, n_frames), kThreads, 0, st>>>(a, cs, sub, bytes);
}
#endif

#else /* LUMA_TU_FAST */

/* Instantiations of the tuned kernels (see luma_fast.cuh).  variant = 10 * PF + MINB:
 *   PF 0 = loads straight into registers, 1 = next tile prefetched into registers, 8 = next tile staged in
 *   shared memory by a tensor-map (TMA) copy, 3..5 = timing diagnostics that skip part of the arithmetic;
 *   MINB = resident blocks per SM the register allocation is held to.
 * Every configuration exists as variant 4 (PF 0, the default); the headline instantiation (Lu'v', 4:2:0,
 * 16-bit planes, walk 1) is additionally compiled in more variants so that scripts/sweep.py can re-measure
 * the choice. */
template <bool SUB, int BYTES, int PF, bool PRESC, int FASTC>
static enc_fn pick_walk(int walk)
{
#if LUMA_TU_CS == 0 || LUMA_TU_CS == 3
    if (walk == 0) /* direct search table over [1e-4, 1e8] (positive searched values: Lu'v' Y, XYZ) */
        return encode_fast_kernel<kCS, SUB, BYTES, 0, PF, 4, PRESC, FASTC>;
#else
    if (walk == 0)
        return nullptr;
#endif
    if (walk == -1) /* direct search table over the thresholds' own range (both clamps on the device) */
        return encode_fast_kernel<kCS, SUB, BYTES, -1, PF, 4, PRESC, FASTC>;
    if (walk == -4) /* direct table in global memory (very wide LUTs: 13-16 bits) */
        return encode_fast_kernel<kCS, SUB, BYTES, -4, PF, 4, PRESC, FASTC>;
    if (walk == -3) /* 64-bit direct table, two thresholds per bucket (wide LUTs such as PQ-12) */
        return encode_fast_kernel<kCS, SUB, BYTES, -3, PF, 4, PRESC, FASTC>;
#if LUMA_TU_CS == 2
    if (walk == -2) /* CS_YCBCR: plane 0 searched by v = (219 y' + 16)/255 in the v-keyed table (no statistics) */
        return encode_fast_kernel<kCS, SUB, BYTES, -2, PF, 4, PRESC, FASTC>;
#else
    if (walk == -2)
        return nullptr;
#endif
#if LUMA_TU_CS == 0
    /* the headline colour space gets the exact walk length */
    if (walk <= 1)
        return encode_fast_kernel<kCS, SUB, BYTES, 1, PF, 4, PRESC, FASTC>;
    if (walk == 2)
        return encode_fast_kernel<kCS, SUB, BYTES, 2, PF, 4, PRESC, FASTC>;
#endif
    if (walk <= 4)
        return encode_fast_kernel<kCS, SUB, BYTES, 4, PF, 4, PRESC, FASTC>;
    return nullptr;
}

template <int PF, bool PRESC>
static enc_fn pick_enc_cfg(bool sub, int bytes, int walk, bool screened)
{
#if LUMA_TU_CS == 0
    if (sub && screened) /* Lu'v' 4:2:0 with screened chroma (luma_fast.cuh FASTC) */
        return bytes == 2 ? pick_walk<true, 2, PF, PRESC, 2>(walk) : pick_walk<true, 1, PF, PRESC, 2>(walk);
#endif
    (void)screened;
    if (sub)
        return bytes == 2 ? pick_walk<true, 2, PF, PRESC, 0>(walk) : pick_walk<true, 1, PF, PRESC, 0>(walk);
    return bytes == 2 ? pick_walk<false, 2, PF, PRESC, 0>(walk) : pick_walk<false, 1, PF, PRESC, 0>(walk);
}

enc_fn LUMA_CAT(get_encode_fast_cs, LUMA_TU_CS)(bool sub, int bytes, int walk, int variant, bool prescale)
{
    if (variant == 4)
        return prescale ? pick_enc_cfg<0, true>(sub, bytes, walk, false) : pick_enc_cfg<0, false>(sub, bytes, walk, false);
    if (variant == kEncVariantScreened) /* Lu'v' 4:2:0: screened chroma, queued redo, L2 prefetch two tiles ahead */
        return prescale ? pick_enc_cfg<6, true>(sub, bytes, walk, true) : pick_enc_cfg<6, false>(sub, bytes, walk, true);
    if (prescale)
        return nullptr; /* the tuning variants exist for preScaling == 1 only */
#if LUMA_TU_CS == 0
    if (sub && bytes == 2 && walk == 0) {
        switch (variant) {
        /* screened chroma (the default is 67): 6 = warp-wide redo, 7 = queued redo, 26 / 27 = those + L2 prefetch of the
         * next tile, 67 = queued + prefetch two tiles ahead, 86 / 87 = tensor-map staging.  Sustained encode of 32 4K
         * frames on B200 (scripts/sweep.py --preroll 0.6, profiles/r02_sweep_screened.log): exact chain 740 us, 6: 714,
         * 7: 721, 26: 689, 27: 669, 86: 675, 87: 670, 67: 660 (decode: 654) */
        case 6: return encode_fast_kernel<kCS, true, 2, 0, 0, 4, false, 1>;
        case 7: return encode_fast_kernel<kCS, true, 2, 0, 0, 4, false, 2>;
        case 26: return encode_fast_kernel<kCS, true, 2, 0, 2, 4, false, 1>;
        case 27: return encode_fast_kernel<kCS, true, 2, 0, 2, 4, false, 2>;
        case 86: return encode_fast_kernel<kCS, true, 2, 0, 8, 4, false, 1>;
        case 87: return encode_fast_kernel<kCS, true, 2, 0, 8, 4, false, 2>;
        case 64: return encode_fast_kernel<kCS, true, 2, 0, 6, 4>; /* exact chain + prefetch two tiles ahead */
        case 3: return encode_fast_kernel<kCS, true, 2, 0, 0, 3>;
        case 5: return encode_fast_kernel<kCS, true, 2, 0, 0, 5>;
        case 84: return encode_fast_kernel<kCS, true, 2, 0, 8, 4>; /* tensor-map staging */
        case 34: return encode_fast_kernel<kCS, true, 2, 0, 3, 4>; /* diagnostics: no colour, no search */
        case 44: return encode_fast_kernel<kCS, true, 2, 0, 4, 4>; /* diagnostics: no search */
        case 54: return encode_fast_kernel<kCS, true, 2, 0, 5, 4>; /* diagnostics: no colour */
        default: break;
        }
    }
    if (sub && bytes == 2 && walk == 1) {
        switch (variant) {
        case 2: return encode_fast_kernel<kCS, true, 2, 1, 0, 2>;
        case 3: return encode_fast_kernel<kCS, true, 2, 1, 0, 3>;
        case 12: return encode_fast_kernel<kCS, true, 2, 1, 1, 2>;
        case 13: return encode_fast_kernel<kCS, true, 2, 1, 1, 3>;
        default: break;
        }
    }
#endif
    return nullptr;
}

dec_fn LUMA_CAT(get_decode_fast_cs, LUMA_TU_CS)(bool sub, int bytes, int variant)
{
    if (variant == 4) {
        if (sub)
            return bytes == 2 ? decode_fast_kernel<kCS, true, 2, 0, 4> : decode_fast_kernel<kCS, true, 1, 0, 4>;
        return bytes == 2 ? decode_fast_kernel<kCS, false, 2, 0, 4> : decode_fast_kernel<kCS, false, 1, 0, 4>;
    }
    if (variant == kDecVariantPrefetch) { /* the default: + L2 prefetch of the next tile's code words (656 -> 644 us per 32 4K frames) */
        if (sub)
            return bytes == 2 ? decode_fast_kernel<kCS, true, 2, 2, 4> : decode_fast_kernel<kCS, true, 1, 2, 4>;
        return bytes == 2 ? decode_fast_kernel<kCS, false, 2, 2, 4> : decode_fast_kernel<kCS, false, 1, 2, 4>;
    }
    if (variant == kDecVariantGlobalLut && bytes == 2) /* 14-16-bit LUTs: luma LUT read in place, + L2 prefetch */
        return sub ? decode_fast_kernel<kCS, true, 2, 18, 4> : decode_fast_kernel<kCS, false, 2, 18, 4>;
#if LUMA_TU_CS == 0
    if (sub && bytes == 2) {
        switch (variant) {
        case 3: return decode_fast_kernel<kCS, true, 2, 0, 3>;
        case 5: return decode_fast_kernel<kCS, true, 2, 0, 5>;
        case 64: return decode_fast_kernel<kCS, true, 2, 6, 4>; /* L2 prefetch two tiles ahead */
        case 13: return decode_fast_kernel<kCS, true, 2, 1, 3>;
        case 14: return decode_fast_kernel<kCS, true, 2, 1, 4>;
        case 15: return decode_fast_kernel<kCS, true, 2, 1, 5>;
        default: break;
        }
    }
#endif
    return nullptr;
}

#endif

} // namespace lumacu

/*
 * luma_dispatch.h -- host-visible table of the kernel instantiations.
 *
 * The kernels are compiled in one translation unit per colour space (luma_kern_tu.cu built with
 * -DLUMA_TU_CS=<0..3> and -DLUMA_TU_FAST=<0|1>) so that the build parallelises; lumacu.cu picks an
 * instantiation through these getters and launches it.
 */
#pragma once

#include "luma_kernels_decl.cuh"

namespace lumacu {

typedef void (*enc_fn)(const EncArgs);
typedef void (*dec_fn)(const DecArgs);

/* generic kernels: any colour space / profile / alignment / search mode */
#define LUMA_DECL_GENERIC(CSV)                                         \
    enc_fn get_encode_generic_cs##CSV(bool sub, int bytes, bool vec);  \
    dec_fn get_decode_generic_cs##CSV(bool sub, int bytes, bool vec);
LUMA_DECL_GENERIC(0)
LUMA_DECL_GENERIC(1)
LUMA_DECL_GENERIC(2)
LUMA_DECL_GENERIC(3)
#undef LUMA_DECL_GENERIC

/* tuned kernels (luma_fast.cuh); walk is the bucket walk length the search was planned with.
 * variant = 10 * PF + MINB (see luma_kern_tu.cu); every configuration exists as variant 4, the headline
 * configuration in a few more for the tuning sweep.
 * Return NULL when there is no instantiation for the request. */
constexpr int kEncVariantPlain = 4, kDecVariantPlain = 4;
constexpr int kDecVariantGlobalLut = 1024; /* decode: luma LUT stays in global memory (luma_fast.cuh PF bit 4) */
constexpr int kDecVariantPrefetch = 24; /* decode: plain loads + L2 prefetch of the next tile (luma_fast.cuh PF 2) */
constexpr int kEncVariantScreened = 67; /* Lu'v' 4:2:0: screened chroma, queued redo, L2 prefetch two tiles ahead (luma_fast.cuh FASTC 2, PF 6) */
constexpr unsigned kEncStagedSmemBytes = 6u * 512u * 8u; /* luma_fast.cuh kEncStageBlock */
#define LUMA_DECL_FAST(CSV)                                          \
    enc_fn get_encode_fast_cs##CSV(bool sub, int bytes, int walk, int variant, bool prescale); \
    dec_fn get_decode_fast_cs##CSV(bool sub, int bytes, int variant);
LUMA_DECL_FAST(0)
LUMA_DECL_FAST(1)
LUMA_DECL_FAST(2)
LUMA_DECL_FAST(3)
#undef LUMA_DECL_FAST

/* element-wise API kernels (compiled in the CS 1 generic unit) */
void launch_transform(int cs, bool fwd, unsigned blocks, cudaStream_t st, float *c0, float *c1, float *c2, size_t n,
                      float sc, float l_max);
void launch_quantize(unsigned blocks, size_t smem, cudaStream_t st, const QuantDev &q, const float *in, float *out,
                     size_t n, int use_lut);
void launch_dequantize(unsigned blocks, cudaStream_t st, const QuantDev &q, const float *in, float *out, size_t n,
                       int use_lut);
const void *quantize_kernel_ptr();
/* CS_YCBCR PQ tables (luma_pq_tables.cuh): pqd = kPqTabPqdBytes, pqe = kPqTabPqeBytes of device memory */
constexpr size_t kPqTabPqdBytes = (size_t)((0x3F800000u - 0x3B800000u + 1u + 31u) / 32u) * 16u;
constexpr size_t kPqTabPqeBytes = (size_t)(0x3F8147AEu - 0x3F55C28Fu + 1u) * 4u;
void launch_build_pq_tables(unsigned blocks, cudaStream_t st, void *pqd, float *pqe, float l_max);
void launch_check_lmax_division(unsigned blocks, cudaStream_t st, float l_max, float rc, uint32_t *bad);
/* v-keyed luma search table of CS_YCBCR encode: kVdTabBytes of device memory + one flag word (non-zero = unusable) */
constexpr size_t kVdTabBytes = (size_t)((((0x3F800000u >> 13) - (0x3D000000u >> 13) + 1u) + 3u) & ~3u) * 4u;
/* PQ encode of every half-float bit pattern (65 536 floats) for one preScaling / Lmax */
void launch_build_pqh(cudaStream_t st, const QuantDev &q, float *tab, float sc, int prescale, float l_max);
void launch_build_vdtab(cudaStream_t st, const QuantDev &q, uint32_t *tab, uint32_t *bad, float l_max);
void launch_test_frame(unsigned blocks, cudaStream_t st, float *rgb, uint32_t w, uint32_t h);
void launch_half_rgba_to_frame(unsigned blocks, cudaStream_t st, const void *rgba, float *rgb, size_t n, size_t plane_stride, int mode);
void launch_frame_to_half_rgba(unsigned blocks, cudaStream_t st, const float *rgb, void *rgba, size_t n, size_t plane_stride);
void launch_pfs_channels(bool to_rgb, unsigned blocks, cudaStream_t st, const float *a0, const float *a1, const float *a2,
                         float *o0, float *o1, float *o2, size_t n);
void launch_display_linear(unsigned blocks, unsigned n_frames, cudaStream_t st, const DecArgs &a, int cs, int sub, int bytes);

} // namespace lumacu

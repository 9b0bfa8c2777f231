/*
 * luma_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the Luma HDRv per-pixel HDR<->integer transform,
 * used as the parity checker for the CUDA path.  Only tests/, the smoke entry
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product path (lumahdrv_b200/csrc, lumahdrv_b200/host) never links it.
 *
 * Parity pinning: the restatement is checked bit-for-bit against the compiled
 * reference itself (oracle/_ref/libluma_ref.so, built in place from
 * /root/reference) and against FNV-1a plane/LUT hashes derived from the
 * reference (tests/golden/).  See oracle/README.md.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference checkout).
 */
#ifndef LUMA_ORACLE_H
#define LUMA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* enum order is wire format: include/luma/luma_quantizer.h:95-96 */
enum { LO_PTF_PSI = 0, LO_PTF_PQ = 1, LO_PTF_LOG = 2, LO_PTF_JND_HDRVDP = 3, LO_PTF_LINEAR = 4 };
enum { LO_CS_LUV = 0, LO_CS_RGB = 1, LO_CS_YCBCR = 2, LO_CS_XYZ = 3 };

typedef struct lo_quantizer {
    int ptf;
    int color_space;
    unsigned bitdepth, bitdepth_color;
    unsigned max_val, max_val_color;
    float l_max, l_min;
    float *mapping; /* max_val + 1 entries */
} lo_quantizer;

/* src/luma_quantizer.cpp:45-58 */
void lo_init(lo_quantizer *q);
void lo_free(lo_quantizer *q);

/* src/luma_quantizer.cpp:172-212 (+ :114-169).  Returns 0, or -1 when a table
 * PTF is requested from a build without the reference's ptfs/ tables. */
int lo_set_quantizer(lo_quantizer *q, int ptf, unsigned bitdepth, int cs,
                     unsigned bitdepth_c, float max_lum, float min_lum);

/* src/luma_quantizer.cpp:215-244 / :247-264 */
float lo_quantize(const lo_quantizer *q, float val, unsigned ch);
float lo_dequantize(const lo_quantizer *q, float val, unsigned ch);

/* src/luma_quantizer.cpp:267-482; frame is planar f32, 3 planes of w*h
 * (include/luma/luma_frame.h:83-86).  Returns 1 on success, 0 on unknown cs. */
int lo_transform_color_space(const lo_quantizer *q, float *frame, unsigned w,
                             unsigned h, int to_cs, float sc);

/* src/luma_quantizer.cpp:485-501 / :504-510 */
float lo_transform_pq(const lo_quantizer *q, float val, int encode);
float lo_transform_log(const lo_quantizer *q, float val, int encode);

/* src/luma_encoder.cpp:260-317, run for planes 0..2 (:196-201) on a frame that
 * already went through lo_transform_color_space(...,1,...).  profile: 0=420/8,
 * 1=444/8, 2=420/16, 3=444/16.  avg_out[plane] receives the sequential fp32
 * mean the reference computes (only plane 0 is ever reported, :314-316). */
void lo_pack_planes(const lo_quantizer *q, const float *frame, unsigned w,
                    unsigned h, int profile, uint8_t *const planes[3],
                    const int strides[3], float avg_out[3]);

/* src/luma_decoder.cpp:205-240: planes -> 3 full-resolution f32 planes (no
 * colour transform yet). */
void lo_unpack_planes(const lo_quantizer *q, const uint8_t *const planes[3],
                      const int strides[3], unsigned w, unsigned h, int profile,
                      float *frame);

/* include/luma/luma_encoder.h:142-148 minus run(): colour transform (in place,
 * the reference mutates the caller's frame) + pack. */
void lo_encode(const lo_quantizer *q, float *frame, unsigned w, unsigned h,
               int profile, float pre_scaling, uint8_t *const planes[3],
               const int strides[3], float avg_out[3]);

/* include/luma/luma_decoder.h:143-161 minus run(). */
void lo_decode(const lo_quantizer *q, const uint8_t *const planes[3],
               const int strides[3], unsigned w, unsigned h, int profile,
               float pre_scaling, float *frame);

/* src/exr_interface.cpp:50-70 */
void lo_test_frame(float *frame, unsigned w, unsigned h);

/* Plane geometry helpers (src/luma_encoder.cpp:265-269; libvpx 1.6.1
 * vpx/src/vpx_image.c pitch rule for vpx_img_alloc(...,32)). */
void lo_plane_dims(unsigned w, unsigned h, int profile, int pw[3], int ph[3]);
void lo_vpx_strides(unsigned w, int profile, int align, int strides[3]);

/* FNV-1a 32 (offset 2166136261, prime 16777619) */
uint32_t lo_fnv1a32(const void *data, size_t n, uint32_t seed);
/* hash of a pitched plane's payload bytes only (padding excluded) */
uint32_t lo_hash_plane(const uint8_t *plane, int stride, int row_bytes, int rows);

/* 1 when the PSI / JND-HDR-VDP tables were compiled in */
int lo_have_ptf_tables(void);

#ifdef __cplusplus
}
#endif
#endif

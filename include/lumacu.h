/*
 * lumacu.h -- C ABI of the B200 (sm_100a) implementation of Luma HDRv's
 * per-pixel HDR<->integer transform.
 *
 * This is the drop-in boundary for ONE stage of the reference codec: the path
 *
 *   LumaEncoder::encode  = LumaQuantizer::transformColorSpace(frame,true,sc)
 *                          + LumaEncoder::setVpxChannel x3  (-> LumaQuantizer::quantize)
 *   LumaDecoder::decode  = LumaDecoder::getVpxChannels (-> LumaQuantizer::dequantize)
 *                          + LumaQuantizer::transformColorSpace(frame,false,sc)
 *
 * (reference: include/luma/luma_encoder.h:142-148, src/luma_encoder.cpp:260-317,
 * include/luma/luma_decoder.h:143-161, src/luma_decoder.cpp:205-240,
 * src/luma_quantizer.cpp:172-510).  VP9 (libvpx) and Matroska stay on the host.
 *
 * Conventions
 *  - plain C types only; no exceptions cross this boundary; every call returns
 *    a lumacu_status (0 = OK) and records a message retrievable with
 *    lumacu_last_error().
 *  - float frames are planar f32 exactly like LumaFrame
 *    (include/luma/luma_frame.h:83-86): buffer[c*h*w + y*w + x], c = 0..2.
 *  - integer planes are the vpx_image_t planes the reference fills/reads:
 *    8-bit samples (profiles 0,1) or little-endian 16-bit samples (profiles
 *    2,3) with a byte pitch stride[plane]; profiles 0,2 are 4:2:0 (chroma
 *    planes ((w+1)>>1) x ((h+1)>>1)), profiles 1,3 are 4:4:4
 *    (src/luma_encoder.cpp:121-128,265-269).
 *  - "_dev" entry points take device pointers and a cudaStream_t (passed as
 *    void*) and are asynchronous; the others take host pointers, stage through
 *    pinned buffers and return when the result is in host memory.  A NULL
 *    stream selects the context's own (non-blocking) stream; to launch on the
 *    CUDA default stream pass cudaStreamLegacy (0x1) or cudaStreamPerThread (0x2).
 *  - the host-pointer entry points accept any host memory, but only page-locked memory (lumacu_host_alloc,
 *    lumacu_host_register) moves at PCIe rate; pageable memory is staged by the driver, several times slower.  The
 *    first time a context sees a pageable frame or plane buffer it says so on stderr (LUMACU_QUIET=1 silences it).
 *  - a context is thread-compatible, not thread-safe: one host thread at a time.  The "_dev" entry points may be used on
 *    several caller streams of the same context (the statistics workspace is kept per stream; the quantizer tables are
 *    read-only between lumacu_set_quantizer calls, which synchronise the whole device before replacing them); the
 *    host-pointer entry points have ONE call in flight per context.
 *  - results are bit-identical to the reference CPU path for the integer
 *    planes and for the decoded floats (see DESIGN.md for the one documented
 *    libm dependency: CS_YCBCR calls powf per pixel).
 */
#ifndef LUMACU_H
#define LUMACU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LUMACU_VERSION 200 /* 0.2.0 */

typedef enum lumacu_status {
    LUMACU_OK = 0,
    LUMACU_ERR_INVALID_ARGUMENT = 1,
    LUMACU_ERR_CUDA = 2,
    LUMACU_ERR_NOT_CONFIGURED = 3, /* lumacu_set_quantizer not called yet */
    LUMACU_ERR_UNSUPPORTED = 4,
    LUMACU_ERR_OUT_OF_MEMORY = 5,
    LUMACU_ERR_NO_DEVICE = 6
} lumacu_status;

/* Wire-format enum values of the reference (include/luma/luma_quantizer.h:95-96);
 * they are memcpy'd into Matroska attachments 432/433, so the order is fixed. */
typedef enum lumacu_ptf {
    LUMACU_PTF_PSI = 0,
    LUMACU_PTF_PQ = 1,
    LUMACU_PTF_LOG = 2,
    LUMACU_PTF_JND_HDRVDP = 3,
    LUMACU_PTF_LINEAR = 4
} lumacu_ptf;

typedef enum lumacu_color_space {
    LUMACU_CS_LUV = 0,
    LUMACU_CS_RGB = 1,
    LUMACU_CS_YCBCR = 2,
    LUMACU_CS_XYZ = 3
} lumacu_color_space;

/* Per-frame reduction over plane 0 after the colour transform.  sum/count is
 * the mean the reference compares against 1.0 for its "is input calibrated?"
 * warning (src/luma_encoder.cpp:276,294,314-316); max is an extra output the
 * reference does not have. */
typedef struct lumacu_frame_stats {
    double sum;  /* sum of plane-0 samples (Y for Lu'v'), accumulated in fp64 */
    float max;   /* max of plane-0 samples (NaNs ignored) */
    float min;   /* min of plane-0 samples (NaNs ignored) */
} lumacu_frame_stats;

typedef struct lumacu_ctx lumacu_ctx;

/* ---- library / context ------------------------------------------------------ */
int lumacu_version(void);
const char *lumacu_status_name(int status);
int lumacu_device_count(int *count);
/* 1 when the PSI / JND-HDR-VDP tables (reference include/luma/ptfs/, consumed at
 * build time) were compiled in, 0 otherwise. */
int lumacu_have_ptf_tables(void);

/* Creates a context bound to CUDA device `device` (own stream, pinned staging
 * buffers grown on demand).  Replaces nothing in the reference; it is the state
 * a LumaEncoder / LumaDecoder object carries next to its LumaQuantizer. */
int lumacu_create(int device, lumacu_ctx **out);
int lumacu_destroy(lumacu_ctx *ctx);
/* Message of the last failing call on `ctx` (or of the last failing
 * lumacu_create on this thread when ctx is NULL).  Never NULL. */
const char *lumacu_last_error(const lumacu_ctx *ctx);
int lumacu_device(const lumacu_ctx *ctx);
/* Block until all work queued on the context's own stream has finished. */
int lumacu_synchronize(lumacu_ctx *ctx);
/* The context's own cudaStream_t (as void*). */
void *lumacu_stream(lumacu_ctx *ctx);
/* Page-locked host memory for frames / planes handed to the host-pointer entry
 * points (what LumaFrame::buffer and the vpx image planes should live in for
 * full PCIe rate).  Not tied to a context. */
int lumacu_host_alloc(size_t bytes, void **out);
int lumacu_host_free(void *p);
/* Page-lock memory the caller already owns (e.g. the vpx_image_t planes libvpx allocated,
 * src/luma_encoder.cpp:121-128) so that copies to/from it are direct DMA.  Registering twice is OK. */
int lumacu_host_register(void *p, size_t bytes);
int lumacu_host_unregister(void *p);

/* ---- quantizer -------------------------------------------------------------- */
/* Host-side LUT construction = the table half of LumaQuantizer::setQuantizer
 * (src/luma_quantizer.cpp:172-212, setMapping* :114-169, transformPQ/Log
 * :485-510).  Writes (2^bitdepth) floats to `lut_out` (capacity `cap`, in
 * floats).  Runs on the host so that PQ/LOG entries come from the same libm
 * (powf/log10f) the reference uses. */
int lumacu_build_lut(int ptf, unsigned bitdepth, float max_lum, float min_lum, float *lut_out,
                     size_t cap);

/* ---- quantizer metadata wire format (host only) ---------------------------------- */
/* The reference carries the quantizer from encoder to decoder as seven Matroska attachments whose payloads
 * are the raw host bytes of the values (src/luma_encoder.cpp:78-106, read back at src/luma_decoder.cpp:79-122):
 *   430 u32 ptfBitDepth | 431 u32 colorBitDepth | 432 i32 ptf | 433 i32 colorSpace |
 *   434 float[maxVal] the LUT WITHOUT its last entry (getSize() = maxVal floats are written) |
 *   435 float preScaling | 436 float[2] {maxLum, minLum}
 * lumacu_metadata_pack writes them back to back as records {u32 id, u32 size, payload} (little endian) so that
 * a worker can be configured from a byte stream without libmatroska; lumacu_metadata_unpack does what
 * LumaDecoder::initialize does: take the scalars, rebuild the table from them (lumacu_build_lut) and overlay
 * the stored entries.  Records with other ids are skipped; 430..434 are mandatory like in the reference. */
typedef struct lumacu_metadata {
    uint32_t ptf_bit_depth;
    uint32_t color_bit_depth;
    int32_t ptf;          /* lumacu_ptf */
    int32_t color_space;  /* lumacu_color_space */
    float pre_scaling;
    float max_lum;
    float min_lum;
} lumacu_metadata;
/* lut = the encoder's table (2^ptf_bit_depth entries, e.g. from lumacu_build_lut).  Returns the number of
 * bytes needed in *used (also when blob is NULL or cap is too small: LUMACU_ERR_INVALID_ARGUMENT then). */
int lumacu_metadata_pack(const lumacu_metadata *m, const float *lut, uint32_t lut_len, uint8_t *blob, size_t cap,
                         size_t *used);
/* lut_out receives 2^ptf_bit_depth entries (capacity lut_cap floats); *lut_len is set to that count. */
int lumacu_metadata_unpack(const uint8_t *blob, size_t size, lumacu_metadata *m, float *lut_out, size_t lut_cap,
                           uint32_t *lut_len);

/* Host-side analysis of LumaQuantizer::quantize's luma branch
 * (src/luma_quantizer.cpp:219-235).  For a finite, strictly increasing LUT the
 * reference's "bisect, then pick the nearer neighbour with fp32 differences" is
 * a monotone step function of val; thr_keys[k-1] receives the order-preserving
 * integer key (sign-magnitude -> unsigned) of the smallest float whose code is
 * >= k, for k = 1 .. lut_len-1.  Returns 1 on success, 0 when the LUT is not
 * finite/strictly increasing (the kernels then replay the reference loop
 * literally).  Pure host code; exposed so that tests can pin the thresholds
 * against the oracle without a GPU. */
int lumacu_derive_thresholds(const float *lut, uint32_t lut_len, uint32_t *thr_keys);
/* Bucket table geometry chosen for a threshold list: bucket = key >> shift,
 * table covers [base, base + n_buckets), walk = max thresholds per bucket. */
int lumacu_plan_buckets(const uint32_t *thr_keys, uint32_t n_thr, uint32_t *shift, uint32_t *base,
                        uint32_t *n_buckets, uint32_t *walk);

/* Device-side quantizer state = the fields LumaQuantizer keeps
 * (include/luma/luma_quantizer.h:120-126): the code->luminance LUT
 * (lut_len = maxVal+1 entries), maxValColor = 2^colorBits-1, the colour space
 * and Lmax (used per pixel only by CS_YCBCR's PQ).  May be called again at any
 * time, e.g. after the decoder overwrote the LUT from attachment 434
 * (src/luma_decoder.cpp:121-122).  The encode side additionally derives exact
 * decision thresholds from the LUT (see DESIGN.md "luma search"). */
int lumacu_set_quantizer(lumacu_ctx *ctx, const float *lut, uint32_t lut_len,
                         uint32_t max_val_color, int color_space, float max_lum);

/* Single-process multi-GPU: every context in ctxs[0..n) receives ctxs[root]'s quantizer.  The host-side derivation
 * (LUT, exact decision thresholds, search tables) is NOT repeated: the root's device tables travel to each peer in one
 * peer-to-peer copy (NVLink / NVSwitch between B200s), so all GPUs search with bit-identical tables; call it again
 * after the root's table changed, e.g. when the decoder overlaid attachment 434 (src/luma_decoder.cpp:121-122).
 * The reference has one quantizer per process and nothing to replace here; processes on different GPUs exchange
 * lumacu_metadata_pack blobs instead (NCCL broadcast, lumahdrv_b200/shard.py). */
int lumacu_broadcast_quantizer(lumacu_ctx *const ctxs[], int n, int root);
/* Host copy of the quantizer a context currently holds; any output pointer may be NULL. */
int lumacu_get_quantizer(const lumacu_ctx *ctx, float *lut_out, size_t cap, uint32_t *lut_len,
                         uint32_t *max_val_color, int *color_space, float *max_lum);

/* ---- whole-frame transform, host memory ------------------------------------- */
/* LumaEncoder::encode minus run() (include/luma/luma_encoder.h:142-148):
 * rgb (3*w*h f32, NOT modified unless write_back != 0, in which case it
 * receives the colour-transformed planes exactly like the reference's in-place
 * side effect, src/luma_quantizer.cpp:281-313) -> planes[0..2].
 * w and h must be even and > 0 (src/luma_encoder.cpp:118-119).
 * stats may be NULL. */
int lumacu_encode(lumacu_ctx *ctx, float *rgb, uint32_t w, uint32_t h, int profile,
                  float pre_scaling, uint8_t *const planes[3], const int32_t strides[3],
                  int write_back, lumacu_frame_stats *stats);

/* LumaDecoder::decode minus run() (include/luma/luma_decoder.h:143-161):
 * planes -> rgb (3*w*h f32). */
int lumacu_decode(lumacu_ctx *ctx, const uint8_t *const planes[3], const int32_t strides[3],
                  uint32_t w, uint32_t h, int profile, float pre_scaling, float *rgb);

/* EXR at either end without the f32 detour over PCIe.  The reference's drivers read their frames from OpenEXR files as
 * half-float Imf::Rgba pixels (ExrInterface::readFrame, src/exr_interface.cpp:73-143; lumaenc.cpp) and lumadec writes them
 * back the same way (ExrInterface::writeFrame, :157-187); between the file and LumaEncoder::encode / after
 * LumaDecoder::decode the CPU expands / rounds every pixel.  These two calls take / deliver the half pixels themselves
 * (w*h*8 bytes, host memory): the bus carries 8 B/px instead of 12 and the pixel loops run on the device, band by band
 * inside the same three-stream pipeline.  Results are those of the reference's own sequence: planes identical to
 * readFrame's loop followed by lumacu_encode (channels = Imf::RgbaChannels as for lumacu_half_rgba_to_frame_dev);
 * pixels identical to lumacu_decode followed by writeFrame's loop (lumacu_frame_to_half_rgba_dev). */
int lumacu_encode_half_rgba(lumacu_ctx *ctx, const void *rgba_half, uint32_t w, uint32_t h, int channels, int profile,
                            float pre_scaling, uint8_t *const planes[3], const int32_t strides[3],
                            lumacu_frame_stats *stats);
int lumacu_decode_half_rgba(lumacu_ctx *ctx, const uint8_t *const planes[3], const int32_t strides[3], uint32_t w,
                            uint32_t h, int profile, float pre_scaling, void *rgba_half);

/* Asynchronous forms of the two calls above: the copies and the kernel are queued on the context's streams and the
 * call returns at once, so that the host thread can do something else meanwhile -- run libvpx on the previous
 * frame's planes (what LumaEncoder::run does between two encode() calls, include/luma/luma_encoder.h:142-148), or
 * drive the next GPU.  ONE call in flight per context: any later host-pointer call on the same context first
 * completes it.
 *   lumacu_wait_input  returns when the call's INPUT buffer (rgb for encode -- unless write_back -- / planes for
 *                      decode) has been read completely and may be reused;
 *   lumacu_wait        returns when the results are in host memory (and *stats is filled);
 *   lumacu_pending     1 while a call is in flight. */
int lumacu_encode_async(lumacu_ctx *ctx, float *rgb, uint32_t w, uint32_t h, int profile,
                        float pre_scaling, uint8_t *const planes[3], const int32_t strides[3],
                        int write_back, lumacu_frame_stats *stats);
int lumacu_decode_async(lumacu_ctx *ctx, const uint8_t *const planes[3], const int32_t strides[3],
                        uint32_t w, uint32_t h, int profile, float pre_scaling, float *rgb);
int lumacu_wait_input(lumacu_ctx *ctx);
int lumacu_wait(lumacu_ctx *ctx);
int lumacu_pending(const lumacu_ctx *ctx);

/* The two halves of the reference's unfused path, for callers that use them separately:
 * lumacu_quantize_planes = LumaEncoder::setChannels (src/luma_encoder.cpp:196-201,260-317): `frame` is
 * ALREADY colour-transformed (e.g. by lumacu_transform_color_space); [2x2 mean,] quantize, pack.
 * lumacu_dequantize_planes = LumaDecoder::getVpxChannels (src/luma_decoder.cpp:205-240): unpack,
 * dequantize, [2x2 replicate]; the result still has to go through the inverse colour transform. */
int lumacu_quantize_planes(lumacu_ctx *ctx, const float *frame, uint32_t w, uint32_t h, int profile,
                           uint8_t *const planes[3], const int32_t strides[3],
                           lumacu_frame_stats *stats);
int lumacu_dequantize_planes(lumacu_ctx *ctx, const uint8_t *const planes[3],
                             const int32_t strides[3], uint32_t w, uint32_t h, int profile,
                             float *frame);

/* LumaQuantizer::transformColorSpace(frame, toCs, sc) (src/luma_quantizer.cpp:
 * 267-482): in-place colour transform of a planar f32 frame. */
int lumacu_transform_color_space(lumacu_ctx *ctx, float *frame, uint32_t w, uint32_t h, int to_cs,
                                 float sc);

/* Element-wise LumaQuantizer::quantize / dequantize (src/luma_quantizer.cpp:
 * 215-264) over n values of channel `ch`; codes are returned as floats like
 * the reference does. */
int lumacu_quantize(lumacu_ctx *ctx, const float *in, float *out, size_t n, unsigned ch);
int lumacu_dequantize(lumacu_ctx *ctx, const float *in, float *out, size_t n, unsigned ch);

/* ---- whole-frame transform, device memory (asynchronous) -------------------- */
/* Batched: n_frames frames per launch.  Frame f reads d_rgb + f*rgb_frame_stride
 * (in floats; 0 means 3*w*h) and writes d_planes[p] + f*plane_frame_stride[p]
 * (bytes).  d_stats (device pointer, n_frames entries) may be NULL.
 * d_rgb_out, if not NULL, receives the colour-transformed frame (the
 * reference's in-place side effect); it may alias d_rgb. */
int lumacu_encode_dev(lumacu_ctx *ctx, const float *d_rgb, float *d_rgb_out, uint32_t w,
                      uint32_t h, int profile, float pre_scaling, uint8_t *const d_planes[3],
                      const int32_t strides[3], uint32_t n_frames, size_t rgb_frame_stride,
                      const size_t plane_frame_stride[3], lumacu_frame_stats *d_stats,
                      void *stream);

int lumacu_decode_dev(lumacu_ctx *ctx, const uint8_t *const d_planes[3], const int32_t strides[3],
                      uint32_t w, uint32_t h, int profile, float pre_scaling, float *d_rgb,
                      uint32_t n_frames, size_t rgb_frame_stride,
                      const size_t plane_frame_stride[3], void *stream);

int lumacu_transform_color_space_dev(lumacu_ctx *ctx, float *d_frame, uint32_t w, uint32_t h,
                                     int to_cs, float sc, void *stream);
int lumacu_quantize_dev(lumacu_ctx *ctx, const float *d_in, float *d_out, size_t n, unsigned ch,
                        void *stream);
int lumacu_dequantize_dev(lumacu_ctx *ctx, const float *d_in, float *d_out, size_t n, unsigned ch,
                          void *stream);

/* ---- display decode ------------------------------------------------------------- */
/* What the reference's player does in its fragment shader (src/lumaplay_dequantizer.frag:70-157, parameters set
 * at lumaplay.cpp:395-411): dequantise + inverse colour transform (here: exactly the LumaDecoder::decode path,
 * nearest-neighbour chroma like the CPU decoder rather than the shader's bilinear texture fetch), then
 *   RGB * exposure / scaling        (scaling = preScaling / user_scaling; or the 8-bit "LDR simulation")
 *   optional sigmoid tone curve     v^0.8 / (v^0.8 + 0.8^0.8)
 *   display gamma                   v^(1/gamma)
 * into an 8-bit RGBA image (alpha = 255).  Floating-point pow makes this path approximate by nature, like the GLSL
 * original; tests compare against a numpy restatement with a 1-LSB tolerance. */
typedef struct lumacu_display_params {
    float exposure;     /* lumaplay's exposure multiplier, 1 = none */
    float gamma;        /* display gamma, e.g. 2.2 */
    float user_scaling; /* lumaplay's userScaling, 1 = none */
    int do_tmo;         /* sigmoid tone curve on/off */
    int ldr_sim;        /* 8-bit LDR simulation on/off */
    int filter;         /* 0: sample like the CPU decoder (nearest-neighbour chroma, exact LUT entries, exact
                         * LumaDecoder::decode arithmetic); 1: sample like the player itself -- its textures are
                         * GL_LINEAR, CLAMP_TO_EDGE (lumaplay.cpp:258-259), so at 1:1 scale 4:2:0 chroma is bilinear
                         * with weights 1/4 : 3/4 per direction and the luminance fetch from the one-short LUT
                         * texture (lumaplay.cpp:371) is the mean of lut[code-1] and lut[code]; the rest of the
                         * shader in plain fp32 (PQ with its built-in L = 10000) */
} lumacu_display_params;
int lumacu_display(lumacu_ctx *ctx, const uint8_t *const planes[3], const int32_t strides[3], uint32_t w,
                   uint32_t h, int profile, float pre_scaling, const lumacu_display_params *params,
                   uint8_t *rgba, int32_t rgba_pitch);
int lumacu_display_dev(lumacu_ctx *ctx, const uint8_t *const d_planes[3], const int32_t strides[3], uint32_t w,
                       uint32_t h, int profile, float pre_scaling, const lumacu_display_params *params,
                       uint8_t *d_rgba, int32_t rgba_pitch, uint32_t n_frames,
                       const size_t plane_frame_stride[3], size_t rgba_frame_stride, void *stream);

/* ---- frame sources on the device ---------------------------------------------- */
/* ExrInterface::testFrame (src/exr_interface.cpp:50-70): the reference's synthetic HDR test pattern generated
 * in device memory (3*w*h f32, planar), bit-identical to the host function. */
int lumacu_test_frame_dev(lumacu_ctx *ctx, float *d_rgb, uint32_t w, uint32_t h, void *stream);
/* The pixel loop of ExrInterface::readFrame (src/exr_interface.cpp:73-143): interleaved half-float RGBA pixels
 * (Imf::Rgba, 8 bytes each, device memory) -> planar f32 frame.  channels = Imf::RgbaChannels of the file:
 * 1 (R), 2 (G), 4 (B) replicate one channel, 7 (RGB) / 15 (RGBA) copy three; anything else fails like the
 * reference ("luminance only frames not yet supported"). */
int lumacu_half_rgba_to_frame_dev(lumacu_ctx *ctx, const void *d_rgba_half, uint32_t w, uint32_t h, int channels,
                                  float *d_rgb, void *stream);
/* The pixel loop of ExrInterface::writeFrame (src/exr_interface.cpp:157-187), the sink lumadec writes decoded frames
 * to: planar f32 frame -> interleaved half-float RGBA pixels (8 bytes each, alpha 0).  float -> half as Imf's half(float)
 * does it: round to nearest even, overflow to infinity, NaN keeps sign and top payload bits.  OpenEXR is not part of the
 * reference tree ("parity unpinned" against it; the test compares with numpy's float16 cast, the same function). */
int lumacu_frame_to_half_rgba_dev(lumacu_ctx *ctx, const float *d_rgb, uint32_t w, uint32_t h, void *d_rgba_half,
                                  void *stream);

/* PfsInterface::readFrame / writeFrame (src/pfs_interface.cpp:57-113, :115-152) minus the stream parsing: a PFS frame
 * carries X, Y, Z as three separate w*h float arrays; the reference runs pfstools' pfs::transformColorSpace
 * (CS_XYZ -> CS_RGB at :84, CS_RGB -> CS_XYZ at :140) and copies the channels into / out of the planar frame.
 * pfstools is not part of the reference tree ("parity unpinned"): its D65 matrices carry the same nine constants as
 * the reference's own xyz2rgbMat / rgb2xyzMat (include/luma/luma_quantizer.h:79-87), applied per pixel as
 * m0*a + m1*b + m2*c in float without clamping -- which is what these kernels do.  d_rgb is 3*w*h floats. */
int lumacu_pfs_xyz_to_frame_dev(lumacu_ctx *ctx, const float *d_x, const float *d_y, const float *d_z, uint32_t w,
                                uint32_t h, float *d_rgb, void *stream);
int lumacu_frame_to_pfs_xyz_dev(lumacu_ctx *ctx, const float *d_rgb, uint32_t w, uint32_t h, float *d_x, float *d_y,
                                float *d_z, void *stream);

/* ---- introspection (used by bench.py / tests) -------------------------------- */
/* Number of kernels this context has launched so far. */
uint64_t lumacu_launch_count(const lumacu_ctx *ctx);
/* Describes how the luma search was configured by the last set_quantizer:
 * mode 0 = bucket table + threshold walk (shared memory), 1 = binary search
 * over the thresholds, 2 = literal replica of the reference's bisection over
 * the LUT (LUT not strictly increasing).  This is the GENERIC kernels' search; the tuned kernels replace modes 0 and 1
 * by direct tables where one exists (shared memory up to 12-bit LUTs, global memory for 13-16 bits). */
int lumacu_search_info(const lumacu_ctx *ctx, int *mode, uint32_t *n_buckets, uint32_t *shift,
                       uint32_t *walk);

/* Kernel selection.  path 0 (default): use the tuned kernels (packed fp32x2 math, shared-memory
 * search tables) whenever their preconditions hold, the generic kernels otherwise; path 1: always
 * the generic kernels (a literal transcription of the reference loops; used by the tests to
 * cross-check the two).  Both produce identical bits. */
int lumacu_set_kernel_path(lumacu_ctx *ctx, int path);
/* CS_YCBCR: the tuned kernels read two of the per-pixel PQ powers from exhaustive device-built tables (45 MB per
 * context, L2-resident; DESIGN.md "YCbCr").  enable = 0 makes them evaluate every powf per pixel instead (tests and
 * sweeps; same bits either way).  Default: enabled. */
int lumacu_set_pq_tables(lumacu_ctx *ctx, int enable);
/* 1 if the last encode/decode launch on this context ran a tuned kernel, 0 if generic. */
int lumacu_last_kernel_path(const lumacu_ctx *ctx);
/* Host-pointer entry points cut a frame into row bands whose H2D copy, kernel and D2H copy overlap on three
 * streams (DESIGN.md "host staging").  0 = automatic (about one band per 8 MiB of frame, at most 8). */
int lumacu_set_host_bands(lumacu_ctx *ctx, int bands);
/* Tuning sweep / tests: pick an instantiation of the tuned kernels instead of the default (0).  All variants produce
 * identical bits; unknown or inapplicable ones fall back to the default family.
 *   enc_variant   4  exact chroma chain, plain loads (the round-1 algorithm; every configuration)
 *                67  screened chroma + queued redo + L2 prefetch two tiles ahead (any Lu'v' 4:2:0 quantizer; the default
 *                    for 8-bit chroma)
 *                 6, 7, 26, 27, 86, 87, 64, 84, 3, 5, 12, 13, 34, 44, 54   headline configuration only: redo policy,
 *                    prefetch distance, tensor-map staging, occupancy and arithmetic-skipping diagnostics
 *                    (lumahdrv_b200/csrc/luma_kern_tu.cu lists them)
 *            + 1000  bucket + threshold luma search where a direct search table would be used
 *            + 2000  the (32-bit-entry) direct search table read from global memory instead of staged into shared memory
 *   dec_variant   4  plain loads;  24  + L2 prefetch of the next tile's code words (the default);  64, 3, 5, 13-15 headline only
 *            + 2000  CS_YCBCR: the green of both pixels of a pair is evaluated (two exact powf) instead of every other one
 *   blocks_per_sm_cap   cap on the resident blocks per SM of the persistent grid (0 = what the occupancy calculator allows),
 *                    + 100 * T sizes the blocks of a multi-frame launch for T tiles per thread. */
int lumacu_set_tuning(lumacu_ctx *ctx, int enc_variant, int dec_variant, int blocks_per_sm_cap);

#ifdef __cplusplus
}
#endif
#endif /* LUMACU_H */

/*
 * vpx_stub.h -- TEST DOUBLE, part of the oracle (test infrastructure only).
 *
 * A minimal stand-in for the slice of the libvpx 1.6.1 API that the reference's
 * src/luma_encoder.cpp and src/luma_decoder.cpp touch, so that those two files
 * (and with them the *real* private plane loops LumaEncoder::setVpxChannel /
 * LumaDecoder::getVpxChannels) compile unmodified from /root/reference without
 * building libvpx.  The "codec" is a lossless loopback: encode serialises the
 * raw planes into one packet, decode re-materialises them in an image with a
 * different, wider pitch (as libvpx's decoder does), which is exactly what the
 * reference's lossLess=1 VP9 round trip yields for the integer planes.
 *
 * Written from scratch for this repo; nothing here is libvpx source.
 */
#ifndef LUMA_ORACLE_VPX_STUB_H
#define LUMA_ORACLE_VPX_STUB_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- images --------------------------------------------------------------- */
#define VPX_IMG_FMT_PLANAR 0x100
#define VPX_IMG_FMT_HIGHBITDEPTH 0x800
typedef enum vpx_img_fmt {
    VPX_IMG_FMT_NONE = 0,
    VPX_IMG_FMT_I420 = VPX_IMG_FMT_PLANAR | 2,
    VPX_IMG_FMT_I444 = VPX_IMG_FMT_PLANAR | 6,
    VPX_IMG_FMT_I42016 = VPX_IMG_FMT_I420 | VPX_IMG_FMT_HIGHBITDEPTH,
    VPX_IMG_FMT_I44416 = VPX_IMG_FMT_I444 | VPX_IMG_FMT_HIGHBITDEPTH
} vpx_img_fmt_t;

typedef struct vpx_image {
    vpx_img_fmt_t fmt;
    unsigned int w, h, bit_depth;
    unsigned int d_w, d_h;
    unsigned int x_chroma_shift, y_chroma_shift;
    unsigned char *planes[4];
    int stride[4];
    unsigned char *img_data;
    size_t img_bytes;
} vpx_image_t;

vpx_image_t *vpx_img_alloc(vpx_image_t *img, vpx_img_fmt_t fmt, unsigned int d_w,
                           unsigned int d_h, unsigned int align);
void vpx_img_free(vpx_image_t *img);

/* ---- codec plumbing ------------------------------------------------------- */
typedef int vpx_codec_err_t;
#define VPX_CODEC_OK 0
typedef struct vpx_codec_iface {
    const char *name;
    int is_encoder;
} vpx_codec_iface_t;
typedef const void *vpx_codec_iter_t;

typedef struct vpx_codec_ctx {
    const vpx_codec_iface_t *iface;
    void *priv;
} vpx_codec_ctx_t;

const vpx_codec_iface_t *vpx_codec_vp9_cx(void);
const vpx_codec_iface_t *vpx_codec_vp9_dx(void);
const char *vpx_codec_iface_name(const vpx_codec_iface_t *iface);
vpx_codec_err_t vpx_codec_destroy(vpx_codec_ctx_t *ctx);
vpx_codec_err_t vpx_codec_control(vpx_codec_ctx_t *ctx, int ctrl_id, int value);

/* ---- encoder -------------------------------------------------------------- */
typedef enum { VPX_BITS_8 = 8, VPX_BITS_10 = 10, VPX_BITS_12 = 12 } vpx_bit_depth_t;
enum vpx_enc_pass { VPX_RC_ONE_PASS, VPX_RC_FIRST_PASS, VPX_RC_LAST_PASS };
enum vpx_rc_mode { VPX_VBR, VPX_CBR, VPX_CQ, VPX_Q };
enum vpx_kf_mode { VPX_KF_FIXED, VPX_KF_AUTO, VPX_KF_DISABLED = 0 };
typedef struct vpx_rational {
    int num, den;
} vpx_rational_t;

typedef struct vpx_codec_enc_cfg {
    unsigned int g_usage, g_threads, g_profile, g_w, g_h;
    vpx_bit_depth_t g_bit_depth;
    unsigned int g_input_bit_depth;
    vpx_rational_t g_timebase;
    unsigned int g_error_resilient;
    enum vpx_enc_pass g_pass;
    unsigned int g_lag_in_frames;
    enum vpx_rc_mode rc_end_usage;
    unsigned int rc_target_bitrate, rc_min_quantizer, rc_max_quantizer;
    enum vpx_kf_mode kf_mode;
    unsigned int kf_min_dist, kf_max_dist;
} vpx_codec_enc_cfg_t;

#define VPX_CODEC_USE_HIGHBITDEPTH 0x40000
#define VPX_EFLAG_FORCE_KF 1
#define VPX_DL_GOOD_QUALITY 1000000
#define VPX_FRAME_IS_KEY 0x1
enum { VP9E_SET_LOSSLESS = 32, VP9E_SET_COLOR_SPACE = 46 };

enum vpx_codec_cx_pkt_kind { VPX_CODEC_CX_FRAME_PKT, VPX_CODEC_STATS_PKT };
typedef struct vpx_codec_cx_pkt {
    enum vpx_codec_cx_pkt_kind kind;
    union {
        struct {
            void *buf;
            size_t sz;
            unsigned int flags;
        } frame;
    } data;
} vpx_codec_cx_pkt_t;

vpx_codec_err_t vpx_codec_enc_config_default(const vpx_codec_iface_t *iface,
                                             vpx_codec_enc_cfg_t *cfg, unsigned int usage);
vpx_codec_err_t vpx_codec_enc_init(vpx_codec_ctx_t *ctx, const vpx_codec_iface_t *iface,
                                   const vpx_codec_enc_cfg_t *cfg, long flags);
vpx_codec_err_t vpx_codec_encode(vpx_codec_ctx_t *ctx, const vpx_image_t *img, long pts,
                                 unsigned long duration, long flags, unsigned long deadline);
const vpx_codec_cx_pkt_t *vpx_codec_get_cx_data(vpx_codec_ctx_t *ctx, vpx_codec_iter_t *iter);

/* ---- decoder -------------------------------------------------------------- */
typedef struct vpx_codec_dec_cfg {
    unsigned int threads, w, h;
} vpx_codec_dec_cfg_t;

vpx_codec_err_t vpx_codec_dec_init(vpx_codec_ctx_t *ctx, const vpx_codec_iface_t *iface,
                                   const vpx_codec_dec_cfg_t *cfg, long flags);
vpx_codec_err_t vpx_codec_decode(vpx_codec_ctx_t *ctx, const uint8_t *data,
                                 unsigned int data_sz, void *user_priv, long deadline);
vpx_image_t *vpx_codec_get_frame(vpx_codec_ctx_t *ctx, vpx_codec_iter_t *iter);

/* ---- loopback packet format (stub-private, used by the harness too) -------- */
typedef struct vpx_stub_pkt_hdr {
    uint32_t magic; /* 'LVPX' */
    uint32_t fmt, d_w, d_h;
    uint32_t row_bytes[3], rows[3];
} vpx_stub_pkt_hdr_t;
#define VPX_STUB_MAGIC 0x5850564cu
/* extra bytes of pitch the stub decoder adds per luma row (mimics libvpx border) */
#define VPX_STUB_DEC_BORDER 64

#ifdef __cplusplus
}
#endif
#endif

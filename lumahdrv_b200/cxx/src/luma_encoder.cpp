/*
 * luma_encoder.cpp -- LumaEncoder of the drop-in facade.
 *
 * Host glue with the behaviour of the reference (reference
 * src/luma_encoder.cpp:63-257: profile fix-up, quantizer set-up, the seven
 * Matroska attachments 430..436 that carry the quantizer to the decoder, VP9
 * configuration, frame submission, flush), written for this code base; the
 * per-pixel work the reference does in transformColorSpace + setVpxChannel
 * (src/luma_quantizer.cpp:269-373, src/luma_encoder.cpp:260-317) is one call
 * into the CUDA layer that fills m_rawFrame.planes[] directly.
 */
#include "luma_encoder.h"

#include "../../../include/lumacu.h"

#include <chrono>
#include "luma_exception.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

void throw_status(lumacu_ctx *ctx, int rc, const char *what)
{
    std::string msg = std::string(what) + ": " + lumacu_status_name(rc) + ": " + lumacu_last_error(ctx);
    throw LumaException(msg.c_str());
}

/* one attachment = id + raw bytes of a trivially copyable value (the reference memcpy's host
 * representations, so the wire format is "whatever the struct bytes are") */
template <typename T>
void attach(MkvInterface &w, unsigned int id, const T *data, size_t count, const char *label)
{
    /* MkvInterface keeps the pointer until writeAttachments(); give it storage that outlives us */
    T *copy = new T[count];
    memcpy(copy, data, count * sizeof(T));
    w.addAttachment(id, (const binary *)copy, (unsigned int)(count * sizeof(T)), label);
}

bool env_flag(const char *name)
{
    const char *e = getenv(name);
    return e && e[0] && e[0] != '0';
}

} // namespace

LumaEncoder::LumaEncoder()
    : m_frameCount(0), m_strict(env_flag("LUMA_STRICT_SIDE_EFFECT")), m_haveImage(false), m_haveCodec(false),
      m_lastMean(0.0), m_registered(NULL)
{
    memset(&m_codec, 0, sizeof(m_codec));
    memset(&m_rawFrame, 0, sizeof(m_rawFrame));
}

LumaEncoder::~LumaEncoder()
{
    if (m_registered)
        lumacu_host_unregister(m_registered);
    if (m_haveImage)
        vpx_img_free(&m_rawFrame);
    if (m_haveCodec && vpx_codec_destroy(&m_codec))
        fprintf(stderr, "Failed to destroy vpx codec\n");
}

bool LumaEncoder::initialize(const char *outputFile, const unsigned int w, const unsigned int h, bool verbose)
{
    LumaEncoderParams &p = m_params;
    LumaEncoderBase::initialize(outputFile, w, h, p.maxLum, p.minLum);

    /* profiles 0/1 are the 8-bit containers, 2/3 the high-bit-depth ones */
    if (p.bitDepth == 8 && p.profile > 1)
        p.profile -= 2;
    else if (p.bitDepth > 8 && p.profile < 2)
        p.profile += 2;

    m_quant.setQuantizer(p.ptf, p.ptfBitDepth, p.colorSpace, p.colorBitDepth, p.maxLum, p.minLum);

    /* quantizer metadata for the decoder: attachments 430..436.  434 carries getSize() = maxVal floats,
     * i.e. one entry short of the table, exactly like the reference -- the decoder rebuilds the table
     * first and then overlays these bytes. */
    attach(m_writer, 430, &p.ptfBitDepth, 1, "PTF bit depth");
    attach(m_writer, 431, &p.colorBitDepth, 1, "Color bit depth");
    attach(m_writer, 432, &p.ptf, 1, "PTF description");
    attach(m_writer, 433, &p.colorSpace, 1, "Color space");
    attach(m_writer, 434, m_quant.getMapping(), m_quant.getSize(), "PTF");
    attach(m_writer, 435, &p.preScaling, 1, "Scaling");
    const float range[2] = {p.maxLum, p.minLum};
    attach(m_writer, 436, range, 2, "Luminance range");
    m_writer.writeAttachments();
    m_writer.setFramerate(p.fps);
    m_writer.setVerbose(verbose);

    if (w == 0 || h == 0 || (w & 1u) || (h & 1u))
        throw LumaException("Invalid frame size");

    static const struct {
        vpx_img_fmt_t fmt;
        const char *err;
    } kFormats[4] = {{VPX_IMG_FMT_I420, "Failed to allocate 8 bit 420 image"},
                     {VPX_IMG_FMT_I444, "Failed to allocate 8 bit 444 image"},
                     {VPX_IMG_FMT_I42016, "Failed to allocate 16 bit 420 image"},
                     {VPX_IMG_FMT_I44416, "Failed to allocate 16 bit 444 image"}};
    if (p.profile > 3)
        throw LumaException("Invalid encoding profile");
    if (!vpx_img_alloc(&m_rawFrame, kFormats[p.profile].fmt, w, h, 32))
        throw LumaException(kFormats[p.profile].err);
    m_haveImage = true;
    /* page-lock the image libvpx allocated so the planes come back from the GPU by direct DMA */
    if (!env_flag("LUMA_NO_REGISTER_VPX") && m_rawFrame.planes[0] && m_rawFrame.planes[2]) {
        const unsigned chromaRows = (p.profile % 2 == 0) ? (h + 1) / 2 : h;
        unsigned char *end = m_rawFrame.planes[2] + (size_t)m_rawFrame.stride[2] * chromaRows;
        if (end > m_rawFrame.planes[0] && lumacu_host_register(m_rawFrame.planes[0], (size_t)(end - m_rawFrame.planes[0])) == 0)
            m_registered = m_rawFrame.planes[0];
    }

    const vpx_codec_iface_t *iface = vpx_codec_vp9_cx();
    vpx_codec_enc_cfg_t cfg;
    if (vpx_codec_enc_config_default(iface, &cfg, 0))
        throw LumaException("Failed to get default codec config");
    cfg.g_w = w;
    cfg.g_h = h;
    cfg.g_threads = 6;
    cfg.g_profile = p.profile;
    cfg.g_timebase.num = 1;
    cfg.g_timebase.den = 25;
    cfg.g_error_resilient = 0;
    cfg.g_pass = VPX_RC_ONE_PASS;
    cfg.g_lag_in_frames = 0;
    cfg.rc_end_usage = VPX_Q; /* constant quality: min == max quantizer */
    cfg.rc_min_quantizer = cfg.rc_max_quantizer = p.quantizerScale;
    cfg.rc_target_bitrate = p.bitrate;
    cfg.kf_mode = VPX_KF_AUTO;
    cfg.kf_max_dist = 25;
    unsigned int depth = 12;
    if (p.bitDepth == 8 || p.profile < 2)
        depth = 8;
    else if (p.bitDepth == 10)
        depth = 10;
    cfg.g_bit_depth = depth == 8 ? VPX_BITS_8 : depth == 10 ? VPX_BITS_10 : VPX_BITS_12;

    const bool ranged = p.ptf == LumaQuantizer::PTF_PQ || p.ptf == LumaQuantizer::PTF_LOG ||
                        p.ptf == LumaQuantizer::PTF_LINEAR;
    const char *rule = "-------------------------------------------------------------------\n";
    fprintf(stderr, "Encoding options:\n%s", rule);
    fprintf(stderr, "Transfer function (PTF):   %s\n", LumaQuantizer::name(p.ptf).c_str());
    fprintf(stderr, "Color space:               %s\n", LumaQuantizer::name(p.colorSpace).c_str());
    fprintf(stderr, "PTF bit depth:             %d\n", p.ptfBitDepth);
    fprintf(stderr, "Color bit depth:           %d\n", p.colorBitDepth);
    if (ranged)
        fprintf(stderr, "Encoding luminance range:  %.4f-%.2f\n", m_quant.getMinLum(), m_quant.getMaxLum());
    fprintf(stderr, "Encoding profile:          %d (4%s)\n", p.profile, (p.profile % 2 == 0) ? "22" : "44");
    fprintf(stderr, "Encoding bit depth:        %u\n", depth);
    fprintf(stderr, "Codec:                     %s\n", vpx_codec_iface_name(iface));
    fprintf(stderr, "Pixel transform:           CUDA (lumacu %d)\n", lumacu_version());
    fprintf(stderr, "Output:                    %s\n%s\n", outputFile, rule);

    if (vpx_codec_enc_init(&m_codec, iface, &cfg, p.profile < 2 ? 0 : VPX_CODEC_USE_HIGHBITDEPTH))
        throw LumaException("Failed to initialize vpxEncoder\n");
    m_haveCodec = true;
    /* BT.2020 signalling for third-party decoders when the planes really are YCbCr */
    if (p.colorSpace == LumaQuantizer::CS_YCBCR && vpx_codec_control(&m_codec, VP9E_SET_COLOR_SPACE, 5))
        fprintf(stderr, "Warning! Failed to set color space of encoder. Color primaries may not be recognized "
                        "during decoding.\n\n");
    if (p.lossLess && vpx_codec_control(&m_codec, VP9E_SET_LOSSLESS, 1))
        throw LumaException("Failed to use lossless mode\n");

    m_quant.device(); /* fail now, not at the first frame, if there is no usable GPU */
    m_initialized = true;
    return true;
}

/* "Is the input calibrated?" check of the reference (src/luma_encoder.cpp:276,294,314-316) */
void LumaEncoder::meanLuminanceCheck(double sum, size_t count)
{
    m_lastMean = count ? sum / (double)count : 0.0;
    if (m_lastMean <= 1.0)
        fprintf(stderr, "Warning! Mean luminance is %f cd/m2. Is the input calibrated to physical units?\n",
                m_lastMean);
}

void LumaEncoder::setChannels(LumaFrame *frame)
{
    if (!m_initialized)
        throw LumaException("LumaEncoder::setChannels: encoder is not initialized");
    if (!frame || !frame->buffer || frame->width != m_rawFrame.d_w || frame->height != m_rawFrame.d_h)
        throw LumaException("Invalid frame size");
    lumacu_ctx *ctx = m_quant.device();
    lumacu_frame_stats st;
    const int32_t strides[3] = {m_rawFrame.stride[0], m_rawFrame.stride[1], m_rawFrame.stride[2]};
    const int rc = lumacu_quantize_planes(ctx, frame->buffer, frame->width, frame->height, (int)m_params.profile,
                                          m_rawFrame.planes, strides, &st);
    if (rc != LUMACU_OK)
        throw_status(ctx, rc, "LumaEncoder::setChannels");
    meanLuminanceCheck(st.sum, (size_t)frame->width * frame->height);
}

bool LumaEncoder::encode(LumaFrame *frame)
{
    if (!m_initialized)
        throw LumaException("LumaEncoder::encode: encoder is not initialized");
    if (!frame || !frame->buffer || frame->width != m_rawFrame.d_w || frame->height != m_rawFrame.d_h)
        throw LumaException("Invalid frame size");
    lumacu_ctx *ctx = m_quant.device();
    lumacu_frame_stats st;
    const int32_t strides[3] = {m_rawFrame.stride[0], m_rawFrame.stride[1], m_rawFrame.stride[2]};
    /* LUMA_FACADE_TIMING=1: where a call's time goes -- the transform (this library) or run() (codec + container) */
    static const bool timing = getenv("LUMA_FACADE_TIMING") && getenv("LUMA_FACADE_TIMING")[0] != '0';
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    const int rc = lumacu_encode(ctx, frame->buffer, frame->width, frame->height, (int)m_params.profile,
                                 m_params.preScaling, m_rawFrame.planes, strides, m_strict ? 1 : 0, &st);
    if (rc != LUMACU_OK)
        throw_status(ctx, rc, "LumaEncoder::encode");
    meanLuminanceCheck(st.sum, (size_t)frame->width * frame->height);
    if (!timing)
        return run();
    const std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
    const bool ok = run();
    const std::chrono::steady_clock::time_point t2 = std::chrono::steady_clock::now();
    fprintf(stderr, "facade-timing encode: transform %.3f ms, run() (codec + container) %.3f ms\n",
            std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
    return ok;
}

bool LumaEncoder::run()
{
    const bool key = m_params.keyframeInterval > 0 && m_frameCount % m_params.keyframeInterval == 0;
    submit(&m_rawFrame, (int)m_frameCount++, key ? VPX_EFLAG_FORCE_KF : 0);
    return true;
}

void LumaEncoder::finish()
{
    while (submit(NULL, -1, 0)) { /* drain the codec */
    }
    LumaEncoderBase::finish();
}

/* hands one image (or a flush request) to VP9 and moves every finished packet into the container */
int LumaEncoder::submit(vpx_image_t *img, int frame_index, int flags)
{
    if (vpx_codec_encode(&m_codec, img, frame_index, 1, flags, VPX_DL_GOOD_QUALITY) != VPX_CODEC_OK)
        fprintf(stderr, "Failed to encode frame\n");
    int packets = 0;
    vpx_codec_iter_t it = NULL;
    for (const vpx_codec_cx_pkt_t *pkt; (pkt = vpx_codec_get_cx_data(&m_codec, &it)) != NULL;) {
        packets = 1;
        if (pkt->kind != VPX_CODEC_CX_FRAME_PKT)
            continue;
        m_writer.addFrame((const uint8_t *)pkt->data.frame.buf, (unsigned int)pkt->data.frame.sz,
                          (pkt->data.frame.flags & VPX_FRAME_IS_KEY) != 0);
    }
    return packets;
}

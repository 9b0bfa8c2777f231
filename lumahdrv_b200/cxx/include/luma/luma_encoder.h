/*
 * luma_encoder.h -- LumaEncoder of the drop-in C++ facade.
 *
 * Source-compatible with the reference header (reference
 * include/luma/luma_encoder.h:59-176): the parameter structs with the same
 * fields and defaults, LumaEncoderBase with its virtual interface, and
 * LumaEncoder with getParams / setParams / initialized / initialize / encode /
 * run / setChannels / finish.  The reference's drivers (lumaenc.cpp:184-265,
 * test/test_simple_enc.cpp) compile against it unmodified.
 *
 * What changed underneath: encode() no longer walks the frame four times on
 * one CPU core (transformColorSpace + 3 x setVpxChannel,
 * src/luma_quantizer.cpp:269-373, src/luma_encoder.cpp:260-317); it hands the
 * frame to one fused CUDA kernel (lumacu_encode, include/lumacu.h) that writes
 * straight into the vpx_image_t planes, then calls run() -- VP9 and Matroska
 * stay on the host exactly as in the reference.
 */
#ifndef LUMA_ENCODER_H
#define LUMA_ENCODER_H

#include "luma_frame.h"
#include "luma_quantizer.h"
#include "mkv_interface.h"

#include "vp8cx.h"
#include "vpx_encoder.h"

struct LumaEncoderParamsBase
{
    LumaEncoderParamsBase()
        : quantizerScale(2), ptfBitDepth(11), colorBitDepth(8), preScaling(1.0f), fps(25.0f),
          minLum(0.005f), maxLum(1e4f), ptf(LumaQuantizer::PTF_PQ), colorSpace(LumaQuantizer::CS_LUV)
    {
    }

    unsigned int quantizerScale, ptfBitDepth, colorBitDepth;
    float preScaling, fps, minLum, maxLum;
    LumaQuantizer::ptf_t ptf;
    LumaQuantizer::colorSpace_t colorSpace;
};

class LumaEncoderBase
{
public:
    LumaEncoderBase() : m_initialized(false) {}
    virtual ~LumaEncoderBase() {}

    /* opens the Matroska writer (reference luma_encoder.h:84-91) */
    virtual bool initialize(const char *outputFile, const unsigned int w, const unsigned int h,
                            const float ma, const float mi, bool verbose = 0)
    {
        (void)verbose;
        m_writer.openWrite(outputFile, w, h, ma, mi);
        return true;
    }

    virtual bool run() = 0;
    virtual void setChannels(LumaFrame *frame) = 0;
    virtual bool encode(LumaFrame *frame) = 0;
    virtual void finish() { m_writer.close(); }

    bool initialized() { return m_initialized; }

protected:
    bool m_initialized;
    LumaQuantizer m_quant;
    MkvInterface m_writer;
};

struct LumaEncoderParams : LumaEncoderParamsBase
{
    LumaEncoderParams() : bitrate(10000), profile(2), keyframeInterval(0), bitDepth(12), lossLess(false) {}

    unsigned int bitrate, profile, keyframeInterval, bitDepth;
    bool lossLess;
};

class LumaEncoder : public LumaEncoderBase
{
public:
    LumaEncoder();
    ~LumaEncoder();

    bool initialize(const char *outputFile, const unsigned int w, const unsigned int h, bool verbose = 0);
    bool run();
    /* quantises an already colour-transformed frame into the vpx planes (reference
     * src/luma_encoder.cpp:196-201): the second half of the reference's unfused encode() */
    void setChannels(LumaFrame *frame);
    /* colour transform + quantisation in one GPU pass, then run().  The reference transforms the
     * caller's frame in place as a side effect; no shipped driver reads it afterwards, so by default
     * the frame is left untouched (saves a 12 B/pixel device->host copy).  Set the environment
     * variable LUMA_STRICT_SIDE_EFFECT=1 (or call setStrictSideEffect) to reproduce the side effect. */
    bool encode(LumaFrame *frame);
    void finish();

    LumaEncoderParams getParams() { return m_params; }
    void setParams(LumaEncoderParams params) { m_params = params; }

    /* ---- additions ---- */
    void setStrictSideEffect(bool on) { m_strict = on; }
    void setDevice(int device) { m_quant.setDevice(device); } /* which GPU runs the transform (before initialize) */
    LumaQuantizer *getQuantizer() { return &m_quant; } /* e.g. for LumaQuantizer::broadcast */
    vpx_image_t *getRawFrame() { return &m_rawFrame; } /* the planes the last encode() produced */
    double lastMeanLuminance() const { return m_lastMean; }

private:
    int submit(vpx_image_t *img, int frame_index, int flags);
    void meanLuminanceCheck(double sum, size_t count);

    vpx_codec_ctx_t m_codec;
    vpx_image_t m_rawFrame;
    unsigned int m_frameCount;
    LumaEncoderParams m_params;
    bool m_strict, m_haveImage, m_haveCodec;
    double m_lastMean;
    void *m_registered; /* page-locked range of m_rawFrame (lumacu_host_register) */
};

#endif // LUMA_ENCODER_H

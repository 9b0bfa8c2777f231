/*
 * luma_decoder.h -- LumaDecoder of the drop-in C++ facade.
 *
 * Source-compatible with the reference header (reference
 * include/luma/luma_decoder.h:60-175): parameter structs, LumaDecoderBase
 * (seekToTime, getQuantizer, getReader, getFrame, initialized) and LumaDecoder
 * (constructor that opens the file, initialize, run, decode, getBuffer,
 * getParams / setParams).  lumadec.cpp:90-165, test/test_simple_dec.cpp and
 * lumaplay.cpp's use of getBuffer()/getQuantizer() compile unmodified.
 *
 * decode() = run() (Matroska + VP9 on the host, as before) followed by one
 * fused CUDA kernel (lumacu_decode) that replaces getVpxChannels +
 * transformColorSpace(frame,false,sc) (src/luma_decoder.cpp:205-240,
 * src/luma_quantizer.cpp:374-479) and fills the decoder-owned LumaFrame.
 */
#ifndef LUMA_DECODER_H
#define LUMA_DECODER_H

#include "luma_frame.h"
#include "luma_quantizer.h"

#include <vector>
#include "mkv_interface.h"

#include "vp8dx.h"
#include "vpx_decoder.h"

struct LumaDecoderParamsBase
{
    LumaDecoderParamsBase()
        : ptf(LumaQuantizer::PTF_PSI), colorSpace(LumaQuantizer::CS_LUV), preScaling(1.0f), minLum(0.005f),
          maxLum(1e4f)
    {
    }

    LumaQuantizer::ptf_t ptf;
    LumaQuantizer::colorSpace_t colorSpace;
    float preScaling, minLum, maxLum;
};

class LumaDecoderBase
{
public:
    LumaDecoderBase(const char *inputFile = NULL, bool verbose = 0) : m_initialized(false), m_input(inputFile)
    {
        (void)verbose;
    }
    virtual ~LumaDecoderBase() {}

    virtual bool initialize(const char *inputFile, bool verbose = 0) = 0;
    virtual bool run() = 0;
    void seekToTime(float tm, bool absolute = false) { m_reader.seekToTime(tm, absolute); }

    virtual LumaFrame *decode() = 0;
    LumaQuantizer *getQuantizer() { return &m_quant; }
    MkvInterface *getReader() { return &m_reader; }
    LumaFrame *getFrame() { return &m_frame; }
    bool initialized() { return m_initialized; }

protected:
    bool m_initialized;
    const char *m_input;
    LumaQuantizer m_quant;
    MkvInterface m_reader;
    LumaFrame m_frame;
};

struct LumaDecoderParams : LumaDecoderParamsBase
{
    LumaDecoderParams() : ptfBitDepth(11), colorBitDepth(8), highBitDepth(true), stride(NULL), profile(2)
    {
        for (int i = 0; i < 3; i++)
            width[i] = height[i] = 0;
    }

    unsigned int ptfBitDepth, colorBitDepth;
    bool highBitDepth;
    int *stride, profile, width[3], height[3];
};

class LumaDecoder : public LumaDecoderBase
{
public:
    LumaDecoder(const char *inputFile = NULL, bool verbose = 0);
    ~LumaDecoder();

    bool initialize(const char *inputFile, bool verbose = 0);
    bool run();
    LumaFrame *decode();

    unsigned char **getBuffer() { return m_vpxFrame->planes; }
    LumaDecoderParams getParams() { return m_params; }
    void setParams(LumaDecoderParams params) { m_params = params; }
    void setDevice(int device) { m_quant.setDevice(device); } /* addition: which GPU runs the transform */

private:
    void registerPlanes();

    vpx_codec_ctx_t m_codec;
    vpx_image_t *m_vpxFrame;
    LumaDecoderParams m_params;
    bool m_firstFrame, m_haveCodec;
    std::vector<unsigned char *> m_registered; /* page-locked libvpx frame buffers, least recently seen first */
    size_t m_registeredBytes;                  /* size of each registered range (changes with the frame geometry) */
};

#endif // LUMA_DECODER_H

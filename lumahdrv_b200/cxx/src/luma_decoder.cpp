/*
 * luma_decoder.cpp -- LumaDecoder of the drop-in facade.
 *
 * Host glue with the behaviour of the reference (reference
 * src/luma_decoder.cpp:46-202: read attachments 430..436, rebuild the
 * quantizer and overlay the stored table, VP9 decoder set-up, pull the first
 * frame to learn the plane geometry, frame pump); the per-pixel work of
 * getVpxChannels + transformColorSpace(false) (src/luma_decoder.cpp:205-240,
 * src/luma_quantizer.cpp:374-479) is one call into the CUDA layer reading
 * m_vpxFrame->planes[] with the decoder's own pitches.
 */
#include "luma_decoder.h"

#include "../../../include/lumacu.h"

#include <chrono>
#include "luma_exception.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace {

void throw_status(lumacu_ctx *ctx, int rc, const char *what)
{
    std::string msg = std::string(what) + ": " + lumacu_status_name(rc) + ": " + lumacu_last_error(ctx);
    throw LumaException(msg.c_str());
}

template <typename T>
T read_as(const binary *bytes, size_t index = 0)
{
    T v;
    memcpy(&v, bytes + index * sizeof(T), sizeof(T));
    return v;
}

} // namespace

LumaDecoder::LumaDecoder(const char *inputFile, bool verbose)
    : LumaDecoderBase(inputFile, verbose), m_vpxFrame(NULL), m_firstFrame(false), m_haveCodec(false), m_registeredBytes(0)
{
    memset(&m_codec, 0, sizeof(m_codec));
    if (inputFile != NULL)
        initialize(inputFile, verbose);
}

/* libvpx hands out frames from a small pool of buffers it owns; page-lock each buffer the first time it is seen
 * (direct DMA instead of a staged copy of the 3 B/px).  The buffers are libvpx's, so the registration is kept honest:
 * everything is unregistered as soon as the frame geometry changes (libvpx reallocates its pool then), and at most 16
 * ranges are held, the least recently seen one being dropped first -- a buffer libvpx no longer hands out does not
 * stay page-locked for long.  LUMA_NO_REGISTER_VPX=1 disables registration altogether (staged copies, no pinning of
 * memory this class does not own). */
void LumaDecoder::registerPlanes()
{
    static const bool off = getenv("LUMA_NO_REGISTER_VPX") && getenv("LUMA_NO_REGISTER_VPX")[0] != '0';
    if (off || !m_vpxFrame || !m_vpxFrame->planes[0] || !m_vpxFrame->planes[2])
        return;
    unsigned char *lo = m_vpxFrame->planes[0];
    const unsigned chromaRows = m_vpxFrame->y_chroma_shift ? (m_vpxFrame->d_h + 1) >> 1 : m_vpxFrame->d_h;
    unsigned char *hi = m_vpxFrame->planes[2] + (size_t)m_vpxFrame->stride[2] * chromaRows;
    const size_t bytes = hi > lo ? (size_t)(hi - lo) : 0;
    if (bytes != m_registeredBytes) { /* new geometry: libvpx has a new pool */
        for (size_t i = 0; i < m_registered.size(); i++)
            lumacu_host_unregister(m_registered[i]);
        m_registered.clear();
        m_registeredBytes = bytes;
    }
    for (size_t i = 0; i < m_registered.size(); i++)
        if (m_registered[i] == lo) { /* seen before: move to the back (most recently used) */
            m_registered.erase(m_registered.begin() + (long)i);
            m_registered.push_back(lo);
            return;
        }
    if (!bytes)
        return;
    if (m_registered.size() >= 16) {
        lumacu_host_unregister(m_registered.front());
        m_registered.erase(m_registered.begin());
    }
    if (lumacu_host_register(lo, bytes) == 0)
        m_registered.push_back(lo);
}

LumaDecoder::~LumaDecoder()
{
    for (size_t i = 0; i < m_registered.size(); i++)
        lumacu_host_unregister(m_registered[i]);
    if (m_haveCodec && vpx_codec_destroy(&m_codec))
        fprintf(stderr, "Failed to destroy vpx codec\n");
}

bool LumaDecoder::initialize(const char *inputFile, bool verbose)
{
    if (inputFile == NULL)
        return false;
    m_input = inputFile;
    m_reader.openRead(inputFile);
    m_reader.setVerbose(verbose);

    /* quantizer metadata written by LumaEncoder::initialize */
    enum { HAVE_PTF_BITS = 1, HAVE_COLOR_BITS = 2, HAVE_PTF = 4, HAVE_CS = 8, HAVE_TABLE = 16 };
    unsigned have = 0;
    const binary *table = NULL;
    unsigned int tableBytes = 0;
    binary *data = NULL;
    unsigned int id = 0, size = 0;
    for (unsigned int i = 0; m_reader.getAttachment(i, &data, id, size); i++) {
        switch (id) {
        case 430: m_params.ptfBitDepth = read_as<unsigned int>(data); have |= HAVE_PTF_BITS; break;
        case 431: m_params.colorBitDepth = read_as<unsigned int>(data); have |= HAVE_COLOR_BITS; break;
        case 432: m_params.ptf = read_as<LumaQuantizer::ptf_t>(data); have |= HAVE_PTF; break;
        case 433: m_params.colorSpace = read_as<LumaQuantizer::colorSpace_t>(data); have |= HAVE_CS; break;
        case 434: table = data; tableBytes = size; have |= HAVE_TABLE; break;
        case 435: m_params.preScaling = read_as<float>(data); break;
        case 436: m_params.maxLum = read_as<float>(data, 0); m_params.minLum = read_as<float>(data, 1); break;
        default: break;
        }
    }
    if (have != (HAVE_PTF_BITS | HAVE_COLOR_BITS | HAVE_PTF | HAVE_CS | HAVE_TABLE))
        throw LumaException(("Failed to locate Luma HDRv meta data in '" + std::string(inputFile) + "'").c_str());

    /* rebuild the table from the parameters, then overlay the stored entries (the file holds one entry
     * less than the table, see LumaEncoder::initialize); never write past the table */
    m_quant.setQuantizer(m_params.ptf, m_params.ptfBitDepth, m_params.colorSpace, m_params.colorBitDepth,
                         m_params.maxLum, m_params.minLum);
    const size_t room = ((size_t)m_quant.getSize() + 1) * sizeof(float);
    memcpy(const_cast<float *>(m_quant.getMapping()), table, tableBytes < room ? tableBytes : room);

    const vpx_codec_iface_t *iface = vpx_codec_vp9_dx();
    const bool ranged = m_params.ptf == LumaQuantizer::PTF_PQ || m_params.ptf == LumaQuantizer::PTF_LOG ||
                        m_params.ptf == LumaQuantizer::PTF_LINEAR;
    const char *rule = "-------------------------------------------------------------------\n";
    fprintf(stderr, "\nDecoding options:\n%s", rule);
    fprintf(stderr, "Transfer function (PTF):   %s\n", LumaQuantizer::name(m_params.ptf).c_str());
    if (ranged)
        fprintf(stderr, "Encoding luminance range:  %.4f-%.2f\n", m_quant.getMinLum(), m_quant.getMaxLum());
    fprintf(stderr, "Color space:               %s\n", LumaQuantizer::name(m_params.colorSpace).c_str());
    fprintf(stderr, "PTF bit depth:             %d\n", m_params.ptfBitDepth);
    fprintf(stderr, "Color bit depth:           %d\n", m_params.colorBitDepth);
    fprintf(stderr, "Codec:                     %s\n", vpx_codec_iface_name(iface));
    fprintf(stderr, "Pixel transform:           CUDA (lumacu %d)\n%s\n", lumacu_version(), rule);

    vpx_codec_dec_cfg_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.threads = 4;
    if (vpx_codec_dec_init(&m_codec, iface, &cfg, 0))
        fprintf(stderr, "Failed to initialize decoder.\n");
    else
        m_haveCodec = true;

    /* pull the first frame now: its format tells profile, pitches and plane sizes */
    m_firstFrame = false;
    m_initialized = true;
    if (!run()) {
        m_initialized = false;
        return false;
    }
    m_firstFrame = true; /* the next run() hands this frame out instead of decoding another */

    const vpx_image_t *f = m_vpxFrame;
    m_params.highBitDepth = (f->fmt & VPX_IMG_FMT_HIGHBITDEPTH) != 0;
    m_params.stride = m_vpxFrame->stride;
    const int hi = m_params.highBitDepth ? 2 : 0;
    m_params.profile = hi + (f->x_chroma_shift ? 0 : 1);
    for (int p = 0; p < 3; p++) {
        const unsigned xs = p ? f->x_chroma_shift : 0, ys = p ? f->y_chroma_shift : 0;
        m_params.width[p] = xs ? (int)((f->d_w + 1) >> xs) : (int)f->d_w;
        m_params.height[p] = ys ? (int)((f->d_h + 1) >> ys) : (int)f->d_h;
    }
    m_quant.device(); /* fail early without a GPU */
    return true;
}

bool LumaDecoder::run()
{
    if (!m_initialized && !initialize(m_input))
        return false;
    if (m_firstFrame) {
        m_firstFrame = false;
        return true;
    }
    m_vpxFrame = NULL;
    if (!m_reader.readFrame())
        return false; /* end of stream */
    unsigned int bytes = 0;
    const uint8_t *packet = m_reader.getFrame(bytes);
    if (vpx_codec_decode(&m_codec, packet, bytes, NULL, 0))
        throw LumaException("Failed to decode frame");
    vpx_codec_iter_t it = NULL;
    m_vpxFrame = vpx_codec_get_frame(&m_codec, &it);
    if (m_vpxFrame == NULL)
        throw LumaException("Failed to get decoded frame");
    return true;
}

LumaFrame *LumaDecoder::decode()
{
    static const bool timing = getenv("LUMA_FACADE_TIMING") && getenv("LUMA_FACADE_TIMING")[0] != '0';
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    if (!run())
        return NULL;
    const std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
    struct Report { /* printed when decode() returns, whichever way */
        bool on;
        std::chrono::steady_clock::time_point a, b;
        ~Report()
        {
            if (on)
                fprintf(stderr, "facade-timing decode: run() (container + codec) %.3f ms, transform %.3f ms\n",
                        std::chrono::duration<double, std::milli>(b - a).count(),
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - b).count());
        }
    } report = {timing, t0, t1};
    if (m_frame.width != m_vpxFrame->d_w || m_frame.height != m_vpxFrame->d_h || !m_frame.buffer) {
        m_frame.width = m_vpxFrame->d_w;
        m_frame.height = m_vpxFrame->d_h;
        m_frame.channels = 3;
        if (!m_frame.init())
            throw LumaException("Cannot allocate memory for the decoded frame");
    }
    lumacu_ctx *ctx = m_quant.device();
    const int32_t strides[3] = {m_vpxFrame->stride[0], m_vpxFrame->stride[1], m_vpxFrame->stride[2]};
    registerPlanes();
    const int hi = (m_vpxFrame->fmt & VPX_IMG_FMT_HIGHBITDEPTH) ? 2 : 0;
    const int profile = hi + (m_vpxFrame->x_chroma_shift ? 0 : 1);
    const int rc = lumacu_decode(ctx, m_vpxFrame->planes, strides, m_frame.width, m_frame.height, profile,
                                 m_params.preScaling, m_frame.buffer);
    if (rc != LUMACU_OK)
        throw_status(ctx, rc, "LumaDecoder::decode");
    return &m_frame;
}

/*
 * luma_quantizer.cpp -- LumaQuantizer of the drop-in facade, over the C ABI.
 *
 * Mirrors the behaviour of the reference class (reference
 * src/luma_quantizer.cpp:45-264 for construction, names, LUT set-up and the
 * scalar calls; :267-482 for transformColorSpace) without containing any of
 * its arithmetic: the LUT comes from lumacu_build_lut (host libm, reference
 * formulas), everything per-pixel from the CUDA kernels.
 */
#include "luma_quantizer.h"

#include "../../../include/lumacu.h"
#include "luma_exception.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

int env_device()
{
    const char *e = getenv("LUMA_CUDA_DEVICE");
    return e ? atoi(e) : 0;
}

void throw_status(lumacu_ctx *ctx, int rc, const char *what)
{
    std::string msg = std::string(what) + ": " + lumacu_status_name(rc) + ": " + lumacu_last_error(ctx);
    throw LumaException(msg.c_str());
}

} // namespace

LumaQuantizer::LumaQuantizer()
    : m_colorSpace(CS_LUV), m_Lmax(10000.0f), m_Lmin(0.005f), m_maxVal(0), m_maxValColor(0), m_bitdepth(0),
      m_bitdepthColor(0), m_device(-1), m_ctx(NULL), m_uploadedCs(CS_LUV), m_uploadedMaxValColor(0), m_uploadedLmax(0.0f)
{
}

LumaQuantizer::~LumaQuantizer()
{
    if (m_ctx)
        lumacu_destroy(m_ctx);
}

std::string LumaQuantizer::name(ptf_t ptf)
{
    switch (ptf) {
    case PTF_PQ: return "Perceptual quantizer (PQ, SMPTE ST 2084)";
    case PTF_LOG: return "Logarithmic";
    case PTF_JND_HDRVDP: return "JND HDR-VDP";
    case PTF_PSI: return "Perceptual - Ferwerda's t.v.i.";
    case PTF_LINEAR: return "Linear scaling";
    }
    return "Undefined";
}

std::string LumaQuantizer::name(colorSpace_t cs)
{
    switch (cs) {
    case CS_LUV: return "Lu'v'";
    case CS_RGB: return "RGB";
    case CS_YCBCR: return "YCbCr (ITU-R BT.2020)";
    case CS_XYZ: return "XYZ";
    }
    return "Undefined";
}

void LumaQuantizer::setQuantizer(ptf_t ptf, unsigned int bitdepth, colorSpace_t cs, unsigned int bitdepthC,
                                 float maxLum, float minLum)
{
    /* depths come straight from attachments 430/431 in the decoder (src/luma_decoder.cpp:79-100): refuse what the
     * 16-bit container cannot hold before sizing the table from them */
    if (bitdepth < 1 || bitdepth > 16 || bitdepthC < 1 || bitdepthC > 16)
        throw LumaException("LumaQuantizer::setQuantizer: bit depths must be in [1,16]");
    m_colorSpace = cs;
    m_bitdepth = bitdepth;
    m_bitdepthColor = bitdepthC;
    m_Lmax = maxLum;
    m_Lmin = minLum;
    /* (int)pow(2, b) - 1 like the reference (:180,:183); exact for the depths the CLI admits */
    m_maxVal = (unsigned int)((int)std::pow(2.0f, (float)bitdepth) - 1);
    m_maxValColor = (unsigned int)((int)std::pow(2.0f, (float)bitdepthC) - 1);
    m_mapping.assign((size_t)m_maxVal + 1, 0.0f);
    const int rc = lumacu_build_lut((int)ptf, bitdepth, maxLum, minLum, &m_mapping[0], m_mapping.size());
    if (rc != LUMACU_OK)
        throw_status(NULL, rc, "LumaQuantizer::setQuantizer");
}

/* create the context on first use and (re)upload the quantizer whenever the host-side state differs
 * from what the device holds -- this also catches writes through getMapping() */
void LumaQuantizer::sync() const
{
    if (m_mapping.empty())
        throw LumaException("LumaQuantizer: setQuantizer() has not been called");
    if (!m_ctx) {
        const int rc = lumacu_create(m_device >= 0 ? m_device : env_device(), &m_ctx);
        if (rc != LUMACU_OK) {
            m_ctx = NULL;
            throw_status(NULL, rc, "LumaQuantizer: no usable CUDA device (this build has no CPU path)");
        }
    }
    const bool same = m_uploaded.size() == m_mapping.size() && m_uploadedCs == m_colorSpace &&
                      m_uploadedMaxValColor == m_maxValColor && m_uploadedLmax == m_Lmax &&
                      memcmp(&m_uploaded[0], &m_mapping[0], m_mapping.size() * sizeof(float)) == 0;
    if (same)
        return;
    const int rc = lumacu_set_quantizer(m_ctx, &m_mapping[0], (uint32_t)m_mapping.size(), m_maxValColor,
                                        (int)m_colorSpace, m_Lmax);
    if (rc != LUMACU_OK)
        throw_status(m_ctx, rc, "LumaQuantizer: lumacu_set_quantizer");
    m_uploaded = m_mapping;
    m_uploadedCs = m_colorSpace;
    m_uploadedMaxValColor = m_maxValColor;
    m_uploadedLmax = m_Lmax;
}

void LumaQuantizer::setDevice(int device)
{
    if (m_ctx && lumacu_device(m_ctx) != device) { /* move: drop the old context, the next use re-uploads */
        lumacu_destroy(m_ctx);
        m_ctx = NULL;
        m_uploaded.clear();
    }
    m_device = device;
}

void LumaQuantizer::broadcast(LumaQuantizer *const q[], int n, int root)
{
    if (!q || n < 1 || root < 0 || root >= n || !q[root])
        throw LumaException("LumaQuantizer::broadcast: bad arguments");
    LumaQuantizer *src = q[root];
    src->sync(); /* context exists, tables uploaded */
    std::vector<lumacu_ctx *> ctxs((size_t)n, (lumacu_ctx *)NULL);
    for (int i = 0; i < n; i++) {
        if (!q[i])
            throw LumaException("LumaQuantizer::broadcast: NULL object");
        if (!q[i]->m_ctx) {
            const int rc = lumacu_create(q[i]->m_device >= 0 ? q[i]->m_device : env_device(), &q[i]->m_ctx);
            if (rc != LUMACU_OK) {
                q[i]->m_ctx = NULL;
                throw_status(NULL, rc, "LumaQuantizer: no usable CUDA device (this build has no CPU path)");
            }
        }
        ctxs[(size_t)i] = q[i]->m_ctx;
    }
    const int rc = lumacu_broadcast_quantizer(&ctxs[0], n, root);
    if (rc != LUMACU_OK)
        throw_status(src->m_ctx, rc, "LumaQuantizer::broadcast");
    for (int i = 0; i < n; i++) {
        LumaQuantizer *d = q[i];
        if (d == src)
            continue;
        d->m_colorSpace = src->m_colorSpace;
        d->m_mapping = src->m_mapping;
        d->m_Lmax = src->m_Lmax;
        d->m_Lmin = src->m_Lmin;
        d->m_maxVal = src->m_maxVal;
        d->m_maxValColor = src->m_maxValColor;
        d->m_bitdepth = src->m_bitdepth;
        d->m_bitdepthColor = src->m_bitdepthColor;
        d->m_uploaded = src->m_uploaded; /* what the device holds now: no re-derivation at the next use */
        d->m_uploadedCs = src->m_uploadedCs;
        d->m_uploadedMaxValColor = src->m_uploadedMaxValColor;
        d->m_uploadedLmax = src->m_uploadedLmax;
    }
}

lumacu_ctx *LumaQuantizer::device() const
{
    sync();
    return m_ctx;
}

void LumaQuantizer::quantizeN(const float *in, float *out, size_t n, unsigned int ch) const
{
    const int rc = lumacu_quantize(device(), in, out, n, ch);
    if (rc != LUMACU_OK)
        throw_status(m_ctx, rc, "LumaQuantizer::quantize");
}

void LumaQuantizer::dequantizeN(const float *in, float *out, size_t n, unsigned int ch) const
{
    const int rc = lumacu_dequantize(device(), in, out, n, ch);
    if (rc != LUMACU_OK)
        throw_status(m_ctx, rc, "LumaQuantizer::dequantize");
}

float LumaQuantizer::quantize(const float val, const unsigned int ch) const
{
    float out = 0.0f;
    quantizeN(&val, &out, 1, ch);
    return out;
}

float LumaQuantizer::dequantize(const float val, const unsigned int ch) const
{
    float out = 0.0f;
    dequantizeN(&val, &out, 1, ch);
    return out;
}

bool LumaQuantizer::transformColorSpace(LumaFrame *frame, bool toCs, float sc)
{
    if (!frame || !frame->buffer)
        return false;
    if ((int)m_colorSpace < 0 || (int)m_colorSpace > 3) {
        /* reference :368-371 / :474-477 */
        fprintf(stderr, "Error: color space not recognized\n");
        return false;
    }
    const int rc = lumacu_transform_color_space(device(), frame->buffer, frame->width, frame->height, toCs ? 1 : 0, sc);
    if (rc != LUMACU_OK)
        throw_status(m_ctx, rc, "LumaQuantizer::transformColorSpace");
    return true;
}

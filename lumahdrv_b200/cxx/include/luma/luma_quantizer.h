/*
 * luma_quantizer.h -- LumaQuantizer of the drop-in C++ facade.
 *
 * Same public interface as the reference class (reference
 * include/luma/luma_quantizer.h:89-111): enums ptf_t / colorSpace_t (their
 * ORDER is wire format, memcpy'd into Matroska attachments 432/433 by
 * src/luma_encoder.cpp:86-92), name(), setQuantizer(), quantize(),
 * dequantize(), transformColorSpace(), getMapping(), getSize(), getMaxLum(),
 * getMinLum().  The arithmetic runs on the GPU through the C ABI in
 * include/lumacu.h; there is no CPU implementation behind this class.
 */
#ifndef LUMA_QUANTIZER_H
#define LUMA_QUANTIZER_H

#include "luma_frame.h"

#include <cstddef>
#include <string>
#include <vector>

struct lumacu_ctx;

class LumaQuantizer
{
public:
    enum ptf_t { PTF_PSI, PTF_PQ, PTF_LOG, PTF_JND_HDRVDP, PTF_LINEAR };
    enum colorSpace_t { CS_LUV, CS_RGB, CS_YCBCR, CS_XYZ };

    LumaQuantizer();
    ~LumaQuantizer();

    static std::string name(ptf_t ptf);
    static std::string name(colorSpace_t cs);

    /* reference src/luma_quantizer.cpp:172-212; the LUT is built on the host with the host libm */
    void setQuantizer(ptf_t ptf, unsigned int bitdepth, colorSpace_t cs, unsigned int bitdepthC,
                      float maxLum, float minLum);

    /* reference src/luma_quantizer.cpp:215-264.  One value per call means one kernel launch per
     * call: these exist for API completeness; bulk callers use quantizeN / dequantizeN. */
    float quantize(const float val, const unsigned int ch) const;
    float dequantize(const float val, const unsigned int ch) const;
    void quantizeN(const float *in, float *out, size_t n, unsigned int ch) const;
    void dequantizeN(const float *in, float *out, size_t n, unsigned int ch) const;

    /* reference src/luma_quantizer.cpp:267-482: in place, false + stderr on an unknown colour space */
    bool transformColorSpace(LumaFrame *frame, bool toCs, float sc);

    /* The reference hands out its internal table and LumaDecoder::initialize writes through the
     * pointer (src/luma_decoder.cpp:122).  Same here: the device copy is refreshed lazily whenever
     * the host table no longer matches what was uploaded. */
    const float *getMapping() { return m_mapping.empty() ? NULL : &m_mapping[0]; }
    unsigned int getSize() { return m_maxVal; } /* maxVal, not the entry count (reference :109) */
    float getMaxLum() { return m_Lmax; }
    float getMinLum() { return m_Lmin; }

    /* ---- additions used by LumaEncoder / LumaDecoder of this facade ---- */
    lumacu_ctx *device() const;   /* context with the current quantizer uploaded; throws LumaException */
    /* CUDA device this object computes on (default: environment variable LUMA_CUDA_DEVICE, else 0).  Must be called
     * before the first use; one object per GPU (on its own host thread) is how a process drives several GPUs. */
    void setDevice(int device);
    colorSpace_t colorSpace() const { return m_colorSpace; }
    /* One process, several GPUs: q[root]'s quantizer (host table AND the device search tables derived from it) is
     * copied to every other object, device to device (lumacu_broadcast_quantizer), instead of each object
     * re-deriving its own.  Call after setQuantizer / LumaDecoder::initialize on the root. */
    static void broadcast(LumaQuantizer *const q[], int n, int root);

private:
    LumaQuantizer(const LumaQuantizer &);
    LumaQuantizer &operator=(const LumaQuantizer &);
    void sync() const;

    colorSpace_t m_colorSpace;
    std::vector<float> m_mapping;
    float m_Lmax, m_Lmin;
    unsigned int m_maxVal, m_maxValColor, m_bitdepth, m_bitdepthColor;

    int m_device; /* -1 = take LUMA_CUDA_DEVICE */
    mutable lumacu_ctx *m_ctx;
    mutable std::vector<float> m_uploaded; /* what the device currently holds */
    mutable colorSpace_t m_uploadedCs;
    mutable unsigned int m_uploadedMaxValColor;
    mutable float m_uploadedLmax;
};

#endif // LUMA_QUANTIZER_H

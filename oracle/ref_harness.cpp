/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE ONLY: C entry points over the
 * UNMODIFIED reference classes, compiled in place from /root/reference
 * (src/luma_quantizer.cpp, src/luma_encoder.cpp, src/luma_decoder.cpp) into
 * oracle/_ref/libluma_ref.so by oracle/Makefile.  libvpx / libmatroska are
 * replaced by the loopback test doubles under oracle/ref_stubs/, so the real
 * LumaEncoder::encode -> transformColorSpace + private setVpxChannel and the
 * real LumaDecoder::decode -> private getVpxChannels + transformColorSpace run
 * exactly as in the reference's lossLess=1 pipeline.
 *
 * Used to (1) validate oracle/luma_oracle.c, (2) generate tests/golden/, and
 * (3) serve as bench.py's CPU baseline (cpu_baseline.kind = "reference").
 * Never linked by the product path.
 */
#include "luma_decoder.h"
#include "luma_encoder.h"
#include "luma_exception.h"

#include <fcntl.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <new>

namespace {

int g_quiet = 1;

/* the reference prints banners/warnings on stderr; mute them while g_quiet */
struct StderrMute {
    int saved;
    StderrMute() : saved(-1)
    {
        if (!g_quiet)
            return;
        fflush(stderr);
        saved = dup(2);
        int nul = open("/dev/null", O_WRONLY);
        if (nul >= 0) {
            dup2(nul, 2);
            close(nul);
        }
    }
    ~StderrMute()
    {
        if (saved >= 0) {
            fflush(stderr);
            dup2(saved, 2);
            close(saved);
        }
    }
};

/* Borrow caller memory as a LumaFrame without copying (fields are public,
 * include/luma/luma_frame.h:88-89); detach before the destructor runs. */
struct BorrowedFrame {
    LumaFrame f;
    BorrowedFrame(float *data, unsigned w, unsigned h)
    {
        f.width = w;
        f.height = h;
        f.channels = 3;
        f.buffer = data;
    }
    ~BorrowedFrame() { f.buffer = NULL; }
};

} // namespace

extern "C" {

typedef struct lref_params {
    int ptf, color_space;
    unsigned ptf_bits, color_bits;
    unsigned profile, bit_depth;
    float pre_scaling, max_lum, min_lum;
} lref_params;

void lref_set_quiet(int q) { g_quiet = q; }

/* ---------------- LumaQuantizer (include/luma/luma_quantizer.h:89-126) ------ */
void *lref_quant_new(void) { return new (std::nothrow) LumaQuantizer(); }
void lref_quant_free(void *q) { delete (LumaQuantizer *)q; }
void lref_quant_set(void *q, int ptf, unsigned bits, int cs, unsigned cbits, float max_lum,
                    float min_lum)
{
    ((LumaQuantizer *)q)
        ->setQuantizer((LumaQuantizer::ptf_t)ptf, bits, (LumaQuantizer::colorSpace_t)cs, cbits,
                       max_lum, min_lum);
}
unsigned lref_quant_size(void *q) { return ((LumaQuantizer *)q)->getSize(); }
const float *lref_quant_mapping(void *q) { return ((LumaQuantizer *)q)->getMapping(); }
float lref_quant_quantize(void *q, float v, unsigned ch) { return ((LumaQuantizer *)q)->quantize(v, ch); }
float lref_quant_dequantize(void *q, float v, unsigned ch)
{
    return ((LumaQuantizer *)q)->dequantize(v, ch);
}
void lref_quant_quantize_n(void *q, const float *in, float *out, size_t n, unsigned ch)
{
    const LumaQuantizer *lq = (const LumaQuantizer *)q;
    for (size_t i = 0; i < n; i++)
        out[i] = lq->quantize(in[i], ch);
}
void lref_quant_dequantize_n(void *q, const float *in, float *out, size_t n, unsigned ch)
{
    const LumaQuantizer *lq = (const LumaQuantizer *)q;
    for (size_t i = 0; i < n; i++)
        out[i] = lq->dequantize(in[i], ch);
}
int lref_quant_transform(void *q, float *frame, unsigned w, unsigned h, int to_cs, float sc)
{
    StderrMute mute;
    BorrowedFrame bf(frame, w, h);
    return ((LumaQuantizer *)q)->transformColorSpace(&bf.f, to_cs != 0, sc) ? 1 : 0;
}

/* ---------------- LumaEncoder stream (real private loops) ------------------- */
struct lref_encoder {
    LumaEncoder enc;
    std::string name;
    unsigned w, h;
};

static LumaEncoderParams to_enc_params(const lref_params *p)
{
    LumaEncoderParams ep;
    ep.ptf = (LumaQuantizer::ptf_t)p->ptf;
    ep.colorSpace = (LumaQuantizer::colorSpace_t)p->color_space;
    ep.ptfBitDepth = p->ptf_bits;
    ep.colorBitDepth = p->color_bits;
    ep.profile = p->profile;
    ep.bitDepth = p->bit_depth;
    ep.preScaling = p->pre_scaling;
    ep.maxLum = p->max_lum;
    ep.minLum = p->min_lum;
    ep.lossLess = 1;
    return ep;
}

/* returns NULL and fills err (if given) when the reference throws */
void *lref_encoder_new(const lref_params *p, unsigned w, unsigned h, char *err, size_t errlen)
{
    StderrMute mute;
    lref_encoder *e = new (std::nothrow) lref_encoder();
    if (!e)
        return NULL;
    char nm[64];
    snprintf(nm, sizeof(nm), "mem:enc:%p", (void *)e);
    e->name = nm;
    e->w = w;
    e->h = h;
    try {
        e->enc.setParams(to_enc_params(p));
        e->enc.initialize(e->name.c_str(), w, h);
    } catch (std::exception &ex) {
        if (err && errlen)
            snprintf(err, errlen, "%s", ex.what());
        mkv_stub_erase(e->name.c_str());
        delete e;
        return NULL;
    }
    return e;
}

void lref_encoder_free(void *h)
{
    lref_encoder *e = (lref_encoder *)h;
    if (!e)
        return;
    mkv_stub_erase(e->name.c_str());
    delete e;
}

/* the profile after the reference's bit-depth fix-up (src/luma_encoder.cpp:69-72) */
int lref_encoder_profile(void *h) { return (int)((lref_encoder *)h)->enc.getParams().profile; }

/* LumaEncoder::encode(frame) (include/luma/luma_encoder.h:142-148): mutates
 * `frame` in place like the reference; the integer planes the encoder handed
 * to the codec are copied (payload rows only) into planes[] with strides[]. */
int lref_encoder_encode(void *h, float *frame, uint8_t *const planes[3], const int strides[3])
{
    StderrMute mute;
    lref_encoder *e = (lref_encoder *)h;
    BorrowedFrame bf(frame, e->w, e->h);
    try {
        if (!e->enc.encode(&bf.f))
            return -1;
    } catch (...) {
        return -2;
    }
    MkvStubFile *file = mkv_stub_find(e->name.c_str());
    size_t n = mkv_stub_frame_count(e->name.c_str());
    if (!file || !n)
        return -3;
    const std::vector<uint8> *pkt = mkv_stub_frame(e->name.c_str(), n - 1);
    vpx_stub_pkt_hdr_t hdr;
    memcpy(&hdr, pkt->data(), sizeof(hdr));
    const uint8 *src = pkt->data() + sizeof(hdr);
    for (int p = 0; p < 3; p++)
        for (uint32_t y = 0; y < hdr.rows[p]; y++, src += hdr.row_bytes[p])
            if (planes && planes[p])
                memcpy(planes[p] + (size_t)y * strides[p], src, hdr.row_bytes[p]);
    /* keep the container from growing: drop the payload we just consumed */
    const_cast<std::vector<uint8> *>(pkt)->clear();
    const_cast<std::vector<uint8> *>(pkt)->shrink_to_fit();
    return 0;
}

/* ---------------- LumaDecoder stream ---------------------------------------- */
struct lref_decoder {
    LumaDecoder *dec;
    std::string name;
    lref_params params;
    unsigned w, h;
    int fmt;
};

void *lref_decoder_new(const lref_params *p, unsigned w, unsigned h, char *err, size_t errlen)
{
    StderrMute mute;
    lref_decoder *d = new (std::nothrow) lref_decoder();
    if (!d)
        return NULL;
    char nm[64];
    snprintf(nm, sizeof(nm), "mem:dec:%p", (void *)d);
    d->name = nm;
    d->dec = NULL;
    d->params = *p;
    d->w = w;
    d->h = h;
    /* let the reference encoder write the metadata attachments 430..436 */
    try {
        LumaEncoder enc;
        enc.setParams(to_enc_params(p));
        enc.initialize(d->name.c_str(), w, h);
        const unsigned prof = enc.getParams().profile;
        d->fmt = prof == 0 ? VPX_IMG_FMT_I420
                           : prof == 1 ? VPX_IMG_FMT_I444 : prof == 2 ? VPX_IMG_FMT_I42016 : VPX_IMG_FMT_I44416;
    } catch (std::exception &ex) {
        if (err && errlen)
            snprintf(err, errlen, "%s", ex.what());
        mkv_stub_erase(d->name.c_str());
        delete d;
        return NULL;
    }
    return d;
}

void lref_decoder_free(void *h)
{
    lref_decoder *d = (lref_decoder *)h;
    if (!d)
        return;
    delete d->dec;
    mkv_stub_erase(d->name.c_str());
    delete d;
}

/* LumaDecoder::decode() (include/luma/luma_decoder.h:143-161) on one frame
 * whose integer planes are given; out receives 3*w*h floats.  dec_strides (if
 * non-NULL) receives the pitches the decoder actually read with. */
int lref_decoder_decode(void *h, const uint8_t *const planes[3], const int strides[3], float *out,
                        int dec_strides[3])
{
    StderrMute mute;
    lref_decoder *d = (lref_decoder *)h;
    vpx_stub_pkt_hdr_t hdr;
    memset(&hdr, 0, sizeof(hdr));
    hdr.magic = VPX_STUB_MAGIC;
    hdr.fmt = (uint32_t)d->fmt;
    hdr.d_w = d->w;
    hdr.d_h = d->h;
    const int sub = ((d->fmt & 0xff) == 2);
    const int bytes = (d->fmt & VPX_IMG_FMT_HIGHBITDEPTH) ? 2 : 1;
    size_t total = sizeof(hdr);
    for (int p = 0; p < 3; p++) {
        unsigned pw = (p && sub) ? (d->w + 1) >> 1 : d->w;
        unsigned ph = (p && sub) ? (d->h + 1) >> 1 : d->h;
        hdr.row_bytes[p] = pw * bytes;
        hdr.rows[p] = ph;
        total += (size_t)hdr.row_bytes[p] * ph;
    }
    std::vector<uint8> pkt(total);
    memcpy(pkt.data(), &hdr, sizeof(hdr));
    uint8 *dst = pkt.data() + sizeof(hdr);
    for (int p = 0; p < 3; p++)
        for (uint32_t y = 0; y < hdr.rows[p]; y++, dst += hdr.row_bytes[p])
            memcpy(dst, planes[p] + (size_t)y * strides[p], hdr.row_bytes[p]);
    mkv_stub_append_frame(d->name.c_str(), pkt.data(), pkt.size());

    LumaFrame *res = NULL;
    try {
        if (!d->dec)
            d->dec = new LumaDecoder(d->name.c_str()); /* initialize() consumes packet 0 */
        res = d->dec->decode();
    } catch (...) {
        return -2;
    }
    if (!res)
        return -1;
    if (dec_strides)
        for (int p = 0; p < 3; p++)
            dec_strides[p] = d->dec->getParams().stride[p];
    memcpy(out, res->buffer, sizeof(float) * 3 * (size_t)d->w * d->h);
    /* drop consumed payloads */
    MkvStubFile *file = mkv_stub_find(d->name.c_str());
    (void)file;
    size_t n = mkv_stub_frame_count(d->name.c_str());
    if (n) {
        std::vector<uint8> *last = const_cast<std::vector<uint8> *>(mkv_stub_frame(d->name.c_str(), n - 1));
        last->clear();
        last->shrink_to_fit();
    }
    return 0;
}

} /* extern "C" */

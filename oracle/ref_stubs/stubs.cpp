/*
 * stubs.cpp -- TEST DOUBLES for the oracle/_ref build (test infrastructure).
 * Implements the loopback "codec" declared in vpx_stub.h and the in-memory
 * container declared in mkv_interface_stub.h.  Written from scratch.
 */
#include "mkv_interface_stub.h"
#include "vpx_stub.h"

#include <cstdlib>
#include <map>
#include <mutex>

/* =========================== image allocation ============================== */
/* Pitch rule of libvpx 1.6.1 vpx_img_alloc as the reference relies on it
 * (src/luma_encoder.cpp:121-128, stride_align 32): the sample width is rounded
 * up to the chroma grid, then to `align` samples; 16-bit formats double the
 * byte pitch; subsampled chroma planes get half the luma pitch. */
extern "C" vpx_image_t *vpx_img_alloc(vpx_image_t *img, vpx_img_fmt_t fmt, unsigned int d_w,
                                      unsigned int d_h, unsigned int align)
{
    if (!img || !d_w || !d_h)
        return NULL;
    const int sub = ((fmt & 0xff) == 2);
    const int bytes = (fmt & VPX_IMG_FMT_HIGHBITDEPTH) ? 2 : 1;
    if (!align)
        align = 1;
    memset(img, 0, sizeof(*img));
    img->fmt = fmt;
    img->d_w = d_w;
    img->d_h = d_h;
    img->x_chroma_shift = img->y_chroma_shift = sub ? 1 : 0;
    img->w = sub ? ((d_w + 1) & ~1u) : d_w;
    img->h = sub ? ((d_h + 1) & ~1u) : d_h;
    img->bit_depth = bytes * 8;
    unsigned s = (img->w + align - 1) & ~(align - 1);
    img->stride[0] = (int)(s * bytes);
    img->stride[1] = img->stride[2] = sub ? img->stride[0] >> 1 : img->stride[0];
    const size_t ysz = (size_t)img->stride[0] * img->h;
    const size_t csz = (size_t)img->stride[1] * (sub ? img->h >> 1 : img->h);
    img->img_bytes = ysz + 2 * csz;
    img->img_data = (unsigned char *)calloc(img->img_bytes, 1);
    if (!img->img_data)
        return NULL;
    img->planes[0] = img->img_data;
    img->planes[1] = img->planes[0] + ysz;
    img->planes[2] = img->planes[1] + csz;
    return img;
}

extern "C" void vpx_img_free(vpx_image_t *img)
{
    if (img && img->img_data) {
        free(img->img_data);
        img->img_data = NULL;
    }
}

/* ============================ loopback codec ================================ */
namespace {
struct EncState {
    std::vector<uint8_t> pkt_bytes;
    vpx_codec_cx_pkt_t pkt;
    bool pending;
};
struct DecState {
    vpx_image_t img;
    std::vector<uint8_t> storage;
    bool have;
};
const vpx_codec_iface_t kEncIface = {"luma-oracle lossless loopback (encoder stand-in)", 1};
const vpx_codec_iface_t kDecIface = {"luma-oracle lossless loopback (decoder stand-in)", 0};

void plane_geometry(const vpx_image_t *img, uint32_t row_bytes[3], uint32_t rows[3])
{
    const int bytes = (img->fmt & VPX_IMG_FMT_HIGHBITDEPTH) ? 2 : 1;
    for (int p = 0; p < 3; p++) {
        unsigned w = img->d_w, h = img->d_h;
        if (p && img->x_chroma_shift)
            w = (w + 1) >> img->x_chroma_shift;
        if (p && img->y_chroma_shift)
            h = (h + 1) >> img->y_chroma_shift;
        row_bytes[p] = w * bytes;
        rows[p] = h;
    }
}
} // namespace

extern "C" const vpx_codec_iface_t *vpx_codec_vp9_cx(void) { return &kEncIface; }
extern "C" const vpx_codec_iface_t *vpx_codec_vp9_dx(void) { return &kDecIface; }
extern "C" const char *vpx_codec_iface_name(const vpx_codec_iface_t *iface)
{
    return iface ? iface->name : "?";
}

extern "C" vpx_codec_err_t vpx_codec_enc_config_default(const vpx_codec_iface_t *,
                                                        vpx_codec_enc_cfg_t *cfg, unsigned int)
{
    memset(cfg, 0, sizeof(*cfg));
    return VPX_CODEC_OK;
}

extern "C" vpx_codec_err_t vpx_codec_enc_init(vpx_codec_ctx_t *ctx, const vpx_codec_iface_t *iface,
                                              const vpx_codec_enc_cfg_t *, long)
{
    ctx->iface = iface;
    EncState *s = new EncState();
    s->pending = false;
    ctx->priv = s;
    return VPX_CODEC_OK;
}

extern "C" vpx_codec_err_t vpx_codec_dec_init(vpx_codec_ctx_t *ctx, const vpx_codec_iface_t *iface,
                                              const vpx_codec_dec_cfg_t *, long)
{
    ctx->iface = iface;
    DecState *s = new DecState();
    s->have = false;
    ctx->priv = s;
    return VPX_CODEC_OK;
}

extern "C" vpx_codec_err_t vpx_codec_destroy(vpx_codec_ctx_t *ctx)
{
    if (!ctx || !ctx->iface)
        return 1;
    if (ctx->iface->is_encoder)
        delete (EncState *)ctx->priv;
    else
        delete (DecState *)ctx->priv;
    ctx->priv = NULL;
    return VPX_CODEC_OK;
}

extern "C" vpx_codec_err_t vpx_codec_control(vpx_codec_ctx_t *, int, int) { return VPX_CODEC_OK; }

extern "C" vpx_codec_err_t vpx_codec_encode(vpx_codec_ctx_t *ctx, const vpx_image_t *img, long,
                                            unsigned long, long flags, unsigned long)
{
    EncState *s = (EncState *)ctx->priv;
    s->pending = false;
    if (!img) /* flush: nothing buffered in a zero-lag loopback */
        return VPX_CODEC_OK;
    vpx_stub_pkt_hdr_t hdr;
    memset(&hdr, 0, sizeof(hdr));
    hdr.magic = VPX_STUB_MAGIC;
    hdr.fmt = (uint32_t)img->fmt;
    hdr.d_w = img->d_w;
    hdr.d_h = img->d_h;
    plane_geometry(img, hdr.row_bytes, hdr.rows);
    size_t total = sizeof(hdr);
    for (int p = 0; p < 3; p++)
        total += (size_t)hdr.row_bytes[p] * hdr.rows[p];
    s->pkt_bytes.resize(total);
    memcpy(s->pkt_bytes.data(), &hdr, sizeof(hdr));
    uint8_t *dst = s->pkt_bytes.data() + sizeof(hdr);
    for (int p = 0; p < 3; p++)
        for (uint32_t y = 0; y < hdr.rows[p]; y++, dst += hdr.row_bytes[p])
            memcpy(dst, img->planes[p] + (size_t)y * img->stride[p], hdr.row_bytes[p]);
    s->pkt.kind = VPX_CODEC_CX_FRAME_PKT;
    s->pkt.data.frame.buf = s->pkt_bytes.data();
    s->pkt.data.frame.sz = total;
    s->pkt.data.frame.flags = (flags & VPX_EFLAG_FORCE_KF) ? VPX_FRAME_IS_KEY : 0;
    s->pending = true;
    return VPX_CODEC_OK;
}

extern "C" const vpx_codec_cx_pkt_t *vpx_codec_get_cx_data(vpx_codec_ctx_t *ctx,
                                                            vpx_codec_iter_t *iter)
{
    EncState *s = (EncState *)ctx->priv;
    if (*iter || !s->pending)
        return NULL;
    *iter = s;
    s->pending = false;
    return &s->pkt;
}

extern "C" vpx_codec_err_t vpx_codec_decode(vpx_codec_ctx_t *ctx, const uint8_t *data,
                                            unsigned int data_sz, void *, long)
{
    DecState *s = (DecState *)ctx->priv;
    s->have = false;
    if (!data || data_sz < sizeof(vpx_stub_pkt_hdr_t))
        return 1;
    vpx_stub_pkt_hdr_t hdr;
    memcpy(&hdr, data, sizeof(hdr));
    if (hdr.magic != VPX_STUB_MAGIC)
        return 1;
    vpx_image_t *img = &s->img;
    memset(img, 0, sizeof(*img));
    img->fmt = (vpx_img_fmt_t)hdr.fmt;
    img->d_w = hdr.d_w;
    img->d_h = hdr.d_h;
    const int sub = ((hdr.fmt & 0xff) == 2);
    img->x_chroma_shift = img->y_chroma_shift = sub ? 1 : 0;
    /* decoder-side images carry a border: wider pitch than the encoder's */
    size_t total = 0;
    size_t offs[3];
    for (int p = 0; p < 3; p++) {
        int pitch = (int)((hdr.row_bytes[p] + 31u) & ~31u) + (p ? VPX_STUB_DEC_BORDER / 2 : VPX_STUB_DEC_BORDER);
        img->stride[p] = pitch;
        offs[p] = total;
        total += (size_t)pitch * hdr.rows[p];
    }
    s->storage.assign(total, 0xA5); /* poison padding so stride bugs show */
    const uint8_t *src = data + sizeof(hdr);
    for (int p = 0; p < 3; p++) {
        img->planes[p] = s->storage.data() + offs[p];
        for (uint32_t y = 0; y < hdr.rows[p]; y++, src += hdr.row_bytes[p])
            memcpy(img->planes[p] + (size_t)y * img->stride[p], src, hdr.row_bytes[p]);
    }
    s->have = true;
    return VPX_CODEC_OK;
}

extern "C" vpx_image_t *vpx_codec_get_frame(vpx_codec_ctx_t *ctx, vpx_codec_iter_t *iter)
{
    DecState *s = (DecState *)ctx->priv;
    if (*iter || !s->have)
        return NULL;
    *iter = s;
    return &s->img;
}

/* ========================= in-memory container ============================== */
struct MkvStubFile {
    std::vector<std::vector<binary> > att_data;
    std::vector<unsigned int> att_id;
    std::vector<std::vector<uint8> > frames;
};

static std::map<std::string, MkvStubFile> &registry()
{
    static std::map<std::string, MkvStubFile> r;
    return r;
}
/* the map itself is shared by every thread of a multi-GPU driver; each "file" is only touched by its owner, and
 * std::map never moves its nodes, so only look-ups and insertions need the lock */
static std::mutex &registry_mutex()
{
    static std::mutex m;
    return m;
}

MkvStubFile *mkv_stub_find(const char *name)
{
    std::lock_guard<std::mutex> lock(registry_mutex());
    std::map<std::string, MkvStubFile>::iterator it = registry().find(name ? name : "");
    return it == registry().end() ? NULL : &it->second;
}
void mkv_stub_erase(const char *name)
{
    std::lock_guard<std::mutex> lock(registry_mutex());
    registry().erase(name ? name : "");
}
size_t mkv_stub_frame_count(const char *name)
{
    MkvStubFile *f = mkv_stub_find(name);
    return f ? f->frames.size() : 0;
}
const std::vector<uint8> *mkv_stub_frame(const char *name, size_t idx)
{
    MkvStubFile *f = mkv_stub_find(name);
    return (f && idx < f->frames.size()) ? &f->frames[idx] : NULL;
}
void mkv_stub_append_frame(const char *name, const uint8 *data, size_t n)
{
    MkvStubFile *f = mkv_stub_find(name);
    if (f)
        f->frames.push_back(std::vector<uint8>(data, data + n));
}

MkvInterface::MkvInterface() : m_file(NULL), m_readPos(0), m_verbose(false), m_frameDuration(40.0f) {}
MkvInterface::~MkvInterface() {}

void MkvInterface::openWrite(const char *outputFile, const unsigned int, const unsigned int,
                             const float, const float)
{
    std::lock_guard<std::mutex> lock(registry_mutex());
    registry()[outputFile] = MkvStubFile();
    m_file = &registry()[outputFile];
}

void MkvInterface::openRead(const char *inputFile)
{
    m_file = mkv_stub_find(inputFile);
    m_readPos = 0;
}

void MkvInterface::close() {}

void MkvInterface::addAttachment(unsigned int uid, const binary *buffer, unsigned int buffer_size,
                                 const char *)
{
    if (!m_file)
        return;
    m_file->att_id.push_back(uid);
    m_file->att_data.push_back(std::vector<binary>(buffer, buffer + buffer_size));
}

void MkvInterface::writeAttachments() {}

void MkvInterface::addFrame(const uint8 *frame_buffer, unsigned int buffer_size, bool)
{
    if (m_file)
        m_file->frames.push_back(std::vector<uint8>(frame_buffer, frame_buffer + buffer_size));
}

bool MkvInterface::getAttachment(unsigned int ind, binary **buffer, unsigned int &id,
                                 unsigned int &buffer_size)
{
    if (!m_file || ind >= m_file->att_id.size())
        return false;
    *buffer = m_file->att_data[ind].data();
    id = m_file->att_id[ind];
    buffer_size = (unsigned int)m_file->att_data[ind].size();
    return true;
}

bool MkvInterface::readFrame()
{
    if (!m_file || m_readPos >= m_file->frames.size())
        return false;
    m_readPos++;
    return true;
}

const uint8 *MkvInterface::getFrame(unsigned int &buffer_size)
{
    if (!m_file || m_readPos == 0 || m_readPos > m_file->frames.size()) {
        buffer_size = 0;
        return NULL;
    }
    buffer_size = (unsigned int)m_file->frames[m_readPos - 1].size();
    return m_file->frames[m_readPos - 1].data();
}

bool MkvInterface::seekToTime(float, bool) { return false; }

/*
 * luma_oracle.c -- TEST INFRASTRUCTURE ONLY (see luma_oracle.h).
 *
 * Plain-C restatement of the reference CPU algorithm for the per-pixel
 * HDR<->integer transform.  It is the checker, never the thing shipped or
 * measured as the product.  Build: oracle/Makefile (gcc -O2 -ffp-contract=off;
 * never -march=native / -ffast-math: FMA contraction changes integer planes).
 *
 * libm dependency: powf / log10f from the host glibc (the reference does not
 * vendor or pin a libm; src/luma_quantizer.cpp:493-499,507-509).
 */
#include "luma_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef LUMA_HAVE_PTF_TABLES
/* Constant tables are consumed from the reference checkout at build time
 * (include/luma/luma_quantizer.h:55-77); they are data, not retyped here. */
static const float tab_psi_10[] = {
#include "ptfs/ptf_jnd_ferwerda_10bit.h"
};
static const float tab_psi_11[] = {
#include "ptfs/ptf_jnd_ferwerda_11bit.h"
};
static const float tab_psi_12[] = {
#include "ptfs/ptf_jnd_ferwerda_12bit.h"
};
static const float tab_vdp_10[] = {
#include "ptfs/ptf_jnd_hdrvdp_10bit.h"
};
static const float tab_vdp_11[] = {
#include "ptfs/ptf_jnd_hdrvdp_11bit.h"
};
static const float tab_vdp_12[] = {
#include "ptfs/ptf_jnd_hdrvdp_12bit.h"
};
int lo_have_ptf_tables(void) { return 1; }
#else
int lo_have_ptf_tables(void) { return 0; }
#endif

/* std::min / std::max are compare-selects (NaN handling differs from fminf):
 * std::min(a,b) = (b<a)?b:a ; std::max(a,b) = (a<b)?b:a */
static inline float sel_min(float a, float b) { return (b < a) ? b : a; }
static inline float sel_max(float a, float b) { return (a < b) ? b : a; }

/* include/luma/luma_quantizer.h:79-87 */
static const float M_RGB2XYZ[3][3] = {{0.412424f, 0.357579f, 0.180464f},
                                      {0.212656f, 0.715158f, 0.072186f},
                                      {0.019332f, 0.119193f, 0.950444f}};
static const float M_XYZ2RGB[3][3] = {{3.240708f, -1.537259f, -0.498570f},
                                      {-0.969257f, 1.875995f, 0.041555f},
                                      {0.055636f, -0.203996f, 1.057069f}};

void lo_init(lo_quantizer *q)
{
    memset(q, 0, sizeof(*q));
    q->l_max = 10000.0f;
    q->l_min = 0.005f;
    q->color_space = LO_CS_LUV;
}

void lo_free(lo_quantizer *q)
{
    free(q->mapping);
    q->mapping = NULL;
}

/* src/luma_quantizer.cpp:485-501.  The PQ constants are the rounded literals
 * the reference uses, narrowed to float. */
float lo_transform_pq(const lo_quantizer *q, float val, int encode)
{
    const float L = q->l_max;
    const float m = 78.8438, n = 0.1593, c1 = 0.8359, c2 = 18.8516, c3 = 18.6875;
    if (encode) {
        float Lp = powf(val / L, n);
        return powf((c1 + c2 * Lp) / (1 + c3 * Lp), m);
    } else {
        float Vp = powf(val, 1.0f / m);
        return L * powf(sel_max(0.0f, (Vp - c1)) / (c2 - c3 * Vp), 1.0f / n);
    }
}

/* src/luma_quantizer.cpp:504-510 */
float lo_transform_log(const lo_quantizer *q, float val, int encode)
{
    if (encode)
        return (log10f(val) - log10f(q->l_min)) / (log10f(q->l_max) - log10f(q->l_min));
    return powf(10.0f, val * (log10f(q->l_max) - log10f(q->l_min)) + log10f(q->l_min));
}

/* src/luma_quantizer.cpp:172-212 */
int lo_set_quantizer(lo_quantizer *q, int ptf, unsigned bitdepth, int cs,
                     unsigned bitdepth_c, float max_lum, float min_lum)
{
    free(q->mapping);
    q->mapping = NULL;
    q->ptf = ptf;
    q->bitdepth = bitdepth;
    q->max_val = (unsigned)((int)powf(2.0f, (float)bitdepth) - 1);
    q->color_space = cs;
    q->bitdepth_color = bitdepth_c;
    q->max_val_color = (unsigned)((int)powf(2.0f, (float)bitdepth_c) - 1);
    q->l_max = max_lum;
    q->l_min = min_lum;
    q->mapping = (float *)malloc(sizeof(float) * ((size_t)q->max_val + 1));
    if (!q->mapping)
        return -2;

    const float *table = NULL;
    switch (ptf) {
    case LO_PTF_PQ: /* :114-118 */
        for (size_t i = 0; i <= q->max_val; i++)
            q->mapping[i] = lo_transform_pq(q, (float)i / q->max_val, 0);
        return 0;
    case LO_PTF_LOG: /* :121-125 */
        for (size_t i = 0; i <= q->max_val; i++)
            q->mapping[i] = lo_transform_log(q, (float)i / q->max_val, 0);
        return 0;
    case LO_PTF_LINEAR: /* :200-203 */
        for (size_t i = 0; i <= q->max_val; i++)
            q->mapping[i] = q->l_max * ((float)i / q->max_val);
        return 0;
#ifdef LUMA_HAVE_PTF_TABLES
    case LO_PTF_JND_HDRVDP: /* :128-147; any depth other than 10/11 reads the 12-bit table */
        table = bitdepth == 10 ? tab_vdp_10 : bitdepth == 11 ? tab_vdp_11 : tab_vdp_12;
        break;
    case LO_PTF_PSI: /* :150-169 */
    default:
        table = bitdepth == 10 ? tab_psi_10 : bitdepth == 11 ? tab_psi_11 : tab_psi_12;
        break;
#else
    default:
        (void)table;
        return -1;
#endif
    }
#ifdef LUMA_HAVE_PTF_TABLES
    /* the reference would read past the table for bitdepth > 12; refuse */
    if (q->max_val + 1 > 4096u && bitdepth != 10 && bitdepth != 11)
        return -3;
    for (size_t i = 0; i <= q->max_val; i++)
        q->mapping[i] = table[i];
    return 0;
#endif
}

/* src/luma_quantizer.cpp:215-244 */
float lo_quantize(const lo_quantizer *q, float val, unsigned ch)
{
    if (ch == 0 || q->color_space == LO_CS_RGB || q->color_space == LO_CS_XYZ) {
        int lo = 0, hi = (int)q->max_val;
        while (lo + 1 < hi) {
            int mid = (lo + hi) / 2;
            if (val < q->mapping[mid])
                hi = mid;
            else
                lo = mid;
        }
        return (val - q->mapping[lo] < q->mapping[hi] - val) ? (float)lo : (float)hi;
    }
    float res = floorf((float)q->max_val_color * val + 0.5f);
    return sel_max(0.0f, sel_min((float)q->max_val_color, res));
}

/* src/luma_quantizer.cpp:247-264 */
float lo_dequantize(const lo_quantizer *q, float val, unsigned ch)
{
    if (ch == 0 || q->color_space == LO_CS_RGB || q->color_space == LO_CS_XYZ) {
        if (val < 0)
            return q->mapping[0];
        if (val >= q->max_val)
            return q->mapping[q->max_val];
        return q->mapping[(int)val];
    }
    return sel_max(val / (float)q->max_val_color, 1e-10f);
}

static inline float clamp_xyz(float v) { return sel_max(sel_min(v, 100000000.0f), 0.0001f); }

static inline float dot3(const float m[3], float a, float b, float c)
{
    return m[0] * a + m[1] * b + m[2] * c; /* ((m0*a)+(m1*b))+(m2*c) */
}

/* src/luma_quantizer.cpp:267-482 */
int lo_transform_color_space(const lo_quantizer *q, float *frame, unsigned w,
                             unsigned h, int to_cs, float sc)
{
    const size_t n = (size_t)w * h;
    float *c0 = frame, *c1 = frame + n, *c2 = frame + 2 * n;

    if (to_cs) {
        switch (q->color_space) {
        case LO_CS_XYZ: /* :273-290 */
            for (size_t i = 0; i < n; i++) {
                float R = c0[i] * sc, G = c1[i] * sc, B = c2[i] * sc;
                c0[i] = clamp_xyz(dot3(M_RGB2XYZ[0], R, G, B));
                c1[i] = clamp_xyz(dot3(M_RGB2XYZ[1], R, G, B));
                c2[i] = clamp_xyz(dot3(M_RGB2XYZ[2], R, G, B));
            }
            return 1;
        case LO_CS_LUV: /* :291-316 */
            for (size_t i = 0; i < n; i++) {
                float R = c0[i] * sc, G = c1[i] * sc, B = c2[i] * sc;
                float X = clamp_xyz(dot3(M_RGB2XYZ[0], R, G, B));
                float Y = clamp_xyz(dot3(M_RGB2XYZ[1], R, G, B));
                float Z = clamp_xyz(dot3(M_RGB2XYZ[2], R, G, B));
                float sum = X + Y + Z;
                float x = X / sum;
                float y = Y / sum;
                c0[i] = Y;
                c1[i] = 4.0f * x / (-2.0f * x + 12.0f * y + 3.0f) * 410.f / 255.0f;
                c2[i] = 9.0f * y / (-2.0f * x + 12.0f * y + 3.0f) * 410.f / 255.0f;
            }
            return 1;
        case LO_CS_YCBCR: /* :317-354, BT.2020 on PQ-encoded R'G'B' */
            for (size_t i = 0; i < n; i++) {
                float R = lo_transform_pq(q, sel_max(c0[i] * sc, 1e-10f), 1);
                float G = lo_transform_pq(q, sel_max(c1[i] * sc, 1e-10f), 1);
                float B = lo_transform_pq(q, sel_max(c2[i] * sc, 1e-10f), 1);
                float y = 0.2627f * R + 0.6780f * G + 0.0593f * B;
                c0[i] = lo_transform_pq(q, (219.0f * y + 16.0f) / 255.0f, 0);
                c1[i] = (224.0f * ((B - y) / 1.8814f) + 128.0f) / 255.0f;
                c2[i] = (224.0f * ((R - y) / 1.4746f) + 128.0f) / 255.0f;
            }
            return 1;
        case LO_CS_RGB: /* :355-367 */
            for (size_t i = 0; i < n; i++) {
                c0[i] *= sc;
                c1[i] *= sc;
                c2[i] *= sc;
            }
            return 1;
        default:
            return 0;
        }
    }

    switch (q->color_space) {
    case LO_CS_XYZ: /* :378-395 */
        for (size_t i = 0; i < n; i++) {
            float X = c0[i], Y = c1[i], Z = c2[i];
            c0[i] = dot3(M_XYZ2RGB[0], X, Y, Z) / sc;
            c1[i] = dot3(M_XYZ2RGB[1], X, Y, Z) / sc;
            c2[i] = dot3(M_XYZ2RGB[2], X, Y, Z) / sc;
        }
        return 1;
    case LO_CS_LUV: /* :396-421 */
        for (size_t i = 0; i < n; i++) {
            float L = c0[i];
            float u = c1[i] * 255.0f / 410.0f;
            float v = c2[i] * 255.0f / 410.0f;
            float x = 9.0f * u / (6.0f * u - 16.0f * v + 12.0f);
            float y = 4.0f * v / (6.0f * u - 16.0f * v + 12.0f);
            float Y = clamp_xyz(L);
            float X = clamp_xyz(x / y * L);
            float Z = clamp_xyz((1.0f - x - y) / y * L);
            c0[i] = dot3(M_XYZ2RGB[0], X, Y, Z) / sc;
            c1[i] = dot3(M_XYZ2RGB[1], X, Y, Z) / sc;
            c2[i] = dot3(M_XYZ2RGB[2], X, Y, Z) / sc;
        }
        return 1;
    case LO_CS_RGB: /* :422-435 */
        for (size_t i = 0; i < n; i++) {
            c0[i] /= sc;
            c1[i] /= sc;
            c2[i] /= sc;
        }
        return 1;
    case LO_CS_YCBCR: /* :436-473 */
        for (size_t i = 0; i < n; i++) {
            float y = lo_transform_pq(q, c0[i], 1);
            y = (255.0f * y - 16.0f) / 219.0f;
            float blue = y + 1.8814f * (255.0f * c1[i] - 128.0f) / 224.0f;
            float red = y + 1.4746f * (255.0f * c2[i] - 128.0f) / 224.0f;
            float green = (y - 0.2627f * red - 0.0593f * blue) / 0.6780f;
            red = sel_max(0.0f, sel_min(1.0f, red));
            green = sel_max(0.0f, sel_min(1.0f, green));
            blue = sel_max(0.0f, sel_min(1.0f, blue));
            c0[i] = lo_transform_pq(q, red, 0) / sc;
            c1[i] = lo_transform_pq(q, green, 0) / sc;
            c2[i] = lo_transform_pq(q, blue, 0) / sc;
        }
        return 1;
    default:
        return 0;
    }
}

/* src/luma_encoder.cpp:265-269 (+ profile<->format table :121-128) */
void lo_plane_dims(unsigned w, unsigned h, int profile, int pw[3], int ph[3])
{
    const int sub = (profile == 0 || profile == 2);
    pw[0] = (int)w;
    ph[0] = (int)h;
    for (int p = 1; p < 3; p++) {
        pw[p] = sub ? (int)((w + 1) >> 1) : (int)w;
        ph[p] = sub ? (int)((h + 1) >> 1) : (int)h;
    }
}

/* libvpx 1.6.1 vpx_img_alloc pitch rule as used at src/luma_encoder.cpp:121-128 */
void lo_vpx_strides(unsigned w, int profile, int align, int strides[3])
{
    const int sub = (profile == 0 || profile == 2);
    const int bytes = profile > 1 ? 2 : 1;
    unsigned aw = sub ? ((w + 1u) & ~1u) : w;
    unsigned s = (aw + (unsigned)align - 1u) & ~((unsigned)align - 1u);
    strides[0] = (int)(s * bytes);
    strides[1] = strides[2] = sub ? strides[0] >> 1 : strides[0];
}

/* src/luma_encoder.cpp:260-317 for one plane */
static float pack_plane(const lo_quantizer *q, const float *src, unsigned fw,
                        unsigned fh, int profile, int plane, uint8_t *buf, int stride)
{
    int pw[3], ph[3];
    lo_plane_dims(fw, fh, profile, pw, ph);
    const int w = pw[plane], h = ph[plane];
    const int m = profile > 1 ? 2 : 1;
    const int subsample = plane && (profile == 2 || profile == 0);
    float avg = 0.0f;

    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            float res;
            if (subsample) {
                size_t i1 = 2 * (size_t)x + 4 * (size_t)y * w;
                size_t i2 = i1 + 2 * (size_t)w;
                res = 0.25f * (src[i1] + src[i1 + 1] + src[i2] + src[i2 + 1]);
            } else {
                res = src[x + (size_t)y * w];
                avg += res;
            }
            res = lo_quantize(q, res, (unsigned)plane);
            if (profile > 1) {
                /* "unsigned char bl = res/256; bh = res - bl*256" : LE u16 */
                unsigned char hi8 = (unsigned char)(int)(res / 256);
                unsigned char lo8 = (unsigned char)(int)(res - hi8 * 256);
                buf[m * x + (size_t)y * stride + 1] = hi8;
                buf[m * x + (size_t)y * stride] = lo8;
            } else {
                buf[m * x + (size_t)y * stride] = (unsigned char)(int)res;
            }
        }
    }
    return avg / (w * h);
}

void lo_pack_planes(const lo_quantizer *q, const float *frame, unsigned w,
                    unsigned h, int profile, uint8_t *const planes[3],
                    const int strides[3], float avg_out[3])
{
    const size_t n = (size_t)w * h;
    for (int p = 0; p < 3; p++) {
        float a = pack_plane(q, frame + p * n, w, h, profile, p, planes[p], strides[p]);
        if (avg_out)
            avg_out[p] = a;
    }
}

/* src/luma_decoder.cpp:205-240 */
void lo_unpack_planes(const lo_quantizer *q, const uint8_t *const planes[3],
                      const int strides[3], unsigned fw, unsigned fh, int profile,
                      float *frame)
{
    int pw[3], ph[3];
    lo_plane_dims(fw, fh, profile, pw, ph);
    const size_t n = (size_t)fw * fh;
    for (int plane = 0; plane < 3; plane++) {
        float *dest = frame + plane * n;
        const uint8_t *buf = planes[plane];
        const int w = pw[plane], h = ph[plane], stride = strides[plane];
        const int upsample = plane && (profile == 2 || profile == 0);
        for (int y = 0; y < h; y++) {
            for (int x = 0; x < w; x++) {
                float val;
                if (profile > 1)
                    val = lo_dequantize(q, buf[2 * x + (size_t)y * stride + 1] * 256.0f +
                                               buf[2 * x + (size_t)y * stride],
                                        (unsigned)plane);
                else
                    val = lo_dequantize(q, buf[x + (size_t)y * stride], (unsigned)plane);
                if (upsample) {
                    size_t i1 = 2 * (size_t)x + 4 * (size_t)y * w;
                    size_t i2 = i1 + 2 * (size_t)w;
                    dest[i1] = dest[i1 + 1] = dest[i2] = dest[i2 + 1] = val;
                } else {
                    dest[x + (size_t)y * w] = val;
                }
            }
        }
    }
}

void lo_encode(const lo_quantizer *q, float *frame, unsigned w, unsigned h,
               int profile, float pre_scaling, uint8_t *const planes[3],
               const int strides[3], float avg_out[3])
{
    lo_transform_color_space(q, frame, w, h, 1, pre_scaling);
    lo_pack_planes(q, frame, w, h, profile, planes, strides, avg_out);
}

void lo_decode(const lo_quantizer *q, const uint8_t *const planes[3],
               const int strides[3], unsigned w, unsigned h, int profile,
               float pre_scaling, float *frame)
{
    lo_unpack_planes(q, planes, strides, w, h, profile, frame);
    lo_transform_color_space(q, frame, w, h, 0, pre_scaling);
}

/* src/exr_interface.cpp:50-70: synthetic HDR pattern (ramps, steps, checker) */
void lo_test_frame(float *frame, unsigned w, unsigned h)
{
    const size_t n = (size_t)w * h;
    float *r = frame, *g = frame + n, *b = frame + 2 * n;
    for (size_t y = 0; y < h; y++) {
        for (size_t x = 0; x < w; x++) {
            const size_t i = x + y * w;
            if (y < h / 5) {
                float v = (y < h / 10) ? 10000.0f * ((float)(x * x)) / (w * w)
                                       : 10000.0f * ((20 * x) / w) / 20.0f;
                r[i] = g[i] = b[i] = v;
            } else {
                const size_t band = (20 * y / h) % 2;
                r[i] = 10000.0f * (band ^ ((30 * x / w) % 2));
                g[i] = 10000.0f * band * ((float)(y * y)) / (h * h);
                b[i] = 10000.0f * band * ((float)(x * x)) / (w * w);
            }
        }
    }
}

uint32_t lo_fnv1a32(const void *data, size_t n, uint32_t seed)
{
    const uint8_t *p = (const uint8_t *)data;
    uint32_t hsh = seed;
    for (size_t i = 0; i < n; i++) {
        hsh ^= p[i];
        hsh *= 16777619u;
    }
    return hsh;
}

uint32_t lo_hash_plane(const uint8_t *plane, int stride, int row_bytes, int rows)
{
    uint32_t hsh = 2166136261u;
    for (int y = 0; y < rows; y++)
        hsh = lo_fnv1a32(plane + (size_t)y * stride, (size_t)row_bytes, hsh);
    return hsh;
}

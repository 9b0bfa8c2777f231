/*
 * luma_kernels_decl.cuh -- argument blocks shared by the kernels and the host launcher.
 */
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace lumacu {

enum { CS_LUV = 0, CS_RGB = 1, CS_YCBCR = 2, CS_XYZ = 3 };
enum { SEARCH_BUCKET = 0, SEARCH_BINARY = 1, SEARCH_LITERAL = 2 };

/* Quantizer state as the kernels see it (device pointers into the context). */
struct QuantDev {
    const float *lut;       /* code -> luminance, max_val + 1 entries (reference m_mapping) */
    const uint32_t *thr;    /* ordered keys of the max_val decision thresholds + pad sentinels */
    const uint16_t *bucket; /* first candidate code per key bucket (SEARCH_BUCKET) */
    uint32_t max_val;
    uint32_t max_val_color;
    float max_val_f;
    float max_val_color_f;
    float l_max;
    int search_mode;
    uint32_t shift, base, nbm1, walk; /* bucket = clamp(key >> shift, base, base + nbm1) - base */
    uint32_t thr_count;               /* max_val + pad */
    uint32_t smem_tables;             /* 1: stage thr (+bucket) in shared memory; 0: read from global */
    uint32_t smem_lut;                /* decode: 1 = stage the LUT in shared memory */
    const float *ctab;                /* chroma code -> u'/v' (LUV) or Cb/Cr (YCBCR), max_val_color + 1 entries */
    /* direct search table for values known to lie in [1e-4, 1e8] (Lu'v' Y, XYZ): one 32-bit entry per key
     * bucket, code = (dtab[(key >> d_shift) - d_lo] + key) >> 16 (luma_fast.cuh search_direct); NULL when the
     * LUT does not qualify */
    const uint32_t *dtab;
    uint32_t d_shift, d_lo, d_n;
    uint32_t d_lo_key, d_hi_key; /* keys are clamped to [d_lo_key, d_hi_key] first (d_lo_key 0 = no lower clamp needed) */
    uint32_t d_double;           /* 1: 64-bit entries {A, B}, up to two thresholds per bucket (wide LUTs; lumacu.cu) */
    uint32_t d_global;           /* 1: too large for shared memory (13-16 bit LUTs): the kernels read dtab from global memory */
    /* CS_YCBCR decode: code -> ((255 PQenc(lut[code])) - 16) / 219 (src/luma_quantizer.cpp:447-448), host-built with the
     * host libm: two of the eight per-pixel powf calls become a table read.  NULL for the other colour spaces. */
    const float *ylut;
    /* CS_YCBCR, tuned kernels: exhaustive L2-resident tables of PQ decode (step table over every float in [2^-8, 1]) and
     * of the outer power of PQ encode (one entry per float in [0.835, 1.01]); luma_pq_tables.cuh.  NULL = evaluate. */
    const uint4 *pqd;
    const float *pqe;
    /* CS_YCBCR encode: direct search table keyed on v = (219 y' + 16)/255 (kVdEntries entries, luma_pq_tables.cuh (3));
     * NULL when it could not be built for this LUT */
    const uint32_t *vdtab;
    /* CS_YCBCR encode: RN(1 / l_max) when val / l_max may be computed as q = val * rc, q += (val - q * l_max) * rc (two
     * FMAs) -- proved equal to the IEEE quotient for EVERY float val >= 1e-10 by an exhaustive device check at
     * set_quantizer time (luma_pq_tables.cuh check_lmax_division_kernel); 0 = divide */
    float lmax_rc;
    uint32_t tune_flags; /* tuning sweeps: bit 0 = CS_YCBCR decode evaluates the green of both pixels of a pair */
};

constexpr int kThreads = 256;

struct StatsPartial {
    double sum;
    float mx;
    float mn;
};

struct FrameStatsDev { /* same layout as lumacu_frame_stats */
    double sum;
    float mx;
    float mn;
};

struct EncArgs {
    QuantDev q;
    const float *rgb;
    float *rgb_out; /* nullable: colour-transformed frame (reference's in-place side effect) */
    size_t rgb_plane_stride; /* floats between the R, G, B planes */
    size_t rgb_frame_stride; /* floats between frames */
    size_t out_plane_stride, out_frame_stride;
    uint32_t w, h;
    uint8_t *plane[3];
    int32_t stride[3];
    size_t plane_frame_stride[3];
    float sc;
    int prescale; /* sc != 1 */
    StatsPartial *partial; /* [frames][gridDim.x], nullable together with stats */
    uint32_t *counter;     /* [frames] */
    FrameStatsDev *stats;  /* [frames] */
    float2 nz;             /* (-0.0f, -0.0f), see luma_fast.cuh mul2_nc */
    int passthrough;       /* generic kernels: the frame is already colour-transformed (setChannels) */
    /* screened-chroma kernels (luma_fast.cuh FASTC): t = screen_k * (sum of the 2x2 block's X/D resp. Y/D) + 0.5 with
     * screen_k1 = RN(maxC/4 * 4 * 410/255), screen_k2 = RN(maxC/4 * 9 * 410/255) */
    float screen_k1, screen_k2;
    /* CS_YCBCR tuned kernels: PQenc(max(c * sc, 1e-10)) for every half-float bit pattern c (luma_pq_tables.cuh (4)); NULL = off */
    const float *pqh;
    /* tensor-map staged kernels: CUtensorMap over the frame batch, dims {w, h, 3 planes, frames} of f32,
     * box {128, 2, 3, 1} (opaque 128 bytes so that this header does not need cuda.h) */
    alignas(64) unsigned char rgb_tmap[128];
};

struct DecArgs {
    QuantDev q;
    const uint8_t *plane[3];
    int32_t stride[3];
    size_t plane_frame_stride[3];
    float *rgb;
    size_t rgb_plane_stride;
    size_t rgb_frame_stride;
    uint32_t w, h;
    float sc;
    int prescale;
    float2 nz; /* (-0.0f, -0.0f), see luma_fast.cuh mul2_nc */
    int passthrough; /* generic kernels: stop after dequantisation + chroma replication (getVpxChannels) */
    /* generic kernels, display mode (lumacu_display*): instead of storing the float frame, apply the display tail of
     * the reference's player shader (src/lumaplay_dequantizer.frag:141-156) and store 8-bit RGBA */
    uint8_t *rgba;        /* NULL = normal decode */
    int32_t rgba_pitch;   /* bytes per output row */
    size_t rgba_frame_stride;
    float disp_exposure, disp_scaling, disp_inv_gamma;
    int disp_tmo, disp_ldr;
};

} // namespace lumacu

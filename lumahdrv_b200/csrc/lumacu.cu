/*
 * lumacu.cu -- C ABI (include/lumacu.h) over the sm_100a kernels.
 *
 * Host responsibilities kept here:
 *   - LUT construction with the host libm, exactly as
 *     LumaQuantizer::setQuantizer does (reference src/luma_quantizer.cpp:172-212);
 *   - derivation of the exact decision thresholds of LumaQuantizer::quantize
 *     (src/luma_quantizer.cpp:219-235) and of the bucket table that turns the
 *     reference's 11-step bisection into one table read plus <= `walk` compares;
 *   - the direct search tables (one or two thresholds per bucket) built from them;
 *   - launch configuration (persistent grid = SM count x resident blocks), kernel
 *     selection, the device-built CS_YCBCR tables;
 *   - staging for the host-pointer entry points (banded, optionally asynchronous),
 *     the GPU-to-GPU quantizer broadcast, error reporting.
 * There is no CPU implementation of the transform in this library: without a
 * CUDA device every compute entry point fails with LUMACU_ERR_NO_DEVICE/_CUDA.
 */
#include "../../include/lumacu.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda.h> /* CUtensorMap + enums only; the encoder entry point is resolved at run time */

#include <algorithm>
#include <new>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "luma_dispatch.h"

using namespace lumacu;

#ifdef LUMA_HAVE_PTF_TABLES
/* Constant perceptual-transfer-function tables are consumed from the reference
 * checkout at build time (include/luma/luma_quantizer.h:55-77): data, not code. */
static const float k_ptf_psi_10[] = {
#include "ptfs/ptf_jnd_ferwerda_10bit.h"
};
static const float k_ptf_psi_11[] = {
#include "ptfs/ptf_jnd_ferwerda_11bit.h"
};
static const float k_ptf_psi_12[] = {
#include "ptfs/ptf_jnd_ferwerda_12bit.h"
};
static const float k_ptf_vdp_10[] = {
#include "ptfs/ptf_jnd_hdrvdp_10bit.h"
};
static const float k_ptf_vdp_11[] = {
#include "ptfs/ptf_jnd_hdrvdp_11bit.h"
};
static const float k_ptf_vdp_12[] = {
#include "ptfs/ptf_jnd_hdrvdp_12bit.h"
};
#endif

namespace {

constexpr uint32_t kMaxBuckets = 8192; /* u16 heads: <= 16 KB of shared memory */
constexpr uint32_t kMaxWalk = 8;
constexpr size_t kMaxSmemTables = 96 * 1024;
constexpr size_t kMaxSmemLut = 64 * 1024;

thread_local std::string g_create_error;

struct DeviceBuffer {
    void *p = nullptr;
    size_t cap = 0;
};

/* What the internal launchers need beyond the public *_dev signatures. */
struct LaunchOpts {
    bool passthrough = false;      /* skip the colour transform: LumaEncoder::setChannels / LumaDecoder::getVpxChannels halves */
    size_t rgb_plane_stride = 0;   /* floats between the R, G, B planes (0 = w*h): row bands of a larger frame */
    /* display launch (decode only): tone curve + gamma -> 8-bit RGBA instead of the float frame */
    const lumacu_display_params *display = nullptr;
    uint8_t *rgba = nullptr;
    int32_t rgba_pitch = 0;
    size_t rgba_frame_stride = 0;
};

} // namespace

struct lumacu_ctx {
    int device = -1;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    std::string err;
    uint64_t launches = 0;
    std::unordered_map<const void *, std::pair<size_t, int>> occupancy; /* kernel -> (smem, blocks per SM) */
    bool force_generic = false; /* tests: run the literal kernels */
    bool last_fast = false;     /* the last encode/decode launch used a tuned kernel */
    int enc_variant = 0, dec_variant = 0; /* tuning sweep: which instantiation of the tuned kernels (0 = default) */
    int grid_cap = 0;                     /* tuning sweep: cap on resident blocks per SM (0 = occupancy) */
    int grid_tpt = 0;                     /* tuning sweep: tiles per thread of a multi-frame launch (0 = default) */
    bool lmax_div_checked = false;        /* CS_YCBCR: val / Lmax by reciprocal + FMA verified for lmax_div_lmax (lmax_div_rc 0 = it is not exact) */
    float lmax_div_lmax = 0.0f, lmax_div_rc = 0.0f;
    bool no_direct = false;               /* tuning sweep / tests: bucket + threshold search even when the direct table exists */
    bool global_direct = false;           /* tuning sweep / tests: read the (32-bit) direct table from global memory, no staging */

    /* quantizer */
    bool configured = false;
    int color_space = 0;
    QuantDev q{};
    std::vector<float> h_lut;
    DeviceBuffer d_tables; /* lut | thr | bucket | ctab | dtab | ylut */
    DeviceBuffer d_pq;       /* CS_YCBCR: pqd | pqe (luma_pq_tables.cuh), built on the device for pq_lmax */
    DeviceBuffer d_vd;       /* CS_YCBCR: v-keyed luma search table + flag word, rebuilt with every quantizer */
    DeviceBuffer d_pqh;      /* CS_YCBCR: PQ encode of every half-float input value for (pqh_sc, pqh_lmax) */
    float pqh_sc = 0.0f, pqh_lmax = 0.0f;
    bool pqh_valid = false;
    float pq_lmax = 0.0f;
    bool pq_valid = false, pq_off = false; /* pq_off: tests / sweeps run the tuned kernels without the tables */
    size_t tables_bytes = 0; /* bytes of d_tables in use (what lumacu_broadcast_quantizer copies to the peers) */
    float max_lum = 0.0f;
    size_t smem_enc = 0, smem_dec = 0;
    size_t smem_dec_fast = 0; /* lut + chroma table; 0 = fast decode unavailable */
    bool dec_global_lut = false; /* tuned decode with the luma LUT in global memory, chroma table (smem_dec_ctab bytes) in shared */
    size_t smem_dec_ctab = 0;
    bool fast_enc_ok = false;

    /* stats workspace */
    /* block partials + per-frame arrival counters of the statistics reduction: one pair PER STREAM the context has
     * launched on, so that launches on different caller streams do not share them */
    struct StatsWs {
        DeviceBuffer partial, counter;
    };
    std::unordered_map<cudaStream_t, StatsWs> stats_ws;

    /* staging for the host-pointer entry points: row bands flow H2D (s_in) -> kernel (stream) -> D2H (s_out) */
    DeviceBuffer d_rgb, d_planes, d_stats, d_aux;
    DeviceBuffer d_half; /* interleaved half RGBA staging of the *_half_rgba host entry points (8 B/px) */
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_k;
    int host_bands = 0;               /* tuning: number of row bands per host-pointer call (0 = automatic) */
    void *h_pin = nullptr;
    size_t h_pin_cap = 0;
    /* asynchronous host-pointer call in flight (lumacu_encode_async / lumacu_decode_async -> lumacu_wait) */
    bool pending = false;
    lumacu_frame_stats *pending_stats = nullptr; /* caller's stats to fill at lumacu_wait (NULL = none) */
    int pending_bands = 0;
    cudaEvent_t ev_input = nullptr; /* recorded after the last H2D copy of the call: the caller's input is reusable */
    bool pageable_warned = false;   /* the one-time note about pageable caller memory has been printed */
};

static int finish_pending(lumacu_ctx *ctx);

namespace {

int fail(lumacu_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_create_error = buf;
    return code;
}

#define CU_TRY(ctx, expr)                                                                         \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? LUMACU_ERR_OUT_OF_MEMORY : LUMACU_ERR_CUDA, \
                        "%s failed: %s", #expr, cudaGetErrorString(e_));                         \
    } while (0)

/* No exception crosses the C ABI (lumacu.h): every entry point is a function-try-block ending in this. */
int on_exception(lumacu_ctx *ctx, bool oom) noexcept
{
    try {
        return fail(ctx, oom ? LUMACU_ERR_OUT_OF_MEMORY : LUMACU_ERR_CUDA,
                    oom ? "host allocation failed (std::bad_alloc)" : "unexpected C++ exception");
    } catch (...) {
        return oom ? LUMACU_ERR_OUT_OF_MEMORY : LUMACU_ERR_CUDA;
    }
}
#define LUMACU_CATCH(ctx_expr)                                  \
    catch (const std::bad_alloc &) { return on_exception(ctx_expr, true); } \
    catch (...) { return on_exception(ctx_expr, false); }

int reserve(lumacu_ctx *ctx, DeviceBuffer &b, size_t bytes)
{
    if (bytes <= b.cap)
        return LUMACU_OK;
    if (b.p) {
        /* launches on caller-provided streams may still use the old buffer */
        CU_TRY(ctx, cudaDeviceSynchronize());
        cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    CU_TRY(ctx, cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return LUMACU_OK;
}

/* ---- order-preserving integer key of a float (same mapping as the device) ------ */
inline uint32_t f2u(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
inline float u2f(uint32_t u)
{
    float f;
    memcpy(&f, &u, 4);
    return f;
}
inline uint32_t key_of(float f)
{
    uint32_t b = f2u(f);
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
inline float float_of_key(uint32_t k)
{
    uint32_t b = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
    return u2f(b);
}

/* the reference's final decision for the bracket (l, l+1):
 * "val - m[l] < m[r] - val ? l : r" (src/luma_quantizer.cpp:232) */
inline bool picks_upper(float val, float ml, float mr)
{
    volatile float a = val - ml; /* volatile: keep fp32 rounding, no contraction/excess precision */
    volatile float b = mr - val;
    return !(a < b);
}

} // namespace

/* ================================ host-only helpers ================================= */

extern "C" int lumacu_version(void) { return LUMACU_VERSION; }

extern "C" const char *lumacu_status_name(int status)
{
    switch (status) {
    case LUMACU_OK: return "LUMACU_OK";
    case LUMACU_ERR_INVALID_ARGUMENT: return "LUMACU_ERR_INVALID_ARGUMENT";
    case LUMACU_ERR_CUDA: return "LUMACU_ERR_CUDA";
    case LUMACU_ERR_NOT_CONFIGURED: return "LUMACU_ERR_NOT_CONFIGURED";
    case LUMACU_ERR_UNSUPPORTED: return "LUMACU_ERR_UNSUPPORTED";
    case LUMACU_ERR_OUT_OF_MEMORY: return "LUMACU_ERR_OUT_OF_MEMORY";
    case LUMACU_ERR_NO_DEVICE: return "LUMACU_ERR_NO_DEVICE";
    default: return "LUMACU_ERR_UNKNOWN";
    }
}

extern "C" int lumacu_device_count(int *count)
try {
    if (!count)
        return LUMACU_ERR_INVALID_ARGUMENT;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
        return fail(nullptr, LUMACU_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return LUMACU_OK;
}
LUMACU_CATCH(nullptr)

extern "C" int lumacu_have_ptf_tables(void)
{
#ifdef LUMA_HAVE_PTF_TABLES
    return 1;
#else
    return 0;
#endif
}

/* PQ / log curves with the host libm, float arithmetic as in the reference
 * (src/luma_quantizer.cpp:485-510; rounded PQ constants narrowed to float). */
static float host_pq_decode(float val, float L)
{
    const float m = 78.8438, n = 0.1593, c1 = 0.8359, c2 = 18.8516, c3 = 18.6875;
    float Vp = powf(val, 1.0f / m);
    float num = Vp - c1;
    num = (0.0f < num) ? num : 0.0f;
    return L * powf(num / (c2 - c3 * Vp), 1.0f / n);
}

static float host_pq_encode(float val, float L)
{
    const float m = 78.8438, n = 0.1593, c1 = 0.8359, c2 = 18.8516, c3 = 18.6875;
    volatile float Lp = powf(val / L, n);
    volatile float num = c1 + c2 * Lp;
    volatile float den = 1.0f + c3 * Lp;
    return powf(num / den, m);
}

extern "C" int lumacu_build_lut(int ptf, unsigned bitdepth, float max_lum, float min_lum, float *lut_out, size_t cap)
try {
    if (!lut_out || bitdepth > 16)
        return fail(nullptr, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_build_lut: bad arguments");
    const unsigned max_val = (unsigned)((int)powf(2.0f, (float)bitdepth) - 1);
    if (cap < (size_t)max_val + 1)
        return fail(nullptr, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_build_lut: capacity %zu < %u", cap, max_val + 1);
    const float *table = nullptr;
    switch (ptf) {
    case LUMACU_PTF_PQ:
        for (unsigned i = 0; i <= max_val; i++)
            lut_out[i] = host_pq_decode((float)i / max_val, max_lum);
        return LUMACU_OK;
    case LUMACU_PTF_LOG: {
        for (unsigned i = 0; i <= max_val; i++) {
            float v = (float)i / max_val;
            lut_out[i] = powf(10.0f, v * (log10f(max_lum) - log10f(min_lum)) + log10f(min_lum));
        }
        return LUMACU_OK;
    }
    case LUMACU_PTF_LINEAR:
        for (unsigned i = 0; i <= max_val; i++)
            lut_out[i] = max_lum * ((float)i / max_val);
        return LUMACU_OK;
#ifdef LUMA_HAVE_PTF_TABLES
    case LUMACU_PTF_JND_HDRVDP:
        table = bitdepth == 10 ? k_ptf_vdp_10 : bitdepth == 11 ? k_ptf_vdp_11 : k_ptf_vdp_12;
        break;
    case LUMACU_PTF_PSI:
        table = bitdepth == 10 ? k_ptf_psi_10 : bitdepth == 11 ? k_ptf_psi_11 : k_ptf_psi_12;
        break;
#else
    case LUMACU_PTF_JND_HDRVDP:
    case LUMACU_PTF_PSI:
        (void)table;
        return fail(nullptr, LUMACU_ERR_UNSUPPORTED,
                    "lumacu_build_lut: library was built without the reference's ptfs/ tables");
#endif
    default:
        return fail(nullptr, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_build_lut: unknown ptf %d", ptf);
    }
#ifdef LUMA_HAVE_PTF_TABLES
    /* any depth other than 10/11 reads the 12-bit table (reference quirk, :131-143); the
     * reference would read past its end for depths above 12 -- refuse instead */
    if (max_val + 1 > 4096u)
        return fail(nullptr, LUMACU_ERR_UNSUPPORTED, "lumacu_build_lut: table PTFs hold at most 4096 entries");
    for (unsigned i = 0; i <= max_val; i++)
        lut_out[i] = table[i];
    return LUMACU_OK;
#endif
}
LUMACU_CATCH(nullptr)

/* Thresholds of the reference's nearest-code decision.  Returns 1 and fills
 * thr_keys[0 .. lut_len-2] (ordered keys, strictly increasing) when the LUT is
 * finite and strictly increasing; returns 0 otherwise (literal search needed). */
extern "C" int lumacu_derive_thresholds(const float *lut, uint32_t lut_len, uint32_t *thr_keys)
{
    if (!lut || lut_len < 2 || !thr_keys)
        return 0;
    for (uint32_t i = 0; i < lut_len; i++)
        if (!isfinite(lut[i]) || (i && !(lut[i - 1] < lut[i])))
            return 0;
    for (uint32_t k = 1; k < lut_len; k++) {
        const float ml = lut[k - 1], mr = lut[k];
        /* predicate picks_upper is false at ml, true at mr, monotone in between */
        uint32_t lo = key_of(ml), hi = key_of(mr);
        while (hi - lo > 1) {
            uint32_t mid = lo + ((hi - lo) >> 1);
            if (picks_upper(float_of_key(mid), ml, mr))
                hi = mid;
            else
                lo = mid;
        }
        thr_keys[k - 1] = hi;
    }
    return 1;
}

/* Chooses the bucket table: the finest key shift whose table has at most
 * kMaxBuckets entries; walk = the largest number of thresholds in one bucket. */
extern "C" int lumacu_plan_buckets(const uint32_t *thr_keys, uint32_t n_thr, uint32_t *shift, uint32_t *base,
                                   uint32_t *n_buckets, uint32_t *walk)
{
    if (!thr_keys || !n_thr)
        return 0;
    for (uint32_t s = 8; s <= 31; s++) {
        const uint32_t b0 = thr_keys[0] >> s, b1 = thr_keys[n_thr - 1] >> s;
        const uint32_t nb = b1 - b0 + 1;
        if (nb > kMaxBuckets)
            continue;
        uint32_t w = 0, run = 0, cur = b0;
        for (uint32_t i = 0; i < n_thr; i++) {
            uint32_t b = thr_keys[i] >> s;
            if (b != cur) {
                cur = b;
                run = 0;
            }
            run++;
            w = std::max(w, run);
        }
        *shift = s;
        *base = b0;
        *n_buckets = nb;
        *walk = w;
        return 1;
    }
    return 0;
}

/* ================================ metadata wire format =============================== */

namespace {
void put_record(uint8_t *&p, uint32_t id, const void *payload, uint32_t size)
{
    memcpy(p, &id, 4);
    memcpy(p + 4, &size, 4);
    memcpy(p + 8, payload, size);
    p += 8 + size;
}
} // namespace

extern "C" int lumacu_metadata_pack(const lumacu_metadata *m, const float *lut, uint32_t lut_len, uint8_t *blob, size_t cap,
                                    size_t *used)
try {
    if (!m || !lut || lut_len < 1 || !used)
        return fail(nullptr, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_metadata_pack: NULL argument");
    const uint32_t table_bytes = (lut_len - 1) * 4u; /* getSize() = maxVal floats, one short of the table */
    const size_t need = 7 * 8 + 4 + 4 + 4 + 4 + (size_t)table_bytes + 4 + 8;
    *used = need;
    if (!blob || cap < need)
        return fail(nullptr, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_metadata_pack: need %zu bytes", need);
    uint8_t *p = blob;
    put_record(p, 430, &m->ptf_bit_depth, 4);
    put_record(p, 431, &m->color_bit_depth, 4);
    put_record(p, 432, &m->ptf, 4);
    put_record(p, 433, &m->color_space, 4);
    put_record(p, 434, lut, table_bytes);
    put_record(p, 435, &m->pre_scaling, 4);
    const float range[2] = {m->max_lum, m->min_lum};
    put_record(p, 436, range, 8);
    return LUMACU_OK;
}
LUMACU_CATCH(nullptr)

extern "C" int lumacu_metadata_unpack(const uint8_t *blob, size_t size, lumacu_metadata *m, float *lut_out, size_t lut_cap,
                                      uint32_t *lut_len)
try {
    if (!blob || !m || !lut_out || !lut_len)
        return fail(nullptr, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_metadata_unpack: NULL argument");
    /* defaults of LumaDecoderParams (include/luma/luma_decoder.h:60-70,112-120) for the optional records */
    lumacu_metadata r{11, 8, LUMACU_PTF_PSI, LUMACU_CS_LUV, 1.0f, 1e4f, 0.005f};
    unsigned have = 0;
    const uint8_t *table = nullptr;
    uint32_t table_bytes = 0;
    for (size_t off = 0; off + 8 <= size;) {
        uint32_t id, n;
        memcpy(&id, blob + off, 4);
        memcpy(&n, blob + off + 4, 4);
        if (off + 8 + (size_t)n > size)
            return fail(nullptr, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_metadata_unpack: record %u overruns the blob", id);
        const uint8_t *pl = blob + off + 8;
        off += 8 + (size_t)n;
        const bool scalar_ok = n >= 4;
        switch (id) {
        case 430: if (scalar_ok) { memcpy(&r.ptf_bit_depth, pl, 4); have |= 1; } break;
        case 431: if (scalar_ok) { memcpy(&r.color_bit_depth, pl, 4); have |= 2; } break;
        case 432: if (scalar_ok) { memcpy(&r.ptf, pl, 4); have |= 4; } break;
        case 433: if (scalar_ok) { memcpy(&r.color_space, pl, 4); have |= 8; } break;
        case 434: table = pl; table_bytes = n; have |= 16; break;
        case 435: if (scalar_ok) memcpy(&r.pre_scaling, pl, 4); break;
        case 436: if (n >= 8) { memcpy(&r.max_lum, pl, 4); memcpy(&r.min_lum, pl + 4, 4); } break;
        default: break;
        }
    }
    if (have != 31u)
        return fail(nullptr, LUMACU_ERR_INVALID_ARGUMENT, "Failed to locate Luma HDRv meta data"); /* src/luma_decoder.cpp:117 */
    if (r.ptf_bit_depth < 1 || r.ptf_bit_depth > 16)
        return fail(nullptr, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_metadata_unpack: PTF bit depth %u", r.ptf_bit_depth);
    const size_t n_lut = (size_t)1 << r.ptf_bit_depth;
    if (lut_cap < n_lut)
        return fail(nullptr, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_metadata_unpack: table needs %zu floats", n_lut);
    /* setQuantizer(...) then memcpy(getMapping(), mapping, mapping_size) (src/luma_decoder.cpp:121-122), without
     * writing past the table */
    const int rc = lumacu_build_lut(r.ptf, r.ptf_bit_depth, r.max_lum, r.min_lum, lut_out, lut_cap);
    if (rc)
        return rc;
    memcpy(lut_out, table, std::min<size_t>(table_bytes, n_lut * 4));
    *m = r;
    *lut_len = (uint32_t)n_lut;
    return LUMACU_OK;
}
LUMACU_CATCH(nullptr)

/* ================================ context ============================================ */

extern "C" int lumacu_create(int device, lumacu_ctx **out)
try {
    if (!out)
        return fail(nullptr, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, LUMACU_ERR_NO_DEVICE, "lumacu_create: no CUDA device (%s)",
                    e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n)
        return fail(nullptr, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_create: device %d out of range [0,%d)", device, n);
    lumacu_ctx *ctx = new (std::nothrow) lumacu_ctx();
    if (!ctx)
        return fail(nullptr, LUMACU_ERR_OUT_OF_MEMORY, "lumacu_create: host allocation failed");
    ctx->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, LUMACU_ERR_CUDA, "lumacu_create: %s", cudaGetErrorString(e));
    }
    if (prop.major != 10) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return fail(nullptr, LUMACU_ERR_UNSUPPORTED, "lumacu_create: device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
    }
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking)) != cudaSuccess) {
        lumacu_destroy(ctx);
        return fail(nullptr, LUMACU_ERR_CUDA, "lumacu_create: %s", cudaGetErrorString(e));
    }
    *out = ctx;
    return LUMACU_OK;
}
LUMACU_CATCH(nullptr)

extern "C" int lumacu_destroy(lumacu_ctx *ctx)
{
    if (!ctx)
        return LUMACU_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->s_out)
        cudaStreamSynchronize(ctx->s_out); /* an asynchronous call the caller never waited for */
    for (auto &kv : ctx->stats_ws)
        for (DeviceBuffer *b : {&kv.second.partial, &kv.second.counter})
            if (b->p)
                cudaFree(b->p);
    for (DeviceBuffer *b : {&ctx->d_tables, &ctx->d_pq, &ctx->d_vd, &ctx->d_pqh, &ctx->d_rgb, &ctx->d_planes, &ctx->d_stats,
                            &ctx->d_aux, &ctx->d_half})
        if (b->p)
            cudaFree(b->p);
    if (ctx->h_pin)
        cudaFreeHost(ctx->h_pin);
    if (ctx->ev_input)
        cudaEventDestroy(ctx->ev_input);
    for (cudaEvent_t ev : ctx->ev_in)
        cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->ev_k)
        cudaEventDestroy(ev);
    if (ctx->s_in)
        cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out)
        cudaStreamDestroy(ctx->s_out);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return LUMACU_OK;
}

extern "C" const char *lumacu_last_error(const lumacu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
extern "C" int lumacu_device(const lumacu_ctx *ctx) { return ctx ? ctx->device : -1; }
extern "C" uint64_t lumacu_launch_count(const lumacu_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int lumacu_synchronize(lumacu_ctx *ctx)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return finish_pending(ctx);
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

extern "C" void *lumacu_stream(lumacu_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int lumacu_host_register(void *p, size_t bytes)
try {
    if (!p || !bytes)
        return LUMACU_ERR_INVALID_ARGUMENT;
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return LUMACU_OK;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, LUMACU_ERR_CUDA, "cudaHostRegister(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return LUMACU_OK;
}
LUMACU_CATCH(nullptr)

extern "C" int lumacu_host_unregister(void *p)
try {
    if (!p)
        return LUMACU_OK;
    if (cudaHostUnregister(p) != cudaSuccess)
        cudaGetLastError();
    return LUMACU_OK;
}
LUMACU_CATCH(nullptr)

extern "C" int lumacu_host_alloc(size_t bytes, void **out)
try {
    if (!out)
        return LUMACU_ERR_INVALID_ARGUMENT;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        *out = nullptr;
        cudaGetLastError();
        return fail(nullptr, LUMACU_ERR_OUT_OF_MEMORY, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return LUMACU_OK;
}
LUMACU_CATCH(nullptr)

extern "C" int lumacu_host_free(void *p)
try {
    if (p)
        cudaFreeHost(p);
    return LUMACU_OK;
}
LUMACU_CATCH(nullptr)

/* ================================ quantizer ========================================== */

/* CS_YCBCR: the device-built tables of the tuned kernels (luma_pq_tables.cuh) for the quantizer `q` (whose LUT, thresholds
 * and search tables are already in device memory): the two PQ tables depend on Lmax only and are kept across
 * quantizers; the v-keyed luma search table is rebuilt every time.  Fills q.pqd / q.pqe / q.vdtab (NULL = the kernels
 * evaluate that piece per pixel).  ~1 ms. */
static int build_ycbcr_tables(lumacu_ctx *ctx, QuantDev &q, float max_lum)
{
    q.pqd = nullptr, q.pqe = nullptr, q.vdtab = nullptr;
    q.lmax_rc = 0.0f;
    ctx->pqh_valid = false; /* rebuilt by the next CS_YCBCR encode launch */
    /* val / Lmax of PQ encode as a multiplication by the reciprocal plus one FMA correction -- only if that equals the
     * IEEE quotient for every operand the path can produce, which the device checks by exhaustion (1.35e9 operands, < 1 ms;
     * result cached per Lmax) */
    if (max_lum > 0.0f && std::isfinite(max_lum)) {
        if (!ctx->lmax_div_checked || memcmp(&ctx->lmax_div_lmax, &max_lum, sizeof(float)) != 0) {
            ctx->lmax_div_checked = false;
            if (reserve(ctx, ctx->d_vd, kVdTabBytes + 16) == LUMACU_OK) {
                uint32_t *flag = (uint32_t *)((unsigned char *)ctx->d_vd.p + kVdTabBytes);
                const float rc = 1.0f / max_lum;
                uint32_t bad = 1;
                CU_TRY(ctx, cudaMemsetAsync(flag, 0, 4, ctx->stream));
                launch_check_lmax_division((unsigned)ctx->sm_count * 8u, ctx->stream, max_lum, rc, flag);
                CU_TRY(ctx, cudaGetLastError());
                CU_TRY(ctx, cudaMemcpyAsync(&bad, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
                CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
                ctx->launches++;
                ctx->lmax_div_lmax = max_lum;
                ctx->lmax_div_rc = bad ? 0.0f : rc;
                ctx->lmax_div_checked = true;
            } else {
                ctx->err.clear();
            }
        }
        if (ctx->lmax_div_checked)
            q.lmax_rc = ctx->lmax_div_rc;
    }
    if (!ctx->pq_valid || memcmp(&ctx->pq_lmax, &max_lum, sizeof(float)) != 0) {
        ctx->pq_valid = false;
        if (reserve(ctx, ctx->d_pq, kPqTabPqdBytes + kPqTabPqeBytes) != LUMACU_OK) {
            ctx->err.clear(); /* no room for the tables: PQ is evaluated per pixel instead */
            return LUMACU_OK;
        }
        launch_build_pq_tables((unsigned)ctx->sm_count * 8u, ctx->stream, ctx->d_pq.p,
                               (float *)((unsigned char *)ctx->d_pq.p + kPqTabPqdBytes), max_lum);
        CU_TRY(ctx, cudaGetLastError());
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->launches += 2;
        ctx->pq_lmax = max_lum;
        ctx->pq_valid = true;
    }
    q.pqd = (const uint4 *)ctx->d_pq.p;
    q.pqe = (const float *)((unsigned char *)ctx->d_pq.p + kPqTabPqdBytes);
    if (reserve(ctx, ctx->d_vd, kVdTabBytes + 16) != LUMACU_OK) {
        ctx->err.clear();
        return LUMACU_OK;
    }
    uint32_t *flag = (uint32_t *)((unsigned char *)ctx->d_vd.p + kVdTabBytes);
    CU_TRY(ctx, cudaMemsetAsync(flag, 0, 4, ctx->stream));
    launch_build_vdtab(ctx->stream, q, (uint32_t *)ctx->d_vd.p, flag, max_lum);
    CU_TRY(ctx, cudaGetLastError());
    uint32_t bad = 1;
    CU_TRY(ctx, cudaMemcpyAsync(&bad, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->launches++;
    if (!bad)
        q.vdtab = (const uint32_t *)ctx->d_vd.p;
    return LUMACU_OK;
}

extern "C" int lumacu_set_quantizer(lumacu_ctx *ctx, const float *lut, uint32_t lut_len, uint32_t max_val_color,
                                    int color_space, float max_lum)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!lut || lut_len < 1 || lut_len > 65536u)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_set_quantizer: lut_len %u not in [1,65536]", lut_len);
    if (color_space < 0 || color_space > 3)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_set_quantizer: unknown colour space %d", color_space);
    /* 2^colorBitDepth - 1 with a colour depth the 16-bit container can hold (attachment 431 is untrusted input) */
    if (max_val_color < 1 || max_val_color > 65535u)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_set_quantizer: max_val_color %u not in [1,65535]", max_val_color);
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (int rcp = finish_pending(ctx))
        return rcp;

    const uint32_t max_val = lut_len - 1;
    std::vector<uint32_t> thr(max_val ? max_val : 1);
    std::vector<uint16_t> bucket;
    uint32_t shift = 0, base = 0, nb = 0, walk = 0;
    int mode = SEARCH_LITERAL;
    if (max_val >= 1 && lumacu_derive_thresholds(lut, lut_len, thr.data())) {
        mode = SEARCH_BINARY;
        if (lumacu_plan_buckets(thr.data(), max_val, &shift, &base, &nb, &walk) && walk <= kMaxWalk) {
            mode = SEARCH_BUCKET;
            bucket.assign(nb + (nb & 1u) + 2u, 0);
            /* head[b] = number of thresholds in lower buckets */
            uint32_t j = 0;
            for (uint32_t b = 0; b < nb; b++) {
                while (j < max_val && (thr[j] >> shift) - base < b)
                    j++;
                bucket[b] = (uint16_t)j;
            }
        }
    }
    const uint32_t pad = std::max(walk, 4u); /* the tuned kernels read up to 4 thresholds from any bucket head */
    const uint32_t thr_count = (mode == SEARCH_LITERAL) ? 0u : max_val + pad;
    if (mode != SEARCH_LITERAL)
        thr.resize(thr_count, 0xFFFFFFFFu);

    /* chroma code -> first chroma term of the inverse transform, with the reference's own expressions
     * (dequantize: src/luma_quantizer.cpp:261; Lu'v': :406-407).  Host float arithmetic, no contraction. */
    std::vector<float> ctab;
    if (color_space == CS_LUV || color_space == CS_YCBCR) {
        ctab.resize((size_t)max_val_color + 1);
        const float mc = (float)max_val_color;
        for (uint32_t i = 0; i <= max_val_color; i++) {
            volatile float c = (float)i / mc;
            c = (c < 1e-10f) ? 1e-10f : c; /* std::max(val/maxC, 1e-10f) */
            if (color_space == CS_LUV) {
                volatile float t = c * 255.0f;
                c = t / 410.0f;
            }
            ctab[i] = c;
        }
    }

    /* Direct search table (tuned encode kernels, colour spaces whose searched values are clamped to
     * [1e-4, 1e8]: Lu'v' luminance, XYZ).  For such a value the raw float bits `key` are an ordered key.  Cut the
     * key range into buckets of 2^S keys holding at most ONE decision threshold each, and store per bucket
     *     entry = (c0 << 16) + 0x10000 - thr_low - (bucket << S)        (mod 2^32)
     * with c0 = number of thresholds below the bucket and thr_low = offset of the bucket's threshold inside it
     * (2^S when it has none).  Then entry + key = (c0 << 16) + 0x10000 + (key_low - thr_low), whose upper half is
     * c0 + 1 when key_low >= thr_low and c0 otherwise: code = (entry + key) >> 16 -- one shared-memory read and one
     * add per sample.  NaN (canonical 0x7fffffff after the clamp) and anything an ulp above 1e8 are folded onto
     * the 1e8 bucket by an unsigned min on the device, which needs every threshold to be <= 1e8.  The table starts
     * one bucket below 1e-4 because a 2x2 mean of four 1e-4 samples may round an ulp below 1e-4. */
    std::vector<uint32_t> dtab;
    uint32_t d_shift = 0, d_lo = 0, d_lo_key = 0, d_hi_key = 0;
    /* key space of the table: raw float bits where the searched values are positive (Lu'v' Y, XYZ), otherwise the
     * ordered keys the thresholds are already stored in (RGB, YCbCr luminance: any sign, any NaN) */
    const bool raw_keys = (color_space == CS_LUV || color_space == CS_XYZ);
    const uint32_t flip = raw_keys ? 0x80000000u : 0u;
    if (mode == SEARCH_BUCKET && thr[0] > 0x80000000u && max_val <= 32767u) {
        const uint32_t k_lo = f2u(1e-4f), k_hi = f2u(1e8f);
        const uint32_t t_first = thr[0] ^ flip, t_last = thr[max_val - 1] ^ flip;
        for (uint32_t S = 16; S >= 12 && dtab.empty(); S--) {
            bool ok = true;
            for (uint32_t j = 1; j < max_val && ok; j++)
                ok = ((thr[j] ^ flip) >> S) != ((thr[j - 1] ^ flip) >> S);
            if (!ok)
                continue; /* two thresholds share a bucket: finer buckets */
            /* (A) raw keys only: the whole clamp range [1e-4, 1e8] (one bucket more below: a 2x2 mean of four 1e-4
             *     samples may round an ulp below 1e-4): the device only needs the upper clamp;
             * (B) just the thresholds' own range plus a bucket either side: smaller, needs both clamps on the device
             *     (LUTs such as LOG-12 whose buckets are too fine for (A) to fit in shared memory, and every
             *     ordered-key table). */
            /* (A) ends one bucket above the last threshold: everything beyond it has the code max_val anyway, and the
             * device's upper key clamp (needed for NaN) folds it onto that bucket -- for PQ-11 with Lmax = 1e4 this
             * drops the 13 binades up to 1e8 (40.8 -> 27.6 KB staged per block) */
            uint32_t lo = (k_lo >> S) - 1u, hi = std::min(k_hi >> S, (t_last >> S) + 1u);
            const bool trimmed = hi < (k_hi >> S);
            bool clamp_lo = false;
            if (!raw_keys || t_last > k_hi || (size_t)(hi - lo + 1u) * 4 > 48 * 1024) {
                lo = (t_first >> S) - 1u;
                hi = (t_last >> S) + 1u;
                clamp_lo = true;
                if ((size_t)(hi - lo + 1u) * 4 > 48 * 1024)
                    break; /* finer buckets only get bigger */
            }
            const uint32_t n = hi - lo + 1u;
            dtab.resize(n);
            uint32_t j = 0; /* thresholds below the current bucket */
            for (uint32_t b = 0; b < n; b++) {
                const uint32_t kb = lo + b;
                while (j < max_val && ((thr[j] ^ flip) >> S) < kb)
                    j++;
                uint32_t thr_low = 1u << S;
                if (j < max_val && ((thr[j] ^ flip) >> S) == kb)
                    thr_low = (thr[j] ^ flip) - (kb << S);
                dtab[b] = (j << 16) + 0x10000u - thr_low - (kb << S);
            }
            d_shift = S;
            d_lo = lo;
            d_lo_key = clamp_lo ? (lo << S) : 0u;
            d_hi_key = (clamp_lo || trimmed) ? (((hi + 1u) << S) - 1u) : k_hi;
            dtab.resize((dtab.size() + 3) & ~(size_t)3, 0u); /* the kernels stage it with 128-bit loads */
        }
    }

    /* Wide LUTs (PQ-12, ...): no bucket width puts at most ONE threshold in a bucket within the shared-memory budget, but
     * at most TWO works.  Same idea with 64-bit entries {A, B}:
     *     A = (c0 << 16) + 0x10000 - t1 - (bucket << S),   B = 0x10000 - t2 - (bucket << S)
     * (t1 <= t2 = offsets of the bucket's thresholds, 2^S when absent): code = ((A + key) >> 16) + ((B + key) >> 16) --
     * one 64-bit shared-memory read and three adds per sample instead of a bucket head plus a four-threshold walk.
     * Always used with both key clamps (luma_fast.cuh WALK -3). */
    uint32_t d_double = 0;
    if (dtab.empty() && mode == SEARCH_BUCKET && thr[0] > 0x80000000u && max_val >= 2u && max_val <= 32767u) {
        const uint32_t k_lo = f2u(1e-4f);
        const uint32_t t_first = thr[0] ^ flip, t_last = thr[max_val - 1] ^ flip;
        for (uint32_t S = 16; S >= 12 && dtab.empty(); S--) {
            bool ok = true;
            for (uint32_t j = 2; j < max_val && ok; j++)
                ok = ((thr[j] ^ flip) >> S) != ((thr[j - 2] ^ flip) >> S);
            if (!ok)
                continue;
            uint32_t lo = (t_first >> S) - 1u, hi = (t_last >> S) + 1u;
            if (raw_keys) /* searched values are clamped to >= 1e-4 (less an ulp): nothing below is ever asked.  The top
                           * stays one bucket above the LAST threshold even beyond 1e8: NaN is folded there by the
                           * upper key clamp and must come out as max_val */
                lo = std::max(lo, (k_lo >> S) - 1u);
            if (hi < lo || (size_t)(hi - lo + 1u) * 8 > 56 * 1024)
                break; /* finer buckets only get bigger */
            const uint32_t n = hi - lo + 1u;
            dtab.resize((size_t)n * 2);
            uint32_t j = 0;
            for (uint32_t b = 0; b < n; b++) {
                const uint32_t kb = lo + b;
                while (j < max_val && ((thr[j] ^ flip) >> S) < kb)
                    j++;
                uint32_t t1 = 1u << S, t2 = 1u << S;
                if (j < max_val && ((thr[j] ^ flip) >> S) == kb)
                    t1 = (thr[j] ^ flip) - (kb << S);
                if (j + 1 < max_val && ((thr[j + 1] ^ flip) >> S) == kb)
                    t2 = (thr[j + 1] ^ flip) - (kb << S);
                dtab[2 * b] = (j << 16) + 0x10000u - t1 - (kb << S);
                dtab[2 * b + 1] = 0x10000u - t2 - (kb << S);
            }
            d_double = 1;
            d_shift = S;
            d_lo = lo;
            d_lo_key = lo << S;
            d_hi_key = ((hi + 1u) << S) - 1u;
            dtab.resize((dtab.size() + 3) & ~(size_t)3, 0u);
        }
    }

    /* Very wide LUTs (13-16 bits, e.g. `lumaenc -pb 16`, lumaenc.cpp:132): neither shared-memory table fits and the bucket
     * walk would be too long (binary search mode).  The one-threshold-per-bucket table still exists -- it is just large
     * (PQ-16: 2^10 floats per bucket, 0.9 MB); it stays in global memory, L1 / L2 resident, and the tuned kernels read it
     * with one read-only load per sample (luma_fast.cuh WALK -4). */
    uint32_t d_global = 0;
    if (dtab.empty() && mode != SEARCH_LITERAL && max_val >= 1u && thr[0] > 0x80000000u) {
        const uint32_t k_lo = f2u(1e-4f);
        const uint32_t t_first = thr[0] ^ flip, t_last = thr[max_val - 1] ^ flip;
        for (uint32_t S = 16; S >= 6 && dtab.empty(); S--) {
            bool ok = true;
            for (uint32_t j = 1; j < max_val && ok; j++)
                ok = ((thr[j] ^ flip) >> S) != ((thr[j - 1] ^ flip) >> S);
            if (!ok)
                continue;
            uint32_t lo = (t_first >> S) - 1u, hi = (t_last >> S) + 1u;
            if (raw_keys)
                lo = std::max(lo, (k_lo >> S) - 1u);
            if (hi < lo || (size_t)(hi - lo + 1u) * 4 > ((size_t)8 << 20))
                break;
            const uint32_t n = hi - lo + 1u;
            dtab.resize(n);
            uint32_t j = 0;
            for (uint32_t b = 0; b < n; b++) {
                const uint32_t kb = lo + b;
                while (j < max_val && ((thr[j] ^ flip) >> S) < kb)
                    j++;
                uint32_t thr_low = 1u << S;
                if (j < max_val && ((thr[j] ^ flip) >> S) == kb)
                    thr_low = (thr[j] ^ flip) - (kb << S);
                dtab[b] = (j << 16) + 0x10000u - thr_low - (kb << S);
            }
            d_global = 1;
            d_shift = S;
            d_lo = lo;
            d_lo_key = lo << S;
            d_hi_key = ((hi + 1u) << S) - 1u;
            dtab.resize((dtab.size() + 3) & ~(size_t)3, 0u);
        }
    }

    /* CS_YCBCR decode: per-code y' = ((255 PQenc(lut[code])) - 16) / 219 with the reference's own float expression
     * (src/luma_quantizer.cpp:447-448, 493-494) and the host libm */
    std::vector<float> ylut;
    if (color_space == CS_YCBCR) {
        ylut.resize(lut_len);
        for (uint32_t i = 0; i < lut_len; i++) {
            volatile float y = host_pq_encode(lut[i], max_lum);
            volatile float t = 255.0f * y;
            t = t - 16.0f;
            ylut[i] = t / 219.0f;
        }
    }

    /* one device allocation: lut | thr | bucket | ctab | dtab | ylut */
    const size_t off_thr = ((size_t)lut_len * 4 + 15) & ~(size_t)15;
    const size_t off_bucket = (off_thr + (size_t)thr_count * 4 + 15) & ~(size_t)15;
    const size_t off_ctab = (off_bucket + bucket.size() * 2 + 15) & ~(size_t)15;
    const size_t off_dtab = (off_ctab + ctab.size() * 4 + 15) & ~(size_t)15;
    const size_t off_ylut = (off_dtab + dtab.size() * 4 + 15) & ~(size_t)15;
    const size_t total = off_ylut + ylut.size() * 4 + 16;
    int rc = reserve(ctx, ctx->d_tables, total);
    if (rc)
        return rc;
    /* previous launches (on the context's stream or on a caller's) may still read the old tables */
    CU_TRY(ctx, cudaDeviceSynchronize());
    unsigned char *d = (unsigned char *)ctx->d_tables.p;
    CU_TRY(ctx, cudaMemcpy(d, lut, (size_t)lut_len * 4, cudaMemcpyHostToDevice));
    if (thr_count)
        CU_TRY(ctx, cudaMemcpy(d + off_thr, thr.data(), (size_t)thr_count * 4, cudaMemcpyHostToDevice));
    if (!bucket.empty())
        CU_TRY(ctx, cudaMemcpy(d + off_bucket, bucket.data(), bucket.size() * 2, cudaMemcpyHostToDevice));
    if (!ctab.empty())
        CU_TRY(ctx, cudaMemcpy(d + off_ctab, ctab.data(), ctab.size() * 4, cudaMemcpyHostToDevice));
    if (!dtab.empty())
        CU_TRY(ctx, cudaMemcpy(d + off_dtab, dtab.data(), dtab.size() * 4, cudaMemcpyHostToDevice));
    if (!ylut.empty())
        CU_TRY(ctx, cudaMemcpy(d + off_ylut, ylut.data(), ylut.size() * 4, cudaMemcpyHostToDevice));

    QuantDev q{};
    q.lut = (const float *)d;
    q.thr = (const uint32_t *)(d + off_thr);
    q.bucket = (const uint16_t *)(d + off_bucket);
    q.ctab = ctab.empty() ? nullptr : (const float *)(d + off_ctab);
    q.dtab = dtab.empty() ? nullptr : (const uint32_t *)(d + off_dtab);
    q.d_shift = d_shift;
    q.d_lo = d_lo;
    q.d_n = (uint32_t)dtab.size();
    q.d_double = d_double;
    q.d_global = d_global;
    q.d_lo_key = d_lo_key;
    q.d_hi_key = d_hi_key;
    q.ylut = ylut.empty() ? nullptr : (const float *)(d + off_ylut);
    q.max_val = max_val;
    q.max_val_color = max_val_color;
    q.max_val_f = (float)max_val;
    q.max_val_color_f = (float)max_val_color;
    q.l_max = max_lum;
    q.search_mode = mode;
    q.shift = shift;
    q.base = base;
    q.nbm1 = nb ? nb - 1 : 0;
    q.walk = walk;
    q.thr_count = thr_count;
    size_t smem_enc = 0;
    if (mode != SEARCH_LITERAL) {
        smem_enc = (size_t)thr_count * 4 + (mode == SEARCH_BUCKET ? ((size_t)(nb + 2) / 2) * 4 : 0);
        if (smem_enc > kMaxSmemTables)
            smem_enc = 0;
    }
    q.smem_tables = smem_enc ? 1u : 0u;
    size_t smem_dec = (size_t)lut_len * 4;
    if (smem_dec > kMaxSmemLut)
        smem_dec = 0;
    q.smem_lut = smem_dec ? 1u : 0u;

    /* tuned kernels: bucket search out of shared memory (encode); LUT + chroma table in shared memory (decode) */
    size_t smem_dec_fast = ((size_t)lut_len + ctab.size()) * 4;
    if (smem_dec_fast > kMaxSmemLut)
        smem_dec_fast = 0;
    /* (the tuned search also wants every threshold to be a positive float: key > key(+0)) */
    ctx->fast_enc_ok = (mode == SEARCH_BUCKET && smem_enc != 0 && thr[0] > 0x80000000u) || d_global != 0;
    ctx->smem_dec_fast = smem_dec_fast;
    /* LUT too large for shared memory (14-16 bits) but the chroma table fits: tuned decode with the LUT read in place */
    ctx->dec_global_lut = smem_dec_fast == 0 && ctab.size() * 4 <= kMaxSmemLut;
    ctx->smem_dec_ctab = ctab.size() * 4;

    if (color_space == CS_YCBCR && (rc = build_ycbcr_tables(ctx, q, max_lum)) != LUMACU_OK)
        return rc;

    ctx->q = q;
    ctx->tables_bytes = total;
    ctx->max_lum = max_lum;
    ctx->smem_enc = smem_enc;
    ctx->smem_dec = smem_dec;
    ctx->color_space = color_space;
    ctx->h_lut.assign(lut, lut + lut_len);
    ctx->configured = true;
    return LUMACU_OK;
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

/* Single-process multi-GPU: every context receives ctxs[root]'s quantizer -- the host LUT and scalars, and the DEVICE
 * tables (LUT, decision thresholds, bucket heads, direct search table, chroma / y' tables) by one peer-to-peer copy per
 * destination (NVLink/NVSwitch when the devices are peers, staged by the driver otherwise), so that the host-side
 * derivation runs once and every GPU searches with bit-identical tables.  Replaces nothing in the reference (it has
 * one quantizer per process); it is the in-process form of the LUT broadcast SURVEY 8e asks for.  Processes on
 * different GPUs exchange lumacu_metadata_pack blobs instead (lumahdrv_b200/shard.py: NCCL broadcast). */
extern "C" int lumacu_broadcast_quantizer(lumacu_ctx *const ctxs[], int n, int root)
try {
    if (!ctxs || n < 1 || root < 0 || root >= n)
        return fail(nullptr, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_broadcast_quantizer: bad arguments");
    for (int i = 0; i < n; i++)
        if (!ctxs[i])
            return fail(nullptr, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_broadcast_quantizer: ctxs[%d] is NULL", i);
    lumacu_ctx *src = ctxs[root];
    if (!src->configured)
        return fail(src, LUMACU_ERR_NOT_CONFIGURED, "lumacu_broadcast_quantizer: the root context has no quantizer");
    const unsigned char *src_base = (const unsigned char *)src->d_tables.p;
    CU_TRY(src, cudaSetDevice(src->device));
    CU_TRY(src, cudaDeviceSynchronize()); /* the root's tables are complete and not being replaced */
    for (int i = 0; i < n; i++) {
        lumacu_ctx *dst = ctxs[i];
        if (dst == src)
            continue;
        CU_TRY(dst, cudaSetDevice(dst->device));
        int rc = finish_pending(dst);
        if (rc || (rc = reserve(dst, dst->d_tables, src->tables_bytes)))
            return rc;
        CU_TRY(dst, cudaDeviceSynchronize()); /* launches on the destination may still read its old tables */
        if (dst->device == src->device)
            CU_TRY(dst, cudaMemcpyAsync(dst->d_tables.p, src_base, src->tables_bytes, cudaMemcpyDeviceToDevice, dst->stream));
        else
            CU_TRY(dst, cudaMemcpyPeerAsync(dst->d_tables.p, dst->device, src_base, src->device, src->tables_bytes, dst->stream));
        CU_TRY(dst, cudaStreamSynchronize(dst->stream));
        /* same table layout, other base address */
        QuantDev q = src->q;
        unsigned char *dst_base = (unsigned char *)dst->d_tables.p;
        auto rebase = [&](const void *p) -> const void * {
            return p ? (const void *)(dst_base + ((const unsigned char *)p - src_base)) : nullptr;
        };
        q.lut = (const float *)rebase(q.lut);
        q.thr = (const uint32_t *)rebase(q.thr);
        q.bucket = (const uint16_t *)rebase(q.bucket);
        q.ctab = (const float *)rebase(q.ctab);
        q.dtab = (const uint32_t *)rebase(q.dtab);
        q.ylut = (const float *)rebase(q.ylut);
        q.pqd = nullptr, q.pqe = nullptr, q.vdtab = nullptr;
        /* built on the spot by the same deterministic kernels rather than copied (45 MB) */
        if (src->color_space == CS_YCBCR && (rc = build_ycbcr_tables(dst, q, src->max_lum)) != LUMACU_OK)
            return rc;
        dst->q = q;
        dst->tables_bytes = src->tables_bytes;
        dst->max_lum = src->max_lum;
        dst->color_space = src->color_space;
        dst->h_lut = src->h_lut;
        dst->smem_enc = src->smem_enc;
        dst->smem_dec = src->smem_dec;
        dst->smem_dec_fast = src->smem_dec_fast;
        dst->dec_global_lut = src->dec_global_lut;
        dst->smem_dec_ctab = src->smem_dec_ctab;
        dst->fast_enc_ok = src->fast_enc_ok;
        dst->configured = true;
    }
    return LUMACU_OK;
}
LUMACU_CATCH(nullptr)

/* Host copy of the quantizer a context holds (what lumacu_set_quantizer / lumacu_broadcast_quantizer left there):
 * lut_out receives up to `cap` floats; *lut_len, *max_val_color, *color_space, *max_lum may be NULL. */
extern "C" int lumacu_get_quantizer(const lumacu_ctx *ctx, float *lut_out, size_t cap, uint32_t *lut_len, uint32_t *max_val_color,
                                    int *color_space, float *max_lum)
try {
    if (!ctx || !ctx->configured)
        return LUMACU_ERR_NOT_CONFIGURED;
    if (lut_len)
        *lut_len = (uint32_t)ctx->h_lut.size();
    if (max_val_color)
        *max_val_color = ctx->q.max_val_color;
    if (color_space)
        *color_space = ctx->color_space;
    if (max_lum)
        *max_lum = ctx->max_lum;
    if (lut_out)
        memcpy(lut_out, ctx->h_lut.data(), std::min(cap, ctx->h_lut.size()) * sizeof(float));
    return LUMACU_OK;
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

extern "C" int lumacu_search_info(const lumacu_ctx *ctx, int *mode, uint32_t *n_buckets, uint32_t *shift, uint32_t *walk)
try {
    if (!ctx || !ctx->configured)
        return LUMACU_ERR_NOT_CONFIGURED;
    if (mode)
        *mode = ctx->q.search_mode;
    if (n_buckets)
        *n_buckets = ctx->q.search_mode == SEARCH_BUCKET ? ctx->q.nbm1 + 1 : 0;
    if (shift)
        *shift = ctx->q.shift;
    if (walk)
        *walk = ctx->q.walk;
    return LUMACU_OK;
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

extern "C" int lumacu_set_kernel_path(lumacu_ctx *ctx, int path)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (path != 0 && path != 1)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_set_kernel_path: path %d not in {0,1}", path);
    ctx->force_generic = (path == 1);
    return LUMACU_OK;
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

extern "C" int lumacu_set_pq_tables(lumacu_ctx *ctx, int enable)
{
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    ctx->pq_off = !enable;
    return LUMACU_OK;
}

extern "C" int lumacu_last_kernel_path(const lumacu_ctx *ctx) { return (ctx && ctx->last_fast) ? 1 : 0; }

extern "C" int lumacu_set_host_bands(lumacu_ctx *ctx, int bands)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (bands < 0 || bands > 64)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_set_host_bands: %d not in [0, 64]", bands);
    ctx->host_bands = bands;
    return LUMACU_OK;
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

extern "C" int lumacu_set_tuning(lumacu_ctx *ctx, int enc_variant, int dec_variant, int blocks_per_sm_cap)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (enc_variant < 0 || dec_variant < 0 || blocks_per_sm_cap < 0)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_set_tuning: negative argument");
    ctx->no_direct = enc_variant >= 1000 && enc_variant < 2000; /* 1000 + variant: bucket + threshold search instead of the direct table */
    ctx->global_direct = enc_variant >= 2000;                   /* 2000 + variant: direct table read from global memory (no staging) */
    ctx->enc_variant = enc_variant % 1000;
    ctx->dec_variant = dec_variant; /* 2000 + v: CS_YCBCR decode evaluates both greens of a pixel pair (luma_pq_tables.cuh) */
    ctx->grid_cap = blocks_per_sm_cap % 100;     /* blocks_per_sm_cap = cap + 100 * tiles_per_thread */
    ctx->grid_tpt = blocks_per_sm_cap / 100;
    return LUMACU_OK;
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

/* ================================ launches ============================================ */
namespace {

enc_fn pick_enc(int cs, bool sub, int bytes, bool vec)
{
    switch (cs) {
    case CS_LUV: return get_encode_generic_cs0(sub, bytes, vec);
    case CS_RGB: return get_encode_generic_cs1(sub, bytes, vec);
    case CS_YCBCR: return get_encode_generic_cs2(sub, bytes, vec);
    default: return get_encode_generic_cs3(sub, bytes, vec);
    }
}
dec_fn pick_dec(int cs, bool sub, int bytes, bool vec)
{
    switch (cs) {
    case CS_LUV: return get_decode_generic_cs0(sub, bytes, vec);
    case CS_RGB: return get_decode_generic_cs1(sub, bytes, vec);
    case CS_YCBCR: return get_decode_generic_cs2(sub, bytes, vec);
    default: return get_decode_generic_cs3(sub, bytes, vec);
    }
}
enc_fn pick_enc_fast(int cs, bool sub, int bytes, int walk, int variant, bool prescale)
{
    switch (cs) {
    case CS_LUV: return get_encode_fast_cs0(sub, bytes, walk, variant, prescale);
    case CS_RGB: return get_encode_fast_cs1(sub, bytes, walk, variant, prescale);
    case CS_YCBCR: return get_encode_fast_cs2(sub, bytes, walk, variant, prescale);
    default: return get_encode_fast_cs3(sub, bytes, walk, variant, prescale);
    }
}
dec_fn pick_dec_fast(int cs, bool sub, int bytes, int variant)
{
    switch (cs) {
    case CS_LUV: return get_decode_fast_cs0(sub, bytes, variant);
    case CS_RGB: return get_decode_fast_cs1(sub, bytes, variant);
    case CS_YCBCR: return get_decode_fast_cs2(sub, bytes, variant);
    default: return get_decode_fast_cs3(sub, bytes, variant);
    }
}

/* CUtensorMap over a batch of planar f32 frames for the tensor-map staged encode kernel: dims (fastest first)
 * {w, h, 3 planes, n_frames}, box {128, 2, 3, 1} -- one copy brings the 128-pixel x 2-row x 3-plane footprint
 * of a warp's 32 tiles.  cuTensorMapEncodeTiled is resolved through the runtime so libcuda is not linked. */
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_rgb_tensor_map(lumacu_ctx *ctx, const float *d_rgb, uint32_t w, uint32_t h, size_t plane_stride, size_t frame_stride,
                        uint32_t n_frames, unsigned char out[128])
{
    static encode_tiled_fn encode = nullptr;
    if (!encode) {
        void *fp = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr) != cudaSuccess || !fp) {
            (void)cudaGetLastError();
            return fail(ctx, LUMACU_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available");
        }
        encode = (encode_tiled_fn)fp;
    }
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    alignas(64) CUtensorMap tm;
    const cuuint64_t dims[4] = {w, h, 3, n_frames};
    const cuuint64_t strides[3] = {(cuuint64_t)w * 4, (cuuint64_t)plane_stride * 4, (cuuint64_t)frame_stride * 4};
    const cuuint32_t box[4] = {128, 2, 3, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)d_rgb, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(ctx, LUMACU_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    memcpy(out, &tm, 128);
    return LUMACU_OK;
}

/* Default variants of the tuned kernels, from the sweep on B200 (DESIGN.md "kernel tuning"). */
constexpr int kEncDefaultVariant = kEncVariantPlain;
constexpr int kDecDefaultVariant = kDecVariantPrefetch;

inline bool aligned(const void *p, size_t a) { return ((uintptr_t)p & (a - 1)) == 0; }

/* persistent grid: resident blocks on the whole chip, capped by the work */
/* multi-frame launches: tiles per thread a block is sized for (sweep on B200, 4K frames: encode 8/16/32 tiles ->
 * 654/642/641 us per 32 frames and 176/172/174 us per 8; decode 8/16/32 -> 645/649/687 us per 32 frames) */
constexpr uint32_t kEncTilesPerThread = 16, kDecTilesPerThread = 8;

int grid_for(lumacu_ctx *ctx, const void *fn, size_t smem, uint32_t ntiles, uint32_t n_frames, uint32_t tpt_default, uint32_t *gx)
{
    /* occupancy (and the opt-in for > 48 KB of dynamic shared memory) is queried once per kernel and size */
    int per_sm = 0;
    auto it = ctx->occupancy.find(fn);
    if (it != ctx->occupancy.end() && it->second.first == smem) {
        per_sm = it->second.second;
    } else {
        /* the 48 KB default limit covers dynamic + STATIC shared memory (mbarrier, powf tables, reduction scratch):
         * 48 KB of tables exactly (e.g. a 13-bit LUT + a 12-bit chroma table) already needs the opt-in */
        cudaFuncAttributes fa;
        CU_TRY(ctx, cudaFuncGetAttributes(&fa, fn));
        if (smem + fa.sharedSizeBytes > 48 * 1024)
            CU_TRY(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, smem));
        if (per_sm < 1)
            return fail(ctx, LUMACU_ERR_CUDA, "kernel does not fit on an SM (smem %zu)", smem);
        ctx->occupancy[fn] = std::make_pair(smem, per_sm);
    }
    if (ctx->grid_cap > 0)
        per_sm = std::min(per_sm, ctx->grid_cap);
    uint32_t resident = (uint32_t)per_sm * (uint32_t)ctx->sm_count;
    uint32_t need = (ntiles + kThreads - 1) / kThreads;
    uint32_t g = resident;
    if (n_frames > 1) {
        /* several frames per launch (grid.y = frame): blocks are scheduled frame-major, so the chip sweeps
         * the batch roughly one frame at a time (few concurrent DRAM streams -- measured faster than
         * spreading the resident slots over all frames at once), with >= ~8 tiles per thread so that the
         * table staging stays amortised */
        uint32_t per_frame = (resident + n_frames - 1) / n_frames;
        const uint32_t tpt = ctx->grid_tpt > 0 ? (uint32_t)ctx->grid_tpt : tpt_default;
        uint32_t coarse = (need + tpt - 1) / tpt;
        g = std::max(per_frame, std::min(coarse, resident));
    } else if (ctx->grid_tpt > 0) { /* tuning sweep: single-frame launches sized by tiles per thread (may exceed one wave) */
        g = std::max(1u, (need + (uint32_t)ctx->grid_tpt - 1) / (uint32_t)ctx->grid_tpt);
    }
    *gx = std::max(1u, std::min(g, need));
    return LUMACU_OK;
}

int check_profile(lumacu_ctx *ctx, int profile, uint32_t w, uint32_t h, bool encode)
{
    if (profile < 0 || profile > 3)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "profile %d not in [0,3]", profile);
    if (w == 0 || h == 0)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "frame size %ux%u is empty", w, h);
    const bool sub = (profile == 0 || profile == 2);
    /* the reference encoder refuses odd sizes (src/luma_encoder.cpp:118-119); its 4:2:0
     * decoder loop indexes as if the width were even (src/luma_decoder.cpp:229-234) */
    if ((encode || sub) && ((w | h) & 1u))
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "frame size %ux%u must be even", w, h);
    if ((uint64_t)w * h > 0x7fffffffull)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "frame size %ux%u too large", w, h);
    return LUMACU_OK;
}

} // namespace

static int encode_launch(lumacu_ctx *ctx, const float *d_rgb, float *d_rgb_out, uint32_t w, uint32_t h, int profile,
                         float pre_scaling, uint8_t *const d_planes[3], const int32_t strides[3], uint32_t n_frames,
                         size_t rgb_frame_stride, const size_t plane_frame_stride[3], lumacu_frame_stats *d_stats, void *stream,
                         const LaunchOpts &opt)
{
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!ctx->configured)
        return fail(ctx, LUMACU_ERR_NOT_CONFIGURED, "lumacu_encode_dev: lumacu_set_quantizer has not been called");
    if (!d_rgb || !d_planes || !strides || !d_planes[0] || !d_planes[1] || !d_planes[2])
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_encode_dev: NULL pointer argument");
    int rc = check_profile(ctx, profile, w, h, true);
    if (rc)
        return rc;
    if (n_frames == 0)
        return LUMACU_OK;
    if (n_frames > 65535u)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_encode_dev: at most 65535 frames per launch");
    const bool sub = (profile == 0 || profile == 2);
    const int bytes = profile > 1 ? 2 : 1;
    const uint32_t cw = sub ? (w + 1) >> 1 : w, chh = sub ? (h + 1) >> 1 : h;
    if (strides[0] < (int32_t)(w * bytes) || strides[1] < (int32_t)(cw * bytes) || strides[2] < (int32_t)(cw * bytes))
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_encode_dev: stride smaller than a row");
    if (n_frames > 1 && !plane_frame_stride)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_encode_dev: plane_frame_stride required for n_frames > 1");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;

    EncArgs a{};
    a.q = ctx->q;
    if (ctx->pq_off) /* every powf and every division evaluated the long way */
        a.q.pqd = nullptr, a.q.pqe = nullptr, a.q.vdtab = nullptr, a.q.lmax_rc = 0.0f;
    a.rgb = d_rgb;
    a.rgb_out = d_rgb_out;
    a.rgb_plane_stride = opt.rgb_plane_stride ? opt.rgb_plane_stride : (size_t)w * h;
    a.rgb_frame_stride = rgb_frame_stride ? rgb_frame_stride : (size_t)3 * w * h;
    a.out_plane_stride = a.rgb_plane_stride;
    a.out_frame_stride = a.rgb_frame_stride;
    a.w = w;
    a.h = h;
    a.sc = pre_scaling;
    a.prescale = (pre_scaling != 1.0f) ? 1 : 0;
    a.nz = make_float2(-0.0f, -0.0f);
    a.screen_k1 = (float)(0.25 * (double)ctx->q.max_val_color * 4.0 * 410.0 / 255.0);
    a.screen_k2 = (float)(0.25 * (double)ctx->q.max_val_color * 9.0 * 410.0 / 255.0);
    bool vec = (w % 4 == 0) && aligned(d_rgb, 16) && (a.rgb_frame_stride % 4 == 0) && (a.rgb_plane_stride % 4 == 0) &&
               (!d_rgb_out || aligned(d_rgb_out, 16));
    for (int p = 0; p < 3; p++) {
        a.plane[p] = d_planes[p];
        a.stride[p] = strides[p];
        a.plane_frame_stride[p] = plane_frame_stride ? plane_frame_stride[p] : (size_t)strides[p] * (p ? chh : h);
        const size_t al = (size_t)((p && sub) ? 2 : 4) * bytes; /* bytes stored per thread and row */
        vec = vec && aligned(d_planes[p], al) && (strides[p] % al == 0) && (a.plane_frame_stride[p] % al == 0);
    }
    enc_fn fn = nullptr;
    const bool small32 = (uint64_t)w * h * 4 < (1ull << 32) && (uint64_t)strides[0] * h < (1ull << 32);
    a.passthrough = opt.passthrough ? 1 : 0;
    size_t smem = ctx->smem_enc;
    if (vec && small32 && !d_rgb_out && ctx->fast_enc_ok && !ctx->force_generic && !opt.passthrough) {
        int variant = ctx->enc_variant ? ctx->enc_variant : kEncDefaultVariant;
        /* Lu'v' 4:2:0 with 8-bit chroma (the reference's default colour depth): screened chroma by default -- identical
         * results, a third fewer pipe cycles, 0.3 % of the tiles redone.  The share of redone tiles grows with the
         * chroma depth and with the samples per tile, and each redo is a scattered re-read plus the whole exact chain:
         * measured on 4K frames, screened vs exact chain -- 10-bit 4:2:0 5.75 vs 6.11 TB/s, 12-bit 4:2:0 4.6 vs 6.2, and a
         * 4:4:4 build (16 chroma samples per tile) 5.1 vs 6.2 at 8 bits -- so everything else keeps the exact chain
         * (lumacu_set_tuning(67) still selects the screened kernel for any 4:2:0 Lu'v' quantizer; same bits). */
        if (!ctx->enc_variant && ctx->color_space == CS_LUV && sub && ctx->q.max_val_color <= 255u)
            variant = kEncVariantScreened;
        const int pf = variant / 10;
        const bool staged = (pf == 8);
        if (staged && (w % 128u) != 0) /* tensor-map staging: a warp must not straddle image rows */
            variant = kEncVariantPlain;
        if (variant / 10 == 8 &&
            make_rgb_tensor_map(ctx, d_rgb, w, h, a.rgb_plane_stride, a.rgb_frame_stride, n_frames, a.rgb_tmap) != LUMACU_OK)
            variant = kEncVariantPlain;
        /* walk 0 = direct search table (one LDS per sample); otherwise bucket heads + <= walk threshold compares */
        /* the 64-bit wide-LUT table (up to 56 KB: three resident blocks instead of four) pays off for the screened kernel,
         * which is latency-bound (PQ-12 4:2:0: 6.2 vs 5.4 TB/s), not for the exact-chain kernels, which are pipe-bound and
         * want the occupancy (PQ-12 4:4:4: 5.0 vs 5.65 TB/s with the bucket + threshold walk) */
        const bool direct = ctx->q.dtab && !ctx->no_direct && (!ctx->q.d_double || variant == kEncVariantScreened);
        /* -1: direct table that needs the lower clamp too; -3: 64-bit entries, two thresholds per bucket; -4: table in
         * global memory */
        const int walk_direct = (ctx->q.d_global || (ctx->global_direct && !ctx->q.d_double)) ? -4 : ctx->q.d_double ? -3 : (ctx->q.d_lo_key ? -1 : 0);
        int walk = direct ? walk_direct : (int)ctx->q.walk;
        /* CS_YCBCR without statistics: plane 0 is searched by v in the v-keyed table (-2) */
        const bool v_keyed = ctx->color_space == CS_YCBCR && a.q.vdtab && !d_stats && !ctx->no_direct && variant == kEncVariantPlain;
        if (v_keyed)
            walk = -2;
        /* the bucket + threshold walk needs its tables in shared memory (fast_enc_ok may rest on the global table alone) */
        const bool walk_ok = ctx->q.search_mode == SEARCH_BUCKET && ctx->smem_enc != 0;
        fn = (walk > 0 && !walk_ok) ? nullptr : pick_enc_fast(ctx->color_space, sub, bytes, walk, variant, a.prescale != 0);
        if (!fn && direct && walk_ok) { /* tuning variants exist for one search flavour only */
            walk = (int)ctx->q.walk;
            fn = pick_enc_fast(ctx->color_space, sub, bytes, walk, variant, a.prescale != 0);
        }
        if (!fn && variant != kEncVariantPlain) { /* e.g. no screened instantiation for this search flavour */
            variant = kEncVariantPlain;
            walk = direct ? walk_direct : (int)ctx->q.walk;
            fn = (walk > 0 && !walk_ok) ? nullptr : pick_enc_fast(ctx->color_space, sub, bytes, walk, variant, a.prescale != 0);
        }
        if (fn && walk <= 0)
            smem = walk == -2 ? kVdTabBytes : walk == -4 ? 0 : (size_t)ctx->q.d_n * 4; /* d_n is a multiple of 4 entries */
        if (fn && staged && variant != kEncVariantPlain)
            smem += kEncStagedSmemBytes;
    }
    ctx->last_fast = fn != nullptr;
    if (fn && ctx->color_space == CS_YCBCR && a.q.pqe) {
        /* half-float input table for this launch's preScaling (EXR-sourced frames: no powf left on the forward path);
         * built on the launch stream right before the kernel that reads it */
        if (!ctx->pqh_valid || memcmp(&ctx->pqh_sc, &pre_scaling, 4) != 0 || memcmp(&ctx->pqh_lmax, &ctx->q.l_max, 4) != 0) {
            ctx->pqh_valid = false;
            if (reserve(ctx, ctx->d_pqh, 65536 * sizeof(float)) == LUMACU_OK) {
                launch_build_pqh(st, a.q, (float *)ctx->d_pqh.p, pre_scaling, a.prescale, ctx->q.l_max);
                CU_TRY(ctx, cudaGetLastError());
                CU_TRY(ctx, cudaStreamSynchronize(st)); /* rare (preScaling / Lmax changed); later launches may use another stream */
                ctx->launches++;
                ctx->pqh_sc = pre_scaling;
                ctx->pqh_lmax = ctx->q.l_max;
                ctx->pqh_valid = true;
            } else {
                ctx->err.clear();
            }
        }
        a.pqh = ctx->pqh_valid ? (const float *)ctx->d_pqh.p : nullptr;
    }
    if (!fn)
        fn = pick_enc(ctx->color_space, sub, bytes, vec);
    const uint32_t ntiles = ((w + 3) / 4) * ((h + 1) / 2);
    uint32_t gx = 1;
    rc = grid_for(ctx, (const void *)fn, smem, ntiles, n_frames, kEncTilesPerThread, &gx);
    if (rc)
        return rc;
    if (d_stats) {
        lumacu_ctx::StatsWs &ws = ctx->stats_ws[st];
        rc = reserve(ctx, ws.partial, (size_t)n_frames * gx * sizeof(StatsPartial));
        if (rc)
            return rc;
        if ((size_t)n_frames * 4 > ws.counter.cap) {
            rc = reserve(ctx, ws.counter, std::max<size_t>((size_t)n_frames * 4, 4096));
            if (rc)
                return rc;
            CU_TRY(ctx, cudaMemsetAsync(ws.counter.p, 0, ws.counter.cap, st));
        }
        a.partial = (StatsPartial *)ws.partial.p;
        a.counter = (uint32_t *)ws.counter.p;
        a.stats = (FrameStatsDev *)d_stats;
    }
    fn<<<dim3(gx, n_frames), kThreads, smem, st>>>(a);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return LUMACU_OK;
}

extern "C" int lumacu_encode_dev(lumacu_ctx *ctx, const float *d_rgb, float *d_rgb_out, uint32_t w, uint32_t h,
                                 int profile, float pre_scaling, uint8_t *const d_planes[3], const int32_t strides[3],
                                 uint32_t n_frames, size_t rgb_frame_stride, const size_t plane_frame_stride[3],
                                 lumacu_frame_stats *d_stats, void *stream)
try {
    return encode_launch(ctx, d_rgb, d_rgb_out, w, h, profile, pre_scaling, d_planes, strides, n_frames, rgb_frame_stride,
                         plane_frame_stride, d_stats, stream, LaunchOpts());
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

static int decode_launch(lumacu_ctx *ctx, const uint8_t *const d_planes[3], const int32_t strides[3], uint32_t w, uint32_t h,
                         int profile, float pre_scaling, float *d_rgb, uint32_t n_frames, size_t rgb_frame_stride,
                         const size_t plane_frame_stride[3], void *stream, const LaunchOpts &opt)
{
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!ctx->configured)
        return fail(ctx, LUMACU_ERR_NOT_CONFIGURED, "lumacu_decode_dev: lumacu_set_quantizer has not been called");
    if ((!d_rgb && !opt.display) || !d_planes || !strides || !d_planes[0] || !d_planes[1] || !d_planes[2])
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_decode_dev: NULL pointer argument");
    int rc = check_profile(ctx, profile, w, h, false);
    if (rc)
        return rc;
    if (n_frames == 0)
        return LUMACU_OK;
    if (n_frames > 65535u)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_decode_dev: at most 65535 frames per launch");
    const bool sub = (profile == 0 || profile == 2);
    const int bytes = profile > 1 ? 2 : 1;
    const uint32_t cw = sub ? (w + 1) >> 1 : w, chh = sub ? (h + 1) >> 1 : h;
    if (strides[0] < (int32_t)(w * bytes) || strides[1] < (int32_t)(cw * bytes) || strides[2] < (int32_t)(cw * bytes))
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_decode_dev: stride smaller than a row");
    if (n_frames > 1 && !plane_frame_stride)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_decode_dev: plane_frame_stride required for n_frames > 1");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;

    DecArgs a{};
    a.q = ctx->q;
    if (ctx->pq_off)
        a.q.pqd = nullptr, a.q.pqe = nullptr;
    a.q.tune_flags = ctx->dec_variant >= 2000 ? 1u : 0u;
    a.rgb = d_rgb;
    a.rgb_plane_stride = opt.rgb_plane_stride ? opt.rgb_plane_stride : (size_t)w * h;
    a.rgb_frame_stride = rgb_frame_stride ? rgb_frame_stride : (size_t)3 * w * h;
    a.w = w;
    a.h = h;
    a.sc = pre_scaling;
    a.prescale = (pre_scaling != 1.0f) ? 1 : 0;
    a.nz = make_float2(-0.0f, -0.0f);
    /* the vector kernels process whole 2-row x 4-column tiles: odd heights (legal for 4:4:4 decode) take the scalar
     * kernel, which masks the missing row */
    bool vec = (w % 4 == 0) && (h % 2 == 0) && aligned(d_rgb, 16) && (a.rgb_frame_stride % 4 == 0) && (a.rgb_plane_stride % 4 == 0);
    for (int p = 0; p < 3; p++) {
        a.plane[p] = d_planes[p];
        a.stride[p] = strides[p];
        a.plane_frame_stride[p] = plane_frame_stride ? plane_frame_stride[p] : (size_t)strides[p] * (p ? chh : h);
        const size_t al = (size_t)((p && sub) ? 2 : 4) * bytes;
        vec = vec && aligned(d_planes[p], al) && (strides[p] % al == 0) && (a.plane_frame_stride[p] % al == 0);
    }
    dec_fn fn = nullptr;
    size_t smem = ctx->smem_dec;
    const bool small32 = (uint64_t)w * h * 4 < (1ull << 32) && (uint64_t)strides[0] * h < (1ull << 32);
    a.passthrough = opt.passthrough ? 1 : 0;
    if (opt.display) {
        a.rgba = opt.rgba;
        a.rgba_pitch = opt.rgba_pitch;
        a.rgba_frame_stride = opt.rgba_frame_stride;
        a.disp_exposure = opt.display->exposure;
        a.disp_scaling = pre_scaling / opt.display->user_scaling; /* lumaplay.cpp:406 */
        a.disp_inv_gamma = 1.0f / opt.display->gamma;
        a.disp_tmo = opt.display->do_tmo ? 1 : 0;
        a.disp_ldr = opt.display->ldr_sim ? 1 : 0;
        a.prescale = 0; /* the player folds preScaling into `scaling` instead of dividing the frame */
        vec = (w % 4 == 0) && (h % 2 == 0);
        for (int p = 0; p < 3; p++) {
            const size_t al = (size_t)((p && sub) ? 2 : 4) * bytes;
            vec = vec && aligned(d_planes[p], al) && (strides[p] % al == 0) && (a.plane_frame_stride[p] % al == 0);
        }
        vec = vec && aligned(a.rgba, 16) && (a.rgba_frame_stride % 16 == 0);
    }
    if (opt.display && opt.display->filter == 1) { /* the player's own GL_LINEAR sampling: one thread per output pixel */
        const size_t npx = (size_t)w * h;
        const unsigned blocks = (unsigned)std::min<size_t>((npx + kThreads - 1) / kThreads, (size_t)ctx->sm_count * 8);
        launch_display_linear(blocks, n_frames, st, a, ctx->color_space, sub ? 1 : 0, bytes);
        CU_TRY(ctx, cudaGetLastError());
        ctx->last_fast = false;
        ctx->launches++;
        return LUMACU_OK;
    }
    if (vec && small32 && ctx->smem_dec_fast && !ctx->force_generic && !opt.passthrough && !opt.display) {
        fn = pick_dec_fast(ctx->color_space, sub, bytes, (ctx->dec_variant % 2000) ? ctx->dec_variant % 2000 : kDecDefaultVariant);
        if (!fn)
            fn = pick_dec_fast(ctx->color_space, sub, bytes, kDecVariantPlain);
        smem = ctx->smem_dec_fast;
    } else if (vec && small32 && ctx->dec_global_lut && bytes == 2 && !ctx->force_generic && !opt.passthrough && !opt.display) {
        fn = pick_dec_fast(ctx->color_space, sub, bytes, kDecVariantGlobalLut);
        smem = ctx->smem_dec_ctab;
    }
    ctx->last_fast = fn != nullptr;
    if (!fn)
        fn = pick_dec(ctx->color_space, sub, bytes, vec);
    const uint32_t ntiles = ((w + 3) / 4) * ((h + 1) / 2);
    uint32_t gx = 1;
    rc = grid_for(ctx, (const void *)fn, smem, ntiles, n_frames, kDecTilesPerThread, &gx);
    if (rc)
        return rc;
    fn<<<dim3(gx, n_frames), kThreads, smem, st>>>(a);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return LUMACU_OK;
}

extern "C" int lumacu_decode_dev(lumacu_ctx *ctx, const uint8_t *const d_planes[3], const int32_t strides[3], uint32_t w,
                                 uint32_t h, int profile, float pre_scaling, float *d_rgb, uint32_t n_frames,
                                 size_t rgb_frame_stride, const size_t plane_frame_stride[3], void *stream)
try {
    return decode_launch(ctx, d_planes, strides, w, h, profile, pre_scaling, d_rgb, n_frames, rgb_frame_stride, plane_frame_stride,
                         stream, LaunchOpts());
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

/* ---- display decode ------------------------------------------------------------------- */
extern "C" int lumacu_display_dev(lumacu_ctx *ctx, const uint8_t *const d_planes[3], const int32_t strides[3], uint32_t w,
                                  uint32_t h, int profile, float pre_scaling, const lumacu_display_params *params,
                                  uint8_t *d_rgba, int32_t rgba_pitch, uint32_t n_frames, const size_t plane_frame_stride[3],
                                  size_t rgba_frame_stride, void *stream)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!params || !d_rgba)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_display_dev: NULL pointer argument");
    if (rgba_pitch < (int32_t)(w * 4) || (rgba_pitch & 3))
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_display_dev: rgba_pitch must be a multiple of 4 and >= 4*w");
    if (!(params->gamma > 0.0f) || !(params->user_scaling > 0.0f))
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_display_dev: gamma and user_scaling must be positive");
    if (params->filter != 0 && params->filter != 1)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_display_dev: filter %d not in {0,1}", params->filter);
    if (!aligned(d_rgba, 4))
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_display_dev: rgba must be 4-byte aligned");
    LaunchOpts opt;
    opt.display = params;
    opt.rgba = d_rgba;
    opt.rgba_pitch = rgba_pitch;
    opt.rgba_frame_stride = rgba_frame_stride ? rgba_frame_stride : (size_t)rgba_pitch * h;
    return decode_launch(ctx, d_planes, strides, w, h, profile, pre_scaling, nullptr, n_frames, 0, plane_frame_stride, stream, opt);
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

/* ---- frame sources on the device ------------------------------------------------------------------------ */
extern "C" int lumacu_test_frame_dev(lumacu_ctx *ctx, float *d_rgb, uint32_t w, uint32_t h, void *stream)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!d_rgb || !w || !h)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_test_frame_dev: NULL frame or empty size");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const size_t n = (size_t)w * h;
    const unsigned blocks = (unsigned)std::min<size_t>((n + kThreads - 1) / kThreads, (size_t)ctx->sm_count * 8);
    launch_test_frame(blocks, st, d_rgb, w, h);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return LUMACU_OK;
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

extern "C" int lumacu_half_rgba_to_frame_dev(lumacu_ctx *ctx, const void *d_rgba_half, uint32_t w, uint32_t h, int channels,
                                             float *d_rgb, void *stream)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!d_rgba_half || !d_rgb || !w || !h)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_half_rgba_to_frame_dev: NULL pointer or empty size");
    if (channels != 1 && channels != 2 && channels != 4 && channels != 7 && channels != 15)
        /* src/exr_interface.cpp:138-140 */
        return fail(ctx, LUMACU_ERR_UNSUPPORTED, "Reading of luminance only frames not yet supported");
    if (!aligned(d_rgba_half, 8))
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_half_rgba_to_frame_dev: pixels must be 8-byte aligned");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const size_t n = (size_t)w * h;
    const unsigned blocks = (unsigned)std::min<size_t>((n + kThreads - 1) / kThreads, (size_t)ctx->sm_count * 8);
    launch_half_rgba_to_frame(blocks, st, d_rgba_half, d_rgb, n, n, channels);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return LUMACU_OK;
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

extern "C" int lumacu_frame_to_half_rgba_dev(lumacu_ctx *ctx, const float *d_rgb, uint32_t w, uint32_t h, void *d_rgba_half,
                                             void *stream)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!d_rgba_half || !d_rgb || !w || !h)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_frame_to_half_rgba_dev: NULL pointer or empty size");
    if (!aligned(d_rgba_half, 8))
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_frame_to_half_rgba_dev: pixels must be 8-byte aligned");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const size_t n = (size_t)w * h;
    const unsigned blocks = (unsigned)std::min<size_t>((n + kThreads - 1) / kThreads, (size_t)ctx->sm_count * 8);
    launch_frame_to_half_rgba(blocks, st, d_rgb, d_rgba_half, n, n);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return LUMACU_OK;
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

/* PFS frame source / sink (src/pfs_interface.cpp:57-113, :115-152): three separate channel arrays <-> planar frame */
static int pfs_channels(lumacu_ctx *ctx, bool to_rgb, const float *a0, const float *a1, const float *a2, float *o0, float *o1,
                        float *o2, uint32_t w, uint32_t h, void *stream, const char *who)
{
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!a0 || !a1 || !a2 || !o0 || !w || !h)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "%s: NULL pointer or empty size", who);
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const size_t n = (size_t)w * h;
    const unsigned blocks = (unsigned)std::min<size_t>((n + kThreads - 1) / kThreads, (size_t)ctx->sm_count * 8);
    launch_pfs_channels(to_rgb, blocks, st, a0, a1, a2, o0, o1, o2, n);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return LUMACU_OK;
}

extern "C" int lumacu_pfs_xyz_to_frame_dev(lumacu_ctx *ctx, const float *d_x, const float *d_y, const float *d_z, uint32_t w,
                                           uint32_t h, float *d_rgb, void *stream)
try {
    const size_t n = (size_t)w * h;
    return pfs_channels(ctx, true, d_x, d_y, d_z, d_rgb, d_rgb ? d_rgb + n : nullptr, d_rgb ? d_rgb + 2 * n : nullptr, w, h, stream,
                        "lumacu_pfs_xyz_to_frame_dev");
}
LUMACU_CATCH(ctx)

extern "C" int lumacu_frame_to_pfs_xyz_dev(lumacu_ctx *ctx, const float *d_rgb, uint32_t w, uint32_t h, float *d_x, float *d_y,
                                           float *d_z, void *stream)
try {
    const size_t n = (size_t)w * h;
    if (!d_y || !d_z)
        return ctx ? fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_frame_to_pfs_xyz_dev: NULL pointer") : LUMACU_ERR_INVALID_ARGUMENT;
    return pfs_channels(ctx, false, d_rgb, d_rgb ? d_rgb + n : nullptr, d_rgb ? d_rgb + 2 * n : nullptr, d_x, d_y, d_z, w, h, stream,
                        "lumacu_frame_to_pfs_xyz_dev");
}
LUMACU_CATCH(ctx)

extern "C" int lumacu_transform_color_space_dev(lumacu_ctx *ctx, float *d_frame, uint32_t w, uint32_t h, int to_cs, float sc,
                                                void *stream)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!ctx->configured)
        return fail(ctx, LUMACU_ERR_NOT_CONFIGURED, "lumacu_transform_color_space_dev: quantizer not set");
    if (!d_frame)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_transform_color_space_dev: NULL frame");
    const size_t n = (size_t)w * h;
    if (!n)
        return LUMACU_OK;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const uint32_t blocks = (uint32_t)std::min<size_t>((n + kThreads - 1) / kThreads, (size_t)ctx->sm_count * 16);
    float *c0 = d_frame, *c1 = d_frame + n, *c2 = d_frame + 2 * n;
    const float L = ctx->q.l_max;
    launch_transform(ctx->color_space, to_cs != 0, blocks, st, c0, c1, c2, n, sc, L);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return LUMACU_OK;
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

static bool channel_uses_lut(const lumacu_ctx *ctx, unsigned ch)
{
    return ch == 0 || ctx->color_space == CS_RGB || ctx->color_space == CS_XYZ;
}

extern "C" int lumacu_quantize_dev(lumacu_ctx *ctx, const float *d_in, float *d_out, size_t n, unsigned ch, void *stream)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!ctx->configured)
        return fail(ctx, LUMACU_ERR_NOT_CONFIGURED, "lumacu_quantize_dev: quantizer not set");
    if (!n)
        return LUMACU_OK;
    if (!d_in || !d_out)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_quantize_dev: NULL pointer");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const void *fn = quantize_kernel_ptr();
    if (ctx->smem_enc > 32 * 1024) /* the 48 KB default also has to hold the kernel's static shared memory */
        CU_TRY(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_enc));
    const uint32_t blocks = (uint32_t)std::min<size_t>((n + kThreads - 1) / kThreads, (size_t)ctx->sm_count * 4);
    launch_quantize(blocks, ctx->smem_enc, st, ctx->q, d_in, d_out, n, channel_uses_lut(ctx, ch) ? 1 : 0);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return LUMACU_OK;
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

extern "C" int lumacu_dequantize_dev(lumacu_ctx *ctx, const float *d_in, float *d_out, size_t n, unsigned ch, void *stream)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!ctx->configured)
        return fail(ctx, LUMACU_ERR_NOT_CONFIGURED, "lumacu_dequantize_dev: quantizer not set");
    if (!n)
        return LUMACU_OK;
    if (!d_in || !d_out)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_dequantize_dev: NULL pointer");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const uint32_t blocks = (uint32_t)std::min<size_t>((n + kThreads - 1) / kThreads, (size_t)ctx->sm_count * 8);
    launch_dequantize(blocks, st, ctx->q, d_in, d_out, n, channel_uses_lut(ctx, ch) ? 1 : 0);
    CU_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return LUMACU_OK;
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

/* ================================ host-pointer entry points ========================== */
namespace {

void plane_geometry(uint32_t w, uint32_t h, int profile, uint32_t pw[3], uint32_t ph[3])
{
    const bool sub = (profile == 0 || profile == 2);
    pw[0] = w;
    ph[0] = h;
    pw[1] = pw[2] = sub ? (w + 1) >> 1 : w;
    ph[1] = ph[2] = sub ? (h + 1) >> 1 : h;
}

/* device-side plane layout used by the host entry points: 256-byte aligned pitches */
void device_plane_layout(uint32_t w, uint32_t h, int profile, int32_t dstride[3], size_t off[3], size_t *total)
{
    uint32_t pw[3], ph[3];
    plane_geometry(w, h, profile, pw, ph);
    const int bytes = profile > 1 ? 2 : 1;
    size_t o = 0;
    for (int p = 0; p < 3; p++) {
        dstride[p] = (int32_t)(((size_t)pw[p] * bytes + 255) & ~(size_t)255);
        off[p] = o;
        o += (size_t)dstride[p] * ph[p];
    }
    *total = o;
}

} // namespace

namespace {

/* Row bands for the host-pointer entry points.  A frame is cut into nb horizontal bands on even-row
 * boundaries (2x2 chroma blocks stay intact); band b's H2D copy, kernel and D2H copy run on three streams
 * chained by events, so the copy of band b+1 overlaps the kernel of band b and the read-back of band b-1:
 * the call costs about max(H2D, D2H) instead of H2D + kernel + D2H.  Small frames use one band. */
int band_count(const lumacu_ctx *ctx, uint32_t w, uint32_t h)
{
    if (ctx->host_bands > 0)
        return std::max(1, std::min<int>(ctx->host_bands, (int)(h / 2)));
    /* about one band per 8 MiB of float frame, at most 8 (measured on B200 + PCIe Gen5, scripts/e2e_probe.py: best
     * counts 2-3 at 720p, 3-4 at 1080p, 4 at 1440p, 8 at 4K and 8K; each band costs ~10 API calls) */
    const size_t bytes = (size_t)w * h * 12;
    const size_t nb = (bytes + ((size_t)8 << 20) - 1) / ((size_t)8 << 20);
    return (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(8, nb), h / 64));
}

int ensure_events(lumacu_ctx *ctx, int nb)
{
    while ((int)ctx->ev_in.size() < nb) {
        cudaEvent_t a = nullptr, b = nullptr;
        CU_TRY(ctx, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        ctx->ev_in.push_back(a);
        CU_TRY(ctx, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        ctx->ev_k.push_back(b);
    }
    return LUMACU_OK;
}

inline uint32_t band_row(uint32_t h, int nb, int b) /* first row of band b; even; band nb starts at h */
{
    if (b >= nb)
        return h;
    return (uint32_t)(((uint64_t)(h / 2) * b / nb) * 2);
}

} // namespace

/* Completes the asynchronous host-pointer call in flight on this context, if any: waits for its last D2H copy and
 * folds the per-band statistics into the caller's lumacu_frame_stats.  Every host-pointer entry point starts with it
 * (the staging buffers and streams are per context: ONE call in flight). */
static int finish_pending(lumacu_ctx *ctx)
{
    if (!ctx->pending)
        return LUMACU_OK;
    ctx->pending = false;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->s_out));
    if (ctx->pending_stats) {
        const lumacu_frame_stats *hs = (const lumacu_frame_stats *)ctx->h_pin;
        lumacu_frame_stats *stats = ctx->pending_stats;
        *stats = hs[0];
        for (int b = 1; b < ctx->pending_bands; b++) {
            stats->sum += hs[b].sum;
            stats->max = fmaxf(stats->max, hs[b].max);
            stats->min = fminf(stats->min, hs[b].min);
        }
        ctx->pending_stats = nullptr;
    }
    return LUMACU_OK;
}

/* The host-pointer entry points run at PCIe rate only from page-locked memory; with pageable memory the driver stages
 * every copy through its own bounce buffer (several times slower, and the "asynchronous" copies block).  That is legal,
 * so the call proceeds -- but not silently: once per context a note goes to stderr (the reference reports its own
 * warnings there, src/luma_encoder.cpp:314-316), unless LUMACU_QUIET is set. */
static void note_if_pageable(lumacu_ctx *ctx, const void *p, const char *what)
{
    if (ctx->pageable_warned)
        return;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    if (at.type != cudaMemoryTypeUnregistered)
        return;
    ctx->pageable_warned = true;
    const char *q = getenv("LUMACU_QUIET");
    if (q && q[0] && q[0] != '0')
        return;
    fprintf(stderr, "lumacu: %s is pageable host memory: copies are staged by the driver (slow). Allocate it with "
                    "lumacu_host_alloc or page-lock it with lumacu_host_register. (LUMACU_QUIET=1 silences this note.)\n", what);
}

/* A host-pointer call that fails half way (some bands already queued) must not return while copies into or out of the
 * caller's buffers are still in flight: drain the three streams unless the call got as far as handing over to
 * finish_pending. */
struct DrainUnlessQueued {
    lumacu_ctx *ctx;
    bool queued = false;
    explicit DrainUnlessQueued(lumacu_ctx *c) : ctx(c) {}
    ~DrainUnlessQueued()
    {
        if (queued)
            return;
        cudaStreamSynchronize(ctx->s_in);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->s_out);
        cudaGetLastError();
    }
};

static int ensure_async_state(lumacu_ctx *ctx)
{
    if (!ctx->h_pin) {
        CU_TRY(ctx, cudaHostAlloc(&ctx->h_pin, 4096, cudaHostAllocPortable));
        ctx->h_pin_cap = 4096;
    }
    if (!ctx->ev_input)
        CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_input, cudaEventDisableTiming));
    return LUMACU_OK;
}

/* half_src != NULL: the frame arrives as interleaved half-float RGBA pixels (Imf::Rgba, 8 B/px over PCIe instead of 12) and
 * is expanded on the device by ExrInterface::readFrame's pixel loop (half_channels = Imf::RgbaChannels) band by band;
 * rgb is unused then. */
static int host_encode(lumacu_ctx *ctx, float *rgb, uint32_t w, uint32_t h, int profile, float pre_scaling,
                       uint8_t *const planes[3], const int32_t strides[3], int write_back, lumacu_frame_stats *stats,
                       bool passthrough, bool async = false, const void *half_src = nullptr, int half_channels = 7)
{
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!ctx->configured)
        return fail(ctx, LUMACU_ERR_NOT_CONFIGURED, "lumacu_encode: lumacu_set_quantizer has not been called");
    if ((!rgb && !half_src) || !planes || !strides || !planes[0] || !planes[1] || !planes[2])
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_encode: NULL pointer argument");
    if (half_src && half_channels != 1 && half_channels != 2 && half_channels != 4 && half_channels != 7 && half_channels != 15)
        return fail(ctx, LUMACU_ERR_UNSUPPORTED, "Reading of luminance only frames not yet supported"); /* src/exr_interface.cpp:138-140 */
    int rc = check_profile(ctx, profile, w, h, true);
    if (rc)
        return rc;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if ((rc = finish_pending(ctx)) || (rc = ensure_async_state(ctx)))
        return rc;
    note_if_pageable(ctx, half_src ? half_src : (const void *)rgb, "the frame passed to lumacu_encode");
    note_if_pageable(ctx, planes[0], "the plane buffer passed to lumacu_encode");
    const size_t npx = (size_t)w * h;
    uint32_t pw[3], ph[3];
    plane_geometry(w, h, profile, pw, ph);
    const int bytes = profile > 1 ? 2 : 1;
    const bool sub = (profile == 0 || profile == 2);
    for (int p = 0; p < 3; p++)
        if (strides[p] < (int32_t)(pw[p] * bytes))
            return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_encode: stride[%d] smaller than a row", p);
    int32_t dstride[3];
    size_t off[3], ptotal;
    device_plane_layout(w, h, profile, dstride, off, &ptotal);
    const int nb = band_count(ctx, w, h);
    if ((rc = reserve(ctx, ctx->d_rgb, npx * 12)) || (rc = reserve(ctx, ctx->d_planes, ptotal)) ||
        (rc = reserve(ctx, ctx->d_stats, sizeof(lumacu_frame_stats) * 64)) || (rc = ensure_events(ctx, nb)) ||
        (half_src && (rc = reserve(ctx, ctx->d_half, npx * 8))))
        return rc;
    float *d_rgb = (float *)ctx->d_rgb.p;
    uint8_t *dp[3] = {(uint8_t *)ctx->d_planes.p + off[0], (uint8_t *)ctx->d_planes.p + off[1],
                      (uint8_t *)ctx->d_planes.p + off[2]};
    lumacu_frame_stats *d_stats = (lumacu_frame_stats *)ctx->d_stats.p;
    LaunchOpts opt;
    opt.passthrough = passthrough;
    opt.rgb_plane_stride = npx; /* a band's planes are still a whole frame apart */
    DrainUnlessQueued drain(ctx);
    for (int b = 0; b < nb; b++) {
        const uint32_t y0 = band_row(h, nb, b), y1 = band_row(h, nb, b + 1), rows = y1 - y0;
        if (half_src) /* the band's pixels are contiguous */
            CU_TRY(ctx, cudaMemcpyAsync((uint8_t *)ctx->d_half.p + (size_t)y0 * w * 8, (const uint8_t *)half_src + (size_t)y0 * w * 8,
                                        (size_t)rows * w * 8, cudaMemcpyHostToDevice, ctx->s_in));
        else /* the band's rows of the three planes in one strided copy (pitch = one plane) */
            CU_TRY(ctx, cudaMemcpy2DAsync(d_rgb + (size_t)y0 * w, npx * 4, rgb + (size_t)y0 * w, npx * 4, (size_t)rows * w * 4, 3,
                                          cudaMemcpyHostToDevice, ctx->s_in));
        CU_TRY(ctx, cudaEventRecord(ctx->ev_in[b], ctx->s_in));
        if (b == nb - 1 && !write_back)
            CU_TRY(ctx, cudaEventRecord(ctx->ev_input, ctx->s_in)); /* the caller's frame has been read completely */
        CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[b], 0));
        if (half_src) {
            const size_t n = (size_t)rows * w;
            const unsigned blocks = (unsigned)std::min<size_t>((n + kThreads - 1) / kThreads, (size_t)ctx->sm_count * 8);
            launch_half_rgba_to_frame(blocks, ctx->stream, (const uint8_t *)ctx->d_half.p + (size_t)y0 * w * 8, d_rgb + (size_t)y0 * w, n,
                                      npx, half_channels);
            CU_TRY(ctx, cudaGetLastError());
            ctx->launches++;
        }
        const uint32_t cy0 = sub ? y0 >> 1 : y0;
        uint8_t *bp[3] = {dp[0] + (size_t)y0 * dstride[0], dp[1] + (size_t)cy0 * dstride[1], dp[2] + (size_t)cy0 * dstride[2]};
        float *band = d_rgb + (size_t)y0 * w;
        rc = encode_launch(ctx, band, write_back ? band : nullptr, w, rows, profile, pre_scaling, bp, dstride, 1, 0, nullptr,
                           stats ? d_stats + b : nullptr, ctx->stream, opt);
        if (rc)
            return rc;
        CU_TRY(ctx, cudaEventRecord(ctx->ev_k[b], ctx->stream));
        CU_TRY(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_k[b], 0));
        for (int p = 0; p < 3; p++) {
            const uint32_t py0 = (p && sub) ? y0 >> 1 : y0, prows = (p && sub) ? rows >> 1 : rows;
            CU_TRY(ctx, cudaMemcpy2DAsync(planes[p] + (size_t)py0 * strides[p], (size_t)strides[p], dp[p] + (size_t)py0 * dstride[p],
                                          (size_t)dstride[p], (size_t)pw[p] * bytes, prows, cudaMemcpyDeviceToHost, ctx->s_out));
        }
        if (write_back)
            CU_TRY(ctx, cudaMemcpy2DAsync(rgb + (size_t)y0 * w, npx * 4, band, npx * 4, (size_t)rows * w * 4, 3,
                                          cudaMemcpyDeviceToHost, ctx->s_out));
    }
    if (stats)
        CU_TRY(ctx, cudaMemcpyAsync(ctx->h_pin, d_stats, sizeof(lumacu_frame_stats) * nb, cudaMemcpyDeviceToHost, ctx->s_out));
    if (write_back) /* with the in-place side effect the caller's frame is busy until the last D2H copy */
        CU_TRY(ctx, cudaEventRecord(ctx->ev_input, ctx->s_out));
    drain.queued = true;
    ctx->pending = true;
    ctx->pending_stats = stats;
    ctx->pending_bands = nb;
    return async ? LUMACU_OK : finish_pending(ctx);
}

extern "C" int lumacu_encode(lumacu_ctx *ctx, float *rgb, uint32_t w, uint32_t h, int profile, float pre_scaling,
                             uint8_t *const planes[3], const int32_t strides[3], int write_back,
                             lumacu_frame_stats *stats)
try {
    return host_encode(ctx, rgb, w, h, profile, pre_scaling, planes, strides, write_back, stats, false);
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

/* half_dst != NULL: the decoded frame leaves as interleaved half-float RGBA pixels (ExrInterface::writeFrame's pixel loop
 * run on the device band by band: 8 B/px over PCIe instead of 12); rgb is unused then. */
static int host_decode(lumacu_ctx *ctx, const uint8_t *const planes[3], const int32_t strides[3], uint32_t w, uint32_t h,
                       int profile, float pre_scaling, float *rgb, bool passthrough, bool async = false, void *half_dst = nullptr)
{
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!ctx->configured)
        return fail(ctx, LUMACU_ERR_NOT_CONFIGURED, "lumacu_decode: lumacu_set_quantizer has not been called");
    if ((!rgb && !half_dst) || !planes || !strides || !planes[0] || !planes[1] || !planes[2])
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_decode: NULL pointer argument");
    int rc = check_profile(ctx, profile, w, h, false);
    if (rc)
        return rc;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if ((rc = finish_pending(ctx)) || (rc = ensure_async_state(ctx)))
        return rc;
    note_if_pageable(ctx, half_dst ? (const void *)half_dst : (const void *)rgb, "the frame passed to lumacu_decode");
    note_if_pageable(ctx, planes[0], "the plane buffer passed to lumacu_decode");
    const size_t npx = (size_t)w * h;
    uint32_t pw[3], ph[3];
    plane_geometry(w, h, profile, pw, ph);
    const int bytes = profile > 1 ? 2 : 1;
    const bool sub = (profile == 0 || profile == 2);
    for (int p = 0; p < 3; p++)
        if (strides[p] < (int32_t)(pw[p] * bytes))
            return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_decode: stride[%d] smaller than a row", p);
    int32_t dstride[3];
    size_t off[3], ptotal;
    device_plane_layout(w, h, profile, dstride, off, &ptotal);
    /* the reference accepts odd decoded sizes in principle (plane dims are rounded up); bands need even rows */
    const int nb = (h % 2 == 0) ? band_count(ctx, w, h) : 1;
    if ((rc = reserve(ctx, ctx->d_rgb, npx * 12)) || (rc = reserve(ctx, ctx->d_planes, ptotal)) || (rc = ensure_events(ctx, nb)) ||
        (half_dst && (rc = reserve(ctx, ctx->d_half, npx * 8))))
        return rc;
    float *d_rgb = (float *)ctx->d_rgb.p;
    uint8_t *dp[3] = {(uint8_t *)ctx->d_planes.p + off[0], (uint8_t *)ctx->d_planes.p + off[1],
                      (uint8_t *)ctx->d_planes.p + off[2]};
    LaunchOpts opt;
    opt.passthrough = passthrough;
    opt.rgb_plane_stride = npx;
    DrainUnlessQueued drain(ctx);
    for (int b = 0; b < nb; b++) {
        const uint32_t y0 = nb == 1 ? 0 : band_row(h, nb, b), y1 = nb == 1 ? h : band_row(h, nb, b + 1), rows = y1 - y0;
        const uint8_t *bp[3];
        for (int p = 0; p < 3; p++) {
            const uint32_t py0 = (p && sub) ? y0 >> 1 : y0, prows = (p && sub) ? (rows + 1) >> 1 : rows;
            CU_TRY(ctx, cudaMemcpy2DAsync(dp[p] + (size_t)py0 * dstride[p], (size_t)dstride[p], planes[p] + (size_t)py0 * strides[p],
                                          (size_t)strides[p], (size_t)pw[p] * bytes, prows, cudaMemcpyHostToDevice, ctx->s_in));
            bp[p] = dp[p] + (size_t)py0 * dstride[p];
        }
        CU_TRY(ctx, cudaEventRecord(ctx->ev_in[b], ctx->s_in));
        if (b == nb - 1)
            CU_TRY(ctx, cudaEventRecord(ctx->ev_input, ctx->s_in)); /* the caller's planes have been read completely */
        CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[b], 0));
        float *band = d_rgb + (size_t)y0 * w;
        rc = decode_launch(ctx, bp, dstride, w, rows, profile, pre_scaling, band, 1, 0, nullptr, ctx->stream, opt);
        if (rc)
            return rc;
        if (half_dst) {
            const size_t n = (size_t)rows * w;
            const unsigned blocks = (unsigned)std::min<size_t>((n + kThreads - 1) / kThreads, (size_t)ctx->sm_count * 8);
            launch_frame_to_half_rgba(blocks, ctx->stream, band, (uint8_t *)ctx->d_half.p + (size_t)y0 * w * 8, n, npx);
            CU_TRY(ctx, cudaGetLastError());
            ctx->launches++;
        }
        CU_TRY(ctx, cudaEventRecord(ctx->ev_k[b], ctx->stream));
        CU_TRY(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_k[b], 0));
        if (half_dst)
            CU_TRY(ctx, cudaMemcpyAsync((uint8_t *)half_dst + (size_t)y0 * w * 8, (const uint8_t *)ctx->d_half.p + (size_t)y0 * w * 8,
                                        (size_t)rows * w * 8, cudaMemcpyDeviceToHost, ctx->s_out));
        else
            CU_TRY(ctx, cudaMemcpy2DAsync(rgb + (size_t)y0 * w, npx * 4, band, npx * 4, (size_t)rows * w * 4, 3, cudaMemcpyDeviceToHost,
                                          ctx->s_out));
    }
    drain.queued = true;
    ctx->pending = true;
    ctx->pending_stats = nullptr;
    ctx->pending_bands = nb;
    return async ? LUMACU_OK : finish_pending(ctx);
}

extern "C" int lumacu_decode(lumacu_ctx *ctx, const uint8_t *const planes[3], const int32_t strides[3], uint32_t w,
                             uint32_t h, int profile, float pre_scaling, float *rgb)
try {
    return host_decode(ctx, planes, strides, w, h, profile, pre_scaling, rgb, false);
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

/* ---- EXR-sourced / EXR-bound frames without the f32 detour over PCIe ---------------------------------------------- */
extern "C" int lumacu_encode_half_rgba(lumacu_ctx *ctx, const void *rgba_half, uint32_t w, uint32_t h, int channels, int profile,
                                       float pre_scaling, uint8_t *const planes[3], const int32_t strides[3],
                                       lumacu_frame_stats *stats)
try {
    if (ctx && !rgba_half)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_encode_half_rgba: NULL pointer argument");
    return host_encode(ctx, nullptr, w, h, profile, pre_scaling, planes, strides, 0, stats, false, false, rgba_half, channels);
}
LUMACU_CATCH(ctx)

extern "C" int lumacu_decode_half_rgba(lumacu_ctx *ctx, const uint8_t *const planes[3], const int32_t strides[3], uint32_t w,
                                       uint32_t h, int profile, float pre_scaling, void *rgba_half)
try {
    if (ctx && !rgba_half)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_decode_half_rgba: NULL pointer argument");
    return host_decode(ctx, planes, strides, w, h, profile, pre_scaling, nullptr, false, false, rgba_half);
}
LUMACU_CATCH(ctx)

/* ---- asynchronous pair: queue the call, return; lumacu_wait_input / lumacu_wait complete it ------------------ */
extern "C" int lumacu_encode_async(lumacu_ctx *ctx, float *rgb, uint32_t w, uint32_t h, int profile, float pre_scaling,
                                   uint8_t *const planes[3], const int32_t strides[3], int write_back,
                                   lumacu_frame_stats *stats)
try {
    return host_encode(ctx, rgb, w, h, profile, pre_scaling, planes, strides, write_back, stats, false, true);
}
LUMACU_CATCH(ctx)

extern "C" int lumacu_decode_async(lumacu_ctx *ctx, const uint8_t *const planes[3], const int32_t strides[3], uint32_t w,
                                   uint32_t h, int profile, float pre_scaling, float *rgb)
try {
    return host_decode(ctx, planes, strides, w, h, profile, pre_scaling, rgb, false, true);
}
LUMACU_CATCH(ctx)

extern "C" int lumacu_wait_input(lumacu_ctx *ctx)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!ctx->pending)
        return LUMACU_OK;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaEventSynchronize(ctx->ev_input));
    return LUMACU_OK;
}
LUMACU_CATCH(ctx)

extern "C" int lumacu_wait(lumacu_ctx *ctx)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    return finish_pending(ctx);
}
LUMACU_CATCH(ctx)

extern "C" int lumacu_pending(const lumacu_ctx *ctx) { return (ctx && ctx->pending) ? 1 : 0; }

extern "C" int lumacu_display(lumacu_ctx *ctx, const uint8_t *const planes[3], const int32_t strides[3], uint32_t w, uint32_t h,
                              int profile, float pre_scaling, const lumacu_display_params *params, uint8_t *rgba,
                              int32_t rgba_pitch)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!ctx->configured)
        return fail(ctx, LUMACU_ERR_NOT_CONFIGURED, "lumacu_display: lumacu_set_quantizer has not been called");
    if (!rgba || !planes || !strides || !planes[0] || !planes[1] || !planes[2])
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_display: NULL pointer argument");
    int rc = check_profile(ctx, profile, w, h, false);
    if (rc)
        return rc;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if ((rc = finish_pending(ctx)))
        return rc;
    uint32_t pw[3], ph[3];
    plane_geometry(w, h, profile, pw, ph);
    const int bytes = profile > 1 ? 2 : 1;
    for (int p = 0; p < 3; p++)
        if (strides[p] < (int32_t)(pw[p] * bytes))
            return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_display: stride[%d] smaller than a row", p);
    int32_t dstride[3];
    size_t off[3], ptotal;
    device_plane_layout(w, h, profile, dstride, off, &ptotal);
    const size_t dpitch = ((size_t)w * 4 + 255) & ~(size_t)255;
    if ((rc = reserve(ctx, ctx->d_planes, ptotal)) || (rc = reserve(ctx, ctx->d_aux, dpitch * h)))
        return rc;
    uint8_t *dp[3] = {(uint8_t *)ctx->d_planes.p + off[0], (uint8_t *)ctx->d_planes.p + off[1],
                      (uint8_t *)ctx->d_planes.p + off[2]};
    for (int p = 0; p < 3; p++)
        CU_TRY(ctx, cudaMemcpy2DAsync(dp[p], (size_t)dstride[p], planes[p], (size_t)strides[p], (size_t)pw[p] * bytes, ph[p],
                                      cudaMemcpyHostToDevice, ctx->stream));
    rc = lumacu_display_dev(ctx, dp, dstride, w, h, profile, pre_scaling, params, (uint8_t *)ctx->d_aux.p, (int32_t)dpitch, 1,
                            nullptr, 0, ctx->stream);
    if (rc)
        return rc;
    CU_TRY(ctx, cudaMemcpy2DAsync(rgba, (size_t)rgba_pitch, ctx->d_aux.p, dpitch, (size_t)w * 4, h, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LUMACU_OK;
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

extern "C" int lumacu_quantize_planes(lumacu_ctx *ctx, const float *frame, uint32_t w, uint32_t h, int profile,
                                      uint8_t *const planes[3], const int32_t strides[3], lumacu_frame_stats *stats)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    return host_encode(ctx, const_cast<float *>(frame), w, h, profile, 1.0f, planes, strides, 0, stats, true);
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

extern "C" int lumacu_dequantize_planes(lumacu_ctx *ctx, const uint8_t *const planes[3], const int32_t strides[3],
                                        uint32_t w, uint32_t h, int profile, float *frame)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    return host_decode(ctx, planes, strides, w, h, profile, 1.0f, frame, true);
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

extern "C" int lumacu_transform_color_space(lumacu_ctx *ctx, float *frame, uint32_t w, uint32_t h, int to_cs, float sc)
try {
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!ctx->configured)
        return fail(ctx, LUMACU_ERR_NOT_CONFIGURED, "lumacu_transform_color_space: quantizer not set");
    if (!frame)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "lumacu_transform_color_space: NULL frame");
    const size_t bytes = (size_t)w * h * 12;
    if (!bytes)
        return LUMACU_OK;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = finish_pending(ctx);
    if (rc || (rc = reserve(ctx, ctx->d_rgb, bytes)))
        return rc;
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_rgb.p, frame, bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = lumacu_transform_color_space_dev(ctx, (float *)ctx->d_rgb.p, w, h, to_cs, sc, ctx->stream);
    if (rc)
        return rc;
    CU_TRY(ctx, cudaMemcpyAsync(frame, ctx->d_rgb.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LUMACU_OK;
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

static int elementwise_host(lumacu_ctx *ctx, const float *in, float *out, size_t n, unsigned ch, bool quant)
{
    if (!ctx)
        return LUMACU_ERR_INVALID_ARGUMENT;
    if (!ctx->configured)
        return fail(ctx, LUMACU_ERR_NOT_CONFIGURED, "quantizer not set");
    if (!n)
        return LUMACU_OK;
    if (!in || !out)
        return fail(ctx, LUMACU_ERR_INVALID_ARGUMENT, "NULL pointer");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = finish_pending(ctx);
    if (rc || (rc = reserve(ctx, ctx->d_aux, n * 8)))
        return rc;
    float *d_in = (float *)ctx->d_aux.p, *d_out = d_in + n;
    CU_TRY(ctx, cudaMemcpyAsync(d_in, in, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    rc = quant ? lumacu_quantize_dev(ctx, d_in, d_out, n, ch, ctx->stream)
               : lumacu_dequantize_dev(ctx, d_in, d_out, n, ch, ctx->stream);
    if (rc)
        return rc;
    CU_TRY(ctx, cudaMemcpyAsync(out, d_out, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LUMACU_OK;
}

extern "C" int lumacu_quantize(lumacu_ctx *ctx, const float *in, float *out, size_t n, unsigned ch)
try {
    return elementwise_host(ctx, in, out, n, ch, true);
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

extern "C" int lumacu_dequantize(lumacu_ctx *ctx, const float *in, float *out, size_t n, unsigned ch)
try {
    return elementwise_host(ctx, in, out, n, ch, false);
}
LUMACU_CATCH(const_cast<lumacu_ctx *>(ctx))

/*
 * luma_kernels.cuh -- the fused sm_100a kernels of the HDR<->integer transform.
 *
 *   encode_kernel : planar f32 RGB -> {colour transform -> luma LUT search |
 *                   chroma 2x2 mean -> round/clamp} -> pitched u8 / LE-u16 planes
 *                   (+ per-frame sum/max/min of plane 0), one pass over HBM.
 *                   Replaces LumaQuantizer::transformColorSpace(frame,true,sc)
 *                   + LumaEncoder::setVpxChannel x3 (reference
 *                   src/luma_quantizer.cpp:269-373, src/luma_encoder.cpp:260-317).
 *   decode_kernel : pitched planes -> LUT gather / chroma scale -> 2x2 replicate
 *                   -> inverse colour -> planar f32 RGB.  Replaces
 *                   LumaDecoder::getVpxChannels + transformColorSpace(frame,false,sc)
 *                   (src/luma_decoder.cpp:205-240, src/luma_quantizer.cpp:374-479).
 *
 * Work decomposition: one thread owns a 2-row x 4-column pixel tile, so a 4:2:0
 * chroma block (2x2) is thread-local, each row-plane access of a warp is one
 * contiguous 512 B (f32) / 256 B (u16) segment, and all global accesses are
 * 128/64/32-bit vectors.  Blocks are persistent (grid-stride over tiles) so the
 * search tables are staged into shared memory once per block.
 */
#pragma once

#include <cuda_fp16.h>

#include "luma_device.cuh"
#include "luma_kernels_decl.cuh"

namespace lumacu {

/* ---- streaming global accesses (each byte is touched once) --------------------- */
__device__ __forceinline__ float4 ld_stream4(const float *p) { return __ldcs(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void st_stream4(float *p, float4 v) { __stcs(reinterpret_cast<float4 *>(p), v); }

/* Stage the search tables (thresholds + bucket heads) into shared memory. */
__device__ __forceinline__ SearchCtx make_search_ctx(const QuantDev &q, unsigned char *smem)
{
    SearchCtx s;
    s.lut = q.lut;
    s.max_val = q.max_val;
    s.shift = q.shift;
    s.base = q.base;
    s.nbm1 = q.nbm1;
    s.walk = q.walk;
    s.mode = q.search_mode;
    s.thr = q.thr;
    s.bucket = q.bucket;
    if (q.smem_tables) {
        uint32_t *thr_s = reinterpret_cast<uint32_t *>(smem);
        for (uint32_t i = threadIdx.x; i < q.thr_count; i += blockDim.x)
            thr_s[i] = q.thr[i];
        s.thr = thr_s;
        if (q.search_mode == SEARCH_BUCKET) {
            /* bucket heads are u16; copy them as u32 pairs (table is padded to an even count) */
            uint32_t *b_s = thr_s + q.thr_count;
            const uint32_t *b_g = reinterpret_cast<const uint32_t *>(q.bucket);
            const uint32_t n32 = (q.nbm1 + 2) >> 1;
            for (uint32_t i = threadIdx.x; i < n32; i += blockDim.x)
                b_s[i] = b_g[i];
            s.bucket = reinterpret_cast<const uint16_t *>(b_s);
        }
        __syncthreads();
    }
    return s;
}

__device__ __forceinline__ void atomic_noop() {}

/* Block-wide reduction of the per-thread plane-0 statistics; the last block of a
 * frame folds all block partials in a fixed order (deterministic result). */
__device__ __forceinline__ void finish_stats(const EncArgs &a, uint32_t frame, double sum, float mx, float mn)
{
    __shared__ double s_sum[kThreads / 32];
    __shared__ float s_mx[kThreads / 32];
    __shared__ float s_mn[kThreads / 32];
    __shared__ uint32_t s_last;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, o);
        mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_down_sync(0xffffffffu, mn, o));
    }
    if (lane == 0) {
        s_sum[warp] = sum;
        s_mx[warp] = mx;
        s_mn[warp] = mn;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < kThreads / 32; ++i) {
            sum += s_sum[i];
            mx = fmaxf(mx, s_mx[i]);
            mn = fminf(mn, s_mn[i]);
        }
        StatsPartial p;
        p.sum = sum;
        p.mx = mx;
        p.mn = mn;
        a.partial[(size_t)frame * gridDim.x + blockIdx.x] = p;
        __threadfence();
        s_last = (atomicAdd(&a.counter[frame], 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last && warp == 0) {
        __threadfence();
        double t = 0.0;
        float tx = -INFINITY, tn = INFINITY;
        const volatile StatsPartial *pp = a.partial + (size_t)frame * gridDim.x;
        for (uint32_t i = lane; i < gridDim.x; i += 32) {
            t += pp[i].sum;
            tx = fmaxf(tx, pp[i].mx);
            tn = fminf(tn, pp[i].mn);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            t += __shfl_down_sync(0xffffffffu, t, o);
            tx = fmaxf(tx, __shfl_down_sync(0xffffffffu, tx, o));
            tn = fminf(tn, __shfl_down_sync(0xffffffffu, tn, o));
        }
        if (lane == 0) {
            a.stats[frame].sum = t;
            a.stats[frame].mx = tx;
            a.stats[frame].mn = tn;
            a.counter[frame] = 0; /* self-cleaning for the next launch */
        }
    }
}

/* pack four codes of one row into the plane's sample container */
template <int BYTES>
__device__ __forceinline__ void store_codes4(uint8_t *row, uint32_t x, const uint32_t c[4], bool vec, uint32_t nvalid)
{
    if (BYTES == 2) {
        if (vec) {
            uint2 v;
            v.x = (c[0] & 0xffffu) | (c[1] << 16);
            v.y = (c[2] & 0xffffu) | (c[3] << 16);
            __stcs(reinterpret_cast<uint2 *>(row + 2 * (size_t)x), v);
        } else {
            for (uint32_t i = 0; i < nvalid; ++i) {
                row[2 * (size_t)(x + i)] = (uint8_t)(c[i] & 0xffu);
                row[2 * (size_t)(x + i) + 1] = (uint8_t)((c[i] >> 8) & 0xffu);
            }
        }
    } else {
        if (vec) {
            uint32_t v = (c[0] & 0xffu) | ((c[1] & 0xffu) << 8) | ((c[2] & 0xffu) << 16) | (c[3] << 24);
            __stcs(reinterpret_cast<uint32_t *>(row + x), v);
        } else {
            for (uint32_t i = 0; i < nvalid; ++i)
                row[x + i] = (uint8_t)(c[i] & 0xffu);
        }
    }
}

template <int BYTES>
__device__ __forceinline__ void store_codes2(uint8_t *row, uint32_t x, const uint32_t c[2], bool vec, uint32_t nvalid)
{
    if (BYTES == 2) {
        if (vec) {
            __stcs(reinterpret_cast<uint32_t *>(row + 2 * (size_t)x), (c[0] & 0xffffu) | (c[1] << 16));
        } else {
            for (uint32_t i = 0; i < nvalid; ++i) {
                row[2 * (size_t)(x + i)] = (uint8_t)(c[i] & 0xffu);
                row[2 * (size_t)(x + i) + 1] = (uint8_t)((c[i] >> 8) & 0xffu);
            }
        }
    } else {
        if (vec) {
            *reinterpret_cast<uint16_t *>(row + x) = (uint16_t)((c[0] & 0xffu) | ((c[1] & 0xffu) << 8));
        } else {
            for (uint32_t i = 0; i < nvalid; ++i)
                row[x + i] = (uint8_t)(c[i] & 0xffu);
        }
    }
}

template <int BYTES>
__device__ __forceinline__ void load_codes4(const uint8_t *row, uint32_t x, uint32_t c[4], bool vec, uint32_t nvalid)
{
    if (BYTES == 2) {
        if (vec) {
            uint2 v = __ldcs(reinterpret_cast<const uint2 *>(row + 2 * (size_t)x));
            c[0] = v.x & 0xffffu;
            c[1] = v.x >> 16;
            c[2] = v.y & 0xffffu;
            c[3] = v.y >> 16;
        } else {
            for (uint32_t i = 0; i < 4; ++i)
                c[i] = (i < nvalid) ? ((uint32_t)row[2 * (size_t)(x + i)] | ((uint32_t)row[2 * (size_t)(x + i) + 1] << 8))
                                    : 0u;
        }
    } else {
        if (vec) {
            uint32_t v = __ldcs(reinterpret_cast<const uint32_t *>(row + x));
            c[0] = v & 0xffu;
            c[1] = (v >> 8) & 0xffu;
            c[2] = (v >> 16) & 0xffu;
            c[3] = v >> 24;
        } else {
            for (uint32_t i = 0; i < 4; ++i)
                c[i] = (i < nvalid) ? (uint32_t)row[x + i] : 0u;
        }
    }
}

template <int BYTES>
__device__ __forceinline__ void load_codes2(const uint8_t *row, uint32_t x, uint32_t c[2], bool vec, uint32_t nvalid)
{
    if (BYTES == 2) {
        if (vec) {
            uint32_t v = __ldcs(reinterpret_cast<const uint32_t *>(row + 2 * (size_t)x));
            c[0] = v & 0xffffu;
            c[1] = v >> 16;
        } else {
            for (uint32_t i = 0; i < 2; ++i)
                c[i] = (i < nvalid) ? ((uint32_t)row[2 * (size_t)(x + i)] | ((uint32_t)row[2 * (size_t)(x + i) + 1] << 8))
                                    : 0u;
        }
    } else {
        if (vec) {
            uint32_t v = __ldcs(reinterpret_cast<const uint16_t *>(row + x));
            c[0] = v & 0xffu;
            c[1] = (v >> 8) & 0xffu;
        } else {
            for (uint32_t i = 0; i < 2; ++i)
                c[i] = (i < nvalid) ? (uint32_t)row[x + i] : 0u;
        }
    }
}

/* =============================== encode ========================================= */
template <int CS, bool SUB, int BYTES, bool VEC>
__global__ void __launch_bounds__(kThreads) encode_kernel(const EncArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (CS == CS_YCBCR)
        powf_tables_stage();
    const SearchCtx s = make_search_ctx(a.q, smem_raw);
    constexpr bool LUT_ALL = (CS == CS_RGB || CS == CS_XYZ); /* every plane goes through the LUT search */
    /* always the full key transform with its explicit NaN test: the shortcut `bits ^ 0x80000000` for
     * positive values is compiled into a float negation (FADD -x, -0), which canonicalises a NaN instead
     * of flipping its sign bit and would send NaN to code 0 instead of max_val */
    constexpr bool POS = false;

    const uint32_t frame = blockIdx.y;
    const float *rgb = a.rgb + (size_t)frame * a.rgb_frame_stride;
    float *rgb_out = a.rgb_out ? a.rgb_out + (size_t)frame * a.out_frame_stride : nullptr;
    uint8_t *pl0 = a.plane[0] + (size_t)frame * a.plane_frame_stride[0];
    uint8_t *pl1 = a.plane[1] + (size_t)frame * a.plane_frame_stride[1];
    uint8_t *pl2 = a.plane[2] + (size_t)frame * a.plane_frame_stride[2];
    const uint32_t w = a.w, h = a.h;
    const float max_c = a.q.max_val_color_f;
    const float l_max = a.q.l_max;

    const uint32_t tpr = (w + 3) >> 2;
    const uint32_t ntiles = tpr * ((h + 1) >> 1);

    double sum = 0.0;
    float mx = -INFINITY, mn = INFINITY;

    for (uint32_t t = blockIdx.x * kThreads + threadIdx.x; t < ntiles; t += gridDim.x * kThreads) {
        const uint32_t ty = t / tpr;
        const uint32_t x0 = (t - ty * tpr) * 4u, y0 = ty * 2u;
        const uint32_t nx = VEC ? 4u : min(4u, w - x0);
        const uint32_t ny = VEC ? 2u : min(2u, h - y0);

        float c[3][2][4];
        /* ---- load: 6 x 128-bit per thread, contiguous 512 B per warp and row-plane */
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float *src = rgb + (size_t)p * a.rgb_plane_stride + (size_t)(y0 + r) * w + x0;
                if (VEC) {
                    float4 v = ld_stream4(src);
                    c[p][r][0] = v.x;
                    c[p][r][1] = v.y;
                    c[p][r][2] = v.z;
                    c[p][r][3] = v.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        c[p][r][i] = ((uint32_t)r < ny && (uint32_t)i < nx) ? src[i] : 1.0f;
                }
            }
        }

        /* ---- colour transform, per pixel */
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (a.passthrough)
                    continue; /* LumaEncoder::setChannels: the caller already ran transformColorSpace */
                float R = c[0][r][i], G = c[1][r][i], B = c[2][r][i];
                if (a.prescale) {
                    R = __fmul_rn(R, a.sc);
                    G = __fmul_rn(G, a.sc);
                    B = __fmul_rn(B, a.sc);
                }
                color_forward<CS>(R, G, B, l_max, c[0][r][i], c[1][r][i], c[2][r][i]);
            }
        }

        /* ---- optional write-back of the transformed frame (reference side effect) */
        if (rgb_out) {
#pragma unroll
            for (int p = 0; p < 3; ++p) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    float *dst = rgb_out + (size_t)p * a.out_plane_stride + (size_t)(y0 + r) * w + x0;
                    if (VEC) {
                        st_stream4(dst, make_float4(c[p][r][0], c[p][r][1], c[p][r][2], c[p][r][3]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if ((uint32_t)r < ny && (uint32_t)i < nx)
                                dst[i] = c[p][r][i];
                    }
                }
            }
        }

        /* ---- plane-0 statistics (src/luma_encoder.cpp:276,294,314-316) */
        if (a.stats) {
            float part[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (VEC) {
                    part[r] = (c[0][r][0] + c[0][r][1]) + (c[0][r][2] + c[0][r][3]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        mx = fmaxf(mx, c[0][r][i]);
                        mn = fminf(mn, c[0][r][i]);
                    }
                } else {
                    part[r] = 0.0f;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if ((uint32_t)r < ny && (uint32_t)i < nx) {
                            part[r] += c[0][r][i];
                            mx = fmaxf(mx, c[0][r][i]);
                            mn = fminf(mn, c[0][r][i]);
                        }
                }
            }
            sum += (double)part[0] + (double)part[1];
        }

        /* ---- plane 0: LUT search, pack, store */
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            uint32_t code[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                code[i] = search_code<POS>(s, c[0][r][i]);
            if (VEC || (uint32_t)r < ny)
                store_codes4<BYTES>(pl0 + (size_t)(y0 + r) * a.stride[0], x0, code, VEC, nx);
        }

        /* ---- planes 1, 2 */
        if (SUB) {
#pragma unroll
            for (int p = 1; p < 3; ++p) {
                uint32_t code[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    /* 0.25f*(((a+b)+c)+d), src/luma_encoder.cpp:287-290 */
                    float m = __fmul_rn(
                        0.25f, __fadd_rn(__fadd_rn(__fadd_rn(c[p][0][2 * j], c[p][0][2 * j + 1]), c[p][1][2 * j]),
                                         c[p][1][2 * j + 1]));
                    code[j] = LUT_ALL ? search_code<POS>(s, m) : quantize_chroma(m, max_c);
                }
                uint8_t *row = (p == 1 ? pl1 : pl2) + (size_t)ty * a.stride[p];
                store_codes2<BYTES>(row, x0 >> 1, code, VEC, (nx + 1) >> 1);
            }
        } else {
#pragma unroll
            for (int p = 1; p < 3; ++p) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    uint32_t code[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        code[i] = LUT_ALL ? search_code<POS>(s, c[p][r][i]) : quantize_chroma(c[p][r][i], max_c);
                    if (VEC || (uint32_t)r < ny)
                        store_codes4<BYTES>((p == 1 ? pl1 : pl2) + (size_t)(y0 + r) * a.stride[p], x0, code, VEC, nx);
                }
            }
        }
    }

    if (a.stats)
        finish_stats(a, frame, sum, mx, mn);
}

/* Display tail of the reference's player (src/lumaplay_dequantizer.frag:141-156) for one colour component:
 *   ldrSim: v = exposure * max(1, min(256, floor(256 v / scaling))) / 256      else: v = v * exposure / scaling
 *   doTmo:  v = v^0.8 / (v^0.8 + 0.8^0.8)                                      (sigmoid tone curve, n = sig = 0.8)
 *   out    = v^(1/gamma), written to an 8-bit unorm framebuffer: clamp to [0, 1], round(255 v); NaN -> 0.
 * Plain fp32 with CUDA's powf: like the GLSL original this is not a bit-exact path (tests allow 1 LSB). */
__device__ __forceinline__ uint32_t display_unorm8(float v, const DecArgs &a)
{
    if (a.disp_ldr)
        v = a.disp_exposure * fmaxf(1.0f, fminf(256.0f, floorf(256.0f * v / a.disp_scaling))) / 256.0f;
    else
        v = v * a.disp_exposure / a.disp_scaling;
    if (a.disp_tmo) {
        const float t = powf(v, 0.8f);
        v = t / (t + 0.83651630373780790557f); /* 0.8^0.8 */
    }
    v = powf(v, a.disp_inv_gamma);
    v = fminf(fmaxf(v, 0.0f), 1.0f); /* fmaxf drops NaN -> 0 */
    return (uint32_t)__float2int_rn(v * 255.0f);
}

/* =============================== decode ========================================= */
template <int CS, bool SUB, int BYTES, bool VEC>
__global__ void __launch_bounds__(kThreads) decode_kernel(const DecArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (CS == CS_YCBCR)
        powf_tables_stage();
    constexpr bool LUT_ALL = (CS == CS_RGB || CS == CS_XYZ);

    const float *lut = a.q.lut;
    if (a.q.smem_lut) {
        float *lut_s = reinterpret_cast<float *>(smem_raw);
        for (uint32_t i = threadIdx.x; i <= a.q.max_val; i += blockDim.x)
            lut_s[i] = a.q.lut[i];
        __syncthreads();
        lut = lut_s;
    }

    const uint32_t frame = blockIdx.y;
    const uint8_t *pl0 = a.plane[0] + (size_t)frame * a.plane_frame_stride[0];
    const uint8_t *pl1 = a.plane[1] + (size_t)frame * a.plane_frame_stride[1];
    const uint8_t *pl2 = a.plane[2] + (size_t)frame * a.plane_frame_stride[2];
    float *rgb = a.rgb + (size_t)frame * a.rgb_frame_stride;
    const uint32_t w = a.w, h = a.h;
    const uint32_t max_val = a.q.max_val;
    const float max_c = a.q.max_val_color_f;
    const float l_max = a.q.l_max;

    const uint32_t tpr = (w + 3) >> 2;
    const uint32_t ntiles = tpr * ((h + 1) >> 1);

    for (uint32_t t = blockIdx.x * kThreads + threadIdx.x; t < ntiles; t += gridDim.x * kThreads) {
        const uint32_t ty = t / tpr;
        const uint32_t x0 = (t - ty * tpr) * 4u, y0 = ty * 2u;
        const uint32_t nx = VEC ? 4u : min(4u, w - x0);
        const uint32_t ny = VEC ? 2u : min(2u, h - y0);

        uint32_t code0[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (VEC || (uint32_t)r < ny)
                load_codes4<BYTES>(pl0 + (size_t)(y0 + r) * a.stride[0], x0, code0[r], VEC, nx);
            else
                code0[r][0] = code0[r][1] = code0[r][2] = code0[r][3] = 0u;
        }

        /* chroma: one ChromaInv per 2x2 block (SUB) or per pixel */
        ChromaInv chr[2][4];
        if (SUB) {
            uint32_t cc[2][2];
            load_codes2<BYTES>(pl1 + (size_t)ty * a.stride[1], x0 >> 1, cc[0], VEC, (nx + 1) >> 1);
            load_codes2<BYTES>(pl2 + (size_t)ty * a.stride[2], x0 >> 1, cc[1], VEC, (nx + 1) >> 1);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float c1, c2;
                if (LUT_ALL) {
                    c1 = lut[min(cc[0][j], max_val)];
                    c2 = lut[min(cc[1][j], max_val)];
                } else {
                    c1 = dequantize_chroma((float)cc[0][j], max_c);
                    c2 = dequantize_chroma((float)cc[1][j], max_c);
                }
                ChromaInv ci;
                if (a.passthrough) {
                    ci.a = c1;
                    ci.b = c2;
                } else {
                    ci = chroma_inverse<CS>(c1, c2);
                }
                chr[0][2 * j] = chr[0][2 * j + 1] = chr[1][2 * j] = chr[1][2 * j + 1] = ci;
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                uint32_t cc1[4] = {0u, 0u, 0u, 0u}, cc2[4] = {0u, 0u, 0u, 0u};
                if (VEC || (uint32_t)r < ny) {
                    load_codes4<BYTES>(pl1 + (size_t)(y0 + r) * a.stride[1], x0, cc1, VEC, nx);
                    load_codes4<BYTES>(pl2 + (size_t)(y0 + r) * a.stride[2], x0, cc2, VEC, nx);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float c1, c2;
                    if (LUT_ALL) {
                        c1 = lut[min(cc1[i], max_val)];
                        c2 = lut[min(cc2[i], max_val)];
                    } else {
                        c1 = dequantize_chroma((float)cc1[i], max_c);
                        c2 = dequantize_chroma((float)cc2[i], max_c);
                    }
                    if (a.passthrough) {
                        chr[r][i].a = c1;
                        chr[r][i].b = c2;
                    } else {
                        chr[r][i] = chroma_inverse<CS>(c1, c2);
                    }
                }
            }
        }

        float o[3][2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float c0 = lut[min(code0[r][i], max_val)];
                float R, G, B;
                if (a.passthrough) { /* LumaDecoder::getVpxChannels only */
                    o[0][r][i] = c0;
                    o[1][r][i] = chr[r][i].a;
                    o[2][r][i] = chr[r][i].b;
                    continue;
                }
                color_inverse<CS>(c0, chr[r][i], l_max, R, G, B);
                if (a.prescale) {
                    R = __fdiv_rn(R, a.sc);
                    G = __fdiv_rn(G, a.sc);
                    B = __fdiv_rn(B, a.sc);
                }
                o[0][r][i] = R;
                o[1][r][i] = G;
                o[2][r][i] = B;
            }
        }

        if (a.rgba) { /* display mode: tone curve + gamma, 8-bit RGBA out */
            uint8_t *img = a.rgba + (size_t)frame * a.rgba_frame_stride;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                uint32_t px[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    px[i] = display_unorm8(o[0][r][i], a) | (display_unorm8(o[1][r][i], a) << 8) |
                            (display_unorm8(o[2][r][i], a) << 16) | 0xff000000u;
                uint32_t *dst = reinterpret_cast<uint32_t *>(img + (size_t)(y0 + r) * a.rgba_pitch) + x0;
                if (VEC && (a.rgba_pitch & 15) == 0) {
                    __stcs(reinterpret_cast<uint4 *>(dst), make_uint4(px[0], px[1], px[2], px[3]));
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if ((uint32_t)r < ny && (uint32_t)i < nx)
                            dst[i] = px[i];
                }
            }
            continue;
        }

#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                float *dst = rgb + (size_t)p * a.rgb_plane_stride + (size_t)(y0 + r) * w + x0;
                if (VEC) {
                    st_stream4(dst, make_float4(o[p][r][0], o[p][r][1], o[p][r][2], o[p][r][3]));
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if ((uint32_t)r < ny && (uint32_t)i < nx)
                            dst[i] = o[p][r][i];
                }
            }
        }
    }
}

#ifdef LUMA_TU_ELEMENTWISE
/* ====================== element-wise API kernels (one translation unit only) ===== */
/* LumaQuantizer::transformColorSpace in place on a planar frame of n pixels. */
template <int CS, bool FWD>
__global__ void __launch_bounds__(kThreads) transform_kernel(float *c0, float *c1, float *c2, size_t n, float sc, float l_max)
{
    if (CS == CS_YCBCR)
        powf_tables_stage();
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        float a = c0[i], b = c1[i], c = c2[i];
        float x, y, z;
        if (FWD) {
            color_forward<CS>(__fmul_rn(a, sc), __fmul_rn(b, sc), __fmul_rn(c, sc), l_max, x, y, z);
        } else {
            ChromaInv ci = chroma_inverse<CS>(b, c);
            color_inverse<CS>(a, ci, l_max, x, y, z);
            x = __fdiv_rn(x, sc);
            y = __fdiv_rn(y, sc);
            z = __fdiv_rn(z, sc);
        }
        c0[i] = x;
        c1[i] = y;
        c2[i] = z;
    }
}

/* LumaQuantizer::quantize over an array (codes returned as floats). */
__global__ void __launch_bounds__(kThreads) quantize_kernel(const QuantDev q, const float *in, float *out, size_t n, int use_lut)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SearchCtx s = make_search_ctx(q, smem_raw);
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const float v = in[i];
        out[i] = use_lut ? (float)search_code<false>(s, v) : (float)quantize_chroma(v, q.max_val_color_f);
    }
}

/* LumaQuantizer::dequantize over an array (src/luma_quantizer.cpp:247-264). */
__global__ void __launch_bounds__(kThreads) dequantize_kernel(const QuantDev q, const float *in, float *out, size_t n, int use_lut)
{
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const float v = in[i];
        float r;
        if (use_lut) {
            if (v < 0.0f)
                r = q.lut[0];
            else if (v >= q.max_val_f)
                r = q.lut[q.max_val];
            else if (v == v)
                r = q.lut[(int)v];
            else
                r = q.lut[0]; /* NaN index: undefined behaviour in the reference; pinned to entry 0 here */
        } else {
            r = dequantize_chroma(v, q.max_val_color_f);
        }
        out[i] = r;
    }
}

/* ---- frame sources on the device ------------------------------------------------------ */
/* ExrInterface::testFrame (src/exr_interface.cpp:50-70): the reference's synthetic HDR pattern, written straight
 * into HBM.  Integer sub-expressions are size_t / unsigned divisions in the reference; every float operation is a
 * separately rounded IEEE mul or div in the reference's left-to-right order, e.g.
 * 10000.0f*((float)(x*x))/(w*w) = ((1e4f * float(x*x)) / float(w*w)) with w*w an unsigned int product. */
__global__ void __launch_bounds__(kThreads) test_frame_kernel(float *rgb, uint32_t w, uint32_t h)
{
    const size_t n = (size_t)w * h;
    const float ww = (float)(uint32_t)(w * w), hh = (float)(uint32_t)(h * h); /* unsigned int products, as in the reference */
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const unsigned long long y = i / w, x = i - y * w;
        float r, g, b;
        if (y < h / 5) {
            const float v = (y < h / 10) ? __fdiv_rn(__fmul_rn(10000.0f, (float)(x * x)), ww)
                                         : __fdiv_rn(__fmul_rn(10000.0f, (float)((20ull * x) / w)), 20.0f);
            r = g = b = v;
        } else {
            const unsigned long long band = (20ull * y / h) % 2ull, col = (30ull * x / w) % 2ull;
            r = __fmul_rn(10000.0f, (float)(band ^ col));
            const float on = __fmul_rn(10000.0f, (float)band);
            g = __fdiv_rn(__fmul_rn(on, (float)(y * y)), hh);
            b = __fdiv_rn(__fmul_rn(on, (float)(x * x)), ww);
        }
        __stcs(rgb + i, r);
        __stcs(rgb + n + i, g);
        __stcs(rgb + 2 * n + i, b);
    }
}

/* ExrInterface::readFrame's pixel loop (src/exr_interface.cpp:73-143): interleaved half-float RGBA (Imf::Rgba, 8
 * bytes per pixel) to the planar f32 LumaFrame layout; mode = Imf::RgbaChannels of the file: WRITE_R (1), WRITE_G
 * (2), WRITE_B (4) replicate that channel into all three planes, WRITE_RGB (7) / WRITE_RGBA (15) copy r, g, b.
 * half -> float is exact.  One thread converts two pixels (one 128-bit load). */
__device__ __forceinline__ float half_bits_to_float(uint32_t h16)
{
    return __half2float(__ushort_as_half((unsigned short)h16));
}
__global__ void __launch_bounds__(kThreads) half_rgba_to_frame_kernel(const uint2 *rgba, float *rgb, size_t n, size_t plane_stride, int mode)
{
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const uint2 v = __ldcs(rgba + i);
        float r = half_bits_to_float(v.x & 0xffffu), g = half_bits_to_float(v.x >> 16), b = half_bits_to_float(v.y & 0xffffu);
        if (mode == 1)
            g = b = r;
        else if (mode == 2)
            r = b = g;
        else if (mode == 4)
            r = g = b;
        __stcs(rgb + i, r);
        __stcs(rgb + plane_stride + i, g);
        __stcs(rgb + 2 * plane_stride + i, b);
    }
}

/* ExrInterface::writeFrame's pixel loop (src/exr_interface.cpp:157-187): planar f32 frame -> interleaved Imf::Rgba
 * pixels, p.r = float (Imf's half(float): round to nearest even, overflow to infinity, a NaN keeps its sign and its top
 * ten payload bits with the lowest forced to 1 if they are all zero -- OpenEXR is not part of the reference tree, "parity
 * unpinned"; numpy's float32 -> float16 cast is the same function and is what the test compares with), p.a = 0. */
__device__ __forceinline__ uint32_t float_to_half_bits(float f)
{
    const uint32_t u = __float_as_uint(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) { /* NaN */
        const uint32_t m = (u & 0x007fffffu) >> 13;
        return ((u >> 16) & 0x8000u) | 0x7c00u | m | (m == 0u ? 1u : 0u);
    }
    return (uint32_t)__half_as_ushort(__float2half_rn(f));
}
__global__ void __launch_bounds__(kThreads) frame_to_half_rgba_kernel(const float *__restrict__ rgb, uint2 *__restrict__ rgba, size_t n,
                                                                      size_t plane_stride)
{
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const float r = __ldcs(rgb + i), g = __ldcs(rgb + plane_stride + i), b = __ldcs(rgb + 2 * plane_stride + i);
        uint2 v;
        v.x = float_to_half_bits(r) | (float_to_half_bits(g) << 16);
        v.y = float_to_half_bits(b); /* alpha = half(0) */
        __stcs(rgba + i, v);
    }
}

/* PfsInterface::readFrame / writeFrame (src/pfs_interface.cpp:57-113, :115-152): a PFS stream carries the colour
 * channels X, Y, Z as three separate float arrays; the reference converts them with pfstools'
 * pfs::transformColorSpace(CS_XYZ -> CS_RGB) (:84) -- resp. CS_RGB -> CS_XYZ (:140) -- and memcpy's the three channels
 * into the LumaFrame planes (:100-102).  pfstools (libpfs, src/pfs/colorspace.cpp, NOT part of the reference tree:
 * "parity unpinned") multiplies every pixel by its D65 matrices xyz2rgbD65Mat / rgb2xyzD65Mat, which hold the same
 * nine constants as the reference's own xyz2rgbMat / rgb2xyzMat (include/luma/luma_quantizer.h:79-87), evaluated
 * m0*a + m1*b + m2*c left to right in float, no clamp.  TO_RGB selects the direction. */
template <bool TO_RGB>
__global__ void __launch_bounds__(kThreads) pfs_channels_kernel(const float *__restrict__ a0, const float *__restrict__ a1,
                                                                const float *__restrict__ a2, float *__restrict__ o0,
                                                                float *__restrict__ o1, float *__restrict__ o2, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const float a = __ldcs(a0 + i), b = __ldcs(a1 + i), c = __ldcs(a2 + i);
        float x, y, z;
        if (TO_RGB) {
            x = dot3(LUMA_I00, LUMA_I01, LUMA_I02, a, b, c);
            y = dot3(LUMA_I10, LUMA_I11, LUMA_I12, a, b, c);
            z = dot3(LUMA_I20, LUMA_I21, LUMA_I22, a, b, c);
        } else {
            x = dot3(LUMA_M00, LUMA_M01, LUMA_M02, a, b, c);
            y = dot3(LUMA_M10, LUMA_M11, LUMA_M12, a, b, c);
            z = dot3(LUMA_M20, LUMA_M21, LUMA_M22, a, b, c);
        }
        __stcs(o0 + i, x);
        __stcs(o1 + i, y);
        __stcs(o2 + i, z);
    }
}

/* ---- the player's own sampling (lumacu_display_params.filter = 1) --------------------------------------------
 * The reference's player draws the decoder's planes as GL_LINEAR textures (lumaplay.cpp:258-259 sets MIN/MAG filter
 * GL_LINEAR and CLAMP_TO_EDGE for every texture, the LUT included) and dequantises in the fragment shader
 * (src/lumaplay_dequantizer.frag:70-157).  At 1:1 scale that means, for output pixel (x, y):
 *   - plane 0 is sampled at its own texel centres: the code itself;
 *   - a half-resolution chroma plane is sampled at u = x/2 - 1/4, v = y/2 - 1/4 (texel units): bilinear weights 1/4 : 3/4
 *     towards the nearer chroma sample in each direction, clamped at the plane's edges;
 *   - the LUT texture is uploaded with getSize() = maxVal texels (one short, lumaplay.cpp:371) and fetched at
 *     code / maxVal: texel coordinate code - 1/2, i.e. the MEAN of lut[code-1] and lut[code] (lut[0] for code 0,
 *     lut[maxVal-1] for code >= maxVal);
 *   - chroma is code / maxValColor without the CPU path's 1e-10 floor; the PQ used for YCbCr has L = 10000 built in.
 * One thread per output pixel; plain fp32 like the shader (approximate by nature). */
__device__ __forceinline__ float disp_code(const uint8_t *pl, int32_t stride, int bytes, uint32_t x, uint32_t y)
{
    const uint8_t *p = pl + (size_t)y * stride + (size_t)x * bytes;
    return bytes == 2 ? (float)(p[0] | (p[1] << 8)) : (float)p[0];
}
__device__ __forceinline__ float disp_lut_linear(const float *lut, uint32_t max_val, float code)
{
    /* texture1D(texM, clamp(code, 0, maxVal) / maxVal) on a maxVal-texel GL_LINEAR, CLAMP_TO_EDGE texture */
    const float t = fminf(fmaxf(code, 0.0f), (float)max_val) - 0.5f;
    const float fl = floorf(t), f = t - fl;
    const int last = (int)max_val - 1;
    const int i0 = min(max((int)fl, 0), last), i1 = min(max((int)fl + 1, 0), last);
    return lut[i0] + f * (lut[i1] - lut[i0]);
}
__device__ __forceinline__ float disp_pq(float val, bool encode)
{
    const float L = 10000.0f, m = 78.8438f, n = 0.1593f, c1 = 0.8359f, c2 = 18.8516f, c3 = 18.6875f;
    if (encode) {
        const float Lp = powf(val / L, n);
        return powf((c1 + c2 * Lp) / (1.0f + c3 * Lp), m);
    }
    const float Vp = powf(val, 1.0f / m);
    return L * powf(fmaxf(0.0f, Vp - c1) / (c2 - c3 * Vp), 1.0f / n);
}
__global__ void __launch_bounds__(kThreads) display_linear_kernel(const DecArgs a, int cs, int sub, int bytes)
{
    const size_t n = (size_t)a.w * a.h;
    const uint32_t frame = blockIdx.y;
    const uint8_t *pl0 = a.plane[0] + (size_t)frame * a.plane_frame_stride[0];
    const uint8_t *pl1 = a.plane[1] + (size_t)frame * a.plane_frame_stride[1];
    const uint8_t *pl2 = a.plane[2] + (size_t)frame * a.plane_frame_stride[2];
    uint8_t *out = a.rgba + (size_t)frame * a.rgba_frame_stride;
    const uint32_t cw = sub ? (a.w + 1) >> 1 : a.w, chh = sub ? (a.h + 1) >> 1 : a.h;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const uint32_t y = (uint32_t)(i / a.w), x = (uint32_t)(i - (size_t)y * a.w);
        float c0 = disp_code(pl0, a.stride[0], bytes, x, y), c1, c2;
        if (sub) {
            const float u = 0.5f * (float)x - 0.25f, v = 0.5f * (float)y - 0.25f;
            const float uf = floorf(u), vf = floorf(v), fx = u - uf, fy = v - vf;
            const uint32_t x0 = (uint32_t)max((int)uf, 0), x1 = min((uint32_t)((int)uf + 1), cw - 1);
            const uint32_t y0 = (uint32_t)max((int)vf, 0), y1 = min((uint32_t)((int)vf + 1), chh - 1);
            const float p00 = disp_code(pl1, a.stride[1], bytes, x0, y0), p10 = disp_code(pl1, a.stride[1], bytes, x1, y0);
            const float p01 = disp_code(pl1, a.stride[1], bytes, x0, y1), p11 = disp_code(pl1, a.stride[1], bytes, x1, y1);
            const float q00 = disp_code(pl2, a.stride[2], bytes, x0, y0), q10 = disp_code(pl2, a.stride[2], bytes, x1, y0);
            const float q01 = disp_code(pl2, a.stride[2], bytes, x0, y1), q11 = disp_code(pl2, a.stride[2], bytes, x1, y1);
            const float pa = p00 + fx * (p10 - p00), pb = p01 + fx * (p11 - p01);
            const float qa = q00 + fx * (q10 - q00), qb = q01 + fx * (q11 - q01);
            c1 = pa + fy * (pb - pa);
            c2 = qa + fy * (qb - qa);
        } else {
            c1 = disp_code(pl1, a.stride[1], bytes, x, y);
            c2 = disp_code(pl2, a.stride[2], bytes, x, y);
        }
        c0 = disp_lut_linear(a.q.lut, a.q.max_val, c0);
        if (cs == CS_RGB || cs == CS_XYZ) {
            c1 = disp_lut_linear(a.q.lut, a.q.max_val, c1);
            c2 = disp_lut_linear(a.q.lut, a.q.max_val, c2);
        } else {
            c1 = c1 / a.q.max_val_color_f;
            c2 = c2 / a.q.max_val_color_f;
        }
        float R, G, B;
        if (cs == CS_LUV) {
            const float L = c0, u = c1 * 255.0f / 410.0f, v = c2 * 255.0f / 410.0f;
            const float den = 6.0f * u - 16.0f * v + 12.0f;
            const float xx = 9.0f * u / den, yy = 4.0f * v / den;
            const float Y = fmaxf(fminf(L, 100000000.0f), 0.0001f);
            const float X = fmaxf(fminf(xx / yy * L, 100000000.0f), 0.0001f);
            const float Z = fmaxf(fminf((1.0f - xx - yy) / yy * L, 100000000.0f), 0.0001f);
            R = LUMA_I00 * X + LUMA_I01 * Y + LUMA_I02 * Z;
            G = LUMA_I10 * X + LUMA_I11 * Y + LUMA_I12 * Z;
            B = LUMA_I20 * X + LUMA_I21 * Y + LUMA_I22 * Z;
        } else if (cs == CS_RGB) {
            R = c0, G = c1, B = c2;
        } else if (cs == CS_YCBCR) {
            float yv = disp_pq(c0, true);
            yv = (255.0f * yv - 16.0f) / 219.0f;
            B = yv + 1.8814f * (255.0f * c1 - 128.0f) / 224.0f;
            R = yv + 1.4746f * (255.0f * c2 - 128.0f) / 224.0f;
            G = (yv - 0.2627f * R - 0.0593f * B) / 0.6780f;
            R = disp_pq(fmaxf(0.0f, fminf(1.0f, R)), false);
            G = disp_pq(fmaxf(0.0f, fminf(1.0f, G)), false);
            B = disp_pq(fmaxf(0.0f, fminf(1.0f, B)), false);
        } else {
            R = LUMA_I00 * c0 + LUMA_I01 * c1 + LUMA_I02 * c2;
            G = LUMA_I10 * c0 + LUMA_I11 * c1 + LUMA_I12 * c2;
            B = LUMA_I20 * c0 + LUMA_I21 * c1 + LUMA_I22 * c2;
        }
        const uint32_t px = display_unorm8(R, a) | (display_unorm8(G, a) << 8) | (display_unorm8(B, a) << 16) | 0xff000000u;
        *reinterpret_cast<uint32_t *>(out + (size_t)y * a.rgba_pitch + (size_t)x * 4) = px;
    }
}

#endif /* LUMA_TU_ELEMENTWISE */

} // namespace lumacu

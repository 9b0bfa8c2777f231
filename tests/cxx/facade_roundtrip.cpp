/*
 * facade_roundtrip.cpp -- TEST DRIVER (test infrastructure, not product code).
 *
 * One translation unit, compiled twice by tests/cxx/Makefile with the SAME flags and the same
 * loopback codec / in-memory container test doubles (oracle/ref_stubs):
 *
 *   build/facade_roundtrip   against lumahdrv_b200/cxx (our LumaEncoder / LumaDecoder over the CUDA C ABI)
 *   build/ref_roundtrip      against the UNMODIFIED reference sources under $(REF)
 *
 * It only uses the reference's public API, the way test/test_simple_enc.cpp:27-69 and
 * test/test_simple_dec.cpp:17-40 do: setParams -> initialize -> encode(&frame) per frame -> finish, then
 * LumaDecoder(file) -> decode() until NULL, reading getBuffer()/getParams() for the integer planes.
 * Both binaries print one line per frame with FNV-1a hashes of the planes and of the decoded floats;
 * the GPU test asserts the two outputs are byte-identical.
 *
 *   usage: <bin> w h ptf cs ptfBits colorBits profile bitDepth preScaling nframes [maxLum minLum]
 */
#include <luma_decoder.h>
#include <luma_encoder.h>
#include <luma_exception.h>

#include "exr_interface.h"

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static uint32_t fnv(const void *p, size_t n, uint32_t h = 2166136261u)
{
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) {
        h ^= b[i];
        h *= 16777619u;
    }
    return h;
}

static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static uint64_t g_state;
static inline uint64_t xorshift64()
{
    g_state ^= g_state << 13;
    g_state ^= g_state >> 7;
    g_state ^= g_state << 17;
    return g_state;
}

/* frame 0 is the reference's test pattern; the others are noise spanning far more than the coded range,
 * with a few special values sprinkled in */
static void fill_frame(LumaFrame &frame, unsigned w, unsigned h, int index)
{
    if (index == 0) {
        ExrInterface::testFrame(frame, w, h);
        return;
    }
    frame.width = w;
    frame.height = h;
    frame.channels = 3;
    frame.init();
    g_state = 0x9E3779B97F4A7C15ull + (uint64_t)index;
    const size_t n = (size_t)3 * w * h;
    for (size_t i = 0; i < n; i++) {
        const uint64_t r = xorshift64();
        /* exponent -14 .. +17 (2^-14 = 6e-5 .. 2^17 = 1.3e5), 23 random mantissa bits */
        const uint32_t e = 127u - 14u + (uint32_t)((r >> 40) % 32u);
        const uint32_t bits = (e << 23) | (uint32_t)(r & 0x7FFFFFu);
        float v;
        memcpy(&v, &bits, 4);
        const unsigned sel = (unsigned)((r >> 56) & 0xFF);
        if (sel == 0)
            v = 0.0f;
        else if (sel == 1)
            v = -v;
        frame.buffer[i] = v;
    }
}

int main(int argc, char **argv)
{
    if (argc < 11) {
        fprintf(stderr, "usage: %s w h ptf cs ptfBits colorBits profile bitDepth preScaling nframes [maxLum minLum]\n", argv[0]);
        return 2;
    }
    const unsigned w = atoi(argv[1]), h = atoi(argv[2]);
    const int nframes = atoi(argv[10]);
    try {
        LumaEncoder encoder;
        LumaEncoderParams params = encoder.getParams();
        params.ptf = (LumaQuantizer::ptf_t)atoi(argv[3]);
        params.colorSpace = (LumaQuantizer::colorSpace_t)atoi(argv[4]);
        params.ptfBitDepth = atoi(argv[5]);
        params.colorBitDepth = atoi(argv[6]);
        params.profile = atoi(argv[7]);
        params.bitDepth = atoi(argv[8]);
        params.preScaling = (float)atof(argv[9]);
        if (argc > 12) {
            params.maxLum = (float)atof(argv[11]);
            params.minLum = (float)atof(argv[12]);
        }
        params.lossLess = 1;
        encoder.setParams(params);
        const char *file = "roundtrip.mkv"; /* lives in the in-memory container double */
        for (int f = 0; f < nframes; f++) {
            LumaFrame frame;
            fill_frame(frame, w, h, f);
            printf("in %d %08x\n", f, fnv(frame.buffer, (size_t)3 * w * h * sizeof(float)));
            if (!encoder.initialized())
                encoder.initialize(file, frame.width, frame.height);
            const double t0 = now_ms();
            encoder.encode(&frame);
            fprintf(stderr, "time encode %d %.3f ms\n", f, now_ms() - t0); /* stderr: stdout must stay comparable */
        }
        encoder.finish();

        /* the quantizer metadata as it went into the container (attachments 430..436, src/luma_encoder.cpp:78-106) */
        {
            MkvInterface reader;
            reader.openRead(file);
            binary *data = NULL;
            unsigned int id = 0, size = 0;
            for (unsigned int i = 0; reader.getAttachment(i, &data, id, size); i++)
                printf("att %u size %u %08x\n", id, size, fnv(data, size));
        }

        LumaDecoder decoder(file);
        for (int f = 0;; f++) {
            const double t0 = now_ms();
            LumaFrame *frame = decoder.decode();
            if (frame == NULL)
                break;
            fprintf(stderr, "time decode %d %.3f ms\n", f, now_ms() - t0);
            LumaDecoderParams dp = decoder.getParams();
            unsigned char **planes = decoder.getBuffer();
            uint32_t ph[3];
            for (int p = 0; p < 3; p++) {
                uint32_t hsh = 2166136261u;
                const size_t row = (size_t)dp.width[p] * (dp.highBitDepth ? 2 : 1);
                for (int y = 0; y < dp.height[p]; y++)
                    hsh = fnv(planes[p] + (size_t)y * dp.stride[p], row, hsh);
                ph[p] = hsh;
            }
            printf("frame %d %ux%u profile %d planes %08x %08x %08x floats %08x\n", f, frame->width, frame->height,
                   dp.profile, ph[0], ph[1], ph[2],
                   fnv(frame->buffer, (size_t)3 * frame->width * frame->height * sizeof(float)));
        }
        /* the scalar API of LumaQuantizer (include/luma/luma_quantizer.h:100-101) */
        LumaQuantizer *q = decoder.getQuantizer();
        const float probes[] = {-1.0f, 0.0f, 1e-4f, 0.005f, 1.0f, 100.0f, 1000.0f, 9999.0f, 1e5f};
        for (unsigned i = 0; i < sizeof(probes) / sizeof(probes[0]); i++)
            printf("q %g -> %g %g | dq -> %g\n", probes[i], q->quantize(probes[i], 0), q->quantize(probes[i] * 1e-4f, 1),
                   q->dequantize(q->quantize(probes[i], 0), 0));
        printf("size %u max %g min %g\n", q->getSize(), q->getMaxLum(), q->getMinLum());
    } catch (LumaException &e) {
        printf("LumaException: %s\n", e.what());
        return 1;
    }
    return 0;
}

/*
 * multi_gpu.cpp -- TEST DRIVER: one LumaEncoder + LumaDecoder pair per worker, frame f of the stream handled by
 * worker f mod N (frames are independent; nothing but the quantizer is shared).  Prints one hash line
 * per frame, in frame order; the output must not depend on N, on the mode, or on how many GPUs the box has.
 *
 *   usage: multi_gpu <n_workers> <n_frames> [w h [mode]]
 *
 * Worker g computes on CUDA device g mod (number of devices): on a one-GPU box N workers are N independent
 * contexts sharing the device, which exercises the same code paths.
 *
 *   mode "threads" (default)  one host thread per worker, through the reference's class API (LumaEncoder::encode,
 *                             LumaDecoder::decode); worker 0's quantizer -- host LUT and device search tables -- is
 *                             shared with the others by LumaQuantizer::broadcast (lumacu_broadcast_quantizer)
 *                             instead of every object deriving its own.
 *   mode "async"              ONE host thread drives all workers through the C ABI: lumacu_encode_async on every
 *                             context, then lumacu_wait on every context (same for decode); quantizer set on
 *                             context 0 and broadcast.
 *
 * Built a second time against the UNMODIFIED reference sources (-DLUMA_REFERENCE_BUILD, CPU, "threads" mode only):
 * that binary's output is the expected output.
 */
#include <luma_decoder.h>
#include <luma_encoder.h>
#include <luma_exception.h>

#ifndef LUMA_REFERENCE_BUILD
#include <lumacu.h>
#endif

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static uint32_t fnv(const void *p, size_t n, uint32_t h = 2166136261u)
{
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) {
        h ^= b[i];
        h *= 16777619u;
    }
    return h;
}

static void fill(LumaFrame &frame, unsigned w, unsigned h, int index)
{
    frame.width = w;
    frame.height = h;
    frame.channels = 3;
    frame.init();
    uint64_t s = 0x9E3779B97F4A7C15ull + (uint64_t)index;
    const size_t n = (size_t)3 * w * h;
    for (size_t i = 0; i < n; i++) {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        const uint32_t bits = ((127u - 8u + (uint32_t)((s >> 40) % 22u)) << 23) | (uint32_t)(s & 0x7FFFFFu);
        memcpy(&frame.buffer[i], &bits, 4);
    }
}

static int device_of(int worker)
{
#ifndef LUMA_REFERENCE_BUILD
    int n = 0;
    if (lumacu_device_count(&n) == LUMACU_OK && n > 0)
        return worker % n;
#endif
    (void)worker;
    return 0;
}

static std::string hash_line(int f, uint32_t planes, uint32_t floats)
{
    char buf[128];
    snprintf(buf, sizeof(buf), "frame %d planes %08x floats %08x", f, planes, floats);
    return buf;
}

/* ---- mode "threads": the reference's class API, one object pair per worker -------------------------------- */
static int run_threads(int n_workers, int n_frames, unsigned w, unsigned h, std::vector<std::string> &lines)
{
    std::vector<std::string> errors(n_workers);
    std::vector<LumaEncoder> encs(n_workers);
    std::vector<std::string> files(n_workers);
    try {
        for (int g = 0; g < n_workers; g++) {
            files[g] = "gpu" + std::to_string(g) + ".mkv"; /* in-memory container double */
#ifndef LUMA_REFERENCE_BUILD
            encs[g].setDevice(device_of(g));
#endif
            LumaEncoderParams p = encs[g].getParams();
            p.lossLess = 1;
            encs[g].setParams(p);
            encs[g].initialize(files[g].c_str(), w, h);
        }
#ifndef LUMA_REFERENCE_BUILD
        /* worker 0 derives the search tables once; the others receive them device to device */
        std::vector<LumaQuantizer *> qs;
        for (int g = 0; g < n_workers; g++)
            qs.push_back(encs[g].getQuantizer());
        LumaQuantizer::broadcast(&qs[0], n_workers, 0);
#endif
    } catch (LumaException &e) {
        printf("LumaException: %s\n", e.what());
        return 1;
    }
    std::vector<std::thread> workers;
    for (int g = 0; g < n_workers; g++)
        workers.emplace_back([&, g]() {
            try {
                LumaEncoder &enc = encs[g];
                std::vector<int> mine;
                for (int f = g; f < n_frames; f += n_workers)
                    mine.push_back(f);
                for (size_t i = 0; i < mine.size(); i++) {
                    LumaFrame frame;
                    fill(frame, w, h, mine[i]);
                    enc.encode(&frame);
                }
                enc.finish();
                if (mine.empty())
                    return;
                LumaDecoder dec;
#ifndef LUMA_REFERENCE_BUILD
                dec.setDevice(device_of(g));
#endif
                dec.initialize(files[g].c_str());
                for (size_t i = 0; i < mine.size(); i++) {
                    LumaFrame *out = dec.decode();
                    if (!out)
                        throw LumaException("decoder ran out of frames");
                    LumaDecoderParams dp = dec.getParams();
                    uint32_t ph = 2166136261u;
                    for (int pl = 0; pl < 3; pl++)
                        for (int y = 0; y < dp.height[pl]; y++)
                            ph = fnv(dec.getBuffer()[pl] + (size_t)y * dp.stride[pl], (size_t)dp.width[pl] * 2, ph);
                    lines[mine[i]] = hash_line(mine[i], ph, fnv(out->buffer, (size_t)3 * w * h * sizeof(float)));
                }
            } catch (LumaException &e) {
                errors[g] = e.what();
            }
        });
    for (size_t i = 0; i < workers.size(); i++)
        workers[i].join();
    for (int g = 0; g < n_workers; g++)
        if (!errors[g].empty()) {
            printf("worker %d: LumaException: %s\n", g, errors[g].c_str());
            return 1;
        }
    return 0;
}

#ifndef LUMA_REFERENCE_BUILD
/* ---- mode "async": one host thread, N contexts, asynchronous C ABI ----------------------------------------- */
#define CK(ctx, expr)                                                                  \
    do {                                                                               \
        const int rc_ = (expr);                                                        \
        if (rc_ != LUMACU_OK) {                                                        \
            printf("%s: %s: %s\n", #expr, lumacu_status_name(rc_), lumacu_last_error(ctx)); \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

static int run_async(int n_workers, int n_frames, unsigned w, unsigned h, std::vector<std::string> &lines)
{
    /* LumaEncoderParams defaults: PQ, Lu'v', 11/8 bit, profile 2, 0.005 .. 10000 cd/m2 (include/luma/luma_encoder.h:110-118) */
    std::vector<lumacu_ctx *> ctx(n_workers, (lumacu_ctx *)NULL);
    for (int g = 0; g < n_workers; g++)
        CK(NULL, lumacu_create(device_of(g), &ctx[g]));
    std::vector<float> lut(2048);
    CK(NULL, lumacu_build_lut(LUMACU_PTF_PQ, 11, 10000.0f, 0.005f, &lut[0], lut.size()));
    CK(ctx[0], lumacu_set_quantizer(ctx[0], &lut[0], 2048, 255, LUMACU_CS_LUV, 10000.0f));
    CK(ctx[0], lumacu_broadcast_quantizer(&ctx[0], n_workers, 0));

    const int32_t strides[3] = {(int32_t)(((w + 31) & ~31u) * 2), (int32_t)((w + 31) & ~31u), (int32_t)((w + 31) & ~31u)};
    const size_t psz[3] = {(size_t)strides[0] * h, (size_t)strides[1] * (h / 2), (size_t)strides[2] * (h / 2)};
    struct Slot {
        float *in, *out;
        uint8_t *planes[3];
        lumacu_frame_stats st;
    };
    std::vector<Slot> slot(n_workers);
    for (int g = 0; g < n_workers; g++) {
        void *p = NULL;
        CK(NULL, lumacu_host_alloc((size_t)3 * w * h * 4, &p));
        slot[g].in = (float *)p;
        CK(NULL, lumacu_host_alloc((size_t)3 * w * h * 4, &p));
        slot[g].out = (float *)p;
        for (int pl = 0; pl < 3; pl++) {
            CK(NULL, lumacu_host_alloc(psz[pl], &p));
            slot[g].planes[pl] = (uint8_t *)p;
            memset(p, 0, psz[pl]);
        }
    }
    for (int f0 = 0; f0 < n_frames; f0 += n_workers) {
        const int live = std::min(n_workers, n_frames - f0);
        for (int g = 0; g < live; g++) { /* queue an encode on every GPU ... */
            LumaFrame frame;
            fill(frame, w, h, f0 + g);
            memcpy(slot[g].in, frame.buffer, (size_t)3 * w * h * 4);
            CK(ctx[g], lumacu_encode_async(ctx[g], slot[g].in, w, h, 2, 1.0f, slot[g].planes, strides, 0, &slot[g].st));
            if (!lumacu_pending(ctx[g])) {
                printf("lumacu_encode_async left nothing pending\n");
                return 1;
            }
        }
        for (int g = 0; g < live; g++) { /* ... the input may be refilled once it has been read ... */
            CK(ctx[g], lumacu_wait_input(ctx[g]));
            memset(slot[g].in, 0xFF, 64); /* scribble: the transform must not read it any more */
        }
        for (int g = 0; g < live; g++) /* ... then collect */
            CK(ctx[g], lumacu_wait(ctx[g]));
        for (int g = 0; g < live; g++)
            CK(ctx[g], lumacu_decode_async(ctx[g], slot[g].planes, strides, w, h, 2, 1.0f, slot[g].out));
        for (int g = 0; g < live; g++)
            CK(ctx[g], lumacu_wait(ctx[g]));
        for (int g = 0; g < live; g++) {
            if (!(slot[g].st.sum > 0.0) || !(slot[g].st.max >= slot[g].st.min)) {
                printf("frame %d: statistics not filled by lumacu_wait\n", f0 + g);
                return 1;
            }
            uint32_t ph = 2166136261u;
            for (int pl = 0; pl < 3; pl++) {
                const unsigned pw = pl ? w / 2 : w, phh = pl ? h / 2 : h;
                for (unsigned y = 0; y < phh; y++)
                    ph = fnv(slot[g].planes[pl] + (size_t)y * strides[pl], (size_t)pw * 2, ph);
            }
            lines[f0 + g] = hash_line(f0 + g, ph, fnv(slot[g].out, (size_t)3 * w * h * sizeof(float)));
        }
    }
    for (int g = 0; g < n_workers; g++) {
        lumacu_host_free(slot[g].in);
        lumacu_host_free(slot[g].out);
        for (int pl = 0; pl < 3; pl++)
            lumacu_host_free(slot[g].planes[pl]);
        lumacu_destroy(ctx[g]);
    }
    return 0;
}
#endif

int main(int argc, char **argv)
{
    const int n_workers = argc > 1 ? atoi(argv[1]) : 1, n_frames = argc > 2 ? atoi(argv[2]) : 8;
    const unsigned w = argc > 4 ? atoi(argv[3]) : 1280, h = argc > 4 ? atoi(argv[4]) : 720;
    const std::string mode = argc > 5 ? argv[5] : "threads";
    if (n_workers < 1 || n_frames < 0 || (w & 1) || (h & 1)) {
        printf("usage: multi_gpu <n_workers> <n_frames> [w h [threads|async]]\n");
        return 2;
    }
    std::vector<std::string> lines(n_frames);
    int rc;
#ifndef LUMA_REFERENCE_BUILD
    if (mode == "async")
        rc = run_async(n_workers, n_frames, w, h, lines);
    else
#endif
        rc = run_threads(n_workers, n_frames, w, h, lines);
    if (rc)
        return rc;
    for (int f = 0; f < n_frames; f++)
        printf("%s\n", lines[f].c_str());
    return 0;
}

/*
 * multi_gpu.cpp -- TEST DRIVER: one LumaEncoder + LumaDecoder pair per GPU, each on its own host thread, frame f of
 * the stream handled by GPU f mod N (frames are independent; nothing but the quantizer parameters is
 * shared).  Prints one hash line per frame, in frame order; the output must not depend on N.
 *
 *   usage: multi_gpu <n_gpus> <n_frames> [w h]
 */
#include <luma_decoder.h>
#include <luma_encoder.h>
#include <luma_exception.h>

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

int main(int argc, char **argv)
{
    const int n_gpus = argc > 1 ? atoi(argv[1]) : 1, n_frames = argc > 2 ? atoi(argv[2]) : 8;
    const unsigned w = argc > 4 ? atoi(argv[3]) : 1280, h = argc > 4 ? atoi(argv[4]) : 720;
    std::vector<std::string> lines(n_frames);
    std::vector<std::string> errors(n_gpus);
    std::vector<std::thread> workers;
    for (int g = 0; g < n_gpus; g++)
        workers.emplace_back([&, g]() {
            try {
                const std::string file = "gpu" + std::to_string(g) + ".mkv"; /* in-memory container double */
                LumaEncoder enc;
                enc.setDevice(g);
                LumaEncoderParams p = enc.getParams();
                p.lossLess = 1;
                enc.setParams(p);
                std::vector<int> mine;
                for (int f = g; f < n_frames; f += n_gpus)
                    mine.push_back(f);
                for (size_t i = 0; i < mine.size(); i++) {
                    LumaFrame frame;
                    fill(frame, w, h, mine[i]);
                    if (!enc.initialized())
                        enc.initialize(file.c_str(), w, h);
                    enc.encode(&frame);
                }
                enc.finish();
                if (mine.empty())
                    return;
                LumaDecoder dec;
                dec.setDevice(g);
                dec.initialize(file.c_str());
                for (size_t i = 0; i < mine.size(); i++) {
                    LumaFrame *out = dec.decode();
                    if (!out)
                        throw LumaException("decoder ran out of frames");
                    LumaDecoderParams dp = dec.getParams();
                    uint32_t ph = 2166136261u;
                    for (int pl = 0; pl < 3; pl++)
                        for (int y = 0; y < dp.height[pl]; y++)
                            ph = fnv(dec.getBuffer()[pl] + (size_t)y * dp.stride[pl], (size_t)dp.width[pl] * 2, ph);
                    char buf[128];
                    snprintf(buf, sizeof(buf), "frame %d planes %08x floats %08x", mine[i], ph,
                             fnv(out->buffer, (size_t)3 * w * h * sizeof(float)));
                    lines[mine[i]] = buf;
                }
            } catch (LumaException &e) {
                errors[g] = e.what();
            }
        });
    for (size_t i = 0; i < workers.size(); i++)
        workers[i].join();
    for (int g = 0; g < n_gpus; g++)
        if (!errors[g].empty()) {
            printf("GPU %d: LumaException: %s\n", g, errors[g].c_str());
            return 1;
        }
    for (int f = 0; f < n_frames; f++)
        printf("%s\n", lines[f].c_str());
    return 0;
}

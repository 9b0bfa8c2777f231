/* exr_interface.cpp -- see exr_interface.h (OpenEXR-free stand-in, raw float container). */
#include "exr_interface.h"

#include "luma_exception.h"

#include <cstring>

/* The reference's test pattern (src/exr_interface.cpp:58-66), restated: the top fifth of the frame is grey
 * -- a quadratic ramp 1e4*(x/w)^2 in its upper half, a 20-step staircase below it -- and the rest is a
 * 20x30 checkerboard in R whose "on" rows carry vertical / horizontal quadratic ramps in G / B.
 * Integer divisions and the float expression order follow the reference so the values are identical. */
bool ExrInterface::testFrame(LumaFrame &frame, unsigned int w, unsigned int h)
{
    frame.width = w;
    frame.height = h;
    frame.channels = 3;
    if (!frame.init())
        throw LumaException("Cannot allocate memory for input frame");
    float *R = frame.getChannel(0), *G = frame.getChannel(1), *B = frame.getChannel(2);
    const float peak = 10000.0f;
    for (size_t y = 0; y < h; y++) {
        const size_t band = 20 * y / h;        /* 20 horizontal bands */
        const bool on = (band % 2) != 0;
        for (size_t x = 0; x < w; x++) {
            const size_t i = x + y * w;
            const float xx = peak * ((float)(x * x)) / (w * w); /* ((1e4 * x^2) / w^2), size_t denominator */
            if (y < h / 10) {
                R[i] = G[i] = B[i] = xx;
            } else if (y < h / 5) {
                R[i] = G[i] = B[i] = peak * ((20 * x) / w) / 20.0f;
            } else {
                const size_t col = 30 * x / w; /* 30 vertical bands */
                R[i] = peak * (float)((band % 2) ^ (col % 2));
                G[i] = peak * (on ? 1 : 0) * ((float)(y * y)) / (h * h);
                B[i] = peak * (on ? 1 : 0) * ((float)(x * x)) / (w * w);
            }
        }
    }
    return true;
}

bool ExrInterface::writeFrame(const char *outputFile, LumaFrame &frame)
{
    if (!outputFile || !frame.buffer)
        return false;
    FILE *f = fopen(outputFile, "wb");
    if (!f)
        return false;
    fprintf(f, "LUMAF32 %u %u %u\n", frame.width, frame.height, frame.channels);
    const size_t n = (size_t)frame.width * frame.height * frame.channels;
    const bool ok = fwrite(frame.buffer, sizeof(float), n, f) == n;
    fclose(f);
    return ok;
}

bool ExrInterface::readFrame(const char *inputFile, LumaFrame &frame)
{
    FILE *f = inputFile ? fopen(inputFile, "rb") : NULL;
    if (!f)
        return false;
    unsigned int w = 0, h = 0, c = 0;
    char nl = 0;
    bool ok = fscanf(f, "LUMAF32 %u %u %u%c", &w, &h, &c, &nl) == 4 && nl == '\n' && w && h && c == 3;
    if (ok) {
        frame.width = w;
        frame.height = h;
        frame.channels = c;
        ok = frame.init();
        const size_t n = (size_t)w * h * c;
        ok = ok && fread(frame.buffer, sizeof(float), n, f) == n;
    }
    fclose(f);
    return ok;
}

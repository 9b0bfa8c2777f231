/*
 * luma_frame.h -- planar float32 frame of the drop-in C++ facade.
 *
 * Public layout and methods are those of the reference's LumaFrame (reference
 * include/luma/luma_frame.h:51-90): fields height, width, channels, buffer;
 * buffer[c*height*width + y*width + x]; init() / clear() / getChannel().
 *
 * Difference (private to the implementation): the pixel storage comes from a
 * small cache of PAGE-LOCKED host blocks (lumacu_frame_alloc) instead of
 * new[], because LumaEncoder::encode / LumaDecoder::decode copy the frame over
 * PCIe and the reference's drivers construct a fresh LumaFrame per video frame
 * (lumaenc.cpp:211, test/test_simple_enc.cpp:53).  When no CUDA driver is
 * present the allocator degrades to malloc, so code that only moves frames
 * around (readers, writers) keeps working.
 */
#ifndef LUMAFRAME_H
#define LUMAFRAME_H

#include <cstddef>

extern "C" {
/* implemented in libluma_b200 (src/luma_frame_pool.cpp) */
float *lumacu_frame_alloc(size_t n_floats);
void lumacu_frame_free(float *p);
}

struct LumaFrame
{
    LumaFrame(unsigned int w = 0, unsigned int h = 0, unsigned int c = 3)
        : height(h), width(w), channels(c), buffer(NULL)
    {
        if (height && width && channels)
            init();
    }

    ~LumaFrame() { clear(); }

    void clear()
    {
        if (buffer != NULL) {
            lumacu_frame_free(buffer);
            buffer = NULL;
        }
    }

    bool init()
    {
        if (!height && !width && !channels)
            return false;
        clear();
        buffer = lumacu_frame_alloc((size_t)channels * height * width);
        return buffer != NULL;
    }

    float *getChannel(unsigned int c) { return buffer + (size_t)c * height * width; }

    unsigned int height, width, channels;
    float *buffer;

private:
    LumaFrame(const LumaFrame &);            /* the reference's frame is not safely copyable either */
    LumaFrame &operator=(const LumaFrame &); /* (implicit copy would double-free buffer) */
};

#endif // LUMAFRAME_H

/*
 * exr_interface.h -- OpenEXR-free stand-in for the reference's ExrInterface
 * (reference include/exr_interface.h:55-61), so that the reference's drivers
 * (lumaenc.cpp, lumadec.cpp, test/test_simple_enc.cpp, test/test_simple_dec.cpp)
 * build on machines without OpenEXR -- such as the B200 image.  Same class
 * name, same three static methods, same argument meaning.
 *
 *   testFrame   the reference's synthetic HDR pattern (src/exr_interface.cpp:50-70)
 *   readFrame / writeFrame
 *               a trivial raw container instead of OpenEXR: the text header
 *               "LUMAF32 <width> <height> <channels>\n" followed by the planar
 *               little-endian float32 samples of LumaFrame::buffer.
 *
 * Where OpenEXR exists, build the drivers against the reference's own
 * src/exr_interface.cpp instead; nothing in the encoder/decoder depends on
 * which one is linked.
 */
#ifndef EXR_INTERFACE_H
#define EXR_INTERFACE_H

#include <cstddef>
#include <stdio.h>

#include "luma_frame.h"

class ExrInterface
{
public:
    static bool readFrame(const char *inputFile, LumaFrame &frame);
    static bool writeFrame(const char *outputFile, LumaFrame &frame);
    static bool testFrame(LumaFrame &frame, unsigned int w = 1280, unsigned int h = 720);
};

#endif // EXR_INTERFACE_H

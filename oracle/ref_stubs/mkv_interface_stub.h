/*
 * mkv_interface_stub.h -- TEST DOUBLE, part of the oracle (test infrastructure).
 *
 * In-memory stand-in for the reference's MkvInterface (include/luma/
 * mkv_interface.h:76-157): same public method names and argument meaning, but
 * "files" live in a process-global map keyed by file name, so the reference's
 * LumaEncoder / LumaDecoder compile and run unmodified without libebml /
 * libmatroska.  Force-included (-include) ahead of the reference headers; it
 * claims the MKV_INTERFACE_H guard so the real header is skipped.
 */
#ifndef MKV_INTERFACE_H
#define MKV_INTERFACE_H

#include <cstddef>
#include <cstdio>
#include <cstring>
#include <stdint.h>
#include <string>
#include <vector>

typedef unsigned char binary;
typedef uint8_t uint8;

struct MkvStubFile;

class MkvInterface
{
public:
    MkvInterface();
    ~MkvInterface();

    void openWrite(const char *outputFile, const unsigned int w, const unsigned int h,
                   const float maxL, const float minL);
    void openRead(const char *inputFile);
    void close();
    void addAttachment(unsigned int uid, const binary *buffer, unsigned int buffer_size,
                       const char *description = "--");
    void addFrame(const uint8 *frame_buffer, unsigned int buffer_size, bool isKey = true);
    bool readFrame();
    bool getAttachment(unsigned int ind, binary **buffer, unsigned int &id,
                       unsigned int &buffer_size);
    void writeAttachments();
    const uint8 *getFrame(unsigned int &buffer_size);
    bool seekToTime(float tm, bool absolute = false);

    void setFramerate(float fps) { m_frameDuration = 1000.0f / fps; }
    void setVerbose(bool verbose) { m_verbose = verbose; }
    int getCurrentTime() { return 0; }
    int getFrameDuration() { return (int)m_frameDuration; }
    int getDuration() { return 0; }

private:
    MkvStubFile *m_file;
    size_t m_readPos;
    bool m_verbose;
    float m_frameDuration;
};

/* harness-side access to the in-memory container */
MkvStubFile *mkv_stub_find(const char *name);
void mkv_stub_erase(const char *name);
size_t mkv_stub_frame_count(const char *name);
const std::vector<uint8> *mkv_stub_frame(const char *name, size_t idx);
void mkv_stub_append_frame(const char *name, const uint8 *data, size_t n);

#endif // MKV_INTERFACE_H

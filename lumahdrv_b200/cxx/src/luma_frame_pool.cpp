/*
 * luma_frame_pool.cpp -- storage behind LumaFrame (include/luma/luma_frame.h).
 *
 * The reference allocates every frame with new[] (reference
 * include/luma/luma_frame.h:62-81) and its drivers build a fresh LumaFrame per
 * video frame.  Here the storage is page-locked (lumacu_host_alloc) so the
 * PCIe copies inside LumaEncoder::encode / LumaDecoder::decode run at full
 * rate, and freed blocks are kept in a small cache because pinning ~100 MB per
 * frame would cost more than the transform itself.
 */
#include "../../../include/lumacu.h"

#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

namespace {

struct Block {
    size_t bytes;
    bool pinned;
};

struct Pool {
    std::mutex mu;
    std::map<void *, Block> live;                 /* handed out */
    std::multimap<size_t, std::pair<void *, bool> > idle; /* bytes -> (ptr, pinned) */
    size_t idle_bytes = 0;
    static constexpr size_t kMaxIdleBytes = (size_t)3 << 30; /* keep at most 3 GiB parked */

    ~Pool()
    {
        /* process exit: the CUDA runtime may already be gone; leave pinned blocks to the OS */
        for (auto &kv : idle)
            if (!kv.second.second)
                free(kv.second.first);
    }
};

Pool &pool()
{
    static Pool *p = new Pool(); /* intentionally leaked: frames may be destroyed during static teardown */
    return *p;
}

bool pinning_disabled()
{
    const char *e = getenv("LUMA_PINNED_FRAMES");
    return e && e[0] == '0';
}

} // namespace

extern "C" float *lumacu_frame_alloc(size_t n_floats)
{
    const size_t bytes = n_floats * sizeof(float);
    if (!bytes)
        return NULL;
    Pool &P = pool();
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.idle.find(bytes);
        if (it != P.idle.end()) {
            void *p = it->second.first;
            P.live[p] = Block{bytes, it->second.second};
            P.idle_bytes -= bytes;
            P.idle.erase(it);
            return (float *)p;
        }
    }
    void *p = NULL;
    bool pinned = false;
    if (!pinning_disabled() && lumacu_host_alloc(bytes, &p) == LUMACU_OK && p)
        pinned = true;
    else
        p = malloc(bytes);
    if (!p)
        return NULL;
    std::lock_guard<std::mutex> lk(P.mu);
    P.live[p] = Block{bytes, pinned};
    return (float *)p;
}

extern "C" void lumacu_frame_free(float *ptr)
{
    if (!ptr)
        return;
    Pool &P = pool();
    Block b;
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.live.find(ptr);
        if (it == P.live.end())
            return; /* not ours (e.g. a caller installed its own buffer); the caller frees it */
        b = it->second;
        P.live.erase(it);
        if (P.idle_bytes + b.bytes <= Pool::kMaxIdleBytes) {
            P.idle.insert(std::make_pair(b.bytes, std::make_pair((void *)ptr, b.pinned)));
            P.idle_bytes += b.bytes;
            return;
        }
    }
    if (b.pinned)
        lumacu_host_free(ptr);
    else
        free(ptr);
}

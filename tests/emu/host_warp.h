// tests/emu/host_warp.h -- TEST-ONLY: the warp of eik_fast.cuh as EIKF_HOST_LANES host threads.  Every lane is a thread,
// the lane-interleaved arrays have stride EIKF_HOST_LANES, and the warp collectives (__any_sync, __reduce_min/max_sync,
// __syncwarp) are a sense-reversing barrier over one slot per lane -- so the code that depends on what the OTHER lanes do
// (ranges that are unions over the lanes, lock-step sweeps, hand-over in split mode) runs on the CPU as it does on the GPU.
#pragma once
#include <atomic>
#include <thread>
#ifndef EIKF_HOST_LANES
#define EIKF_HOST_LANES 8
#endif
constexpr int LS = EIKF_HOST_LANES;
namespace host_warp {
struct Team {
    std::atomic<int> arrived{0};
    std::atomic<int> phase{0};
    int slot[EIKF_HOST_LANES];
    int site[EIKF_HOST_LANES];      // source line of the collective each lane is at
    std::atomic<int> mismatches{0}; // collectives that the lanes reached from different call sites (undefined on the GPU)
};
inline Team& team() { static Team t; return t; }
inline int& lane() { static thread_local int l = 0; return l; }
inline void barrier()
{
    Team& t = team();
    const int ph = t.phase.load(std::memory_order_acquire);
    if (t.arrived.fetch_add(1, std::memory_order_acq_rel) == EIKF_HOST_LANES - 1) {
        t.arrived.store(0, std::memory_order_relaxed);
        t.phase.store(ph + 1, std::memory_order_release);
    } else {
        int spins = 0;
        while (t.phase.load(std::memory_order_acquire) == ph)
            if (++spins > 2000) { std::this_thread::yield(); spins = 0; }
    }
}
// op: 0 = any, 1 = min, 2 = max
inline int reduce(int v, int op, int site)
{
    Team& t = team();
    t.slot[lane()] = v;
    t.site[lane()] = site;
    barrier();
    int r = t.slot[0];
    for (int i = 1; i < EIKF_HOST_LANES; i++) {
        const int x = t.slot[i];
        r = (op == 0) ? (r | x) : (op == 1) ? (x < r ? x : r) : (x > r ? x : r);
        if (lane() == 0 && t.site[i] != t.site[0]) t.mismatches.fetch_add(1);
    }
    barrier();
    return r;
}
inline void sync(int site) { reduce(0, 0, site); }
}   // namespace host_warp
// every lane of the warp must be at the SAME collective (on the GPU anything else is undefined): the call site is checked
#define EIKF_ANY(p) (host_warp::reduce((p) ? 1 : 0, 0, __LINE__) != 0)
#define EIKF_SYNC() host_warp::sync(__LINE__)
#define EIKF_MIN(v) host_warp::reduce((v), 1, __LINE__)
#define EIKF_MAX(v) host_warp::reduce((v), 2, __LINE__)

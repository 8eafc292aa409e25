// TEST INFRASTRUCTURE ONLY.  Just enough of the CUDA device vocabulary to compile the kernel headers
// of nanospring_b200/csrc (*_kernels.cuh) for the HOST and run them with real threads: every lane of
// a warp is an OS thread, warp collectives (__shfl_*_sync, __ballot_sync, __any/__all_sync,
// __reduce_add_sync, __syncwarp) meet on a barrier, atomics are host atomics.
//   emu_launch        one warp at a time: for warp-level kernels ("shared memory" is a buffer the harness
//                     passes in; __syncthreads is only meaningful with one warp per block)
//   emu_launch_block  all warps of a block alive at once: __syncthreads is a block barrier, `__shared__`
//                     variables are function-local statics (blocks run one after the other)
// All state is `inline` (one instance per program or shared library), so several harnesses can be
// linked into one executable (tests/cpp/tsan_driver.cpp, the ThreadSanitizer run) as well as loaded
// as separate libraries by the Python tests.
#pragma once
#include <stdint.h>

#include <barrier>
#include <functional>
#include <thread>
#include <vector>

#define __global__ inline
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x) alignas(x)

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local emu_dim3 blockIdx, threadIdx;
inline emu_dim3 gridDim, blockDim;

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

template <typename T> inline T __ldg(const T *p) { return *p; }
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
inline uint32_t __vcmpeq4(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i)
        if (((a >> (8 * i)) & 0xFF) == ((b >> (8 * i)) & 0xFF)) r |= 0xFFu << (8 * i);
    return r;
}
inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline uint32_t __funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t sh) {   // shift clamped to 32
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh > 32 ? 32 : sh));
}
inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
    const uint64_t src = ((uint64_t)b << 32) | a;              // selector nibbles 0..7 only (no sign replication)
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((src >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31));
}
inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) {
    return (uint32_t)((((((uint64_t)hi) << 32) | lo) << (sh & 31)) >> 32);
}
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}

struct EmuWarp {
    std::barrier<> bar{32};
    uint32_t slot[32];
    uint64_t slot64[32];
};
inline thread_local EmuWarp *emu_warp = nullptr;
inline thread_local int emu_lane = 0;

inline uint32_t __shfl_up_sync(uint32_t, uint32_t v, int d) {
    EmuWarp *w = emu_warp;
    w->slot[emu_lane] = v;
    w->bar.arrive_and_wait();
    const uint32_t r = emu_lane >= d ? w->slot[emu_lane - d] : v;
    w->bar.arrive_and_wait();
    return r;
}
inline uint64_t emu_exchange64(uint64_t v, int src) {
    EmuWarp *w = emu_warp;
    w->slot64[emu_lane] = v;
    w->bar.arrive_and_wait();
    const uint64_t r = w->slot64[src & 31];
    w->bar.arrive_and_wait();
    return r;
}
inline uint64_t __shfl_sync(uint32_t, uint64_t v, int src) { return emu_exchange64(v, src); }
inline unsigned long long __shfl_sync(uint32_t, unsigned long long v, int src) { return emu_exchange64(v, src); }
inline uint32_t __shfl_sync(uint32_t, uint32_t v, int src) { return (uint32_t)emu_exchange64(v, src); }
inline unsigned long long __shfl_xor_sync(uint32_t, unsigned long long v, int m) { return emu_exchange64(v, emu_lane ^ m); }
inline uint32_t __shfl_xor_sync(uint32_t, uint32_t v, int m) { return (uint32_t)emu_exchange64(v, emu_lane ^ m); }
inline uint64_t __shfl_xor_sync(uint32_t, uint64_t v, int m) { return emu_exchange64(v, emu_lane ^ m); }
inline void __syncwarp() {
    emu_warp->bar.arrive_and_wait();
}
// emu_launch: one warp runs at a time, so this is only meaningful with ONE warp per block
// (emu_launch(grid, 32, ...)); emu_launch_block runs all warps of a block concurrently and meets here
inline thread_local std::barrier<> *emu_block_bar = nullptr;
inline void __syncthreads() {
    if (emu_block_bar) emu_block_bar->arrive_and_wait();
    else __syncwarp();
}
inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline uint32_t atomicCAS(uint32_t *p, uint32_t expected, uint32_t desired) {
    __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return expected;                                           // the value found, like the device intrinsic
}
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { }
    return old;
}
inline uint64_t min(uint64_t a, uint64_t b) { return a < b ? a : b; }
inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
inline uint32_t __ballot_sync(uint32_t, bool pred) {
    EmuWarp *w = emu_warp;
    w->slot[emu_lane] = pred ? 1u : 0u;
    w->bar.arrive_and_wait();
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r |= w->slot[i] << i;
    w->bar.arrive_and_wait();
    return r;
}
inline bool __any_sync(uint32_t m, bool pred) { return __ballot_sync(m, pred) != 0; }
inline bool __all_sync(uint32_t m, bool pred) { return __ballot_sync(m, !pred) == 0; }
inline uint32_t __reduce_add_sync(uint32_t, uint32_t v) {
    EmuWarp *w = emu_warp;
    w->slot[emu_lane] = v;
    w->bar.arrive_and_wait();
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r += w->slot[i];
    w->bar.arrive_and_wait();
    return r;
}

// kernel<<<grid, block>>>(...) : body() is the kernel call with its arguments bound
inline void emu_launch(unsigned grid, unsigned block, const std::function<void()> &body) {
    gridDim.x = grid;
    blockDim.x = block;
    for (unsigned b = 0; b < grid; ++b)
        for (unsigned w0 = 0; w0 < block; w0 += 32) {
            EmuWarp warp;
            std::vector<std::thread> lanes;
            for (int l = 0; l < 32; ++l)
                lanes.emplace_back([&, l] {
                    emu_warp = &warp;
                    emu_lane = l;
                    blockIdx.x = b;
                    threadIdx.x = w0 + l;
                    body();
                });
            for (auto &t : lanes) t.join();
        }
}

// The same with all warps of a block alive at once (block-level cooperation: __syncthreads, shared
// counters).  `__shared__` variables are function-local statics: blocks run one after the other, so
// one instance serves them all.  block must be a multiple of 32.
#define __shared__ static
inline void emu_launch_block(unsigned grid, unsigned block, const std::function<void()> &body) {
    gridDim.x = grid;
    blockDim.x = block;
    const unsigned nwarps = block / 32;
    for (unsigned b = 0; b < grid; ++b) {
        std::vector<EmuWarp> warps(nwarps);
        std::barrier<> block_bar((std::ptrdiff_t)block);
        std::vector<std::thread> threads;
        for (unsigned t = 0; t < block; ++t)
            threads.emplace_back([&, t] {
                emu_warp = &warps[t / 32];
                emu_lane = (int)(t % 32);
                emu_block_bar = &block_bar;
                blockIdx.x = b;
                threadIdx.x = t;
                body();
                emu_block_bar = nullptr;
            });
        for (auto &th : threads) th.join();
    }
}

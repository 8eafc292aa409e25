// Device code of the sketch kernels (see sketch.cu for the reference lines they replace and for the
// shortcut the filter kernel takes).  Kept free of runtime-API includes so that
// tests/cpp/sketch_host_emul.cpp can compile the SAME kernels for the host (NSMH_HOST_EMUL: lock-step
// warp emulation, one warp per block, the bulk copy replaced by a memcpy) and check them against the
// oracle without a GPU.
#pragma once
#include <stdint.h>
#ifdef NSMH_HOST_EMUL
#include <string.h>
#endif

#include "nsmh_constants.h"
#include "sketch_device.cuh"

namespace nsmh {

struct SketchArgs {
    const uint64_t *off;        // [n_reads+1] global base offsets
    const uint32_t *W;          // packed stream
    uint64_t *sk;               // [n_reads][n]
    const uint32_t *tile_start; // [n_reads+1] exclusive scan of tiles per read
    const uint32_t *tile_read;  // [num_tiles] read of every tile (nullptr: binary search)
    const uint64_t *rnd;        // [n]
    const uint16_t *ftab_first;         // per prefix: first hash that targets it | second one << 8 (0xFF = none)
    const uint8_t *ftab_next, *ftab_hit3;
    unsigned long long *counters;   // [0] fix-ups
    unsigned int *tile_queue;       // next tile to hand out (filter kernel: warps take tiles dynamically)
    uint32_t n_reads, k, n;
    int lambda_log2;
    uint32_t tile_words;        // words (16 k-mer starts each) per tile
};

// ---- row init + tile counts ---------------------------------------------------
__global__ void __launch_bounds__(256)
sketch_init_kernel(SketchArgs a, uint32_t *__restrict__ tile_cnt) {
    // one warp per read: rows are written with coalesced stores, no per-element division
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < a.n_reads; i += warps) {
        const uint64_t b0 = a.off[i], len = a.off[i + 1] - b0;
        // ReadFilter.cpp:119-124: untouched (zero) when len-k+1 < 0, all-ones otherwise
        const uint64_t v = (len + 1 < a.k) ? 0ULL : ~0ULL;
        uint64_t *row = a.sk + (size_t)i * a.n;
        for (uint32_t l = lane; l < a.n; l += 32) row[l] = v;
        if (lane == 0) {
            uint32_t tiles = 0;
            if (len >= a.k) {
                uint64_t nk = len - a.k + 1;
                // tiles start on a 4-word (16-byte) boundary: the bulk copies need it
                uint64_t w0 = (b0 / kWordBases) & ~3ULL, w1 = (b0 + nk - 1) / kWordBases;
                tiles = (uint32_t)((w1 - w0 + a.tile_words) / a.tile_words);
            }
            tile_cnt[i] = tiles;
        }
    }
}

// tile -> read map: one thread per read writes its (few) tiles
__global__ void __launch_bounds__(256)
sketch_tile_map_kernel(const uint32_t *__restrict__ ts, uint32_t n_reads, uint32_t *__restrict__ tile_read) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_reads; i += gridDim.x * blockDim.x)
        for (uint32_t t = ts[i]; t < ts[i + 1]; ++t) tile_read[t] = i;
}

__device__ __forceinline__ uint32_t find_read_of_tile(const uint32_t *__restrict__ ts, uint32_t n_reads,
                                                      uint32_t tile) {
    uint32_t lo = 0, hi = n_reads;   // largest i with ts[i] <= tile (reads without tiles skipped)
    while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (ts[mid] <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ TileGeom tile_geom(const SketchArgs &a, uint32_t tile) {
    TileGeom g;
    g.read = a.tile_read ? a.tile_read[tile] : find_read_of_tile(a.tile_start, a.n_reads, tile);
    g.rb = a.off[g.read];
    g.nk = a.off[g.read + 1] - g.rb - a.k + 1;
    uint64_t w0 = (g.rb / kWordBases) & ~3ULL, w1 = (g.rb + g.nk - 1) / kWordBases;
    g.w_begin = w0 + (uint64_t)(tile - a.tile_start[g.read]) * a.tile_words;
    g.w_end = g.w_begin + a.tile_words < w1 + 1 ? g.w_begin + a.tile_words : w1 + 1;
    return g;
}

__device__ __forceinline__ int filter_bits(uint64_t nk, int lambda_log2, int max_bits, uint32_t k) {
    int b = 63 - __clzll((long long)nk) - lambda_log2;
    b = b < 0 ? 0 : b;
    b = b > max_bits ? max_bits : b;
    b = b > 2 * (int)k ? 2 * (int)k : b;
    return b;
}

// ---- bulk copy + mbarrier (sm_90+/sm_100a PTX) -------------------------------------
#ifndef NSMH_HOST_EMUL
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}

#define NSMH_FILTER_SMEM_PARAM
#define NSMH_FILTER_SMEM_DECL extern __shared__ __align__(16) uint8_t smem[];
#else
// host emulation (tests/cpp/sketch_host_emul.cpp): the bulk copy is a memcpy by the issuing lane, which
// the warp then meets at a barrier; `smem` is a buffer the harness passes in
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t) { *bar = 0; }
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *, uint32_t) {}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *) { memcpy(dst, src, bytes); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t *, uint32_t) {
    __syncwarp();
    return true;
}
#define NSMH_FILTER_SMEM_PARAM , uint8_t *smem
#define NSMH_FILTER_SMEM_DECL
#endif

// shared memory of the filter kernel: block-wide tables, then one private area per warp
struct FilterSmem {
    uint32_t stage_words;    // per warp: staged packed words (tile_words + 8)
    uint32_t mask_words;     // per warp: hit masks (tile_words / 2)
    size_t tab_bytes, warp_bytes;
    __host__ __device__ FilterSmem(uint32_t n, uint32_t tile_words) {
        stage_words = tile_words + 8;
        mask_words = tile_words / 2;
        size_t t = kFilter3TabSize + (4 << kFilter3MaxBits) + (size_t)(kFilterMaxBits + 1) * n + (size_t)n * 8 + 8;
        tab_bytes = (t + 15) & ~(size_t)15;
        warp_bytes = ((size_t)stage_words * 4 + (size_t)mask_words * 4 + (size_t)n * 16 + 16 + 15) & ~(size_t)15;
    }
};

// ---- filter kernel ------------------------------------------------------------------
// a.tile_words is a multiple of 64: in every step a lane owns two adjacent words = 32 positions.

// Plain 64-bit accesses to the warp's minima that several lanes may make to the same word in the same
// instruction (phase 2 below: one of the colliding stores wins, every writer re-reads).  volatile on the
// device so that they are neither cached in registers nor moved across the warp barriers; relaxed
// atomics in the host emulation (where every lane is an OS thread).
#ifndef NSMH_HOST_EMUL
__device__ __forceinline__ void racy_store64(uint64_t *p, uint64_t v) { *reinterpret_cast<volatile uint64_t *>(p) = v; }
__device__ __forceinline__ uint64_t racy_load64(const uint64_t *p) { return *reinterpret_cast<const volatile uint64_t *>(p); }
#else
__device__ __forceinline__ void racy_store64(uint64_t *p, uint64_t v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
__device__ __forceinline__ uint64_t racy_load64(const uint64_t *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
#endif

// hit mask of the 32 positions that start in words i0, i0+1 of the staged tile (bit 31-q = position q)
__device__ __forceinline__ uint32_t filter_hits32(const uint32_t *sw, const uint8_t *s_hit3, uint32_t i0, int rshift3,
                                                  uint32_t lo_pos, uint32_t hi_pos) {
    const uint2 w01 = *reinterpret_cast<const uint2 *>(sw + i0);
    const uint32_t w0 = w01.x, w1 = w01.y, w2 = sw[i0 + 2];
    uint32_t hi = 0, lo = 0;
#pragma unroll
    for (int t = 0; t < 5; ++t) {          // positions 0..14
        const uint32_t v = t ? __funnelshift_l(w1, w0, 6 * t) : w0;
        hi = hi * 8 + s_hit3[__funnelshift_rc(v, 1u, rshift3)];   // (1<<(b+4)) | window
    }
#pragma unroll
    for (int t = 5; t < 11; ++t) {         // positions 15..32
        const uint32_t v = t == 5 ? __funnelshift_l(w1, w0, 30) : __funnelshift_l(w2, w1, 6 * t - 32);
        lo = lo * 8 + s_hit3[__funnelshift_rc(v, 1u, rshift3)];
    }
    uint32_t m = (hi << 17) | (lo >> 1);
    const uint32_t pb = i0 * kWordBases;
    if (pb < lo_pos || pb + 32 > hi_pos) {     // read / tile edges
        const uint32_t first = lo_pos > pb ? lo_pos - pb : 0u;          // valid q in [first, last)
        const uint32_t last = hi_pos > pb ? min(hi_pos - pb, 32u) : 0u;
        const uint32_t keep_hi = first >= 32 ? 0u : 0xFFFFFFFFu >> first;
        const uint32_t keep_lo = last == 0 ? 0u : 0xFFFFFFFFu << (32 - last);
        m &= keep_hi & keep_lo;
    }
    return m;
}

// One block per SM, a warp per tile.  Per tile:
//   stage    the tile's packed words arrive in shared memory through ONE bulk copy (cp.async.bulk + mbarrier)
//   phase 1  32 positions per lane and step: funnel shifts + table lookups (one lookup answers three
//            consecutive positions) give a hit mask; the hits of the whole warp are appended to ONE list
//            of tile positions (prefix sum of the lanes' counts, then every lane writes its own)
//   phase 2  the list is consumed 32 hits at a time, one per lane, whatever lane found them: k-mer from
//            the staged words, hash(es) that target its prefix, 64-bit XOR / compare against the warp's
//            minima.  A smaller value is written with a PLAIN store: when lanes collide on a hash one
//            store wins, all writers re-read after a warp barrier and the losers that are still smaller
//            try again (shared-memory atomics cost 2 cycles per lane, 64-bit min is a compare-and-swap
//            loop on top; the ncu capture of the first version of this kernel showed the shared-memory
//            pipe 87 % busy, a third of it in phase 2's atomics and scattered mask reads)
//   flush    one 64-bit atomicMin per (tile, hash) combines the tiles of a read
// The list holds tile_words positions (16 bit each); a tile with more hits (low-complexity sequence)
// is consumed in several rounds, and a single step with more hits than the list holds (> tile_words of
// its 1024 positions) is walked lane by lane with atomics instead.
__global__ void __launch_bounds__(1024)
sketch_filter_kernel(SketchArgs a NSMH_FILTER_SMEM_PARAM) {
    NSMH_FILTER_SMEM_DECL
    const FilterSmem L(a.n, a.tile_words);
    uint8_t *s_hit3 = smem;                                   // kFilter3TabSize
    uint16_t *s_first = reinterpret_cast<uint16_t *>(s_hit3 + kFilter3TabSize);   // 2^(kFilter3MaxBits+1) entries
    uint8_t *s_next = s_hit3 + kFilter3TabSize + (4 << kFilter3MaxBits);          // (kFilterMaxBits+1) * n
    uint64_t *s_rlo = reinterpret_cast<uint64_t *>(smem + ((kFilter3TabSize + (4 << kFilter3MaxBits) +
                                                            (size_t)(kFilterMaxBits + 1) * a.n + 7) & ~(size_t)7));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *mine = smem + L.tab_bytes + (size_t)warp * L.warp_bytes;
    uint32_t *sw = reinterpret_cast<uint32_t *>(mine);                      // staged words
    uint16_t *list = reinterpret_cast<uint16_t *>(sw + L.stage_words);      // hit positions (tile-relative)
    ulonglong2 *my_min = reinterpret_cast<ulonglong2 *>(sw + L.stage_words + L.mask_words);   // {rand[l] & mask, running minimum}
    uint64_t *bar = reinterpret_cast<uint64_t *>(my_min + a.n);
    const uint32_t cap = 2 * L.mask_words;                                  // list entries

    const uint64_t mask = kmer_mask(a.k);
    {
        const uint4 *g3 = reinterpret_cast<const uint4 *>(a.ftab_hit3);
        const uint4 *gf = reinterpret_cast<const uint4 *>(a.ftab_first);
        uint4 *s3 = reinterpret_cast<uint4 *>(s_hit3), *sf = reinterpret_cast<uint4 *>(s_first);
        for (int t = threadIdx.x; t < kFilter3TabSize / 16; t += blockDim.x) s3[t] = g3[t];
        for (int t = threadIdx.x; t < (4 << kFilter3MaxBits) / 16; t += blockDim.x) sf[t] = gf[t];
        for (uint32_t t = threadIdx.x; t < (kFilterMaxBits + 1) * a.n; t += blockDim.x) s_next[t] = a.ftab_next[t];
        for (uint32_t t = threadIdx.x; t < a.n; t += blockDim.x) s_rlo[t] = a.rnd[t] & mask;
        if (lane == 0) {
            mbar_init(bar, 1);
            fence_proxy_async();
        }
    }
    __syncthreads();

    const uint32_t num_tiles = a.tile_start[a.n_reads];
    const int kshift = 64 - 2 * (int)a.k;
    uint32_t phase = 0;

    // Tiles are handed out dynamically (their sizes differ by orders of magnitude, e.g. with
    // ultra-long reads).  The next index is requested while the current tile is processed,
    // so the atomic's latency is never waited for.
    uint32_t my_next = 0;
    if (lane == 0) my_next = atomicAdd(a.tile_queue, 1u);
    for (;;) {
        const uint32_t tile = __shfl_sync(0xffffffffu, my_next, 0);
        if (tile >= num_tiles) break;
        if (lane == 0) my_next = atomicAdd(a.tile_queue, 1u);
        const TileGeom g = tile_geom(a, tile);
        const uint32_t nw = (uint32_t)(g.w_end - g.w_begin);
        // ---- stage the tile: words [w_begin, w_end + 3) rounded up to 16 bytes, one bulk copy ----
        __syncwarp();                       // everybody is done with the previous tile's words
        if (lane == 0) {
            const uint32_t bytes = ((nw + 3 + 3) & ~3u) * 4;
            fence_proxy_async();
            mbar_expect_tx(bar, bytes);
            bulk_g2s(sw, a.W + g.w_begin, bytes, bar);
        }
        const int b = filter_bits(g.nk, a.lambda_log2, kFilter3MaxBits, a.k);
        const int rshift3 = 32 - (b + 4);       // window of b+4 bits answers positions j, j+1, j+2
        const int rshift = 32 - b;
        const uint8_t *nxt = s_next + (size_t)b * a.n;
        for (uint32_t l = lane; l < a.n; l += 32) my_min[l] = make_ulonglong2(s_rlo[l], ~0ULL);
        // valid k-mer starts of the tile, relative to its first base
        const uint64_t t0 = g.w_begin * kWordBases;
        const uint32_t lo_pos = g.rb > t0 ? (uint32_t)(g.rb - t0) : 0u;
        const uint32_t hi_pos = (uint32_t)min((uint64_t)nw * kWordBases, g.rb + g.nk - t0);
        const uint32_t steps = (nw + 63) / 64;
        while (!mbar_try_wait(bar, phase)) { }
        phase ^= 1;

        uint32_t total = 0;                 // hits in the list
        for (uint32_t it = 0;; ++it) {
            // ---- phase 1: this step's hit mask, the lanes' counts prefix-summed ----
            const bool done = it == steps;
            const uint32_t i0 = it * 64 + 2 * lane;
            uint32_t m = 0, incl = 0, wt = 0;
            if (!done) {
                m = filter_hits32(sw, s_hit3, i0, rshift3, lo_pos, hi_pos);
                incl = __popc(m);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                wt = __shfl_sync(0xffffffffu, incl, 31);
            }
            if (done || total + wt > cap) {
                // ---- phase 2: consume the list, one hit per lane and round ----
                __syncwarp();               // list entries (and, for the first round, the initial minima) are in place
                for (uint32_t base = 0; base < total; base += 32) {
                    const uint32_t i = base + lane;
                    bool want = false;
                    uint64_t x = 0, y = 0, *mp = nullptr;
                    uint32_t l2 = 0xFFu;
                    if (i < total) {
                        const uint32_t pos = list[i];
                        const uint32_t wi = pos >> 4;
                        uint32_t h32;
                        x = kmer_at(sw[wi], sw[wi + 1], sw[wi + 2], (int)(pos & 15), kshift, h32);
                        const uint32_t ll = s_first[__funnelshift_rc(h32, 1u, rshift)];   // a hit: at least one hash
                        const uint32_t l = ll & 0xFFu;
                        l2 = ll >> 8;
                        mp = reinterpret_cast<uint64_t *>(&my_min[l].y);
                        const ulonglong2 rm = my_min[l];
                        y = x ^ rm.x;
                        want = y < rm.y;
                        if (want) racy_store64(mp, y);
                    }
                    __syncwarp();           // all stores of the round are done: did mine survive, or a smaller one?
                    if (want && y < racy_load64(mp)) atomicMin(reinterpret_cast<unsigned long long *>(mp), (unsigned long long)y);
                    // further hashes with the same prefix (3 % of the hits): atomics, no plain store is in flight now
                    for (; l2 != 0xFFu; l2 = nxt[l2])
                        atomicMin(reinterpret_cast<unsigned long long *>(&my_min[l2].y), (unsigned long long)(x ^ my_min[l2].x));
                    __syncwarp();           // the next round's plain stores stay clear of this round's atomics
                }
                total = 0;
                if (done) break;
            }
            if (wt > cap) {
                // more hits in ONE step than the list holds: every lane walks its own mask
                while (m) {
                    const int q = __clz(m);
                    m &= ~(0x80000000u >> q);
                    const uint32_t wi = i0 + (q >> 4);
                    uint32_t h32;
                    const uint64_t x = kmer_at(sw[wi], sw[wi + 1], sw[wi + 2], q & 15, kshift, h32);
                    uint32_t l = s_first[__funnelshift_rc(h32, 1u, rshift)] & 0xFFu;
                    do {
                        const uint64_t y = x ^ my_min[l].x;
                        atomicMin(reinterpret_cast<unsigned long long *>(&my_min[l].y), (unsigned long long)y);
                        l = nxt[l];
                    } while (l != 0xFFu);
                }
                __syncwarp();
            } else {
                uint32_t slot = total + incl - __popc(m);
                const uint32_t pb = i0 * kWordBases;
                while (m) {
                    const int q = __clz(m);
                    m &= ~(0x80000000u >> q);
                    list[slot++] = (uint16_t)(pb + q);
                }
                total += wt;
            }
        }
        __syncwarp();
        for (uint32_t l = lane; l < a.n; l += 32) {
            const uint64_t v = my_min[l].y;
            if (v != ~0ULL)
                atomicMin(reinterpret_cast<unsigned long long *>(a.sk + (size_t)g.read * a.n + l),
                          (a.rnd[l] & ~mask) | v);
        }
    }
}

// ---- exact fix-up: (read, hash) pairs that no k-mer matched on the filter prefix ----
// Pass 1 streams the sketch matrix once and lists the entries that still hold the initial
// all-ones value; pass 2 recomputes each listed entry over all k-mers of its read, one warp
// per entry, 512 positions per warp step (coalesced word loads, 16 k-mers per lane).
__global__ void __launch_bounds__(256)
sketch_missing_kernel(SketchArgs a, uint32_t *__restrict__ list, unsigned int *__restrict__ count) {
    const int lane = threadIdx.x & 31;
    const uint64_t total = (uint64_t)a.n_reads * a.n;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (total + stride - 1) / stride;
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (uint64_t r = 0; r < rounds; ++r, t += stride) {     // whole warps stay in the loop for the ballot
        const bool miss = t < total && a.sk[t] == ~0ULL;
        const uint32_t m = __ballot_sync(0xffffffffu, miss);
        if (m) {
            const int leader = __ffs(m) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(count, (unsigned int)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (miss) list[base + __popc(m & ((1u << lane) - 1))] = (uint32_t)t;
        }
    }
}

// One warp per listed entry.  The scan works on the leading 32 bits of y = k-mer ^ rlo only
// (one funnel shift, one xor, one 32-bit min per position): y32 is a prefix of y, so the
// 64-bit minimum lies in a word whose 32-bit minimum equals the global one.  Every lane
// remembers the word of its own minimum; the winning words are then redone in 64 bits.  If a
// lane saw its minimum in two different words (a 32-bit tie, probability ~2^-32 per pair) the
// entry falls back to a plain 64-bit scan, so the result is exact in every case.
// U = words per lane and step (4 by default; NSMH_FIXUP_WIDTH=8 is an experiment: half as many
// dependent L2 round trips for the long reads that make up the kernel's tail).
// vals (optional): entry e's value goes to vals[e] instead of the sketch matrix - for the build that runs beside
// this kernel and must keep seeing all-ones for the listed entries (table_insert_list_kernel stores them later).
template <int U>
__global__ void __launch_bounds__(256)
sketch_fixup_kernel(SketchArgs a, const uint32_t *__restrict__ list, const unsigned int *__restrict__ count,
                    uint64_t *__restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    const uint64_t mask = kmer_mask(a.k);
    const int kshift = 64 - 2 * (int)a.k;
    const int s_lo = 2 * (int)a.k > 32 ? 2 * (int)a.k - 32 : 0;      // y32 = y >> s_lo
    const int sh = 2 * (int)a.k >= 32 ? 0 : 32 - 2 * (int)a.k;       // window >> sh = leading bits of the k-mer
    const uint32_t todo = *count;
    for (uint32_t e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < todo; e += warps) {
        const uint32_t t = list[e];
        const uint32_t i = t / a.n, lf = t - i * a.n;
        const uint64_t rb = a.off[i], len = a.off[i + 1] - rb;
        if (len < a.k) {                  // len == k-1: all-ones is the reference's value (ReadFilter.cpp:119-124)
            if (vals && lane == 0) vals[e] = ~0ULL;
            continue;
        }
        TileGeom g;
        g.read = i;
        g.rb = rb;
        g.nk = len - a.k + 1;
        g.w_begin = rb / kWordBases;
        g.w_end = (rb + g.nk - 1) / kWordBases + 1;
        const uint64_t r = a.rnd[lf], rlo = r & mask;
        const uint32_t t32 = (uint32_t)(rlo >> s_lo);
        uint32_t m = 0xFFFFFFFFu;
        uint64_t mw = 0;
        bool has = false, tie = false;
        // U words per lane and step, all loads first: the scan is a chain of L2 round trips
        for (uint64_t wb = g.w_begin + lane; wb < g.w_end; wb += 32 * U) {
            uint32_t w0[U], w1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint64_t w = wb + 32 * u;
                const bool in = w < g.w_end;
                w0[u] = in ? __ldg(a.W + w) : 0u;
                w1[u] = in ? __ldg(a.W + w + 1) : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint64_t w = wb + 32 * u;
                if (w >= g.w_end) break;
                int lo, hi;
                valid_range(g, w, lo, hi);
                if (lo >= hi) continue;
                uint32_t local = 0xFFFFFFFFu;
                if (lo == 0 && hi == kWordBases) {
#pragma unroll
                    for (int j = 0; j < kWordBases; ++j) {
                        const uint32_t x = j ? __funnelshift_l(w1[u], w0[u], 2 * j) : w0[u];
                        local = min(local, (x >> sh) ^ t32);
                    }
                } else {
                    for (int j = lo; j < hi; ++j) local = min(local, (__funnelshift_l(w1[u], w0[u], 2 * j) >> sh) ^ t32);
                }
                if (!has || local < m) { m = local; mw = w; tie = false; has = true; }
                else if (local == m) tie = true;
            }
        }
        uint32_t gm = has ? m : 0xFFFFFFFFu;
#pragma unroll
        for (int o = 16; o; o >>= 1) gm = min(gm, __shfl_xor_sync(0xffffffffu, gm, o));
        const bool cand = has && m == gm;
        uint64_t best = ~0ULL;
        if (__any_sync(0xffffffffu, cand && tie)) {
            for (uint64_t w = g.w_begin + lane; w < g.w_end; w += 32) {
                int lo, hi;
                valid_range(g, w, lo, hi);
                const uint64_t v = word_min64(a.W, w, lo, hi, kshift, rlo);
                best = v < best ? v : best;
            }
        } else if (cand) {
            int lo, hi;
            valid_range(g, mw, lo, hi);
            best = word_min64(a.W, mw, lo, hi, kshift, rlo);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other < best ? other : best;
        }
        if (lane == 0) {
            if (vals) vals[e] = (r & ~mask) | best;
            else a.sk[t] = (r & ~mask) | best;
            atomicAdd(a.counters, 1ULL);
        }
    }
}

// ---- brute force: the reference's operation count --------------------------------
// Warp per tile; hashes in register chunks of HC; every lane rolls the 16 k-mers of
// its word and keeps HC running minima; shuffles combine lanes; one 64-bit atomic
// min per (tile, hash) combines tiles of a read.
template <int HC>
__global__ void __launch_bounds__(256)
sketch_brute_kernel(SketchArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t num_tiles = a.tile_start[a.n_reads];
    const uint64_t mask = kmer_mask(a.k);
    const int kshift = 64 - 2 * (int)a.k;
    for (uint32_t tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < num_tiles;
         tile += warps) {
        const TileGeom g = tile_geom(a, tile);
        for (uint32_t c0 = 0; c0 < a.n; c0 += HC) {
            uint64_t r[HC], m[HC];
#pragma unroll
            for (int h = 0; h < HC; ++h) {
                r[h] = c0 + h < a.n ? (a.rnd[c0 + h] & mask) : 0ULL;
                m[h] = ~0ULL;
            }
            for (uint64_t wb = g.w_begin; wb < g.w_end; wb += 32) {
                const uint64_t w = wb + lane;
                if (w >= g.w_end) continue;
                int lo, hi;
                valid_range(g, w, lo, hi);
                if (lo >= hi) continue;
                const uint32_t w0 = __ldg(a.W + w), w1 = __ldg(a.W + w + 1), w2 = __ldg(a.W + w + 2);
#pragma unroll 4
                for (int j = 0; j < 16; ++j) {
                    // positions outside [lo,hi) re-evaluate a valid neighbour: min is idempotent
                    int jj = j < lo ? lo : (j >= hi ? hi - 1 : j);
                    uint32_t h32;
                    const uint64_t x = kmer_at(w0, w1, w2, jj, kshift, h32);
#pragma unroll
                    for (int h = 0; h < HC; ++h) {
                        uint64_t y = x ^ r[h];
                        m[h] = y < m[h] ? y : m[h];
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < HC; ++h) {
                uint64_t v = m[h];
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    uint64_t other = __shfl_xor_sync(0xffffffffu, v, o);
                    v = other < v ? other : v;
                }
                if (lane == 0 && c0 + h < a.n && v != ~0ULL)
                    atomicMin(reinterpret_cast<unsigned long long *>(a.sk + (size_t)g.read * a.n + c0 + h),
                              (a.rnd[c0 + h] & ~mask) | v);
            }
        }
    }
}

} // namespace nsmh

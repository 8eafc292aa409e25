// The online query in ONE kernel: ReadFilter::getFilteredReads(const std::string&, results)
// (src/ReadFilter.cpp:85-97), the call the consensus builder makes twice per window from every OpenMP
// thread (src/Consensus.cpp:180-191).  The bulk path spreads a query over ~15 launches, five small copies
// and three host round trips - fine for 10^5 reads at once, far too slow for one 10 kb window.  Here a block
// per string does everything:
//   1. reads the ASCII window straight from the caller's pinned host buffer (mapped memory, no copy call),
//      2-bit packs it into shared memory (code (c&2)|((c&4)>>2), pack_kernels.cuh's layout)
//   2. sketches it: a warp per hash function at a time; every lane turns its words into the 16 leading
//      32-bit windows once and takes the minimum of (window ^ target) for the warp's hash functions; the
//      winning word is redone in 64 bits, a 32-bit tie falls back to a full 64-bit scan (the exact scheme of
//      sketch_fixup_kernel) - string2Sketch, ReadFilter.cpp:117-131, incl. the short-string rules
//   3. probes the n tables and thresholds the gathered ids (count_queries, the bulk lookup's own code)
//   4. writes count, position and ids into mapped host memory: the host only waits for the stream.
// One launch + one synchronisation per call.  Strings beyond kOnlineMaxBases, queries whose id lists need
// the heavy tiers and results beyond the buffer are flagged and repeated on the general path by the host.
#pragma once
#include <stdint.h>

#include "nsmh_constants.h"
#include "query_kernels.cuh"
#include "sketch_device.cuh"

namespace nsmh {

constexpr uint32_t kOnlineMaxBases = 1u << 17;      // 32 KB of packed words in shared memory
constexpr int kOnlineThreads = 512;
constexpr uint32_t kOnlineOverflowIds = 8192;       // room behind the fixed places for longer result lists

// the tables probed with keys that live in shared memory
struct OnlineSrc {
    static constexpr bool kInlinePairs = false;
    ProbeSrc t;
    const uint64_t *keys;           // [n] shared memory
    using Pending = ProbeSrc::Pending;
    using Ctx = ProbeSrc::Ctx;
    __device__ __forceinline__ uint32_t subs() const { return t.n; }
    __device__ __forceinline__ void prefetch(uint32_t, int) const {}
    __device__ __forceinline__ Ctx context(uint32_t j) const { return t.context(j); }
    __device__ __forceinline__ Pending begin(const Ctx &x, uint32_t, uint32_t j) const {
        Pending p;
        p.j = j;
        p.key = keys[j];
        p.b = p.key == kEmptyKey ? (t.cap >> 1) : slot_index(p.key, t.cap >> 1);
        ldg256(x.region + 2 * p.b, p.sa, p.sb, p.sc, p.sd);
        return p;
    }
    __device__ __forceinline__ Pending begin(uint32_t q, uint32_t j) const { return begin(context(j), q, j); }
    __device__ __forceinline__ ListRef finish(Pending p) const { return t.finish(p); }
    __device__ __forceinline__ ListRef get(uint32_t q, uint32_t j) const { return finish(begin(q, j)); }
};

constexpr uint32_t kOnlineIdsPerQuery = kFixedIds + kOnlineOverflowIds;     // result ids a query can return through this path

struct OnlineArgs {
    const uint8_t *text;            // the strings (mapped host memory or device memory); every start 16-byte aligned,
    const uint64_t *start, *len;    // [nq] byte offset and length of string q; 16 readable bytes past every end
    const uint64_t *rnd;            // [n]
    ProbeSrc tables;
    // where the host reads the answer: qcount [nq], qpos [nq] (position inside the query's own kOnlineIdsPerQuery
    // ids of tmp_ids; ~0: the general path must take over), tmp_ids [nq][kOnlineIdsPerQuery]
    uint32_t *qcount;
    uint64_t *qpos;
    uint32_t *tmp_ids;
    uint32_t *heavy_scratch;        // [nq] device
    unsigned long long *counters;   // [nq][8] device scratch
    uint64_t *sketch_out;           // optional [nq][n]
    uint32_t k, thr;
};

// shared memory: packed words (len/16 + 4), n keys, the lookup's warp buffer
__host__ __device__ __forceinline__ size_t online_smem_bytes(uint32_t max_len, uint32_t n) {
    return ((size_t)(max_len / kWordBases + 4) * 4 + 15 & ~(size_t)15) + (size_t)n * 8 + (size_t)kWarpWords * 4;
}

__global__ void __launch_bounds__(kOnlineThreads)
online_query_kernel(OnlineArgs a, uint32_t max_len) {
    extern __shared__ __align__(16) uint8_t o_smem[];
    uint32_t *W = reinterpret_cast<uint32_t *>(o_smem);
    const size_t w_bytes = (size_t)(max_len / kWordBases + 4) * 4 + 15 & ~(size_t)15;
    uint64_t *keys = reinterpret_cast<uint64_t *>(o_smem + w_bytes);
    uint32_t *cbuf = reinterpret_cast<uint32_t *>(keys + a.tables.n);
    const uint32_t q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kOnlineThreads / 32;
    const uint64_t len = a.len[q];
    const uint32_t n = a.tables.n, k = a.k;
    unsigned long long *counters = a.counters + 8 * (size_t)q;
    if (threadIdx.x < 8) counters[threadIdx.x] = 0;

    // ---- 1. pack: thread t makes word t, t + 512, ... from ONE 16-byte load (bases past the end: zero bits) ----
    const uint32_t nwords = (uint32_t)((len + kWordBases - 1) / kWordBases);
    const uint4 *src = reinterpret_cast<const uint4 *>(a.text + a.start[q]);
    for (uint32_t w = threadIdx.x; w < nwords + 3; w += kOnlineThreads) {
        uint32_t v = 0;
        if (w < nwords) {
            const uint4 c16 = src[w];
            const uint32_t cc[4] = {c16.x, c16.y, c16.z, c16.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // 4 bytes -> 4 codes (c&2)|((c&4)>>2), first byte most significant
                const uint32_t t = ((cc[i] & 0x02020202u) | ((cc[i] >> 2) & 0x01010101u));
                v = (v << 8) | ((t & 0xFF) << 6) | (((t >> 8) & 0xFF) << 4) | (((t >> 16) & 0xFF) << 2) | (t >> 24);
            }
            const uint64_t p0 = (uint64_t)w * kWordBases;
            if (p0 + kWordBases > len) v &= ~0u << (2 * (uint32_t)(p0 + kWordBases - len));
        }
        W[w] = v;
    }
    __syncthreads();

    // ---- 2. sketch (ReadFilter.cpp:117-131) ----
    const uint64_t mask = kmer_mask(k);
    if (len + 1 < k) {
        for (uint32_t l = threadIdx.x; l < n; l += kOnlineThreads) keys[l] = 0;             // untouched, zero-initialised sketch
    } else if (len + 1 == k) {
        for (uint32_t l = threadIdx.x; l < n; l += kOnlineThreads) keys[l] = ~0ULL;
    } else {
        TileGeom g;
        g.read = 0;
        g.rb = 0;
        g.nk = len - k + 1;
        g.w_begin = 0;
        g.w_end = (g.nk - 1) / kWordBases + 1;
        const int kshift = 64 - 2 * (int)k;
        const int s_lo = 2 * (int)k > 32 ? 2 * (int)k - 32 : 0;      // y32 = y >> s_lo
        const int sh = 2 * (int)k >= 32 ? 0 : 32 - 2 * (int)k;       // window >> sh = leading bits of the k-mer
        constexpr int HC = 2;                                           // hash functions per warp and pass
        for (uint32_t l0 = warp * HC; l0 < n; l0 += nwarps * HC) {
            uint64_t r[HC], rlo[HC];
            uint32_t t32[HC], m[HC];
            uint64_t mw[HC];
            bool has[HC], tie[HC];
#pragma unroll
            for (int h = 0; h < HC; ++h) {
                r[h] = a.rnd[min(l0 + h, n - 1)];
                rlo[h] = r[h] & mask;
                t32[h] = (uint32_t)(rlo[h] >> s_lo);
                m[h] = 0xFFFFFFFFu;
                mw[h] = 0;
                has[h] = tie[h] = false;
            }
            for (uint64_t w = lane; w < g.w_end; w += 32) {
                const uint32_t w0 = W[w], w1 = W[w + 1];
                int lo, hi;
                valid_range(g, w, lo, hi);
                uint32_t local[HC];
#pragma unroll
                for (int h = 0; h < HC; ++h) local[h] = 0xFFFFFFFFu;
                if (lo == 0 && hi == kWordBases) {
#pragma unroll
                    for (int j = 0; j < kWordBases; ++j) {
                        const uint32_t x = (j ? __funnelshift_l(w1, w0, 2 * j) : w0) >> sh;
#pragma unroll
                        for (int h = 0; h < HC; ++h) local[h] = min(local[h], x ^ t32[h]);
                    }
                } else {
                    for (int j = lo; j < hi; ++j) {
                        const uint32_t x = __funnelshift_l(w1, w0, 2 * j) >> sh;
#pragma unroll
                        for (int h = 0; h < HC; ++h) local[h] = min(local[h], x ^ t32[h]);
                    }
                }
                if (lo < hi) {
#pragma unroll
                    for (int h = 0; h < HC; ++h) {
                        if (!has[h] || local[h] < m[h]) { m[h] = local[h]; mw[h] = w; tie[h] = false; has[h] = true; }
                        else if (local[h] == m[h]) tie[h] = true;
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < HC; ++h) {
                uint32_t gm = has[h] ? m[h] : 0xFFFFFFFFu;
#pragma unroll
                for (int o = 16; o; o >>= 1) gm = min(gm, __shfl_xor_sync(0xffffffffu, gm, o));
                const bool cand = has[h] && m[h] == gm;
                uint64_t best = ~0ULL;
                if (__any_sync(0xffffffffu, cand && tie[h])) {
                    for (uint64_t w = lane; w < g.w_end; w += 32) {
                        int lo, hi;
                        valid_range(g, w, lo, hi);
                        const uint64_t v = word_min64<false>(W, w, lo, hi, kshift, rlo[h]);
                        best = v < best ? v : best;
                    }
                } else if (cand) {
                    int lo, hi;
                    valid_range(g, mw[h], lo, hi);
                    best = word_min64<false>(W, mw[h], lo, hi, kshift, rlo[h]);
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
                    best = other < best ? other : best;
                }
                if (lane == 0 && l0 + h < n) keys[l0 + h] = (r[h] & ~mask) | best;
            }
        }
    }
    __syncthreads();
    if (a.sketch_out)
        for (uint32_t l = threadIdx.x; l < n; l += kOnlineThreads) a.sketch_out[(size_t)q * n + l] = keys[l];

    // ---- 3. + 4. probe, count, hand over (one warp; the bulk lookup's code) ----
    if (warp == 0) {
        OnlineSrc os;
        os.t = a.tables;
        os.keys = keys;
        CountArgs ca;               // this query alone: its own result area, count, position and scratch
        ca.qcount = a.qcount + q;
        ca.qpos = a.qpos + q;
        ca.tmp_ids = a.tmp_ids + (size_t)q * kOnlineIdsPerQuery;
        ca.tmp_cap = kOnlineIdsPerQuery;
        ca.heavy_list = a.heavy_scratch + q;
        ca.counters = counters;
        ca.nq = 1;
        ca.thr = a.thr;
        if (n <= 64) count_queries<OnlineSrc, 2>(os, ca, cbuf, 0, 0x7FFFFFFFu);
        else count_queries<OnlineSrc, kRegListsMax>(os, ca, cbuf, 0, 0x7FFFFFFFu);
    }
}

} // namespace nsmh

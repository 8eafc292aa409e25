// MinHash sketching on the device.  Replaces the hot loop of
//   MinHashReadFilter::initialize        (src/ReadFilter.cpp:31-44)
//   MinHashReadFilter::string2Sketch     (src/ReadFilter.cpp:117-131)
//   MinHashReadFilter::string2KMers      (src/ReadFilter.cpp:138-152)
//   MinHashReadFilter::hashKMer          (src/ReadFilter.cpp:133-136)
// Result per read and hash l (bit-exact with the reference):
//   sketch[l] = min over forward k-mers x of (x XOR rand[l])   (u64 compare)
//   len < k-1 -> 0 (row never written by the reference), len == k-1 -> ~0.
//
// Two kernels produce the same numbers:
//
//  * sketch_brute_kernel — the reference's operation count: every k-mer against
//    every hash (n XOR+MIN pairs per k-mer), minima kept in registers, combined
//    with warp shuffles.  This is the kernel the INT32-pipe roofline describes.
//
//  * sketch_filter_kernel (default) — an exact shortcut.  With K = 2k and
//    r = rand[l] & (2^K-1):  x XOR rand[l] = (rand[l] & ~mask) | (x ^ r), so only
//    y = x ^ r matters, and y < 2^(K-b) exactly when the top b bits of x equal
//    the top b bits of r.  If any k-mer of the read matches hash l on those b
//    bits, the minimum is among the matching k-mers.  The top b bits of the
//    k-mer starting at base p are just the b-bit window of the packed stream at
//    p, so the scan is funnel shifts + shared-memory table lookups (one lookup
//    answers three consecutive positions) and touches the 64-bit path only for
//    the ~n*lambda positions per read that hit, lambda = 2^lambda_log2 .. twice
//    that being the expected k-mers per bucket (b = floor(log2 #kmers) - lambda_log2).
//    Hits of a 512-position block are compacted and processed 32 at a time so all
//    lanes stay busy.  A (read, hash) pair whose bucket stayed empty (probability
//    ~e^-lambda on random sequence, certain on e.g. homopolymers) is rescanned
//    exhaustively by sketch_fixup_kernel, which makes the result unconditional.
#include <cstdlib>

#include "nsmh_internal.cuh"

namespace nsmh {

struct SketchArgs {
    const uint64_t *off;        // [n_reads+1] global base offsets
    const uint32_t *W;          // packed stream
    uint64_t *sk;               // [n_reads][n]
    const uint32_t *tile_start; // [n_reads+1] exclusive scan of tiles per read
    const uint32_t *tile_read;  // [num_tiles] read of every tile (nullptr: binary search)
    const uint64_t *rnd;        // [n]
    const uint8_t *ftab_hit, *ftab_first, *ftab_next, *ftab_hit3;
    unsigned long long *counters;   // [0] fix-ups
    uint32_t n_reads, k, n;
    int lambda_log2;
    uint32_t tile_words;        // words (16 k-mer starts each) per tile
};

__device__ __forceinline__ uint64_t kmer_mask(uint32_t k) { return (1ULL << (2 * k)) - 1; }

// ---- row init + tile counts ---------------------------------------------------
__global__ void __launch_bounds__(256)
sketch_init_kernel(SketchArgs a, uint32_t *__restrict__ tile_cnt) {
    const uint64_t total = (uint64_t)a.n_reads * a.n;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
         t += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t i = (uint32_t)(t / a.n);
        uint32_t l = (uint32_t)(t - (uint64_t)i * a.n);
        uint64_t b0 = a.off[i], len = a.off[i + 1] - b0;
        // ReadFilter.cpp:119-124: untouched (zero) when len-k+1 < 0, all-ones otherwise
        a.sk[t] = (len + 1 < a.k) ? 0ULL : ~0ULL;
        if (l == 0) {
            uint32_t tiles = 0;
            if (len >= a.k) {
                uint64_t nk = len - a.k + 1;
                uint64_t w0 = b0 / kWordBases, w1 = (b0 + nk - 1) / kWordBases;
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

struct TileGeom {
    uint32_t read;
    uint64_t rb;        // first base of the read (global)
    uint64_t nk;        // number of k-mers
    uint64_t w_begin, w_end;   // word range of this tile (global word indices)
};

__device__ __forceinline__ TileGeom tile_geom(const SketchArgs &a, uint32_t tile) {
    TileGeom g;
    g.read = a.tile_read ? a.tile_read[tile] : find_read_of_tile(a.tile_start, a.n_reads, tile);
    g.rb = a.off[g.read];
    g.nk = a.off[g.read + 1] - g.rb - a.k + 1;
    uint64_t w0 = g.rb / kWordBases, w1 = (g.rb + g.nk - 1) / kWordBases;
    g.w_begin = w0 + (uint64_t)(tile - a.tile_start[g.read]) * a.tile_words;
    g.w_end = g.w_begin + a.tile_words < w1 + 1 ? g.w_begin + a.tile_words : w1 + 1;
    return g;
}

// valid k-mer start positions of word w: j in [lo, hi)
__device__ __forceinline__ void valid_range(const TileGeom &g, uint64_t w, int &lo, int &hi) {
    uint64_t p0 = w * kWordBases;
    lo = g.rb > p0 ? (int)(g.rb - p0) : 0;
    uint64_t end = g.rb + g.nk;   // one past the last k-mer start
    hi = end >= p0 + kWordBases ? kWordBases : (end > p0 ? (int)(end - p0) : 0);
}

__device__ __forceinline__ int filter_bits(uint64_t nk, int lambda_log2, int max_bits, uint32_t k) {
    int b = 63 - __clzll((long long)nk) - lambda_log2;
    b = b < 0 ? 0 : b;
    b = b > max_bits ? max_bits : b;
    b = b > 2 * (int)k ? 2 * (int)k : b;
    return b;
}

// 64-bit k-mer starting at base j of word w0 (w1, w2 are the following words)
__device__ __forceinline__ uint64_t kmer_at(uint32_t w0, uint32_t w1, uint32_t w2, int j, int kshift,
                                            uint32_t &h32) {
    h32 = __funnelshift_l(w1, w0, 2 * j);
    uint32_t l32 = __funnelshift_l(w2, w1, 2 * j);
    return (((uint64_t)h32 << 32) | l32) >> kshift;
}

// ---- filter kernel: 3 positions per lookup, dense hit rounds ---------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
sketch_filter_kernel(SketchArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *s_hit3 = smem;                                   // kFilter3TabSize
    uint8_t *s_first = s_hit3 + kFilter3TabSize;              // 2^(kFilter3MaxBits+1)
    ulonglong2 *s_min = reinterpret_cast<ulonglong2 *>(s_first + (2 << kFilter3MaxBits));   // WARPS * n pairs
    uint64_t *s_rlo = reinterpret_cast<uint64_t *>(s_min + (size_t)WARPS * a.n);            // n
    uint16_t *s_list = reinterpret_cast<uint16_t *>(s_rlo + a.n);                           // WARPS * 512
    uint8_t *s_next = reinterpret_cast<uint8_t *>(s_list + WARPS * 512);

    const uint64_t mask = kmer_mask(a.k);
    {
        const uint4 *g3 = reinterpret_cast<const uint4 *>(a.ftab_hit3);
        const uint4 *gf = reinterpret_cast<const uint4 *>(a.ftab_first);
        uint4 *s3 = reinterpret_cast<uint4 *>(s_hit3), *sf = reinterpret_cast<uint4 *>(s_first);
        for (int t = threadIdx.x; t < kFilter3TabSize / 16; t += WARPS * 32) s3[t] = g3[t];
        for (int t = threadIdx.x; t < (2 << kFilter3MaxBits) / 16; t += WARPS * 32) sf[t] = gf[t];
        for (uint32_t t = threadIdx.x; t < (kFilterMaxBits + 1) * a.n; t += WARPS * 32)
            s_next[t] = a.ftab_next[t];
        for (uint32_t t = threadIdx.x; t < a.n; t += WARPS * 32) s_rlo[t] = a.rnd[t] & mask;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ulonglong2 *my_min = s_min + (size_t)warp * a.n;
    uint16_t *my_list = s_list + warp * 512;
    const uint32_t num_tiles = a.tile_start[a.n_reads];
    const int kshift = 64 - 2 * (int)a.k;

    for (uint32_t tile = blockIdx.x * WARPS + warp; tile < num_tiles; tile += gridDim.x * WARPS) {
        const TileGeom g = tile_geom(a, tile);
        const int b = filter_bits(g.nk, a.lambda_log2, kFilter3MaxBits, a.k);
        const int rshift3 = 32 - (b + 4);       // window of b+4 bits answers positions j, j+1, j+2
        const int rshift = 32 - b;
        const uint8_t *nxt = s_next + (size_t)b * a.n;

        for (uint32_t l = lane; l < a.n; l += 32) my_min[l] = make_ulonglong2(s_rlo[l], ~0ULL);
        __syncwarp();

        // software pipeline: the words of the next 512-position block are in flight while
        // this one is scanned
        uint32_t n0 = 0, n1 = 0, n2 = 0;
        if (g.w_begin + lane < g.w_end) {
            n0 = __ldg(a.W + g.w_begin + lane);
            n1 = __ldg(a.W + g.w_begin + lane + 1);
            n2 = __ldg(a.W + g.w_begin + lane + 2);
        }
        for (uint64_t wb = g.w_begin; wb < g.w_end; wb += 32) {
            const uint64_t w = wb + lane;
            const uint32_t w0 = n0, w1 = n1, w2 = n2;
            n0 = n1 = n2 = 0;
            if (w + 32 < g.w_end) {
                n0 = __ldg(a.W + w + 32);
                n1 = __ldg(a.W + w + 33);
                n2 = __ldg(a.W + w + 34);
            }
            int lo = 0, hi = 0;
            if (w < g.w_end) valid_range(g, w, lo, hi);
            // phase 1: six lookups cover positions 0..17; bits beyond 15 are dropped
            uint32_t hits = 0;
#pragma unroll
            for (int t = 0; t < 6; ++t) {
                uint32_t v = t ? __funnelshift_l(w1, w0, 6 * t) : w0;
                uint32_t idx = __funnelshift_rc(v, 1u, rshift3);   // (1<<(b+4)) | window
                hits = hits * 8 + s_hit3[idx];
            }
            hits = (hits >> 2) & (0xFFFFu >> lo) & (0xFFFFu << (16 - hi)) & 0xFFFFu;

            // phase 2: compact the block's hits, then 32 per round with every lane busy
            const uint32_t cnt = __popc(hits);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (total == 0) continue;
            uint32_t pos = incl - cnt;
            while (hits) {
                const int top = 31 - __clz(hits);
                hits ^= 1u << top;
                my_list[pos++] = (uint16_t)((lane << 4) | (15 - top));
            }
            __syncwarp();
            for (uint32_t e0 = 0; e0 < total; e0 += 32) {
                const uint32_t e = e0 + lane;
                const uint32_t ent = e < total ? my_list[e] : 0u;
                const int src = ent >> 4, j = ent & 15;
                const uint32_t a0 = __shfl_sync(0xffffffffu, w0, src);
                const uint32_t a1 = __shfl_sync(0xffffffffu, w1, src);
                const uint32_t a2 = __shfl_sync(0xffffffffu, w2, src);
                if (e < total) {
                    uint32_t h32;
                    const uint64_t x = kmer_at(a0, a1, a2, j, kshift, h32);
                    uint32_t l = s_first[__funnelshift_rc(h32, 1u, rshift)];
                    do {
                        const ulonglong2 rm = my_min[l];        // {rand[l] & mask, running minimum}
                        const uint32_t ln = nxt[l];
                        const uint64_t y = x ^ rm.x;
                        if (y < rm.y) atomicMin(&my_min[l].y, (unsigned long long)y);
                        l = ln;
                    } while (l != 0xFFu);
                }
            }
            __syncwarp();
        }
        __syncwarp();
        for (uint32_t l = lane; l < a.n; l += 32) {
            uint64_t v = my_min[l].y;
            if (v != ~0ULL)
                atomicMin(reinterpret_cast<unsigned long long *>(a.sk + (size_t)g.read * a.n + l),
                          (a.rnd[l] & ~mask) | v);
        }
        __syncwarp();
    }
}

// ---- first version of the filter kernel (one lookup per position, hits handled
//      inline by the lane that found them); kept for A/B measurements -------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
sketch_filter1_kernel(SketchArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *s_hit = smem;
    uint8_t *s_first = s_hit + kFilterTabSize;
    uint64_t *s_rlo = reinterpret_cast<uint64_t *>(s_first + kFilterTabSize);
    uint64_t *s_min = s_rlo + a.n;
    uint8_t *s_next = reinterpret_cast<uint8_t *>(s_min + (size_t)WARPS * a.n);

    const uint64_t mask = kmer_mask(a.k);
    {
        const uint4 *gh = reinterpret_cast<const uint4 *>(a.ftab_hit);
        const uint4 *gf = reinterpret_cast<const uint4 *>(a.ftab_first);
        uint4 *sh = reinterpret_cast<uint4 *>(s_hit), *sf = reinterpret_cast<uint4 *>(s_first);
        for (int t = threadIdx.x; t < kFilterTabSize / 16; t += WARPS * 32) {
            sh[t] = gh[t];
            sf[t] = gf[t];
        }
        for (uint32_t t = threadIdx.x; t < (kFilterMaxBits + 1) * a.n; t += WARPS * 32)
            s_next[t] = a.ftab_next[t];
        for (uint32_t t = threadIdx.x; t < a.n; t += WARPS * 32) s_rlo[t] = a.rnd[t] & mask;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t *my_min = s_min + (size_t)warp * a.n;
    const uint32_t num_tiles = a.tile_start[a.n_reads];
    const int kshift = 64 - 2 * (int)a.k;

    for (uint32_t tile = blockIdx.x * WARPS + warp; tile < num_tiles; tile += gridDim.x * WARPS) {
        const TileGeom g = tile_geom(a, tile);
        const int b = filter_bits(g.nk, a.lambda_log2, kFilterMaxBits, a.k);
        const int rshift = 32 - b;
        const uint8_t *nxt = s_next + (size_t)b * a.n;

        for (uint32_t l = lane; l < a.n; l += 32) my_min[l] = ~0ULL;
        __syncwarp();

        for (uint64_t wb = g.w_begin; wb < g.w_end; wb += 32) {
            const uint64_t w = wb + lane;
            uint32_t w0 = 0, w1 = 0, w2 = 0;
            int lo = 0, hi = 0;
            if (w < g.w_end) {
                w0 = __ldg(a.W + w);
                w1 = __ldg(a.W + w + 1);
                w2 = __ldg(a.W + w + 2);
                valid_range(g, w, lo, hi);
            }
            uint32_t hits = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                uint32_t v = j ? __funnelshift_l(w1, w0, 2 * j) : w0;
                uint32_t idx = __funnelshift_rc(v, 1u, rshift);   // (1<<b) | top b bits
                hits = hits * 2 + s_hit[idx];
            }
            hits &= (0xFFFFu >> lo) & (0xFFFFu << (16 - hi)) & 0xFFFFu;
            while (hits) {
                const int top = 31 - __clz(hits);
                hits ^= 1u << top;
                uint32_t h32;
                const uint64_t x = kmer_at(w0, w1, w2, 15 - top, kshift, h32);
                uint32_t l = s_first[__funnelshift_rc(h32, 1u, rshift)];
                do {
                    uint64_t y = x ^ s_rlo[l];
                    if (y < my_min[l]) atomicMin(reinterpret_cast<unsigned long long *>(my_min + l), y);
                    l = nxt[l];
                } while (l != 0xFFu);
            }
        }
        __syncwarp();
        for (uint32_t l = lane; l < a.n; l += 32) {
            uint64_t v = my_min[l];
            if (v != ~0ULL)
                atomicMin(reinterpret_cast<unsigned long long *>(a.sk + (size_t)g.read * a.n + l),
                          (a.rnd[l] & ~mask) | v);
        }
        __syncwarp();
    }
}

// ---- exact fix-up: (read, hash) pairs that no k-mer matched on the filter prefix ----
// One warp per read checks the row; every missing hash is recomputed over all k-mers of
// the read, 512 positions per warp step (coalesced word loads, 16 k-mers per lane).
__global__ void __launch_bounds__(256)
sketch_fixup_kernel(SketchArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    const uint64_t mask = kmer_mask(a.k);
    const int kshift = 64 - 2 * (int)a.k;
    for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < a.n_reads; i += warps) {
        const uint64_t rb = a.off[i], len = a.off[i + 1] - rb;
        if (len < a.k) continue;
        TileGeom g;
        g.read = i;
        g.rb = rb;
        g.nk = len - a.k + 1;
        g.w_begin = rb / kWordBases;
        g.w_end = (rb + g.nk - 1) / kWordBases + 1;
        for (uint32_t l0 = 0; l0 < a.n; l0 += 32) {
            uint32_t l = l0 + lane;
            bool miss = l < a.n && a.sk[(size_t)i * a.n + l] == ~0ULL;
            uint32_t todo = __ballot_sync(0xffffffffu, miss);
            while (todo) {
                const uint32_t lf = l0 + (__ffs(todo) - 1);
                todo &= todo - 1;
                const uint64_t r = a.rnd[lf], rlo = r & mask;
                uint64_t best = ~0ULL;
                for (uint64_t wb = g.w_begin; wb < g.w_end; wb += 32) {
                    const uint64_t w = wb + lane;
                    if (w >= g.w_end) continue;
                    int lo, hi;
                    valid_range(g, w, lo, hi);
                    const uint32_t w0 = __ldg(a.W + w), w1 = __ldg(a.W + w + 1), w2 = __ldg(a.W + w + 2);
                    for (int j = lo; j < hi; ++j) {
                        uint32_t h32;
                        const uint64_t y = kmer_at(w0, w1, w2, j, kshift, h32) ^ rlo;
                        best = y < best ? y : best;
                    }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
                    best = other < best ? other : best;
                }
                if (lane == 0) {
                    a.sk[(size_t)i * a.n + lf] = (r & ~mask) | best;
                    atomicAdd(a.counters, 1ULL);
                }
            }
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

// ---- host side --------------------------------------------------------------------
// Per prefix width b: hit/first tables at [2^b, 2^(b+1)), a per-b chain of hashes that
// share a target prefix, and the 3-position table at [2^(b+4), 2^(b+5)).
int build_filter_tables(nsmh_ctx *c) {
    const uint32_t n = c->n, k = c->k;
    std::vector<uint8_t> hit(kFilterTabSize, 0), first(kFilterTabSize, 0xFF);
    std::vector<uint8_t> next((size_t)(kFilterMaxBits + 1) * (n ? n : 1), 0xFF);
    std::vector<uint8_t> hit3(kFilter3TabSize, 0);
    if (n <= 255) {
        const uint64_t mask = (1ULL << (2 * k)) - 1;
        for (int b = 0; b <= kFilterMaxBits && b <= 2 * (int)k; ++b) {
            for (uint32_t l = 0; l < n; ++l) {
                uint64_t rlo = c->rand[l] & mask;
                uint32_t t = b ? (uint32_t)(rlo >> (2 * k - b)) : 0u;
                uint32_t idx = (1u << b) | t;
                hit[idx] = 1;
                next[(size_t)b * n + l] = first[idx];
                first[idx] = (uint8_t)l;
            }
        }
        for (int b = 0; b <= kFilter3MaxBits && b <= 2 * (int)k; ++b) {
            const uint32_t pm = (1u << b) - 1, base = 1u << b;
            for (uint32_t wv = 0; wv < (1u << (b + 4)); ++wv) {
                uint32_t p0 = (wv >> 4) & pm, p1 = (wv >> 2) & pm, p2 = wv & pm;
                hit3[(1u << (b + 4)) | wv] =
                    (uint8_t)((hit[base | p0] << 2) | (hit[base | p1] << 1) | hit[base | p2]);
            }
        }
    }
    const char *e = getenv("NSMH_LAMBDA_LOG2");
    if (e && *e) c->lambda_log2 = atoi(e) < 0 ? 0 : (atoi(e) > 8 ? 8 : atoi(e));
    e = getenv("NSMH_TILE_WORDS");
    if (e && *e && atoi(e) >= 32) c->tile_words = (uint32_t)atoi(e);
    e = getenv("NSMH_SKETCH_VARIANT");
    if (e && *e) c->sketch_variant = atoi(e) ? 1 : 0;
    NSMH_TRY(c->d_ftab_hit.ensure(hit.size(), c->stream));
    NSMH_TRY(c->d_ftab_first.ensure(first.size(), c->stream));
    NSMH_TRY(c->d_ftab_next.ensure(next.size(), c->stream));
    NSMH_TRY(c->d_ftab_hit3.ensure(hit3.size(), c->stream));
    NSMH_CK(cudaMemcpyAsync(c->d_ftab_hit.p, hit.data(), hit.size(), cudaMemcpyHostToDevice, c->stream));
    NSMH_CK(cudaMemcpyAsync(c->d_ftab_first.p, first.data(), first.size(), cudaMemcpyHostToDevice, c->stream));
    NSMH_CK(cudaMemcpyAsync(c->d_ftab_next.p, next.data(), next.size(), cudaMemcpyHostToDevice, c->stream));
    NSMH_CK(cudaMemcpyAsync(c->d_ftab_hit3.p, hit3.data(), hit3.size(), cudaMemcpyHostToDevice, c->stream));
    NSMH_CK(cudaStreamSynchronize(c->stream));   // host vectors die here
    return NSMH_OK;
}

int sketch_reads(nsmh_ctx *c, const ReadSet &rs, uint64_t *d_sketches, DevBuf &tile_start,
                 DevBuf &cub_tmp, int mode, cudaStream_t s, uint32_t *launches, cudaEvent_t ev0,
                 cudaEvent_t ev1) {
    if (rs.num_reads == 0) return NSMH_OK;
    constexpr int WARPS = 8;
    SketchArgs a;
    a.off = rs.d_offsets();
    a.W = rs.packed.as<uint32_t>();
    a.sk = d_sketches;
    a.rnd = c->d_rand.as<uint64_t>();
    a.ftab_hit = c->d_ftab_hit.as<uint8_t>();
    a.ftab_first = c->d_ftab_first.as<uint8_t>();
    a.ftab_next = c->d_ftab_next.as<uint8_t>();
    a.ftab_hit3 = c->d_ftab_hit3.as<uint8_t>();
    a.counters = c->counters.as<unsigned long long>();
    a.n_reads = rs.num_reads;
    a.k = c->k;
    a.n = c->n;
    a.lambda_log2 = c->lambda_log2;
    a.tile_words = c->tile_words;

    // tile_start: [0..n_reads] exclusive scan, followed by the per-read counts
    // upper bound on the number of tiles: ceil(words_i / T) <= words_i / T + 1 per read, and the
    // reads' word ranges overlap by at most one word each
    const size_t max_tiles = (size_t)((rs.num_words + rs.num_reads) / c->tile_words) + rs.num_reads + 1;
    NSMH_TRY(tile_start.ensure((((size_t)rs.num_reads + 1) * 2 + max_tiles) * sizeof(uint32_t), s));
    uint32_t *cnt = tile_start.as<uint32_t>() + rs.num_reads + 1;
    uint32_t *ts = tile_start.as<uint32_t>();
    uint32_t *tile_read = cnt + rs.num_reads + 1;
    a.tile_start = ts;
    a.tile_read = tile_read;
    NSMH_CK(cudaMemsetAsync(cnt + rs.num_reads, 0, sizeof(uint32_t), s));
    const uint64_t total = (uint64_t)rs.num_reads * c->n;
    int blocks = (int)((total + 255) / 256 < (uint64_t)c->num_sms * 8 ? (total + 255) / 256
                                                                       : (uint64_t)c->num_sms * 8);
    sketch_init_kernel<<<blocks, 256, 0, s>>>(a, cnt);
    ++*launches;
    NSMH_CK(cudaGetLastError());
    size_t tmp_bytes = 0;
    NSMH_CK(cub_exclusive_sum_u32(nullptr, tmp_bytes, cnt, ts, (size_t)rs.num_reads + 1, s));
    NSMH_TRY(cub_tmp.ensure(tmp_bytes, s));
    NSMH_CK(cub_exclusive_sum_u32(cub_tmp.p, tmp_bytes, cnt, ts, (size_t)rs.num_reads + 1, s));
    sketch_tile_map_kernel<<<(rs.num_reads + 255) / 256, 256, 0, s>>>(ts, rs.num_reads, tile_read);
    *launches += 3;
    NSMH_CK(cudaGetLastError());

    if (mode == 0 && c->n > 255) mode = 1;   // chain tables index hashes with one byte
    if (ev0) NSMH_CK(cudaEventRecord(ev0, s));
    if (mode == 0) {
        static bool attr_set = false;
        if (!attr_set) {
            NSMH_CK(cudaFuncSetAttribute(sketch_filter_kernel<WARPS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            NSMH_CK(cudaFuncSetAttribute(sketch_filter1_kernel<WARPS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set = true;
        }
        const size_t common = (size_t)c->n * 8 + (size_t)WARPS * c->n * 8 + (size_t)(kFilterMaxBits + 1) * c->n + 16;
        int occ = 0;
        if (c->sketch_variant == 0) {
            const size_t smem = kFilter3TabSize + (2 << kFilter3MaxBits) + common + (size_t)WARPS * c->n * 8 +
                                WARPS * 512 * sizeof(uint16_t);
            NSMH_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sketch_filter_kernel<WARPS>, WARPS * 32, smem));
            if (occ < 1) return fail(NSMH_EINVAL, "sketch: n too large for the filter kernel's shared memory");
            sketch_filter_kernel<WARPS><<<c->num_sms * occ, WARPS * 32, smem, s>>>(a);
        } else {
            const size_t smem = 2 * (size_t)kFilterTabSize + common;
            NSMH_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sketch_filter1_kernel<WARPS>, WARPS * 32, smem));
            if (occ < 1) return fail(NSMH_EINVAL, "sketch: n too large for the filter kernel's shared memory");
            sketch_filter1_kernel<WARPS><<<c->num_sms * occ, WARPS * 32, smem, s>>>(a);
        }
        ++*launches;
        NSMH_CK(cudaGetLastError());
        if (ev1) NSMH_CK(cudaEventRecord(ev1, s));
        sketch_fixup_kernel<<<c->num_sms * 8, 256, 0, s>>>(a);
        ++*launches;
        NSMH_CK(cudaGetLastError());
    } else {
        int occ = 0;
        NSMH_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sketch_brute_kernel<8>, 256, 0));
        sketch_brute_kernel<8><<<c->num_sms * (occ > 0 ? occ : 1), 256, 0, s>>>(a);
        ++*launches;
        NSMH_CK(cudaGetLastError());
        if (ev1) NSMH_CK(cudaEventRecord(ev1, s));
    }
    return NSMH_OK;
}

} // namespace nsmh

// Device helpers of the sketch code that more than one translation unit needs (the sketch kernels of
// sketch_kernels.cuh and the online query kernel of online_kernels.cuh): no kernels here.
#pragma once
#include <stdint.h>

#include "nsmh_constants.h"

namespace nsmh {

__device__ __forceinline__ uint64_t kmer_mask(uint32_t k) { return (1ULL << (2 * k)) - 1; }


struct TileGeom {
    uint32_t read;
    uint64_t rb;        // first base of the read (global)
    uint64_t nk;        // number of k-mers
    uint64_t w_begin, w_end;   // word range of this tile (global word indices)
};


// valid k-mer start positions of word w: j in [lo, hi)
__device__ __forceinline__ void valid_range(const TileGeom &g, uint64_t w, int &lo, int &hi) {
    uint64_t p0 = w * kWordBases;
    lo = g.rb > p0 ? (int)(g.rb - p0) : 0;
    uint64_t end = g.rb + g.nk;   // one past the last k-mer start
    hi = end >= p0 + kWordBases ? kWordBases : (end > p0 ? (int)(end - p0) : 0);
}


// 64-bit k-mer starting at base j of word w0 (w1, w2 are the following words)
__device__ __forceinline__ uint64_t kmer_at(uint32_t w0, uint32_t w1, uint32_t w2, int j, int kshift,
                                            uint32_t &h32) {
    h32 = __funnelshift_l(w1, w0, 2 * j);
    uint32_t l32 = __funnelshift_l(w2, w1, 2 * j);
    return (((uint64_t)h32 << 32) | l32) >> kshift;
}


// min over the k-mers starting in word w (positions [lo, hi)) of (k-mer ^ rlo), full 64 bits.
// GLOBAL: W is global memory (read-only path); false: any address space (the online kernel's shared memory).
template <bool GLOBAL = true>
__device__ __forceinline__ uint64_t word_min64(const uint32_t *__restrict__ W, uint64_t w, int lo, int hi,
                                               int kshift, uint64_t rlo) {
    const uint32_t w0 = GLOBAL ? __ldg(W + w) : W[w], w1 = GLOBAL ? __ldg(W + w + 1) : W[w + 1],
                   w2 = GLOBAL ? __ldg(W + w + 2) : W[w + 2];
    uint64_t best = ~0ULL;
    for (int j = lo; j < hi; ++j) {
        uint32_t h32;
        const uint64_t y = kmer_at(w0, w1, w2, j, kshift, h32) ^ rlo;
        best = y < best ? y : best;
    }
    return best;
}


} // namespace nsmh

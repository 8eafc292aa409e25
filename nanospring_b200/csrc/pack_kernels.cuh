// Device code of the 2-bit packers (see pack.cu for the layout and the reference lines it replaces).
// Kept free of runtime-API includes so that tests/cpp/pack_prefilter_host_emul.cpp can compile the
// SAME kernels for the host and check them against the oracle without a GPU.
#pragma once
#include <stdint.h>

#include "nsmh_constants.h"

namespace nsmh {

__device__ __forceinline__ uint32_t codes4(uint32_t x) {
    // 4 ASCII bytes (first base in the low byte) -> 8 bits, first base in bits 7..6
    uint32_t t = (x & 0x02020202u) | ((x >> 2) & 0x01010101u);
    return ((t << 6) | (t >> 4) | (t >> 14) | (t >> 24)) & 0xFFu;
}

// One thread per output word: 16 ASCII bytes in (one 128-bit load), 4 bytes out.
__global__ void __launch_bounds__(256)
pack_ascii_kernel(const uint8_t *__restrict__ src, uint64_t first_word, uint64_t num_bases,
                  uint32_t *__restrict__ W, int aligned16) {
    const uint64_t nwords = (num_bases + 15) / 16;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < nwords;
         t += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t x[4];
        const uint64_t b0 = t * 16;
        if (aligned16 && b0 + 16 <= num_bases) {
            uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + b0));
            x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t w = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint64_t b = b0 + q * 4 + j;
                    // bytes past the end pack as code 0
                    uint32_t c = b < num_bases ? src[b] : 0u;
                    w |= c << (8 * j);
                }
                x[q] = w;
            }
        }
        uint32_t word = (codes4(x[0]) << 24) | (codes4(x[1]) << 16) | (codes4(x[2]) << 8) | codes4(x[3]);
        W[first_word + t] = word;
    }
}

// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t read_of_base(const uint64_t *__restrict__ off, uint32_t n_reads,
                                                 uint64_t g) {
    // largest i with off[i] <= g < off[i+1]  (empty reads are skipped)
    uint32_t lo = 0, hi = n_reads;   // answer in [lo, hi)
    while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (off[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// Output word t of the reverse-complement read set: read i keeps its base range,
// rc[j] = fwd[L-1-j] ^ 1  (A0<->T1, C2<->G3).
__global__ void __launch_bounds__(256)
pack_rc_kernel(const uint64_t *__restrict__ off, uint32_t n_reads, uint64_t total_bases,
               const uint32_t *__restrict__ Wsrc, uint32_t *__restrict__ W) {
    const uint64_t nwords = (total_bases + 15) / 16;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < nwords;
         t += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t g = t * 16;
        uint32_t i = read_of_base(off, n_reads, g);
        uint64_t rb = off[i], re = off[i + 1];
        uint32_t word = 0;
#pragma unroll 1
        for (int j = 0; j < 16; ++j, ++g) {
            uint32_t code = 0;
            if (g < total_bases) {
                while (g >= re) { ++i; rb = re; re = off[i + 1]; }
                uint64_t src = rb + (re - 1 - g);
                uint32_t sw = Wsrc[src >> 4];
                int sj = (int)(src & 15);
                code = ((sw >> (30 - 2 * sj)) & 3u) ^ 1u;
            }
            word = (word << 2) | code;
        }
        W[t] = word;
    }
}

// Reads that arrive already packed the reference's way (DnaBitset: 4 bases/byte,
// first base in bits 7..6, each read byte aligned; src/dnaToBits.cpp:11-36) are
// re-laid into the continuous stream: words [w_begin, w_end) of it.  A lane owns kDnaWordsPerLane words
// 32 words apart (the warp's stores stay coalesced), finds the read of its first word by bisection and
// walks forward from there.  A word that lies inside one read - all but one in several hundred - is two
// aligned 32-bit loads and one funnel shift; a word across a read boundary goes base by base.
// `src` must be 4-byte aligned and readable 8 bytes past the last read's last byte.
constexpr int kDnaWordsPerLane = 8;

__global__ void __launch_bounds__(256)
pack_dnabitset_kernel(const uint64_t *__restrict__ off, uint32_t n_reads, uint64_t total_bases,
                      const uint8_t *__restrict__ src, const uint64_t *__restrict__ src_off,
                      uint32_t *__restrict__ W, uint64_t w_begin, uint64_t w_end) {
    const uint32_t *src32 = reinterpret_cast<const uint32_t *>(src);
    const uint64_t span = 32ULL * kDnaWordsPerLane;            // words per warp and pass
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t t0 = w_begin + warp * span; t0 < w_end; t0 += warps * span) {
        uint64_t t = t0 + lane;
        if (t >= w_end) continue;
        uint32_t i = read_of_base(off, n_reads, t * 16 < total_bases ? t * 16 : total_bases - 1);
        uint64_t rb = off[i], re = off[i + 1];
#pragma unroll 1
        for (int j = 0; j < kDnaWordsPerLane && t < w_end; ++j, t += 32) {
            uint64_t g = t * 16;
            while (g >= re && i + 1 < n_reads) { ++i; rb = re; re = off[i + 1]; }
            uint32_t word;
            if (g + 16 <= re) {
                const uint64_t lj = g - rb;
                const uint64_t a = src_off[i] + (lj >> 2);      // byte that holds base g
                const uint32_t sh = 2 * (uint32_t)(lj & 3) + 8 * (uint32_t)(a & 3);
                const uint32_t x0 = __byte_perm(src32[a >> 2], 0, 0x0123);          // big-endian: first base on top
                const uint32_t x1 = __byte_perm(src32[(a >> 2) + 1], 0, 0x0123);
                word = __funnelshift_l(x1, x0, sh);               // sh <= 30
            } else {
                word = 0;
                uint32_t ii = i;
                uint64_t b = rb, e = re;
#pragma unroll 1
                for (int q = 0; q < 16; ++q, ++g) {
                    uint32_t code = 0;
                    if (g < total_bases) {
                        while (g >= e) { ++ii; b = e; e = off[ii + 1]; }
                        const uint64_t lj = g - b;
                        code = (src[src_off[ii] + (lj >> 2)] >> (6 - 2 * (int)(lj & 3))) & 3u;
                    }
                    word = (word << 2) | code;
                }
            }
            W[t] = word;
        }
    }
}

} // namespace nsmh

// Device code of the candidate pre-filters (see prefilter.cu for what they replace in the reference).
// Kept free of runtime-API includes so that tests/cpp/pack_prefilter_host_emul.cpp can compile the
// SAME kernels for the host and check them against the oracle without a GPU.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/nsmh.h"
#include "nsmh_constants.h"

namespace nsmh {

constexpr int kRepShifts = 6;          // Consensus.cpp:412 "I choose 6 here; it is tunable"
constexpr int kRepGroupsPerWarp = 16;  // 32 words per group: a warp streams 8192 bases between queue visits

__device__ __forceinline__ uint32_t code_at(const uint32_t *__restrict__ W, uint64_t g) {
    return (__ldg(W + (g >> 4)) >> (30 - 2 * (int)(g & 15))) & 3u;
}

// bit (30 - 2q) set for positions q in [qa, qb) of a word, 0 <= qa <= qb <= 16
__device__ __forceinline__ uint32_t pos_mask(int qa, int qb) {
    if (qa >= qb) return 0u;
    const uint32_t from = qa == 0 ? 0xFFFFFFFFu : (0xFFFFFFFFu >> (2 * qa));
    const uint32_t upto = qb == 16 ? 0u : (0xFFFFFFFFu >> (2 * qb));
    return (from & ~upto) & 0x55555555u;
}

// largest i with off[i] <= g  (reads without bases are skipped over)
__device__ __forceinline__ uint32_t read_of(const uint64_t *__restrict__ off, uint32_t n_reads, uint64_t g) {
    uint32_t lo = 0, hi = n_reads;
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (off[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ void flush_counts(uint32_t *__restrict__ cnt, uint32_t read, uint32_t (&c)[kRepShifts], int lane) {
#pragma unroll
    for (int s = 0; s < kRepShifts; ++s) {
        uint32_t v = c[s];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == s && v) atomicAdd(cnt + (size_t)read * kRepShifts + s, v);
        c[s] = 0;
    }
}

__global__ void __launch_bounds__(256)
repetitive_count_kernel(const uint64_t *__restrict__ off, uint32_t n_reads, uint64_t num_words,
                        const uint32_t *__restrict__ W, uint32_t *__restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const uint64_t span = 32ull * kRepGroupsPerWarp;                 // words per warp visit
    for (uint64_t w_begin = (blockIdx.x * (uint64_t)(blockDim.x >> 5) + (threadIdx.x >> 5)) * span; w_begin < num_words;
         w_begin += warps * span) {
        uint32_t c[kRepShifts] = {0, 0, 0, 0, 0, 0};
        uint32_t cur = read_of(off, n_reads, w_begin * kWordBases);  // warp-uniform
        uint64_t cur_end = off[cur + 1];
        for (int gidx = 0; gidx < kRepGroupsPerWarp; ++gidx) {
            const uint64_t w = w_begin + 32ull * gidx + lane;
            const uint64_t g0 = (w_begin + 32ull * gidx) * kWordBases;     // first base of the group
            if (g0 >= num_words * kWordBases) break;
            const uint32_t w0 = w < num_words ? __ldg(W + w) : 0u;
            const uint32_t w1 = w < num_words ? __ldg(W + w + 1) : 0u;     // pad words follow the stream
            // fast path: all 512 positions of the group are interior positions of read `cur`
            if (g0 >= off[cur] && g0 + 32 * kWordBases + kRepShifts <= cur_end) {
#pragma unroll
                for (int s = 1; s <= kRepShifts; ++s) {
                    const uint32_t e = ~(w0 ^ __funnelshift_l(w1, w0, 2 * s));
                    c[s - 1] += __popc(e & (e >> 1) & 0x55555555u);
                }
                continue;
            }
            // slow path (a read boundary is near): leave the registers of `cur`, then every lane
            // walks the reads that own pieces of its word
            flush_counts(cnt, cur, c, lane);
            if (w < num_words) {
                const uint64_t p0 = w * kWordBases;
                uint32_t i = read_of(off, n_reads, p0);
                uint64_t q = p0;                                      // next position to account for
                while (q < p0 + kWordBases && i < n_reads) {
                    const uint64_t rb = off[i], re = off[i + 1];
                    if (re <= q) { ++i; continue; }
                    // interior positions of read i inside this word: [max(q, rb), min(p0+16, re-6))
                    const uint64_t a = q > rb ? q : rb;
                    const uint64_t lim = re - rb > kRepShifts ? re - kRepShifts : rb;
                    const uint64_t b = lim < p0 + kWordBases ? lim : p0 + kWordBases;
                    if (a < b) {
                        const uint32_t m = pos_mask((int)(a - p0), (int)(b - p0));
#pragma unroll
                        for (int s = 1; s <= kRepShifts; ++s) {
                            const uint32_t e = ~(w0 ^ __funnelshift_l(w1, w0, 2 * s));
                            const uint32_t v = __popc(e & (e >> 1) & m);
                            if (v) atomicAdd(cnt + (size_t)i * kRepShifts + (s - 1), v);
                        }
                    }
                    q = re < p0 + kWordBases ? re : p0 + kWordBases;
                    if (re <= p0 + kWordBases) ++i;
                }
            }
            // the group may have moved the warp into a later read
            const uint64_t next = g0 + 32 * kWordBases;
            if (next >= cur_end) {
                cur = read_of(off, n_reads, next < num_words * kWordBases ? next : num_words * kWordBases - 1);
                cur_end = off[cur + 1];
            }
        }
        flush_counts(cnt, cur, c, lane);
    }
}

// flags[i]: bit 0 = repetitive (Consensus.cpp:405-424), bit 1 = shorter than 32 bases (Consensus.cpp:213)
__global__ void __launch_bounds__(256)
repetitive_flag_kernel(const uint64_t *__restrict__ off, uint32_t n_reads, const uint32_t *__restrict__ W,
                       const uint32_t *__restrict__ cnt, uint8_t *__restrict__ flags) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_reads; i += gridDim.x * blockDim.x) {
        const uint64_t rb = off[i], L = off[i + 1] - rb;
        uint32_t c[kRepShifts];
#pragma unroll
        for (int s = 0; s < kRepShifts; ++s) c[s] = cnt[(size_t)i * kRepShifts + s];
        // positions whose partner (j + s) % L may wrap: j in [L-6, L), or all of a read of <= 6 bases
        for (uint64_t j = L > kRepShifts ? L - kRepShifts : 0; j < L; ++j) {
            const uint32_t a = code_at(W, rb + j);
#pragma unroll
            for (int s = 1; s <= kRepShifts; ++s)
                if (a == code_at(W, rb + (j + s) % L)) ++c[s - 1];
        }
        bool rep = false;
#pragma unroll
        for (int s = 0; s < kRepShifts; ++s) rep = rep || ((double)c[s] > 0.7 * (double)L);   // Consensus.cpp:420
        flags[i] = (uint8_t)((rep ? NSMH_FLAG_REPETITIVE : 0) | (L < 32 ? NSMH_FLAG_SHORT : 0));
    }
}

// ---- dropping flagged candidates from the bulk CSR -------------------------------------------
// warp per query row: count the surviving ids / write them in order
__global__ void __launch_bounds__(256)
csr_keep_count_kernel(const uint64_t *__restrict__ off, const uint32_t *__restrict__ ids, uint32_t nq,
                      const uint8_t *__restrict__ flags, uint32_t n_flags, uint32_t drop, uint32_t *__restrict__ keep) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < nq; q += warps) {
        uint32_t c = 0;
        for (uint64_t t = off[q] + lane; t < off[q + 1]; t += 32) {
            const uint32_t id = ids[t];
            c += (id < n_flags && (flags[id] & drop)) ? 0u : 1u;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) keep[q] = c;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) keep[nq] = 0;
}

__global__ void __launch_bounds__(256)
csr_keep_write_kernel(const uint64_t *__restrict__ off, const uint32_t *__restrict__ ids, uint32_t nq,
                      const uint8_t *__restrict__ flags, uint32_t n_flags, uint32_t drop,
                      const uint64_t *__restrict__ new_off, uint32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < nq; q += warps) {
        uint64_t dst = new_off[q];
        const uint64_t b = off[q], e = off[q + 1];
        for (uint64_t t0 = b; t0 < e; t0 += 32) {         // whole warp iterates: ballot keeps the order
            const uint64_t t = t0 + lane;
            uint32_t id = 0;
            bool keep = false;
            if (t < e) {
                id = ids[t];
                keep = !(id < n_flags && (flags[id] & drop));
            }
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (keep) out[dst + __popc(m & ((1u << lane) - 1))] = id;
            dst += __popc(m);
        }
    }
}

} // namespace nsmh

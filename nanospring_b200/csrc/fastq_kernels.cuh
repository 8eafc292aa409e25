// Device code of the FASTQ ingest (see fastq.cu for the pipeline).  Kept free of runtime-API
// includes so that tests/cpp/fastq_host_emul.cpp can compile the SAME kernels for the host
// (lock-step warp emulation) and check them against the oracle without a GPU.
#pragma once
#include <stdint.h>

namespace nsmh {

constexpr int kFqTileBytes = 512;      // bytes per warp step: 32 lanes x 16
constexpr int kFqPackIters = 16;       // words per lane and chunk in fastq_pack_kernel

__device__ __forceinline__ uint32_t fq_codes4(uint32_t x) {
    // 4 ASCII bytes (first base in the low byte) -> 8 bits, first base in bits 7..6 (pack.cu codes4)
    uint32_t t = (x & 0x02020202u) | ((x >> 2) & 0x01010101u);
    return ((t << 6) | (t >> 4) | (t >> 14) | (t >> 24)) & 0xFFu;
}

// 16 text bytes starting at `pos` (multiple of 16), bytes past the end read as 0
__device__ __forceinline__ uint4 fq_load16(const uint8_t *__restrict__ text, uint64_t bytes, uint64_t pos,
                                           int aligned16) {
    if (aligned16 && pos + 16 <= bytes) return __ldg(reinterpret_cast<const uint4 *>(text + pos));
    uint32_t x[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint64_t b = pos + j;
        const uint32_t c = b < bytes ? text[b] : 0u;
        x[j >> 2] |= c << (8 * (j & 3));
    }
    return make_uint4(x[0], x[1], x[2], x[3]);
}

// bit j set <=> byte j of the 16 is '\n'
__device__ __forceinline__ uint32_t fq_newline_mask(uint4 v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t e = __vcmpeq4(w[q], 0x0A0A0A0Au);       // 0xFF in every matching byte
        const uint32_t bits = (e & 0x01u) | ((e >> 7) & 0x02u) | ((e >> 14) & 0x04u) | ((e >> 21) & 0x08u);
        m |= bits << (4 * q);
    }
    return m;
}

__global__ void __launch_bounds__(256)
fastq_count_newlines_kernel(const uint8_t *__restrict__ text, uint64_t bytes, int aligned16, uint64_t ntiles,
                            uint32_t *__restrict__ tile_cnt) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < ntiles; t += warps) {
        const uint64_t pos = t * kFqTileBytes + (uint64_t)lane * 16;
        uint32_t c = 0;
        if (pos < bytes) c = __popc(fq_newline_mask(fq_load16(text, bytes, pos, aligned16)));
        c = __reduce_add_sync(0xFFFFFFFFu, c);
        if (lane == 0) tile_cnt[t] = c;
    }
}

__global__ void __launch_bounds__(256)
fastq_write_newlines_kernel(const uint8_t *__restrict__ text, uint64_t bytes, int aligned16, uint64_t ntiles,
                            const uint64_t *__restrict__ tile_base, uint64_t *__restrict__ nl) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < ntiles; t += warps) {
        const uint64_t base = tile_base[t];
        if (tile_base[t + 1] == base) continue;               // warp-uniform: no newline in this tile
        const uint64_t pos = t * kFqTileBytes + (uint64_t)lane * 16;
        uint32_t m = 0;
        if (pos < bytes) m = fq_newline_mask(fq_load16(text, bytes, pos, aligned16));
        const uint32_t c = __popc(m);
        uint32_t incl = c;                                     // inclusive warp scan of the lane counts
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += v;
        }
        uint64_t o = base + (incl - c);
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            nl[o++] = pos + j;
        }
    }
}

// Record i = lines 4i .. 4i+3; line j (j >= 1) starts at nl[j-1] + 1 and ends at nl[j] (or at `bytes`
// for a last line without '\n').  num_lines = newlines + (text does not end in '\n').
__global__ void __launch_bounds__(256)
fastq_read_table_kernel(const uint64_t *__restrict__ nl, uint64_t newlines, uint64_t num_lines, uint64_t bytes,
                        uint32_t num_reads, uint64_t *__restrict__ src_start, uint32_t *__restrict__ len32,
                        unsigned long long *__restrict__ too_long) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > num_reads) return;
    if (i == num_reads) {                                      // sentinel for the exclusive sum
        len32[i] = 0;
        return;
    }
    const uint64_t seq_line = 4 * i + 1;
    uint64_t b, e;
    if (seq_line < num_lines) {
        b = nl[seq_line - 1] + 1;
        e = seq_line < newlines ? nl[seq_line] : bytes;
    } else if (num_lines > newlines) {
        // the text ends inside this record's header line: the reference's second getline fails
        // without clearing its string, so the header bytes are what gets stored (ReadData.cpp:177-191)
        b = i ? nl[4 * i - 1] + 1 : 0;
        e = bytes;
    } else {
        b = e = bytes;                                         // header line complete, nothing after it
    }
    const uint64_t len = e - b;
    if (len > 0xFFFFFFFFull) atomicAdd(too_long, 1ull);
    src_start[i] = b;
    len32[i] = (uint32_t)len;
}

__device__ __forceinline__ uint32_t fq_read_of_base(const uint64_t *__restrict__ off, uint32_t n_reads, uint64_t g) {
    uint32_t lo = 0, hi = n_reads;      // largest i with off[i] <= g (empty reads before it are skipped)
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (off[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// A warp owns chunks of 32 * kFqPackIters consecutive output words; lane l packs words
// first + l, first + 32 + l, ...  The read of a lane's first word comes from one binary search,
// later words walk forward (base offsets only grow).  A word that lies inside one read takes its
// 16 bytes from 5 aligned 32-bit loads (consecutive lanes read consecutive 16-byte pieces of the
// same line); words that straddle reads, and the text's last bytes, go base by base.
__global__ void __launch_bounds__(256)
fastq_pack_kernel(const uint8_t *__restrict__ text, uint64_t safe_bytes, const uint64_t *__restrict__ off,
                  const uint64_t *__restrict__ src_start, uint32_t n_reads, uint64_t total_bases,
                  uint32_t *__restrict__ W) {
    const uint64_t nwords = (total_bases + 15) / 16;
    const uint64_t chunk_words = 32ull * kFqPackIters;
    const uint64_t nchunks = (nwords + chunk_words - 1) / chunk_words;
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t ch = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ch < nchunks; ch += warps) {
        uint64_t t = ch * chunk_words + lane;
        if (t >= nwords) continue;
        uint32_t i = fq_read_of_base(off, n_reads, t * 16);
        uint64_t rb = off[i], re = off[i + 1], sb = src_start[i];
        for (int it = 0; it < kFqPackIters && t < nwords; ++it, t += 32) {
            uint64_t g = t * 16;
            while (g >= re && i + 1 < n_reads) {
                ++i;
                rb = re;
                re = off[i + 1];
                sb = src_start[i];
            }
            uint32_t word;
            const uint64_t p = sb + (g - rb);
            const uint64_t pa = p & ~3ull;
            if (g + 16 <= re && pa + 20 <= safe_bytes) {
                const uint32_t *a = reinterpret_cast<const uint32_t *>(text + pa);
                const uint32_t w0 = __ldg(a), w1 = __ldg(a + 1), w2 = __ldg(a + 2), w3 = __ldg(a + 3), w4 = __ldg(a + 4);
                const uint32_t sh = (uint32_t)(p & 3) * 8;
                const uint32_t x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh);
                const uint32_t x2 = __funnelshift_r(w2, w3, sh), x3 = __funnelshift_r(w3, w4, sh);
                word = (fq_codes4(x0) << 24) | (fq_codes4(x1) << 16) | (fq_codes4(x2) << 8) | fq_codes4(x3);
            } else {
                word = 0;
#pragma unroll 1
                for (int j = 0; j < 16; ++j, ++g) {
                    uint32_t code = 0;
                    if (g < total_bases) {
                        while (g >= re) {
                            ++i;
                            rb = re;
                            re = off[i + 1];
                            sb = src_start[i];
                        }
                        const uint32_t c = text[sb + (g - rb)];
                        code = (c & 2u) | ((c & 4u) >> 2);
                    }
                    word = (word << 2) | code;
                }
            }
            W[t] = word;
        }
    }
}

// out[j] = "ATCG"[code of global base b0 + j], j < nb; a thread per 16 output bytes
__global__ void __launch_bounds__(256)
unpack_ascii_kernel(const uint32_t *__restrict__ W, uint64_t b0, uint64_t nb, uint8_t *__restrict__ out) {
    const uint64_t groups = (nb + 15) / 16;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < groups;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t b = b0 + t * 16;
        const uint32_t sh = (uint32_t)(b & 15) * 2;
        // 16 codes starting at base b, first base most significant (W has kPackPadWords zero words behind it)
        const uint32_t lo = W[(b >> 4) + 1], hi = W[b >> 4];
        const uint32_t v = __funnelshift_l(lo, hi, sh);
        uint32_t x[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t code = (v >> (30 - 2 * (4 * q + j))) & 3u;
                w |= ((0x47435441u >> (8 * code)) & 0xFFu) << (8 * j);     // 'A' 'T' 'C' 'G'
            }
            x[q] = w;
        }
        if (t * 16 + 16 <= nb) {
            *reinterpret_cast<uint4 *>(out + t * 16) = make_uint4(x[0], x[1], x[2], x[3]);
        } else {
            for (uint64_t j = t * 16; j < nb; ++j) out[j] = (uint8_t)(x[(j >> 2) & 3] >> (8 * (j & 3)));
        }
    }
}

} // namespace nsmh

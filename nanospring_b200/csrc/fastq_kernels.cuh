// Device code of the FASTQ ingest (see fastq.cu for the pipeline).  Kept free of runtime-API
// includes so that tests/cpp/fastq_host_emul.cpp can compile the SAME kernels for the host
// (lock-step warp emulation) and check them against the oracle without a GPU.
#pragma once
#include <stdint.h>

namespace nsmh {

constexpr int kFqTileBytes = 512;      // bytes per warp step: 32 lanes x 16
constexpr int kFqCountTiles = 4;       // tiles per warp step in fastq_count_newlines_kernel
constexpr int kFqPackIters = 32;       // words per lane and chunk in fastq_pack_kernel (default of its `iters`)
constexpr int kFqPackUnroll = 4;       // ... of which this many are gathered together (divides iters)

// 4 ASCII bytes (first base in the low byte) -> 2-bit codes (c & 2) | ((c & 4) >> 2) in the TOP byte,
// first base in bits 31..30.  The four code fields of t sit at bits 0, 8, 16, 24; one multiply by
// 2^30 + 2^20 + 2^10 + 1 moves them to bits 30, 28, 26, 24 (all partial products land on distinct
// bit positions, so nothing carries; what would land above bit 31 drops out).
__device__ __forceinline__ uint32_t fq_codes4_top(uint32_t x) {
    const uint32_t t = (x & 0x02020202u) | ((x >> 2) & 0x01010101u);
    return t * 0x40100401u;
}

// four times that: 16 bases -> one packed word, first base most significant
__device__ __forceinline__ uint32_t fq_pack16(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) {
    const uint32_t hi = __byte_perm(fq_codes4_top(x0), fq_codes4_top(x1), 0x0037);   // [.. .. c(x0) c(x1)]
    const uint32_t lo = __byte_perm(fq_codes4_top(x2), fq_codes4_top(x3), 0x0037);
    return __byte_perm(lo, hi, 0x5410);
}

// 0x80 in every byte of x that equals '\n', 0 elsewhere (exact for any byte values)
__device__ __forceinline__ uint32_t fq_newline_bytes(uint32_t x) {
    const uint32_t y = x ^ 0x0A0A0A0Au;                        // zero bytes <=> newlines
    return ~(((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y | 0x7F7F7F7Fu);
}

// number of '\n' among the 16 bytes
__device__ __forceinline__ uint32_t fq_newline_count(uint4 v) {
    const uint32_t z = fq_newline_bytes(v.x) | (fq_newline_bytes(v.y) >> 1) | (fq_newline_bytes(v.z) >> 2) |
                       (fq_newline_bytes(v.w) >> 3);
    return (uint32_t)__popc(z);
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
        // flags at bits 0, 8, 16, 24 -> bits 21..24 of the product (distinct positions, no carries)
        const uint32_t bits = (((fq_newline_bytes(w[q]) >> 7) * 0x00204081u) >> 21) & 0xFu;
        m |= bits << (4 * q);
    }
    return m;
}

// A warp counts kFqCountTiles consecutive tiles per step: the loads of all of them are issued before
// the first byte compare, so a lane keeps kFqCountTiles x 16 bytes in flight (one tile per step left
// the kernel waiting on DRAM latency at half the copy bandwidth).
__global__ void __launch_bounds__(256)
fastq_count_newlines_kernel(const uint8_t *__restrict__ text, uint64_t bytes, int aligned16, uint64_t ntiles,
                            uint32_t *__restrict__ tile_cnt) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t w0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (uint64_t t0 = w0 * kFqCountTiles; t0 < ntiles; t0 += warps * kFqCountTiles) {
        uint4 v[kFqCountTiles];
        if (aligned16 && (t0 + kFqCountTiles) * kFqTileBytes <= bytes) {       // warp-uniform: the common case
            const uint4 *src = reinterpret_cast<const uint4 *>(text + t0 * kFqTileBytes) + lane;
#pragma unroll
            for (int u = 0; u < kFqCountTiles; ++u) v[u] = __ldg(src + u * (kFqTileBytes / 16));
        } else {
#pragma unroll 1
            for (int u = 0; u < kFqCountTiles; ++u) {
                const uint64_t pos = (t0 + u) * kFqTileBytes + (uint64_t)lane * 16;
                const uint4 x = pos < bytes ? fq_load16(text, bytes, pos, aligned16) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int q = 0; q < kFqCountTiles; ++q)
                    if (q == u) v[q] = x;
            }
        }
        uint32_t mine = 0;
#pragma unroll
        for (int u = 0; u < kFqCountTiles; ++u) {
            const uint32_t c = __reduce_add_sync(0xFFFFFFFFu, fq_newline_count(v[u]));
            if (lane == u) mine = c;
        }
        if (lane < kFqCountTiles && t0 + lane < ntiles) tile_cnt[t0 + lane] = mine;
    }
}

// A warp looks at 32 tiles per step (lane l reads the bases of tile t0 + l, coalesced) and only
// walks through the ones that hold a newline - on nanopore FASTQ about one tile in ten; their text
// is read a second time, the rest costs 8 bytes of tile_base per tile.
__global__ void __launch_bounds__(256)
fastq_write_newlines_kernel(const uint8_t *__restrict__ text, uint64_t bytes, int aligned16, uint64_t ntiles,
                            const uint64_t *__restrict__ tile_base, uint64_t *__restrict__ nl) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t w0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (uint64_t t0 = w0 * 32; t0 < ntiles; t0 += warps * 32) {
        uint64_t my_base = 0, my_next = 0;
        if (t0 + lane < ntiles) {
            my_base = tile_base[t0 + lane];
            my_next = tile_base[t0 + lane + 1];
        }
        uint32_t has = __ballot_sync(0xFFFFFFFFu, my_next != my_base);
        while (has) {                                          // warp-uniform
            const int src = __ffs(has) - 1;
            has &= has - 1;
            const uint64_t base = __shfl_sync(0xFFFFFFFFu, my_base, src);
            const uint64_t pos = (t0 + src) * kFqTileBytes + (uint64_t)lane * 16;
            uint32_t m = 0;
            if (pos < bytes) m = fq_newline_mask(fq_load16(text, bytes, pos, aligned16));
            const uint32_t c = __popc(m);
            uint32_t incl = c;                                 // inclusive warp scan of the lane counts
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

// A warp owns chunks of 32 * iters consecutive output words; lane l packs words
// first + l, first + 32 + l, ...  The read of a lane's first word comes from one binary search,
// later words walk forward (base offsets only grow).  A word that lies inside one read takes its
// 16 bytes from 5 aligned 32-bit loads (consecutive lanes read consecutive 16-byte pieces of the
// same line); words that straddle reads, and the text's last bytes, go base by base.  Words are
// handled kFqPackUnroll at a time: first the read cursor is advanced for each of them (rarely more
// than a compare), then all their loads are issued, then the codes are extracted - 80 bytes in
// flight per lane instead of 20.
// (Measured on B200: gathering the two enclosing 16-byte blocks with 128-bit loads instead - 8 L1
// wavefronts per warp-word instead of 20 - is not faster, profiles/r1_ingest_sweep_s8.jsonl; the kernel
// is bound by instruction issue, not by the load path.)
__global__ void __launch_bounds__(256)
fastq_pack_kernel(const uint8_t *__restrict__ text, uint64_t safe_bytes, const uint64_t *__restrict__ off,
                  const uint64_t *__restrict__ src_start, uint32_t n_reads, uint64_t total_bases,
                  uint32_t *__restrict__ W, uint32_t iters) {
    const uint64_t nwords = (total_bases + 15) / 16;
    const uint64_t chunk_words = 32ull * iters;
    const uint64_t nchunks = (nwords + chunk_words - 1) / chunk_words;
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t tb = reinterpret_cast<uint64_t>(text);
    for (uint64_t ch = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ch < nchunks; ch += warps) {
        uint64_t t = ch * chunk_words + lane;
        if (t >= nwords) continue;
        uint32_t i = fq_read_of_base(off, n_reads, t * 16);
        uint64_t rb = off[i], re = off[i + 1], sb = src_start[i];
        for (uint32_t it = 0; it < iters && t < nwords; it += kFqPackUnroll, t += 32ull * kFqPackUnroll) {
            uint64_t a[kFqPackUnroll];                         // aligned address of the first load
            uint32_t sh[kFqPackUnroll], word[kFqPackUnroll];   // byte offset of the word behind it
            bool fast[kFqPackUnroll];
#pragma unroll
            for (int u = 0; u < kFqPackUnroll; ++u) {
                const uint64_t tu = t + 32ull * u;
                a[u] = 0;
                sh[u] = 0;
                word[u] = 0;
                fast[u] = false;
                if (tu >= nwords) continue;
                uint64_t g = tu * 16;
                while (g >= re && i + 1 < n_reads) {
                    ++i;
                    rb = re;
                    re = off[i + 1];
                    sb = src_start[i];
                }
                const uint64_t addr = tb + sb + (g - rb);
                const uint64_t base = addr & ~3ull;
                if (g + 16 <= re && safe_bytes && base + 20 <= tb + safe_bytes) {
                    fast[u] = true;
                    a[u] = base;
                    sh[u] = (uint32_t)(addr - base);
                } else {
                    uint32_t w = 0;
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
                        w = (w << 2) | code;
                    }
                    word[u] = w;
                }
            }
            uint32_t w[kFqPackUnroll][5];
#pragma unroll
            for (int u = 0; u < kFqPackUnroll; ++u)
#pragma unroll
                for (int q = 0; q < 5; ++q) w[u][q] = fast[u] ? __ldg(reinterpret_cast<const uint32_t *>(a[u]) + q) : 0u;
#pragma unroll
            for (int u = 0; u < kFqPackUnroll; ++u) {
                const uint64_t tu = t + 32ull * u;
                if (tu >= nwords) continue;
                if (fast[u]) {
                    const uint32_t e0 = w[u][0], e1 = w[u][1], e2 = w[u][2], e3 = w[u][3], e4 = w[u][4];
                    const uint32_t bs = (sh[u] & 3) * 8;
                    const uint32_t x0 = __funnelshift_r(e0, e1, bs), x1 = __funnelshift_r(e1, e2, bs);
                    const uint32_t x2 = __funnelshift_r(e2, e3, bs), x3 = __funnelshift_r(e3, e4, bs);
                    word[u] = fq_pack16(x0, x1, x2, x3);
                }
                W[tu] = word[u];
            }
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

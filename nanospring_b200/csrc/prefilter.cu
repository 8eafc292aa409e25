// Candidate pre-filters of the consensus builder, computed on the device from the packed reads
// (SURVEY section 8(f), row N3).  Replaces
//   Consensus::checkRepetitive  (src/Consensus.cpp:405-424)  one getRead + 6 shifted self-compares per read
//   Consensus::initialize       (src/Consensus.cpp:426-442)  isRepetitive[] for all reads (omp parallel)
//   the length gate             (src/Consensus.cpp:213)      `readStr1.size() < 32` -> skip the candidate
// and applies them to the bulk candidate CSR before it leaves the device
// (Consensus.cpp:204-216: `if (isRepetitive[r]) continue; ... if (size < 32) continue;`).
//
// checkRepetitive(read): for shift i in 1..6, c_i = #{ j in [0,L) : read[j] == read[(j+i) % L] };
// repetitive iff some c_i > 0.7 * (double)L.  The strings the reference compares come out of its
// 2-bit store (ReadData::getRead -> DnaBitset::to_string, dnaToBits.cpp:81-98), so comparing bases
// is comparing 2-bit codes, which is what the packed stream holds.
//
// Two kernels:
//  * repetitive_count_kernel  streams the packed words once (0.25 B/base, HBM bound).  A position
//    j is "interior" when j + 6 < L: its six partners are the next six bases of the stream, so a
//    word's 16 positions are compared with one funnel shift + xor + popcount per shift.  Counts
//    are kept in registers while a warp stays inside one read and added to cnt[read][6] with one
//    atomic per shift when it leaves it.
//  * repetitive_flag_kernel   one thread per read: the (at most 6) positions whose partners wrap
//    around the end of the read are done with the reference's modulo, then the threshold test in
//    double precision exactly as written in the reference.
#include <algorithm>

#include "nsmh_internal.cuh"

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

int compute_read_flags(nsmh_ctx *c) {
    cudaStream_t s = c->stream;
    const ReadSet &rs = c->reads;
    NSMH_TRY(c->read_flags.ensure(std::max<size_t>(rs.num_reads, 1), s));
    if (rs.num_reads == 0) return NSMH_OK;
    DevBuf cnt;
    NSMH_TRY(cnt.ensure((size_t)rs.num_reads * kRepShifts * sizeof(uint32_t), s));
    NSMH_CK(cudaMemsetAsync(cnt.p, 0, (size_t)rs.num_reads * kRepShifts * sizeof(uint32_t), s));
    if (rs.num_words) {
        const uint64_t visits = (rs.num_words + 32ull * kRepGroupsPerWarp - 1) / (32ull * kRepGroupsPerWarp);
        const int blocks = (int)std::min<uint64_t>((visits + 7) / 8, (uint64_t)c->num_sms * 8);
        repetitive_count_kernel<<<blocks, 256, 0, s>>>(rs.d_offsets(), rs.num_reads, rs.num_words,
                                                       rs.packed.as<uint32_t>(), cnt.as<uint32_t>());
        ++c->launches;
        NSMH_CK(cudaGetLastError());
    }
    const int blocks = (int)std::min<uint64_t>(((uint64_t)rs.num_reads + 255) / 256, (uint64_t)c->num_sms * 8);
    repetitive_flag_kernel<<<blocks, 256, 0, s>>>(rs.d_offsets(), rs.num_reads, rs.packed.as<uint32_t>(),
                                                  cnt.as<uint32_t>(), c->read_flags.as<uint8_t>());
    ++c->launches;
    NSMH_CK(cudaGetLastError());
    cnt.release(s);
    return NSMH_OK;
}

// Compacts ws.out_off / ws.out_ids in place (through ws.tmp_ids), dropping ids whose flags
// intersect `drop`.  flags index = id - id_base (ids outside the local reads are kept).
int drop_flagged_candidates(nsmh_ctx *c, QueryWs &ws, uint32_t drop, cudaStream_t s) {
    const uint32_t nq = ws.last_nq;
    if (nq == 0 || ws.last_total == 0) return NSMH_OK;
    NSMH_TRY(ws.qcount.ensure(((size_t)nq + 1) * sizeof(uint32_t), s));
    NSMH_TRY(ws.hstart.ensure(((size_t)nq + 1) * sizeof(uint64_t), s));
    NSMH_TRY(ws.tmp_ids.ensure(ws.last_total * sizeof(uint32_t), s));
    const int blocks = (int)std::min<uint64_t>(((uint64_t)nq + 7) / 8, (uint64_t)c->num_sms * 8);
    csr_keep_count_kernel<<<blocks, 256, 0, s>>>(ws.out_off.as<uint64_t>(), ws.out_ids.as<uint32_t>(), nq,
                                                 c->read_flags.as<uint8_t>(), c->reads.num_reads, drop,
                                                 ws.qcount.as<uint32_t>());
    NSMH_CK(cudaGetLastError());
    size_t tmp_bytes = 0;
    NSMH_CK(cub_exclusive_sum_u32_to_u64(nullptr, tmp_bytes, ws.qcount.as<uint32_t>(), ws.hstart.as<uint64_t>(), (size_t)nq + 1, s));
    NSMH_TRY(ws.cub_tmp.ensure(tmp_bytes, s));
    NSMH_CK(cub_exclusive_sum_u32_to_u64(ws.cub_tmp.p, tmp_bytes, ws.qcount.as<uint32_t>(), ws.hstart.as<uint64_t>(), (size_t)nq + 1, s));
    csr_keep_write_kernel<<<blocks, 256, 0, s>>>(ws.out_off.as<uint64_t>(), ws.out_ids.as<uint32_t>(), nq,
                                                 c->read_flags.as<uint8_t>(), c->reads.num_reads, drop,
                                                 ws.hstart.as<uint64_t>(), ws.tmp_ids.as<uint32_t>());
    NSMH_CK(cudaGetLastError());
    ws.launches += 4;
    uint64_t total = 0;
    NSMH_CK(cudaMemcpyAsync(&total, ws.hstart.as<uint64_t>() + nq, sizeof total, cudaMemcpyDeviceToHost, s));
    NSMH_CK(cudaMemcpyAsync(ws.out_off.p, ws.hstart.p, ((size_t)nq + 1) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
    NSMH_CK(cudaStreamSynchronize(s));
    if (total) NSMH_CK(cudaMemcpyAsync(ws.out_ids.p, ws.tmp_ids.p, total * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    NSMH_CK(cudaStreamSynchronize(s));
    ws.last_total = total;
    return NSMH_OK;
}

} // namespace nsmh

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
#include "prefilter_kernels.cuh"

namespace nsmh {

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
// intersect `drop`.  flags index = candidate id: nsmh_query_all_drop only accepts single-GPU bulk results,
// where ids are ids of the loaded reads (ids >= num_reads cannot occur there and would be kept).
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

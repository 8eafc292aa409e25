// FASTQ text -> packed reads on the device.  Replaces the reference's host loader
//   ReadData::loadFromFastqFile_lowmem   (src/ReadData.cpp:156-221; the CLI fixes low_mem = true,
//                                         src/main.cpp:40): getline x4 per record, the second line is
//                                         the read, DnaBitset::load_from_string + to_file per read
//   ReadData::getRead                    (src/ReadData.cpp:225-235): mutex + seekg + unpack per call
// The inflated text is copied to the device once; everything else happens there:
//
//   fastq_count_newlines_kernel   a warp per 4 x 512-byte tiles, 128-bit loads, byte compares (HBM: text once)
//   (CUB exclusive sum over the tile counts)
//   fastq_write_newlines_kernel   ordered byte positions of all '\n' (only tiles that have one are re-read)
//   fastq_read_table_kernel       a thread per record: byte range of its second line = the read
//   (CUB exclusive sum over the read lengths -> base offsets)
//   fastq_pack_kernel             a thread per packed u32 word: 16 read bytes gathered from the text
//                                 (5 aligned 32-bit loads + funnel shifts), 2-bit codes out
//   unpack_ascii_kernel           "ATCG"[code] for nsmh_get_reads_ascii (what getRead returns)
//
// Record semantics are those of the reference loop, including its malformed-input corners
// (oracle/minhash_oracle.c orc_fastq_index restates them; tests/golden/fastq_golden.npz holds the
// outputs of the unmodified src/ReadData.cpp): lines are split at '\n' only, a final line without
// '\n' counts, a trailing '\n' does not open a new line, record i is lines 4i..4i+3 whatever they
// start with, a missing second line gives an empty read - except when the text ends inside the
// header line itself (no '\n' after it), where the reference stores the header bytes as the read.
#include <algorithm>
#include <cstdlib>

#include "nsmh_internal.cuh"
#include "fastq_kernels.cuh"

namespace nsmh {

static int grid_warps(uint64_t warps_wanted, int num_sms) {
    const uint64_t blocks = (warps_wanted + 7) / 8;            // 8 warps per 256-thread block
    const uint64_t cap = (uint64_t)num_sms * 8;                // 8 resident blocks of 256 threads per SM
    return (int)std::max<uint64_t>(1, std::min(blocks, cap));
}

// d_text: device pointer to the whole text; safe_bytes >= bytes: how far the allocation may be
// read.  Fills c->reads (offsets + packed stream).  Synchronises the stream.
int parse_fastq_device(nsmh_ctx *c, const uint8_t *d_text, uint64_t bytes, uint64_t safe_bytes, int last_byte) {
    cudaStream_t s = c->stream;
    ReadSet &rs = c->reads;
    if (rs.external_offsets) { rs.offsets.p = nullptr; rs.offsets.cap = 0; rs.external_offsets = false; }
    const int aligned16 = (reinterpret_cast<uintptr_t>(d_text) & 15) == 0;
    if (reinterpret_cast<uintptr_t>(d_text) & 3) safe_bytes = 0;      // 32-bit gathers need a 4-aligned base
    const uint64_t ntiles = (bytes + kFqTileBytes - 1) / kFqTileBytes;
    DevBuf tile_cnt, tile_base, nl, src, len32, tmp, flag;
    int rc = NSMH_OK;
    cudaError_t e = cudaSuccess;
#define FQ_CK(call) do { if (!rc && (e = (call)) != cudaSuccess) rc = cuda_fail(e, #call, __FILE__, __LINE__); } while (0)
#define FQ_TRY(expr) do { if (!rc) rc = (expr); } while (0)
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    for (auto &x : ev) FQ_CK(cudaEventCreate(&x));
    FQ_CK(cudaEventRecord(ev[0], s));
    FQ_TRY(tile_cnt.ensure((ntiles + 1) * sizeof(uint32_t), s));
    FQ_TRY(tile_base.ensure((ntiles + 1) * sizeof(uint64_t), s));
    FQ_TRY(flag.ensure(sizeof(unsigned long long), s));
    FQ_CK(cudaMemsetAsync(flag.p, 0, sizeof(unsigned long long), s));
    if (!rc) FQ_CK(cudaMemsetAsync(tile_cnt.as<uint32_t>() + ntiles, 0, sizeof(uint32_t), s));
    if (!rc && ntiles) {
        fastq_count_newlines_kernel<<<grid_warps((ntiles + kFqCountTiles - 1) / kFqCountTiles, c->num_sms), 256, 0, s>>>(d_text, bytes, aligned16, ntiles,
                                                                                  tile_cnt.as<uint32_t>());
        ++c->launches;
        FQ_CK(cudaGetLastError());
    }
    size_t tmp_bytes = 0;
    FQ_CK(cub_exclusive_sum_u32_to_u64(nullptr, tmp_bytes, tile_cnt.as<uint32_t>(), tile_base.as<uint64_t>(), ntiles + 1, s));
    FQ_TRY(tmp.ensure(tmp_bytes, s));
    FQ_CK(cub_exclusive_sum_u32_to_u64(tmp.p, tmp_bytes, tile_cnt.as<uint32_t>(), tile_base.as<uint64_t>(), ntiles + 1, s));
    c->launches += 2;
    uint64_t newlines = 0;
    if (!rc) FQ_CK(cudaMemcpyAsync(&newlines, tile_base.as<uint64_t>() + ntiles, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    FQ_CK(cudaStreamSynchronize(s));
    uint64_t num_lines = 0, num_reads64 = 0;
    if (!rc) {
        num_lines = newlines + ((bytes > 0 && last_byte != '\n') ? 1 : 0);
        num_reads64 = (num_lines + 3) / 4;
        // ReadData.cpp:184-186: "Too many reads for read_t type to handle."
        if (num_reads64 >= 0xFFFFFFFFull) rc = fail(NSMH_EINVAL, "load_fastq: too many reads for 32-bit read ids");
    }
    const uint32_t num_reads = (uint32_t)num_reads64;
    FQ_TRY(nl.ensure(std::max<uint64_t>(newlines, 1) * sizeof(uint64_t), s));
    if (!rc && newlines) {
        fastq_write_newlines_kernel<<<grid_warps((ntiles + 31) / 32, c->num_sms), 256, 0, s>>>(d_text, bytes, aligned16, ntiles,
                                                                                  tile_base.as<uint64_t>(), nl.as<uint64_t>());
        ++c->launches;
        FQ_CK(cudaGetLastError());
    }
    FQ_TRY(src.ensure(std::max<uint64_t>(num_reads, 1) * sizeof(uint64_t), s));
    FQ_TRY(len32.ensure(((uint64_t)num_reads + 1) * sizeof(uint32_t), s));
    FQ_TRY(rs.offsets.ensure(((uint64_t)num_reads + 1) * sizeof(uint64_t), s));
    if (!rc) {
        const uint64_t threads = (uint64_t)num_reads + 1;
        fastq_read_table_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(
            nl.as<uint64_t>(), newlines, num_lines, bytes, num_reads, src.as<uint64_t>(), len32.as<uint32_t>(),
            flag.as<unsigned long long>());
        ++c->launches;
        FQ_CK(cudaGetLastError());
    }
    tmp_bytes = 0;
    FQ_CK(cub_exclusive_sum_u32_to_u64(nullptr, tmp_bytes, len32.as<uint32_t>(), rs.offsets.as<uint64_t>(), (size_t)num_reads + 1, s));
    FQ_TRY(tmp.ensure(tmp_bytes, s));
    FQ_CK(cub_exclusive_sum_u32_to_u64(tmp.p, tmp_bytes, len32.as<uint32_t>(), rs.offsets.as<uint64_t>(), (size_t)num_reads + 1, s));
    c->launches += 2;
    uint64_t total_bases = 0;
    unsigned long long too_long = 0;
    if (!rc) FQ_CK(cudaMemcpyAsync(&total_bases, rs.offsets.as<uint64_t>() + num_reads, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    if (!rc) FQ_CK(cudaMemcpyAsync(&too_long, flag.p, sizeof too_long, cudaMemcpyDeviceToHost, s));
    FQ_CK(cudaStreamSynchronize(s));
    if (!rc && too_long) rc = fail(NSMH_EINVAL, "load_fastq: a read is longer than 2^32-1 bases");
    if (!rc) {
        rs.num_reads = num_reads;
        rc = alloc_packed(rs, total_bases, s);
    }
    FQ_CK(cudaEventRecord(ev[1], s));
    if (!rc && rs.num_words) {
        // words per lane and chunk: a multiple of kFqPackUnroll (NSMH_FQ_PACK_ITERS: tuning runs only)
        uint32_t iters = kFqPackIters;
        const char *ev_iters = getenv("NSMH_FQ_PACK_ITERS");
        if (ev_iters && *ev_iters && atoi(ev_iters) > 0)
            iters = (uint32_t)std::min(4096, (atoi(ev_iters) + kFqPackUnroll - 1) / kFqPackUnroll * kFqPackUnroll);
        const uint64_t chunks = (rs.num_words + 32ull * iters - 1) / (32ull * iters);
        auto kernel = fastq_pack_kernel;
        int occ = 0;
        FQ_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, 0));
        const uint64_t resident = (uint64_t)c->num_sms * (occ > 0 ? occ : 1);
        const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((chunks + 7) / 8, resident));
        if (!rc) {
            kernel<<<blocks, 256, 0, s>>>(d_text, safe_bytes, rs.d_offsets(), src.as<uint64_t>(), num_reads, total_bases,
                                          rs.packed.as<uint32_t>(), iters);
            ++c->launches;
            FQ_CK(cudaGetLastError());
        }
    }
    FQ_CK(cudaEventRecord(ev[2], s));
    FQ_CK(cudaStreamSynchronize(s));
#undef FQ_CK
#undef FQ_TRY
    DevBuf *bufs[] = {&tile_cnt, &tile_base, &nl, &src, &len32, &tmp, &flag};
    for (DevBuf *b : bufs) b->release(s);
    float ms = 0;
    if (!rc) {
        if (cudaEventElapsedTime(&ms, ev[0], ev[2]) == cudaSuccess) c->stats.fastq_parse_ms = ms; else cudaGetLastError();
        if (cudaEventElapsedTime(&ms, ev[1], ev[2]) == cudaSuccess) c->stats.fastq_pack_ms = ms; else cudaGetLastError();
    }
    for (auto &x : ev) if (x) cudaEventDestroy(x);
    return rc;
}

int unpack_ascii_device(nsmh_ctx *c, uint64_t b0, uint64_t nb, uint8_t *d_out, cudaStream_t s) {
    if (!nb) return NSMH_OK;
    const uint64_t groups = (nb + 15) / 16;
    const uint64_t blocks = std::min<uint64_t>((groups + 255) / 256, (uint64_t)c->num_sms * 8);
    unpack_ascii_kernel<<<(unsigned)blocks, 256, 0, s>>>(c->reads.packed.as<uint32_t>(), b0, nb, d_out);
    ++c->launches;
    NSMH_CK(cudaGetLastError());
    return NSMH_OK;
}

} // namespace nsmh

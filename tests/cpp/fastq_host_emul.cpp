// TEST INFRASTRUCTURE ONLY.  Runs the FASTQ-ingest kernels of nanospring_b200/csrc/fastq_kernels.cuh
// on the host (cuda_host_shim.h: lock-step warp emulation) in the order parse_fastq_device
// (fastq.cu) launches them, with the two CUB prefix sums replaced by loops.  tests/test_fastq_emul.py
// compares the result with the oracle; it is a logic check of the device code for containers without
// a GPU, never a product path (the product is libnsmh.so and needs a GPU).
#include "cuda_host_shim.h"

#include <cstring>

#include "../../nanospring_b200/csrc/fastq_kernels.cuh"

using namespace nsmh;

extern "C" {

// text must stay readable up to safe_bytes (>= bytes).  Capacities: offsets >= newlines/4 + 3,
// words >= bytes/16 + 2.  Returns 0, or -1 when a capacity is too small / a read is too long.
int fq_emul_parse(const uint8_t *text, uint64_t bytes, uint64_t safe_bytes, unsigned grid, uint32_t pack_iters, uint64_t *offsets,
                  uint64_t offsets_cap, uint32_t *words, uint64_t words_cap, uint32_t *num_reads_out,
                  uint64_t *newlines_out) {
    const int aligned16 = (reinterpret_cast<uintptr_t>(text) & 15) == 0;
    if (reinterpret_cast<uintptr_t>(text) & 3) safe_bytes = 0;
    const uint64_t ntiles = (bytes + kFqTileBytes - 1) / kFqTileBytes;
    std::vector<uint32_t> tile_cnt(ntiles + 1, 0);
    std::vector<uint64_t> tile_base(ntiles + 1, 0);
    if (ntiles)
        emu_launch(grid, 64, [&] { fastq_count_newlines_kernel(text, bytes, aligned16, ntiles, tile_cnt.data()); });
    for (uint64_t t = 0; t < ntiles; ++t) tile_base[t + 1] = tile_base[t] + tile_cnt[t];
    const uint64_t newlines = tile_base[ntiles];
    const int last = bytes ? text[bytes - 1] : 0;
    const uint64_t num_lines = newlines + ((bytes > 0 && last != '\n') ? 1 : 0);
    const uint64_t num_reads = (num_lines + 3) / 4;
    if (num_reads + 1 > offsets_cap) return -1;
    std::vector<uint64_t> nl(newlines + 1, ~0ull), src(num_reads + 1, 0);
    std::vector<uint32_t> len32(num_reads + 1, 0xDEADBEEF);
    if (newlines)
        emu_launch(grid, 64, [&] { fastq_write_newlines_kernel(text, bytes, aligned16, ntiles, tile_base.data(), nl.data()); });
    unsigned long long too_long = 0;
    emu_launch((unsigned)((num_reads + 1 + 63) / 64), 64, [&] {
        fastq_read_table_kernel(nl.data(), newlines, num_lines, bytes, (uint32_t)num_reads, src.data(), len32.data(), &too_long);
    });
    if (too_long) return -1;
    offsets[0] = 0;
    for (uint64_t i = 0; i < num_reads; ++i) offsets[i + 1] = offsets[i] + len32[i];
    const uint64_t total = offsets[num_reads];
    const uint64_t nwords = (total + 15) / 16;
    if (nwords > words_cap) return -1;
    if (nwords)
        emu_launch(grid, 64, [&] {
            const uint32_t iters = pack_iters ? pack_iters : (uint32_t)kFqPackIters;
            fastq_pack_kernel(text, safe_bytes, offsets, src.data(), (uint32_t)num_reads, total, words, iters);
        });
    *num_reads_out = (uint32_t)num_reads;
    *newlines_out = newlines;
    return 0;
}

// "ATCG"[code] of bases [b0, b0 + nb) of a packed stream (8 zero pad words behind it)
void fq_emul_unpack(const uint32_t *words, uint64_t b0, uint64_t nb, uint8_t *out, unsigned grid) {
    if (nb) emu_launch(grid, 64, [&] { unpack_ascii_kernel(words, b0, nb, out); });
}

}  // extern "C"

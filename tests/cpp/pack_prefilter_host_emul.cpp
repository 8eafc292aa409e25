// TEST INFRASTRUCTURE ONLY.  Runs the 2-bit packers (nanospring_b200/csrc/pack_kernels.cuh) and the
// candidate pre-filter kernels (csrc/prefilter_kernels.cuh) on the host with cuda_host_shim.h, in the
// order pack.cu / prefilter.cu launch them (the CUB prefix sum of the CSR compaction replaced by a
// loop).  tests/test_pack_prefilter_emul.py compares the results with the oracle; a logic check for
// the container without a GPU, never a product path.
#include "cuda_host_shim.h"

#include <cstring>

#include "../../nanospring_b200/csrc/pack_kernels.cuh"
#include "../../nanospring_b200/csrc/prefilter_kernels.cuh"

using namespace nsmh;

extern "C" {

// ASCII bases -> packed words (W has room for ceil(n/16) + kPackPadWords words)
void pp_emul_pack_ascii(const uint8_t *src, uint64_t num_bases, uint32_t *W, unsigned grid) {
    const int aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    if (num_bases) emu_launch(grid, 64, [&] { pack_ascii_kernel(src, 0, num_bases, W, aligned); });
}

// reverse complement of every read of a packed read set
void pp_emul_pack_rc(const uint64_t *off, uint32_t n_reads, const uint32_t *Wsrc, uint32_t *W, unsigned grid) {
    const uint64_t total = off[n_reads];
    if (total) emu_launch(grid, 64, [&] { pack_rc_kernel(off, n_reads, total, Wsrc, W); });
}

// reads in the reference's DnaBitset layout (byte aligned per read) -> continuous stream
void pp_emul_pack_dnabitset(const uint64_t *off, uint32_t n_reads, const uint8_t *src, const uint64_t *src_off,
                            uint32_t *W, unsigned grid) {
    const uint64_t total = off[n_reads];
    if (!total) return;
    // two word ranges, as the chunked loader converts them (the cut is anywhere, also at 0 and at the end)
    const uint64_t nwords = (total + 15) / 16, cut = (nwords * 5) / 8;
    emu_launch(grid, 64, [&] { pack_dnabitset_kernel(off, n_reads, total, src, src_off, W, 0, cut); });
    emu_launch(grid, 64, [&] { pack_dnabitset_kernel(off, n_reads, total, src, src_off, W, cut, nwords); });
}

// NSMH_FLAG_* of every read (W must be followed by kPackPadWords zero words)
void pp_emul_read_flags(const uint64_t *off, uint32_t n_reads, const uint32_t *W, uint8_t *flags, unsigned grid) {
    if (!n_reads) return;
    const uint64_t num_words = (off[n_reads] + 15) / 16;
    std::vector<uint32_t> cnt((size_t)n_reads * kRepShifts, 0);
    if (num_words) emu_launch(grid, 64, [&] { repetitive_count_kernel(off, n_reads, num_words, W, cnt.data()); });
    emu_launch(grid, 64, [&] { repetitive_flag_kernel(off, n_reads, W, cnt.data(), flags); });
}

// CSR compaction: drop ids whose flags intersect `drop`; returns the new total (new_off [nq+1], out >= old total)
uint64_t pp_emul_csr_drop(const uint64_t *off, const uint32_t *ids, uint32_t nq, const uint8_t *flags, uint32_t n_flags,
                          uint32_t drop, uint64_t *new_off, uint32_t *out, unsigned grid) {
    std::vector<uint32_t> keep((size_t)nq + 1, 0xDEADBEEF);
    emu_launch(grid, 64, [&] { csr_keep_count_kernel(off, ids, nq, flags, n_flags, drop, keep.data()); });
    new_off[0] = 0;
    for (uint32_t q = 0; q < nq; ++q) new_off[q + 1] = new_off[q] + keep[q];
    if (nq) emu_launch(grid, 64, [&] { csr_keep_write_kernel(off, ids, nq, flags, n_flags, drop, new_off, out); });
    return new_off[nq];
}

}  // extern "C"

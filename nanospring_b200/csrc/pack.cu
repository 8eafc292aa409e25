// On-device 2-bit packing of reads.  Replaces the reference's host path
//   DnaBitset::DnaBitset / load_from_string   (src/dnaToBits.cpp:11-36, 46-71)
//   ReadData::getRead -> DnaBitset::to_string (src/ReadData.cpp:225-235, dnaToBits.cpp:81-98)
// Code: (c&2)|((c&4)>>2)  => A0 T1 C2 G3 for ANY byte (src/ReadFilter.cpp:113-115).
//
// Device layout (free to choose): ONE continuous stream of
// 2-bit codes over all reads; u32 word w holds global bases 16w..16w+15, base j
// in bits 31-2j..30-2j, so a k-mer is a funnel shift of adjacent words.
// Reverse complement works on the codes: the reference stores 2 bits per base and
// getRead() hands out "ATCG"[code] (dnaToBits.cpp:81-98), so the strings the caller
// reverse-complements (Consensus.cpp:181-184) only ever hold A/T/C/G and
// toComplement (ReadData.cpp:247-260) flips every base: code ^ 1.
#include "nsmh_internal.cuh"
#include "pack_kernels.cuh"

namespace nsmh {

int alloc_packed(ReadSet &rs, uint64_t total_bases, cudaStream_t s) {
    rs.total_bases = total_bases;
    rs.num_words = (total_bases + 15) / 16;
    NSMH_TRY(rs.packed.ensure((rs.num_words + kPackPadWords) * sizeof(uint32_t), s));
    NSMH_CK(cudaMemsetAsync(rs.packed.as<uint32_t>() + rs.num_words, 0, kPackPadWords * sizeof(uint32_t), s));
    return NSMH_OK;
}

// Packs global bases [first_base, first_base+num_bases); first_base % 16 == 0 and
// d_bases points at base `first_base`.
int pack_ascii(ReadSet &rs, const char *d_bases, uint64_t first_base, uint64_t num_bases,
               cudaStream_t s, uint32_t *launches) {
    if (num_bases == 0) return NSMH_OK;
    if (first_base % 16) return fail(NSMH_EINVAL, "pack_ascii: chunk start not word aligned");
    const uint64_t nwords = (num_bases + 15) / 16;
    int blocks = (int)((nwords + 255) / 256 < 148 * 16 ? (nwords + 255) / 256 : 148 * 16);
    int aligned = (reinterpret_cast<uintptr_t>(d_bases) & 15) == 0;
    pack_ascii_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const uint8_t *>(d_bases),
                                             first_base / 16, num_bases, rs.packed.as<uint32_t>(),
                                             aligned);
    if (launches) ++*launches;
    NSMH_CK(cudaGetLastError());
    return NSMH_OK;
}

int pack_reverse_complement(const ReadSet &src, ReadSet &dst, cudaStream_t s, uint32_t *launches) {
    dst.num_reads = src.num_reads;
    NSMH_TRY(alloc_packed(dst, src.total_bases, s));
    if (src.num_words == 0) return NSMH_OK;
    int blocks = (int)((src.num_words + 255) / 256 < 148 * 16 ? (src.num_words + 255) / 256 : 148 * 16);
    pack_rc_kernel<<<blocks, 256, 0, s>>>(src.d_offsets(), src.num_reads, src.total_bases,
                                          src.packed.as<uint32_t>(), dst.packed.as<uint32_t>());
    if (launches) ++*launches;
    NSMH_CK(cudaGetLastError());
    return NSMH_OK;
}

int pack_from_dnabitset(ReadSet &rs, const uint8_t *d_src, const uint64_t *d_src_byte_off,
                        cudaStream_t s, uint32_t *launches, uint64_t w_begin, uint64_t w_end) {
    if (w_end > rs.num_words) w_end = rs.num_words;
    if (w_begin >= w_end) return NSMH_OK;
    const uint64_t per_block = 8ULL * 32 * kDnaWordsPerLane;       // 8 warps, one pass
    const uint64_t want = (w_end - w_begin + per_block - 1) / per_block;
    const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
    pack_dnabitset_kernel<<<blocks, 256, 0, s>>>(rs.d_offsets(), rs.num_reads, rs.total_bases, d_src,
                                                 d_src_byte_off, rs.packed.as<uint32_t>(), w_begin, w_end);
    if (launches) ++*launches;
    NSMH_CK(cudaGetLastError());
    return NSMH_OK;
}

} // namespace nsmh

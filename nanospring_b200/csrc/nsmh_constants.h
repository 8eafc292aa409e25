// Layout constants shared by the host code (nsmh_internal.cuh) and the kernel headers that are also
// compiled for the host by the emulation tests (no CUDA includes here).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/nsmh.h"

namespace nsmh {

constexpr int kWordBases = 16;          // bases per packed u32 word, first base in bits 31..30
constexpr int kPackPadWords = 8;        // zero words after the last packed word (k-mer window overrun)
constexpr int kTileWords = 640;         // default words (10240 k-mer start positions) per sketch tile: 32 warps per SM fit
constexpr int kFilterMaxBits = 12;      // largest prefix width b of the sketch filter tables
constexpr int kFilterTabSize = 1 << (kFilterMaxBits + 1);  // table for b lives at [2^b, 2^(b+1))
constexpr int kFilterLambdaLog2 = 2;    // default: b = floor(log2(#kmers)) - 2  => 4..8 k-mers expected per bucket
constexpr int kFilter3MaxBits = 11;     // 3-positions-per-lookup tables: window of b+4 bits
constexpr int kFilter3TabSize = 1 << (kFilter3MaxBits + 5);   // table for b lives at [2^(b+4), 2^(b+5))
constexpr uint64_t kEmptyKey = ~0ULL;   // empty marker of the hash tables (key ~0 has its own slot)

// One 16-byte hash-table slot = one 32-byte sector per probe.  Slots start as
// all-ones: key ~0 is the empty marker and cntm1 (group size - 1) wraps to 0 on
// the first insert.  val = the read id itself for a group of one, else the start
// of the group in Tables::ids.
struct __align__(16) Slot {
    uint64_t key;
    uint32_t val;
    uint32_t cntm1;
};

// home slot / bucket of a key among `cap` of them (cap < 2^32, any value)
__host__ __device__ __forceinline__ uint64_t slot_index(uint64_t key, uint64_t cap) {
    const uint64_t h = (key * 0x9E3779B97F4A7C15ULL) >> 32;
    return (h * cap) >> 32;
}

// slots per table region: cap (even) + the slot of key ~0, padded so every region starts on a sector
__host__ __device__ __forceinline__ uint64_t region_stride(uint64_t cap) { return cap + 2; }

// ---- multi-GPU over peer memory (kernels in query_kernels.cuh / multigpu_kernels.cuh) ----
constexpr int kMgMaxRanks = NSMH_MG_MAX_RANKS;
// Probe results travel as one u64 per (read, hash): {id | group start} | (group size) << 32.
// In the arena of the rank that owns the reads they are blocked by table owner and, inside an owner's
// block, by groups of kPeerCols hash functions: group b is a dense array [local row][kPeerCols] (an owner
// whose number of hash functions is not a multiple of kPeerCols leaves the last columns of its last group
// unused).  A block of the probe kernel produces the tile [256 rows][kPeerCols] in shared memory and sends
// it as ONE contiguous, 64-byte-aligned run per read owner - a bulk copy (cp.async.bulk shared -> global)
// when the tile has one owner - so the NVLink traffic is full 128-byte lines.
constexpr int kProbeCols = 4;      // probe kernels: adjacent hash functions per thread (4 x 8 B of keys = one sector)
constexpr int kProbeRows = 256;    // ... and queries per block = threads per block
constexpr int kPeerCols = 8;
__host__ __device__ __forceinline__ uint32_t peer_padded_cols(uint32_t nc) { return (nc + kPeerCols - 1) / kPeerCols * kPeerCols; }
// index of (local row q, hash function jj of an owner) inside that owner's block of a rank with `rows` reads
__host__ __device__ __forceinline__ size_t peer_result_index(uint32_t rows, uint32_t q, uint32_t jj) {
    return (size_t)rows * (jj / kPeerCols * kPeerCols) + (size_t)q * kPeerCols + jj % kPeerCols;
}
constexpr int kInboxMaxGroup = 32;           // groups of up to this many ids are pushed to the read owner's inbox
constexpr uint32_t kInboxFlag = 0x80000000u; // in the size field: "val is a position in your inbox"
constexpr uint32_t kPairFlag = 0x40000000u;  // in the size field: a group of two, ids in bits 0..30 and 31..61 of the word
struct PeerDst {
    uint64_t *pr[kMgMaxRanks];        // probe-result area in rank r's arena
    uint32_t *inbox[kMgMaxRanks];     // this rank's segment of rank r's inbox
    uint32_t inbox_cap[kMgMaxRanks];  // ... and its capacity in ids
    uint32_t *cursor;                 // [kMgMaxRanks] local: ids pushed to rank r so far (zeroed per run)
    uint32_t row_end[kMgMaxRanks];    // global row index one past rank r's rows
    uint32_t world, col0, ncols;      // hash functions this rank owns: [col0, col0 + ncols)
    uint32_t pcol0;                   // where this rank's block starts in a destination's pr, in (padded) columns
    uint32_t chunk0;                  // first 256-row chunk that holds a read of this rank (< number of chunks)
};
// Probe results stored locally by the table owners + where the owners keep their group ids.
struct PeerLists {
    const uint64_t *pr;                     // local probe-result area (layout above)
    const uint32_t *inbox;                  // local inbox: segment o (inbox_cap ids) is filled by rank o
    uint32_t inbox_cap;
    const uint32_t *ids[kMgMaxRanks];       // ids array of the rank that owns the hash function
    uint32_t col_end[kMgMaxRanks];          // hash functions [col_end[r-1], col_end[r]) belong to rank r
    uint32_t world, n, rows;
};

} // namespace nsmh

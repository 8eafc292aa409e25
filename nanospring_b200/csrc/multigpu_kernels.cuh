// Device code and arena layout of the multi-GPU path over peer memory (see multigpu.cu for the
// stages).  Kept free of runtime-API includes so that tests/cpp/multigpu_host_emul.cpp can run a whole
// "world" of ranks on the host - arenas as host buffers, the ranks' kernels one after the other -
// and check every rank's candidate lists against the oracle without a GPU.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "nsmh_constants.h"
#include "nsmh_ldst.cuh"

namespace nsmh {

// dynamic shared memory of the kernels here: a parameter in the host emulation (like sketch_filter_kernel)
#ifndef NSMH_HOST_EMUL
#define NSMH_MG_SMEM_PARAM
#define NSMH_MG_SMEM_DECL extern __shared__ __align__(16) uint8_t mg_smem[];
#else
#define NSMH_MG_SMEM_PARAM , uint8_t *mg_smem
#define NSMH_MG_SMEM_DECL
#endif

// ---- who owns which hash functions -------------------------------------------------------------
// col_end[r] = one past the last hash function of rank r; the first n % world ranks own one more than the
// others (15/15/15/15 of 60 at 4 ranks, 8/8/8/8/7/7/7/7 at 8).  false: more ranks than hash functions.
inline bool mg_split_columns(uint32_t n, uint32_t world, uint32_t *col_end) {
    if (n < world) return false;
    uint32_t cend = 0;
    for (uint32_t r = 0; r < world; ++r) {
        cend += n / world + (r < n % world ? 1u : 0u);
        col_end[r] = cend;
    }
    return true;
}

// padded column offset of every owner's block in pr (pcol[o]) and their total
inline uint32_t mg_padded_offsets(const uint32_t *col_end, uint32_t world, uint32_t *pcol) {
    uint32_t acc = 0;
    for (uint32_t o = 0; o < world; ++o) {
        pcol[o] = acc;
        acc += peer_padded_cols(col_end[o] - (o ? col_end[o - 1] : 0u));
    }
    return acc;
}

// ---- one rank's arena ----------------------------------------------------------------------------
//   m      [total_rows][ncols] u64   sketch columns of the hash functions this rank owns, all reads
//   pr     [rows][pr_cols]     u64   probe results of this rank's reads, blocked by table owner and, inside an
//                                    owner's block, by groups of kPeerCols hash functions (nsmh_constants.h);
//                                    pr_cols = sum over owners of their hash functions rounded up to kPeerCols
//   ids    [total_rows*ncols]  u32   group members of the owned tables
//   inbox  [world][inbox_cap]  u32   small groups pushed along by the table owners
//   flags  2 x kMgMaxRanks epochs, error flag, kMgMaxRanks inbox cursors
struct MgLayout {
    uint64_t off_m, off_pr, off_ids, off_inbox, off_flags, arena_bytes;
    uint32_t inbox_cap;
};

inline MgLayout mg_layout(uint32_t total_rows, uint32_t ncols, uint32_t my_rows, uint32_t pr_cols, uint32_t max_cols,
                          uint32_t world, long long inbox_cap_override) {
    auto align256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
    auto max_sz = [](size_t a, size_t b) { return a > b ? a : b; };
    MgLayout t;
    const size_t items = max_sz((size_t)total_rows * ncols, 1);
    const size_t local = max_sz((size_t)my_rows * pr_cols, 1);
    size_t off = 0;
    t.off_m = off;      off = align256(off + items * sizeof(uint64_t));
    t.off_pr = off;     off = align256(off + local * sizeof(uint64_t));
    t.off_ids = off;    off = align256(off + items * sizeof(uint32_t));
    // inbox: one segment per source rank, room for 2 ids per (local read, hash of that rank)
    uint64_t cap = 2ull * my_rows * max_cols;
    if (cap < 64) cap = 64;
    if (cap > (1ull << 30)) cap = 1ull << 30;
    if (inbox_cap_override > 0) cap = (uint64_t)(inbox_cap_override < (1ll << 30) ? inbox_cap_override : (1ll << 30));
    t.inbox_cap = (uint32_t)(cap + 3 & ~3ull);          // segments and the bulk copies into them stay 16-byte aligned
    t.off_inbox = off;  off = align256(off + (size_t)world * t.inbox_cap * sizeof(uint32_t));
    t.off_flags = off;  off = align256(off + (3 * kMgMaxRanks + 8) * sizeof(uint32_t));
    t.arena_bytes = off;
    return t;
}

// ---- stage 1: sketch rows -> column blocks in the owners' arenas ----------------------------------
struct ScatterArgs {
    uint64_t *m[kMgMaxRanks];         // column block of rank o: [total_rows][ncols_o]
    uint32_t col_end[kMgMaxRanks];
    uint32_t world, row0;             // row0: global row of local row 0
};

// A block takes kScatterRows consecutive local rows: the tile [rows][n] is contiguous in the sketch matrix
// and comes in with coalesced 128-bit loads; for every owner o the sub-block [rows][ncols_o] is contiguous in
// the owner's column block too (M_o is row-major with ncols_o columns), so it leaves as a dense run of
// 8-byte stores - full 128-byte lines over NVLink.  (The first version stored one 32-byte piece per thread
// at a stride of ncols_o * 8 bytes: half-written lines, and twice the time at 8 ranks than at 2.)
constexpr int kScatterRows = 32;

__global__ void __launch_bounds__(256)
mg_scatter_columns_kernel(const uint64_t *__restrict__ S, uint32_t rows, uint32_t n, ScatterArgs a NSMH_MG_SMEM_PARAM) {
    NSMH_MG_SMEM_DECL
    uint64_t *tile = reinterpret_cast<uint64_t *>(mg_smem);        // [kScatterRows][n]
    const uint32_t tiles = (rows + kScatterRows - 1) / kScatterRows;
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const uint32_t i0 = t * kScatterRows, nr = min((uint32_t)kScatterRows, rows - i0);
        const uint32_t words = nr * n;
        const uint64_t *src = S + (size_t)i0 * n;
        __syncthreads();                    // the previous tile has left
        if ((n & 1) == 0) {                 // rows are 16-byte aligned
            for (uint32_t w = threadIdx.x * 2; w < words; w += blockDim.x * 2) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(src + w);
                tile[w] = v.x;
                tile[w + 1] = v.y;
            }
        } else {
            for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) tile[w] = src[w];
        }
        __syncthreads();
        uint32_t cb = 0;
        for (uint32_t o = 0; o < a.world; ++o) {
            const uint32_t nc = a.col_end[o] - cb;
            uint64_t *dst = a.m[o] + (size_t)(a.row0 + i0) * nc;
            const uint32_t cnt = nr * nc;
            for (uint32_t e = threadIdx.x; e < cnt; e += blockDim.x) {
                const uint32_t i = e / nc, c = e - i * nc;
                dst[e] = tile[i * n + cb + c];
            }
            cb = a.col_end[o];
        }
    }
}

// Entries whose values arrived after the scatter (nsmh_mg_sketch_run: the sketch's exact fix-up pass runs on the
// second stream beside mg_scatter_columns_kernel, which sent all-ones for them): list[e] = local row * n + column,
// vals[e] the value.  It goes into the local sketch matrix and into the owner's column block.
__global__ void __launch_bounds__(256)
mg_scatter_list_kernel(uint64_t *__restrict__ S, uint32_t n, const uint32_t *__restrict__ list,
                       const unsigned int *__restrict__ count, const uint64_t *__restrict__ vals, ScatterArgs a) {
    const uint32_t todo = *count;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < todo; e += gridDim.x * blockDim.x) {
        const uint64_t v = vals[e];
        if (v == ~0ULL) continue;           // a read of k-1 bases: all-ones is its value, and that was sent
        const uint32_t t = list[e];
        const uint32_t row = t / n, l = t - row * n;
        S[t] = v;
        uint32_t o = 0, cb = 0;
        while (o + 1 < a.world && l >= a.col_end[o]) { cb = a.col_end[o]; ++o; }
        const uint32_t nc = a.col_end[o] - cb;
        a.m[o][(size_t)(a.row0 + row) * nc + (l - cb)] = v;
    }
}

} // namespace nsmh

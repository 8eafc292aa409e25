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

// ---- who owns which hash functions -------------------------------------------------------------
// Hash functions are handed out in units of 4 (one 32-byte sector of a sketch row) when possible.
// col_end[r] = one past the last hash function of rank r.  false: more ranks than units.
inline bool mg_split_columns(uint32_t n, uint32_t world, uint32_t *col_end) {
    const uint32_t unit = (n % 4 == 0 && n / 4 >= world) ? 4 : 1;
    const uint32_t units = n / unit;
    if (units < world) return false;
    uint32_t cend = 0;
    for (uint32_t r = 0; r < world; ++r) {
        cend += (units / world + (r < units % world ? 1u : 0u)) * unit;
        col_end[r] = cend;
    }
    return true;
}

// ---- one rank's arena ----------------------------------------------------------------------------
//   m      [total_rows][ncols] u64   sketch columns of the hash functions this rank owns, all reads
//   pr     [rows][n_total]     u64   probe results of this rank's reads, blocked by table owner
//   ids    [total_rows*ncols]  u32   group members of the owned tables
//   inbox  [world][inbox_cap]  u32   small groups pushed along by the table owners
//   flags  2 x kMgMaxRanks epochs, error flag, kMgMaxRanks inbox cursors
struct MgLayout {
    uint64_t off_m, off_pr, off_ids, off_inbox, off_flags, arena_bytes;
    uint32_t inbox_cap;
};

inline MgLayout mg_layout(uint32_t total_rows, uint32_t ncols, uint32_t my_rows, uint32_t n_total, uint32_t max_cols,
                          uint32_t world, long long inbox_cap_override) {
    auto align256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
    auto max_sz = [](size_t a, size_t b) { return a > b ? a : b; };
    MgLayout t;
    const size_t items = max_sz((size_t)total_rows * ncols, 1);
    const size_t local = max_sz((size_t)my_rows * n_total, 1);
    size_t off = 0;
    t.off_m = off;      off = align256(off + items * sizeof(uint64_t));
    t.off_pr = off;     off = align256(off + local * sizeof(uint64_t));
    t.off_ids = off;    off = align256(off + items * sizeof(uint32_t));
    // inbox: one segment per source rank, room for 2 ids per (local read, hash of that rank)
    uint64_t cap = 2ull * my_rows * max_cols;
    if (cap < 64) cap = 64;
    if (cap > (1ull << 30)) cap = 1ull << 30;
    if (inbox_cap_override > 0) cap = (uint64_t)(inbox_cap_override < (1ll << 30) ? inbox_cap_override : (1ll << 30));
    t.inbox_cap = (uint32_t)cap;
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

// One warp per local row (strided), a lane per hash function: loads are the contiguous sketch
// row, stores are runs of ncols_o * 8 bytes in the owner's memory.
__global__ void __launch_bounds__(256)
mg_scatter_columns_kernel(const uint64_t *__restrict__ S, uint32_t rows, uint32_t n, ScatterArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t w0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (uint32_t j = lane; j < n; j += 32) {
        uint32_t o = 0;
        while (o + 1 < a.world && j >= a.col_end[o]) ++o;
        const uint32_t cb = o ? a.col_end[o - 1] : 0u, nc = a.col_end[o] - cb;
        uint64_t *dst = a.m[o] + (size_t)a.row0 * nc + (j - cb);
        uint32_t i = w0;
        for (; i + 3 * warps < rows; i += 4 * warps) {          // four loads in flight per lane
            const uint64_t v0 = __ldg(S + (size_t)i * n + j), v1 = __ldg(S + (size_t)(i + warps) * n + j);
            const uint64_t v2 = __ldg(S + (size_t)(i + 2 * warps) * n + j), v3 = __ldg(S + (size_t)(i + 3 * warps) * n + j);
            dst[(size_t)i * nc] = v0;
            dst[(size_t)(i + warps) * nc] = v1;
            dst[(size_t)(i + 2 * warps) * nc] = v2;
            dst[(size_t)(i + 3 * warps) * nc] = v3;
        }
        for (; i < rows; i += warps) dst[(size_t)i * nc] = __ldg(S + (size_t)i * n + j);
    }
}

// The same when every rank owns a multiple of 4 hash functions: a thread moves 4 adjacent
// columns of one row = one 32-byte sector in (256-bit load) and one out (256-bit store).
__global__ void __launch_bounds__(256)
mg_scatter_columns4_kernel(const uint64_t *__restrict__ S, uint32_t rows, uint32_t n, ScatterArgs a) {
    const uint32_t groups = n >> 2;
    const uint64_t total = (uint64_t)rows * groups, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total; t += stride) {
        const uint32_t i = (uint32_t)(t / groups), j = (uint32_t)(t - (uint64_t)i * groups) << 2;
        uint64_t v0, v1, v2, v3;
        ldg256(S + t * 4, v0, v1, v2, v3);
        uint32_t o = 0;
        while (o + 1 < a.world && j >= a.col_end[o]) ++o;
        const uint32_t cb = o ? a.col_end[o - 1] : 0u, nc = a.col_end[o] - cb;
        uint64_t *d = a.m[o] + (size_t)(a.row0 + i) * nc + (j - cb);
        stg256(d, v0, v1, v2, v3);
    }
}

} // namespace nsmh

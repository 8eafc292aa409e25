// TEST INFRASTRUCTURE ONLY.  Builds the per-hash-function tables with the kernels of
// nanospring_b200/csrc/table_kernels.cuh on the host (cuda_host_shim.h; whole 256-thread blocks run
// concurrently as OS threads, NSMH_HOST_EMUL swaps the three PTX accesses for host atomics) in the
// order build_tables (table.cu) launches them, then probes them with the lookup kernel's body
// (query_kernels.cuh: count_body<ProbeSrc>).  tests/test_table_emul.py compares key counts and candidate
// lists with the oracle; a logic check for the container without a GPU, never a product path.
#define NSMH_HOST_EMUL 1
#include "cuda_host_shim.h"

#include <algorithm>
#include <cstring>

#include "../../nanospring_b200/csrc/table_kernels.cuh"
#include "../../nanospring_b200/csrc/query_kernels.cuh"

using namespace nsmh;

namespace {
struct EmulTables {
    std::vector<Slot> slots;
    std::vector<uint32_t> ids;
    uint64_t cap = 0;
    uint32_t rows = 0, n = 0;
} g_t;
}  // namespace

extern "C" {

// sk [rows][n] -> tables kept inside the library (one set at a time).  max_blocks bounds the grid of the
// insert kernel (build_tables uses the resident blocks of the device).  Returns 0.
// deferred entries (nsmh_sketch_build): list [count] = row * n + column of the entries of sk that still hold all-ones,
// vals [count] their keys; sk receives them.  list == nullptr: the plain build.
static int build_impl(uint64_t *sk, uint32_t rows, uint32_t n, unsigned max_blocks, const uint32_t *list,
                      unsigned int count, const uint64_t *vals);

int table_emul_build(const uint64_t *sk, uint32_t rows, uint32_t n, unsigned max_blocks) {
    return build_impl(const_cast<uint64_t *>(sk), rows, n, max_blocks, nullptr, 0, nullptr);
}

int table_emul_build_deferred(uint64_t *sk, uint32_t rows, uint32_t n, unsigned max_blocks, const uint32_t *list,
                              unsigned int count, const uint64_t *vals) {
    return build_impl(sk, rows, n, max_blocks, list, count, vals);
}

static int build_impl(uint64_t *sk, uint32_t rows, uint32_t n, unsigned max_blocks, const uint32_t *list,
                      unsigned int count, const uint64_t *vals) {
    g_t = EmulTables();
    g_t.rows = rows;
    g_t.n = n;
    const uint64_t cap = std::max<uint64_t>(16, 2ULL * rows);
    const uint64_t nslots = (uint64_t)n * region_stride(cap);
    g_t.cap = cap;
    g_t.slots.resize(nslots);
    memset(g_t.slots.data(), 0xFF, nslots * sizeof(Slot));
    const uint64_t items = (uint64_t)rows * n;
    g_t.ids.assign(items ? items : 1, 0xDEADBEEFu);
    if (!items) return 0;
    const uint64_t units = (uint64_t)((rows + kBuildRows - 1) / kBuildRows) * ((n + kBuildCols - 1) / kBuildCols);
    const uint32_t blocks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(units, max_blocks));
    const uint32_t seg_cap = kBuildRows * kBuildCols;           // one segment per work unit
    const size_t nseg = (size_t)units * seg_cap;
    std::vector<uint32_t> multi(std::max<size_t>(nseg, 1) * 4, 0);
    std::vector<unsigned int> tmp(8 + 2 * (size_t)units, 0);
    BuildArgs a;
    a.sk = sk;
    a.slots = g_t.slots.data();
    a.ids = g_t.ids.data();
    a.m_slot = multi.data();
    a.m_id = a.m_slot + nseg;
    a.m_rank = a.m_id + nseg;
    a.g_slot = a.m_rank + nseg;
    a.counters = tmp.data();
    a.seg_count = a.counters + 8;
    a.cap = cap;
    a.rows = rows;
    a.n = n;
    a.seg_cap = seg_cap;
    a.segments = (uint32_t)units;
    a.skip_empty = list ? 1u : 0u;
    emu_launch_block(blocks, kBuildRows, [&] { table_insert_kernel(a); });
    if (list) emu_launch(2, 256, [&] { table_insert_list_kernel(a, list, &count, vals, sk); });
    emu_launch(2, 256, [&] { table_groups_kernel(a); });
    emu_launch(2, 256, [&] { table_fill_kernel(a); });
    return 0;
}

// distinct keys of table j (BBHashMap::numKeys, BBHashMap.cpp:35)
uint32_t table_emul_num_keys(uint32_t j) {
    uint32_t out = 0;
    const Slot *region = g_t.slots.data() + (uint64_t)j * region_stride(g_t.cap);
    emu_launch(1, 64, [&] { table_count_keys_kernel(region, g_t.cap + 1, &out); });
    return out;
}

// nq query sketches [nq][n] against the tables through the lookup kernel's body.  Outputs as in
// count_emul_run (query_host_emul.cpp); counters [4] zeroed by the caller.
void table_emul_query(const uint64_t *qsk, uint32_t nq, uint32_t thr, unsigned grid, uint32_t *qcount, uint64_t *qpos,
                      uint32_t *tmp_ids, uint64_t tmp_cap, uint32_t *heavy_list, unsigned long long *counters) {
    ProbeSrc src;
    src.qsk = qsk;
    src.slots = g_t.slots.data();
    src.ids = g_t.ids.data();
    src.pval = nullptr;
    src.pcnt = nullptr;
    src.cap = g_t.cap;
    src.n = g_t.n;
    CountArgs a;
    a.qcount = qcount;
    a.qpos = qpos;
    a.tmp_ids = tmp_ids;
    a.tmp_cap = tmp_cap;
    a.heavy_list = heavy_list;
    a.counters = counters;
    a.nq = nq;
    a.thr = thr ? thr : 1;
    std::vector<uint32_t> smem((size_t)grid * kLookupWarps * kWarpWords + 4, 0xA5A5A5A5u);
    uint32_t *base = smem.data();
    while (reinterpret_cast<uintptr_t>(base) & 15) ++base;
    emu_launch(grid, kLookupWarps * 32, [&] { count_body(src, a, base + (size_t)blockIdx.x * kLookupWarps * kWarpWords); });
}

}  // extern "C"

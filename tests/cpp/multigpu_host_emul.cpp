// TEST INFRASTRUCTURE ONLY.  A whole "world" of ranks of the peer-memory multi-GPU path on the host:
// every rank's arena is a host buffer laid out by mg_layout (csrc/multigpu_kernels.cuh, the function
// nsmh_mg_init uses), and the stages of nsmh_mg_run (csrc/multigpu.cu) run rank after rank with the
// device code compiled for the host (cuda_host_shim.h): scatter of the sketch columns to the table
// owners, table build per owner, probe_to_peers (results and small groups stored into the read owners'
// arenas), counting with PeerSrc (larger groups read from the owners' ids).  The device-side flag
// barriers are not needed: a stage finishes on all ranks before the next begins.
// tests/test_multigpu_emul.py compares every rank's candidate lists with the oracle.
#define NSMH_HOST_EMUL 1
#define NSMH_STAGE_IDS 96       // probe_to_peers_kernel: some tiles overflow the staging area of their inbox ids, some do not
#include "cuda_host_shim.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "../../nanospring_b200/csrc/table_kernels.cuh"
#include "../../nanospring_b200/csrc/query_kernels.cuh"
#include "../../nanospring_b200/csrc/multigpu_kernels.cuh"

using namespace nsmh;

extern "C" {

// S [total_rows][n]: the global sketch matrix, rows of rank r = [sum(rows[:r]), ...).  Per-rank outputs are
// concatenated in rank order: qcount [total_rows + world] (rank r's block has rows[r] + 1 entries),
// qpos [total_rows], tmp [world][tmp_cap], heavy [total_rows + world], counters [world][4].
// Returns 0; -1 when the world does not fit (more ranks than hash functions).
int mg_emul_run(const uint64_t *S, const uint32_t *rows, uint32_t world, uint32_t n, uint32_t thr,
                long long inbox_cap_override, unsigned grid, uint32_t *qcount, uint64_t *qpos, uint32_t *tmp,
                uint64_t tmp_cap, uint32_t *heavy, unsigned long long *counters, uint32_t *col_end_out) {
    uint32_t col_end[kMgMaxRanks] = {0}, row_end[kMgMaxRanks] = {0};
    if (world < 1 || world > (uint32_t)kMgMaxRanks || !mg_split_columns(n, world, col_end)) return -1;
    uint64_t total64 = 0;
    for (uint32_t r = 0; r < world; ++r) {
        total64 += rows[r];
        row_end[r] = (uint32_t)total64;
        col_end_out[r] = col_end[r];
    }
    const uint32_t total_rows = (uint32_t)total64;
    uint32_t max_cols = 0;
    for (uint32_t r = 0; r < world; ++r) max_cols = std::max(max_cols, col_end[r] - (r ? col_end[r - 1] : 0u));
    uint32_t pcol[kMgMaxRanks] = {0};
    const uint32_t pr_cols = mg_padded_offsets(col_end, world, pcol);
    std::vector<MgLayout> lay(world);
    std::vector<uint8_t *> arena(world, nullptr);
    for (uint32_t r = 0; r < world; ++r) {
        const uint32_t ncols = col_end[r] - (r ? col_end[r - 1] : 0u);
        lay[r] = mg_layout(total_rows, ncols, rows[r], pr_cols, max_cols, world, inbox_cap_override);
        arena[r] = static_cast<uint8_t *>(aligned_alloc(256, (lay[r].arena_bytes + 255) & ~(size_t)255));
        memset(arena[r], 0xA5, lay[r].arena_bytes);                               // nothing may rely on zeroes ...
        memset(arena[r] + lay[r].off_flags, 0, (3 * kMgMaxRanks + 8) * sizeof(uint32_t));   // ... but the flags / cursors
    }
    // ---- stage 1: scatter the sketch columns to their owners ----
    for (uint32_t r = 0; r < world; ++r) {
        if (!rows[r]) continue;
        ScatterArgs sa;
        for (uint32_t o = 0; o < (uint32_t)kMgMaxRanks; ++o) {
            sa.m[o] = o < world ? reinterpret_cast<uint64_t *>(arena[o] + lay[o].off_m) : nullptr;
            sa.col_end[o] = o < world ? col_end[o] : 0u;
        }
        sa.world = world;
        sa.row0 = r ? row_end[r - 1] : 0u;
        // nsmh_mg_sketch_run: some entries are still all-ones when the columns are scattered (the sketch's fix-up pass
        // runs beside the scatter); their values follow from a list (mg_scatter_list_kernel).  Every 7th entry here.
        std::vector<uint64_t> Sr(S + (size_t)sa.row0 * n, S + (size_t)(sa.row0 + rows[r]) * n);
        std::vector<uint32_t> list;
        std::vector<uint64_t> vals;
        for (size_t t = 0; t < Sr.size(); ++t)
            if (t % 7 == 3 || Sr[t] == ~0ULL) {                 // sketch_missing_kernel lists every all-ones entry
                list.push_back((uint32_t)t);
                vals.push_back(Sr[t]);
                Sr[t] = ~0ULL;
            }
        const unsigned int count = (unsigned int)list.size();
        list.push_back(0);
        vals.push_back(0);
        std::vector<uint64_t> tile((size_t)kScatterRows * n + 2, 0xA5A5A5A5A5A5A5A5ull);
        emu_launch_block(grid, 256, [&] { mg_scatter_columns_kernel(Sr.data(), rows[r], n, sa, reinterpret_cast<uint8_t *>(tile.data())); });
        emu_launch(2, 64, [&] { mg_scatter_list_kernel(Sr.data(), n, list.data(), &count, vals.data(), sa); });
        if (memcmp(Sr.data(), S + (size_t)sa.row0 * n, Sr.size() * sizeof(uint64_t)) != 0) return -7;
    }
    // ---- stage 2: every owner builds its tables over all rows ----
    const uint64_t cap = std::max<uint64_t>(16, 2ULL * total_rows);
    std::vector<std::vector<Slot>> slots(world);
    for (uint32_t r = 0; r < world; ++r) {
        const uint32_t col0 = r ? col_end[r - 1] : 0u, ncols = col_end[r] - col0;
        slots[r].resize((size_t)ncols * region_stride(cap));
        memset(slots[r].data(), 0xFF, slots[r].size() * sizeof(Slot));
        if (!total_rows) continue;
        const uint64_t units = (uint64_t)((total_rows + kBuildRows - 1) / kBuildRows) * ((ncols + kBuildCols - 1) / kBuildCols);
        const uint32_t blocks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(units, 3));
        const uint32_t seg_cap = kBuildRows * kBuildCols;           // one segment per work unit
        const size_t nseg = (size_t)units * seg_cap;
        std::vector<uint32_t> multi(std::max<size_t>(nseg, 1) * 4, 0);
        std::vector<unsigned int> btmp(8 + 2 * (size_t)units, 0);
        BuildArgs a;
        a.sk = reinterpret_cast<const uint64_t *>(arena[r] + lay[r].off_m);
        a.slots = slots[r].data();
        a.ids = reinterpret_cast<uint32_t *>(arena[r] + lay[r].off_ids);
        a.m_slot = multi.data();
        a.m_id = a.m_slot + nseg;
        a.m_rank = a.m_id + nseg;
        a.g_slot = a.m_rank + nseg;
        a.counters = btmp.data();
        a.seg_count = a.counters + 8;
        a.cap = cap;
        a.rows = total_rows;
        a.n = ncols;
        a.seg_cap = seg_cap;
        a.segments = (uint32_t)units;
        a.skip_empty = 0;
        emu_launch_block(blocks, kBuildRows, [&] { table_insert_kernel(a); });
        emu_launch(2, 256, [&] { table_groups_kernel(a); });
        emu_launch(2, 256, [&] { table_fill_kernel(a); });
    }
    // ---- stage 3: every owner probes its tables for all rows and stores the results at the read owners ----
    for (uint32_t r = 0; r < world && total_rows; ++r) {
        const uint32_t col0 = r ? col_end[r - 1] : 0u, ncols = col_end[r] - col0;
        ProbeSrc src;
        src.qsk = reinterpret_cast<const uint64_t *>(arena[r] + lay[r].off_m);
        src.slots = slots[r].data();
        src.ids = reinterpret_cast<const uint32_t *>(arena[r] + lay[r].off_ids);
        src.pval = nullptr;
        src.pcnt = nullptr;
        src.cap = cap;
        src.n = ncols;
        PeerDst pd;
        for (uint32_t o = 0; o < (uint32_t)kMgMaxRanks; ++o) {
            const bool in = o < world;
            pd.pr[o] = in ? reinterpret_cast<uint64_t *>(arena[o] + lay[o].off_pr) : nullptr;
            pd.inbox[o] = in ? reinterpret_cast<uint32_t *>(arena[o] + lay[o].off_inbox) + (size_t)r * lay[o].inbox_cap : nullptr;
            pd.inbox_cap[o] = in ? lay[o].inbox_cap : 0u;
            pd.row_end[o] = in ? row_end[o] : 0u;
        }
        pd.cursor = reinterpret_cast<uint32_t *>(arena[r] + lay[r].off_flags) + 2 * kMgMaxRanks + 8;
        pd.world = world;
        pd.col0 = col0;
        pd.ncols = ncols;
        pd.pcol0 = pcol[r];
        pd.chunk0 = std::min<uint32_t>(r ? row_end[r - 1] : 0u, total_rows - 1) / kProbeRows;
        emu_launch_block(grid, kProbeRows, [&] { probe_to_peers_kernel(src, total_rows, pd); });
    }
    // ---- stage 4: every rank thresholds its own reads ----
    size_t qc_off = 0, q_off = 0;
    for (uint32_t r = 0; r < world; ++r) {
        PeerSrc src;
        src.L.pr = reinterpret_cast<const uint64_t *>(arena[r] + lay[r].off_pr);
        src.L.inbox = reinterpret_cast<const uint32_t *>(arena[r] + lay[r].off_inbox);
        src.L.inbox_cap = lay[r].inbox_cap;
        for (uint32_t o = 0; o < (uint32_t)kMgMaxRanks; ++o) {
            src.L.ids[o] = o < world ? reinterpret_cast<const uint32_t *>(arena[o] + lay[o].off_ids) : nullptr;
            src.L.col_end[o] = o < world ? col_end[o] : 0u;
        }
        src.L.world = world;
        src.L.n = n;
        src.L.rows = rows[r];
        CountArgs a;
        a.qcount = qcount + qc_off;
        a.qpos = qpos + q_off;
        a.tmp_ids = tmp + (size_t)r * tmp_cap;
        a.tmp_cap = tmp_cap;
        a.heavy_list = heavy + qc_off;
        a.counters = counters + 4 * (size_t)r;
        a.nq = rows[r];
        a.thr = thr ? thr : 1;
        if (rows[r]) {
            std::vector<uint32_t> smem((size_t)grid * kLookupWarps * kWarpWords + 4, 0xA5A5A5A5u);
            uint32_t *base = smem.data();
            while (reinterpret_cast<uintptr_t>(base) & 15) ++base;
            emu_launch(grid, kLookupWarps * 32, [&] { count_body(src, a, base + (size_t)blockIdx.x * kLookupWarps * kWarpWords); });
        }
        qc_off += rows[r] + 1;
        q_off += rows[r];
    }
    for (uint32_t r = 0; r < world; ++r) free(arena[r]);
    return 0;
}

}  // extern "C"

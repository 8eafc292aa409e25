// Per-hash-function sketch tables on the device.  Replaces
//   MinHashReadFilter::populateHashTables   (src/ReadFilter.cpp:159-172)
//   BBHashMap::initialize                    (src/BBHashMap.cpp:10-99)
// The reference builds, per hash function j, a minimal perfect hash over the
// distinct keys of sketch column j plus CSR arrays keys[] / startPosInReadIds[] /
// readIds[]; keys are stored and verified on lookup (BBHashMap.cpp:105-106), so
// the structure is an exact dictionary key -> list of read ids and its answers do
// not depend on the MPHF.
//
// Here: one open-addressing region of 16-byte slots {key, val, cnt-1} per hash function
// (cap = 2*reads slots + one extra slot for the key that equals the empty marker ~0).
// The two slots of a 32-byte sector form a bucket: a probe fetches one sector and checks
// both, so nearly every probe sequence ends after a single DRAM access.
//   pass 1  insert : linear probing; an empty slot is claimed with ONE 128-bit
//                    compare-and-swap that writes key, read id and count together, so a
//                    group of one (most groups) is finished by a single atomic and needs
//                    no second access at lookup time.  Later arrivals bump the count
//                    (their rank inside the group) and are appended to a compact
//                    "multi" list; the second arrival also registers the group.
//   pass 2  groups : every registered group gets a range of `ids` from a warp-
//                    aggregated cursor; the inlined first id moves to ids[start]
//   pass 3  fill   : the multi list writes ids[start + rank]
// Passes 2 and 3 run over the compact lists only (device-side counts, no host sync).
// Work is ordered by hash function: a block handles 256 rows x 4 adjacent columns (one
// 32-byte sector per row) and consecutive blocks stay in the same columns, so the
// regions being filled (a few MB each) are L2 resident while they are hammered with
// atomics instead of spreading random sectors over the whole table.
// Inside a group the id order is arbitrary; every consumer sorts (ReadFilter.cpp:73).
#include <algorithm>
#include <cstdlib>

#include "nsmh_internal.cuh"
#include "table_kernels.cuh"

namespace nsmh {

// Clear the tables for `rows` reads on the copy stream, so that the (bandwidth-bound) memset
// overlaps whatever the main stream does next (nsmh_sketch calls this before sketching: the
// tables do not depend on the sketches until the first insert).
// slots per read in every table region (even).  2 = load factor 0.5: 10 % of the keys do not fit their home
// bucket, so nearly every query (60 probes) pays a second dependent bucket access; NSMH_TABLE_SLOTS_PER_READ
// changes it for A/B runs.
static uint64_t table_cap(uint32_t rows) {
    static const uint64_t per_read = [] {
        const char *e = getenv("NSMH_TABLE_SLOTS_PER_READ");
        const long v = e && *e ? atol(e) : 2;
        return (uint64_t)(v < 2 ? 2 : (v > 16 ? 16 : v)) & ~1ULL;
    }();
    return std::max<uint64_t>(16, per_read * rows);
}

// in_order: clear on the context's own stream (the pipelined loaders: that stream idles until the first chunk
// has arrived, and the copy stream is busy with the bus) instead of the copy stream (beside the sketch kernels)
int preclear_tables(nsmh_ctx *c, uint32_t rows, bool in_order) {
    Tables &T = c->tables;
    // an earlier clear may still run on the copy stream: nothing below may free the buffer under it
    if (c->precleared_rows) NSMH_CK(cudaStreamWaitEvent(c->stream, c->ev_cleared, 0));
    c->precleared_rows = 0;
    const uint64_t cap = table_cap(rows);
    const uint64_t nslots = (uint64_t)c->n * region_stride(cap);
    if (rows == 0 || nslots >= (1ULL << 32)) return NSMH_OK;      // build_tables reports the error
    NSMH_TRY(T.slots.ensure(nslots * sizeof(Slot), c->stream));
    if (in_order) {
        NSMH_CK(cudaMemsetAsync(T.slots.p, 0xFF, nslots * sizeof(Slot), c->stream));
        NSMH_CK(cudaEventRecord(c->ev_cleared, c->stream));
    } else {
        NSMH_CK(cudaEventRecord(c->ev_order, c->stream));              // allocation + earlier readers
        NSMH_CK(cudaStreamWaitEvent(c->copy_stream, c->ev_order, 0));
        NSMH_CK(cudaMemsetAsync(T.slots.p, 0xFF, nslots * sizeof(Slot), c->copy_stream));
        NSMH_CK(cudaEventRecord(c->ev_cleared, c->copy_stream));
    }
    c->precleared_rows = rows;
    c->precleared_ptr = T.slots.p;
    return NSMH_OK;
}

// NSMH_BUILD_LEAVE_BLOCKS: blocks per SM the insert kernel leaves free when the sketch fix-up runs beside it.
// (The insert kernel's blocks fill the register file: without room the fix-up only runs before or after it.  One
// block less - 4 of 5 - and two fix-up blocks per SM was the best of the combinations measured,
// profiles/r2_sketch_build_overlap_s31.txt; the insert kernel's time follows its resident blocks, so the gain is
// 3-4 % of the step, not the 8 % a free overlap would give.)
static int build_leave_blocks() {
    const char *e = getenv("NSMH_BUILD_LEAVE_BLOCKS");
    return e && *e ? std::max(0, atoi(e)) : 1;
}

// defer (nsmh_sketch_build): the sketch entries that are still all-ones are being recomputed on the second stream;
// the insert kernel leaves them out (and leaves that stream room on every SM), table_insert_list_kernel adds them
// once they are there.  The tables are the same set of groups either way.
int build_tables(nsmh_ctx *c, SketchDeferred *defer) {
    if (defer && !defer->pending) defer = nullptr;
    Tables &T = c->tables;
    cudaStream_t s = c->stream;
    const uint32_t n = c->n, rows = c->table_reads;
    T.built = false;
    if (c->precleared_rows) NSMH_CK(cudaStreamWaitEvent(s, c->ev_cleared, 0));   // see preclear_tables
    const uint64_t cap = table_cap(rows);   // even; load factor <= 0.5
    const uint64_t nslots = (uint64_t)n * region_stride(cap);
    const uint64_t items = (uint64_t)rows * n;
    if (nslots >= (1ULL << 32) || items >= (1ULL << 32) - (1ULL << 22))
        return fail(NSMH_EINVAL, "build: reads*n too large for 32-bit slot indices");
    T.cap = cap;
    T.table_reads = rows;
    const size_t ni = (size_t)(items ? items : 1);
    // persistent grid: exactly the blocks that are resident together, so that consecutive work
    // units (same hash functions) really run at the same time and their regions stay in L2
    int occ = 0;
    NSMH_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, table_insert_kernel, kBuildRows, 0));
    const uint64_t units = (uint64_t)((rows + kBuildRows - 1) / kBuildRows) * ((n + kBuildCols - 1) / kBuildCols);
    if (defer && occ > 2) occ = std::max(2, occ - build_leave_blocks());
    const uint32_t blocks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(units, (uint64_t)c->num_sms * (occ > 0 ? occ : 1)));
    const uint32_t seg_cap = kBuildRows * kBuildCols;           // one segment per work unit
    const size_t nseg = (size_t)units * seg_cap;
    NSMH_TRY(T.slots.ensure(nslots * sizeof(Slot), s));
    NSMH_TRY(T.ids.ensure(ni * sizeof(uint32_t), s));
    NSMH_TRY(c->build_multi.ensure(std::max<size_t>(nseg, 1) * 4 * sizeof(uint32_t), s));
    NSMH_TRY(c->build_tmp.ensure((8 + 2 * (size_t)units) * sizeof(unsigned int), s));
    if (!(c->precleared_rows == rows && c->precleared_ptr == T.slots.p))         // else: cleared during nsmh_sketch
        NSMH_CK(cudaMemsetAsync(T.slots.p, 0xFF, nslots * sizeof(Slot), s));
    c->precleared_rows = 0;
    NSMH_CK(cudaMemsetAsync(c->build_tmp.p, 0, 8 * sizeof(unsigned int), s));
    if (items) {
        BuildArgs a;
        a.sk = c->table_sketches;
        a.slots = T.slots.as<Slot>();
        a.ids = T.ids.as<uint32_t>();
        a.m_slot = c->build_multi.as<uint32_t>();
        a.m_id = a.m_slot + nseg;
        a.m_rank = a.m_id + nseg;
        a.g_slot = a.m_rank + nseg;
        a.counters = c->build_tmp.as<unsigned int>();
        a.seg_count = a.counters + 8;
        a.cap = cap;
        a.rows = rows;
        a.n = n;
        a.seg_cap = seg_cap;
        a.segments = (uint32_t)units;
        a.skip_empty = defer ? 1u : 0u;
        table_insert_kernel<<<blocks, kBuildRows, 0, s>>>(a);
        NSMH_CK(cudaGetLastError());
        if (defer) {
            NSMH_TRY(defer->launch(0));
            NSMH_CK(cudaStreamWaitEvent(s, defer->fixed, 0));
            table_insert_list_kernel<<<c->num_sms * 2, 256, 0, s>>>(a, defer->list, defer->count, defer->vals, defer->sk);
            NSMH_CK(cudaGetLastError());
            ++c->launches;
            defer->pending = false;
        }
        table_groups_kernel<<<c->num_sms * 4, 256, 0, s>>>(a);
        NSMH_CK(cudaGetLastError());
        table_fill_kernel<<<c->num_sms * 4, 256, 0, s>>>(a);
        NSMH_CK(cudaGetLastError());
        c->launches += 3;
    }
    T.built = true;
    return NSMH_OK;
}

int table_num_keys(nsmh_ctx *c, uint32_t j, uint32_t *out) {
    Tables &T = c->tables;
    if (!T.built) return fail(NSMH_ESTATE, "table_num_keys: tables not built");
    if (j >= c->n) return fail(NSMH_EINVAL, "table_num_keys: j out of range");
    cudaStream_t s = c->stream;
    NSMH_TRY(c->build_tmp.ensure(64, s));
    uint32_t *d = c->build_tmp.as<uint32_t>() + 4;
    NSMH_CK(cudaMemsetAsync(d, 0, sizeof(uint32_t), s));
    table_count_keys_kernel<<<c->num_sms, 256, 0, s>>>(T.slots.as<Slot>() + (uint64_t)j * region_stride(T.cap),
                                                       T.cap + 1, d);
    ++c->launches;
    NSMH_CK(cudaGetLastError());
    NSMH_CK(cudaMemcpyAsync(out, d, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    NSMH_CK(cudaStreamSynchronize(s));
    return NSMH_OK;
}

} // namespace nsmh

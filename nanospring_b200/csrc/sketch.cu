// MinHash sketching on the device.  Replaces the hot loop of
//   MinHashReadFilter::initialize        (src/ReadFilter.cpp:31-44)
//   MinHashReadFilter::string2Sketch     (src/ReadFilter.cpp:117-131)
//   MinHashReadFilter::string2KMers      (src/ReadFilter.cpp:138-152)
//   MinHashReadFilter::hashKMer          (src/ReadFilter.cpp:133-136)
// Result per read and hash l (bit-exact with the reference):
//   sketch[l] = min over forward k-mers x of (x XOR rand[l])   (u64 compare)
//   len < k-1 -> 0 (row never written by the reference), len == k-1 -> ~0.
//
// Two kernels produce the same numbers:
//
//  * sketch_brute_kernel — the reference's operation count: every k-mer against
//    every hash (n XOR+MIN pairs per k-mer), minima kept in registers, combined
//    with warp shuffles.  This is the kernel the INT32-pipe roofline describes.
//
//  * sketch_filter_kernel (default) — an exact shortcut.  With K = 2k and
//    r = rand[l] & (2^K-1):  x XOR rand[l] = (rand[l] & ~mask) | (x ^ r), so only
//    y = x ^ r matters, and y < 2^(K-b) exactly when the top b bits of x equal
//    the top b bits of r.  If any k-mer of the read matches hash l on those b
//    bits, the minimum is among the matching k-mers.  The top b bits of the
//    k-mer starting at base p are just the b-bit window of the packed stream at
//    p, so the scan is funnel shifts + shared-memory table lookups (one lookup
//    answers three consecutive positions) and touches the 64-bit path only for
//    the ~n*lambda positions per read that hit, lambda = 2^lambda_log2 .. twice
//    that being the expected k-mers per bucket (b = floor(log2 #kmers) - lambda_log2).
//    A warp owns a tile (<= 640 packed words of one read): the words arrive in
//    shared memory through ONE bulk asynchronous copy (cp.async.bulk + mbarrier),
//    phase 1 turns them into one 32-position hit mask per lane and step and appends
//    the hits of all lanes to one list, and phase 2 consumes the list 32 hits at a
//    time with plain shared-memory loads and stores (no atomics).  A (read, hash) pair
//    whose bucket stayed empty (probability ~e^-lambda on random sequence, certain
//    on e.g. homopolymers) is rescanned exhaustively by sketch_fixup_kernel, which
//    makes the result unconditional.
#include <algorithm>
#include <cstdlib>

#include "nsmh_internal.cuh"
#include "sketch_kernels.cuh"
#include "sketch_tables.h"

namespace nsmh {

// ---- host side (the kernels are in sketch_kernels.cuh) -------------------------------
// uploads the lookup tables of the filter kernel (layout: sketch_tables.h)
int build_filter_tables(nsmh_ctx *c) {
    const FilterTables ft = make_filter_tables(c->rand.data(), c->n, c->k);
    const std::vector<uint16_t> &first = ft.first;
    const std::vector<uint8_t> &next = ft.next, &hit3 = ft.hit3;
    const char *e = getenv("NSMH_LAMBDA_LOG2");
    if (e && *e) c->lambda_log2 = atoi(e) < 0 ? 0 : (atoi(e) > 8 ? 8 : atoi(e));
    e = getenv("NSMH_TILE_WORDS");
    if (e && *e && atoi(e) >= 32) c->tile_words = (uint32_t)atoi(e);
    NSMH_TRY(c->d_ftab_first.ensure(first.size() * sizeof(uint16_t), c->stream));
    NSMH_TRY(c->d_ftab_next.ensure(next.size(), c->stream));
    NSMH_TRY(c->d_ftab_hit3.ensure(hit3.size(), c->stream));
    NSMH_CK(cudaMemcpyAsync(c->d_ftab_first.p, first.data(), first.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
    NSMH_CK(cudaMemcpyAsync(c->d_ftab_next.p, next.data(), next.size(), cudaMemcpyHostToDevice, c->stream));
    NSMH_CK(cudaMemcpyAsync(c->d_ftab_hit3.p, hit3.data(), hit3.size(), cudaMemcpyHostToDevice, c->stream));
    NSMH_CK(cudaStreamSynchronize(c->stream));   // host vectors die here
    return NSMH_OK;
}

// Sketches reads [r0, r1) of `rs` (the whole set by default) into rows r0.. of d_sketches.  A range is what
// the pipelined loaders use: the reads of a chunk are sketched while the next chunk is still on the bus.
// The kernels see the range as a read set of its own - offsets stay absolute positions in the one packed
// stream, read indices are relative to r0.
int sketch_reads(nsmh_ctx *c, const ReadSet &rs, uint64_t *d_sketches, DevBuf &tile_start,
                 DevBuf &cub_tmp, int mode, cudaStream_t s, uint32_t *launches, cudaEvent_t ev0,
                 cudaEvent_t ev1, uint32_t r0, uint32_t r1, SketchDeferred *defer) {
    if (r1 > rs.num_reads) r1 = rs.num_reads;
    if (r0 >= r1) return NSMH_OK;
    const uint32_t nr = r1 - r0;
    SketchArgs a;
    a.off = rs.d_offsets() + r0;
    a.W = rs.packed.as<uint32_t>();
    a.sk = d_sketches + (size_t)r0 * c->n;
    a.rnd = c->d_rand.as<uint64_t>();
    a.ftab_first = c->d_ftab_first.as<uint16_t>();
    a.ftab_next = c->d_ftab_next.as<uint8_t>();
    a.ftab_hit3 = c->d_ftab_hit3.as<uint8_t>();
    a.counters = c->counters.as<unsigned long long>();
    a.n_reads = nr;
    a.k = c->k;
    a.n = c->n;
    a.lambda_log2 = c->lambda_log2;
    a.tile_words = (c->tile_words + 63) & ~63u;     // the filter kernel takes 64 words per step
    if (a.tile_words > 4096) a.tile_words = 4096;
    // ... and keeps a tile per warp in shared memory: at least 8 warps per SM must fit
    const size_t smem_budget = 226 * 1024;
    while (a.tile_words > 64 && FilterSmem(c->n, a.tile_words).tab_bytes + 8 * FilterSmem(c->n, a.tile_words).warp_bytes > smem_budget)
        a.tile_words = ((a.tile_words / 2) + 63) & ~63u;

    // tile_start: [0..n_reads] exclusive scan, followed by the per-read counts
    // upper bound on the number of tiles: ceil(words_i / T) <= words_i / T + 1 per read, and the
    // reads' word ranges overlap by at most one word each
    // (a read's word range is widened by at most 4 words: shared boundary word + 16-byte alignment)
    const size_t max_tiles = (size_t)((rs.num_words + 4 * (uint64_t)nr) / a.tile_words) + nr + 1;
    NSMH_TRY(tile_start.ensure((((size_t)nr + 1) * 2 + max_tiles + 2) * sizeof(uint32_t), s));
    uint32_t *cnt = tile_start.as<uint32_t>() + nr + 1;
    uint32_t *ts = tile_start.as<uint32_t>();
    uint32_t *tile_read = cnt + nr + 1;
    a.tile_start = ts;
    a.tile_read = tile_read;
    a.tile_queue = tile_read + max_tiles;      // per call: concurrent online queries have their own
    NSMH_CK(cudaMemsetAsync(cnt + nr, 0, sizeof(uint32_t), s));
    NSMH_CK(cudaMemsetAsync(a.tile_queue, 0, 2 * sizeof(uint32_t), s));      // + the fix-up list length
    const int blocks = (int)std::min<uint64_t>(((uint64_t)nr + 7) / 8, (uint64_t)c->num_sms * 8);
    sketch_init_kernel<<<blocks, 256, 0, s>>>(a, cnt);
    ++*launches;
    NSMH_CK(cudaGetLastError());
    size_t tmp_bytes = 0;
    NSMH_CK(cub_exclusive_sum_u32(nullptr, tmp_bytes, cnt, ts, (size_t)nr + 1, s));
    // the same scratch buffer later holds the fix-up list (u32 per (read, hash) pair at worst)
    const size_t list_off = (tmp_bytes + 255) & ~(size_t)255;
    if ((uint64_t)nr * c->n >= (1ULL << 32))
        return fail(NSMH_EINVAL, "sketch: reads*n too large for 32-bit pair indices");
    NSMH_TRY(cub_tmp.ensure(list_off + (size_t)nr * c->n * sizeof(uint32_t) + 16, s));
    NSMH_CK(cub_exclusive_sum_u32(cub_tmp.p, tmp_bytes, cnt, ts, (size_t)nr + 1, s));
    sketch_tile_map_kernel<<<(nr + 255) / 256, 256, 0, s>>>(ts, nr, tile_read);
    *launches += 3;
    NSMH_CK(cudaGetLastError());

    if (mode == 0 && c->n > 255) mode = 1;   // chain tables index hashes with one byte
    if (ev0) NSMH_CK(cudaEventRecord(ev0, s));
    if (mode == 0) {
        // per-device attribute: set on every call (cheap) rather than once per process
        auto filter_kernel = sketch_filter_kernel;
        NSMH_CK(cudaFuncSetAttribute(filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        // one block per SM, as many warps as the shared memory holds (each warp owns a tile)
        const FilterSmem L(c->n, a.tile_words);
        const size_t budget = smem_budget;
        if (L.tab_bytes + 8 * L.warp_bytes > budget)
            return fail(NSMH_EINVAL, "sketch: n too large for the filter kernel's shared memory");
        int warps = (int)std::min<size_t>(32, (budget - L.tab_bytes) / L.warp_bytes);
        const size_t smem = L.tab_bytes + (size_t)warps * L.warp_bytes;
        filter_kernel<<<c->num_sms, warps * 32, smem, s>>>(a);
        ++*launches;
        NSMH_CK(cudaGetLastError());
        if (ev1) NSMH_CK(cudaEventRecord(ev1, s));
        uint32_t *miss_list = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(cub_tmp.p) + list_off);
        unsigned int *miss_count = a.tile_queue + 1;
        if (defer) {
            // The fix-up runs on the second stream, beside whatever the caller queues next on `s` (the table
            // insert), and leaves its values in a buffer of their own: the sketch matrix keeps all-ones for the
            // listed entries until table_insert_list_kernel stores and inserts them.
            const size_t entries = (size_t)nr * c->n;
            NSMH_TRY(defer->buf.ensure(entries * (sizeof(uint32_t) + sizeof(uint64_t)) + 16, s));
            defer->vals = defer->buf.as<uint64_t>();
            defer->list = reinterpret_cast<uint32_t *>(defer->vals + entries);
            defer->count = miss_count;
            defer->sk = a.sk;
            defer->n = c->n;
            NSMH_CK(cudaEventRecord(defer->filtered, s));
            // queued by build_tables right after its insert kernel, so that the insert's blocks are resident first
            // and these kernels take what is left of every SM (NSMH_FIXUP_BLOCKS: their blocks per SM)
            const char *fe = getenv("NSMH_FIXUP_BLOCKS");
            const int dflt = fe && *fe && atoi(fe) > 0 ? atoi(fe) : 2;         // measured: profiles/r2_sketch_build_overlap_s31.txt
            const int sms = c->num_sms;
            SketchDeferred *d = defer;
            defer->launch = [a, d, miss_count, dflt, sms](int per_sm) -> int {
                const int grid = sms * (per_sm > 0 ? per_sm : dflt);
                NSMH_CK(cudaStreamWaitEvent(d->aux, d->filtered, 0));
                sketch_missing_kernel<<<grid, 256, 0, d->aux>>>(a, d->list, miss_count);
                sketch_fixup_kernel<4><<<grid, 256, 0, d->aux>>>(a, d->list, miss_count, d->vals);
                NSMH_CK(cudaGetLastError());
                NSMH_CK(cudaEventRecord(d->fixed, d->aux));
                return NSMH_OK;
            };
            defer->pending = true;
        } else {
            sketch_missing_kernel<<<c->num_sms * 8, 256, 0, s>>>(a, miss_list, miss_count);
            sketch_fixup_kernel<4><<<c->num_sms * 8, 256, 0, s>>>(a, miss_list, miss_count, nullptr);
            NSMH_CK(cudaGetLastError());
        }
        *launches += 2;
    } else {
        int occ = 0;
        NSMH_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sketch_brute_kernel<8>, 256, 0));
        sketch_brute_kernel<8><<<c->num_sms * (occ > 0 ? occ : 1), 256, 0, s>>>(a);
        ++*launches;
        NSMH_CK(cudaGetLastError());
        if (ev1) NSMH_CK(cudaEventRecord(ev1, s));
    }
    return NSMH_OK;
}

} // namespace nsmh

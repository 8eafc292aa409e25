// Candidate lookup on the device.  Replaces
//   MinHashReadFilter::getFilteredReads(kMer_t sketch[], results)   (src/ReadFilter.cpp:65-83)
//   BBHashMap::pushMatchesInVector                                  (src/BBHashMap.cpp:101-120)
// for a whole batch of query sketches: n exact probes per query, the id lists of
// all probes are gathered, sorted, and an id is emitted when it occurs at least
// overlapSketchThreshold times.  Output per query is ascending and contains the
// query read itself, exactly like the reference's std::sort + upper_bound loop.
//
// count_kernel: one warp per query.  The id lists of a query come from a "source":
//   ProbeSrc  n tables probed on the spot (one 16-byte slot load per probe); the probe
//             results are stored so the second pass does not probe again;
//   PartsSrc  lists that were gathered elsewhere (multi-GPU: every rank probes the
//             tables it owns for ALL queries, nsmh_probe_lists, and ships the lists to
//             the rank that owns the query, nsmh_count_lists).
// A warp scan places the lists in a warp-private shared-memory buffer, a bitonic
// network sorts it, and run lengths are compared against the threshold.  Pass 1
// (COUNT) gives the per-query result count; after one prefix sum pass 2 (EMIT)
// repeats the cheap on-chip part and writes the CSR in place.  Queries whose lists
// exceed the buffer (short reads that all share the all-zero / all-ones sketch,
// SURVEY S5) take the global path: (query, id) pairs, one radix sort, run flags,
// compaction.
#include <algorithm>
#include <cstdlib>

#include "nsmh_internal.cuh"
#include "query_kernels.cuh"
#include "online_kernels.cuh"

namespace nsmh {

constexpr bool kMidTierDefault = true;   // counting-filter tier between the warp sort and the global sort
// count_kernel: query_kernels.cuh (count_body) - one warp per query, one pass
// RL = lists per lane kept in registers: 2 for n <= 64 (fewer registers, more warps per SM), else 4
// CAP = ids per warp-private sort buffer (kLookupCap, or kLookupCapWide with RL = 4)
template <typename Src, int RL, int CAP>
__global__ void __launch_bounds__(kLookupWarps * 32)
count_kernel(Src src, CountArgs a) {
    extern __shared__ __align__(16) uint32_t s_buf[];
    count_body<Src, RL, CAP>(src, a, s_buf);
}

// tmp_ids (completion order) -> CSR (query order); 8 lanes per query
__global__ void __launch_bounds__(256)
csr_place_kernel(CountArgs a, const uint32_t *__restrict__ mid_ids, const uint64_t *__restrict__ out_off,
                 uint32_t *__restrict__ out_ids) {
    const uint32_t sub = threadIdx.x & 7;
    const uint32_t groups = gridDim.x * (blockDim.x >> 3);
    for (uint32_t q = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3); q < a.nq; q += groups) {
        const uint64_t pos = a.qpos[q];
        if (pos == ~0ULL) continue;          // global-path query: heavy_copy_kernel places it
        const uint32_t cnt = a.qcount[q];
        uint32_t *dst = out_ids + out_off[q];
        const uint32_t *src = (pos & kMidPosFlag) ? mid_ids + (pos & ~kMidPosFlag) : a.tmp_ids + pos;
        for (uint32_t i = sub; i < cnt; i += 8) dst[i] = src[i];
    }
}

// The same placement launched BEFORE the host has seen the counters (the default;
// NSMH_LOOKUP_SPECULATE=0 switches it off): a query whose results did not fit tmp_ids, or whose place lies beyond
// out_ids, is skipped - the host then repeats the placement on the general path.
__global__ void __launch_bounds__(256)
csr_place_guarded_kernel(CountArgs a, const uint64_t *__restrict__ out_off, uint32_t *__restrict__ out_ids,
                         uint64_t out_cap) {
    const uint32_t sub = threadIdx.x & 7;
    const uint32_t groups = gridDim.x * (blockDim.x >> 3);
    for (uint32_t q = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3); q < a.nq; q += groups) {
        const uint64_t pos = a.qpos[q];
        if (pos == ~0ULL) continue;
        const uint32_t cnt = a.qcount[q];
        const uint64_t dst0 = out_off[q];
        if (pos + cnt > a.tmp_cap || dst0 + cnt > out_cap) continue;
        for (uint32_t i = sub; i < cnt; i += 8) out_ids[dst0 + i] = a.tmp_ids[pos + i];
    }
}

// ---------------------------------------------------------------- counting-filter tier --
// query_kernels.cuh: a warp per query that overflowed the warp buffer; exact, no global sort.
template <typename Src>
__global__ void __launch_bounds__(kMidWarps * 32)
mid_count_kernel(Src src, MidArgs m) {
    extern __shared__ __align__(16) uint32_t s_mid[];
    const uint32_t warp = threadIdx.x >> 5;
    mid_count_body(src, m, s_mid + (size_t)warp * kMidWarpWords, blockIdx.x * kMidWarps + warp, gridDim.x * kMidWarps);
}

// ---------------------------------------------------------------- global path --
template <typename Src>
__global__ void __launch_bounds__(256)
heavy_counts_kernel(Src src, const uint32_t *__restrict__ heavy_list, uint64_t items, uint32_t *__restrict__ hc) {
    const uint32_t subs = src.subs();
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < items;
         t += (uint64_t)gridDim.x * blockDim.x)
        hc[t] = src.get(heavy_list[t / subs], (uint32_t)(t % subs)).c;
}

// one warp per (heavy query, list): copy the ids as (local heavy index << 32 | id)
template <typename Src>
__global__ void __launch_bounds__(256)
heavy_gather_kernel(Src src, const uint32_t *__restrict__ heavy_list, uint64_t item0, uint64_t items,
                    uint32_t h0, const uint64_t *__restrict__ hoff, uint64_t pair0, uint64_t *__restrict__ pairs) {
    const int lane = threadIdx.x & 31;
    const uint32_t subs = src.subs();
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t t = item0 + blockIdx.x * (uint64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
         t < item0 + items; t += warps) {
        const uint64_t h = t / subs;
        const ListRef r = src.get(heavy_list[h], (uint32_t)(t % subs));
        uint64_t *dst = pairs + (hoff[t] - pair0);
        const uint64_t tag = (h - h0) << 32;
        if (!r.ptr) { if (lane == 0 && r.c) { dst[0] = tag | r.one; if (r.c == 2) dst[1] = tag | r.two; } }
        else for (uint32_t i = lane; i < r.c; i += 32) dst[i] = tag | r.ptr[i];
    }
}

// sorted pairs -> flag the first element of every run of length >= thr
__global__ void __launch_bounds__(256)
flag_runs_kernel(const uint64_t *__restrict__ pairs, uint64_t T, uint32_t thr,
                 const uint32_t *__restrict__ heavy_list, uint32_t h0, uint8_t *__restrict__ flags,
                 uint32_t *__restrict__ qcount) {
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < T;
         p += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = pairs[p];
        bool head = p == 0 || pairs[p - 1] != v;
        bool ok = head && (thr <= 1 || (p + thr - 1 < T && pairs[p + thr - 1] == v));
        flags[p] = ok;
        if (ok) atomicAdd(qcount + heavy_list[h0 + (uint32_t)(v >> 32)], 1u);
    }
}

__global__ void __launch_bounds__(256)
heavy_result_counts_kernel(const uint32_t *__restrict__ heavy_list, uint32_t nh,
                           const uint32_t *__restrict__ qcount, uint32_t *__restrict__ hcnt) {
    for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < nh; h += gridDim.x * blockDim.x)
        hcnt[h] = qcount[heavy_list[h]];
}

// one warp per heavy query: move its results to their place in the CSR
__global__ void __launch_bounds__(256)
heavy_copy_kernel(const uint32_t *__restrict__ heavy_list, uint32_t nh, const uint32_t *__restrict__ hcnt,
                  const uint64_t *__restrict__ hstart, const uint32_t *__restrict__ hout,
                  const uint64_t *__restrict__ out_off, uint32_t *__restrict__ out_ids) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t h = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); h < nh; h += warps) {
        const uint32_t *src = hout + hstart[h];
        uint32_t *dst = out_ids + out_off[heavy_list[h]];
        for (uint32_t r = lane; r < hcnt[h]; r += 32) dst[r] = src[r];
    }
}

// NSMH_MID_TIER=0 sends every query that overflows the warp buffer to the global sort (A/B runs, tests)
static bool mid_tier_enabled() {
    const char *e = getenv("NSMH_MID_TIER");
    return e && *e ? atoi(e) != 0 : kMidTierDefault;
}

// NSMH_LOOKUP_SPECULATE=0 switches the speculative placement off (see count_and_emit; A/B runs, tests)
static bool speculate_enabled() {
    const char *e = getenv("NSMH_LOOKUP_SPECULATE");
    return e && *e ? atoi(e) != 0 : true;
}

// NSMH_LOOKUP_WIDE=0: the 1024-id sort buffer also for n > 64 (A/B runs)
static bool wide_buffer_enabled() {
    const char *e = getenv("NSMH_LOOKUP_WIDE");
    return e && *e ? atoi(e) != 0 : true;
}

static int grid_for(uint64_t items, int sms, int per_block = 256) {
    uint64_t b = (items + per_block - 1) / per_block;
    uint64_t cap = (uint64_t)sms * 16;
    return (int)(b < cap ? (b ? b : 1) : cap);
}

// Queries that overflowed the warp buffer: global (heavy index, id) pair sort in batches.
template <typename Src>
static int heavy_path(nsmh_ctx *c, QueryWs &ws, const Src &src, uint32_t subs, const uint32_t *hl, uint32_t nh,
                      cudaStream_t s) {
    const uint64_t items = (uint64_t)nh * subs;
    NSMH_TRY(ws.hc.ensure((items + 1) * sizeof(uint32_t), s));
    NSMH_TRY(ws.hoff.ensure((items + 1) * sizeof(uint64_t), s));
    NSMH_CK(cudaMemsetAsync(ws.hc.as<uint32_t>() + items, 0, sizeof(uint32_t), s));
    heavy_counts_kernel<Src><<<grid_for(items, c->num_sms), 256, 0, s>>>(src, hl, items, ws.hc.as<uint32_t>());
    NSMH_CK(cudaGetLastError());
    size_t tmp_bytes = 0;
    NSMH_CK(cub_exclusive_sum_u32_to_u64(nullptr, tmp_bytes, ws.hc.as<uint32_t>(), ws.hoff.as<uint64_t>(), items + 1, s));
    NSMH_TRY(ws.cub_tmp.ensure(tmp_bytes, s));
    NSMH_CK(cub_exclusive_sum_u32_to_u64(ws.cub_tmp.p, tmp_bytes, ws.hc.as<uint32_t>(), ws.hoff.as<uint64_t>(), items + 1, s));
    ws.launches += 3;
    // pair offset of every heavy query, on the host, to cut batches that fit the scratch budget
    std::vector<uint64_t> hstart_pairs((size_t)nh + 1);
    NSMH_CK(cudaMemcpy2DAsync(hstart_pairs.data(), sizeof(uint64_t), ws.hoff.p, (size_t)subs * sizeof(uint64_t),
                              sizeof(uint64_t), nh, cudaMemcpyDeviceToHost, s));
    NSMH_CK(cudaMemcpyAsync(&hstart_pairs[nh], ws.hoff.as<uint64_t>() + items, sizeof(uint64_t),
                            cudaMemcpyDeviceToHost, s));
    NSMH_CK(cudaStreamSynchronize(s));
    size_t free_b = 0, total_b = 0;
    NSMH_CK(cudaMemGetInfo(&free_b, &total_b));
    uint64_t budget = (uint64_t)((free_b + ws.pairs.cap + ws.pairs_alt.cap + ws.flags.cap) * 0.6 / 21.0);
    if (budget < (1u << 20)) budget = 1u << 20;
    uint64_t hout_total = 0;
    NSMH_TRY(ws.nsel.ensure(4 * sizeof(uint64_t), s));
    for (uint32_t h0 = 0; h0 < nh;) {
        uint32_t h1 = h0 + 1;   // at least one query per batch
        while (h1 < nh && hstart_pairs[h1 + 1] - hstart_pairs[h0] <= budget) ++h1;
        const uint64_t pair0 = hstart_pairs[h0], Tn = hstart_pairs[h1] - pair0;
        if (Tn) {
            NSMH_TRY(ws.pairs.ensure(Tn * sizeof(uint64_t), s));
            NSMH_TRY(ws.pairs_alt.ensure(Tn * sizeof(uint64_t), s));
            NSMH_TRY(ws.flags.ensure(Tn, s));
            const uint64_t bitems = (uint64_t)(h1 - h0) * subs;
            heavy_gather_kernel<Src><<<grid_for(bitems, c->num_sms, 8), 256, 0, s>>>(
                src, hl, (uint64_t)h0 * subs, bitems, h0, ws.hoff.as<uint64_t>(), pair0, ws.pairs.as<uint64_t>());
            NSMH_CK(cudaGetLastError());
            int hbits = 1;
            while ((1ULL << hbits) < (uint64_t)(h1 - h0)) ++hbits;
            bool in_alt = false;
            tmp_bytes = 0;
            NSMH_CK(cub_sort_keys_u64(nullptr, tmp_bytes, ws.pairs.as<uint64_t>(), ws.pairs_alt.as<uint64_t>(),
                                      Tn, 0, 32 + hbits, in_alt, s));
            NSMH_TRY(ws.cub_tmp.ensure(tmp_bytes, s));
            NSMH_CK(cub_sort_keys_u64(ws.cub_tmp.p, tmp_bytes, ws.pairs.as<uint64_t>(),
                                      ws.pairs_alt.as<uint64_t>(), Tn, 0, 32 + hbits, in_alt, s));
            const uint64_t *sorted = in_alt ? ws.pairs_alt.as<uint64_t>() : ws.pairs.as<uint64_t>();
            flag_runs_kernel<<<grid_for(Tn, c->num_sms), 256, 0, s>>>(sorted, Tn, c->thr, hl, h0,
                                                                      ws.flags.as<uint8_t>(), ws.qcount.as<uint32_t>());
            NSMH_CK(cudaGetLastError());
            NSMH_TRY(ws.hout.ensure((hout_total + Tn) * sizeof(uint32_t), s, hout_total * sizeof(uint32_t)));
            tmp_bytes = 0;
            NSMH_CK(cub_select_low32_flagged(nullptr, tmp_bytes, sorted, ws.flags.as<uint8_t>(),
                                             ws.hout.as<uint32_t>() + hout_total, ws.nsel.as<uint64_t>(), Tn, s));
            NSMH_TRY(ws.cub_tmp.ensure(tmp_bytes, s));
            NSMH_CK(cub_select_low32_flagged(ws.cub_tmp.p, tmp_bytes, sorted, ws.flags.as<uint8_t>(),
                                             ws.hout.as<uint32_t>() + hout_total, ws.nsel.as<uint64_t>(), Tn, s));
            ws.launches += 6 + (32 + hbits + 7) / 8;
            uint64_t nsel = 0;
            NSMH_CK(cudaMemcpyAsync(&nsel, ws.nsel.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
            NSMH_CK(cudaStreamSynchronize(s));
            hout_total += nsel;
        }
        h0 = h1;
    }
    // where every heavy query's results start in hout
    NSMH_TRY(ws.hcnt.ensure(((size_t)nh + 1) * sizeof(uint32_t), s));
    NSMH_TRY(ws.hstart.ensure(((size_t)nh + 1) * sizeof(uint64_t), s));
    NSMH_CK(cudaMemsetAsync(ws.hcnt.as<uint32_t>() + nh, 0, sizeof(uint32_t), s));
    heavy_result_counts_kernel<<<grid_for(nh, c->num_sms), 256, 0, s>>>(hl, nh, ws.qcount.as<uint32_t>(),
                                                                        ws.hcnt.as<uint32_t>());
    NSMH_CK(cudaGetLastError());
    tmp_bytes = 0;
    NSMH_CK(cub_exclusive_sum_u32_to_u64(nullptr, tmp_bytes, ws.hcnt.as<uint32_t>(), ws.hstart.as<uint64_t>(), (size_t)nh + 1, s));
    NSMH_TRY(ws.cub_tmp.ensure(tmp_bytes, s));
    NSMH_CK(cub_exclusive_sum_u32_to_u64(ws.cub_tmp.p, tmp_bytes, ws.hcnt.as<uint32_t>(), ws.hstart.as<uint64_t>(), (size_t)nh + 1, s));
    ws.launches += 3;
    return NSMH_OK;
}

// one counting pass, prefix sum, placement for nq queries whose lists come from `src`.
// Result CSR in ws.out_off (u64 [nq+1]) / ws.out_ids (u32 [ws.last_total]).  Synchronises `s`.
template <typename Src>
static int count_and_emit(nsmh_ctx *c, QueryWs &ws, const Src &src, uint32_t subs, uint32_t nq, cudaStream_t s) {
    ws.last_nq = nq;
    ws.last_total = 0;
    ws.last_pairs = 0;
    ws.last_heavy = ws.last_sorted = 0;
    NSMH_TRY(ws.out_off.ensure(((size_t)nq + 1) * sizeof(uint64_t), s));
    if (nq == 0) {
        NSMH_CK(cudaMemsetAsync(ws.out_off.p, 0, sizeof(uint64_t), s));
        NSMH_CK(cudaStreamSynchronize(s));
        return NSMH_OK;
    }
    NSMH_TRY(ws.qcount.ensure(((size_t)nq + 1) * sizeof(uint32_t), s));
    NSMH_TRY(ws.qpos.ensure((size_t)nq * sizeof(uint64_t), s));
    NSMH_TRY(ws.heavy_list.ensure((size_t)nq * sizeof(uint32_t), s));
    NSMH_TRY(ws.counters.ensure(8 * sizeof(uint64_t), s));
    // results land in tmp_ids: kFixedIds per query at a fixed place, larger result lists behind them in
    // completion order; the size of that part is a guess that the kernel checks (the cursor keeps counting
    // past the end), so at most one repeat with the exact size
    const size_t fixed_ids = (size_t)nq * kFixedIds;
    if (ws.tmp_ids.cap < (fixed_ids + (size_t)nq * 8) * sizeof(uint32_t))
        NSMH_TRY(ws.tmp_ids.ensure((fixed_ids + (size_t)nq * 8) * sizeof(uint32_t), s));

    const bool wide = subs > 64 && wide_buffer_enabled();
    const size_t smem = (size_t)kLookupWarps * warp_words(wide ? kLookupCapWide : kLookupCap) * sizeof(uint32_t);
    // function attributes are per device: set on every call (cheap) rather than once per process
    auto kernel = subs <= 64 ? count_kernel<Src, 2, kLookupCap>
                  : wide     ? count_kernel<Src, kRegListsMax, kLookupCapWide>
                             : count_kernel<Src, kRegListsMax, kLookupCap>;
    NSMH_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CountArgs a;
    a.qcount = ws.qcount.as<uint32_t>();
    a.qpos = ws.qpos.as<uint64_t>();
    a.heavy_list = ws.heavy_list.as<uint32_t>();
    a.counters = ws.counters.as<unsigned long long>();
    a.nq = nq;
    a.thr = c->thr ? c->thr : 1;    // thr 0 and 1 both emit every gathered id once (ReadFilter.cpp:76-82)
    int occ = 0;
    NSMH_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kLookupWarps * 32, smem));
    const int blocks = (int)std::min<uint64_t>(((uint64_t)nq + kLookupWarps - 1) / kLookupWarps,
                                               (uint64_t)c->num_sms * (occ > 0 ? occ : 1));
    unsigned long long cnt[4] = {0, 0, 0, 0};
    // Prefix sum and placement are queued right behind the counting kernel (measured on B200: 0.235 ->
    // 0.221 ms for the benchmark's lookup; NSMH_LOOKUP_SPECULATE=0 switches it off), before the host knows whether a query overflowed; when none did and the results fit
    // (the common case) the single read-back below is the only host round trip of the lookup.
    const bool speculate = speculate_enabled();
    size_t spec_tmp_bytes = 0;
    if (speculate) {
        NSMH_TRY(ws.out_ids.ensure(ws.tmp_ids.cap, s));
        NSMH_CK(cub_exclusive_sum_u32_to_u64(nullptr, spec_tmp_bytes, ws.qcount.as<uint32_t>(), ws.out_off.as<uint64_t>(), (size_t)nq + 1, s));
        NSMH_TRY(ws.cub_tmp.ensure(spec_tmp_bytes, s));
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        a.tmp_ids = ws.tmp_ids.as<uint32_t>();
        a.tmp_cap = ws.tmp_ids.cap / sizeof(uint32_t);
        NSMH_CK(cudaMemsetAsync(ws.counters.p, 0, 8 * sizeof(uint64_t), s));
        NSMH_CK(cudaMemsetAsync(ws.qcount.as<uint32_t>() + nq, 0, sizeof(uint32_t), s));
        kernel<<<blocks, kLookupWarps * 32, smem, s>>>(src, a);
        ++ws.launches;
        NSMH_CK(cudaGetLastError());
        uint64_t spec_total = 0;
        if (speculate && attempt == 0) {
            NSMH_CK(cub_exclusive_sum_u32_to_u64(ws.cub_tmp.p, spec_tmp_bytes, ws.qcount.as<uint32_t>(), ws.out_off.as<uint64_t>(), (size_t)nq + 1, s));
            csr_place_guarded_kernel<<<grid_for((uint64_t)nq * 8, c->num_sms), 256, 0, s>>>(
                a, ws.out_off.as<uint64_t>(), ws.out_ids.as<uint32_t>(), ws.out_ids.cap / sizeof(uint32_t));
            NSMH_CK(cudaGetLastError());
            ws.launches += 3;
            NSMH_CK(cudaMemcpyAsync(&spec_total, ws.out_off.as<uint64_t>() + nq, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        }
        NSMH_CK(cudaMemcpyAsync(cnt, ws.counters.p, sizeof cnt, cudaMemcpyDeviceToHost, s));
        NSMH_CK(cudaStreamSynchronize(s));
        if (speculate && attempt == 0 && cnt[0] == 0 && fixed_ids + cnt[3] <= a.tmp_cap &&
            spec_total <= ws.out_ids.cap / sizeof(uint32_t)) {
            ws.last_pairs = cnt[1];
            ws.last_total = spec_total;          // == cnt[2]: nothing was skipped by the guards
            return NSMH_OK;
        }
        if (fixed_ids + cnt[3] <= a.tmp_cap) break;
        NSMH_TRY(ws.tmp_ids.ensure((fixed_ids + (size_t)cnt[3]) * sizeof(uint32_t), s));
    }
    uint32_t nh = (uint32_t)cnt[0];
    const uint32_t *hl = ws.heavy_list.as<uint32_t>();
    ws.last_pairs = cnt[1];
    ws.last_heavy = nh;
    // output size: the results counted so far are exact (cnt[2]); a query handed on to the next tiers emits
    // at most (its gathered ids) / thr, and those ids are part of cnt[1]
    NSMH_TRY(ws.out_ids.ensure((size_t)(cnt[2] + cnt[1] / a.thr + 1) * sizeof(uint32_t), s));
    if (nh && mid_tier_enabled()) {
        // counting-filter tier: resolves the heavy queries that have few ids above the threshold
        // (query_kernels.cuh); what it cannot resolve goes on to the global sort.  A result needs thr
        // gathered ids, so cnt[1] / thr bounds what all queries together can emit.
        MidArgs m;
        m.mid_cap = cnt[1] / a.thr + 1;
        NSMH_TRY(ws.mid_ids.ensure((size_t)m.mid_cap * sizeof(uint32_t), s));
        NSMH_TRY(ws.heavy2_list.ensure((size_t)nh * sizeof(uint32_t), s));
        m.heavy_list = hl;
        m.unresolved_list = ws.heavy2_list.as<uint32_t>();
        m.mid_ids = ws.mid_ids.as<uint32_t>();
        m.qcount = a.qcount;
        m.qpos = a.qpos;
        m.counters = a.counters + 4;
        m.nh = nh;
        m.thr = a.thr;
        const size_t mid_smem = (size_t)kMidWarps * kMidWarpWords * sizeof(uint32_t);
        NSMH_CK(cudaFuncSetAttribute(mid_count_kernel<Src>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mid_smem));
        int mocc = 0;
        NSMH_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&mocc, mid_count_kernel<Src>, kMidWarps * 32, mid_smem));
        const int mblocks = (int)std::min<uint64_t>(((uint64_t)nh + kMidWarps - 1) / kMidWarps,
                                                    (uint64_t)c->num_sms * (mocc > 0 ? mocc : 1));
        mid_count_kernel<Src><<<mblocks, kMidWarps * 32, mid_smem, s>>>(src, m);
        ++ws.launches;
        NSMH_CK(cudaGetLastError());
        unsigned long long mc[3] = {0, 0, 0};
        NSMH_CK(cudaMemcpyAsync(mc, m.counters, sizeof mc, cudaMemcpyDeviceToHost, s));
        NSMH_CK(cudaStreamSynchronize(s));
        if (mc[2] || mc[0] > nh) return fail(NSMH_ECUDA, "query: internal error in the counting-filter tier");
        nh = (uint32_t)mc[0];
        hl = ws.heavy2_list.as<uint32_t>();
    }
    ws.last_sorted = nh;
    if (nh) NSMH_TRY(heavy_path(c, ws, src, subs, hl, nh, s));

    size_t tmp_bytes = 0;
    NSMH_CK(cub_exclusive_sum_u32_to_u64(nullptr, tmp_bytes, ws.qcount.as<uint32_t>(), ws.out_off.as<uint64_t>(), (size_t)nq + 1, s));
    NSMH_TRY(ws.cub_tmp.ensure(tmp_bytes, s));
    NSMH_CK(cub_exclusive_sum_u32_to_u64(ws.cub_tmp.p, tmp_bytes, ws.qcount.as<uint32_t>(), ws.out_off.as<uint64_t>(), (size_t)nq + 1, s));
    csr_place_kernel<<<grid_for((uint64_t)nq * 8, c->num_sms), 256, 0, s>>>(a, ws.mid_ids.as<uint32_t>(), ws.out_off.as<uint64_t>(),
                                                                          ws.out_ids.as<uint32_t>());
    ws.launches += 3;
    NSMH_CK(cudaGetLastError());
    if (nh) {
        heavy_copy_kernel<<<grid_for(nh, c->num_sms, 8), 256, 0, s>>>(
            hl, nh, ws.hcnt.as<uint32_t>(), ws.hstart.as<uint64_t>(),
            ws.hout.as<uint32_t>(), ws.out_off.as<uint64_t>(), ws.out_ids.as<uint32_t>());
        ++ws.launches;
        NSMH_CK(cudaGetLastError());
    }
    uint64_t total = 0;
    NSMH_CK(cudaMemcpyAsync(&total, ws.out_off.as<uint64_t>() + nq, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    NSMH_CK(cudaStreamSynchronize(s));
    ws.last_total = total;
    return NSMH_OK;
}

// probe the n tables for nq device-resident sketches [nq][n]; results in ws.pval / ws.pcnt
static int probe_all(nsmh_ctx *c, QueryWs &ws, const uint64_t *d_qsketch, uint32_t nq, cudaStream_t s,
                     StoredSrc &stored) {
    Tables &T = c->tables;
    if (!T.built) return fail(NSMH_ESTATE, "query: tables not built (call nsmh_build)");
    const uint64_t items = (uint64_t)nq * c->n;
    if (items >= (1ULL << 32)) return fail(NSMH_EINVAL, "query: queries*n too large for 32-bit item indices");
    NSMH_TRY(ws.pval.ensure(std::max<uint64_t>(items, 1) * sizeof(uint32_t), s));
    NSMH_TRY(ws.pcnt.ensure(std::max<uint64_t>(items, 1) * sizeof(uint32_t), s));
    ProbeSrc src;
    src.qsk = d_qsketch;
    src.slots = T.slots.as<Slot>();
    src.ids = T.ids.as<uint32_t>();
    src.pval = ws.pval.as<uint32_t>();
    src.pcnt = ws.pcnt.as<uint32_t>();
    src.cap = T.cap;
    src.n = c->n;
    if (items) {
        const uint64_t units = (uint64_t)((nq + kProbeRows - 1) / kProbeRows) * ((c->n + kProbeCols - 1) / kProbeCols);
        // persistent grid = the blocks that are resident together (L2 locality of the regions)
        int occ = 0;
        NSMH_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe_items_kernel, kProbeRows, 0));
        const int blocks = (int)std::min<uint64_t>(units, (uint64_t)c->num_sms * (occ > 0 ? occ : 1));
        probe_items_kernel<<<blocks, kProbeRows, 0, s>>>(src, nq);
        ++ws.launches;
        NSMH_CK(cudaGetLastError());
    }
    stored.pval = src.pval;
    stored.pcnt = src.pcnt;
    stored.ids = src.ids;
    stored.n = c->n;
    return NSMH_OK;
}

// Query nq device-resident sketches [nq][n] against the tables: probing is fused into the
// counting kernel (no probe results in HBM).
int query_sketches_device(nsmh_ctx *c, QueryWs &ws, const uint64_t *d_qsketch, uint32_t nq, cudaStream_t s) {
    Tables &T = c->tables;
    if (!T.built) return fail(NSMH_ESTATE, "query: tables not built (call nsmh_build)");
    const char *e_split = getenv("NSMH_LOOKUP_SPLIT");
    if (e_split && atoi(e_split) != 0) {
        // probe kernel (thread per query x 4 hash functions, ordered by hash function so the region being
        // probed is L2 resident) + counting kernel over the stored results
        StoredSrc stored;
        NSMH_TRY(probe_all(c, ws, d_qsketch, nq, s, stored));
        return count_and_emit(c, ws, stored, c->n, nq, s);
    }
    ProbeSrc src;
    src.qsk = d_qsketch;
    src.slots = T.slots.as<Slot>();
    src.ids = T.ids.as<uint32_t>();
    src.pval = nullptr;
    src.pcnt = nullptr;
    src.cap = T.cap;
    src.n = c->n;
    return count_and_emit(c, ws, src, c->n, nq, s);
}

// Threshold the union of `parts` partial id lists per query (device CSR pieces).
int count_lists_device(nsmh_ctx *c, QueryWs &ws, uint32_t nq, uint32_t parts, const uint64_t *const *d_offsets,
                       const uint32_t *const *d_ids, cudaStream_t s) {
    if (parts == 0 || parts > (uint32_t)kMaxParts) return fail(NSMH_EINVAL, "count_lists: parts must be in 1..16");
    PartsSrc src;
    src.parts = parts;
    for (uint32_t p = 0; p < (uint32_t)kMaxParts; ++p) {
        src.offs[p] = p < parts ? d_offsets[p] : nullptr;
        src.ids[p] = p < parts ? d_ids[p] : nullptr;
    }
    return count_and_emit(c, ws, src, parts, nq, s);
}

// ---------------------------------------------------------------- probe only --
// Multi-GPU building block: probe the n tables for every query and gather the id lists,
// no counting: CSR of concatenated lists in ws.out_off / ws.out_ids.
__global__ void __launch_bounds__(256)
probe_totals_kernel(StoredSrc src, uint32_t nq, uint32_t *__restrict__ qcount) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < nq; q += warps) {
        uint32_t tot = 0;
        for (uint32_t j = lane; j < src.n; j += 32) tot += src.pcnt[(size_t)q * src.n + j];
#pragma unroll
        for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) qcount[q] = tot;
    }
}

__global__ void __launch_bounds__(256)
probe_write_kernel(StoredSrc src, uint32_t nq, const uint64_t *__restrict__ out_off, uint32_t *__restrict__ out_ids) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < nq; q += warps) {
        uint64_t base = out_off[q];
        for (uint32_t j0 = 0; j0 < src.n; j0 += 32) {
            const uint32_t j = j0 + lane;
            const ListRef r = j < src.n ? src.get(q, j) : empty_list();
            const uint32_t incl = warp_incl_scan(r.c, lane);
            uint32_t *dst = out_ids + base + (incl - r.c);
            if (r.c == 1 && !r.ptr) dst[0] = r.one;
            else for (uint32_t i = 0; i < r.c; ++i) dst[i] = r.ptr[i];
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

// ---------------------------------------------------------------- multi-GPU --
int probe_to_peers_device(nsmh_ctx *sub, const uint64_t *d_qsketch, uint32_t nq, const PeerDst &dst,
                          cudaStream_t s, uint32_t *launches) {
    Tables &T = sub->tables;
    if (!T.built) return fail(NSMH_ESTATE, "mg: owned tables not built");
    if ((uint64_t)nq * sub->n >= (1ULL << 32)) return fail(NSMH_EINVAL, "mg: reads*n too large for 32-bit item indices");
    if (nq == 0) return NSMH_OK;
    ProbeSrc src;
    src.qsk = d_qsketch;
    src.slots = T.slots.as<Slot>();
    src.ids = T.ids.as<uint32_t>();
    src.pval = nullptr;
    src.pcnt = nullptr;
    src.cap = T.cap;
    src.n = sub->n;
    const uint64_t units = (uint64_t)((nq + kProbeRows - 1) / kProbeRows) * ((sub->n + kPeerCols - 1) / kPeerCols);
    int occ = 0;
    NSMH_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe_to_peers_kernel, kProbeRows, 0));
    const int blocks = (int)std::min<uint64_t>(units, (uint64_t)sub->num_sms * (occ > 0 ? occ : 1));
    probe_to_peers_kernel<<<blocks, kProbeRows, 0, s>>>(src, nq, dst);
    ++*launches;
    NSMH_CK(cudaGetLastError());
    return NSMH_OK;
}

int count_peer_lists_device(nsmh_ctx *c, QueryWs &ws, const PeerLists &lists, uint32_t nq, cudaStream_t s) {
    PeerSrc src;
    src.L = lists;
    return count_and_emit(c, ws, src, lists.n, nq, s);
}

int probe_lists_device(nsmh_ctx *c, QueryWs &ws, const uint64_t *d_qsketch, uint32_t nq, cudaStream_t s) {
    StoredSrc src;
    NSMH_TRY(probe_all(c, ws, d_qsketch, nq, s, src));
    ws.last_nq = nq;
    ws.last_total = 0;
    NSMH_TRY(ws.out_off.ensure(((size_t)nq + 1) * sizeof(uint64_t), s));
    NSMH_TRY(ws.qcount.ensure(((size_t)nq + 1) * sizeof(uint32_t), s));
    NSMH_CK(cudaMemsetAsync(ws.qcount.as<uint32_t>() + nq, 0, sizeof(uint32_t), s));
    if (nq) {
        probe_totals_kernel<<<grid_for(nq, c->num_sms, 8), 256, 0, s>>>(src, nq, ws.qcount.as<uint32_t>());
        NSMH_CK(cudaGetLastError());
    }
    size_t tmp_bytes = 0;
    NSMH_CK(cub_exclusive_sum_u32_to_u64(nullptr, tmp_bytes, ws.qcount.as<uint32_t>(), ws.out_off.as<uint64_t>(), (size_t)nq + 1, s));
    NSMH_TRY(ws.cub_tmp.ensure(tmp_bytes, s));
    NSMH_CK(cub_exclusive_sum_u32_to_u64(ws.cub_tmp.p, tmp_bytes, ws.qcount.as<uint32_t>(), ws.out_off.as<uint64_t>(), (size_t)nq + 1, s));
    uint64_t total = 0;
    NSMH_CK(cudaMemcpyAsync(&total, ws.out_off.as<uint64_t>() + nq, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    NSMH_CK(cudaStreamSynchronize(s));
    NSMH_TRY(ws.out_ids.ensure(std::max<uint64_t>(total, 1) * sizeof(uint32_t), s));
    if (nq) {
        probe_write_kernel<<<grid_for(nq, c->num_sms, 8), 256, 0, s>>>(src, nq, ws.out_off.as<uint64_t>(),
                                                                        ws.out_ids.as<uint32_t>());
        NSMH_CK(cudaGetLastError());
    }
    ws.launches += 4;
    NSMH_CK(cudaStreamSynchronize(s));
    ws.last_total = total;
    ws.last_pairs = total;
    return NSMH_OK;
}

// The online fast path: ONE kernel reads the strings from mapped host memory, sketches, probes, counts and
// writes the answer back to mapped host memory (online_kernels.cuh); the host waits for the stream once.
// Returns 1 when the general path must take over (long strings, heavy id lists, results beyond the buffer).
static constexpr uint32_t kOnlineMaxQueries = 8;
static bool online_enabled() {
    static const bool on = [] {
        const char *e = getenv("NSMH_ONLINE_FUSED");
        return !(e && *e && atoi(e) == 0);
    }();
    return on;
}

int online_query(nsmh_ctx *c, QueryWs &ws, const char *bases, const uint64_t *offsets, uint32_t nq,
                        uint64_t *offsets_out, uint32_t *ids_out, size_t cap) {
    if (!online_enabled() || nq == 0 || nq > kOnlineMaxQueries || c->n > 32 * (uint32_t)kRegListsMax || c->sketch_mode != 0) return 1;
    uint64_t max_len = 0, text_bytes = 0;
    for (uint32_t i = 0; i < nq; ++i) {
        const uint64_t len = offsets[i + 1] - offsets[i];
        max_len = std::max(max_len, len);
        text_bytes += (len + 15 & ~15ULL) + 16;
    }
    if (max_len > kOnlineMaxBases) return 1;
    cudaStream_t s = ws.stream;
    // mapped block: [start nq][len nq][qpos nq] u64 | [qcount nq] u32 | ids [nq][kOnlineIdsPerQuery] u32 | text
    const size_t head = (size_t)kOnlineMaxQueries * (3 * sizeof(uint64_t) + sizeof(uint32_t));
    const size_t ids_off = (head + 15) & ~(size_t)15;
    const size_t text_off = ids_off + (size_t)kOnlineMaxQueries * kOnlineIdsPerQuery * sizeof(uint32_t);
    const size_t need = text_off + text_bytes + 64;
    if (need > ws.h_online_cap) {
        if (ws.h_online) cudaFreeHost(ws.h_online);
        ws.h_online = nullptr;
        ws.h_online_cap = 0;
        const size_t want = need + need / 2 + (64 << 10);
        NSMH_CK(cudaHostAlloc(reinterpret_cast<void **>(&ws.h_online), want, cudaHostAllocMapped | cudaHostAllocPortable));
        ws.h_online_cap = want;
    }
    NSMH_TRY(ws.online_scratch.ensure((size_t)kOnlineMaxQueries * (8 * sizeof(uint64_t) + sizeof(uint32_t)), s));
    uint64_t *h_start = reinterpret_cast<uint64_t *>(ws.h_online), *h_len = h_start + kOnlineMaxQueries,
             *h_qpos = h_len + kOnlineMaxQueries;
    uint32_t *h_qcount = reinterpret_cast<uint32_t *>(h_qpos + kOnlineMaxQueries);
    uint32_t *h_ids = reinterpret_cast<uint32_t *>(ws.h_online + ids_off);
    uint8_t *h_text = ws.h_online + text_off;
    uint64_t at = 0;
    for (uint32_t i = 0; i < nq; ++i) {
        const uint64_t len = offsets[i + 1] - offsets[i];
        h_start[i] = at;
        h_len[i] = len;
        h_qpos[i] = ~0ULL;
        h_qcount[i] = 0;
        if (len) memcpy(h_text + at, bases + offsets[i], len);
        at += (len + 15 & ~15ULL) + 16;
    }
    uint8_t *d_base = nullptr;
    NSMH_CK(cudaHostGetDevicePointer(reinterpret_cast<void **>(&d_base), ws.h_online, 0));
    OnlineArgs a;
    a.text = d_base + text_off;
    a.start = reinterpret_cast<const uint64_t *>(d_base);
    a.len = a.start + kOnlineMaxQueries;
    a.qpos = reinterpret_cast<uint64_t *>(d_base) + 2 * kOnlineMaxQueries;
    a.qcount = reinterpret_cast<uint32_t *>(a.qpos + kOnlineMaxQueries);
    a.tmp_ids = reinterpret_cast<uint32_t *>(d_base + ids_off);
    a.counters = ws.online_scratch.as<unsigned long long>();
    a.heavy_scratch = reinterpret_cast<uint32_t *>(a.counters + 8 * kOnlineMaxQueries);
    a.rnd = c->d_rand.as<uint64_t>();
    a.sketch_out = nullptr;
    a.k = c->k;
    a.thr = c->thr ? c->thr : 1;
    Tables &T = c->tables;
    a.tables.qsk = nullptr;
    a.tables.slots = T.slots.as<Slot>();
    a.tables.ids = T.ids.as<uint32_t>();
    a.tables.pval = a.tables.pcnt = nullptr;
    a.tables.cap = T.cap;
    a.tables.n = c->n;
    const uint32_t ml = (uint32_t)((max_len + 1023) & ~1023ULL);
    const size_t smem = online_smem_bytes(ml, c->n);
    if (smem > 48 * 1024) NSMH_CK(cudaFuncSetAttribute(online_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    online_query_kernel<<<nq, kOnlineThreads, smem, s>>>(a, ml);
    ++ws.launches;
    NSMH_CK(cudaGetLastError());
    NSMH_CK(cudaStreamSynchronize(s));
    uint64_t total = 0;
    for (uint32_t i = 0; i < nq; ++i) {
        if (h_qpos[i] == ~0ULL || h_qpos[i] + h_qcount[i] > kOnlineIdsPerQuery) return 1;      // heavy tiers / too many results
        total += h_qcount[i];
    }
    uint64_t o = 0;
    for (uint32_t i = 0; i < nq; ++i) {
        if (offsets_out) offsets_out[i] = o;
        const uint32_t *src = h_ids + (size_t)i * kOnlineIdsPerQuery + h_qpos[i];
        for (uint32_t r = 0; r < h_qcount[i]; ++r, ++o)
            if (ids_out && o < cap) ids_out[o] = src[r];
    }
    if (offsets_out) offsets_out[nq] = o;
    ws.last_nq = nq;
    ws.last_total = total;
    if (total > cap) {
        char buf[128];
        snprintf(buf, sizeof buf, "query: output needs %llu ids, capacity is %zu", (unsigned long long)total, cap);
        return fail(NSMH_ERANGE, buf);
    }
    return NSMH_OK;
}


} // namespace nsmh

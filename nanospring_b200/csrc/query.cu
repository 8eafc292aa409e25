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

#include "nsmh_internal.cuh"

namespace nsmh {

constexpr int kLookupCap = 2048;     // ids per warp-private buffer
constexpr int kLookupWarps = 8;
constexpr int kMaxParts = 16;        // PartsSrc: at most this many partial lists per query

__device__ __forceinline__ uint64_t slot_hash_q(uint64_t key, uint32_t log2cap) {
    return (key * 0x9E3779B97F4A7C15ULL) >> (64 - log2cap);
}

// exact probe of table l: group size (0 = absent) and val (the id itself for a
// group of one, else the start of the group in ids)
__device__ __forceinline__ uint32_t probe_slot(const Slot *__restrict__ slots, uint64_t cap,
                                               uint32_t log2cap, uint32_t l, uint64_t key, uint32_t &val) {
    const Slot *region = slots + (uint64_t)l * (cap + 1);
    if (key == kEmptyKey) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(region + cap));
        val = v.z;
        return v.w + 1u;          // untouched slot: 0xFFFFFFFF + 1 = 0
    }
    uint64_t h = slot_hash_q(key, log2cap);
    for (;;) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(region + h));
        const uint64_t kk = ((uint64_t)v.y << 32) | v.x;
        if (kk == key) { val = v.z; return v.w + 1u; }
        if (kk == kEmptyKey) { val = 0; return 0; }
        h = (h + 1) & (cap - 1);
    }
}

// one id list: c ids at ptr, or (ptr == nullptr, c == 1) the single id `one`
struct ListRef {
    const uint32_t *ptr;
    uint32_t c, one;
};

struct ProbeSrc {
    const uint64_t *qsk;     // [nq][n]
    const Slot *slots;
    const uint32_t *ids;
    uint32_t *pval, *pcnt;   // [nq][n] probe results (the id / start in ids, group size)
    uint64_t cap;
    uint32_t log2cap, n;
    __device__ __forceinline__ uint32_t subs() const { return n; }
    template <bool FIRST>
    __device__ __forceinline__ ListRef get(uint32_t q, uint32_t j) const {
        const size_t t = (size_t)q * n + j;
        uint32_t val, c;
        if (FIRST) {
            c = probe_slot(slots, cap, log2cap, j, qsk[t], val);
            pval[t] = val;
            pcnt[t] = c;
        } else {
            c = pcnt[t];
            val = pval[t];
        }
        ListRef r;
        r.c = c;
        r.one = val;
        r.ptr = c == 1 ? nullptr : ids + val;
        return r;
    }
};

struct PartsSrc {
    const uint64_t *offs[kMaxParts];   // each [nq+1]
    const uint32_t *ids[kMaxParts];
    uint32_t parts;
    __device__ __forceinline__ uint32_t subs() const { return parts; }
    template <bool FIRST>
    __device__ __forceinline__ ListRef get(uint32_t q, uint32_t j) const {
        const uint64_t o0 = offs[j][q], o1 = offs[j][q + 1];
        ListRef r;
        r.ptr = ids[j] + o0;
        r.c = (uint32_t)(o1 - o0);
        r.one = 0;
        return r;
    }
};

struct CountArgs {
    uint32_t *qcount;        // [nq+1]
    const uint64_t *out_off; // [nq+1]   (EMIT)
    uint32_t *out_ids;       //          (EMIT)
    uint32_t *heavy_list;    // [nq]
    unsigned long long *counters;   // [0] number of heavy queries, [1] total gathered ids
    uint32_t nq, thr;
};

template <typename Src, bool EMIT>
__global__ void __launch_bounds__(kLookupWarps * 32)
count_kernel(Src src, CountArgs a) {
    extern __shared__ uint32_t s_buf[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *buf = s_buf + (size_t)warp * kLookupCap;
    const uint32_t total_warps = gridDim.x * kLookupWarps;
    const uint32_t subs = src.subs();
    unsigned long long pairs_local = 0;

    for (uint32_t q = blockIdx.x * kLookupWarps + warp; q < a.nq; q += total_warps) {
        // ---- fetch the lists, lay them out in the buffer ----
        uint32_t T = 0;
        for (uint32_t j0 = 0; j0 < subs; j0 += 32) {
            const uint32_t j = j0 + lane;
            ListRef r;
            r.ptr = nullptr;
            r.c = 0;
            r.one = 0;
            if (j < subs) r = src.template get<!EMIT>(q, j);
            uint32_t incl = r.c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t round_total = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t off = T + incl - r.c;
            if ((uint64_t)T + round_total <= kLookupCap) {
                if (r.c == 1) buf[off] = r.ptr ? r.ptr[0] : r.one;
                else if (r.c > 1 && r.c <= 8)
                    for (uint32_t i = 0; i < r.c; ++i) buf[off + i] = r.ptr[i];
                uint32_t big = __ballot_sync(0xffffffffu, r.c > 8);
                while (big) {
                    const int s = __ffs(big) - 1;
                    big &= big - 1;
                    const uint32_t *bp = reinterpret_cast<const uint32_t *>(
                        __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(r.ptr), s));
                    const uint32_t bc = __shfl_sync(0xffffffffu, r.c, s);
                    const uint32_t bo = __shfl_sync(0xffffffffu, off, s);
                    for (uint32_t i = lane; i < bc; i += 32) buf[bo + i] = bp[i];
                }
            }
            // saturate: the sum of the list sizes can exceed 32 bits only in theory
            T = (uint64_t)T + round_total > 0xFFFFFFFFull ? 0xFFFFFFFFu : T + round_total;
        }
        if (!EMIT) pairs_local += lane == 0 ? T : 0;
        if (T > kLookupCap) {                     // global path handles this query
            if (!EMIT && lane == 0) {
                a.qcount[q] = 0;
                a.heavy_list[atomicAdd(a.counters, 1ULL)] = q;
            }
            continue;
        }
        // ---- sort (bitonic network over the padded buffer) ----
        uint32_t P = 32;
        while (P < T) P <<= 1;
        for (uint32_t i = T + lane; i < P; i += 32) buf[i] = 0xFFFFFFFFu;
        __syncwarp();
        for (uint32_t kk = 2; kk <= P; kk <<= 1) {
            for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
                for (uint32_t i = lane; i < P / 2; i += 32) {
                    const uint32_t lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                    const uint32_t hi = lo | j;
                    const uint32_t x = buf[lo], y = buf[hi];
                    const bool asc = (lo & kk) == 0;
                    if ((x > y) == asc) { buf[lo] = y; buf[hi] = x; }
                }
                __syncwarp();
            }
        }
        // ---- run lengths against the threshold (ReadFilter.cpp:76-82) ----
        uint32_t emitted = 0;
        const uint64_t out0 = EMIT ? a.out_off[q] : 0;
        for (uint32_t i0 = 0; i0 < T; i0 += 32) {
            const uint32_t i = i0 + lane;
            bool ok = false;
            uint32_t v = 0;
            if (i < T) {
                v = buf[i];
                const bool head = i == 0 || buf[i - 1] != v;
                ok = head && (a.thr <= 1 || (i + a.thr - 1 < T && buf[i + a.thr - 1] == v));
            }
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            if (EMIT && ok) a.out_ids[out0 + emitted + __popc(m & ((1u << lane) - 1))] = v;
            emitted += __popc(m);
        }
        if (!EMIT && lane == 0) a.qcount[q] = emitted;
        __syncwarp();
    }
    if (!EMIT && lane == 0 && pairs_local) atomicAdd(a.counters + 1, pairs_local);
}

// ---------------------------------------------------------------- global path --
template <typename Src>
__global__ void __launch_bounds__(256)
heavy_counts_kernel(Src src, const uint32_t *__restrict__ heavy_list, uint64_t items, uint32_t *__restrict__ hc) {
    const uint32_t subs = src.subs();
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < items;
         t += (uint64_t)gridDim.x * blockDim.x)
        hc[t] = src.template get<false>(heavy_list[t / subs], (uint32_t)(t % subs)).c;
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
        const ListRef r = src.template get<false>(heavy_list[h], (uint32_t)(t % subs));
        uint64_t *dst = pairs + (hoff[t] - pair0);
        const uint64_t tag = (h - h0) << 32;
        if (!r.ptr) { if (lane == 0 && r.c) dst[0] = tag | r.one; }
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

static int grid_for(uint64_t items, int sms, int per_block = 256) {
    uint64_t b = (items + per_block - 1) / per_block;
    uint64_t cap = (uint64_t)sms * 16;
    return (int)(b < cap ? (b ? b : 1) : cap);
}

// Queries that overflowed the warp buffer: global (heavy index, id) pair sort in batches.
template <typename Src>
static int heavy_path(nsmh_ctx *c, QueryWs &ws, const Src &src, uint32_t subs, uint32_t nh, cudaStream_t s) {
    const uint64_t items = (uint64_t)nh * subs;
    uint32_t *hl = ws.heavy_list.as<uint32_t>();
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

// count pass, prefix sum, emit pass for nq queries whose lists come from `src`.
// Result CSR in ws.out_off (u64 [nq+1]) / ws.out_ids (u32 [ws.last_total]).  Synchronises `s`.
template <typename Src>
static int count_and_emit(nsmh_ctx *c, QueryWs &ws, const Src &src, uint32_t subs, uint32_t nq, cudaStream_t s) {
    ws.last_nq = nq;
    ws.last_total = 0;
    ws.last_pairs = 0;
    NSMH_TRY(ws.out_off.ensure(((size_t)nq + 1) * sizeof(uint64_t), s));
    if (nq == 0) {
        NSMH_CK(cudaMemsetAsync(ws.out_off.p, 0, sizeof(uint64_t), s));
        NSMH_CK(cudaStreamSynchronize(s));
        return NSMH_OK;
    }
    NSMH_TRY(ws.qcount.ensure(((size_t)nq + 1) * sizeof(uint32_t), s));
    NSMH_TRY(ws.heavy_list.ensure((size_t)nq * sizeof(uint32_t), s));
    NSMH_TRY(ws.counters.ensure(4 * sizeof(uint64_t), s));
    NSMH_CK(cudaMemsetAsync(ws.counters.p, 0, 4 * sizeof(uint64_t), s));
    NSMH_CK(cudaMemsetAsync(ws.qcount.as<uint32_t>() + nq, 0, sizeof(uint32_t), s));

    const size_t smem = (size_t)kLookupWarps * kLookupCap * sizeof(uint32_t);
    NSMH_CK(cudaFuncSetAttribute(count_kernel<Src, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NSMH_CK(cudaFuncSetAttribute(count_kernel<Src, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CountArgs a;
    a.qcount = ws.qcount.as<uint32_t>();
    a.out_off = ws.out_off.as<uint64_t>();
    a.out_ids = nullptr;
    a.heavy_list = ws.heavy_list.as<uint32_t>();
    a.counters = ws.counters.as<unsigned long long>();
    a.nq = nq;
    a.thr = c->thr;
    int blocks = (int)std::min<uint64_t>(((uint64_t)nq + kLookupWarps - 1) / kLookupWarps, (uint64_t)c->num_sms * 3);
    count_kernel<Src, false><<<blocks, kLookupWarps * 32, smem, s>>>(src, a);
    ++ws.launches;
    NSMH_CK(cudaGetLastError());
    unsigned long long cnt[2] = {0, 0};
    NSMH_CK(cudaMemcpyAsync(cnt, ws.counters.p, sizeof cnt, cudaMemcpyDeviceToHost, s));
    NSMH_CK(cudaStreamSynchronize(s));
    const uint32_t nh = (uint32_t)cnt[0];
    ws.last_pairs = cnt[1];
    // a result id needs at least one gathered id, so their number bounds the output size
    NSMH_TRY(ws.out_ids.ensure((size_t)std::max<uint64_t>(cnt[1], 1) * sizeof(uint32_t), s));
    a.out_ids = ws.out_ids.as<uint32_t>();
    if (nh) NSMH_TRY(heavy_path(c, ws, src, subs, nh, s));

    size_t tmp_bytes = 0;
    NSMH_CK(cub_exclusive_sum_u32_to_u64(nullptr, tmp_bytes, ws.qcount.as<uint32_t>(), ws.out_off.as<uint64_t>(), (size_t)nq + 1, s));
    NSMH_TRY(ws.cub_tmp.ensure(tmp_bytes, s));
    NSMH_CK(cub_exclusive_sum_u32_to_u64(ws.cub_tmp.p, tmp_bytes, ws.qcount.as<uint32_t>(), ws.out_off.as<uint64_t>(), (size_t)nq + 1, s));
    count_kernel<Src, true><<<blocks, kLookupWarps * 32, smem, s>>>(src, a);
    ws.launches += 3;
    NSMH_CK(cudaGetLastError());
    if (nh) {
        heavy_copy_kernel<<<grid_for(nh, c->num_sms, 8), 256, 0, s>>>(
            ws.heavy_list.as<uint32_t>(), nh, ws.hcnt.as<uint32_t>(), ws.hstart.as<uint64_t>(),
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

static int make_probe_src(nsmh_ctx *c, QueryWs &ws, const uint64_t *d_qsketch, uint32_t nq, cudaStream_t s,
                          ProbeSrc &src) {
    Tables &T = c->tables;
    if (!T.built) return fail(NSMH_ESTATE, "query: tables not built (call nsmh_build)");
    const uint64_t items = std::max<uint64_t>((uint64_t)nq * c->n, 1);
    NSMH_TRY(ws.pval.ensure(items * sizeof(uint32_t), s));
    NSMH_TRY(ws.pcnt.ensure(items * sizeof(uint32_t), s));
    src.qsk = d_qsketch;
    src.slots = T.slots.as<Slot>();
    src.ids = T.ids.as<uint32_t>();
    src.pval = ws.pval.as<uint32_t>();
    src.pcnt = ws.pcnt.as<uint32_t>();
    src.cap = T.cap;
    src.log2cap = T.log2cap;
    src.n = c->n;
    return NSMH_OK;
}

// Query nq device-resident sketches [nq][n] against the tables.
int query_sketches_device(nsmh_ctx *c, QueryWs &ws, const uint64_t *d_qsketch, uint32_t nq, cudaStream_t s) {
    ProbeSrc src;
    NSMH_TRY(make_probe_src(c, ws, d_qsketch, nq, s, src));
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
probe_totals_kernel(ProbeSrc src, uint32_t nq, uint32_t *__restrict__ qcount) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < nq; q += warps) {
        uint32_t tot = 0;
        for (uint32_t j = lane; j < src.n; j += 32) tot += src.get<true>(q, j).c;
#pragma unroll
        for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) qcount[q] = tot;
    }
}

__global__ void __launch_bounds__(256)
probe_write_kernel(ProbeSrc src, uint32_t nq, const uint64_t *__restrict__ out_off, uint32_t *__restrict__ out_ids) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < nq; q += warps) {
        uint64_t base = out_off[q];
        for (uint32_t j0 = 0; j0 < src.n; j0 += 32) {
            const uint32_t j = j0 + lane;
            ListRef r;
            r.ptr = nullptr;
            r.c = 0;
            r.one = 0;
            if (j < src.n) r = src.get<false>(q, j);
            uint32_t incl = r.c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            uint32_t *dst = out_ids + base + (incl - r.c);
            if (r.c == 1 && !r.ptr) dst[0] = r.one;
            else for (uint32_t i = 0; i < r.c; ++i) dst[i] = r.ptr[i];
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

int probe_lists_device(nsmh_ctx *c, QueryWs &ws, const uint64_t *d_qsketch, uint32_t nq, cudaStream_t s) {
    ProbeSrc src;
    NSMH_TRY(make_probe_src(c, ws, d_qsketch, nq, s, src));
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

} // namespace nsmh

// Device code of the candidate lookup (the kernels of query.cu are thin wrappers around it): the
// body of count_kernel (a warp per query: counting table, counting filter + sort path, hand-over of
// larger queries) and the "counting filter" tier for queries whose gathered id lists do not fit the
// warp-private sort buffer (kLookupCap ids).  Everything here is warp-level and templated on the
// source of the id lists, and the file is kept free of runtime-API includes, so that
// tests/cpp/query_host_emul.cpp can compile the SAME code for the host
// (tests/cpp/cuda_host_shim.h, lock-step warp emulation) with lists given as a CSR and check it
// against a plain sort-and-count without a GPU.
//
// Reference semantics (src/ReadFilter.cpp:65-83): the ids of all n probed lists are gathered,
// sorted, and an id is emitted when it occurs at least overlapSketchThreshold times.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "nsmh_constants.h"
#include "nsmh_ldst.cuh"

namespace nsmh {

constexpr int kMaxParts = 16;        // PartsSrc: at most this many partial lists per query
constexpr uint32_t kNoId = 0xFFFFFFFFu;   // read ids are < 2^32-1 (ReadData.cpp:122-124)

// one id list: c ids at ptr, or (ptr == nullptr) the single id `one` (c == 1) / the two ids `one`, `two` (c == 2)
struct ListRef {
    const uint32_t *ptr;
    uint32_t c, one, two;
};

__device__ __forceinline__ ListRef empty_list() {
    ListRef r;
    r.ptr = nullptr;
    r.c = 0;
    r.one = r.two = 0;
    return r;
}

// The n tables, probed for a batch of query sketches.  probe_items_kernel resolves every
// (query, hash) item and stores {val, group size}; everything downstream reads those.

struct ProbeSrc {
    static constexpr bool kInlinePairs = false;
    const uint64_t *qsk;     // [nq][n]
    const Slot *slots;
    const uint32_t *ids;
    uint32_t *pval, *pcnt;   // [nq][n] probe results: the id / the start in ids, the group size
    uint64_t cap;
    uint32_t n;
    // Probing is split in two so that the (random, DRAM-latency) bucket loads of several lists
    // are in flight together: begin() issues the 32-byte load of the home bucket (both slots of
    // one sector), finish() resolves the probe and only rarely has to move to the next bucket.
    struct Pending {
        uint64_t key, b;
        uint64_t sa, sb, sc, sd;    // slot = {key, val | (cnt-1) << 32}, twice
        uint32_t j;
    };
    __device__ __forceinline__ uint32_t subs() const { return n; }
    __device__ __forceinline__ void prefetch(uint32_t q, int lane) const {      // row q of the sketches: n * 8 bytes
        if ((uint32_t)lane * 16 < n) l2_prefetch_line(qsk + (size_t)q * n + lane * 16);
    }
    // what does not change from query to query for list j (kept in registers across the query loop)
    struct Ctx {
        const Slot *region;
    };
    __device__ __forceinline__ Ctx context(uint32_t j) const { return Ctx{slots + (uint64_t)j * region_stride(cap)}; }
    __device__ __forceinline__ Pending begin(const Ctx &x, uint32_t q, uint32_t j) const {
        Pending p;
        p.j = j;
        p.key = __ldg(qsk + (size_t)q * n + j);
        p.b = p.key == kEmptyKey ? (cap >> 1) : slot_index(p.key, cap >> 1);   // key ~0 lives in the extra slot
        ldg256(x.region + 2 * p.b, p.sa, p.sb, p.sc, p.sd);
        return p;
    }
    __device__ __forceinline__ Pending begin(uint32_t q, uint32_t j) const { return begin(context(j), q, j); }
    // group size (0 = absent) and val (the id itself for a group of one, else the start in ids)
    __device__ __forceinline__ ListRef finish(Pending p) const {
        const Slot *region = slots + (uint64_t)p.j * region_stride(cap);
        const uint64_t nb = cap >> 1;
        uint64_t hit = 0;
        bool found = false;
        for (;;) {
            if (p.key == kEmptyKey || p.sa == p.key) { hit = p.sb; found = true; break; }
            if (p.sa == kEmptyKey) break;
            if (p.sc == p.key) { hit = p.sd; found = true; break; }
            if (p.sc == kEmptyKey) break;
            p.b = p.b + 1 == nb ? 0 : p.b + 1;
            ldg256(region + 2 * p.b, p.sa, p.sb, p.sc, p.sd);
        }
        ListRef r;
        r.c = found ? (uint32_t)(hit >> 32) + 1u : 0u;     // untouched extra slot: 0xFFFFFFFF + 1 = 0
        r.one = (uint32_t)hit;
        r.two = 0;
        r.ptr = r.c == 1 ? nullptr : ids + r.one;
        return r;
    }
    __device__ __forceinline__ ListRef get(uint32_t q, uint32_t j) const { return finish(begin(q, j)); }
};

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// `need` ids of inbox space from a 32-bit cursor.  A full inbox is not asked again (the cursor would wrap
// after 2^32 ids and hand out ranges that earlier results already point to): what does not fit stays behind
// and is read remotely, the cursor overshoots the capacity (<= 2^30) by what is in flight at most.
__device__ __forceinline__ uint32_t inbox_take(uint32_t *cursor, uint32_t need, uint32_t cap) {
    if (*reinterpret_cast<volatile uint32_t *>(cursor) >= cap) return cap;
    return atomicAdd(cursor, need);
}


// One thread per (query, 4 adjacent hash functions): the four keys are one 32-byte sector of the
// sketch row, every probe fetches one bucket = the two slots of one sector with ONE 256-bit
// load, and the four results leave as 16-byte stores.  Blocks are ordered by hash function
// (like table_insert_kernel), so the few table regions being probed at any time are L2
// resident: DRAM streams every region once instead of serving 64-byte bursts for random
// sectors all over the table.  The kernel is bound by the rate of 32-byte sector requests
// (tools/micro/atom_bench.cu), hence one request per probe.
__global__ void __launch_bounds__(kProbeRows, 4)
probe_items_kernel(ProbeSrc src, uint32_t nq) {
    const uint32_t chunks = (nq + kProbeRows - 1) / kProbeRows;
    const uint32_t colgroups = (src.n + kProbeCols - 1) / kProbeCols;
    const uint32_t units = chunks * colgroups;      // nq * n < 2^32 (probe_all)
    const uint64_t nb = src.cap >> 1, rstride = region_stride(src.cap);
    const bool vec = (src.n & 3) == 0;      // rows are sector aligned
    for (uint32_t u = blockIdx.x; u < units; u += gridDim.x) {
        const uint32_t cg = u / chunks;
        const uint32_t q = (u - cg * chunks) * kProbeRows + threadIdx.x;
        const uint32_t l0 = cg * kProbeCols;
        if (q >= nq) continue;
        const size_t t0 = (size_t)q * src.n + l0;
        uint64_t key[kProbeCols];
        if (vec) ldg256(src.qsk + t0, key[0], key[1], key[2], key[3]);
        else {
#pragma unroll
            for (int j = 0; j < kProbeCols; ++j) key[j] = l0 + j < src.n ? __ldg(src.qsk + t0 + j) : 0;
        }
        uint64_t b[kProbeCols], sa[kProbeCols], sb[kProbeCols], sc[kProbeCols], sd[kProbeCols];
#pragma unroll
        for (int j = 0; j < kProbeCols; ++j) {
            b[j] = key[j] == kEmptyKey ? nb : slot_index(key[j], nb);   // key ~0 lives in the extra slot
            const uint32_t l = min(l0 + j, src.n - 1);
            ldg256(src.slots + (uint64_t)l * rstride + 2 * b[j], sa[j], sb[j], sc[j], sd[j]);
        }
        uint32_t val[kProbeCols], cnt[kProbeCols];
#pragma unroll
        for (int j = 0; j < kProbeCols; ++j) {
            const Slot *region = src.slots + (uint64_t)min(l0 + j, src.n - 1) * rstride;
            for (;;) {
                // slot = {key, val | (cnt-1) << 32}
                if (key[j] == kEmptyKey || sa[j] == key[j]) { val[j] = (uint32_t)sb[j]; cnt[j] = (uint32_t)(sb[j] >> 32) + 1u; break; }
                if (sa[j] == kEmptyKey) { val[j] = 0; cnt[j] = 0; break; }
                if (sc[j] == key[j]) { val[j] = (uint32_t)sd[j]; cnt[j] = (uint32_t)(sd[j] >> 32) + 1u; break; }
                if (sc[j] == kEmptyKey) { val[j] = 0; cnt[j] = 0; break; }
                b[j] = b[j] + 1 == nb ? 0 : b[j] + 1;
                ldg256(region + 2 * b[j], sa[j], sb[j], sc[j], sd[j]);
            }
        }
        if (vec) {
            *reinterpret_cast<uint4 *>(src.pval + t0) = make_uint4(val[0], val[1], val[2], val[3]);
            *reinterpret_cast<uint4 *>(src.pcnt + t0) = make_uint4(cnt[0], cnt[1], cnt[2], cnt[3]);
        } else {
#pragma unroll
            for (int j = 0; j < kProbeCols; ++j)
                if (l0 + j < src.n) { src.pval[t0 + j] = val[j]; src.pcnt[t0 + j] = cnt[j]; }
        }
    }
}

// The same probe, for the tables this rank owns (src.n hash functions starting at dst.col0) and
// the reads of ALL ranks (q = global row).  A block takes 256 rows x one group of up to kPeerCols hash
// functions (two passes of 4 per thread), collects the results in shared memory and writes the tile out
// as ONE dense run per read owner into that owner's memory (peer mapping over NVLink): consecutive threads
// store consecutive 8-byte words, i.e. full 128-byte lines.  (The first version stored 32 bytes per thread
// at a stride of ncols * 8 bytes - half-written lines - and its time tripled from 2 to 8 ranks.)
// A group of two travels inside the result word; groups of 3..kInboxMaxGroup members are pushed along: their ids
// go into this rank's segment of the read owner's inbox, so the counting kernel over there finds them in its own
// memory instead of paying an NVLink round trip per list.  The ids of a tile are STAGED in shared memory and leave
// as one bulk copy next to the tile (space from a LOCAL cursor, one atomic per tile).  (The first version let every
// thread store its lists' ids straight into the peer's inbox, 4 bytes at a time: measured on 8 GPUs those
// scattered remote stores were 0.15 of the kernel's 0.36 ms, the 16 KB tiles cost nothing -
// profiles/r2_remote_store_split_s26.txt.)  Larger groups, what does not fit the staging area or the inbox, and
// the few tiles whose rows belong to two owners (direct stores there) stay behind and are read remotely on demand.
#ifndef NSMH_STAGE_IDS
#define NSMH_STAGE_IDS 1792            // (the host emulation builds with a small value so that tiles overflow it)
#endif
constexpr int kStageIds = NSMH_STAGE_IDS;      // ids staged per tile (7 per row on average; beyond that lists stay behind)
static_assert(kStageIds % 4 == 0, "the staged ids leave in 16-byte pieces");

__global__ void __launch_bounds__(kProbeRows, 4)
probe_to_peers_kernel(ProbeSrc src, uint32_t nq, PeerDst dst) {
    __align__(128) __shared__ uint64_t s_tile[2][kProbeRows * kPeerCols];     // two tiles: one may still be leaving
    __align__(16) __shared__ uint32_t s_ids[2][kStageIds];                     // ... and the inbox ids that go with them
    __shared__ uint32_t s_wsum[kProbeRows / 32], s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunks = (nq + kProbeRows - 1) / kProbeRows;
    const uint32_t colblocks = (src.n + kPeerCols - 1) / kPeerCols;
    const uint32_t units = chunks * colblocks;
    const uint64_t nb = src.cap >> 1, rstride = region_stride(src.cap);
    const bool vec_in = (src.n & 3) == 0;
    uint32_t round = 0;
    for (uint32_t u = blockIdx.x; u < units; u += gridDim.x, ++round) {
        uint64_t *tile = s_tile[round & 1];
        uint32_t *stage = s_ids[round & 1];
        // every rank starts with the chunks of its own reads and walks on from there: at any time the ranks
        // store into DIFFERENT peers (all of them starting at row 0 put 7 senders on one receiver's links)
        const uint32_t cbk = u / chunks;
        uint32_t chunk = u - cbk * chunks + dst.chunk0;
        chunk = chunk >= chunks ? chunk - chunks : chunk;
        const uint32_t q0 = chunk * kProbeRows, nr = min((uint32_t)kProbeRows, nq - q0);
        const uint32_t q = q0 + threadIdx.x;
        const uint32_t c0 = cbk * kPeerCols, w = min((uint32_t)kPeerCols, src.n - c0);     // this group's hash functions
        const bool active = q < nq;
        uint32_t o_first = 0;
        while (o_first + 1 < dst.world && q0 >= dst.row_end[o_first]) ++o_first;
        const bool one_owner = q0 + nr <= dst.row_end[o_first];     // block-uniform
        uint32_t o = o_first;
        if (active && !one_owner)
            while (o + 1 < dst.world && q >= dst.row_end[o]) ++o;        // owner of read q
        // ---- probe: kPeerCols tables, kProbeCols at a time ----
        uint32_t val[kPeerCols], cnt[kPeerCols];
        uint32_t need = 0;
#pragma unroll
        for (int g = 0; g < kPeerCols / kProbeCols; ++g) {
            const uint32_t l0 = c0 + g * kProbeCols;
#pragma unroll
            for (int j = 0; j < kProbeCols; ++j) val[g * kProbeCols + j] = cnt[g * kProbeCols + j] = 0;
            if (active && l0 < c0 + w) {
                const size_t t0 = (size_t)q * src.n + l0;
                uint64_t key[kProbeCols];
                if (vec_in) ldg256(src.qsk + t0, key[0], key[1], key[2], key[3]);
                else {
#pragma unroll
                    for (int j = 0; j < kProbeCols; ++j) key[j] = l0 + j < src.n ? __ldg(src.qsk + t0 + j) : 0;
                }
                uint64_t b[kProbeCols], sa[kProbeCols], sb[kProbeCols], sc[kProbeCols], sd[kProbeCols];
#pragma unroll
                for (int j = 0; j < kProbeCols; ++j) {
                    b[j] = key[j] == kEmptyKey ? nb : slot_index(key[j], nb);
                    const uint32_t l = min(l0 + j, src.n - 1);
                    ldg256(src.slots + (uint64_t)l * rstride + 2 * b[j], sa[j], sb[j], sc[j], sd[j]);
                }
#pragma unroll
                for (int j = 0; j < kProbeCols; ++j) {
                    const Slot *region = src.slots + (uint64_t)min(l0 + j, src.n - 1) * rstride;
                    uint32_t v = 0, c = 0;
                    for (;;) {
                        if (key[j] == kEmptyKey || sa[j] == key[j]) { v = (uint32_t)sb[j]; c = (uint32_t)(sb[j] >> 32) + 1u; break; }
                        if (sa[j] == kEmptyKey) break;
                        if (sc[j] == key[j]) { v = (uint32_t)sd[j]; c = (uint32_t)(sd[j] >> 32) + 1u; break; }
                        if (sc[j] == kEmptyKey) break;
                        b[j] = b[j] + 1 == nb ? 0 : b[j] + 1;
                        ldg256(region + 2 * b[j], sa[j], sb[j], sc[j], sd[j]);
                    }
                    if (l0 + j >= c0 + w) v = c = 0;
                    val[g * kProbeCols + j] = v;
                    cnt[g * kProbeCols + j] = c;
                    if (c >= 3 && c <= (uint32_t)kInboxMaxGroup) need += c;
                }
            }
        }
        // ---- inbox space for the tile's small groups ----
        // the copies that took these buffers two rounds ago have read them (the issuing thread waits, the
        // barriers below tell the rest)
        if (threadIdx.x == 0) bulk_store_wait_read<1>();
        uint32_t pos;           // where this thread's ids go: in the staging area (one owner) or in the inbox itself
        bool fits;
        uint32_t staged = 0;    // ids of the tile that leave through the staging area
        if (one_owner) {
            const uint32_t incl = warp_incl_scan(need, lane);
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            uint32_t before = 0, total = 0;
#pragma unroll
            for (int ww = 0; ww < kProbeRows / 32; ++ww) {
                const uint32_t t = s_wsum[ww];
                before += ww < warp ? t : 0u;
                total += t;
            }
            pos = before + incl - need;
            staged = min(total, (uint32_t)kStageIds) + 3u & ~3u;       // whole 16-byte pieces (the tail is padding)
            if (threadIdx.x == 0) {
                uint32_t base = 0xFFFFFFFFu;
                if (staged) {
                    base = inbox_take(dst.cursor + o_first, staged, dst.inbox_cap[o_first]);
                    if ((uint64_t)base + staged > dst.inbox_cap[o_first]) base = 0xFFFFFFFFu;
                }
                s_base = base;
            }
            __syncthreads();
            fits = need && s_base != 0xFFFFFFFFu && pos + need <= (uint32_t)kStageIds;
        } else {
            // rows of two owners (at most world - 1 tiles): every thread for itself, straight into the inbox
            const uint32_t take = need + 3u & ~3u;                      // the cursors stay multiples of 4
            pos = take ? inbox_take(dst.cursor + o, take, dst.inbox_cap[o]) : 0u;
            fits = need && (uint64_t)pos + take <= dst.inbox_cap[o];
            __syncthreads();
        }
        const uint32_t base = one_owner ? s_base : 0u;
#pragma unroll
        for (int j = 0; j < kPeerCols; ++j) {
            uint64_t out = (uint64_t)val[j] | ((uint64_t)cnt[j] << 32);
            if (cnt[j] == 2) {
                // a group of two travels inside the result word (ids < 2^31: checked by nsmh_mg_init)
                const uint32_t i0 = src.ids[val[j]], i1 = src.ids[val[j] + 1];
                out = (uint64_t)i0 | ((uint64_t)i1 << 31) | ((uint64_t)kPairFlag << 32);
            } else if (fits && cnt[j] >= 3 && cnt[j] <= (uint32_t)kInboxMaxGroup) {
                uint32_t *box = one_owner ? stage + pos : dst.inbox[o] + pos;
                for (uint32_t i = 0; i < cnt[j]; ++i) box[i] = src.ids[val[j] + i];
                out = (uint64_t)(base + pos) | ((uint64_t)(cnt[j] | kInboxFlag) << 32);
                pos += cnt[j];
            }
            tile[threadIdx.x * kPeerCols + j] = out;        // rows past nq / unused columns: zeroes
        }
        // ---- the tile [rows of the chunk][kPeerCols] leaves: a dense, 64-byte-aligned run per read owner ----
        bulk_store_fence();                 // the tile's stores are visible to the copy engine
        __syncthreads();
        if (one_owner) {
            if (threadIdx.x == 0) {
                const uint32_t r0 = o_first ? dst.row_end[o_first - 1] : 0u, rows_o = dst.row_end[o_first] - r0;
                bulk_store(dst.pr[o_first] + (size_t)rows_o * dst.pcol0 + peer_result_index(rows_o, q0 - r0, c0),
                           tile, nr * kPeerCols * (uint32_t)sizeof(uint64_t));
                if (s_base != 0xFFFFFFFFu) bulk_store(dst.inbox[o_first] + s_base, stage, staged * (uint32_t)sizeof(uint32_t));
                bulk_store_commit();
            }
        } else {
            for (uint32_t e = threadIdx.x; e < nr * kPeerCols; e += kProbeRows) {
                const uint32_t r = e / kPeerCols, c = e % kPeerCols, qq = q0 + r;
                uint32_t oo = 0;
                while (oo + 1 < dst.world && qq >= dst.row_end[oo]) ++oo;
                const uint32_t r0 = oo ? dst.row_end[oo - 1] : 0u, rows_o = dst.row_end[oo] - r0;
                dst.pr[oo][(size_t)rows_o * dst.pcol0 + peer_result_index(rows_o, qq - r0, c0 + c)] = tile[e];
            }
            if (threadIdx.x == 0) bulk_store_commit();          // an empty group keeps the two-round bookkeeping uniform
        }
    }
    if (threadIdx.x == 0) bulk_store_wait_all();        // the shared memory must outlive the copies
}

// id lists from stored probe results
struct StoredSrc {
    static constexpr bool kInlinePairs = false;
    const uint32_t *pval, *pcnt;   // [nq][n]
    const uint32_t *ids;
    uint32_t n;
    struct Pending {
        uint32_t val, c;
    };
    __device__ __forceinline__ uint32_t subs() const { return n; }
    __device__ __forceinline__ void prefetch(uint32_t q, int lane) const {
        if ((uint32_t)lane * 32 < n) {
            l2_prefetch_line(pval + (size_t)q * n + lane * 32);
            l2_prefetch_line(pcnt + (size_t)q * n + lane * 32);
        }
    }
    struct Ctx {};
    __device__ __forceinline__ Ctx context(uint32_t) const { return Ctx{}; }
    __device__ __forceinline__ Pending begin(const Ctx &, uint32_t q, uint32_t j) const { return begin(q, j); }
    __device__ __forceinline__ Pending begin(uint32_t q, uint32_t j) const {
        Pending p;
        p.val = __ldg(pval + (size_t)q * n + j);
        p.c = __ldg(pcnt + (size_t)q * n + j);
        return p;
    }
    __device__ __forceinline__ ListRef finish(Pending p) const {
        ListRef r;
        r.c = p.c;
        r.one = p.val;
        r.two = 0;
        r.ptr = p.c == 1 ? nullptr : ids + p.val;
        return r;
    }
    __device__ __forceinline__ ListRef get(uint32_t q, uint32_t j) const { return finish(begin(q, j)); }
};

struct PartsSrc {
    static constexpr bool kInlinePairs = false;
    const uint64_t *offs[kMaxParts];   // each [nq+1]
    const uint32_t *ids[kMaxParts];
    uint32_t parts;
    struct Pending {
        uint64_t o0, o1;
        uint32_t j;
    };
    __device__ __forceinline__ uint32_t subs() const { return parts; }
    __device__ __forceinline__ void prefetch(uint32_t, int) const {}
    struct Ctx {};
    __device__ __forceinline__ Ctx context(uint32_t) const { return Ctx{}; }
    __device__ __forceinline__ Pending begin(const Ctx &, uint32_t q, uint32_t j) const { return begin(q, j); }
    __device__ __forceinline__ Pending begin(uint32_t q, uint32_t j) const {
        Pending p;
        p.j = j;
        p.o0 = offs[j][q];
        p.o1 = offs[j][q + 1];
        return p;
    }
    __device__ __forceinline__ ListRef finish(Pending p) const {
        ListRef r;
        r.ptr = ids[p.j] + p.o0;
        r.c = (uint32_t)(p.o1 - p.o0);
        r.one = r.two = 0;
        return r;
    }
    __device__ __forceinline__ ListRef get(uint32_t q, uint32_t j) const { return finish(begin(q, j)); }
};

// Multi-GPU (multigpu.cu): the probe results of every (query, hash) were stored here by the
// rank that owns the hash function's table; the ids of groups with two or more members stay
// in the owner's memory and are read through the NVLink peer mapping.
struct PeerSrc {
    static constexpr bool kInlinePairs = true;     // groups of two travel inside the probe result (kPairFlag)
    PeerLists L;
    struct Pending {
        uint64_t v;
        uint32_t o;
    };
    // list j of every query: which rank owns the hash function, and where its results for row 0 are
    struct Ctx {
        const uint64_t *p;
        uint32_t o;
    };
    __device__ __forceinline__ uint32_t subs() const { return L.n; }
    __device__ __forceinline__ void prefetch(uint32_t, int) const {}
    __device__ __forceinline__ Ctx context(uint32_t j) const {
        uint32_t o = 0, pcol = 0;
        while (o + 1 < L.world && j >= L.col_end[o]) {
            pcol += peer_padded_cols(L.col_end[o] - (o ? L.col_end[o - 1] : 0u));
            ++o;
        }
        const uint32_t cb = o ? L.col_end[o - 1] : 0u;
        return Ctx{L.pr + (size_t)L.rows * pcol + peer_result_index(L.rows, 0, j - cb), o};
    }
    __device__ __forceinline__ Pending begin(const Ctx &x, uint32_t q, uint32_t) const {
        Pending p;
        p.o = x.o;
        p.v = __ldg(x.p + (size_t)q * kPeerCols);
        return p;
    }
    __device__ __forceinline__ Pending begin(uint32_t q, uint32_t j) const { return begin(context(j), q, j); }
    __device__ __forceinline__ ListRef finish(Pending p) const {
        ListRef r;
        r.c = (uint32_t)(p.v >> 32);
        r.one = (uint32_t)p.v;
        r.two = 0;
        r.ptr = nullptr;
        if (r.c & kInboxFlag) {                  // the owner pushed the ids into our inbox
            r.c &= ~kInboxFlag;
            r.ptr = L.inbox + (size_t)p.o * L.inbox_cap + r.one;
        } else if (r.c & kPairFlag) {            // a group of two, both ids inside the result word
            r.c = 2;
            r.one = (uint32_t)p.v & 0x7FFFFFFFu;
            r.two = (uint32_t)(p.v >> 31) & 0x7FFFFFFFu;
        } else if (r.c > 1) {
            r.ptr = L.ids[p.o] + r.one;          // NVLink read from the owner
        }
        return r;
    }
    __device__ __forceinline__ ListRef get(uint32_t q, uint32_t j) const { return finish(begin(q, j)); }
};


// ascending bitonic sort of buf[0..P), P a power of two >= 32, by one warp
__device__ __forceinline__ void warp_bitonic_smem(uint32_t *buf, uint32_t P, int lane) {
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
}

// ------------------------------------------------------------------ counting filter, in place --
// The same idea for the ids a warp has already laid out in its sort buffer (count_kernel's sort path,
// 128 < gathered ids <= kLookupCap): count them into 512 16-bit counters, keep the ids whose bucket
// reaches the threshold (stable, in place), and let the bitonic sort run on the survivors only -
// typically tens instead of a thousand.  Exact for the same reason as below; T <= 65535.
constexpr int kWarpFilterBuckets = 512;     // two per word: c16 = kWarpFilterBuckets / 2 words

__device__ __forceinline__ uint32_t warp_filter_bucket(uint32_t id) { return (id * 0x9E3779B1u) >> 23; }   // 9 bits

__device__ __forceinline__ uint32_t warp_filter_ids(uint32_t *buf, uint32_t T, uint32_t thr, uint32_t *c16, int lane) {
    for (int i = lane; i < kWarpFilterBuckets / 2; i += 32) c16[i] = 0;
    __syncwarp();
    for (uint32_t i = lane; i < T; i += 32) {
        const uint32_t b = warp_filter_bucket(buf[i]);
        atomicAdd(c16 + (b >> 1), 1u << (16 * (b & 1)));
    }
    __syncwarp();
    uint32_t S = 0;
    for (uint32_t i0 = 0; i0 < T; i0 += 32) {
        const uint32_t i = i0 + lane;
        bool keep = false;
        uint32_t v = 0;
        if (i < T) {
            v = buf[i];
            const uint32_t b = warp_filter_bucket(v);
            keep = ((c16[b >> 1] >> (16 * (b & 1))) & 0xFFFFu) >= thr;
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, keep);
        __syncwarp();       // every read of this round happens before its writes (S <= i0)
        if (keep) buf[S + __popc(mask & ((1u << lane) - 1))] = v;
        S += __popc(mask);
        __syncwarp();
    }
    return S;
}

// ------------------------------------------------------------------ the lookup kernel's body --
constexpr int kLookupCap = 1024;     // ids per warp-private sort buffer (n <= 64)
constexpr int kLookupCapWide = 2048; // ... for n > 64: twice the lists, twice the ids (at k = 15, n = 120 a third of the
                                     // queries gathered 1024..2048 ids and paid for a second tier; the registers of the
                                     // 4-lists-per-lane variant already limit it to 3 blocks per SM, which 3 x 74 KB fit)
constexpr int kResWords = 256;       // result list of the register path / filter counters of the sort path
constexpr int kRegListsMax = 4;      // lists per lane that are probed once and kept in registers (n <= 128; 2 for n <= 64)
static_assert(kResWords >= kWarpFilterBuckets / 2, "the sort path keeps its filter counters in the result area");
static_assert(kResWords >= 2 * 32 * kRegListsMax, "thr <= 1: every gathered id is a result");
constexpr int kWarpWords = kLookupCap + kResWords + 32;
__host__ __device__ constexpr int warp_words(int cap) { return cap + kResWords + 32; }
constexpr int kLookupWarps = 8;

constexpr int kFixedIds = 16;        // result ids per query that have a fixed place in tmp_ids

struct CountArgs {
    uint32_t *qcount;        // [nq+1] result ids per query
    uint64_t *qpos;          // [nq]   where the query's results start in tmp_ids (~0: heavy query)
    // Results before the prefix sum: query q owns tmp_ids[q * kFixedIds ..); a query with more results
    // takes a range behind those nq * kFixedIds entries from a cursor.  (With one cursor for everything every
    // query waited for the round trip of an atomic on ONE address before it could store its results.)
    uint32_t *tmp_ids;
    uint64_t tmp_cap;        // entries, >= nq * kFixedIds
    uint32_t *heavy_list;    // [nq]
    unsigned long long *counters;   // [0] heavy queries, [1] gathered ids, [2] result ids, [3] cursor behind the fixed places
    uint32_t nq, thr;
};

// One warp per query, ONE pass.
//   gather     up to 32 * kRegLists lists are probed once, kRegLists per lane, all loads in flight together
//   registers  <= 32 * kRegLists gathered ids (every list a single id, or laid out through the buffer): counted
//              with warp votes - the lowest lane that still holds an id broadcasts it, one ballot per
//              register says who else holds it, the population count is its multiplicity; one round
//              per DISTINCT id, no shared-memory traffic (the first version counted in a shared-memory
//              hash table: 2 atomics per id at 2 cycles per lane made the shared-memory pipe, not the
//              DRAM latency of the probes, the kernel's bound)
//   sort path  <= kCap ids: laid out, counting filter in place, bitonic sort, run lengths
//   beyond     handed to the counting-filter tier / the global path
// Results go to tmp_ids in completion order; a prefix sum over qcount and csr_place_kernel then
// produce the CSR.
// queries q_first, q_first + q_stride, ... < a.nq by ONE warp; buf = kWarpWords warp-private words
template <typename Src, int kRegLists = kRegListsMax, int kCap = kLookupCap>
__device__ __forceinline__ void count_queries(Src src, CountArgs a, uint32_t *buf, uint32_t q_first, uint32_t q_stride) {
    constexpr int kRegIds = Src::kInlinePairs ? 2 * kRegLists : kRegLists;      // ids per lane held in registers
    constexpr int kRegMaxIds = 32 * kRegIds;     // gathered ids that are counted in registers
    const int lane = threadIdx.x & 31;
    uint32_t *res = buf + kCap;           // kResWords entries
    const uint32_t total_warps = q_stride;
    const uint32_t subs = src.subs();
    const bool keep_lists = subs <= (uint32_t)kRegMaxIds;
    unsigned long long pairs_local = 0, results_local = 0;
    typename Src::Ctx ctx[kRegLists];
#pragma unroll
    for (int c = 0; c < kRegLists; ++c) ctx[c] = src.context(min((uint32_t)(c * 32 + lane), subs ? subs - 1 : 0u));

    for (uint32_t q = q_first; q < a.nq; q += total_warps) {
        if ((uint64_t)q + total_warps < a.nq) src.prefetch(q + total_warps, lane);    // the next query's keys, into L2
        ListRef r[kRegLists];
        uint32_t v[kRegIds];
        uint32_t T = 0xFFFFFFFFu;           // gathered ids (saturating)
        bool in_regs = false;
        if (keep_lists) {
            typename Src::Pending pend[kRegLists];
#pragma unroll
            for (int c = 0; c < kRegLists; ++c)
                if (c * 32 + lane < (int)subs) pend[c] = src.begin(ctx[c], q, c * 32 + lane);
            unsigned long long mine = 0;
            bool multi = false;
#pragma unroll
            for (int c = 0; c < kRegLists; ++c) {
                r[c] = c * 32 + lane < (int)subs ? src.finish(pend[c]) : empty_list();
                mine += r[c].c;
                if (Src::kInlinePairs) {
                    multi |= r[c].c > 2 || (r[c].c == 2 && r[c].ptr);
                    v[2 * c] = r[c].c == 1 ? (r[c].ptr ? r[c].ptr[0] : r[c].one) : r[c].c == 2 ? r[c].one : kNoId;
                    v[2 * c + 1] = r[c].c == 2 ? r[c].two : kNoId;
                } else {
                    multi |= r[c].c > 1;
                    v[c] = r[c].c == 1 ? (r[c].ptr ? r[c].ptr[0] : r[c].one) : kNoId;
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
            T = mine > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)mine;
            in_regs = T <= (uint32_t)kRegMaxIds && !__any_sync(0xffffffffu, multi);
        }
        uint32_t R = 0;
        const uint32_t *out = res;
        if (!in_regs) {
            // ---- lay the lists out in the buffer ----
            T = 0;
            __syncwarp();                   // the previous query's results have left the buffer
            auto lay = [&](const ListRef &rr) {
                const uint32_t incl = warp_incl_scan(rr.c, lane);
                const uint32_t round_total = __shfl_sync(0xffffffffu, incl, 31);
                const uint32_t off = T + incl - rr.c;
                if ((uint64_t)T + round_total <= (uint32_t)kCap) {
                    if (rr.c == 1) buf[off] = rr.ptr ? rr.ptr[0] : rr.one;
                    else if (rr.c == 2 && !rr.ptr) { buf[off] = rr.one; buf[off + 1] = rr.two; }
                    else if (rr.c > 1 && rr.c <= 8)
                        for (uint32_t i = 0; i < rr.c; ++i) buf[off + i] = rr.ptr[i];
                    uint32_t big = __ballot_sync(0xffffffffu, rr.c > 8);       // lists that long are never inline
                    while (big) {
                        const int sl = __ffs(big) - 1;
                        big &= big - 1;
                        const uint32_t *bp = reinterpret_cast<const uint32_t *>(
                            __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(rr.ptr), sl));
                        const uint32_t bc = __shfl_sync(0xffffffffu, rr.c, sl);
                        const uint32_t bo = __shfl_sync(0xffffffffu, off, sl);
                        for (uint32_t i = lane; i < bc; i += 32) buf[bo + i] = bp[i];
                    }
                }
                // saturate: the sum of the list sizes can exceed 32 bits only in theory
                T = (uint64_t)T + round_total > 0xFFFFFFFFull ? 0xFFFFFFFFu : T + round_total;
            };
            if (keep_lists) {
#pragma unroll
                for (int c = 0; c < kRegLists; ++c)
                    if (c * 32 < (int)subs) lay(r[c]);
            } else {
                for (uint32_t j0 = 0; j0 < subs; j0 += 32) lay(j0 + lane < subs ? src.get(q, j0 + lane) : empty_list());
            }
            if (T > (uint32_t)kCap) {                     // the next tier handles this query
                pairs_local += lane == 0 ? T : 0;
                if (lane == 0) {
                    a.qcount[q] = 0;
                    a.qpos[q] = ~0ULL;
                    a.heavy_list[atomicAdd(a.counters, 1ULL)] = q;
                }
                __syncwarp();
                continue;
            }
            __syncwarp();
            if (T <= (uint32_t)kRegMaxIds) {
#pragma unroll
                for (int c = 0; c < kRegIds; ++c) v[c] = c * 32 + lane < (int)T ? buf[c * 32 + lane] : kNoId;
                in_regs = true;
            }
        }
        if (in_regs) {
            // ---- count in registers: one round per distinct id ----
            uint32_t rem[kRegIds];
#pragma unroll
            for (int c = 0; c < kRegIds; ++c) rem[c] = __ballot_sync(0xffffffffu, v[c] != kNoId);
#pragma unroll
            for (int c = 0; c < kRegIds; ++c) {
                while (rem[c]) {
                    const uint32_t id = __shfl_sync(0xffffffffu, v[c], __ffs(rem[c]) - 1);
                    uint32_t cnt = 0;
#pragma unroll
                    for (int d = c; d < kRegIds; ++d) {         // earlier registers hold no id that is still uncounted
                        const uint32_t eq = __ballot_sync(0xffffffffu, v[d] == id);
                        cnt += __popc(eq);
                        rem[d] &= ~eq;
                    }
                    if (cnt >= a.thr) {
                        if (lane == 0) res[R] = id;
                        ++R;
                    }
                }
            }
            __syncwarp();
            if (R > 1 && R <= 32) {
                // shuffle bitonic sort of up to 32 results
                uint32_t w = lane < (int)R ? res[lane] : kNoId;
#pragma unroll
                for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
                    for (int j = kk >> 1; j > 0; j >>= 1) {
                        const uint32_t o = __shfl_xor_sync(0xffffffffu, w, j);
                        const bool up = (lane & kk) == 0, lower = (lane & j) == 0;
                        w = (lower == up) ? min(w, o) : max(w, o);
                    }
                }
                __syncwarp();
                if (lane < (int)R) res[lane] = w;
                __syncwarp();
            } else if (R > 32) {
                uint32_t P = 64;
                while (P < R) P <<= 1;
                for (uint32_t i = R + lane; i < P; i += 32) res[i] = kNoId;
                __syncwarp();
                warp_bitonic_smem(res, P, lane);
            }
        } else {
            // ---- sort path ----
            // ids whose counting-filter bucket stays below the threshold cannot qualify: drop them
            // before sorting (`res` holds the counters)
            uint32_t Ts = T;
            if (a.thr > 1 && T > 64) Ts = warp_filter_ids(buf, T, a.thr, res, lane);
            uint32_t P = 32;
            while (P < Ts) P <<= 1;
            for (uint32_t i = Ts + lane; i < P; i += 32) buf[i] = kNoId;
            __syncwarp();
            warp_bitonic_smem(buf, P, lane);
            // run lengths against the threshold (ReadFilter.cpp:76-82), compacted in place
            for (uint32_t i0 = 0; i0 < Ts; i0 += 32) {
                const uint32_t i = i0 + lane;
                bool ok = false;
                uint32_t w = 0;
                if (i < Ts) {
                    w = buf[i];
                    const bool head = i == 0 || buf[i - 1] != w;
                    ok = head && (a.thr <= 1 || (i + a.thr - 1 < Ts && buf[i + a.thr - 1] == w));
                }
                const uint32_t m = __ballot_sync(0xffffffffu, ok);
                __syncwarp();       // every read of this round happens before its writes (R <= i0)
                if (ok) buf[R + __popc(m & ((1u << lane) - 1))] = w;
                R += __popc(m);
                __syncwarp();
            }
            out = buf;
        }
        pairs_local += lane == 0 ? T : 0;
        // ---- hand the results over ----
        unsigned long long base = (unsigned long long)q * kFixedIds;
        if (R > (uint32_t)kFixedIds) {
            if (lane == 0) base = (unsigned long long)a.nq * kFixedIds + atomicAdd(a.counters + 3, (unsigned long long)R);
            base = __shfl_sync(0xffffffffu, base, 0);
        }
        if (lane == 0) {
            a.qcount[q] = R;
            a.qpos[q] = base;
        }
        results_local += lane == 0 ? R : 0;
        if (base + R <= a.tmp_cap)
            for (uint32_t i = lane; i < R; i += 32) a.tmp_ids[base + i] = out[i];
        __syncwarp();
    }
    if (lane == 0 && pairs_local) atomicAdd(a.counters + 1, pairs_local);
    if (lane == 0 && results_local) atomicAdd(a.counters + 2, results_local);
}

// the lookup kernel's body: blocks of kLookupWarps warps, a warp per query
template <typename Src, int kRegLists = kRegListsMax, int kCap = kLookupCap>
__device__ __forceinline__ void count_body(Src src, CountArgs a, uint32_t *s_buf) {
    const uint32_t warp = threadIdx.x >> 5;
    count_queries<Src, kRegLists, kCap>(src, a, s_buf + (size_t)warp * warp_words(kCap), blockIdx.x * kLookupWarps + warp,
                                        gridDim.x * kLookupWarps);
}

// ------------------------------------------------------------------ counting-filter tier --
// With small k (or very deep coverage) most table groups hold tens of reads that share a sketch
// value by chance, so a query gathers thousands of ids of which only a handful occur
// overlapSketchThreshold times.  Sorting all of them (globally: 8 bytes per id through several
// radix passes in HBM) is wasted work.  One warp per such query instead
//   1. counts every gathered id into kMidBuckets 16-bit counters in shared memory (a query with at
//      most 65535 ids cannot overflow a counter),
//   2. walks the lists a second time and keeps the ids whose bucket reached the threshold - every
//      occurrence of an id that qualifies survives, because its bucket counts at least its own
//      occurrences; ids that do not qualify survive only through bucket collisions,
//   3. sorts the survivors (<= kMidCap) and thresholds the run lengths exactly like the sort path.
// The result is exact; a query with too many ids or survivors is handed to the global sort path.
constexpr int kMidBuckets = 4096;
constexpr int kMidCap = 1024;
constexpr uint32_t kMidMaxIds = 65535;
constexpr int kMidWarpWords = kMidBuckets / 2 + kMidCap + 32;
constexpr int kMidWarps = 4;
constexpr uint64_t kMidPosFlag = 1ULL << 63;     // in qpos: "results are in mid_ids" (~0 stays "global path")

struct MidArgs {
    const uint32_t *heavy_list;     // [nh] queries that overflowed the warp buffer
    uint32_t *unresolved_list;      // [nh] out: the ones this tier hands on
    uint32_t *mid_ids;              // results of the resolved ones, in completion order
    uint64_t mid_cap;
    uint32_t *qcount;               // [nq]
    uint64_t *qpos;                 // [nq]
    unsigned long long *counters;   // [0] unresolved, [1] cursor in mid_ids, [2] overflow of mid_ids (must stay 0)
    uint32_t nh, thr;
};

__device__ __forceinline__ uint32_t mid_bucket(uint32_t id) { return (id * 0x9E3779B1u) >> 20; }   // 12 bits

// f(id) for every id of the query's lists: short lists by the lane that owns them, longer ones by
// the whole warp (coalesced)
template <typename Src, typename F>
__device__ __forceinline__ void mid_for_each_id(const Src &src, uint32_t q, uint32_t subs, int lane, F f) {
    for (uint32_t j0 = 0; j0 < subs; j0 += 32) {
        const uint32_t j = j0 + lane;
        const ListRef r = j < subs ? src.get(q, j) : empty_list();
        if (r.c == 1) f(r.ptr ? r.ptr[0] : r.one);
        else if (r.c == 2 && !r.ptr) { f(r.one); f(r.two); }
        else if (r.c > 1 && r.c <= 4)
            for (uint32_t i = 0; i < r.c; ++i) f(r.ptr[i]);
        uint32_t big = __ballot_sync(0xffffffffu, r.c > 4);
        while (big) {
            const int sl = __ffs(big) - 1;
            big &= big - 1;
            const uint32_t *bp = reinterpret_cast<const uint32_t *>(
                __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(r.ptr), sl));
            const uint32_t bc = __shfl_sync(0xffffffffu, r.c, sl);
            for (uint32_t i = lane; i < bc; i += 32) f(bp[i]);
        }
    }
}

// One warp: heavy queries first, first + stride, ...; wbuf = kMidWarpWords warp-private words.
template <typename Src>
__device__ __forceinline__ void mid_count_body(const Src &src, const MidArgs &m, uint32_t *wbuf, uint32_t first,
                                               uint32_t stride) {
    const int lane = threadIdx.x & 31;
    uint32_t *cnt = wbuf;                          // kMidBuckets 16-bit counters
    uint32_t *surv = wbuf + kMidBuckets / 2;       // kMidCap ids
    uint32_t *nsurv = surv + kMidCap;
    const uint32_t subs = src.subs();
    for (uint32_t h = first; h < m.nh; h += stride) {
        const uint32_t q = m.heavy_list[h];
        unsigned long long T = 0;
        for (uint32_t j = lane; j < subs; j += 32) T += src.get(q, j).c;
#pragma unroll
        for (int o = 16; o; o >>= 1) T += __shfl_xor_sync(0xffffffffu, T, o);
        bool ok = T <= kMidMaxIds;
        uint32_t S = 0;
        if (ok) {
            for (int i = lane; i < kMidBuckets / 2; i += 32) cnt[i] = 0;
            if (lane == 0) *nsurv = 0;
            __syncwarp();
            mid_for_each_id(src, q, subs, lane, [&](uint32_t id) {
                const uint32_t b = mid_bucket(id);
                atomicAdd(cnt + (b >> 1), 1u << (16 * (b & 1)));
            });
            __syncwarp();
            const uint32_t thr = m.thr;
            mid_for_each_id(src, q, subs, lane, [&](uint32_t id) {
                const uint32_t b = mid_bucket(id);
                if (((cnt[b >> 1] >> (16 * (b & 1))) & 0xFFFFu) >= thr) {
                    const uint32_t p = atomicAdd(nsurv, 1u);
                    if (p < (uint32_t)kMidCap) surv[p] = id;
                }
            });
            __syncwarp();
            S = *nsurv;
            ok = S <= (uint32_t)kMidCap;
        }
        if (!ok) {
            if (lane == 0) m.unresolved_list[atomicAdd(m.counters, 1ULL)] = q;
            __syncwarp();
            continue;
        }
        uint32_t P = 32;
        while (P < S) P <<= 1;
        for (uint32_t i = S + lane; i < P; i += 32) surv[i] = kNoId;
        __syncwarp();
        warp_bitonic_smem(surv, P, lane);
        // run lengths against the threshold (ReadFilter.cpp:76-82), compacted in place
        uint32_t R = 0;
        for (uint32_t i0 = 0; i0 < S; i0 += 32) {
            const uint32_t i = i0 + lane;
            bool keep = false;
            uint32_t v = 0;
            if (i < S) {
                v = surv[i];
                const bool head = i == 0 || surv[i - 1] != v;
                keep = head && (m.thr <= 1 || (i + m.thr - 1 < S && surv[i + m.thr - 1] == v));
            }
            const uint32_t mask = __ballot_sync(0xffffffffu, keep);
            __syncwarp();       // every read of this round happens before its writes (R <= i0)
            if (keep) surv[R + __popc(mask & ((1u << lane) - 1))] = v;
            R += __popc(mask);
            __syncwarp();
        }
        unsigned long long base = 0;
        if (lane == 0) {
            base = atomicAdd(m.counters + 1, (unsigned long long)R);
            if (base + R > m.mid_cap) m.counters[2] = 1ULL;
            m.qcount[q] = R;
            m.qpos[q] = kMidPosFlag | base;
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base + R <= m.mid_cap)
            for (uint32_t i = lane; i < R; i += 32) m.mid_ids[base + i] = surv[i];
        __syncwarp();
    }
}

} // namespace nsmh

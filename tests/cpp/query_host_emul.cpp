// TEST INFRASTRUCTURE ONLY.  Runs the counting-filter tier of the candidate lookup
// (nanospring_b200/csrc/query_kernels.cuh: mid_count_body, the code mid_count_kernel in query.cu wraps)
// on the host with cuda_host_shim.h, over id lists given as a CSR.  tests/test_query_emul.py
// compares the outcome with a plain sort-and-count (what ReadFilter.cpp:65-83 does); a logic check
// for the container without a GPU, never a product path.
#define NSMH_HOST_EMUL 1
#include "cuda_host_shim.h"

#include "../../nanospring_b200/csrc/query_kernels.cuh"

using namespace nsmh;

namespace {
// lists of query q: list_off[q * subs + j] .. list_off[q * subs + j + 1]; every other list of one id
// is handed out in the inlined form the hash-table slots use (ptr == nullptr, id in `one`)
struct CsrSrc {
    const uint64_t *list_off;
    const uint32_t *ids;
    uint32_t nsubs;
    static constexpr bool kInlinePairs = false;
    uint32_t subs() const { return nsubs; }
    void prefetch(uint32_t, int) const {}
    struct Pending {
        uint32_t q, j;
    };
    struct Ctx {};
    Ctx context(uint32_t) const { return Ctx{}; }
    Pending begin(const Ctx &, uint32_t q, uint32_t j) const { return Pending{q, j}; }
    Pending begin(uint32_t q, uint32_t j) const { return Pending{q, j}; }
    ListRef finish(Pending p) const { return get(p.q, p.j); }
    ListRef get(uint32_t q, uint32_t j) const {
        const uint64_t o0 = list_off[(uint64_t)q * nsubs + j], o1 = list_off[(uint64_t)q * nsubs + j + 1];
        ListRef r;
        r.c = (uint32_t)(o1 - o0);
        r.one = r.two = 0;
        r.ptr = ids + o0;
        if (r.c == 1 && (j & 1)) {
            r.one = ids[o0];
            r.ptr = nullptr;
        }
        return r;
    }
};
}  // namespace

extern "C" {

int mid_emul_constants(uint32_t *buckets, uint32_t *cap, uint32_t *max_ids) {
    *buckets = kMidBuckets;
    *cap = kMidCap;
    *max_ids = kMidMaxIds;
    return 0;
}

// All nq queries are "heavy".  qcount/qpos [nq] must come in as 0 / ~0 (what count_kernel leaves for a
// heavy query); counters [3] zeroed.  grid blocks of 2 warps.
void mid_emul_run(const uint64_t *list_off, const uint32_t *ids, uint32_t nq, uint32_t subs, uint32_t thr,
                  unsigned grid, uint32_t *qcount, uint64_t *qpos, uint32_t *mid_ids, uint64_t mid_cap,
                  uint32_t *unresolved, unsigned long long *counters) {
    CsrSrc src{list_off, ids, subs};
    std::vector<uint32_t> heavy(nq);
    for (uint32_t q = 0; q < nq; ++q) heavy[q] = nq - 1 - q;          // any order
    MidArgs m;
    m.heavy_list = heavy.data();
    m.unresolved_list = unresolved;
    m.mid_ids = mid_ids;
    m.mid_cap = mid_cap;
    m.qcount = qcount;
    m.qpos = qpos;
    m.counters = counters;
    m.nh = nq;
    m.thr = thr;
    std::vector<uint32_t> smem((size_t)grid * 2 * kMidWarpWords, 0xA5A5A5A5u);     // never assumed zero
    emu_launch(grid, 64, [&] {
        const uint32_t warp = threadIdx.x >> 5;
        mid_count_body(src, m, smem.data() + ((size_t)blockIdx.x * 2 + warp) * kMidWarpWords, blockIdx.x * 2 + warp,
                       gridDim.x * 2);
    });
}

// The whole lookup kernel body (count_body = what count_kernel in query.cu wraps), blocks of
// kLookupWarps warps.  qcount [nq+1], qpos [nq], heavy_list [nq], counters [4] zeroed by the caller;
// tmp_cap >= nq * kFixedIds (the fixed places), lookup_fixed_ids() returns kFixedIds.
uint32_t lookup_fixed_ids() { return kFixedIds; }
static int count_emul_wide = 0;
void count_emul_set_wide(int wide) { count_emul_wide = wide; }

void count_emul_run(const uint64_t *list_off, const uint32_t *ids, uint32_t nq, uint32_t subs, uint32_t thr,
                    unsigned grid, uint32_t *qcount, uint64_t *qpos, uint32_t *tmp_ids, uint64_t tmp_cap,
                    uint32_t *heavy_list, unsigned long long *counters) {
    CsrSrc src{list_off, ids, subs};
    CountArgs a;
    a.qcount = qcount;
    a.qpos = qpos;
    a.tmp_ids = tmp_ids;
    a.tmp_cap = tmp_cap;
    a.heavy_list = heavy_list;
    a.counters = counters;
    a.nq = nq;
    a.thr = thr ? thr : 1;
    const size_t block_words = (size_t)kLookupWarps * warp_words(count_emul_wide ? kLookupCapWide : kLookupCap);
    std::vector<uint32_t> smem((size_t)grid * block_words + 4, 0xA5A5A5A5u);
    uint32_t *base = smem.data();
    while (reinterpret_cast<uintptr_t>(base) & 15) ++base;                  // the body stores uint4
    if (count_emul_wide)        // what query.cu launches for n > 64: 4 lists per lane in registers, 2048-id buffer
        emu_launch(grid, kLookupWarps * 32, [&] { count_body<CsrSrc, kRegListsMax, kLookupCapWide>(src, a, base + (size_t)blockIdx.x * block_words); });
    else
        emu_launch(grid, kLookupWarps * 32, [&] { count_body(src, a, base + (size_t)blockIdx.x * block_words); });
}

// count_kernel's sort path for ONE query whose T <= 1024 gathered ids are already laid out: counting
// filter in place (warp_filter_ids, counters in a 256-word area), bitonic sort of the survivors,
// run lengths against the threshold.  out gets the emitted ids, *survivors what the filter kept.
void sortpath_emul_run(const uint32_t *ids, uint32_t T, uint32_t thr, int use_filter, uint32_t *out, uint32_t *R_out,
                       uint32_t *survivors) {
    std::vector<uint32_t> buf(1024 + 256 + 32, 0xA5A5A5A5u);
    for (uint32_t i = 0; i < T; ++i) buf[i] = ids[i];
    uint32_t R_shared = 0, S_shared = 0;
    emu_launch(1, 32, [&] {
        const int lane = threadIdx.x & 31;
        uint32_t *b = buf.data(), *res = b + 1024;
        uint32_t Ts = T;
        if (use_filter && thr > 1 && T > 64) {
            __syncwarp();
            Ts = warp_filter_ids(b, T, thr, res, lane);
        }
        uint32_t P = 32;
        while (P < Ts) P <<= 1;
        for (uint32_t i = Ts + lane; i < P; i += 32) b[i] = kNoId;
        __syncwarp();
        warp_bitonic_smem(b, P, lane);
        uint32_t R = 0;
        for (uint32_t i0 = 0; i0 < Ts; i0 += 32) {
            const uint32_t i = i0 + lane;
            bool ok = false;
            uint32_t v = 0;
            if (i < Ts) {
                v = b[i];
                const bool head = i == 0 || b[i - 1] != v;
                ok = head && (thr <= 1 || (i + thr - 1 < Ts && b[i + thr - 1] == v));
            }
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            __syncwarp();
            if (ok) b[R + __popc(m & ((1u << lane) - 1))] = v;
            R += __popc(m);
            __syncwarp();
        }
        if (lane == 0) {
            R_shared = R;
            S_shared = Ts;
        }
    });
    for (uint32_t i = 0; i < R_shared; ++i) out[i] = buf[i];
    *R_out = R_shared;
    *survivors = S_shared;
}

}  // extern "C"

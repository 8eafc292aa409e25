// Per-hash-function sketch tables on the device.  Replaces
//   MinHashReadFilter::populateHashTables   (src/ReadFilter.cpp:159-172)
//   BBHashMap::initialize                    (src/BBHashMap.cpp:10-99)
// The reference builds, per hash function j, a minimal perfect hash over the
// distinct keys of sketch column j plus CSR arrays keys[] / startPosInReadIds[] /
// readIds[]; keys are stored and verified on lookup (BBHashMap.cpp:105-106), so
// the structure is an exact dictionary key -> list of read ids and its answers do
// not depend on the MPHF.
//
// Here: one open-addressing region of cap+1 16-byte slots {key, val, cnt-1} per
// hash function (cap = 2*reads; the extra slot holds the key that equals the empty
// marker ~0).  A probe is ONE 16-byte load = one 32-byte sector.
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

#include "nsmh_internal.cuh"

namespace nsmh {

constexpr int kBuildCols = 4;      // adjacent hash functions per block (4 x 8 B = one sector per row)
constexpr int kBuildRows = 256;    // rows per block = threads per block

struct BuildArgs {
    const uint64_t *sk;     // [rows][n]
    Slot *slots;            // [n][cap+1]
    uint32_t *ids;
    uint32_t *m_slot, *m_id, *m_rank;   // members that arrived second or later
    uint32_t *g_slot;                   // slots of groups with two or more members
    unsigned int *counters;             // [0] ids cursor, [1] multi members, [2] multi groups
    uint64_t cap;
    uint32_t rows, n;
};

// 128-bit compare-and-swap on a slot; returns the previous contents.
__device__ __forceinline__ void slot_cas(Slot *p, uint64_t exp_lo, uint64_t exp_hi, uint64_t new_lo,
                                         uint64_t new_hi, uint64_t &old_lo, uint64_t &old_hi) {
    asm volatile(
        "{\n\t.reg .b128 d, b, c;\n\t"
        "mov.b128 b, {%2, %3};\n\t"
        "mov.b128 c, {%4, %5};\n\t"
        "atom.relaxed.gpu.global.cas.b128 d, [%6], b, c;\n\t"
        "mov.b128 {%0, %1}, d;\n\t}"
        : "=l"(old_lo), "=l"(old_hi)
        : "l"(exp_lo), "l"(exp_hi), "l"(new_lo), "l"(new_hi), "l"(p)
        : "memory");
}

__device__ __forceinline__ uint64_t slot_key_now(const Slot *p) {
    uint64_t k;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(k) : "l"(&p->key) : "memory");
    return k;
}

// returns the slot index (global) and the element's rank inside its group
__device__ __forceinline__ uint32_t insert_one(const BuildArgs &a, uint32_t l, uint64_t key, uint32_t id,
                                               uint32_t &rank) {
    const uint64_t base = (uint64_t)l * (a.cap + 1);
    Slot *region = a.slots + base;
    if (key == kEmptyKey) {
        // slots start as all-ones: the count field holds (group size - 1), wrapping from ~0
        rank = atomicAdd(&region[a.cap].cntm1, 1u) + 1u;
        if (rank == 0) region[a.cap].val = id;
        return (uint32_t)(base + a.cap);
    }
    uint64_t h = slot_index(key, a.cap);
    for (;;) {
        Slot *p = region + h;
        uint64_t cur = slot_key_now(p);
        if (cur == kEmptyKey) {
            uint64_t old_lo, old_hi;
            slot_cas(p, ~0ULL, ~0ULL, key, (uint64_t)id, old_lo, old_hi);   // {key, val = id, cnt-1 = 0}
            if (old_lo == kEmptyKey) { rank = 0; return (uint32_t)(base + h); }
            cur = old_lo;
        }
        if (cur == key) {
            rank = atomicAdd(&p->cntm1, 1u) + 1u;
            return (uint32_t)(base + h);
        }
        h = h + 1 == a.cap ? 0 : h + 1;
    }
}

__global__ void __launch_bounds__(kBuildRows)
table_insert_kernel(BuildArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t chunks = (a.rows + kBuildRows - 1) / kBuildRows;
    const uint32_t colgroups = (a.n + kBuildCols - 1) / kBuildCols;
    const uint64_t units = (uint64_t)chunks * colgroups;
    for (uint64_t u = blockIdx.x; u < units; u += gridDim.x) {
        const uint32_t cg = (uint32_t)(u / chunks);
        const uint32_t row = (uint32_t)(u % chunks) * kBuildRows + threadIdx.x;
        const uint32_t l0 = cg * kBuildCols;
        const bool live = row < a.rows;
        uint64_t keys[kBuildCols];
#pragma unroll
        for (int j = 0; j < kBuildCols; ++j)
            keys[j] = (live && l0 + j < a.n) ? __ldg(a.sk + (size_t)row * a.n + l0 + j) : 0;
#pragma unroll
        for (int j = 0; j < kBuildCols; ++j) {
            uint32_t rank = 0, s = 0;
            if (live && l0 + j < a.n) s = insert_one(a, l0 + j, keys[j], row, rank);
            // warp-aggregated append of the members that were not first in their group
            const uint32_t m = __ballot_sync(0xffffffffu, rank >= 1);
            if (m) {
                const int leader = __ffs(m) - 1;
                uint32_t b = 0;
                if (lane == leader) b = atomicAdd(a.counters + 1, (unsigned int)__popc(m));
                b = __shfl_sync(0xffffffffu, b, leader);
                if (rank >= 1) {
                    const uint32_t pos = b + __popc(m & ((1u << lane) - 1));
                    a.m_slot[pos] = s;
                    a.m_id[pos] = row;
                    a.m_rank[pos] = rank;
                    if (rank == 1) a.g_slot[atomicAdd(a.counters + 2, 1u)] = s;
                }
            }
        }
    }
}

// groups of two or more: allocate the id range, move the inlined first id into it
__global__ void __launch_bounds__(256)
table_groups_kernel(BuildArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t groups = a.counters[2];
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (groups + stride - 1) / stride;
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t r = 0; r < rounds; ++r, g += stride) {   // whole warps stay in the loop for the shuffles
        uint32_t need = 0, s = 0;
        if (g < groups) {
            s = a.g_slot[g];
            need = a.slots[s].cntm1 + 1u;
        }
        uint32_t incl = need;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t base = 0;
        if (total) {
            if (lane == 31) base = atomicAdd(a.counters, total);
            base = __shfl_sync(0xffffffffu, base, 31);
        }
        if (need) {
            const uint32_t b = base + incl - need;
            a.ids[b] = a.slots[s].val;
            a.slots[s].val = b;
        }
    }
}

__global__ void __launch_bounds__(256)
table_fill_kernel(BuildArgs a) {
    const uint32_t members = a.counters[1];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < members; i += gridDim.x * blockDim.x)
        a.ids[a.slots[a.m_slot[i]].val + a.m_rank[i]] = a.m_id[i];
}

__global__ void __launch_bounds__(256)
table_count_keys_kernel(const Slot *__restrict__ slots, uint64_t nslots, uint32_t *out) {
    uint32_t local = 0;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < nslots;
         t += (uint64_t)gridDim.x * blockDim.x)
        local += slots[t].cntm1 != 0xFFFFFFFFu;
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

int build_tables(nsmh_ctx *c) {
    Tables &T = c->tables;
    cudaStream_t s = c->stream;
    const uint32_t n = c->n, rows = c->table_reads;
    T.built = false;
    const uint64_t cap = std::max<uint64_t>(16, 2ULL * rows);   // load factor <= 0.5
    const uint64_t nslots = (uint64_t)n * (cap + 1);
    const uint64_t items = (uint64_t)rows * n;
    if (nslots >= (1ULL << 32) || items >= (1ULL << 32))
        return fail(NSMH_EINVAL, "build: reads*n too large for 32-bit slot indices");
    T.cap = cap;
    T.table_reads = rows;
    const size_t ni = (size_t)(items ? items : 1);
    NSMH_TRY(T.slots.ensure(nslots * sizeof(Slot), s));
    NSMH_TRY(T.ids.ensure(ni * sizeof(uint32_t), s));
    NSMH_TRY(c->build_multi.ensure(ni * 4 * sizeof(uint32_t), s));
    NSMH_TRY(c->build_tmp.ensure(64, s));
    NSMH_CK(cudaMemsetAsync(T.slots.p, 0xFF, nslots * sizeof(Slot), s));
    NSMH_CK(cudaMemsetAsync(c->build_tmp.p, 0, 4 * sizeof(unsigned int), s));
    if (items) {
        BuildArgs a;
        a.sk = c->table_sketches;
        a.slots = T.slots.as<Slot>();
        a.ids = T.ids.as<uint32_t>();
        a.m_slot = c->build_multi.as<uint32_t>();
        a.m_id = a.m_slot + ni;
        a.m_rank = a.m_id + ni;
        a.g_slot = a.m_rank + ni;
        a.counters = c->build_tmp.as<unsigned int>();
        a.cap = cap;
        a.rows = rows;
        a.n = n;
        const uint64_t units = (uint64_t)((rows + kBuildRows - 1) / kBuildRows) * ((n + kBuildCols - 1) / kBuildCols);
        const int blocks = (int)std::min<uint64_t>(units, (uint64_t)c->num_sms * 8);
        table_insert_kernel<<<blocks, kBuildRows, 0, s>>>(a);
        NSMH_CK(cudaGetLastError());
        table_groups_kernel<<<c->num_sms * 2, 256, 0, s>>>(a);
        NSMH_CK(cudaGetLastError());
        table_fill_kernel<<<c->num_sms * 2, 256, 0, s>>>(a);
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
    table_count_keys_kernel<<<c->num_sms, 256, 0, s>>>(T.slots.as<Slot>() + (uint64_t)j * (T.cap + 1),
                                                       T.cap + 1, d);
    ++c->launches;
    NSMH_CK(cudaGetLastError());
    NSMH_CK(cudaMemcpyAsync(out, d, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    NSMH_CK(cudaStreamSynchronize(s));
    return NSMH_OK;
}

} // namespace nsmh

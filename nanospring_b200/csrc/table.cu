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
// hash function (cap = power of two >= 2*reads; the extra slot holds the key that
// equals the empty marker ~0).  A probe is ONE 16-byte load = one 32-byte sector.
//   pass 1  insert : linear probing with atomicCAS on the key; atomicAdd on the
//                    count returns the element's rank inside its group
//   pass 2  leader : rank-0 elements finish their group: a group of one stores the
//                    read id in the slot itself (most groups; no second access at
//                    lookup time), larger groups get a range of `ids` from a
//                    warp-aggregated atomic cursor
//   pass 3  fill   : the remaining elements write ids[val + rank]
// Inside a group the id order is arbitrary; every consumer sorts (ReadFilter.cpp:73).
#include "nsmh_internal.cuh"

namespace nsmh {

__device__ __forceinline__ uint64_t slot_hash(uint64_t key, uint32_t log2cap) {
    return (key * 0x9E3779B97F4A7C15ULL) >> (64 - log2cap);
}

// one thread per (row, hash) item of the sketch matrix
__global__ void __launch_bounds__(256)
table_insert_kernel(const uint64_t *__restrict__ sk, uint64_t items, uint32_t n, uint64_t cap,
                    uint32_t log2cap, Slot *__restrict__ slots, uint32_t *__restrict__ item_slot,
                    uint32_t *__restrict__ item_rank) {
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < items;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t l = (uint32_t)(t % n);
        const uint64_t key = sk[t];
        const uint64_t base = (uint64_t)l * (cap + 1);
        uint64_t s;
        if (key == kEmptyKey) {
            s = base + cap;
        } else {
            uint64_t h = slot_hash(key, log2cap);
            for (;;) {
                s = base + h;
                unsigned long long *kp = reinterpret_cast<unsigned long long *>(&slots[s].key);
                unsigned long long prev = *reinterpret_cast<volatile unsigned long long *>(kp);
                if (prev == key) break;
                if (prev == kEmptyKey) {
                    prev = atomicCAS(kp, (unsigned long long)kEmptyKey, (unsigned long long)key);
                    if (prev == kEmptyKey || prev == key) break;
                }
                h = (h + 1) & (cap - 1);
            }
        }
        item_slot[t] = (uint32_t)s;
        // slots start as all-ones: the count field holds (group size - 1), wrapping from ~0
        item_rank[t] = atomicAdd(&slots[s].cntm1, 1u) + 1u;
    }
}

__global__ void __launch_bounds__(256)
table_leader_kernel(uint64_t items, uint32_t n, Slot *__restrict__ slots,
                    const uint32_t *__restrict__ item_slot, const uint32_t *__restrict__ item_rank,
                    uint32_t *__restrict__ ids, unsigned int *__restrict__ cursor) {
    const int lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (items + stride - 1) / stride;
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (uint64_t r = 0; r < rounds; ++r, t += stride) {   // whole warps stay in the loop for the shuffles
        uint32_t need = 0, s = 0;
        if (t < items && item_rank[t] == 0) {
            s = item_slot[t];
            uint32_t c = slots[s].cntm1 + 1u;
            if (c == 1) slots[s].val = (uint32_t)(t / n);
            else need = c;
        }
        // warp-aggregated allocation of id ranges for groups of two or more
        uint32_t incl = need;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t base = 0;
        if (total) {
            if (lane == 31) base = atomicAdd(cursor, total);
            base = __shfl_sync(0xffffffffu, base, 31);
        }
        if (need) {
            uint32_t b = base + incl - need;
            slots[s].val = b;
            ids[b] = (uint32_t)(t / n);
        }
    }
}

__global__ void __launch_bounds__(256)
table_fill_kernel(uint64_t items, uint32_t n, const Slot *__restrict__ slots,
                  const uint32_t *__restrict__ item_slot, const uint32_t *__restrict__ item_rank,
                  uint32_t *__restrict__ ids) {
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < items;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = item_rank[t];
        if (r) ids[slots[item_slot[t]].val + r] = (uint32_t)(t / n);
    }
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
    uint32_t log2cap = 4;
    while ((1ULL << log2cap) < 2ULL * rows) ++log2cap;
    const uint64_t cap = 1ULL << log2cap;
    const uint64_t nslots = (uint64_t)n * (cap + 1);
    const uint64_t items = (uint64_t)rows * n;
    if (nslots >= (1ULL << 32) || items >= (1ULL << 32))
        return fail(NSMH_EINVAL, "build: reads*n too large for 32-bit slot indices");
    T.cap = cap;
    T.log2cap = log2cap;
    T.table_reads = rows;
    NSMH_TRY(T.slots.ensure(nslots * sizeof(Slot), s));
    NSMH_TRY(T.ids.ensure((items ? items : 1) * sizeof(uint32_t), s));
    NSMH_TRY(c->item_slot.ensure((items ? items : 1) * sizeof(uint32_t), s));
    NSMH_TRY(c->item_rank.ensure((items ? items : 1) * sizeof(uint32_t), s));
    NSMH_TRY(c->build_tmp.ensure(64, s));
    NSMH_CK(cudaMemsetAsync(T.slots.p, 0xFF, nslots * sizeof(Slot), s));
    NSMH_CK(cudaMemsetAsync(c->build_tmp.p, 0, sizeof(unsigned int), s));
    if (items) {
        int blocks = (int)((items + 255) / 256 < (uint64_t)c->num_sms * 16 ? (items + 255) / 256
                                                                           : (uint64_t)c->num_sms * 16);
        table_insert_kernel<<<blocks, 256, 0, s>>>(c->table_sketches, items, n, cap, log2cap,
                                                   T.slots.as<Slot>(), c->item_slot.as<uint32_t>(),
                                                   c->item_rank.as<uint32_t>());
        NSMH_CK(cudaGetLastError());
        table_leader_kernel<<<blocks, 256, 0, s>>>(items, n, T.slots.as<Slot>(),
                                                   c->item_slot.as<uint32_t>(), c->item_rank.as<uint32_t>(),
                                                   T.ids.as<uint32_t>(), c->build_tmp.as<unsigned int>());
        NSMH_CK(cudaGetLastError());
        table_fill_kernel<<<blocks, 256, 0, s>>>(items, n, T.slots.as<Slot>(), c->item_slot.as<uint32_t>(),
                                                 c->item_rank.as<uint32_t>(), T.ids.as<uint32_t>());
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

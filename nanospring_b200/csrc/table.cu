// Per-hash-function sketch tables on the device.  Replaces
//   MinHashReadFilter::populateHashTables   (src/ReadFilter.cpp:159-172)
//   BBHashMap::initialize                    (src/BBHashMap.cpp:10-99)
// The reference builds, per hash function j, a minimal perfect hash over the
// distinct keys of sketch column j plus CSR arrays keys[] / startPosInReadIds[] /
// readIds[]; keys are stored and verified on lookup (BBHashMap.cpp:105-106), so
// the structure is an exact dictionary key -> list of read ids and its answers do
// not depend on the MPHF.  Here: one open-addressing region of cap+1
// slots per hash function (cap = power of two >= 2*reads; the extra slot holds the
// key that equals the empty marker), linear probing with atomicCAS, group sizes by
// atomicAdd whose return value is the element's rank, one prefix sum, one scatter.
// Inside a group the id order is arbitrary; every consumer sorts (ReadFilter.cpp:73).
#include "nsmh_internal.cuh"

namespace nsmh {

__device__ __forceinline__ uint64_t slot_hash(uint64_t key, uint32_t log2cap) {
    return (key * 0x9E3779B97F4A7C15ULL) >> (64 - log2cap);
}

// one thread per (row, hash) item of the sketch matrix
__global__ void __launch_bounds__(256)
table_insert_kernel(const uint64_t *__restrict__ sk, uint64_t items, uint32_t n, uint64_t cap,
                    uint32_t log2cap, unsigned long long *__restrict__ keys,
                    uint32_t *__restrict__ cnt, uint32_t *__restrict__ item_slot,
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
                unsigned long long prev = keys[s];
                if (prev == key) break;
                if (prev == kEmptyKey) {
                    prev = atomicCAS(keys + s, (unsigned long long)kEmptyKey, (unsigned long long)key);
                    if (prev == kEmptyKey || prev == key) break;
                }
                h = (h + 1) & (cap - 1);
            }
        }
        item_slot[t] = (uint32_t)s;
        item_rank[t] = atomicAdd(cnt + s, 1u);
    }
}

__global__ void __launch_bounds__(256)
table_fill_kernel(uint64_t items, uint32_t n, const uint32_t *__restrict__ begin,
                  const uint32_t *__restrict__ item_slot, const uint32_t *__restrict__ item_rank,
                  uint32_t *__restrict__ ids) {
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < items;
         t += (uint64_t)gridDim.x * blockDim.x)
        ids[begin[item_slot[t]] + item_rank[t]] = (uint32_t)(t / n);
}

__global__ void __launch_bounds__(256)
table_count_keys_kernel(const uint32_t *__restrict__ cnt, uint64_t slots, uint32_t *out) {
    uint32_t local = 0;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < slots;
         t += (uint64_t)gridDim.x * blockDim.x)
        local += cnt[t] != 0;
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
    const uint64_t slots = (uint64_t)n * (cap + 1);
    const uint64_t items = (uint64_t)rows * n;
    if (slots >= (1ULL << 32) || items >= (1ULL << 32))
        return fail(NSMH_EINVAL, "build: reads*n too large for 32-bit slot indices");
    T.cap = cap;
    T.log2cap = log2cap;
    T.table_reads = rows;
    NSMH_TRY(T.keys.ensure(slots * sizeof(uint64_t), s));
    NSMH_TRY(T.cnt.ensure((slots + 1) * sizeof(uint32_t), s));
    NSMH_TRY(T.begin.ensure((slots + 1) * sizeof(uint32_t), s));
    NSMH_TRY(T.ids.ensure((items ? items : 1) * sizeof(uint32_t), s));
    NSMH_TRY(c->item_slot.ensure((items ? items : 1) * sizeof(uint32_t), s));
    NSMH_TRY(c->item_rank.ensure((items ? items : 1) * sizeof(uint32_t), s));
    NSMH_CK(cudaMemsetAsync(T.keys.p, 0xFF, slots * sizeof(uint64_t), s));
    NSMH_CK(cudaMemsetAsync(T.cnt.p, 0, (slots + 1) * sizeof(uint32_t), s));
    if (items) {
        int blocks = (int)((items + 255) / 256 < (uint64_t)c->num_sms * 16 ? (items + 255) / 256
                                                                           : (uint64_t)c->num_sms * 16);
        table_insert_kernel<<<blocks, 256, 0, s>>>(c->table_sketches, items, n, cap, log2cap,
                                                   T.keys.as<unsigned long long>(), T.cnt.as<uint32_t>(),
                                                   c->item_slot.as<uint32_t>(), c->item_rank.as<uint32_t>());
        ++c->launches;
        NSMH_CK(cudaGetLastError());
        size_t tmp_bytes = 0;
        NSMH_CK(cub_exclusive_sum_u32(nullptr, tmp_bytes, T.cnt.as<uint32_t>(), T.begin.as<uint32_t>(),
                                      slots + 1, s));
        NSMH_TRY(c->build_tmp.ensure(tmp_bytes, s));
        NSMH_CK(cub_exclusive_sum_u32(c->build_tmp.p, tmp_bytes, T.cnt.as<uint32_t>(),
                                      T.begin.as<uint32_t>(), slots + 1, s));
        c->launches += 2;
        table_fill_kernel<<<blocks, 256, 0, s>>>(items, n, T.begin.as<uint32_t>(),
                                                 c->item_slot.as<uint32_t>(), c->item_rank.as<uint32_t>(),
                                                 T.ids.as<uint32_t>());
        ++c->launches;
        NSMH_CK(cudaGetLastError());
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
    uint32_t *d = c->build_tmp.as<uint32_t>();
    NSMH_CK(cudaMemsetAsync(d, 0, sizeof(uint32_t), s));
    table_count_keys_kernel<<<c->num_sms, 256, 0, s>>>(T.cnt.as<uint32_t>() + (uint64_t)j * (T.cap + 1),
                                                       T.cap + 1, d);
    ++c->launches;
    NSMH_CK(cudaGetLastError());
    NSMH_CK(cudaMemcpyAsync(out, d, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    NSMH_CK(cudaStreamSynchronize(s));
    return NSMH_OK;
}

} // namespace nsmh

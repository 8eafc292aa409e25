// Device code of the table build (see table.cu for the structure and the reference lines it
// replaces).  Kept free of runtime-API includes so that tests/cpp/table_host_emul.cpp can compile the
// SAME kernels for the host (NSMH_HOST_EMUL: whole blocks of OS threads, the three PTX accesses
// replaced by host atomics) and check the tables against the sketch matrix without a GPU.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "nsmh_constants.h"
#include "nsmh_ldst.cuh"

namespace nsmh {

constexpr int kBuildCols = 4;      // adjacent hash functions per block (4 x 8 B = one sector per row)
constexpr int kBuildRows = 256;    // rows per block = threads per block

struct BuildArgs {
    const uint64_t *sk;     // [rows][n]
    Slot *slots;            // [n][cap+1]
    uint32_t *ids;
    // Members that arrived second or later, and the slots of their groups.  Every work unit of
    // the insert kernel appends to its own segment of seg_cap entries (shared-memory cursor):
    // a single global cursor would serialise ~10^5 same-address atomics per build.
    uint32_t *m_slot, *m_id, *m_rank;
    uint32_t *g_slot;
    unsigned int *counters;             // [0] ids cursor, [1] next work unit of the insert kernel
    unsigned int *seg_count;            // [2*segments] members, groups of every segment
    uint64_t cap;
    uint32_t rows, n, seg_cap, segments;
    uint32_t skip_empty;    // 1: all-ones keys are left to table_insert_list_kernel (their values are still being computed)
};

#ifndef NSMH_HOST_EMUL
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

// both slots of a bucket as they are now (256-bit L2-coherent load)
__device__ __forceinline__ void bucket_now(const Slot *p, uint64_t &k0, uint64_t &v0, uint64_t &k1, uint64_t &v1) {
    asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(k0), "=l"(v0), "=l"(k1), "=l"(v1) : "l"(p) : "memory");
}

__device__ __forceinline__ uint64_t slot_key_now(const Slot *p) {
    uint64_t k;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(k) : "l"(&p->key) : "memory");
    return k;
}

#else
// host emulation (tests/cpp/table_host_emul.cpp): the same three accesses with the host's atomics
// (cmpxchg16b through libatomic for the 128-bit compare-and-swap)
__device__ __forceinline__ void slot_cas(Slot *p, uint64_t exp_lo, uint64_t exp_hi, uint64_t new_lo,
                                         uint64_t new_hi, uint64_t &old_lo, uint64_t &old_hi) {
    unsigned __int128 expected = ((unsigned __int128)exp_hi << 64) | exp_lo;
    const unsigned __int128 desired = ((unsigned __int128)new_hi << 64) | new_lo;
    __atomic_compare_exchange_n(reinterpret_cast<unsigned __int128 *>(p), &expected, desired, false, __ATOMIC_RELAXED,
                                __ATOMIC_RELAXED);
    old_lo = (uint64_t)expected;
    old_hi = (uint64_t)(expected >> 64);
}
__device__ __forceinline__ void bucket_now(const Slot *p, uint64_t &k0, uint64_t &v0, uint64_t &k1, uint64_t &v1) {
    const uint64_t *q = reinterpret_cast<const uint64_t *>(p);
    k0 = __atomic_load_n(q, __ATOMIC_RELAXED);
    v0 = __atomic_load_n(q + 1, __ATOMIC_RELAXED);
    k1 = __atomic_load_n(q + 2, __ATOMIC_RELAXED);
    v1 = __atomic_load_n(q + 3, __ATOMIC_RELAXED);
}
__device__ __forceinline__ uint64_t slot_key_now(const Slot *p) { return __atomic_load_n(&p->key, __ATOMIC_RELAXED); }
#endif

// A block takes a unit of 256 rows x 4 adjacent columns; a thread owns the 4 keys of one row
// (one 32-byte sector of the sketch matrix) and works through them as a small state machine:
// one loop iteration = one compare-and-swap of the thread's current key, and a thread whose
// key is placed moves on to its next key without waiting for the other lanes.  A warp
// therefore pays the longest SUM of probe steps over 4 keys among its lanes, not 4 times the
// longest probe sequence, and the loop body stays small (the kernel is bound by the L2
// atomic rate, ~50 G atomics/s whatever their width: tools/micro/atom_bench.cu).
// Buckets are the two slots of one 32-byte sector.
// 5 resident blocks per SM (48 registers, 4 bytes of spill) instead of the 4 the compiler's 50 registers allow: the
// kernel's time follows its resident warps (build 0.244 -> 0.228 ms; 6 and 8 blocks spill more and are slower again,
// profiles/r2_sketch_build_overlap_s31.txt).
#ifndef NSMH_INSERT_MIN_BLOCKS
#define NSMH_INSERT_MIN_BLOCKS 5
#endif
__global__ void __launch_bounds__(kBuildRows, NSMH_INSERT_MIN_BLOCKS)
table_insert_kernel(BuildArgs a) {
    __shared__ unsigned int s_count[2];
    const int lane = threadIdx.x & 31;
    const uint32_t chunks = (a.rows + kBuildRows - 1) / kBuildRows;
    const uint32_t colgroups = (a.n + kBuildCols - 1) / kBuildCols;
    const uint32_t units = chunks * colgroups;     // < 2^32: checked by build_tables
    const uint64_t nb = a.cap >> 1;
    const uint64_t stride = region_stride(a.cap);
    __shared__ unsigned int s_unit;
    if (threadIdx.x < 2) s_count[threadIdx.x] = 0;

    // Units are handed out dynamically, in order (the regions in use at any time stay few and L2 resident;
    // a static round-robin left 17 % of the warp samples of the first version waiting at the exit).
    // Every unit appends to its own segment of seg_cap = rows x columns entries.
    uint32_t prev = 0xFFFFFFFFu;
    for (;;) {
        __syncthreads();                    // the previous unit's appends are done, everybody has read s_unit
        if (threadIdx.x == 0) {
            if (prev != 0xFFFFFFFFu) {
                a.seg_count[2 * (size_t)prev] = s_count[0];
                a.seg_count[2 * (size_t)prev + 1] = s_count[1];
                s_count[0] = s_count[1] = 0;
            }
            s_unit = atomicAdd(a.counters + 1, 1u);
        }
        __syncthreads();
        const uint32_t u = s_unit;
        if (u >= units) break;
        prev = u;
        const size_t seg0 = (size_t)u * a.seg_cap;
        const uint32_t cg = u / chunks;
        const uint32_t row = (u - cg * chunks) * kBuildRows + threadIdx.x;
        const uint32_t l0 = cg * kBuildCols;
        const uint32_t nj = row < a.rows ? min((uint32_t)kBuildCols, a.n - l0) : 0u;   // keys of this thread
        uint64_t keys[kBuildCols];
#pragma unroll
        for (int j = 0; j < kBuildCols; ++j)
            keys[j] = (uint32_t)j < nj ? __ldg(a.sk + (size_t)row * a.n + l0 + j) : 0;
        // ---- fast round: the home buckets of all keys are peeked together, then all claims go out together,
        // so a thread waits for two memory round trips per unit instead of two per key (the kernel is bound
        // by the latency of these dependent accesses, not by their number).  What the round cannot settle
        // (home bucket full, a slot lost to another key meanwhile) is left to the loop below.
        bool done[kBuildCols];
        uint32_t frank[kBuildCols], fslot[kBuildCols];
        {
            uint64_t hk0[kBuildCols], hk1[kBuildCols], got[kBuildCols];
            uint32_t hb[kBuildCols];
            int state[kBuildCols];          // 0 nothing, 1 claim sent, 2 the key is there already
#pragma unroll
            for (int jj = 0; jj < kBuildCols; ++jj) {
                done[jj] = a.skip_empty && (uint32_t)jj < nj && keys[jj] == kEmptyKey;      // not this kernel's business
                frank[jj] = fslot[jj] = 0;
                state[jj] = 0;
                hk0[jj] = hk1[jj] = 0;
                hb[jj] = 0;
                if ((uint32_t)jj < nj && keys[jj] != kEmptyKey) {
                    uint64_t v0, v1;
                    hb[jj] = (uint32_t)slot_index(keys[jj], nb);
                    bucket_now(a.slots + (uint64_t)(l0 + jj) * stride + 2 * (uint64_t)hb[jj], hk0[jj], v0, hk1[jj], v1);
                }
            }
#pragma unroll
            for (int jj = 0; jj < kBuildCols; ++jj) {
                got[jj] = 0;
                if ((uint32_t)jj < nj && keys[jj] != kEmptyKey) {
                    const uint64_t key = keys[jj];
                    const int t = (hk0[jj] == key || hk0[jj] == kEmptyKey) ? 0 : (hk1[jj] == key || hk1[jj] == kEmptyKey) ? 1 : 2;
                    if (t != 2) {
                        Slot *p = a.slots + (uint64_t)(l0 + jj) * stride + 2 * (uint64_t)hb[jj] + t;
                        fslot[jj] = (uint32_t)(p - a.slots);
                        if ((t ? hk1[jj] : hk0[jj]) == kEmptyKey) {
                            uint64_t old_hi;
                            slot_cas(p, ~0ULL, ~0ULL, key, (uint64_t)row, got[jj], old_hi);   // {key, val = id, cnt-1 = 0}
                            state[jj] = 1;
                        } else {
                            state[jj] = 2;
                        }
                    }
                }
            }
#pragma unroll
            for (int jj = 0; jj < kBuildCols; ++jj) {
                if (state[jj] == 1) {
                    if (got[jj] == kEmptyKey) done[jj] = true;                 // rank 0: first of its group
                    else if (got[jj] == keys[jj]) state[jj] = 2;               // the same key arrived meanwhile
                }
                if (state[jj] == 2) {
                    frank[jj] = atomicAdd(&a.slots[fslot[jj]].cntm1, 1u) + 1u;
                    done[jj] = true;
                }
            }
        }
        // members that were not first in their group (warp-aggregated append, as in the loop below)
#pragma unroll
        for (int jj = 0; jj < kBuildCols; ++jj) {
            const bool member = done[jj] && frank[jj] >= 1;
            const uint32_t m = __ballot_sync(0xffffffffu, member);
            if (m) {
                const int leader = __ffs(m) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(&s_count[0], (unsigned int)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (member) {
                    const size_t pos = seg0 + base + __popc(m & ((1u << lane) - 1));
                    a.m_slot[pos] = fslot[jj];
                    a.m_id[pos] = row;
                    a.m_rank[pos] = frank[jj];
                    if (frank[jj] == 1) a.g_slot[seg0 + atomicAdd(&s_count[1], 1u)] = fslot[jj];
                }
            }
        }
        uint32_t j = 0;
        while (j < nj && (j == 0 ? done[0] : j == 1 ? done[1] : j == 2 ? done[2] : done[3])) ++j;
        if (__all_sync(0xffffffffu, j >= nj)) continue;
        bool fresh = true;        // the current key has not been probed yet
        uint64_t key = 0, b = 0;
        Slot *region = a.slots;
        for (;;) {
            uint32_t rank = 0, s = 0;
            bool completed = false;
            if (j < nj) {
                if (fresh) {
                    key = j == 0 ? keys[0] : j == 1 ? keys[1] : j == 2 ? keys[2] : keys[3];
                    region = a.slots + (uint64_t)(l0 + j) * stride;
                    b = slot_index(key, nb);
                    fresh = false;
                }
                Slot *p;
                if (key == kEmptyKey) {
                    // the extra slot; slots start as all-ones: the count field holds
                    // (group size - 1), wrapping from ~0
                    p = region + a.cap;
                    rank = atomicAdd(&p->cntm1, 1u) + 1u;
                    if (rank == 0) p->val = row;
                    completed = true;
                } else {
                    // peek at the bucket (one 32-byte load, L2-coherent), then one atomic on the slot
                    // that holds the key or is the first free one.  The peek may be stale; the
                    // compare-and-swap decides.  (A load that misses L2 followed by an atomic that
                    // hits is faster than an atomic that misses: tools/micro/atom_bench.cu.)
                    uint64_t k0, v0, k1, v1;
                    bucket_now(region + 2 * b, k0, v0, k1, v1);
                    const int t = (k0 == key || k0 == kEmptyKey) ? 0 : (k1 == key || k1 == kEmptyKey) ? 1 : 2;
                    p = region + 2 * b + (t & 1);
                    if (t == 2) {
                        b = b + 1 == nb ? 0 : b + 1;
                    } else {
                        uint64_t cur = t ? k1 : k0;
                        if (cur == kEmptyKey) {
                            uint64_t old_hi;
                            slot_cas(p, ~0ULL, ~0ULL, key, (uint64_t)row, cur, old_hi);   // {key, val = id, cnt-1 = 0}
                            if (cur == kEmptyKey) completed = true;                  // rank 0: first of its group
                        }
                        if (!completed && cur == key) {
                            rank = atomicAdd(&p->cntm1, 1u) + 1u;
                            completed = true;
                        }
                        // otherwise another key took the slot meanwhile: look again
                    }
                }
                if (completed) {
                    s = (uint32_t)(p - a.slots);
                    do ++j; while (j < nj && (j == 1 ? done[1] : j == 2 ? done[2] : j == 3 ? done[3] : false));
                    fresh = true;
                }
            }
            // warp-aggregated append of the members that were not first in their group
            const uint32_t m = __ballot_sync(0xffffffffu, completed && rank >= 1);
            if (m) {
                const int leader = __ffs(m) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(&s_count[0], (unsigned int)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (completed && rank >= 1) {
                    const size_t pos = seg0 + base + __popc(m & ((1u << lane) - 1));
                    a.m_slot[pos] = s;
                    a.m_id[pos] = row;
                    a.m_rank[pos] = rank;
                    if (rank == 1) a.g_slot[seg0 + atomicAdd(&s_count[1], 1u)] = s;
                }
            }
            if (__all_sync(0xffffffffu, j >= nj)) break;
        }
    }
}

// The entries a build with skip_empty left out, from a list: entry list[e] = row * n + hash function, its key in
// vals[e] (stored into the sketch matrix on the way).  Same claims and counters as above, one key per thread; a
// member that is not the first of its group is appended to the segment of the work unit its (row, column) belongs
// to (the unit's own appends are over: this kernel runs after table_insert_kernel), so the two kernels below need
// not know about it.  Thousands of entries, not millions: global atomics on the segment counters are fine here.
__global__ void __launch_bounds__(256)
table_insert_list_kernel(BuildArgs a, const uint32_t *__restrict__ list, const unsigned int *__restrict__ count,
                         const uint64_t *__restrict__ vals, uint64_t *__restrict__ sk) {
    const uint32_t todo = *count;
    const uint32_t chunks = (a.rows + kBuildRows - 1) / kBuildRows;
    const uint64_t nb = a.cap >> 1;
    const uint64_t stride = region_stride(a.cap);
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < todo; e += gridDim.x * blockDim.x) {
        const uint32_t t = list[e];
        const uint32_t row = t / a.n, l = t - row * a.n;
        const uint64_t key = vals[e];
        if (key != kEmptyKey) sk[t] = key;
        Slot *region = a.slots + (uint64_t)l * stride;
        Slot *p;
        uint32_t rank = 0;
        if (key == kEmptyKey) {
            p = region + a.cap;                 // the extra slot, see table_insert_kernel
            rank = atomicAdd(&p->cntm1, 1u) + 1u;
            if (rank == 0) p->val = row;
        } else {
            uint64_t b = slot_index(key, nb);
            for (;;) {
                uint64_t k0, v0, k1, v1;
                bucket_now(region + 2 * b, k0, v0, k1, v1);
                const int tt = (k0 == key || k0 == kEmptyKey) ? 0 : (k1 == key || k1 == kEmptyKey) ? 1 : 2;
                if (tt == 2) { b = b + 1 == nb ? 0 : b + 1; continue; }
                p = region + 2 * b + tt;
                uint64_t cur = tt ? k1 : k0;
                if (cur == kEmptyKey) {
                    uint64_t old_hi;
                    slot_cas(p, ~0ULL, ~0ULL, key, (uint64_t)row, cur, old_hi);     // {key, val = id, cnt-1 = 0}
                    if (cur == kEmptyKey) break;                                    // rank 0: first of its group
                }
                if (cur == key) { rank = atomicAdd(&p->cntm1, 1u) + 1u; break; }
                // another key took the slot meanwhile: look again
            }
        }
        if (rank >= 1) {
            const uint32_t u = (l / kBuildCols) * chunks + row / kBuildRows;
            const size_t seg0 = (size_t)u * a.seg_cap;
            const size_t pos = seg0 + atomicAdd(&a.seg_count[2 * (size_t)u], 1u);
            const uint32_t sidx = (uint32_t)(p - a.slots);
            a.m_slot[pos] = sidx;
            a.m_id[pos] = row;
            a.m_rank[pos] = rank;
            if (rank == 1) a.g_slot[seg0 + atomicAdd(&a.seg_count[2 * (size_t)u + 1], 1u)] = sidx;
        }
    }
}

// groups of two or more: allocate the id range, move the inlined first id into it.
// A warp per segment (one per work unit of the insert kernel, most of them nearly empty).
__global__ void __launch_bounds__(256)
table_groups_kernel(BuildArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); seg < a.segments; seg += warps) {
        const uint32_t groups = a.seg_count[2 * (size_t)seg + 1];
        const uint32_t *g_slot = a.g_slot + (size_t)seg * a.seg_cap;
        for (uint32_t g0 = 0; g0 < groups; g0 += 32) {        // whole warps stay in the loop for the shuffles
            const uint32_t g = g0 + lane;
            uint32_t need = 0, s = 0;
            if (g < groups) {
                s = g_slot[g];
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
}

__global__ void __launch_bounds__(256)
table_fill_kernel(BuildArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); seg < a.segments; seg += warps) {
        const uint32_t members = a.seg_count[2 * (size_t)seg];
        const size_t seg0 = (size_t)seg * a.seg_cap;
        for (uint32_t i = lane; i < members; i += 32)
            a.ids[a.slots[a.m_slot[seg0 + i]].val + a.m_rank[seg0 + i]] = a.m_id[seg0 + i];
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

} // namespace nsmh

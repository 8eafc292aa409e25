// 256-bit global loads / stores (one 32-byte sector per access) shared by the lookup and multi-GPU
// kernels; plain loads / stores when the kernels are compiled for the host emulation.
#pragma once
#include <stdint.h>
#ifdef NSMH_HOST_EMUL
#include <string.h>
#endif

namespace nsmh {

#ifndef NSMH_HOST_EMUL
__device__ __forceinline__ void ldg256(const void *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) {
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
__device__ __forceinline__ void stg256(void *p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
// Hint: bring `bytes` (multiple of 16) starting at p (16-byte aligned) into L2, asynchronously, by ONE thread
// (cp.async.bulk.prefetch.L2, sm_90+).  The table kernels work through the hash functions in order and ask
// for the slice of the NEXT regions while they hammer the current ones with random accesses.
__device__ __forceinline__ void l2_prefetch_bulk(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void l2_prefetch_line(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// Bulk copy shared -> global (cp.async.bulk, sm_90+): ONE thread hands `bytes` (multiple of 16; both addresses
// 16-byte aligned) to the copy engine, which writes full lines - to local memory or to a peer's over NVLink.
// Groups complete in order: wait_read<N> returns when all but the N youngest groups have READ their source.
__device__ __forceinline__ void bulk_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void *dst, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"((uint32_t)__cvta_generic_to_shared(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
#else
// host emulation: the copy happens at once, by the calling thread
__device__ __forceinline__ void bulk_store_fence() {}
__device__ __forceinline__ void bulk_store(void *dst, const void *src_smem, uint32_t bytes) { memcpy(dst, src_smem, bytes); }
__device__ __forceinline__ void bulk_store_commit() {}
template <int N>
__device__ __forceinline__ void bulk_store_wait_read() {}
__device__ __forceinline__ void bulk_store_wait_all() {}
__device__ __forceinline__ void l2_prefetch_bulk(const void *, uint32_t) {}
__device__ __forceinline__ void l2_prefetch_line(const void *) {}
__device__ __forceinline__ void ldg256(const void *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) {
    const uint64_t *q = static_cast<const uint64_t *>(p);
    a = q[0]; b = q[1]; c = q[2]; d = q[3];
}
__device__ __forceinline__ void stg256(void *p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    uint64_t *q = static_cast<uint64_t *>(p);
    q[0] = a; q[1] = b; q[2] = c; q[3] = d;
}
#endif

// Work units are (column group, chunk of rows), column group major.  The unit (cg, chunk) asks for slice
// `chunk` of every region of column group cg + 1: by the time the grid gets there (chunks units later) the
// regions are L2 resident and the random bucket accesses hit.  Called by the first `cols` threads of a block.
__device__ __forceinline__ void prefetch_next_regions(const void *slots, uint64_t region_bytes, uint32_t n, uint32_t cols,
                                                      uint32_t cg, uint32_t chunk, uint32_t chunks, uint32_t t) {
    const uint32_t l = (cg + 1) * cols + t;
    if (t >= cols || l >= n) return;
    const uint64_t slice = ((region_bytes + chunks - 1) / chunks + 15) & ~15ULL;
    const uint64_t off = slice * chunk;
    if (off >= region_bytes) return;
    const uint64_t len = region_bytes - off < slice ? (region_bytes - off) & ~15ULL : slice;
    if (len) l2_prefetch_bulk(static_cast<const uint8_t *>(slots) + (uint64_t)l * region_bytes + off, (uint32_t)len);
}

} // namespace nsmh

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

} // namespace nsmh

// 256-bit global loads / stores (one 32-byte sector per access) shared by the lookup and multi-GPU
// kernels; plain loads / stores when the kernels are compiled for the host emulation.
#pragma once
#include <stdint.h>

namespace nsmh {

#ifndef NSMH_HOST_EMUL
__device__ __forceinline__ void ldg256(const void *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) {
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
__device__ __forceinline__ void stg256(void *p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
#else
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

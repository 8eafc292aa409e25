// Micro-benchmark (development tool): throughput of random 16-byte-slot operations on a table
// the size of the C2 hash tables, to see what bounds table_insert_kernel / probe_items_kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atom_bench atom_bench.cu && ./atom_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

struct __align__(16) Slot { uint64_t key; uint32_t val; uint32_t cnt; };

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
// window = number of slots an item's address is confined to, per group of `group` consecutive items
__device__ __forceinline__ uint64_t addr_of(uint64_t t, uint64_t nslots, uint64_t window, uint64_t group) {
    uint64_t r = mix(t * 0x9E3779B97F4A7C15ULL + 1);
    if (window >= nslots) return r % nslots;
    uint64_t nwin = nslots / window;
    uint64_t w = (t / group) % nwin;
    return w * window + r % window;
}

template <int OP>
__global__ void k(Slot *s, uint64_t items, uint64_t nslots, uint64_t window, uint64_t group, uint32_t *sink) {
    uint32_t acc = 0;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < items; t += (uint64_t)gridDim.x * blockDim.x) {
        Slot *p = s + addr_of(t, nslots, window, group);
        if (OP == 0) { uint4 v = __ldg(reinterpret_cast<const uint4 *>(p)); acc += v.x ^ v.w; }
        if (OP == 1) { acc += (uint32_t)atomicCAS(reinterpret_cast<unsigned long long *>(&p->key), ~0ULL, (unsigned long long)t); }
        if (OP == 2) { acc += atomicAdd(&p->cnt, 1u); }
        if (OP == 3) {
            uint64_t lo, hi;
            asm volatile("{\n\t.reg .b128 d, b, c;\n\tmov.b128 b, {%2, %3};\n\tmov.b128 c, {%4, %5};\n\t"
                         "atom.relaxed.gpu.global.cas.b128 d, [%6], b, c;\n\tmov.b128 {%0, %1}, d;\n\t}"
                         : "=l"(lo), "=l"(hi) : "l"(~0ULL), "l"(~0ULL), "l"((uint64_t)t), "l"((uint64_t)t), "l"(p) : "memory");
            acc += (uint32_t)lo;
        }
        if (OP == 4) {   // load key then CAS128 (the insert's sequence)
            uint64_t kk;
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(kk) : "l"(&p->key) : "memory");
            if (kk == ~0ULL) {
                uint64_t lo, hi;
                asm volatile("{\n\t.reg .b128 d, b, c;\n\tmov.b128 b, {%2, %3};\n\tmov.b128 c, {%4, %5};\n\t"
                             "atom.relaxed.gpu.global.cas.b128 d, [%6], b, c;\n\tmov.b128 {%0, %1}, d;\n\t}"
                             : "=l"(lo), "=l"(hi) : "l"(~0ULL), "l"(~0ULL), "l"((uint64_t)t), "l"((uint64_t)t), "l"(p) : "memory");
                acc += (uint32_t)lo;
            }
        }
        if (OP == 5) { p->key = t; p->val = (uint32_t)t; }   // plain 16-byte store (two stores)
        if (OP == 6) { *reinterpret_cast<uint4 *>(p) = make_uint4((uint32_t)t, 0, (uint32_t)t, 0); }
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int OP>
float run(Slot *s, uint64_t items, uint64_t nslots, uint64_t window, uint64_t group, uint32_t *sink, bool fresh, int blocks) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
        if (fresh) cudaMemset(s, 0xFF, nslots * sizeof(Slot));
        cudaEventRecord(a);
        k<OP><<<blocks, 256>>>(s, items, nslots, window, group, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const uint64_t nslots = 60ULL * 200002, items = 6000000;
    Slot *s; uint32_t *sink;
    cudaMalloc(&s, nslots * sizeof(Slot)); cudaMalloc(&sink, 4);
    const char *names[] = {"ldg128", "cas64", "add32", "cas128", "ld+cas128", "st8+st4", "st128"};
    for (int blocks : {148 * 8, 148 * 16}) {
        for (uint64_t window : {nslots, (uint64_t)200002 * 4, (uint64_t)200002}) {
            const uint64_t group = window >= nslots ? 1 : (window == 200002 ? 100000 : 400000);
            printf("blocks %d window %llu slots (%.1f MB) group %llu\n", blocks, (unsigned long long)window, window * 16 / 1e6, (unsigned long long)group);
            for (int fresh = 0; fresh < 2; ++fresh) {
                float t[7];
                t[0] = run<0>(s, items, nslots, window, group, sink, fresh, blocks);
                t[1] = run<1>(s, items, nslots, window, group, sink, fresh, blocks);
                t[2] = run<2>(s, items, nslots, window, group, sink, fresh, blocks);
                t[3] = run<3>(s, items, nslots, window, group, sink, fresh, blocks);
                t[4] = run<4>(s, items, nslots, window, group, sink, fresh, blocks);
                t[5] = run<5>(s, items, nslots, window, group, sink, fresh, blocks);
                t[6] = run<6>(s, items, nslots, window, group, sink, fresh, blocks);
                printf("  fresh=%d:", fresh);
                for (int i = 0; i < 7; ++i) printf(" %s %.1f us |", names[i], t[i] * 1e3);
                printf("\n");
            }
        }
    }
    return 0;
}

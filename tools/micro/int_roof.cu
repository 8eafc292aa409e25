// Micro-benchmark (development tool): the integer roof SURVEY 8(d) leaves "to be measured" - what one
// B200 sustains on the operations the reference's sketch loop is made of, so that the brute-force
// kernel (every k-mer x every hash: sketch_brute_kernel) and bench.py's `int32_reference_count` can be
// put against a measured number instead of the nominal 148 SMs x 64 ALU lanes x clock.
//
//   make -C tools/micro && tools/micro/int_roof            (prints one JSON line)
//
//   lop3     dependent chains of 32-bit LOP3 (XOR), 8 independent chains per thread: ALU pipe peak
//   imad     the same with IMAD (FMA-pipe integer): the second integer pipe
//   mix      LOP3 and IMAD chains interleaved: what a kernel that balances both pipes can issue
//   xormin64 the reference's pair: y = x ^ r (64 bit), m = min(m, y) (64-bit unsigned), 8 hashes per
//            k-mer in registers - the inner loop of sketch_brute_kernel without its loads
// Rates are per second over the whole device; "lane_ops" counts 32-bit operations (SURVEY 8(d) counts
// a 64-bit XOR as 2 and a 64-bit MIN as 4), "pairs" counts (k-mer, hash) pairs.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kChains = 8;
constexpr int kInner = 256;

template <int MODE>
__global__ void __launch_bounds__(256) int_kernel(uint32_t *sink, uint32_t seed, int outer) {
    uint32_t a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = seed + threadIdx.x * 2654435761u + i;
    const uint32_t b = seed | 1u;
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int it = 0; it < kInner; ++it) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) {
                if (MODE == 0) a[i] = (a[i] ^ b) ^ (a[i] >> 1);              // SHF + LOP3: two ALU ops
                if (MODE == 1) a[i] = a[i] * b + (uint32_t)it;               // IMAD
                if (MODE == 2) a[i] = (i & 1) ? a[i] * b + (uint32_t)it : ((a[i] ^ b) ^ (a[i] >> 1));
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) r ^= a[i];
    if (r == 0x12345678u) sink[0] = r;        // keeps the chains alive
}

// 8 running minima per thread, one new "k-mer" per inner step (a cheap LCG stands in for the roll)
__global__ void __launch_bounds__(256) xormin64_kernel(uint64_t *sink, uint64_t seed, int outer) {
    uint64_t m[kChains], r[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        m[i] = ~0ULL;
        r[i] = seed * (2 * i + 1) + threadIdx.x;
    }
    uint64_t x = seed ^ (blockIdx.x * 1315423911ull + threadIdx.x);
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int it = 0; it < kInner; ++it) {
            x = (x << 2) | ((x >> 61) & 3);       // shift in two bits, like the rolling k-mer
#pragma unroll
            for (int i = 0; i < kChains; ++i) {
                const uint64_t y = x ^ r[i];
                m[i] = y < m[i] ? y : m[i];
            }
        }
    }
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) acc ^= m[i];
    if (acc == 0x123456789abcdefULL) sink[0] = acc;
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) {
        fprintf(stderr, "int_roof: no CUDA device\n");
        return 1;
    }
    void *sink = nullptr;
    cudaMalloc(&sink, 64);
    const int blocks = p.multiProcessorCount * 8, outer = 64;
    const double threads = (double)blocks * 256;
    const double steps = (double)outer * kInner * kChains;           // chain steps per thread
    const float t0 = time_ms([&] { int_kernel<0><<<blocks, 256>>>((uint32_t *)sink, 12345u, outer); });
    const float t1 = time_ms([&] { int_kernel<1><<<blocks, 256>>>((uint32_t *)sink, 12345u, outer); });
    const float t2 = time_ms([&] { int_kernel<2><<<blocks, 256>>>((uint32_t *)sink, 12345u, outer); });
    const float t3 = time_ms([&] { xormin64_kernel<<<blocks, 256>>>((uint64_t *)sink, 0x9E3779B97F4A7C15ULL, outer); });
    const double pairs = threads * steps;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %d, "
           "\"lop3_chain_tops\": %.3f, \"imad_tops\": %.3f, \"mixed_tops\": %.3f, "
           "\"xormin64_gpairs_per_s\": %.2f, \"xormin64_lane_tops_survey_count\": %.3f, "
           "\"note\": \"lop3_chain counts 2 ALU ops per step (shift + 3-input xor); xormin64 counts 6 lane-ops per pair (SURVEY 8(d))\"}\n",
           p.name, p.multiProcessorCount, p.clockRate / 1000,
           threads * steps * 2 / (t0 * 1e-3) / 1e12, threads * steps / (t1 * 1e-3) / 1e12,
           threads * steps * 1.5 / (t2 * 1e-3) / 1e12,
           pairs / (t3 * 1e-3) / 1e9, pairs * 6 / (t3 * 1e-3) / 1e12);
    cudaFree(sink);
    return 0;
}

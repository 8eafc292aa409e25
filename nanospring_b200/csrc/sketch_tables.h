// Host-side construction of the lookup tables of sketch_filter_kernel (shared by sketch.cu and the
// host emulation of the kernels, tests/cpp/sketch_host_emul.cpp).
// Per prefix width b: hit/first tables at [2^b, 2^(b+1)), a per-b chain of hashes that share a target
// prefix (next), and the 3-position table at [2^(b+4), 2^(b+5)) (hit3).
#pragma once
#include <stdint.h>

#include <vector>

#include "nsmh_constants.h"

namespace nsmh {

struct FilterTables {
    std::vector<uint16_t> first;            // head of the chain | its successor << 8
    std::vector<uint8_t> next, hit3;
};

inline FilterTables make_filter_tables(const uint64_t *rand, uint32_t n, uint32_t k) {
    std::vector<uint8_t> hit(kFilterTabSize, 0);
    FilterTables ft;
    std::vector<uint8_t> first(kFilterTabSize, 0xFF);
    std::vector<uint8_t> &next = ft.next, &hit3 = ft.hit3;
    next.assign((size_t)(kFilterMaxBits + 1) * (n ? n : 1), 0xFF);
    hit3.assign(kFilter3TabSize, 0);
    ft.first.assign(kFilterTabSize, 0xFFFF);
    if (n <= 255) {
        const uint64_t mask = (1ULL << (2 * k)) - 1;
        for (int b = 0; b <= kFilterMaxBits && b <= 2 * (int)k; ++b) {
            for (uint32_t l = 0; l < n; ++l) {
                uint64_t rlo = rand[l] & mask;
                uint32_t t = b ? (uint32_t)(rlo >> (2 * k - b)) : 0u;
                uint32_t idx = (1u << b) | t;
                hit[idx] = 1;
                next[(size_t)b * n + l] = first[idx];
                first[idx] = (uint8_t)l;
            }
        }
        for (int b = 0; b <= kFilterMaxBits && b <= 2 * (int)k; ++b)
            for (uint32_t idx = 1u << b; idx < (2u << b); ++idx)
                if (first[idx] != 0xFF) ft.first[idx] = (uint16_t)(first[idx] | (next[(size_t)b * n + first[idx]] << 8));
        for (int b = 0; b <= kFilter3MaxBits && b <= 2 * (int)k; ++b) {
            const uint32_t pm = (1u << b) - 1, base = 1u << b;
            for (uint32_t wv = 0; wv < (1u << (b + 4)); ++wv) {
                uint32_t p0 = (wv >> 4) & pm, p1 = (wv >> 2) & pm, p2 = wv & pm;
                hit3[(1u << (b + 4)) | wv] =
                    (uint8_t)((hit[base | p0] << 2) | (hit[base | p1] << 1) | hit[base | p2]);
            }
        }
    }
    return ft;
}

} // namespace nsmh

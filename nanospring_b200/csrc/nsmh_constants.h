// Layout constants shared by the host code (nsmh_internal.cuh) and the kernel headers that are also
// compiled for the host by the emulation tests (no CUDA includes here).
#pragma once
#include <stdint.h>

namespace nsmh {

constexpr int kWordBases = 16;          // bases per packed u32 word, first base in bits 31..30
constexpr int kPackPadWords = 8;        // zero words after the last packed word (k-mer window overrun)
constexpr int kTileWords = 640;         // default words (10240 k-mer start positions) per sketch tile: 32 warps per SM fit
constexpr int kFilterMaxBits = 12;      // largest prefix width b of the sketch filter tables
constexpr int kFilterTabSize = 1 << (kFilterMaxBits + 1);  // table for b lives at [2^b, 2^(b+1))
constexpr int kFilterLambdaLog2 = 2;    // default: b = floor(log2(#kmers)) - 2  => 4..8 k-mers expected per bucket
constexpr int kFilter3MaxBits = 11;     // 3-positions-per-lookup tables: window of b+4 bits
constexpr int kFilter3TabSize = 1 << (kFilter3MaxBits + 5);   // table for b lives at [2^(b+4), 2^(b+5))

} // namespace nsmh

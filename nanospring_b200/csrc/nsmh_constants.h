// Layout constants shared by the host code (nsmh_internal.cuh) and the kernel headers that are also
// compiled for the host by the emulation tests (no CUDA includes here).
#pragma once
#include <stdint.h>

namespace nsmh {

constexpr int kWordBases = 16;          // bases per packed u32 word, first base in bits 31..30
constexpr int kPackPadWords = 8;        // zero words after the last packed word (k-mer window overrun)

} // namespace nsmh

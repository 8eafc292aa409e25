// TEST INFRASTRUCTURE ONLY (oracle/): C entry points around the UNMODIFIED reference consensus
// translation unit (/root/reference/src/Consensus.cpp), compiled where it lies by
// oracle/Makefile into oracle/_ref/libnsref_consensus.so.  Pins the candidate pre-filter
// restatement (oracle/minhash_oracle.c: orc_read_flags) and the device kernels
// (nanospring_b200/csrc/prefilter.cu) to the reference's own Consensus::checkRepetitive
// (Consensus.cpp:405-424) and Consensus::initialize (Consensus.cpp:426-442).
#define private public
#include "Consensus.h"
#undef private

#include <iostream>
#include <sstream>

void nsref_fill_read_data(ReadData &rD, const char *bases, const uint64_t *offsets, uint32_t numReads);

extern "C" {

// out[i] = isRepetitive[i] after the reference's Consensus::initialize() over these reads.
int nsref_consensus_is_repetitive(const char *bases, const uint64_t *offsets, uint32_t numReads, int threads,
                                  uint8_t *out) {
    try {
        ReadData rD;
        nsref_fill_read_data(rD, bases, offsets, numReads);
        Consensus c;
        c.rD = &rD;
        if (threads > 0) omp_set_num_threads(threads);
        // Consensus::initialize also sizes readStatusLock to numLocks = 2^24 OpenMP locks
        // (Consensus.h:102); that is the reference's own behaviour, ~130 MB, harmless here.
        c.initialize();
        for (uint32_t i = 0; i < numReads; ++i) out[i] = c.isRepetitive[i];
        return 0;
    } catch (const std::exception &e) {
        std::cerr << "nsref_consensus_is_repetitive: " << e.what() << std::endl;
        return 1;
    }
}

// One read through checkRepetitive itself.
int nsref_check_repetitive(const char *s, size_t len) {
    ReadData rD;
    uint64_t offsets[2] = {0, (uint64_t)len};
    nsref_fill_read_data(rD, s, offsets, 1);
    Consensus c;
    c.rD = &rD;
    return c.checkRepetitive(0) ? 1 : 0;
}

}  // extern "C"

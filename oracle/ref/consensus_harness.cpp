// TEST INFRASTRUCTURE ONLY (oracle/): C entry points around the UNMODIFIED reference consensus
// translation unit (/root/reference/src/Consensus.cpp), compiled where it lies by
// oracle/Makefile into oracle/_ref/libnsref_consensus.so.  Pins the candidate pre-filter
// restatement (oracle/minhash_oracle.c: orc_read_flags) and the device kernels
// (nanospring_b200/csrc/prefilter.cu) to the reference's own Consensus::checkRepetitive
// (Consensus.cpp:405-424) and Consensus::initialize (Consensus.cpp:426-442); and exposes the unmodified
// ConsensusGraph::alignRead (nsref_align_read below) as the oracle of the next row, N4.
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

// ---- N4: ConsensusGraph::alignRead (ConsensusGraph.cpp:161-398), unmodified --------------------
// The graph is seeded with `ref` as the caller does (Consensus.cpp:322-323: initialize + calculateMainPathGreedy,
// after which mainPath.path == ref), then `read` is aligned against that main path with the CLI's minimap2
// parameters.  Output: *ok = alignRead's return value; rel_pos / begin_off / end_off as it sets them; the edit
// script as (type, info) pairs - type 0 SAME (info = run length), 1 INSERT, 2 DELETE, 3 SUBSTITUTION (info = the
// base).  Returns the number of edits (also when it exceeds cap; only cap pairs are written), -1 on an exception.
// This is the oracle a device implementation of N4 would be pinned to: the parity contract is this output, byte
// for byte (tests/test_align_oracle.py holds golden vectors generated from it).
long nsref_align_read(const char *ref, size_t ref_len, const char *read, size_t read_len, size_t m_k, size_t m_w,
                      size_t max_chain_iter, int *ok, long *rel_pos, long *begin_off, long *end_off, uint8_t *types,
                      uint64_t *infos, size_t cap) {
    try {
        ConsensusGraph g;
        g.initialize(std::string(ref, ref_len), 0, 0);
        g.calculateMainPathGreedy();
        std::vector<Edit> script;
        ssize_t rp = 0, bo = 0, eo = 0;
        const bool good = g.alignRead(std::string(read, read_len), script, rp, bo, eo, m_k, m_w, max_chain_iter);
        *ok = good ? 1 : 0;
        *rel_pos = (long)rp;
        *begin_off = (long)bo;
        *end_off = (long)eo;
        for (size_t i = 0; i < script.size() && i < cap; ++i) {
            types[i] = (uint8_t)script[i].editType;
            infos[i] = script[i].editType == SAME ? (uint64_t)script[i].editInfo.num
                                                  : (uint64_t)(unsigned char)script[i].editInfo.ins;
        }
        return (long)script.size();
    } catch (const std::exception &e) {
        std::cerr << "nsref_align_read: " << e.what() << std::endl;
        return -1;
    }
}

}  // extern "C"

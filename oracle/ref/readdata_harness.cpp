// TEST INFRASTRUCTURE ONLY (oracle/).  C entry points around the reference's OWN read loader:
// the unmodified src/ReadData.cpp + src/dnaToBits.cpp compiled where they lie (oracle/Makefile,
// target readdata -> oracle/_ref/libnsref_readdata.so).  Used to pin the FASTQ-ingest oracle
// (oracle/minhash_oracle.c, orc_fastq_*) and to generate tests/golden/fastq_golden.npz:
//   ReadData::loadFromFile(fileName, FASTQ|GZIP, low_mem)   ReadData.cpp:12-26, 86-221
//   ReadData::getRead(i, str)                               ReadData.cpp:225-235
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>

#include "ReadData.h"

extern "C" {

// filetype: 0 FASTQ, 2 GZIP (ReadData.h:23).  Returns nullptr on failure.
void *nsrd_load(const char *path, int filetype, int low_mem, const char *temp_dir) {
    ReadData *rd = new ReadData();
    rd->tempDir = temp_dir ? temp_dir : "/tmp";
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());          // the loader prints numReads / avgReadLen / maxReadLen
    bool ok = true;
    try {
        rd->loadFromFile(path, filetype == 2 ? ReadData::GZIP : ReadData::FASTQ, low_mem != 0);
    } catch (...) {
        ok = false;
    }
    std::cout.rdbuf(old);
    if (!ok) { delete rd; return nullptr; }
    return rd;
}

uint32_t nsrd_num_reads(void *h) { return static_cast<ReadData *>(h)->getNumReads(); }
uint64_t nsrd_avg_read_len(void *h) { return static_cast<ReadData *>(h)->avgReadLen; }
uint64_t nsrd_max_read_len(void *h) { return static_cast<ReadData *>(h)->maxReadLen; }

// lengths[i] = size of getRead(i)
void nsrd_lengths(void *h, uint64_t *lengths) {
    ReadData *rd = static_cast<ReadData *>(h);
    std::string s;
    for (read_t i = 0; i < rd->getNumReads(); ++i) {
        rd->getRead(i, s);
        lengths[i] = s.size();
    }
}

// concatenation of getRead(0..N-1) into out (sum of lengths bytes)
void nsrd_all_reads(void *h, char *out) {
    ReadData *rd = static_cast<ReadData *>(h);
    std::string s;
    size_t used = 0;
    for (read_t i = 0; i < rd->getNumReads(); ++i) {
        rd->getRead(i, s);
        std::memcpy(out + used, s.data(), s.size());
        used += s.size();
    }
}

void nsrd_free(void *h) { delete static_cast<ReadData *>(h); }

}  // extern "C"

// TEST INFRASTRUCTURE ONLY (oracle/): Boost-free stand-in for the reference's
// src/ReadData.cpp so that the UNMODIFIED reference translation units
// src/ReadFilter.cpp, src/BBHashMap.cpp and src/dnaToBits.cpp can be compiled
// where they lie under /root/reference (see oracle/Makefile -> oracle/_ref/).
//
// It defines the ReadData members declared in /root/reference/include/ReadData.h
// that the MinHash path touches (ReadFilter.cpp:13-14,26,41): getNumReads,
// getRead, getReadPos, plus the static complement helper and the destructor.
// Reads are held in memory as DnaBitset objects, i.e. the reference's own
// "high memory" representation (ReadData.cpp:86-154), so getRead() goes through
// the reference's own 2-bit pack/unpack (dnaToBits.cpp:11-36, 81-98).
#define private public
#include "ReadData.h"
#undef private

#include <stdexcept>

void ReadData::loadFromFile(const char *, enum Filetype, bool) {
    throw std::runtime_error("read_data_shim: file loading is done by the harness");
}

read_t ReadData::getNumReads() { return numReads; }

void ReadData::getRead(read_t readId, std::string &readStr) {
    readData[readId]->to_string(readStr);
}

std::vector<unsigned long> &ReadData::getReadPos() { return readPos; }

std::vector<std::unique_ptr<DnaBitset>> &ReadData::getReadData() { return readData; }

// A<->T, C<->G, anything else unchanged (ReadData.cpp:247-260).
char ReadData::toComplement(char base) {
    if (base == 'A') return 'T';
    if (base == 'T') return 'A';
    if (base == 'C') return 'G';
    if (base == 'G') return 'C';
    return base;
}

ReadData::~ReadData() {}

// Harness entry: fill a ReadData from concatenated ASCII bases + offsets.
void nsref_fill_read_data(ReadData &rD, const char *bases, const uint64_t *offsets,
                          uint32_t numReads) {
    rD.reads_in_memory = true;
    rD.numReads = numReads;
    rD.readData.clear();
    rD.readPos.clear();
    rD.readData.resize(numReads);
    rD.readPos.assign(numReads, 0);
    size_t total = 0, maxLen = 0;
    for (uint32_t i = 0; i < numReads; ++i) {
        size_t len = offsets[i + 1] - offsets[i];
        rD.readData[i].reset(new DnaBitset(bases + offsets[i], len));
        total += len;
        if (len > maxLen) maxLen = len;
    }
    rD.maxReadLen = maxLen;
    rD.avgReadLen = numReads ? total / numReads : 0;
}

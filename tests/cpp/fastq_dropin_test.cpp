// Drop-in test of the FASTQ ingest against the reference's OWN loader.
//   reference:  ReadData rD; rD.tempDir = ...; rD.loadFromFile(file, filetype, low_mem = true);   (Compressor.cpp:57-60)
//               rF.initialize(rD);                                                               (Compressor.cpp:76)
//   device:     GpuMinHashReadFilter::initializeFromFile(file, filetype)
// The unmodified src/ReadData.cpp (boost::iostreams replaced by oracle/ref/boost_shim on zlib) loads
// the file on the host; the same file goes through nsmh_load_fastq_file.  Checked: read count, every
// read as getRead() returns it, the sketch matrix of GpuMinHashReadFilter::initialize(rD) (reads pulled
// through the reference's getRead) vs initializeFromFile, and a few online queries.
// Built by `make -C oracle fastq_dropin`; run on the GPU box by tests/test_gpu_fastq_dropin.py.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>

#include "GpuMinHashReadFilter.h"

int main(int argc, char **argv) {
    if (argc < 7) {
        std::fprintf(stderr, "usage: %s file gzip(0|1) k n thr tmpdir\n", argv[0]);
        return 2;
    }
    const ReadData::Filetype ft = std::atoi(argv[2]) ? ReadData::GZIP : ReadData::FASTQ;
    const size_t k = std::atoi(argv[3]), n = std::atoi(argv[4]), thr = std::atoi(argv[5]);

    ReadData rD;
    rD.tempDir = argv[6];
    {
        std::streambuf *old = std::cout.rdbuf();
        std::ostringstream sink;
        std::cout.rdbuf(sink.rdbuf());
        rD.loadFromFile(argv[1], ft, true);        // the CLI's mode (main.cpp:40)
        std::cout.rdbuf(old);
    }

    GpuMinHashReadFilter viaGetRead, viaFile;
    for (GpuMinHashReadFilter *g : {&viaGetRead, &viaFile}) {
        g->k = k;
        g->n = n;
        g->overlapSketchThreshold = thr;
        g->randNumbers.resize(n);
        nsmh_rand_from_seed(20261017u, (uint32_t)n, g->randNumbers.data());
    }
    viaGetRead.initialize(rD);
    viaFile.initializeFromFile(argv[1], ft);

    long bad = 0;
    const read_t N = rD.getNumReads();
    if (viaFile.getNumReads() != N) {
        std::printf("read count: reference %u device %u\n", N, viaFile.getNumReads());
        return 1;
    }
    std::string a, b;
    size_t bases = 0, maxLen = 0;
    for (read_t i = 0; i < N; ++i) {
        rD.getRead(i, a);
        viaFile.getRead(i, b);
        if (a != b && ++bad < 5) std::printf("read %u differs (%zu vs %zu bases)\n", i, a.size(), b.size());
        bases += a.size();
        if (a.size() > maxLen) maxLen = a.size();
    }
    if (maxLen != rD.maxReadLen || (N && bases / N != rD.avgReadLen)) ++bad;
    std::vector<uint64_t> s1((size_t)N * n), s2((size_t)N * n);
    if (nsmh_get_sketches(viaGetRead.handle(), s1.data()) || nsmh_get_sketches(viaFile.handle(), s2.data())) return 3;
    if (s1 != s2) {
        ++bad;
        std::printf("sketch matrices differ\n");
    }
    std::vector<read_t> r1, r2;
    long ids = 0;
    for (read_t i = 0; i < N; i += (N / 64 ? N / 64 : 1)) {
        rD.getRead(i, a);
        ReadFilter *f1 = &viaGetRead, *f2 = &viaFile;
        f1->getFilteredReads(a, r1);
        f2->getFilteredReads(a, r2);
        if (r1 != r2) ++bad;
        ids += (long)r1.size();
    }
    std::printf("reads %u bases %zu candidate ids %ld mismatches %ld\n", N, bases, ids, bad);
    if (bad) return 1;
    std::printf("FASTQ DROPIN OK\n");
    return 0;
}

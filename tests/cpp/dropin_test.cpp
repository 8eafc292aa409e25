// Drop-in test of the boundary: the reference's caller pattern
//     ReadFilter *rF = ...;  rF->initialize(rD);                    (Compressor.cpp:69-76)
//     #pragma omp parallel ... rF->getFilteredReads(window, results) (Consensus.cpp:29,189)
// driven through GpuMinHashReadFilter (libnsmh.so, GPU) and, side by side, through the
// reference's own MinHashReadFilter (oracle/_ref/libnsref.so, CPU) with the same injected
// random numbers.  Compiled against the reference's headers by `make -C oracle dropin`;
// run on the GPU box by tests/test_gpu_dropin.py.  Prints "DROPIN OK" on success.
#include <omp.h>

#include <algorithm>
#include <chrono>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>

#include "GpuMinHashReadFilter.h"

void nsref_fill_read_data(ReadData &rD, const char *bases, const uint64_t *offsets, uint32_t numReads);
extern "C" {
void *nsref_create(const char *bases, const uint64_t *offsets, uint32_t numReads, uint32_t k, uint32_t n,
                   uint32_t thr, const uint64_t *randNumbers, int threads, const char *tmpdir,
                   double *sketch_ms, double *build_ms);
size_t nsref_query_string(void *h, const char *s, size_t len, uint32_t *out, size_t cap);
void nsref_destroy(void *h);
}

int main(int argc, char **argv) {
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s reads.bin k n thr tmpdir\n", argv[0]);
        return 2;
    }
    // reads.bin: u32 numReads, u64 offsets[numReads+1], bases
    std::ifstream in(argv[1], std::ios::binary);
    uint32_t numReads = 0;
    in.read(reinterpret_cast<char *>(&numReads), 4);
    std::vector<uint64_t> offsets((size_t)numReads + 1);
    in.read(reinterpret_cast<char *>(offsets.data()), offsets.size() * 8);
    std::string bases(offsets[numReads], '\0');
    in.read(&bases[0], bases.size());
    const size_t k = std::atoi(argv[2]), n = std::atoi(argv[3]), thr = std::atoi(argv[4]);

    ReadData rD;
    nsref_fill_read_data(rD, bases.data(), offsets.data(), numReads);

    GpuMinHashReadFilter gpu;
    gpu.k = k;
    gpu.n = n;
    gpu.overlapSketchThreshold = thr;
    gpu.tempDir = argv[5];
    gpu.randNumbers.resize(n);
    nsmh_rand_from_seed(20261017u, (uint32_t)n, gpu.randNumbers.data());
    // DROPIN_DEVICES="0,0": several devices behind the one filter (the same device named twice on a 1-GPU box)
    if (const char *dv = std::getenv("DROPIN_DEVICES"))
        for (const char *p = dv; *p;) {
            gpu.devices.push_back(std::atoi(p));
            while (*p && *p != ',') ++p;
            if (*p == ',') ++p;
        }
    const bool multi = gpu.devices.size() > 1;
    ReadFilter *rF = &gpu;                 // the caller only ever holds a ReadFilter* (Consensus.h:41)
    const char *mode = std::getenv("DROPIN_MODE");
    if (mode && std::string(mode) == "bitset") {
        // the reads as ReadData holds them: DnaBitset bytes (dnaToBits.cpp:46-71), every read byte aligned
        std::vector<uint8_t> packed;
        std::vector<size_t> lens(numReads);
        for (uint32_t i = 0; i < numReads; ++i) {
            const size_t len = offsets[i + 1] - offsets[i];
            lens[i] = len;
            for (size_t b = 0; b < len; b += 4) {
                uint8_t byte = 0;
                for (size_t j = 0; j < 4 && b + j < len; ++j) {
                    const uint8_t c = (uint8_t)bases[offsets[i] + b + j];
                    byte |= (uint8_t)(((c & 2) | ((c & 4) >> 2)) << (6 - 2 * j));
                }
                packed.push_back(byte);
            }
        }
        const std::string path = std::string(argv[5]) + "/readBitset";      // ReadData.h:127
        {
            std::ofstream bf(path, std::ios::binary);
            bf.write(reinterpret_cast<const char *>(packed.data()), (std::streamsize)packed.size());
        }
        gpu.initializeFromBitsetFile(path, lens);
    } else {
        rF->initialize(rD);
    }

    void *ref = nsref_create(bases.data(), offsets.data(), numReads, (uint32_t)k, (uint32_t)n, (uint32_t)thr,
                             gpu.randNumbers.data(), 0, argv[5], nullptr, nullptr);
    if (!ref) return 3;

    long mismatches = 0, queries = 0, total = 0;
    std::vector<std::vector<double>> lat_gpu(omp_get_max_threads()), lat_ref(omp_get_max_threads());
    const auto wall0 = std::chrono::steady_clock::now();
#pragma omp parallel reduction(+ : mismatches, queries, total)
    {
        std::string s, rc;
        auto &lg = lat_gpu[omp_get_thread_num()];
        auto &lr = lat_ref[omp_get_thread_num()];
        auto us = [](std::chrono::steady_clock::time_point a) {
            return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count();
        };
        std::vector<read_t> res, resRc, want((size_t)numReads + 1);
#pragma omp for schedule(dynamic, 4)
        for (read_t i = 0; i < numReads; ++i) {
            rD.getRead(i, s);
            if (s.size() > 3000) s = s.substr(s.size() / 3, 2500);      // a window, like addRelatedReads
            rc.clear();
            ReadData::toReverseComplement(s.begin(), s.end(), std::inserter(rc, rc.end()));
            auto t0 = std::chrono::steady_clock::now();
            rF->getFilteredReads(s, res);
            lg.push_back(us(t0));
            t0 = std::chrono::steady_clock::now();
            rF->getFilteredReads(rc, resRc);
            lg.push_back(us(t0));
            t0 = std::chrono::steady_clock::now();
            size_t c = nsref_query_string(ref, s.data(), s.size(), want.data(), want.size());
            lr.push_back(us(t0));
            if (c != res.size() || !std::equal(res.begin(), res.end(), want.begin())) ++mismatches;
            total += (long)c;
            c = nsref_query_string(ref, rc.data(), rc.size(), want.data(), want.size());
            if (c != resRc.size() || !std::equal(resRc.begin(), resRc.end(), want.begin())) ++mismatches;
            total += (long)c;
            queries += 2;
            if (i % 16 == 0 && !multi) {   // the paired fast path gives the same two answers
                std::vector<read_t> a, b;
                gpu.getFilteredReadsPair(s, rc, a, b);
                if (a != res || b != resRc) ++mismatches;
            }
        }
    }
    nsref_destroy(ref);
    const double wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
    auto pct = [](std::vector<std::vector<double>> &v, double p) {
        std::vector<double> all;
        for (auto &x : v) all.insert(all.end(), x.begin(), x.end());
        if (all.empty()) return 0.0;
        std::sort(all.begin(), all.end());
        return all[std::min(all.size() - 1, (size_t)(p * all.size()))];
    };
    std::printf("latency per getFilteredReads call, %d OpenMP threads (windows <= 2500 bases): GPU filter p50 %.1f us p99 %.1f us, "
                "reference filter p50 %.1f us p99 %.1f us; whole loop %.3f s\n", omp_get_max_threads(), pct(lat_gpu, 0.5),
                pct(lat_gpu, 0.99), pct(lat_ref, 0.5), pct(lat_ref, 0.99), wall_s);
    std::printf("threads %d queries %ld candidate ids %ld mismatches %ld\n", omp_get_max_threads(), queries,
                total, mismatches);
    if (mismatches) return 1;
    std::printf("DROPIN OK\n");
    return 0;
}

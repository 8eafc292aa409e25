// TEST INFRASTRUCTURE ONLY (oracle/): C entry points around the UNMODIFIED
// reference MinHashReadFilter (/root/reference/src/ReadFilter.cpp,
// src/BBHashMap.cpp, include/BooPHF.h, src/dnaToBits.cpp), compiled where the
// sources lie by oracle/Makefile into oracle/_ref/libnsref.so.
//
// The reference draws its n random numbers from std::random_device inside
// initialize() (ReadFilter.cpp:49-63), so two runs never agree. For
// reproducible goldens this harness REPLAYS the body of
// MinHashReadFilter::initialize (ReadFilter.cpp:11-47) with caller-supplied
// rand[] and calls the reference's own string2Sketch / populateHashTables /
// getFilteredReads. nsref_create_verbatim() instead runs initialize() itself
// and reads the numbers back, to prove the replay loses nothing.
#define private public
#include "ReadFilter.h"
#undef private

#include <chrono>
#include <cstring>
#include <functional>
#include <iostream>
#include <iterator>
#include <omp.h>
#include <random>
#include <sstream>
#include <stdexcept>

void nsref_fill_read_data(ReadData &rD, const char *bases, const uint64_t *offsets,
                          uint32_t numReads);

namespace {

struct Ctx {
    ReadData rD;
    MinHashReadFilter rF;
    std::vector<kMer_t> sketches;
    std::string err;
};

// The reference prints progress to std::cout (ReadFilter.cpp:160,168-170);
// keep the harness' stdout clean for bench.py's single JSON line.
struct MuteCout {
    std::ostringstream sink;
    std::streambuf *old;
    MuteCout() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~MuteCout() { std::cout.rdbuf(old); }
};

double ms_since(std::chrono::high_resolution_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(
               std::chrono::high_resolution_clock::now() - t0).count();
}

void check_libstdcxx_assumptions() {
    // std::hash<uint64_t> is the identity and
    // uniform_int_distribution<unsigned long long>() over mt19937_64 is the raw stream.
    if (std::hash<kMer_t>()(0x123456789abcdefULL) != 0x123456789abcdefULL)
        throw std::runtime_error("std::hash<uint64_t> is not the identity here");
    std::mt19937_64 a(12345), b(12345);
    std::uniform_int_distribution<unsigned long long> dis;
    for (int i = 0; i < 8; ++i)
        if (dis(a) != b()) throw std::runtime_error("uniform_int_distribution != raw mt19937_64");
}

} // namespace

extern "C" {

// Replayed initialize(): sketches all reads, builds the n BBHashMap tables.
void *nsref_create(const char *bases, const uint64_t *offsets, uint32_t numReads,
                   uint32_t k, uint32_t n, uint32_t thr, const uint64_t *randNumbers,
                   int threads, const char *tmpdir, double *sketch_ms, double *build_ms) {
    try {
        check_libstdcxx_assumptions();
        MuteCout mute;
        Ctx *c = new Ctx();
        nsref_fill_read_data(c->rD, bases, offsets, numReads);
        MinHashReadFilter &rF = c->rF;
        rF.k = k;
        rF.n = n;
        rF.overlapSketchThreshold = thr;
        rF.tempDir = tmpdir;
        rF.rD = &c->rD;
        rF.numReads = numReads;
        rF.readPos = &c->rD.getReadPos();
        rF.randNumbers = new kMer_t[n];
        for (uint32_t i = 0; i < n; ++i) rF.randNumbers[i] = randNumbers[i];

        if (threads > 0) omp_set_num_threads(threads);
        c->sketches.assign((size_t)n * numReads, 0);
        size_t maxNumkMers = c->rD.maxReadLen < (size_t)k - 1 ? 0 : c->rD.maxReadLen - k + 1;
        auto t0 = std::chrono::high_resolution_clock::now();
#pragma omp parallel
        {
            std::vector<kMer_t> kMersVec(maxNumkMers), hashesVec(n);
            std::string readStr;
#pragma omp for schedule(dynamic, 16)
            for (read_t i = 0; i < numReads; ++i) {
                c->rD.getRead(i, readStr);
                rF.string2Sketch(readStr, c->sketches.data() + (size_t)i * n, kMersVec, hashesVec);
            }
        }
        if (sketch_ms) *sketch_ms = ms_since(t0);
        t0 = std::chrono::high_resolution_clock::now();
        rF.populateHashTables(c->sketches);
        if (build_ms) *build_ms = ms_since(t0);
        return c;
    } catch (const std::exception &e) {
        std::cerr << "nsref_create: " << e.what() << std::endl;
        return nullptr;
    }
}

// Verbatim initialize() (unseeded RNG); the numbers it drew are copied to rand_out[n].
void *nsref_create_verbatim(const char *bases, const uint64_t *offsets, uint32_t numReads,
                            uint32_t k, uint32_t n, uint32_t thr, int threads,
                            const char *tmpdir, uint64_t *rand_out) {
    try {
        MuteCout mute;
        Ctx *c = new Ctx();
        nsref_fill_read_data(c->rD, bases, offsets, numReads);
        c->rF.k = k;
        c->rF.n = n;
        c->rF.overlapSketchThreshold = thr;
        c->rF.tempDir = tmpdir;
        if (threads > 0) omp_set_num_threads(threads);
        c->rF.initialize(c->rD);
        for (uint32_t i = 0; i < n; ++i) rand_out[i] = c->rF.randNumbers[i];
        return c;
    } catch (const std::exception &e) {
        std::cerr << "nsref_create_verbatim: " << e.what() << std::endl;
        return nullptr;
    }
}

int nsref_has_sketches(void *h) { return !static_cast<Ctx *>(h)->sketches.empty(); }

void nsref_get_sketches(void *h, uint64_t *out) {
    Ctx *c = static_cast<Ctx *>(h);
    std::memcpy(out, c->sketches.data(), c->sketches.size() * sizeof(uint64_t));
}

// One query through the public virtual interface (ReadFilter.h:24-25).
size_t nsref_query_string(void *h, const char *s, size_t len, uint32_t *out, size_t cap) {
    Ctx *c = static_cast<Ctx *>(h);
    std::vector<read_t> res;
    c->rF.getFilteredReads(std::string(s, len), res);
    for (size_t i = 0; i < res.size() && i < cap; ++i) out[i] = res[i];
    return res.size();
}

// Query every read against the tables.
//   mode 0: forward, private getFilteredReads(sketch[]) on the stored sketches
//   mode 1: reverse complement string through the public overload (Consensus.cpp:181-191)
//   mode 2: forward string through the public overload (re-sketches)
// offsets_out has numReads+1 entries; *ids_out is malloc'ed (free with nsref_free).
int nsref_query_all(void *h, int mode, int threads, uint64_t *offsets_out, uint32_t **ids_out,
                    double *ms) {
    Ctx *c = static_cast<Ctx *>(h);
    const uint32_t N = c->rF.numReads;
    const size_t n = c->rF.n;
    if (mode == 0 && c->sketches.empty()) return -1;
    if (threads > 0) omp_set_num_threads(threads);
    std::vector<std::vector<read_t>> res(N);
    auto t0 = std::chrono::high_resolution_clock::now();
#pragma omp parallel
    {
        std::string s, rc;
#pragma omp for schedule(dynamic, 16)
        for (read_t i = 0; i < N; ++i) {
            if (mode == 0) {
                c->rF.getFilteredReads(c->sketches.data() + (size_t)i * n, res[i]);
            } else {
                c->rD.getRead(i, s);
                if (mode == 1) {
                    rc.clear();
                    ReadData::toReverseComplement(s.begin(), s.end(), std::inserter(rc, rc.end()));
                    c->rF.getFilteredReads(rc, res[i]);
                } else {
                    c->rF.getFilteredReads(s, res[i]);
                }
            }
        }
    }
    if (ms) *ms = ms_since(t0);
    uint64_t total = 0;
    for (uint32_t i = 0; i < N; ++i) {
        offsets_out[i] = total;
        total += res[i].size();
    }
    offsets_out[N] = total;
    uint32_t *ids = static_cast<uint32_t *>(malloc((total ? total : 1) * sizeof(uint32_t)));
    for (uint32_t i = 0; i < N; ++i)
        if (!res[i].empty())
            std::memcpy(ids + offsets_out[i], res[i].data(), res[i].size() * sizeof(uint32_t));
    *ids_out = ids;
    return 0;
}

void nsref_free(void *p) { free(p); }

void nsref_destroy(void *h) {
    MuteCout mute;
    delete static_cast<Ctx *>(h);
}

// The objects behind a handle, as the reference's callers hold them (Consensus.h:31-41:
// `ReadData *rD; ReadFilter *rF;`): for tests that run the reference's own consensus stage
// against this filter (tests/cpp/consensus_dropin_test.cpp).
void *nsref_read_filter(void *h) { return static_cast<ReadFilter *>(&static_cast<Ctx *>(h)->rF); }
void *nsref_read_data(void *h) { return &static_cast<Ctx *>(h)->rD; }

// Public static helpers, for the known-answer tests of SURVEY section 8(c).
uint64_t nsref_kmer_to_int(const char *s, size_t len) {
    return MinHashReadFilter::kMerToInt(std::string(s, len));
}

size_t nsref_string2kmers(const char *s, size_t len, uint32_t k, uint64_t *out) {
    std::string str(s, len);
    if (len < k) return 0;
    std::vector<kMer_t> v(len - k + 1);
    MinHashReadFilter::string2KMers(str, k, v);
    std::memcpy(out, v.data(), v.size() * sizeof(uint64_t));
    return v.size();
}

void nsref_rand_from_seed(uint32_t seed, uint32_t n, uint64_t *out) {
    // what generateRandomNumbers() would draw had random_device returned `seed`
    std::mt19937_64 gen(seed);
    std::uniform_int_distribution<unsigned long long> dis;
    for (uint32_t i = 0; i < n; ++i) out[i] = dis(gen);
}

} // extern "C"

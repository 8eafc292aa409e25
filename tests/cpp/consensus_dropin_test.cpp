// The drop-in under the REAL caller.
//
// The reference's consensus stage -- Consensus::generateAndWriteConsensus (src/Consensus.cpp:20),
// whose addRelatedReads (:168-191) issues the forward + reverse-complement getFilteredReads of
// every window of the growing main path, followed by minimap2 alignment of the candidates
// (src/ConsensusGraph.cpp:161-398) -- is compiled UNMODIFIED where it lies (oracle/Makefile,
// target consensus_dropin) and run twice over the same reads:
//     arm "ref": ReadFilter* = the reference's own MinHashReadFilter (CPU, libnsref.so)
//     arm "gpu": ReadFilter* = GpuMinHashReadFilter (libnsmh.so, the B200 engine)
// with the same n random numbers.  With one OpenMP thread the greedy contig growth is
// deterministic, so every output stream the stage writes (Contig.tid.0.* and the metadata file)
// must be byte-identical between the arms: the GPU filter returned the same candidate list for
// every window the consensus builder ever asked about.  A third run drives the GPU arm from
// several OpenMP threads (the production setting, Consensus.cpp:29) and checks that it completes
// and accounts for every read.  Prints "CONSENSUS DROPIN OK" on success.
#include <omp.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "Consensus.h"
#include "GpuMinHashReadFilter.h"
#include "ReadAligner.h"

void nsref_fill_read_data(ReadData &rD, const char *bases, const uint64_t *offsets, uint32_t numReads);
extern "C" {
void *nsref_create(const char *bases, const uint64_t *offsets, uint32_t numReads, uint32_t k, uint32_t n,
                   uint32_t thr, const uint64_t *randNumbers, int threads, const char *tmpdir,
                   double *sketch_ms, double *build_ms);
void *nsref_read_filter(void *h);
void nsref_destroy(void *h);
}

namespace fs = std::filesystem;

struct CountingFilter : public ReadFilter {      // counts what the consensus builder asks
    ReadFilter *inner;
    long calls = 0, ids = 0;
    explicit CountingFilter(ReadFilter *f) : inner(f) {}
    void initialize(ReadData &rD) override { inner->initialize(rD); }
    void getFilteredReads(const std::string &s, std::vector<read_t> &results) override {
        inner->getFilteredReads(s, results);
#pragma omp atomic
        ++calls;
#pragma omp atomic
        ids += (long)results.size();
    }
};

static uint64_t fnv1a(const std::string &bytes) {
    uint64_t h = 0xcbf29ce484222325ULL;
    for (unsigned char c : bytes) h = (h ^ c) * 0x100000001b3ULL;
    return h;
}

// file name -> (size, FNV-1a) of everything the stage left in dir
static std::map<std::string, std::pair<size_t, uint64_t>> digest_dir(const std::string &dir) {
    std::map<std::string, std::pair<size_t, uint64_t>> out;
    for (const auto &e : fs::directory_iterator(dir)) {
        if (!e.is_regular_file()) continue;
        std::ifstream in(e.path(), std::ios::binary);
        std::stringstream ss;
        ss << in.rdbuf();
        const std::string bytes = ss.str();
        out[e.path().filename().string()] = {bytes.size(), fnv1a(bytes)};
    }
    return out;
}

static double run_consensus(ReadData &rD, ReadFilter *rF, const std::string &dir, int threads) {
    fs::remove_all(dir);
    fs::create_directories(dir);
    MergeSortReadAligner rA(21, 10);             // main.cpp:125 (unused: use_sort_merge = false)
    Consensus consensus;
    consensus.rD = &rD;
    consensus.rF = rF;
    consensus.rA = &rA;
    consensus.tempDir = dir + "/";
    consensus.tempFileName = "Contig";
    consensus.numThr = threads;
    consensus.m_k = 20;                          // CLI defaults, main.cpp:63-69
    consensus.m_w = 50;
    consensus.max_chain_iter = 400;
    consensus.edge_threshold = 4000000;
    omp_set_num_threads(threads);                // Compressor.cpp:55
    std::ostringstream sink;                     // the stage prints its statistics to std::cout
    std::streambuf *old = std::cout.rdbuf(sink.rdbuf());
    const double t0 = omp_get_wtime();
    try {
        consensus.generateAndWriteConsensus();
    } catch (...) {
        std::cout.rdbuf(old);
        throw;
    }
    std::cout.rdbuf(old);
    return omp_get_wtime() - t0;
}

int main(int argc, char **argv) {
    if (argc < 7) {
        std::fprintf(stderr, "usage: %s reads.bin k n thr tmpdir ref|gpu|both [threads]\n", argv[0]);
        return 2;
    }
    std::ifstream in(argv[1], std::ios::binary);
    uint32_t numReads = 0;
    in.read(reinterpret_cast<char *>(&numReads), 4);
    std::vector<uint64_t> offsets((size_t)numReads + 1);
    in.read(reinterpret_cast<char *>(offsets.data()), offsets.size() * 8);
    std::string bases(offsets[numReads], '\0');
    in.read(&bases[0], bases.size());
    const size_t k = std::atoi(argv[2]), n = std::atoi(argv[3]), thr = std::atoi(argv[4]);
    const std::string tmp = argv[5], mode = argv[6];
    const int threads = argc > 7 ? std::atoi(argv[7]) : 8;
    std::vector<uint64_t> rnd(n);
    nsmh_rand_from_seed(20261017u, (uint32_t)n, rnd.data());

    ReadData rD;
    nsref_fill_read_data(rD, bases.data(), offsets.data(), numReads);
    std::map<std::string, std::pair<size_t, uint64_t>> dref, dgpu;
    long ref_calls = -1, gpu_calls = -1;
    try {
        if (mode == "ref" || mode == "both") {
            void *ref = nsref_create(bases.data(), offsets.data(), numReads, (uint32_t)k, (uint32_t)n, (uint32_t)thr,
                                     rnd.data(), 0, tmp.c_str(), nullptr, nullptr);
            if (!ref) return 3;
            CountingFilter cf(static_cast<ReadFilter *>(nsref_read_filter(ref)));
            const double s = run_consensus(rD, &cf, tmp + "/ref", 1);
            dref = digest_dir(tmp + "/ref");
            ref_calls = cf.calls;
            std::printf("ref: %.2f s, %ld getFilteredReads calls, %ld candidate ids, %zu files\n", s, cf.calls, cf.ids,
                        dref.size());
            nsref_destroy(ref);
        }
        if (mode == "gpu" || mode == "both") {
            GpuMinHashReadFilter gpu;
            gpu.k = k;
            gpu.n = n;
            gpu.overlapSketchThreshold = thr;
            gpu.tempDir = tmp;
            gpu.randNumbers = rnd;
            // DROPIN_DEVICES="0,1,2,3": the one filter spread over several GPUs of this ONE process (tables replicated,
            // the OpenMP threads' queries round-robin over them); the same device may be named twice on a 1-GPU box
            if (const char *dv = std::getenv("DROPIN_DEVICES"))
                for (const char *q = dv; *q;) {
                    gpu.devices.push_back(std::atoi(q));
                    while (*q && *q != ',') ++q;
                    if (*q == ',') ++q;
                }
            if (!gpu.devices.empty()) std::printf("gpu filter on %zu devices\n", gpu.devices.size());
            gpu.initialize(rD);
            CountingFilter cf(&gpu);
            const double s = run_consensus(rD, &cf, tmp + "/gpu", 1);
            dgpu = digest_dir(tmp + "/gpu");
            gpu_calls = cf.calls;
            std::printf("gpu: %.2f s, %ld getFilteredReads calls, %ld candidate ids, %zu files\n", s, cf.calls, cf.ids,
                        dgpu.size());
            if (threads > 1) {                   // the production setting: concurrent callers
                CountingFilter cm(&gpu);
                const double sm = run_consensus(rD, &cm, tmp + "/gpu_mt", threads);
                std::printf("gpu, %d threads: %.2f s, %ld getFilteredReads calls, %zu files\n", threads, sm, cm.calls,
                            digest_dir(tmp + "/gpu_mt").size());
            }
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "consensus_dropin_test: %s\n", e.what());
        return 4;
    }
    for (const auto &kv : (mode == "gpu" ? dgpu : dref))
        std::printf("  %-28s %10zu bytes  fnv %016llx\n", kv.first.c_str(), kv.second.first,
                    (unsigned long long)kv.second.second);
    if (mode == "both") {
        if (dref.empty() || dref != dgpu || ref_calls != gpu_calls) {
            std::printf("CONSENSUS DROPIN MISMATCH\n");
            for (const auto &kv : dgpu)
                std::printf("  gpu %-24s %10zu bytes  fnv %016llx\n", kv.first.c_str(), kv.second.first,
                            (unsigned long long)kv.second.second);
            return 1;
        }
        std::printf("CONSENSUS DROPIN OK\n");
    }
    return 0;
}

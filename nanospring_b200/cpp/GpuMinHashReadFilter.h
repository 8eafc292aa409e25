// GpuMinHashReadFilter — header-only C++ adaptor that makes libnsmh.so a drop-in
// for the reference's read-overlap filter.
//
// It derives from the reference's abstract interface
//     class ReadFilter { virtual void getFilteredReads(const std::string&, std::vector<read_t>&) = 0;
//                        virtual void initialize(ReadData&) = 0; }        (include/ReadFilter.h:15-30)
// and exposes the same public fields as MinHashReadFilter (include/ReadFilter.h:37-42):
//     k, n, overlapSketchThreshold, tempDir
// so the two call sites of the reference need no other change:
//     src/Compressor.cpp:69-76   construct, assign fields, initialize(rD)
//     src/Consensus.cpp:189      rF->getFilteredReads(string, results)   (inside `omp parallel`)
//
// Compile this header inside the reference tree (it includes the reference's own
// ReadFilter.h / ReadData.h / Types.h) and link with -lnsmh.  Errors of the C ABI
// become std::runtime_error, which the reference's main() already catches
// (src/main.cpp:161-176).  No temp files are written; tempDir is kept for parity.
#ifndef GPU_MINHASH_READ_FILTER_H_
#define GPU_MINHASH_READ_FILTER_H_

#include <cstdio>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ReadFilter.h"   // the reference's header: ReadFilter, ReadData, read_t, kMer_t
#include "nsmh.h"

class GpuMinHashReadFilter : public ReadFilter {
public:
    /** [k]-mer **/
    size_t k = 23;
    /** size of sketch **/
    size_t n = 60;
    size_t overlapSketchThreshold = 6;
    std::string tempDir;
    /** CUDA device ordinal **/
    int device = 0;
    /** Optional: several GPUs for this ONE process (the reference is one process with OpenMP threads,
     *  main.cpp:35).  With two or more entries initialize() splits the reads by bases over the devices,
     *  every device sketches its shard and builds the tables of all reads, and getFilteredReads() sends
     *  each calling thread to one of them (include/nsmh.h, nsmh_multi_*).  Empty: `device` alone. */
    std::vector<int> devices;
    /** Optional: fix the n random numbers (tests / reproducible runs).  Left empty, they
     *  are drawn exactly like MinHashReadFilter::generateRandomNumbers (ReadFilter.cpp:49-63). */
    std::vector<kMer_t> randNumbers;

    GpuMinHashReadFilter() = default;
    GpuMinHashReadFilter(const GpuMinHashReadFilter &) = delete;
    GpuMinHashReadFilter &operator=(const GpuMinHashReadFilter &) = delete;
    ~GpuMinHashReadFilter() override { release(); }

    /** ReadFilter.cpp:11-47: sketch every read of rD and build the n tables, on the GPU(s).
     *  The reads come from rD's own 2-bit store when ReadData exposes it (the two accessors of
     *  a quarter of the bytes, no unpacking, no getRead() mutex), else through
     *  getRead() as ASCII. */
    void initialize(ReadData &rD) override {
        if (initializeFromStore(rD, 0)) return;
        const read_t numReads = rD.getNumReads();
        // Pull the reads once, sequentially (in the CLI's low-memory mode getRead() takes a global
        // mutex, ReadData.cpp:225-235), into one pinned ASCII buffer + offsets.
        std::vector<uint64_t> offsets((size_t)numReads + 1, 0);
        std::string s;
        size_t cap = (size_t)rD.avgReadLen * numReads + (size_t)rD.maxReadLen + 1024, used = 0;
        char *bases = nullptr;
        check(nsmh_host_alloc_near(stagingDevice(), cap, reinterpret_cast<void **>(&bases)));
        try {
            for (read_t i = 0; i < numReads; ++i) {
                rD.getRead(i, s);
                if (used + s.size() > cap) {
                    size_t ncap = (used + s.size()) * 2;
                    char *nb = nullptr;
                    check(nsmh_host_alloc_near(stagingDevice(), ncap, reinterpret_cast<void **>(&nb)));
                    std::copy(bases, bases + used, nb);
                    nsmh_host_free(bases);
                    bases = nb;
                    cap = ncap;
                }
                std::copy(s.begin(), s.end(), bases + used);
                used += s.size();
                offsets[i + 1] = used;
            }
            create();
            if (m_) {
                check(nsmh_multi_load_reads_ascii(m_, bases, offsets.data(), numReads));
                sketchAndBuild();
            } else {
                // load + sketch + build pipelined: a chunk is sketched while the next one crosses PCIe
                check(nsmh_initialize_ascii(h_, bases, offsets.data(), numReads));
            }
        } catch (...) {
            nsmh_host_free(bases);
            throw;
        }
        nsmh_host_free(bases);
    }

    /** initialize() from reads that are already 2-bit packed the reference's way (DnaBitset,
     *  dnaToBits.cpp:11-36): `packed` = the concatenated byte-aligned bitsets, lengths[i] = bases of read i.
     *  This is what ReadData holds in memory (readData, high-memory mode) and in tempDir/readBitset
     *  (low-memory mode, the CLI's only mode: main.cpp:40, ReadData.cpp:156-221). */
    template <typename Len>
    void initializeFromBitset(const uint8_t *packed, const Len *lengths, read_t numReads) {
        std::vector<uint32_t> len32((size_t)numReads);
        for (read_t i = 0; i < numReads; ++i) len32[i] = (uint32_t)lengths[i];
        create();
        if (m_) {
            check(nsmh_multi_load_reads_dnabitset(m_, packed, len32.data(), numReads));
            sketchAndBuild();
        } else {
            check(nsmh_initialize_dnabitset(h_, packed, len32.data(), numReads));
        }
    }

    /** The same from the low-memory temp file (ReadData.cpp:181-204 writes it read after read). */
    template <typename Len>
    void initializeFromBitsetFile(const std::string &path, const std::vector<Len> &lengths) {
        size_t bytes = 0;
        for (auto l : lengths) bytes += ((size_t)l + 3) / 4;
        uint8_t *buf = nullptr;
        check(nsmh_host_alloc_near(stagingDevice(), bytes ? bytes : 1, reinterpret_cast<void **>(&buf)));
        FILE *fp = std::fopen(path.c_str(), "rb");
        const size_t got = fp ? std::fread(buf, 1, bytes, fp) : 0;
        if (fp) std::fclose(fp);
        try {
            if (got != bytes) throw std::runtime_error("GpuMinHashReadFilter: cannot read " + path);
            initializeFromBitset(buf, lengths.data(), (read_t)lengths.size());
        } catch (...) {
            nsmh_host_free(buf);
            throw;
        }
        nsmh_host_free(buf);
    }

    /** ReadFilter.cpp:85-97.  Re-entrant; called concurrently by all OpenMP threads. */
    void getFilteredReads(const std::string &s, std::vector<read_t> &results) override {
        results.clear();
        if (!h_ && !m_) throw std::runtime_error("GpuMinHashReadFilter::getFilteredReads before initialize");
        // a consensus window at 30-60x coverage has tens to hundreds of candidates: start large enough that the
        // query is not repeated (NSMH_ERANGE reports the size needed)
        results.resize(results.capacity() > 1024 ? results.capacity() : 1024);
        for (;;) {
            size_t count = 0;
            const int rc = m_ ? nsmh_multi_query_string(m_, s.data(), s.size(), results.data(), results.size(), &count)
                              : nsmh_query_string(h_, s.data(), s.size(), results.data(), results.size(), &count);
            if (rc == NSMH_ERANGE) {
                results.resize(count);
                continue;
            }
            check(rc);
            results.resize(count);
            return;
        }
    }

    /** A window and its reverse complement in one launch sequence (the pair of calls at
     *  Consensus.cpp:185-191).  Optional fast path; results as the two single calls give. */
    void getFilteredReadsPair(const std::string &fwd, const std::string &rev, std::vector<read_t> &resFwd,
                              std::vector<read_t> &resRev) {
        if (!h_) throw std::runtime_error("GpuMinHashReadFilter::getFilteredReadsPair before initialize (single device only)");
        std::string both = fwd + rev;
        uint64_t off[3] = {0, fwd.size(), fwd.size() + rev.size()}, out_off[3] = {0, 0, 0};
        std::vector<read_t> ids(2048);
        for (;;) {
            int rc = nsmh_query_strings(h_, both.data(), off, 2, out_off, ids.data(), ids.size());
            if (rc == NSMH_ERANGE) {
                ids.resize(out_off[2]);
                continue;
            }
            check(rc);
            break;
        }
        resFwd.assign(ids.begin(), ids.begin() + out_off[1]);
        resRev.assign(ids.begin() + out_off[1], ids.begin() + out_off[2]);
    }

    /** The candidate pre-filters of the consensus builder, computed on the device from the packed
     *  reads: flags[i] & NSMH_FLAG_REPETITIVE == Consensus::checkRepetitive(i), i.e. isRepetitive[i]
     *  of Consensus::initialize (Consensus.cpp:405-442); flags[i] & NSMH_FLAG_SHORT == the
     *  `size() < 32` gate of addRelatedReads (Consensus.cpp:213).  Valid after initialize(). */
    void readFlags(std::vector<uint8_t> &flags) {
        if (!h_) throw std::runtime_error("GpuMinHashReadFilter::readFlags before initialize");
        uint32_t numReads = 0;
        check(nsmh_num_reads(h_, &numReads, nullptr));
        flags.assign(numReads, 0);
        check(nsmh_read_flags(h_, flags.data()));
    }

    /** SURVEY 8(f) N2 - ReadData::loadFromFile(fileName, FASTQ|GZIP, low_mem) (ReadData.cpp:12-26,
     *  156-221) and initialize() in one step, without the host 2-bit store: the file is read (and
     *  inflated) on the host, records are split and packed by kernels (csrc/fastq.cu), and the packed
     *  reads are sketched where they are.  No temp file, no getRead() round trip.  getNumReads() /
     *  getRead() below then serve the reads from the device store. */
    void initializeFromFile(const char *fileName, ReadData::Filetype filetype) {
        if (filetype != ReadData::FASTQ && filetype != ReadData::GZIP)
            throw std::runtime_error("GpuMinHashReadFilter::initializeFromFile: FASTQ or GZIP input only");
        const std::vector<int> keep = devices;
        devices.clear();                    // the device-side read store lives on one device
        try {
            create();
        } catch (...) {
            devices = keep;
            throw;
        }
        devices = keep;
        check(nsmh_load_fastq_file(h_, fileName, filetype == ReadData::GZIP ? 1 : 0));
        check(nsmh_sketch(h_));
        check(nsmh_build(h_));
    }

    /** ReadData::getNumReads (ReadData.cpp:223) of the reads on the device. */
    read_t getNumReads() {
        uint32_t numReads = 0;
        if (h_) check(nsmh_num_reads(h_, &numReads, nullptr));
        return numReads;
    }

    /** ReadData::getRead (ReadData.cpp:225-235) from the device's 2-bit store: "ATCG"[code] per base.
     *  Not re-entrant (a bulk accessor; the reference's own getRead takes a global mutex). */
    void getRead(read_t readId, std::string &readStr) {
        if (!h_) throw std::runtime_error("GpuMinHashReadFilter::getRead before initialize");
        if (hostOffsets_.empty()) {
            hostOffsets_.resize((size_t)getNumReads() + 1);
            check(nsmh_read_offsets(h_, hostOffsets_.data()));
        }
        if ((size_t)readId + 1 >= hostOffsets_.size()) throw std::runtime_error("GpuMinHashReadFilter::getRead: bad read id");
        readStr.resize(hostOffsets_[readId + 1] - hostOffsets_[readId]);
        check(nsmh_get_reads_ascii(h_, readId, 1, readStr.empty() ? nullptr : &readStr[0]));
    }

    /** Generates a sequence of n kMer_t random numbers (ReadFilter.cpp:49-63). */
    void generateRandomNumbers(size_t count) {
        std::random_device rd;
        randNumbers.resize(count);
        check(nsmh_rand_from_seed(rd(), (uint32_t)count, randNumbers.data()));
    }

    nsmh_handle handle() const { return h_; }
    nsmh_multi_handle multiHandle() const { return m_; }

private:
    nsmh_handle h_ = nullptr;
    nsmh_multi_handle m_ = nullptr;
    std::vector<uint64_t> hostOffsets_;     // filled by the first getRead()

    void release() {
        nsmh_destroy(h_);
        nsmh_multi_destroy(m_);
        h_ = nullptr;
        m_ = nullptr;
        hostOffsets_.clear();
    }
    void create() {
        release();
        if (randNumbers.size() != n) generateRandomNumbers(n);
        if (devices.size() > 1)
            check(nsmh_multi_create((uint32_t)k, (uint32_t)n, (uint32_t)overlapSketchThreshold, randNumbers.data(),
                                    devices.data(), (int)devices.size(), &m_));
        else
            check(nsmh_create((uint32_t)k, (uint32_t)n, (uint32_t)overlapSketchThreshold, randNumbers.data(),
                              devices.empty() ? device : devices[0], &h_));
    }
    /** the device whose NUMA node the pinned staging buffers are taken from */
    int stagingDevice() const { return devices.empty() ? device : devices[0]; }
    void sketchAndBuild() {
        if (m_) {
            check(nsmh_multi_sketch(m_));
            check(nsmh_multi_build(m_));
        } else {
            check(nsmh_sketch(h_));
            check(nsmh_build(h_));
        }
    }
    // rD's own 2-bit store, when ReadData has the accessors of INTEGRATION.md section 2 (chosen at compile time)
    template <typename RD>
    auto initializeFromStore(RD &rD, int) -> decltype(rD.getBitsetFilePath(), rD.getReadLengths(), bool()) {
        if (rD.getBitsetFilePath().empty()) return false;          // high-memory mode: no temp file
        initializeFromBitsetFile(rD.getBitsetFilePath(), rD.getReadLengths());
        return true;
    }
    template <typename RD>
    bool initializeFromStore(RD &, long) { return false; }

    static void check(int rc) {
        if (rc != NSMH_OK) throw std::runtime_error(std::string("nsmh: ") + nsmh_last_error());
    }
};

#endif  // GPU_MINHASH_READ_FILTER_H_

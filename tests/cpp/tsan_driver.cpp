// TEST INFRASTRUCTURE ONLY.  Race detection for the device code: links the host-emulation harnesses
// (tests/cpp/*_host_emul.cpp, every lane an OS thread) into one executable built with
// -fsanitize=thread and runs each kernel on small random inputs.  ThreadSanitizer then reports every
// pair of accesses by different lanes to the same shared / global word that is not separated by a
// barrier (__syncwarp, __syncthreads, a warp collective) or done with atomics - the class of bug that
// lock-step emulation hides.  Results are checked by the other emulation tests, not here.
// tests/test_emul_tsan.py builds and runs it (make -C oracle tsan).
#include <string>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

extern "C" {
int sketch_emul_run(const uint32_t *W, const uint64_t *off, uint32_t n_reads, uint32_t k, uint32_t n,
                    const uint64_t *rnd, int mode, int lambda_log2, uint32_t tile_words, unsigned grid, uint64_t *sk,
                    unsigned long long *fixups);
void pp_emul_pack_ascii(const uint8_t *src, uint64_t num_bases, uint32_t *W, unsigned grid);
void pp_emul_read_flags(const uint64_t *off, uint32_t n_reads, const uint32_t *W, uint8_t *flags, unsigned grid);
int table_emul_build(const uint64_t *sk, uint32_t rows, uint32_t n, unsigned max_blocks);
int table_emul_build_deferred(uint64_t *sk, uint32_t rows, uint32_t n, unsigned max_blocks, const uint32_t *list,
                              unsigned int count, const uint64_t *vals);
void table_emul_query(const uint64_t *qsk, uint32_t nq, uint32_t thr, unsigned grid, uint32_t *qcount, uint64_t *qpos,
                      uint32_t *tmp_ids, uint64_t tmp_cap, uint32_t *heavy_list, unsigned long long *counters);
void mid_emul_run(const uint64_t *list_off, const uint32_t *ids, uint32_t nq, uint32_t subs, uint32_t thr,
                  unsigned grid, uint32_t *qcount, uint64_t *qpos, uint32_t *mid_ids, uint64_t mid_cap,
                  uint32_t *unresolved, unsigned long long *counters);
void count_emul_run(const uint64_t *list_off, const uint32_t *ids, uint32_t nq, uint32_t subs, uint32_t thr,
                    unsigned grid, uint32_t *qcount, uint64_t *qpos, uint32_t *tmp_ids, uint64_t tmp_cap,
                    uint32_t *heavy_list, unsigned long long *counters);
int fq_emul_parse(const uint8_t *text, uint64_t bytes, uint64_t safe_bytes, unsigned grid, uint32_t pack_iters, uint64_t *offsets,
                  uint64_t offsets_cap, uint32_t *words, uint64_t words_cap, uint32_t *num_reads_out,
                  uint64_t *newlines_out);
}

int main() {
    std::mt19937_64 g(7);
    // reads: a few lengths incl. homopolymer
    std::vector<uint32_t> lens = {0, 22, 23, 700, 1500, 12000, 64, 3000};
    std::vector<uint64_t> off(lens.size() + 1, 0);
    for (size_t i = 0; i < lens.size(); ++i) off[i + 1] = off[i] + lens[i];
    std::vector<uint8_t> bases(off.back() + 64, 'A');
    for (uint64_t i = 0; i < off.back(); ++i) bases[i] = "ACGT"[g() & 3];
    for (uint64_t i = off[3]; i < off[4]; ++i) bases[i] = 'A';       // homopolymer: fix-ups
    uint8_t *ab = bases.data();
    while (reinterpret_cast<uintptr_t>(ab) & 15) ++ab;               // (alignment of the vector is already 16)
    std::vector<uint32_t> W((off.back() + 15) / 16 + 8, 0);
    pp_emul_pack_ascii(bases.data(), off.back(), W.data(), 2);
    std::vector<uint8_t> flags(lens.size(), 0);
    pp_emul_read_flags(off.data(), (uint32_t)lens.size(), W.data(), flags.data(), 2);
    const uint32_t k = 23, n = 60;
    std::vector<uint64_t> rnd(n);
    for (auto &r : rnd) r = g();
    std::vector<uint64_t> sk(lens.size() * n, 0);
    unsigned long long fix = 0;
    for (int mode : {0, 1}) {
        sketch_emul_run(W.data(), off.data(), (uint32_t)lens.size(), k, n, rnd.data(), mode, 2, 640, 2, sk.data(), &fix);
        printf("sketch mode %d done, fixups %llu\n", mode, fix);
    }
    // tables + probing lookup over a sketch matrix with duplicates
    const uint32_t rows = 600;
    std::vector<uint64_t> S((size_t)rows * n);
    for (uint32_t i = 0; i < rows; ++i)
        for (uint32_t j = 0; j < n; ++j) S[(size_t)i * n + j] = (i % 7 == 0) ? 42 + j : (g() % 2000);
    {
        // nsmh_sketch_build: the all-ones entries are left out by the insert kernel and added from a list afterwards
        std::vector<uint64_t> work(S);
        std::vector<uint32_t> list;
        std::vector<uint64_t> vals;
        for (size_t t = 0; t < work.size(); ++t)
            if (t % 5 == 1) { list.push_back((uint32_t)t); vals.push_back(work[t]); work[t] = ~0ULL; }
        list.push_back(0);
        vals.push_back(0);
        table_emul_build_deferred(work.data(), rows, n, 3, list.data(), (unsigned)list.size() - 1, vals.data());
        printf("deferred build done: %s\n", work == S ? "matrix restored" : "MATRIX DIFFERS");
    }
    table_emul_build(S.data(), rows, n, 3);
    {
        std::vector<uint32_t> qcount(rows + 1, 0), tmp(rows * 64 + 1024, 0), heavy(rows + 1, 0);
        std::vector<uint64_t> qpos(rows, 0);
        unsigned long long c[4] = {0, 0, 0, 0};      // [3]: the cursor behind the fixed result places
        table_emul_query(S.data(), rows, 6, 2, qcount.data(), qpos.data(), tmp.data(), tmp.size(), heavy.data(), c);
        printf("lookup done: heavy %llu pairs %llu results %llu\n", c[0], c[1], c[2]);
    }
    // CSR-source lookup body + counting-filter tier
    {
        const uint32_t nq = 6, subs = 60;
        std::vector<uint64_t> lo(1, 0);
        std::vector<uint32_t> ids;
        for (uint32_t q = 0; q < nq; ++q)
            for (uint32_t j = 0; j < subs; ++j) {
                const uint32_t len = (q == 0) ? 1 : (q == 1) ? 3 : (q == 2) ? 12 : (q == 3) ? 40 : (q == 4) ? 16 : 0;
                for (uint32_t t = 0; t < len; ++t) ids.push_back((t == 0 && j % 2) ? 777 : (uint32_t)(g() % 100000));
                lo.push_back(ids.size());
            }
        ids.push_back(0);
        std::vector<uint32_t> qcount(nq + 1, 0), tmp(ids.size() + 8, 0), heavy(nq + 1, 0), mid(ids.size() + 8, 0), unres(nq + 1, 0);
        std::vector<uint64_t> qpos(nq, ~0ULL);
        unsigned long long c[4] = {0, 0, 0, 0};      // [3]: the cursor behind the fixed result places
        count_emul_run(lo.data(), ids.data(), nq, subs, 6, 1, qcount.data(), qpos.data(), tmp.data(), tmp.size() - 8, heavy.data(), c);
        printf("count body done: heavy %llu\n", c[0]);
        std::fill(qcount.begin(), qcount.end(), 0);
        std::fill(qpos.begin(), qpos.end(), ~0ULL);
        unsigned long long m[4] = {0, 0, 0, 0};
        mid_emul_run(lo.data(), ids.data(), nq, subs, 6, 2, qcount.data(), qpos.data(), mid.data(), mid.size() - 8, unres.data(), m);
        printf("mid tier done: unresolved %llu\n", m[0]);
    }
    // FASTQ ingest
    {
        std::string text;
        for (int r = 0; r < 20; ++r) {
            text += "@r\n";
            const int L = (r * 911) % 3000;
            for (int i = 0; i < L; ++i) text += "ACGT"[g() & 3];
            text += "\n+\n";
            text.append(L, 'I');
            text += "\n";
        }
        std::vector<uint8_t> buf(text.size() + 128, '\n');
        memcpy(buf.data(), text.data(), text.size());
        std::vector<uint64_t> offs(64, 0);
        std::vector<uint32_t> words(text.size() / 16 + 16, 0);
        uint32_t nr = 0;
        uint64_t nl = 0;
        fq_emul_parse(buf.data(), text.size(), text.size() + 64, 2, 0, offs.data(), offs.size(), words.data(), words.size() - 8, &nr, &nl);
        printf("fastq done: %u reads\n", nr);
    }
    return 0;
}

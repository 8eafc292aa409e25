/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C) of NanoSpring's MinHash
 * read-overlap path, used solely as the parity checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing under nanospring_b200/ may call into this file.
 *
 * Parity status: PINNED.  oracle/Makefile also compiles the reference's own
 * translation units (oracle/_ref/libnsref.so); tests/test_oracle.py checks this
 * restatement against that build on the reference's CI file and on synthetic
 * reads, and against the golden checksums of SURVEY.md section 8(c) that were
 * produced by the unmodified reference code.
 *
 * Each function cites the reference lines it restates (paths relative to
 * /root/reference).  The code is written from the behaviour, not copied.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* RNG: std::mt19937_64(seed)() stream, which is what generateRandomNumbers()
 * (src/ReadFilter.cpp:49-63) draws from once random_device has produced the
 * 32-bit seed; uniform_int_distribution<unsigned long long> over the full
 * range returns the raw 64-bit outputs.  Published MT19937-64 algorithm
 * (Matsumoto & Nishimura), parameters as in the C++11 standard.           */
void orc_rand_from_seed(uint32_t seed, uint32_t n, uint64_t *out) {
    enum { NN = 312, MM = 156 };
    static const uint64_t MATRIX_A = 0xB5026F5AA96619E9ULL;
    static const uint64_t UM = 0xFFFFFFFF80000000ULL, LM = 0x7FFFFFFFULL;
    uint64_t mt[NN];
    int mti;
    mt[0] = seed;
    for (mti = 1; mti < NN; mti++)
        mt[mti] = 6364136223846793005ULL * (mt[mti - 1] ^ (mt[mti - 1] >> 62)) + (uint64_t)mti;
    for (uint32_t o = 0; o < n; ++o) {
        if (mti >= NN) {
            int i;
            for (i = 0; i < NN; i++) {
                uint64_t x = (mt[i] & UM) | (mt[(i + 1) % NN] & LM);
                mt[i] = mt[(i + MM) % NN] ^ (x >> 1) ^ ((x & 1ULL) ? MATRIX_A : 0ULL);
            }
            mti = 0;
        }
        uint64_t x = mt[mti++];
        x ^= (x >> 29) & 0x5555555555555555ULL;
        x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
        x ^= (x << 37) & 0xFFF7EEE000000000ULL;
        x ^= (x >> 43);
        out[o] = x;
    }
}

/* ------------------------------------------------------------------------- */
/* src/ReadFilter.cpp:113-115 (and src/dnaToBits.cpp:6-9): A0 T1 C2 G3, any
 * other byte through the same two bit tests.                                 */
uint8_t orc_base_to_int(char base) {
    return (uint8_t)((base & 2) | ((base & 4) >> 2));
}

/* src/ReadFilter.cpp:101-109: first base ends up most significant. */
uint64_t orc_kmer_to_int(const char *s, size_t len) {
    uint64_t v = 0;
    for (size_t i = 0; i < len; ++i) v = (v << 2) | orc_base_to_int(s[i]);
    return v;
}

/* src/ReadFilter.cpp:138-152: rolling forward k-mers, masked to 2k bits.
 * Returns the number written (len-k+1, or 0).  k in [1,31] (k = 32 is UB in
 * the reference).                                                */
size_t orc_string2kmers(const char *s, size_t len, uint32_t k, uint64_t *kmers) {
    if (len < k) return 0;
    const uint64_t mask = (1ULL << (2 * k)) - 1;
    uint64_t cur = orc_kmer_to_int(s, k);
    kmers[0] = cur;
    size_t cnt = len - k + 1;
    for (size_t i = 1; i < cnt; ++i) {
        cur = ((cur << 2) | orc_base_to_int(s[i + k - 1])) & mask;
        kmers[i] = cur;
    }
    return cnt;
}

/* src/ReadFilter.cpp:117-136: sketch[l] = min_i (kmer_i XOR rand[l]) as an
 * unsigned 64-bit compare (std::hash<uint64_t> is the identity in libstdc++).
 * len < k-1: sketch is left untouched; len == k-1: every slot becomes all-ones
 * (SURVEY S5).  Callers zero-initialise, as the reference's vectors do
 * (ReadFilter.cpp:21, :88).                                                 */
void orc_string2sketch(const char *s, size_t len, uint32_t k, uint32_t n, const uint64_t *rnd,
                       uint64_t *sketch) {
    if (len + 1 < (size_t)k) return;
    for (uint32_t l = 0; l < n; ++l) sketch[l] = ~0ULL;
    if (len < k) return;
    const uint64_t mask = (1ULL << (2 * k)) - 1;
    uint64_t cur = orc_kmer_to_int(s, k);
    size_t cnt = len - k + 1;
    for (size_t i = 0;;) {
        for (uint32_t l = 0; l < n; ++l) {
            uint64_t h = cur ^ rnd[l];
            if (h < sketch[l]) sketch[l] = h;
        }
        if (++i == cnt) break;
        cur = ((cur << 2) | orc_base_to_int(s[i + k - 1])) & mask;
    }
}

/* The sketch loop of initialize(), src/ReadFilter.cpp:21, 31-44:
 * row-major [read][hash], zero-initialised.                                 */
void orc_sketch_all(const char *bases, const uint64_t *offsets, uint32_t num_reads, uint32_t k,
                    uint32_t n, const uint64_t *rnd, uint64_t *sketches) {
    memset(sketches, 0, (size_t)num_reads * n * sizeof(uint64_t));
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < (int64_t)num_reads; ++i)
        orc_string2sketch(bases + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), k, n, rnd,
                          sketches + (size_t)i * n);
}

/* include/ReadData.h:163-172 + src/ReadData.cpp:247-260: reverse, then A<->T,
 * C<->G, every other byte unchanged.                                        */
void orc_reverse_complement(const char *s, size_t len, char *out) {
    for (size_t i = 0; i < len; ++i) {
        char c = s[len - 1 - i];
        out[i] = c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c;
    }
}

/* ------------------------------------------------------------------------- */
/* Tables.  src/BBHashMap.cpp:10-99 builds, per hash function j, the distinct
 * keys of column j, a start array and the read ids of each key in ascending
 * read order.  The minimal perfect hash in between is only an index: keys are
 * stored and verified on lookup (BBHashMap.cpp:105-106), so any exact
 * key -> id-list dictionary gives the same answers.  Here: sort
 * (key, id), run-length, binary search.                                     */
typedef struct {
    uint32_t num_keys;
    uint64_t *keys;   /* [num_keys] ascending */
    uint32_t *start;  /* [num_keys + 1] */
    uint32_t *ids;    /* [num_reads] grouped by key, ascending inside a group */
} orc_table;

typedef struct {
    uint32_t n, num_reads;
    orc_table *t;
} orc_tables;

typedef struct { uint64_t key; uint32_t id; } orc_pair;

static int pair_cmp(const void *a, const void *b) {
    const orc_pair *x = (const orc_pair *)a, *y = (const orc_pair *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->id < y->id ? -1 : (x->id > y->id);
}

orc_tables *orc_build_tables(const uint64_t *sketches, uint32_t num_reads, uint32_t n) {
    orc_tables *T = (orc_tables *)calloc(1, sizeof(*T));
    T->n = n;
    T->num_reads = num_reads;
    T->t = (orc_table *)calloc(n ? n : 1, sizeof(orc_table));
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t j = 0; j < (int64_t)n; ++j) {
        orc_pair *p = (orc_pair *)malloc((num_reads ? num_reads : 1) * sizeof(orc_pair));
        for (uint32_t r = 0; r < num_reads; ++r) {
            p[r].key = sketches[(size_t)r * n + j];
            p[r].id = r;
        }
        qsort(p, num_reads, sizeof(orc_pair), pair_cmp);
        orc_table *t = &T->t[j];
        t->keys = (uint64_t *)malloc((num_reads ? num_reads : 1) * sizeof(uint64_t));
        t->start = (uint32_t *)malloc(((size_t)num_reads + 1) * sizeof(uint32_t));
        t->ids = (uint32_t *)malloc((num_reads ? num_reads : 1) * sizeof(uint32_t));
        uint32_t nk = 0;
        for (uint32_t r = 0; r < num_reads; ++r) {
            if (r == 0 || p[r].key != p[r - 1].key) {
                t->keys[nk] = p[r].key;
                t->start[nk] = r;
                nk++;
            }
            t->ids[r] = p[r].id;
        }
        t->start[nk] = num_reads;
        t->num_keys = nk;
        free(p);
    }
    return T;
}

void orc_free_tables(orc_tables *T) {
    if (!T) return;
    for (uint32_t j = 0; j < T->n; ++j) {
        free(T->t[j].keys);
        free(T->t[j].start);
        free(T->t[j].ids);
    }
    free(T->t);
    free(T);
}

uint32_t orc_table_num_keys(const orc_tables *T, uint32_t j) { return T->t[j].num_keys; }

/* src/BBHashMap.cpp:101-120: exact lookup, returns the id range of `key`. */
static int table_find(const orc_table *t, uint64_t key, uint32_t *b, uint32_t *e) {
    uint32_t lo = 0, hi = t->num_keys;
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (t->keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    if (lo == t->num_keys || t->keys[lo] != key) return 0;
    *b = t->start[lo];
    *e = t->start[lo + 1];
    return 1;
}

static int u32_cmp(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : (x > y);
}

/* src/ReadFilter.cpp:65-83: gather the n id lists, sort, keep ids that occur at
 * least `thr` times; output ascending, includes the query read itself.
 * Returns a malloc'ed array (caller frees) and its length in *count.        */
uint32_t *orc_query_sketch(const orc_tables *T, const uint64_t *sketch, uint32_t thr,
                           size_t *count) {
    size_t total = 0;
    for (uint32_t j = 0; j < T->n; ++j) {
        uint32_t b, e;
        if (table_find(&T->t[j], sketch[j], &b, &e)) total += e - b;
    }
    uint32_t *m = (uint32_t *)malloc((total ? total : 1) * sizeof(uint32_t));
    size_t pos = 0;
    for (uint32_t j = 0; j < T->n; ++j) {
        uint32_t b, e;
        if (table_find(&T->t[j], sketch[j], &b, &e)) {
            memcpy(m + pos, T->t[j].ids + b, (size_t)(e - b) * sizeof(uint32_t));
            pos += e - b;
        }
    }
    qsort(m, total, sizeof(uint32_t), u32_cmp);
    size_t out = 0;
    for (size_t i = 0; i < total;) {
        size_t j = i + 1;
        while (j < total && m[j] == m[i]) ++j;
        if (j - i >= (size_t)thr) m[out++] = m[i];
        i = j;
    }
    *count = out;
    return m;
}

/* src/ReadFilter.cpp:85-97: public string overload; zero-initialised sketch,
 * so strings shorter than k-1 query with the all-zero sketch.               */
uint32_t *orc_query_string(const orc_tables *T, const char *s, size_t len, uint32_t k,
                           const uint64_t *rnd, uint32_t thr, size_t *count) {
    uint64_t *sk = (uint64_t *)calloc(T->n ? T->n : 1, sizeof(uint64_t));
    orc_string2sketch(s, len, k, T->n, rnd, sk);
    uint32_t *r = orc_query_sketch(T, sk, thr, count);
    free(sk);
    return r;
}

/* Bulk mode (the loop the production caller spreads over windows,
 * src/Consensus.cpp:180-191): every read queried against the tables, either
 * with its stored sketch (rc = 0) or as its reverse-complement string (rc = 1).
 * offsets_out[num_reads+1]; *ids_out malloc'ed.                              */
int orc_query_all(const orc_tables *T, const char *bases, const uint64_t *offsets,
                  const uint64_t *sketches, uint32_t k, const uint64_t *rnd, uint32_t thr, int rc,
                  uint64_t *offsets_out, uint32_t **ids_out) {
    const uint32_t N = T->num_reads;
    uint32_t **res = (uint32_t **)calloc(N ? N : 1, sizeof(uint32_t *));
    size_t *cnt = (size_t *)calloc(N ? N : 1, sizeof(size_t));
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < (int64_t)N; ++i) {
        if (!rc) {
            res[i] = orc_query_sketch(T, sketches + (size_t)i * T->n, thr, &cnt[i]);
        } else {
            /* The caller never sees the raw input bytes: ReadData stores 2 bits per base
             * (dnaToBits.cpp:11-36) and getRead() returns the letters "ATCG"[code]
             * (ReadData.cpp:225-235, dnaToBits.cpp:81-98).  That string is what gets
             * reverse-complemented, so every base is complemented, N included (N -> G -> C). */
            size_t len = (size_t)(offsets[i + 1] - offsets[i]);
            char *canon = (char *)malloc(len ? len : 1);
            char *buf = (char *)malloc(len ? len : 1);
            for (size_t j = 0; j < len; ++j) canon[j] = "ATCG"[orc_base_to_int(bases[offsets[i] + j])];
            orc_reverse_complement(canon, len, buf);
            res[i] = orc_query_string(T, buf, len, k, rnd, thr, &cnt[i]);
            free(canon);
            free(buf);
        }
    }
    uint64_t total = 0;
    for (uint32_t i = 0; i < N; ++i) {
        offsets_out[i] = total;
        total += cnt[i];
    }
    offsets_out[N] = total;
    uint32_t *ids = (uint32_t *)malloc((total ? total : 1) * sizeof(uint32_t));
    for (uint32_t i = 0; i < N; ++i) {
        memcpy(ids + offsets_out[i], res[i], cnt[i] * sizeof(uint32_t));
        free(res[i]);
    }
    free(res);
    free(cnt);
    *ids_out = ids;
    return 0;
}

void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------------- */
/* Candidate pre-filters of the consensus builder.
 * Consensus::checkRepetitive (src/Consensus.cpp:405-424): the read comes out of the 2-bit
 * store ("ATCG"[code], dnaToBits.cpp:81-98); for every shift i in 1..6 count the positions j
 * with read[j] == read[(j+i) % len]; repetitive when a count exceeds 0.7 * (double)len.
 * The length gate `readStr1.size() < 32` is src/Consensus.cpp:213.
 * flags[r]: bit 0 repetitive, bit 1 shorter than 32 bases.                 */
void orc_read_flags(const char *bases, const uint64_t *offsets, uint32_t num_reads, uint8_t *flags) {
#pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t r = 0; r < num_reads; ++r) {
        const char *s = bases + offsets[r];
        const size_t len = (size_t)(offsets[r + 1] - offsets[r]);
        int rep = 0;
        for (size_t i = 1; i <= 6 && !rep; ++i) {
            size_t same = 0;
            for (size_t j = 0; j < len; ++j)
                if (orc_base_to_int(s[j]) == orc_base_to_int(s[(j + i) % len])) ++same;
            if ((double)same > 0.7 * (double)len) rep = 1;
        }
        flags[r] = (uint8_t)(rep | (len < 32 ? 2 : 0));
    }
}

/* ------------------------------------------------------------------------- */
/* FASTQ ingest: which bytes of a FASTQ text become the reads.
 * Restates the record loop of ReadData::loadFromFastqFile_lowmem
 * (src/ReadData.cpp:177-198; the CLI fixes low_mem = true, src/main.cpp:40):
 *     while (getline(fin, line)) { getline(fin, line); <line is the read>; getline; getline; }
 * with std::getline's behaviour restated explicitly: a call succeeds iff at least one
 * byte (possibly just the '\n') is left; the returned line excludes the '\n'; a last
 * line without '\n' is returned and sets eofbit; and a call made with eofbit already
 * set fails WITHOUT clearing its string argument.  The last point is observable: when
 * the text ends inside a header line (no '\n' after it), the second getline leaves the
 * header bytes in `line` and they are stored as that record's read (the high-memory
 * loader, :113-127, reads into a different string there and stores an empty or stale
 * read instead; that loader is not reachable from the CLI).  No '\r' handling, no
 * check of '@' / '+': every 4 lines, the second one, any bytes.
 * start[i]/len[i] = byte range of read i inside text.  Returns the number of reads
 * (records); only the first `cap` entries are written.                        */
typedef struct { const char *t; size_t n, pos; int eof; } orc_lines;

/* returns 1 and sets [*b, *e) on success; 0 on failure with [*b,*e) untouched */
static int orc_getline(orc_lines *L, size_t *b, size_t *e) {
    if (L->eof) return 0;                       /* sentry fails, string not erased */
    if (L->pos >= L->n) {                       /* erase(), nothing extracted -> failbit|eofbit */
        L->eof = 1;
        *b = *e = L->n;
        return 0;
    }
    const char *nl = memchr(L->t + L->pos, '\n', L->n - L->pos);
    *b = L->pos;
    if (nl) {
        *e = (size_t)(nl - L->t);
        L->pos = *e + 1;
    } else {
        *e = L->n;
        L->pos = L->n;
        L->eof = 1;
    }
    return 1;
}

uint64_t orc_fastq_index(const char *text, size_t bytes, uint64_t *start, uint64_t *len, uint64_t cap) {
    orc_lines L = {text, bytes, 0, 0};
    size_t b = 0, e = 0, tb, te;
    uint64_t reads = 0;
    while (orc_getline(&L, &b, &e)) {           /* header -> line */
        orc_getline(&L, &b, &e);                /* read -> the same string */
        if (reads < cap) {
            start[reads] = b;
            len[reads] = e - b;
        }
        ++reads;
        tb = te = 0;
        orc_getline(&L, &tb, &te);              /* '+' line (dropped) */
        orc_getline(&L, &tb, &te);              /* quality line (dropped) */
    }
    return reads;
}

/* What ReadData::getRead hands back for stored bytes: the 2-bit code of every byte
 * (src/dnaToBits.cpp:46-71) rendered through "ATCG"[code] (src/dnaToBits.cpp:81-98). */
void orc_store_roundtrip(const char *in, size_t n, char *out) {
    static const char tab[4] = {'A', 'T', 'C', 'G'};
    for (size_t i = 0; i < n; ++i) out[i] = tab[orc_base_to_int(in[i])];
}

/* ------------------------------------------------------------------------- */
/* FNV-1a-64 over the little-endian bytes of u64 words: the checksum SURVEY.md
 * section 8(c) uses for its golden table.                                    */
uint64_t orc_fnv1a64_u64(const uint64_t *p, size_t n, uint64_t h) {
    for (size_t i = 0; i < n; ++i) {
        uint64_t v = p[i];
        for (int b = 0; b < 8; ++b) {
            h ^= (v >> (8 * b)) & 0xFF;
            h *= 0x100000001b3ULL;
        }
    }
    return h;
}

/* Candidate-set checksum: per read, its count then each id, widened to u64. */
uint64_t orc_fnv1a64_csr(const uint64_t *offsets, const uint32_t *ids, uint32_t num_reads) {
    uint64_t h = 0xcbf29ce484222325ULL;
    for (uint32_t i = 0; i < num_reads; ++i) {
        uint64_t c = offsets[i + 1] - offsets[i];
        h = orc_fnv1a64_u64(&c, 1, h);
        for (uint64_t j = offsets[i]; j < offsets[i + 1]; ++j) {
            uint64_t v = ids[j];
            h = orc_fnv1a64_u64(&v, 1, h);
        }
    }
    return h;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int t) {
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

/*
 * nsmh.h — C ABI of the B200-native MinHash read-overlap engine (libnsmh.so).
 *
 * This is the drop-in boundary for NanoSpring's read-overlap stage: the entry
 * points below are what an FFI binding of the reference's
 *     class ReadFilter / class MinHashReadFilter   (include/ReadFilter.h:15-139)
 * would bind.  Plain pointers and sizes only; no C++ or torch types.  The C++
 * adaptor nanospring_b200/cpp/GpuMinHashReadFilter.h and the Python mirror
 * nanospring_b200/filter.py sit on top of exactly these symbols.
 *
 * Reference semantics kept bit-for-bit (file:line relative to the reference):
 *  - base code A0 T1 C2 G3 via (c&2)|((c&4)>>2) for ANY byte     ReadFilter.cpp:113-115
 *  - forward k-mers, first base most significant, 1 <= k <= 31    ReadFilter.cpp:138-152
 *  - sketch[l] = min over k-mers x of (x XOR rand[l]), u64        ReadFilter.cpp:117-136
 *  - len <  k-1 : sketch all-zero;  len == k-1 : all-ones         ReadFilter.cpp:21,119-124
 *  - table j maps sketch[.][j] -> read ids                        BBHashMap.cpp:10-120
 *  - query: ids occurring in >= thr of the n probed lists,
 *    ascending, query read itself included                        ReadFilter.cpp:65-83
 *  - the n random numbers are an INPUT (the reference draws them from
 *    std::random_device, ReadFilter.cpp:49-63); nsmh_rand_from_seed() gives the
 *    numbers std::mt19937_64(seed) would have produced.
 *
 * Every function returns NSMH_OK (0) or a negative NSMH_E* code; the message is
 * available from nsmh_last_error() (thread-local).  There is no CPU fallback:
 * without a CUDA device every compute entry point fails with NSMH_ECUDA.
 *
 * Thread safety: one handle may be shared by many host threads for
 * nsmh_query_string / nsmh_query_strings / nsmh_query_sketches after
 * nsmh_build() (the contract of getFilteredReads, called concurrently from the
 * OpenMP region at Consensus.cpp:29,189).  All other calls on one handle must
 * be serialised by the caller (initialize() is single-threaded, Compressor.cpp:76).
 */
#ifndef NSMH_H_
#define NSMH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSMH_OK 0
#define NSMH_EINVAL (-1)   /* bad argument (k outside 1..31, n == 0, null pointer ...) */
#define NSMH_ECUDA (-2)    /* CUDA runtime error / no device */
#define NSMH_ESTATE (-3)   /* call order violated (e.g. query before build) */
#define NSMH_ENOMEM (-4)   /* host or device allocation failed */
#define NSMH_ERANGE (-5)   /* caller's output buffer too small; required size reported */

typedef struct nsmh_ctx *nsmh_handle;

/* ---- lifetime / parameters (MinHashReadFilter fields k, n, overlapSketchThreshold:
 *      ReadFilter.h:37-42; CLI defaults 23/60/6: main.cpp:57-62) ------------------- */
int nsmh_create(uint32_t k, uint32_t n, uint32_t overlap_sketch_thr, const uint64_t *rand_numbers,
                int device, nsmh_handle *out);
/* Change k / n / threshold / random numbers of an existing handle.  Loaded reads stay on the
 * device (the reference loads the reads first, Compressor.cpp:60, and configures the filter
 * afterwards, :69-76); sketches and tables are dropped and must be recomputed. */
int nsmh_set_params(nsmh_handle h, uint32_t k, uint32_t n, uint32_t overlap_sketch_thr,
                    const uint64_t *rand_numbers);
int nsmh_destroy(nsmh_handle h);
const char *nsmh_last_error(void);
const char *nsmh_version(void);

/* First n outputs of std::mt19937_64(seed): what generateRandomNumbers()
 * (ReadFilter.cpp:49-63) draws once random_device has returned `seed`.  Host only. */
int nsmh_rand_from_seed(uint32_t seed, uint32_t n, uint64_t *out);

/* Pinned host memory for staging reads / results (optional; any host memory works). */
int nsmh_host_alloc(size_t bytes, void **out);
int nsmh_host_free(void *p);
/* The same on the NUMA node that `device` hangs off (sysfs numa_node of its PCI function; plain pinned memory
 * when the box has no such information).  On a two-socket host a staging buffer on the far node sends every
 * byte over the socket interconnect before PCIe - with several GPUs loading at once that link sets the rate. */
int nsmh_host_alloc_near(int device, size_t bytes, void **out);
/* Restricts the CALLING thread (and the threads it creates afterwards) to the CPUs of that node; *numa_node
 * (optional) receives the node, -1 when there is nothing to bind to.  For one-process-per-GPU launches. */
int nsmh_bind_thread_near(int device, int *numa_node);

/* ---- reads -> device (replaces ReadData::getRead + DnaBitset on the host,
 *      ReadData.cpp:225-235, dnaToBits.cpp:11-103) --------------------------------- */
/* Concatenated ASCII bases of num_reads reads; read i = bases[offsets[i] .. offsets[i+1]).
 * Host buffers.  Copies host->device in chunks overlapped with the on-device 2-bit pack. */
int nsmh_load_reads_ascii(nsmh_handle h, const char *bases, const uint64_t *offsets,
                          uint32_t num_reads);
/* Same, but both buffers are already device-resident (device pointers). */
int nsmh_load_reads_ascii_device(nsmh_handle h, const char *d_bases, const uint64_t *d_offsets,
                                 uint32_t num_reads, uint64_t total_bases);
/* Reads already 2-bit packed in the reference's DnaBitset layout (4 bases per byte,
 * first base in bits 7..6, every read byte aligned; dnaToBits.cpp:11-36), concatenated.
 * lengths[i] = bases of read i.  Host buffers.  Bytes are taken as exact A/C/G/T. */
int nsmh_load_reads_dnabitset(nsmh_handle h, const uint8_t *packed, const uint32_t *lengths,
                              uint32_t num_reads);
/* MinHashReadFilter::initialize(ReadData&) in one call (ReadFilter.cpp:11-47: sketch every read, then
 * populateHashTables) = nsmh_load_reads_* + nsmh_sketch + nsmh_build, pipelined: the reads that a chunk of
 * the copy completes are packed and sketched while the following chunks are still crossing PCIe, and the
 * tables are built behind the last chunk.  Returns as soon as the last byte has left the caller's buffers;
 * the device finishes in stream order, so any later call on the handle sees the finished tables.  Results
 * are those of the three separate calls, bit for bit (tests/test_gpu_parity.py). */
int nsmh_initialize_ascii(nsmh_handle h, const char *bases, const uint64_t *offsets, uint32_t num_reads);
int nsmh_initialize_dnabitset(nsmh_handle h, const uint8_t *packed, const uint32_t *lengths,
                              uint32_t num_reads);
/* The first two stages only (load + sketch, pipelined the same way): for callers that build the tables
 * elsewhere - the multi-GPU flows below, which partition the tables by hash function across ranks. */
int nsmh_load_sketch_ascii(nsmh_handle h, const char *bases, const uint64_t *offsets, uint32_t num_reads);
int nsmh_load_sketch_dnabitset(nsmh_handle h, const uint8_t *packed, const uint32_t *lengths,
                               uint32_t num_reads);
int nsmh_num_reads(nsmh_handle h, uint32_t *num_reads, uint64_t *total_bases);

/* ---- FASTQ ingest on the device (replaces ReadData::loadFromFile for
 *      FASTQ / GZIP input, ReadData.cpp:12-26, 156-221, and the temp-file + mutex getRead,
 *      ReadData.cpp:225-235).  Record semantics are the reference loader's: lines split at
 *      '\n' only, record i = lines 4i..4i+3, its second line is the read (any bytes, no
 *      '\r' stripping, no '@'/'+' checks), a last line without '\n' counts; a text that ends
 *      inside a header line stores that header's bytes as the last read, as the reference
 *      does.  The reads end up exactly as after nsmh_load_reads_ascii. ------------------- */
/* Whole (inflated) FASTQ text in host memory: copied once, parsed and packed on the device. */
int nsmh_load_fastq(nsmh_handle h, const char *text, size_t bytes);
/* Same with the text already device-resident (device pointer; read-only, may be freed after). */
int nsmh_load_fastq_device(nsmh_handle h, const char *d_text, size_t bytes);
/* ReadData::loadFromFile(path, gzip ? GZIP : FASTQ): the file is read (and inflated with zlib
 * when gzip != 0, concatenated members included) on the host in pinned chunks that stream to the
 * device while the next chunk is produced; no temp file is written. */
int nsmh_load_fastq_file(nsmh_handle h, const char *path, int gzip);
/* Base offsets of the loaded reads: offsets[i] .. offsets[i+1] = read i, host u64 [num_reads+1]. */
int nsmh_read_offsets(nsmh_handle h, uint64_t *offsets);
/* ReadData::getRead for reads first .. first+count-1, concatenated: every base rendered as
 * "ATCG"[code] from the device's 2-bit store (dnaToBits.cpp:81-98).  out: host buffer of
 * offsets[first+count] - offsets[first] bytes.  Serialised with the other non-query calls. */
int nsmh_get_reads_ascii(nsmh_handle h, uint32_t first, uint32_t count, char *out);

/* ---- MinHashReadFilter::initialize() (ReadFilter.cpp:11-47), split in its two stages --- */
/* Sketch every loaded read: sketches[read][hash], row-major u64, on the device. */
int nsmh_sketch(nsmh_handle h);
/* mode: 0 = filtered (prefix-filter kernel + exact fix-up, default), 1 = brute force
 * (every k-mer x every hash, the reference's operation count).  Results are identical. */
int nsmh_set_sketch_mode(nsmh_handle h, int mode);
int nsmh_get_sketches(nsmh_handle h, uint64_t *out /* [num_reads * n], host */);
/* Device address of the local sketch matrix (for NCCL all-gather by the host program). */
int nsmh_sketches_device_ptr(nsmh_handle h, uint64_t **d_ptr);
/* Use an external device-resident sketch matrix [table_reads][n] as the table contents
 * (multi-GPU: the all-gathered sketches).  Local read i is global read id_base + i.
 * The buffer must stay valid until nsmh_build() returns. */
int nsmh_set_table_sketches(nsmh_handle h, const uint64_t *d_sketches, uint32_t table_reads,
                            uint32_t id_base);
/* populateHashTables() (ReadFilter.cpp:159-172): build the n key -> read-id tables. */
int nsmh_build(nsmh_handle h);
/* nsmh_sketch + nsmh_build as one call (the body of MinHashReadFilter::initialize once the reads are on the
 * device, ReadFilter.cpp:21-47): the exact fix-up pass of the sketch runs on a second stream beside the table
 * insert, and the few entries it produces are inserted from a list afterwards.  Sketches and tables are those of
 * the two separate calls (tests/test_gpu_parity.py::test_sketch_build_overlap_equals_separate_calls). */
int nsmh_sketch_build(nsmh_handle h);
/* Number of distinct keys of table j (BBHashMap::numKeys, BBHashMap.cpp:35). */
int nsmh_table_num_keys(nsmh_handle h, uint32_t j, uint32_t *num_keys);

/* ---- bulk query: every loaded read against the tables --------------------------------
 * rc = 0: the read's own sketch (getFilteredReads(sketch[]), ReadFilter.cpp:65-83);
 * rc = 1: the read's reverse-complement string (Consensus.cpp:181-191).
 * Result is CSR, device resident until the next bulk query: offsets[num_reads+1], ids[total]. */
int nsmh_query_all(nsmh_handle h, int rc, uint64_t *total_ids);
int nsmh_query_all_result(nsmh_handle h, uint64_t *offsets /* [num_reads+1] host */,
                          uint32_t *ids /* [total_ids] host */);
int nsmh_query_all_device_ptrs(nsmh_handle h, uint64_t **d_offsets, uint32_t **d_ids);

/* ---- candidate pre-filters of the consensus builder, on the device ----------------------
 * flags[i] of loaded read i:
 *   NSMH_FLAG_REPETITIVE  Consensus::checkRepetitive(i) (Consensus.cpp:405-424): for some shift
 *                         s in 1..6 more than 0.7*len positions j have read[j] == read[(j+s) % len];
 *                         this is isRepetitive[i] of Consensus::initialize (Consensus.cpp:426-442)
 *   NSMH_FLAG_SHORT       len < 32, the length gate of addRelatedReads (Consensus.cpp:213)
 * Computed from the packed reads on the first call after a load (one pass over 0.25 B/base).
 * nsmh_query_all_drop removes from the bulk CSR of the last nsmh_query_all every candidate id
 * whose flags intersect drop_mask - what the caller does one id at a time at
 * Consensus.cpp:204-216 - so that they never leave the device.  Single-GPU bulk results only
 * (candidate ids must be ids of the loaded reads). */
#define NSMH_FLAG_REPETITIVE 1u
#define NSMH_FLAG_SHORT 2u
int nsmh_read_flags(nsmh_handle h, uint8_t *out /* [num_reads] host, may be NULL */);
int nsmh_read_flags_device_ptr(nsmh_handle h, uint8_t **d_flags);
int nsmh_query_all_drop(nsmh_handle h, uint32_t drop_mask, uint64_t *total_ids);

/* ---- multi-GPU building blocks (tables partitioned by hash function across ranks) -----
 * nsmh_probe_lists: probe this handle's n tables for num_queries device-resident sketches
 * [num_queries][n] and gather the id lists WITHOUT counting: CSR of concatenated lists
 * (fetch with nsmh_query_all_result / nsmh_query_all_device_ptrs).
 * nsmh_count_lists: for num_queries queries, `parts` partial lists each (device CSR pieces
 * d_offsets[p][num_queries+1], d_ids[p], e.g. received from the ranks that own the tables),
 * emit the ids that occur in at least overlap_sketch_thr lists, ascending (ReadFilter.cpp:73-82).
 * Result is the bulk CSR, as after nsmh_query_all.  parts <= 16. */
int nsmh_probe_lists(nsmh_handle h, const uint64_t *d_sketches, uint32_t num_queries, uint64_t *total_ids);
int nsmh_count_lists(nsmh_handle h, uint32_t num_queries, uint32_t parts, const uint64_t *const *d_offsets,
                     const uint32_t *const *d_ids, uint64_t *total_ids);

/* ---- multi-GPU, native: one process per GPU, exchanges over NVLink peer memory -------------
 * The reference has no distributed path (ReadFilter.cpp:31-44 is one OpenMP loop over the
 * reads, :163-165 one over the tables); this is the only exchange the path needs.  Rank g owns
 * the tables of a contiguous block of hash functions for ALL reads of all ranks:
 *   sketch (local reads) -> every rank STORES its sketch columns straight into the owners'
 *   memory -> flag barrier -> owners build their tables and probe them for every read,
 *   storing {id | group start, group size} straight into the memory of the rank that owns
 *   the read -> flag barrier -> every rank thresholds its own reads (group ids are read from
 *   the owner's memory on demand).  No host round trip between the stages, no NCCL call.
 * Results are bit-identical to one GPU holding all reads (global read id = rows of the lower
 * ranks + local id).  The work per rank is constant under weak scaling.
 *
 * Usage (all ranks): nsmh_mg_init -> exchange the tokens by any means (e.g. an all-gather of
 * NSMH_MG_TOKEN_BYTES bytes) -> nsmh_mg_connect -> per batch: nsmh_load_reads_* , nsmh_sketch,
 * nsmh_mg_run, nsmh_query_all_result / _device_ptrs.  rows_per_rank[r] = reads of rank r; the
 * batch loaded on this rank must have exactly rows_per_rank[rank] reads.  world <= 16.
 * Every rank must call nsmh_mg_run the same number of times; a peer that never arrives makes
 * the call fail with NSMH_ECUDA after a timeout instead of hanging the GPU. */
#define NSMH_MG_MAX_RANKS 16
#define NSMH_MG_TOKEN_BYTES 256
int nsmh_mg_init(nsmh_handle h, uint32_t rank, uint32_t world, const uint32_t *rows_per_rank,
                 void *token_out /* NSMH_MG_TOKEN_BYTES */);
int nsmh_mg_connect(nsmh_handle h, const void *tokens /* world * NSMH_MG_TOKEN_BYTES, rank order */);
int nsmh_mg_run(nsmh_handle h, uint64_t *total_ids);
/* nsmh_sketch + nsmh_mg_run as one call: the sketch's exact fix-up pass runs on a second stream beside the column
 * scatter, the entries it produces are sent after it.  Same results (bench.py's parity check at N > 1 runs on it). */
int nsmh_mg_sketch_run(nsmh_handle h, uint64_t *total_ids);
/* Device time of the stages of the last nsmh_mg_run, ms: scatter columns, barrier, build owned
 * tables, probe + store to peers, barrier, count. */
int nsmh_mg_stage_ms(nsmh_handle h, float *out /* [6] */);
/* Detach from the peers and free the exchange memory (also done by nsmh_destroy).  All ranks
 * must have finished their last nsmh_mg_run before any rank calls this. */
int nsmh_mg_shutdown(nsmh_handle h);

/* ---- one process, several GPUs (multidev.cu) ------------------------------------------------
 * The reference is ONE process with OpenMP threads (main.cpp:35, Compressor.cpp:55): initialize()
 * once (Compressor.cpp:69-76), getFilteredReads() concurrently from every thread (Consensus.cpp:29,189).
 * A multi handle gives that caller all GPUs of the box: the reads are split by bases over `ndev`
 * devices and sketched in parallel, every device then pulls the other shards' sketch rows over
 * NVLink and builds the FULL tables (global read ids), so that ANY device answers an online query
 * alone - callers are spread over the devices round robin - and the bulk query of shard d runs on
 * device d.  Results are those of one device holding all reads.  devices[] may name a device more
 * than once (tests on one GPU). */
typedef struct nsmh_multi *nsmh_multi_handle;
int nsmh_multi_create(uint32_t k, uint32_t n, uint32_t overlap_sketch_thr, const uint64_t *rand_numbers,
                      const int *devices, int ndev, nsmh_multi_handle *out);
int nsmh_multi_destroy(nsmh_multi_handle m);
int nsmh_multi_num_devices(nsmh_multi_handle m, int *ndev);
/* same inputs as nsmh_load_reads_ascii / _dnabitset; shard boundaries: nsmh_multi_shards */
int nsmh_multi_load_reads_ascii(nsmh_multi_handle m, const char *bases, const uint64_t *offsets,
                                uint32_t num_reads);
int nsmh_multi_load_reads_dnabitset(nsmh_multi_handle m, const uint8_t *packed, const uint32_t *lengths,
                                    uint32_t num_reads);
int nsmh_multi_num_reads(nsmh_multi_handle m, uint32_t *num_reads, uint64_t *total_bases);
int nsmh_multi_shards(nsmh_multi_handle m, uint32_t *first_read /* [ndev + 1] */);
int nsmh_multi_sketch(nsmh_multi_handle m);
int nsmh_multi_get_sketches(nsmh_multi_handle m, uint64_t *out /* [num_reads * n], host */);
int nsmh_multi_build(nsmh_multi_handle m);
/* nsmh_query_string on one of the devices; thread-safe, re-entrant */
int nsmh_multi_query_string(nsmh_multi_handle m, const char *s, size_t len, uint32_t *out, size_t cap,
                            size_t *count);
/* nsmh_query_all / _result over all shards: one CSR in global read order */
int nsmh_multi_query_all(nsmh_multi_handle m, int rc, uint64_t *total_ids);
int nsmh_multi_query_all_result(nsmh_multi_handle m, uint64_t *offsets /* [num_reads+1] host */,
                                uint32_t *ids /* [total_ids] host */);

/* ---- online query: ReadFilter::getFilteredReads(const std::string&, std::vector<read_t>&)
 *      (ReadFilter.h:24-25, ReadFilter.cpp:85-97).  Thread-safe, re-entrant. -------------
 * Writes min(count, cap) ids to out and the full count to *count; returns NSMH_ERANGE
 * if count > cap (call again with a larger buffer). */
int nsmh_query_string(nsmh_handle h, const char *s, size_t len, uint32_t *out, size_t cap,
                      size_t *count);
/* Batch of strings in one launch sequence (e.g. forward + reverse-complement window).
 * offsets_out[num_strings+1]; ids_out capacity cap. */
int nsmh_query_strings(nsmh_handle h, const char *bases, const uint64_t *offsets,
                       uint32_t num_strings, uint64_t *offsets_out, uint32_t *ids_out, size_t cap);
/* Batch of ready-made sketches [num_queries][n] (host). */
int nsmh_query_sketches(nsmh_handle h, const uint64_t *sketches, uint32_t num_queries,
                        uint64_t *offsets_out, uint32_t *ids_out, size_t cap);

/* ---- instrumentation ------------------------------------------------------------------ */
/* Device time (CUDA events on the engine's stream) of the last call of each stage, ms. */
typedef struct {
    float h2d_pack_ms;   /* host loaders: H2D + pack, whole call (nsmh_initialize_* / nsmh_load_sketch_*: up to the last chunk's pack,
                          * the sketches of the earlier chunks included) */
    float pack_ms;       /* pack kernels only */
    float sketch_ms;     /* nsmh_sketch */
    float sketch_main_ms;   /* main sketch kernel only */
    float build_ms;      /* nsmh_build */
    float query_ms;      /* nsmh_query_all */
    uint64_t sketch_fixups;   /* (read,hash) pairs the exact fix-up pass had to rescan */
    uint64_t query_pairs;     /* gathered (query,id) pairs in the last bulk query */
    uint32_t kernel_launches; /* kernels launched by this handle since creation */
    float fastq_parse_ms;     /* nsmh_load_fastq*: device time of parse + pack (text already on the device) */
    float fastq_pack_ms;      /* ... of which the gather-pack kernel */
    float fastq_load_ms;      /* nsmh_load_fastq / _file: whole call incl. file read, inflate and H2D (host clock) */
    uint32_t query_heavy;     /* last bulk query: queries whose gathered ids overflowed the on-chip sort buffer */
    uint32_t query_sorted;    /* ... of which went through the global radix sort (the rest: counting-filter tier) */
} nsmh_stats;
int nsmh_get_stats(nsmh_handle h, nsmh_stats *out);
/* The engine's CUDA stream (cudaStream_t as void*), for event timing by the host program. */
int nsmh_stream(nsmh_handle h, void **stream);
int nsmh_synchronize(nsmh_handle h);

/* ---- synthetic nanopore-like reads (bench / test tooling; follows the reference's
 *      util/old_code/createData.py: random genome, ins/del/sub errors, 50% RC) --------- */
typedef struct {
    uint64_t genome_len;    /* 50e6 in the BASELINE configs */
    uint64_t genome_seed;
    uint64_t read_seed;
    float p_ins, p_del, p_sub;   /* 0.03 / 0.03 / 0.04 */
    float p_rc;                  /* 0.5 */
} nsmh_synth_params;
/* Writes reads first_read .. first_read+num_reads of the synthetic set into bases
 * (ASCII, host).  offsets[num_reads+1] are the caller-chosen read boundaries. */
int nsmh_synth_reads_host(const nsmh_synth_params *p, uint64_t first_read, uint32_t num_reads,
                          const uint64_t *offsets, char *bases);
/* Same content generated directly into device memory (device pointers). */
int nsmh_synth_reads_device(int device, const nsmh_synth_params *p, uint64_t first_read,
                            uint32_t num_reads, const uint64_t *d_offsets, char *d_bases);

#ifdef __cplusplus
}
#endif
#endif /* NSMH_H_ */

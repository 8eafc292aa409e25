// Internal declarations of libnsmh.so (not part of the ABI; see include/nsmh.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <functional>
#include <vector>

#include "../../include/nsmh.h"
#include "nsmh_constants.h"

namespace nsmh {

// ---------------------------------------------------------------- errors --
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define NSMH_CK(call)                                                              \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) return ::nsmh::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)
#define NSMH_TRY(expr)                 \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != NSMH_OK) return rc__; \
    } while (0)

// ---------------------------------------------------------------- buffers --
// Grow-only device buffer on a stream-ordered pool.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool borrowed = false;      // storage owned by somebody else (e.g. the peer-visible arena): never grown or freed
    // keep > 0: preserve the first `keep` bytes when the buffer has to grow
    int ensure(size_t bytes, cudaStream_t s, size_t keep = 0);
    void release(cudaStream_t s);
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

// ------------------------------------------------------------- constants --

// A set of reads resident on the device: the reference's ReadData as far as the
// MinHash path needs it (ReadData.h:26-60), 2-bit packed in ONE continuous
// stream: read i occupies global bases [offsets[i], offsets[i+1]).
struct ReadSet {
    uint32_t num_reads = 0;
    uint64_t total_bases = 0;
    uint64_t num_words = 0;     // ceil(total_bases/16)
    DevBuf offsets;             // u64 [num_reads+1]
    DevBuf packed;              // u32 [num_words + kPackPadWords]
    bool external_offsets = false;
    const uint64_t *d_offsets() const { return offsets.as<uint64_t>(); }
    void release(cudaStream_t s) { offsets.release(s); packed.release(s); }
};

// Hash tables: one open-addressing region of (cap+1) slots per hash function.
struct Tables {
    bool built = false;
    uint32_t table_reads = 0;   // rows of the sketch matrix the tables were built from
    uint64_t cap = 0;           // slots per region (2 * rows); slot `cap` holds key ~0
    DevBuf slots;               // Slot [n*(cap+1)]
    DevBuf ids;                 // u32 read ids of the groups with two or more members
};

// Scratch of one bulk / online query (owns its stream so that concurrent host
// threads never share mutable state).
struct QueryWs {
    cudaStream_t stream = nullptr;
    DevBuf qsketch;     // u64 [nq*n]  (string queries)
    DevBuf pval, pcnt;  // u32 [nq*n]  probe results
    DevBuf heavy_list;  // u32 [nq]    queries that overflow the warp buffer
    DevBuf heavy2_list; // u32 [nh]    ... of which the counting-filter tier could not resolve
    DevBuf mid_ids;     // u32         results of the counting-filter tier, completion order
    DevBuf counters;    // u64 [8]     [0..2] count_kernel, [4..6] mid_count_kernel
    DevBuf hc;          // u32 [nh*n+1], global path
    DevBuf hoff;        // u64 [nh*n+1]
    DevBuf hout;        // u32 results of heavy queries
    DevBuf hcnt;        // u32 [nh+1]
    DevBuf hstart;      // u64 [nh+1]
    DevBuf pairs, pairs_alt;   // u64 [T]
    DevBuf flags;       // u8 [T]
    DevBuf qcount;      // u32 [nq+1]
    DevBuf qpos;        // u64 [nq]    start of the query's results in tmp_ids
    DevBuf tmp_ids;     // u32         results in completion order (before the CSR placement)
    DevBuf out_off;     // u64 [nq+1]
    DevBuf out_ids;     // u32 [total]
    DevBuf nsel;        // u64 [2]
    DevBuf cub_tmp;
    DevBuf str_bases;   // staging for query strings
    ReadSet str_reads;
    DevBuf tile_start;  // sketch scratch
    void *h_pinned = nullptr;
    size_t h_pinned_cap = 0;
    // online fast path (online_kernels.cuh): one mapped host block that the kernel reads the strings from and
    // writes the answer to, and a few words of device scratch
    uint8_t *h_online = nullptr;
    size_t h_online_cap = 0;
    DevBuf online_scratch;
    uint64_t last_total = 0;
    uint64_t last_pairs = 0;
    uint32_t last_nq = 0;
    uint32_t last_heavy = 0;    // queries of the last call that overflowed the warp buffer
    uint32_t last_sorted = 0;   // ... of which went through the global sort
    uint32_t launches = 0;
    uint64_t seen_build_epoch = 0;  // the last build this workspace's stream has been ordered behind
};

struct MgState;

} // namespace nsmh

// nsmh_sketch_build: the exact fix-up of the sketch (entries no k-mer matched on the filter prefix) runs on a second
// stream beside the table insert; its values wait here until table_insert_list_kernel stores and inserts them.
struct SketchDeferred {
    nsmh::DevBuf buf;                         // vals u64 [entries] | list u32 [entries]
    cudaStream_t aux = nullptr;         // the context's copy stream
    cudaEvent_t filtered = nullptr, fixed = nullptr;
    uint64_t *vals = nullptr;
    uint32_t *list = nullptr;           // entry index = read * n + hash
    unsigned int *count = nullptr;
    uint64_t *sk = nullptr;             // the sketch matrix the entries belong to
    uint32_t n = 0;
    bool pending = false;
    std::function<int(int)> launch;     // queues the two fix-up kernels on aux, `blocks per SM` each (0: the build's default);
                                        // called once the kernel they run beside is queued
};

struct nsmh_ctx {
    int device = 0;
    uint32_t k = 0, n = 0, thr = 0;
    std::vector<uint64_t> rand;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev[10] = {};
    bool build_timed = false;
    uint64_t build_epoch = 0;       // builds queued so far; ev[7] marks the end of the last one on `stream`
    SketchDeferred defer;           // nsmh_sketch_build
    bool load_timed = false;        // ev[0], ev[1] bracket a pipelined load that returned before the device finished
    int sketch_mode = 0;
    int num_sms = 148;

    nsmh::DevBuf d_rand;        // u64 [n]
    nsmh::DevBuf d_ftab_first;  // u8 [kFilterTabSize]     first hash of the chain
    nsmh::DevBuf d_ftab_next;   // u8 [(kFilterMaxBits+1)*n] next hash in chain, 0xFF = end
    nsmh::DevBuf d_ftab_hit3;   // u8 [kFilter3TabSize]    3 hit bits per (b+4)-bit window
    int lambda_log2 = nsmh::kFilterLambdaLog2;
    uint32_t tile_words = nsmh::kTileWords;

    nsmh::ReadSet reads;
    bool reads_loaded = false;
    nsmh::DevBuf sketches;      // u64 [num_reads*n]
    bool sketched = false;
    nsmh::DevBuf tile_start;    // u32 [num_reads+1]
    nsmh::DevBuf read_flags;    // u8 [num_reads]  NSMH_FLAG_* (prefilter.cu), valid when flags_valid
    bool flags_valid = false;
    nsmh::DevBuf counters;      // u64 [8] device counters (fix-ups ...)

    const uint64_t *table_sketches = nullptr;   // rows the tables are built from
    uint32_t table_reads = 0, id_base = 0;
    nsmh::Tables tables;
    cudaEvent_t ev_order = nullptr, ev_cleared = nullptr;
    uint32_t precleared_rows = 0;   // tables already cleared for this many rows (preclear_tables)
    void *precleared_ptr = nullptr;
    nsmh::DevBuf build_multi;   // build scratch: members / groups of keys shared by several reads
    nsmh::DevBuf build_tmp;

    nsmh::QueryWs bulk;         // nsmh_query_all workspace (uses ctx stream)
    bool bulk_valid = false;
    std::mutex pool_mu;
    std::vector<nsmh::QueryWs *> pool;   // idle online-query workspaces

    nsmh_stats stats = {};
    uint32_t launches = 0;

    bool owns_stream = true;    // false: `stream` belongs to the parent context (multi-GPU sub-context)
    nsmh::MgState *mg = nullptr;   // multigpu.cu: tables partitioned by hash function across ranks
    nsmh_ctx *mg_sub() const;      // the context of the tables this rank owns (nullptr before nsmh_mg_connect)
    uint32_t mg_total_rows() const;
};

namespace nsmh {

// ---- cub_ops.cu (library primitives: scan / radix sort / select) -----------
cudaError_t cub_exclusive_sum_u32(void *tmp, size_t &tmp_bytes, const uint32_t *in, uint32_t *out,
                                  size_t n, cudaStream_t s);
cudaError_t cub_exclusive_sum_u32_to_u64(void *tmp, size_t &tmp_bytes, const uint32_t *in,
                                         uint64_t *out, size_t n, cudaStream_t s);
cudaError_t cub_sort_keys_u64(void *tmp, size_t &tmp_bytes, uint64_t *keys, uint64_t *alt,
                              size_t n, int begin_bit, int end_bit, bool &result_in_alt,
                              cudaStream_t s);
cudaError_t cub_select_low32_flagged(void *tmp, size_t &tmp_bytes, const uint64_t *in,
                                     const uint8_t *flags, uint32_t *out, uint64_t *num_selected,
                                     size_t n, cudaStream_t s);

// ---- pack.cu -----------------------------------------------------------------
int pack_ascii(ReadSet &rs, const char *d_bases, uint64_t first_base, uint64_t num_bases,
               cudaStream_t s, uint32_t *launches);
int alloc_packed(ReadSet &rs, uint64_t total_bases, cudaStream_t s);
int pack_reverse_complement(const ReadSet &src, ReadSet &dst, cudaStream_t s, uint32_t *launches);
int pack_from_dnabitset(ReadSet &rs, const uint8_t *d_src, const uint64_t *d_src_byte_off,
                        cudaStream_t s, uint32_t *launches, uint64_t w_begin = 0, uint64_t w_end = ~0ULL);

// ---- fastq.cu ----------------------------------------------------------------
int parse_fastq_device(nsmh_ctx *c, const uint8_t *d_text, uint64_t bytes, uint64_t safe_bytes, int last_byte);
int unpack_ascii_device(nsmh_ctx *c, uint64_t b0, uint64_t nb, uint8_t *d_out, cudaStream_t s);

// ---- sketch.cu ---------------------------------------------------------------
int build_filter_tables(nsmh_ctx *c);
int sketch_reads(nsmh_ctx *c, const ReadSet &rs, uint64_t *d_sketches, DevBuf &tile_start,
                 DevBuf &cub_tmp, int mode, cudaStream_t s, uint32_t *launches, cudaEvent_t ev0,
                 cudaEvent_t ev1, uint32_t r0 = 0, uint32_t r1 = ~0u, SketchDeferred *defer = nullptr);

// ---- table.cu ----------------------------------------------------------------
int build_tables(nsmh_ctx *c, SketchDeferred *defer = nullptr);
// multigpu.cu: nsmh_mg_run; defer: see nsmh_mg_sketch_run
int mg_run_impl(nsmh_ctx *c, uint64_t *total_ids, SketchDeferred *defer);
int preclear_tables(nsmh_ctx *c, uint32_t rows, bool in_order = false);
int table_num_keys(nsmh_ctx *c, uint32_t j, uint32_t *out);

// ---- query.cu ----------------------------------------------------------------
// one fused kernel per online query (online_kernels.cuh); 1 = the general path must take over
int online_query(nsmh_ctx *c, QueryWs &ws, const char *bases, const uint64_t *offsets, uint32_t nq, uint64_t *offsets_out,
                 uint32_t *ids_out, size_t cap);
int query_sketches_device(nsmh_ctx *c, QueryWs &ws, const uint64_t *d_qsketch, uint32_t nq,
                          cudaStream_t s);
int probe_lists_device(nsmh_ctx *c, QueryWs &ws, const uint64_t *d_qsketch, uint32_t nq, cudaStream_t s);
int count_lists_device(nsmh_ctx *c, QueryWs &ws, uint32_t nq, uint32_t parts, const uint64_t *const *d_offsets,
                       const uint32_t *const *d_ids, cudaStream_t s);

// ---- prefilter.cu --------------------------------------------------------------
int compute_read_flags(nsmh_ctx *c);
int drop_flagged_candidates(nsmh_ctx *c, QueryWs &ws, uint32_t drop, cudaStream_t s);

// ---- multi-GPU over peer memory (query.cu kernels, multigpu.cu orchestration) ----
int probe_to_peers_device(nsmh_ctx *sub, const uint64_t *d_qsketch, uint32_t nq, const PeerDst &dst,
                          cudaStream_t s, uint32_t *launches);
int count_peer_lists_device(nsmh_ctx *c, QueryWs &ws, const PeerLists &src, uint32_t nq, cudaStream_t s);
void mg_destroy(nsmh_ctx *c);

} // namespace nsmh

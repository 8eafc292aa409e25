// Multi-GPU read-overlap stage, one process per GPU, exchanges over NVLink peer memory.
//
// The reference has no distributed path: MinHashReadFilter::initialize is one OpenMP loop over
// the reads (src/ReadFilter.cpp:31-44) followed by one over the n tables (:163-165), and
// getFilteredReads probes all n tables (:65-83).  Sharding the reads makes sketching local; the
// tables need the sketches of all reads.  Here the tables are partitioned by hash function:
// rank g owns a contiguous block of the n hash functions for ALL reads, so every stage handles
// reads_per_rank * n items per rank whatever the world size (weak scaling), and the data moves
// exactly twice, both times as plain stores into a peer's memory issued by the kernel that
// produced the data:
//
//   1. mg_scatter_columns_kernel   local sketch rows [rows][n] -> column blocks in the owners'
//                                  arenas, M_o[global row][owned column]          (8 B / item)
//   2. mg_barrier_kernel           flags in peer memory (release/acquire at system scope)
//   3. table build (table.cu)      owner builds its tables over all rows of M_g
//   4. probe_to_peers_kernel       owner probes its tables for all rows and stores
//      (query.cu)                  {id | group start, group size} into the arena of the rank
//                                  that owns the read                             (8 B / item)
//   5. mg_barrier_kernel
//   6. count_kernel<PeerSrc>       every rank thresholds its own reads; ids of groups with
//      (query.cu)                  two or more members are read from the owner's arena
//
// Everything runs on the context's stream without host round trips between the stages.  The
// arena is one cudaMalloc block per rank, exported with cudaIpcGetMemHandle; the token the
// ranks exchange carries the handle and the arena's layout.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "nsmh_internal.cuh"
#include "multigpu_kernels.cuh"

namespace nsmh {

constexpr uint64_t kMgMagic = 0x4e534d48504d4731ULL;   // "NSMHPMG1"

struct MgToken {
    cudaIpcMemHandle_t handle;       // 64 bytes
    uint64_t magic;
    uint64_t off_m, off_pr, off_ids, off_inbox, off_flags, arena_bytes;
    uint32_t inbox_cap, pad0;
    uint32_t rank, world, n_total, col0, ncols, rows, total_rows;
    int32_t device;
};
static_assert(sizeof(MgToken) <= NSMH_MG_TOKEN_BYTES, "token does not fit");

struct MgState {
    uint32_t rank = 0, world = 0, n_total = 0;
    uint32_t rows[kMgMaxRanks] = {}, row_end[kMgMaxRanks] = {}, col_end[kMgMaxRanks] = {};
    uint32_t total_rows = 0, col0 = 0, ncols = 0;
    uint8_t *arena = nullptr;
    MgToken self = {};
    MgToken peers[kMgMaxRanks] = {};
    uint8_t *peer_base[kMgMaxRanks] = {};
    bool connected = false;
    nsmh_ctx *sub = nullptr;          // the owned tables: ncols hash functions over total_rows reads
    uint32_t epoch = 0;
    uint64_t timeout_ns = 20ull * 1000 * 1000 * 1000;
    cudaEvent_t ev[7] = {};
    float stage_ms[6] = {};
};

struct BarrierArgs {
    uint32_t *flags[kMgMaxRanks];     // flags block of every rank: [2][kMgMaxRanks] epochs, [1] error, [kMgMaxRanks] inbox cursors
    uint32_t world, rank;
    uint64_t timeout_ns;
};

__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Thread t tells rank t "rank `rank` has finished everything before this kernel" and waits for
// the same message from rank t.  The stores of the preceding kernels (same stream) happen
// before the release store; the kernels after this one start after every acquire load saw the
// peers' epochs, so they see the peers' data.  A peer that never arrives sets the error flag
// after the timeout instead of hanging the device.
__global__ void mg_barrier_kernel(BarrierArgs a, uint32_t epoch, uint32_t which) {
    const uint32_t t = threadIdx.x;
    if (t >= a.world || t == a.rank) return;
    __threadfence_system();
    uint32_t *remote = a.flags[t] + which * kMgMaxRanks + a.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    const uint32_t *mine = a.flags[a.rank] + which * kMgMaxRanks + t;
    const uint64_t t0 = global_ns();
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        if (global_ns() - t0 > a.timeout_ns) {
            a.flags[a.rank][2 * kMgMaxRanks] = 1u;
            break;
        }
        __nanosleep(200);
    }
}

void mg_destroy(nsmh_ctx *c) {
    MgState *m = c->mg;
    if (!m) return;
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (uint32_t r = 0; r < m->world; ++r)
        if (r != m->rank && m->peer_base[r]) cudaIpcCloseMemHandle(m->peer_base[r]);
    if (m->sub) {
        m->sub->tables.ids.p = nullptr;      // borrowed from the arena
        m->sub->tables.ids.cap = 0;
        m->sub->tables.ids.borrowed = false;
        nsmh_destroy(m->sub);
    }
    for (auto &e : m->ev) if (e) cudaEventDestroy(e);
    if (m->arena) cudaFree(m->arena);
    cudaGetLastError();
    delete m;
    c->mg = nullptr;
}

} // namespace nsmh

nsmh_ctx *nsmh_ctx::mg_sub() const { return mg && mg->connected ? mg->sub : nullptr; }
uint32_t nsmh_ctx::mg_total_rows() const { return mg ? mg->total_rows : 0u; }

using namespace nsmh;

extern "C" {

int nsmh_mg_init(nsmh_handle c, uint32_t rank, uint32_t world, const uint32_t *rows_per_rank, void *token_out) {
    if (!c) return fail(NSMH_EINVAL, "null handle");
    if (!rows_per_rank || !token_out) return fail(NSMH_EINVAL, "mg_init: null argument");
    if (world < 1 || world > (uint32_t)kMgMaxRanks || rank >= world)
        return fail(NSMH_EINVAL, "mg_init: world must be in 1..16 and rank < world");
    uint32_t split[kMgMaxRanks];
    if (!mg_split_columns(c->n, world, split)) return fail(NSMH_EINVAL, "mg_init: more ranks than hash functions");
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(c->device) != cudaSuccess) return fail(NSMH_ECUDA, "cudaSetDevice failed");
    mg_destroy(c);
    MgState *m = new (std::nothrow) MgState();
    if (!m) return fail(NSMH_ENOMEM, "mg_init: out of host memory");
    c->mg = m;
    m->rank = rank;
    m->world = world;
    m->n_total = c->n;
    uint64_t total = 0;
    for (uint32_t r = 0; r < world; ++r) {
        m->rows[r] = rows_per_rank[r];
        total += rows_per_rank[r];
        m->row_end[r] = (uint32_t)total;
        m->col_end[r] = split[r];
    }
    int rc = NSMH_OK;
    do {
        // read ids travel in 31 bits when a group of two is sent inside a probe result (kPairFlag)
        if (total >= (1ULL << 30)) { rc = fail(NSMH_EINVAL, "mg_init: more than 2^30 reads"); break; }
        m->total_rows = (uint32_t)total;
        m->col0 = rank ? m->col_end[rank - 1] : 0u;
        m->ncols = m->col_end[rank] - m->col0;
        const char *e = getenv("NSMH_MG_TIMEOUT_MS");
        if (e && *e && atoll(e) > 0) m->timeout_ns = (uint64_t)atoll(e) * 1000000ull;

        MgToken &t = m->self;
        memset(&t, 0, sizeof t);
        const size_t items = std::max<size_t>((size_t)m->total_rows * m->ncols, 1);
        uint32_t max_cols = 0;
        for (uint32_t r = 0; r < world; ++r) max_cols = std::max(max_cols, m->col_end[r] - (r ? m->col_end[r - 1] : 0u));
        const char *ic = getenv("NSMH_MG_INBOX_CAP");        // tests: force the overflow path
        uint32_t pcol[kMgMaxRanks];
        const uint32_t pr_cols = mg_padded_offsets(m->col_end, world, pcol);
        const MgLayout lay = mg_layout(m->total_rows, m->ncols, m->rows[rank], pr_cols, max_cols, world,
                                       ic && *ic ? atoll(ic) : 0);
        t.off_m = lay.off_m;
        t.off_pr = lay.off_pr;
        t.off_ids = lay.off_ids;
        t.off_inbox = lay.off_inbox;
        t.off_flags = lay.off_flags;
        t.inbox_cap = lay.inbox_cap;
        t.arena_bytes = lay.arena_bytes;
        const size_t off = lay.arena_bytes;
        cudaError_t ce = cudaMalloc(reinterpret_cast<void **>(&m->arena), off);
        if (ce != cudaSuccess) { rc = cuda_fail(ce, "cudaMalloc(arena)", __FILE__, __LINE__); break; }
        if ((ce = cudaMemset(m->arena + t.off_flags, 0, (3 * kMgMaxRanks + 8) * sizeof(uint32_t))) != cudaSuccess) { rc = cuda_fail(ce, "memset", __FILE__, __LINE__); break; }
        if ((ce = cudaIpcGetMemHandle(&t.handle, m->arena)) != cudaSuccess) { rc = cuda_fail(ce, "cudaIpcGetMemHandle", __FILE__, __LINE__); break; }
        t.magic = kMgMagic;
        t.rank = rank;
        t.world = world;
        t.n_total = m->n_total;
        t.col0 = m->col0;
        t.ncols = m->ncols;
        t.rows = m->rows[rank];
        t.total_rows = m->total_rows;
        t.device = c->device;
        for (auto &ev : m->ev)
            if ((ce = cudaEventCreate(&ev)) != cudaSuccess) { rc = cuda_fail(ce, "event", __FILE__, __LINE__); break; }
        if (rc) break;

        // the owned tables live in a sub-context that shares this context's stream
        if ((rc = nsmh_create(c->k, m->ncols, c->thr, c->rand.data() + m->col0, c->device, &m->sub))) break;
        cudaStreamSynchronize(m->sub->stream);
        cudaStreamDestroy(m->sub->stream);
        m->sub->stream = c->stream;
        m->sub->bulk.stream = c->stream;
        m->sub->owns_stream = false;
        m->sub->tables.ids.p = m->arena + t.off_ids;
        m->sub->tables.ids.cap = items * sizeof(uint32_t);
        m->sub->tables.ids.borrowed = true;
    } while (0);
    if (rc) {
        std::string keep = nsmh_last_error();
        mg_destroy(c);
        set_error(keep);
    } else {
        memset(token_out, 0, NSMH_MG_TOKEN_BYTES);
        memcpy(token_out, &m->self, sizeof(MgToken));
    }
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

int nsmh_mg_connect(nsmh_handle c, const void *tokens) {
    if (!c) return fail(NSMH_EINVAL, "null handle");
    MgState *m = c->mg;
    if (!m) return fail(NSMH_ESTATE, "mg_connect: call nsmh_mg_init first");
    if (!tokens) return fail(NSMH_EINVAL, "mg_connect: null tokens");
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(c->device) != cudaSuccess) return fail(NSMH_ECUDA, "cudaSetDevice failed");
    int rc = NSMH_OK;
    for (uint32_t r = 0; r < m->world && !rc; ++r) {
        MgToken &t = m->peers[r];
        memcpy(&t, static_cast<const uint8_t *>(tokens) + (size_t)r * NSMH_MG_TOKEN_BYTES, sizeof(MgToken));
        const uint32_t cb = r ? m->col_end[r - 1] : 0u;
        if (t.magic != kMgMagic || t.rank != r || t.world != m->world || t.n_total != m->n_total ||
            t.total_rows != m->total_rows || t.rows != m->rows[r] || t.col0 != cb || t.ncols != m->col_end[r] - cb) {
            rc = fail(NSMH_EINVAL, "mg_connect: token " + std::to_string(r) + " does not match this rank's configuration");
            break;
        }
        if (r == m->rank) {
            m->peer_base[r] = m->arena;
            continue;
        }
        if (m->peer_base[r]) { cudaIpcCloseMemHandle(m->peer_base[r]); m->peer_base[r] = nullptr; }
        void *p = nullptr;
        cudaError_t ce = cudaIpcOpenMemHandle(&p, t.handle, cudaIpcMemLazyEnablePeerAccess);
        if (ce != cudaSuccess) { rc = cuda_fail(ce, "cudaIpcOpenMemHandle", __FILE__, __LINE__); break; }
        m->peer_base[r] = static_cast<uint8_t *>(p);
    }
    m->connected = rc == NSMH_OK;
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

int nsmh_mg_run(nsmh_handle c, uint64_t *total_ids) {
    if (!c) return fail(NSMH_EINVAL, "null handle");
    return nsmh::mg_run_impl(c, total_ids, nullptr);
}

} // extern "C"

// defer (nsmh_mg_sketch_run): the exact fix-up pass of the sketch has not run yet - it is queued on the second
// stream once the column scatter is, and what it produces is sent after it (mg_scatter_list_kernel).
int nsmh::mg_run_impl(nsmh_ctx *c, uint64_t *total_ids, SketchDeferred *defer) {
    MgState *m = c->mg;
    if (!m || !m->connected) return fail(NSMH_ESTATE, "mg_run: call nsmh_mg_init and nsmh_mg_connect first");
    if (!c->sketched) return fail(NSMH_ESTATE, "mg_run: call nsmh_sketch first");
    if (c->reads.num_reads != m->rows[m->rank])
        return fail(NSMH_EINVAL, "mg_run: the loaded batch does not have rows_per_rank[rank] reads");
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(c->device) != cudaSuccess) return fail(NSMH_ECUDA, "cudaSetDevice failed");
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};
    cudaStream_t s = c->stream;
    nsmh_ctx *sub = m->sub;
    const uint32_t rows = m->rows[m->rank];
    const uint32_t epoch = ++m->epoch;
    c->bulk_valid = false;
    c->bulk.stream = s;

    BarrierArgs ba;
    ScatterArgs sa;
    PeerDst pd;
    PeerLists pl;
    for (uint32_t r = 0; r < (uint32_t)kMgMaxRanks; ++r) {
        const bool in = r < m->world;
        uint8_t *base = in ? m->peer_base[r] : nullptr;
        ba.flags[r] = in ? reinterpret_cast<uint32_t *>(base + m->peers[r].off_flags) : nullptr;
        sa.m[r] = in ? reinterpret_cast<uint64_t *>(base + m->peers[r].off_m) : nullptr;
        sa.col_end[r] = in ? m->col_end[r] : 0u;
        pd.pr[r] = in ? reinterpret_cast<uint64_t *>(base + m->peers[r].off_pr) : nullptr;
        pd.inbox[r] = in ? reinterpret_cast<uint32_t *>(base + m->peers[r].off_inbox) + (size_t)m->rank * m->peers[r].inbox_cap : nullptr;
        pd.inbox_cap[r] = in ? m->peers[r].inbox_cap : 0u;
        pd.row_end[r] = in ? m->row_end[r] : 0u;
        pl.ids[r] = in ? reinterpret_cast<const uint32_t *>(base + m->peers[r].off_ids) : nullptr;
        pl.col_end[r] = in ? m->col_end[r] : 0u;
    }
    if (const char *dbg = getenv("NSMH_MG_DEBUG_LOCAL_STORES")) {
        // timing experiment only (results are wrong): the probe results of ALL reads go to this rank's own arena
        // 1: result tiles and inbox ids stay local; 2: only the inbox ids do; 3: only the result tiles do
        const int mode = atoi(dbg);
        if (mode != 0)
            for (uint32_t r = 0; r < m->world; ++r) {
                if (mode != 2) pd.pr[r] = reinterpret_cast<uint64_t *>(m->arena + m->self.off_pr);
                if (mode != 3) pd.inbox[r] = reinterpret_cast<uint32_t *>(m->arena + m->self.off_inbox) + (size_t)m->rank * m->self.inbox_cap;
            }
    }
    ba.world = sa.world = pd.world = pl.world = m->world;
    ba.rank = m->rank;
    ba.timeout_ns = m->timeout_ns;
    sa.row0 = m->rank ? m->row_end[m->rank - 1] : 0u;
    pd.col0 = m->col0;
    pd.ncols = m->ncols;
    {
        uint32_t pcol[kMgMaxRanks];
        mg_padded_offsets(m->col_end, m->world, pcol);
        pd.pcol0 = pcol[m->rank];
    }
    pd.chunk0 = m->total_rows ? std::min<uint32_t>(sa.row0, m->total_rows - 1) / kProbeRows : 0u;
    pl.pr = reinterpret_cast<const uint64_t *>(m->arena + m->self.off_pr);
    pl.n = m->n_total;
    pl.rows = rows;
    pl.inbox = reinterpret_cast<const uint32_t *>(m->arena + m->self.off_inbox);
    pl.inbox_cap = m->self.inbox_cap;
    uint32_t *cursors = reinterpret_cast<uint32_t *>(m->arena + m->self.off_flags) + 2 * kMgMaxRanks + 8;
    pd.cursor = cursors;
    NSMH_CK(cudaMemsetAsync(cursors, 0, kMgMaxRanks * sizeof(uint32_t), s));
    uint64_t *M = reinterpret_cast<uint64_t *>(m->arena + m->self.off_m);

    NSMH_CK(cudaEventRecord(m->ev[0], s));
    if (rows) {
        const size_t smem = (size_t)kScatterRows * c->n * sizeof(uint64_t);
        if (smem > 48 * 1024) NSMH_CK(cudaFuncSetAttribute(mg_scatter_columns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // NSMH_MG_SCATTER_BLOCKS / NSMH_MG_FIXUP_BLOCKS: blocks per SM of the scatter and of the fix-up pass that runs
        // beside it (nsmh_mg_sketch_run).  Measured on 8 GPUs: whatever the split, the two hardly overlap
        // (profiles/r2_mg_sketch_run_overlap_s34.txt), so both keep their usual 8.
        const char *sbe = getenv("NSMH_MG_SCATTER_BLOCKS"), *fbe = getenv("NSMH_MG_FIXUP_BLOCKS");
        const int scatter_per_sm = sbe && *sbe && atoi(sbe) > 0 ? atoi(sbe) : 8;
        const int fixup_per_sm = fbe && *fbe && atoi(fbe) > 0 ? atoi(fbe) : 8;
        const int blocks = (int)std::min<uint64_t>(((uint64_t)rows + kScatterRows - 1) / kScatterRows, (uint64_t)c->num_sms * scatter_per_sm);
        mg_scatter_columns_kernel<<<blocks, 256, smem, s>>>(c->sketches.as<uint64_t>(), rows, c->n, sa);
        ++c->launches;
        NSMH_CK(cudaGetLastError());
        if (defer && defer->pending) {
            NSMH_TRY(defer->launch(fixup_per_sm));
            NSMH_CK(cudaStreamWaitEvent(s, defer->fixed, 0));
            mg_scatter_list_kernel<<<c->num_sms * 2, 256, 0, s>>>(c->sketches.as<uint64_t>(), c->n, defer->list, defer->count,
                                                                 defer->vals, sa);
            ++c->launches;
            NSMH_CK(cudaGetLastError());
            defer->pending = false;
        }
    }
    NSMH_CK(cudaEventRecord(m->ev[1], s));
    mg_barrier_kernel<<<1, 32, 0, s>>>(ba, epoch, 0);
    ++c->launches;
    NSMH_CK(cudaGetLastError());
    NSMH_CK(cudaEventRecord(m->ev[2], s));
    sub->table_sketches = M;
    sub->table_reads = m->total_rows;
    sub->id_base = 0;
    NSMH_TRY(build_tables(sub));
    NSMH_CK(cudaEventRecord(m->ev[3], s));
    NSMH_TRY(probe_to_peers_device(sub, M, m->total_rows, pd, s, &c->launches));
    NSMH_CK(cudaEventRecord(m->ev[4], s));
    mg_barrier_kernel<<<1, 32, 0, s>>>(ba, epoch, 1);
    ++c->launches;
    NSMH_CK(cudaGetLastError());
    NSMH_CK(cudaEventRecord(m->ev[5], s));
    NSMH_TRY(count_peer_lists_device(c, c->bulk, pl, rows, s));
    NSMH_CK(cudaEventRecord(m->ev[6], s));
    uint32_t err = 0;
    NSMH_CK(cudaMemcpyAsync(&err, m->arena + m->self.off_flags + 2 * kMgMaxRanks * sizeof(uint32_t), sizeof err,
                            cudaMemcpyDeviceToHost, s));
    NSMH_CK(cudaStreamSynchronize(s));
    c->launches += sub->launches;
    sub->launches = 0;
    for (int i = 0; i < 6; ++i) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, m->ev[i], m->ev[i + 1]) != cudaSuccess) { cudaGetLastError(); ms = 0; }
        m->stage_ms[i] = ms;
    }
    if (err) {
        cudaMemsetAsync(m->arena + m->self.off_flags + 2 * kMgMaxRanks * sizeof(uint32_t), 0, sizeof err, s);
        return fail(NSMH_ECUDA, "mg_run: a peer rank did not reach the barrier before the timeout");
    }
    c->bulk_valid = true;
    c->stats.query_pairs = c->bulk.last_pairs;
    c->stats.query_heavy = c->bulk.last_heavy;
    c->stats.query_sorted = c->bulk.last_sorted;
    c->stats.build_ms = m->stage_ms[2];
    c->stats.query_ms = m->stage_ms[3] + m->stage_ms[5];
    if (total_ids) *total_ids = c->bulk.last_total;
    return NSMH_OK;
}

extern "C" {

int nsmh_mg_stage_ms(nsmh_handle c, float *out) {
    if (!c || !out) return fail(NSMH_EINVAL, "mg_stage_ms: null argument");
    if (!c->mg) return fail(NSMH_ESTATE, "mg_stage_ms: call nsmh_mg_init first");
    for (int i = 0; i < 6; ++i) out[i] = c->mg->stage_ms[i];
    return NSMH_OK;
}

int nsmh_mg_shutdown(nsmh_handle c) {
    if (!c) return fail(NSMH_EINVAL, "null handle");
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(c->device) != cudaSuccess) return fail(NSMH_ECUDA, "cudaSetDevice failed");
    mg_destroy(c);
    if (prev >= 0) cudaSetDevice(prev);
    return NSMH_OK;
}

} // extern "C"

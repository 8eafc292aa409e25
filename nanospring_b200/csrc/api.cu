// C ABI of libnsmh.so (include/nsmh.h): context management and the host-side
// orchestration of the kernels in pack.cu / sketch.cu / table.cu / query.cu.
// Mirrors the call sequence of the reference:
//   Compressor.cpp:69-76   rF.k/n/overlapSketchThreshold = ...; rF.initialize(rD)
//   ReadFilter.cpp:11-47   initialize(): sketch all reads, populateHashTables()
//   Consensus.cpp:189      rF->getFilteredReads(string, results)  (concurrent callers)
#include <sched.h>
#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstring>

#include "nsmh_internal.cuh"

namespace nsmh {

static thread_local std::string g_err;

void set_error(const std::string &msg) { g_err = msg; }
int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file,
             line, what);
    g_err = buf;
    cudaGetLastError();   // clear sticky-free errors
    return e == cudaErrorMemoryAllocation ? NSMH_ENOMEM : NSMH_ECUDA;
}

int DevBuf::ensure(size_t bytes, cudaStream_t s, size_t keep) {
    if (bytes <= cap && p) return NSMH_OK;
    if (borrowed) return fail(NSMH_ENOMEM, "internal: a borrowed device buffer is too small");
    size_t want = bytes + bytes / 8 + 256;
    void *np = nullptr;
    NSMH_CK(cudaMallocAsync(&np, want, s));
    if (p) {
        if (keep) NSMH_CK(cudaMemcpyAsync(np, p, keep < cap ? keep : cap, cudaMemcpyDeviceToDevice, s));
        NSMH_CK(cudaFreeAsync(p, s));
    }
    p = np;
    cap = want;
    return NSMH_OK;
}

void DevBuf::release(cudaStream_t s) {
    if (p && !borrowed) cudaFreeAsync(p, s);
    borrowed = false;
    p = nullptr;
    cap = 0;
}

static void free_ws(QueryWs &ws, cudaStream_t s) {
    DevBuf *bufs[] = {&ws.qsketch, &ws.pval, &ws.pcnt, &ws.heavy_list, &ws.heavy2_list, &ws.mid_ids, &ws.counters, &ws.hc, &ws.hoff, &ws.hout, &ws.hcnt, &ws.hstart, &ws.pairs, &ws.pairs_alt, &ws.flags,
                      &ws.qcount, &ws.qpos, &ws.tmp_ids, &ws.out_off, &ws.out_ids, &ws.nsel, &ws.cub_tmp, &ws.str_bases,
                      &ws.tile_start, &ws.online_scratch};
    for (DevBuf *b : bufs) b->release(s);
    ws.str_reads.release(s);
    if (ws.h_pinned) cudaFreeHost(ws.h_pinned);
    ws.h_pinned = nullptr;
    ws.h_pinned_cap = 0;
    if (ws.h_online) cudaFreeHost(ws.h_online);
    ws.h_online = nullptr;
    ws.h_online_cap = 0;
}

static int ensure_pinned(QueryWs &ws, size_t bytes) {
    if (bytes <= ws.h_pinned_cap) return NSMH_OK;
    if (ws.h_pinned) cudaFreeHost(ws.h_pinned);
    ws.h_pinned = nullptr;
    ws.h_pinned_cap = 0;
    size_t want = bytes + bytes / 4 + 4096;
    NSMH_CK(cudaMallocHost(&ws.h_pinned, want));
    ws.h_pinned_cap = want;
    return NSMH_OK;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static float elapsed(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); ms = 0; }
    return ms;
}

// MT19937-64 (Matsumoto & Nishimura; the engine behind std::mt19937_64).
static void mt19937_64_stream(uint32_t seed, uint32_t n, uint64_t *out) {
    const int NN = 312, MM = 156;
    std::vector<uint64_t> mt(NN);
    mt[0] = seed;
    for (int i = 1; i < NN; ++i) mt[i] = 6364136223846793005ULL * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i;
    int idx = NN;
    for (uint32_t o = 0; o < n; ++o) {
        if (idx >= NN) {
            for (int i = 0; i < NN; ++i) {
                uint64_t x = (mt[i] & 0xFFFFFFFF80000000ULL) | (mt[(i + 1) % NN] & 0x7FFFFFFFULL);
                uint64_t xa = x >> 1;
                if (x & 1) xa ^= 0xB5026F5AA96619E9ULL;
                mt[i] = mt[(i + MM) % NN] ^ xa;
            }
            idx = 0;
        }
        uint64_t y = mt[idx++];
        y ^= (y >> 29) & 0x5555555555555555ULL;
        y ^= (y << 17) & 0x71D67FFFEDA60000ULL;
        y ^= (y << 37) & 0xFFF7EEE000000000ULL;
        y ^= y >> 43;
        out[o] = y;
    }
}

} // namespace nsmh

using namespace nsmh;

#define CTX_GUARD(h)                                                   \
    if (!(h)) return fail(NSMH_EINVAL, "null handle");                 \
    DeviceGuard guard__((h)->device);                                  \
    if (!guard__.ok) return fail(NSMH_ECUDA, "cudaSetDevice failed")

extern "C" {

const char *nsmh_last_error(void) { return g_err.c_str(); }
const char *nsmh_version(void) { return "nsmh 0.1 (sm_100a)"; }

int nsmh_rand_from_seed(uint32_t seed, uint32_t n, uint64_t *out) {
    if (!out && n) return fail(NSMH_EINVAL, "rand_from_seed: null output");
    mt19937_64_stream(seed, n, out);
    return NSMH_OK;
}

int nsmh_host_alloc(size_t bytes, void **out) {
    if (!out) return fail(NSMH_EINVAL, "host_alloc: null output");
    NSMH_CK(cudaMallocHost(out, bytes ? bytes : 1));
    return NSMH_OK;
}

// ---- host memory next to a GPU -------------------------------------------------------------------
// On a two-socket box half of the GPUs hang off the other socket: a pinned buffer on the wrong node sends
// every H2D byte across the socket interconnect first, and with eight ranks copying at once that link, not
// PCIe, sets the end-to-end rate.  The node of a device comes from sysfs (its PCI function's numa_node).
static int device_numa_cpus(int device, cpu_set_t *set) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char *q = bus; *q; ++q) *q = (char)tolower((unsigned char)*q);
    char path[160];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *fh = fopen(path, "r");
    if (!fh) return -1;
    int node = -1;
    if (fscanf(fh, "%d", &node) != 1) node = -1;
    fclose(fh);
    if (node < 0) return -1;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    fh = fopen(path, "r");
    if (!fh) return -1;
    CPU_ZERO(set);
    int a = 0, b = 0, count = 0;
    while (fscanf(fh, "%d", &a) == 1) {          // "0-31,64-95"
        b = a;
        int ch = fgetc(fh);
        if (ch == '-') {
            if (fscanf(fh, "%d", &b) != 1) break;
            ch = fgetc(fh);
        }
        for (int cpu = a; cpu <= b && cpu < CPU_SETSIZE; ++cpu) { CPU_SET(cpu, set); ++count; }
        if (ch != ',') break;
    }
    fclose(fh);
    if (!count) return -1;
    // keep only what this process may run on (containers hand out a subset)
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof allowed, &allowed) == 0) {
        cpu_set_t both;
        CPU_AND(&both, set, &allowed);
        if (CPU_COUNT(&both) == 0) return -1;
        *set = both;
    }
    return node;
}

int nsmh_bind_thread_near(int device, int *numa_node) {
    cpu_set_t set;
    const int node = device_numa_cpus(device, &set);
    if (numa_node) *numa_node = node;
    if (node < 0) return NSMH_OK;                // single node / no topology information: nothing to do
    if (sched_setaffinity(0, sizeof set, &set) != 0) { if (numa_node) *numa_node = -1; }
    return NSMH_OK;
}

int nsmh_host_alloc_near(int device, size_t bytes, void **out) {
    if (!out) return fail(NSMH_EINVAL, "host_alloc_near: null output");
    cpu_set_t near, before;
    const bool have_before = sched_getaffinity(0, sizeof before, &before) == 0;
    const bool moved = have_before && device_numa_cpus(device, &near) >= 0 && sched_setaffinity(0, sizeof near, &near) == 0;
    // pages are taken from the node of the CPU that pins them (default policy: local allocation)
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (moved) sched_setaffinity(0, sizeof before, &before);
    if (e != cudaSuccess) return cuda_fail(e, "cudaHostAlloc", __FILE__, __LINE__);
    return NSMH_OK;
}

int nsmh_host_free(void *p) {
    if (p) NSMH_CK(cudaFreeHost(p));
    return NSMH_OK;
}

int nsmh_create(uint32_t k, uint32_t n, uint32_t thr, const uint64_t *rand_numbers, int device,
                nsmh_handle *out) {
    if (!out) return fail(NSMH_EINVAL, "create: null output handle");
    *out = nullptr;
    // k == 32 is undefined behaviour in the reference (1ull << 64, ReadFilter.cpp:145); k == 0 is
    // rejected by its CLI (main.cpp:120-121).
    if (k < 1 || k > 31) return fail(NSMH_EINVAL, "create: k must be in 1..31");
    if (n < 1) return fail(NSMH_EINVAL, "create: n must be >= 1");
    if (!rand_numbers) return fail(NSMH_EINVAL, "create: rand_numbers is null");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(NSMH_ECUDA, "create: no CUDA device available (this engine has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(NSMH_EINVAL, "create: device index out of range");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(NSMH_ECUDA, "create: cudaSetDevice failed");
    nsmh_ctx *c = new (std::nothrow) nsmh_ctx();
    if (!c) return fail(NSMH_ENOMEM, "create: out of host memory");
    c->device = device;
    c->k = k;
    c->n = n;
    c->thr = thr;
    c->rand.assign(rand_numbers, rand_numbers + n);
    int rc = NSMH_OK;
    do {
        cudaDeviceProp prop;
        if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { rc = cuda_fail(e, "props", __FILE__, __LINE__); break; }
        c->num_sms = prop.multiProcessorCount;
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) { rc = cuda_fail(e, "stream", __FILE__, __LINE__); break; }
        if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) { rc = cuda_fail(e, "stream", __FILE__, __LINE__); break; }
        for (auto &ev : c->ev)
            if ((e = cudaEventCreate(&ev)) != cudaSuccess) { rc = cuda_fail(e, "event", __FILE__, __LINE__); break; }
        if (rc) break;
        if ((e = cudaEventCreateWithFlags(&c->ev_order, cudaEventDisableTiming)) != cudaSuccess) { rc = cuda_fail(e, "event", __FILE__, __LINE__); break; }
        if ((e = cudaEventCreateWithFlags(&c->ev_cleared, cudaEventDisableTiming)) != cudaSuccess) { rc = cuda_fail(e, "event", __FILE__, __LINE__); break; }
        // keep freed blocks in the pool: the bench re-runs the same sizes every step
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thresh = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
        }
        if ((rc = c->d_rand.ensure(n * sizeof(uint64_t), c->stream))) break;
        if ((rc = c->counters.ensure(8 * sizeof(uint64_t), c->stream))) break;
        if ((e = cudaMemcpyAsync(c->d_rand.p, c->rand.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) { rc = cuda_fail(e, "memcpy", __FILE__, __LINE__); break; }
        if ((e = cudaMemsetAsync(c->counters.p, 0, 8 * sizeof(uint64_t), c->stream)) != cudaSuccess) { rc = cuda_fail(e, "memset", __FILE__, __LINE__); break; }
        if ((rc = build_filter_tables(c))) break;
        c->bulk.stream = c->stream;
    } while (0);
    if (rc) {
        nsmh_destroy(c);
        return rc;
    }
    *out = c;
    return NSMH_OK;
}

int nsmh_set_params(nsmh_handle c, uint32_t k, uint32_t n, uint32_t thr, const uint64_t *rand_numbers) {
    CTX_GUARD(c);
    if (k < 1 || k > 31) return fail(NSMH_EINVAL, "set_params: k must be in 1..31");
    if (n < 1) return fail(NSMH_EINVAL, "set_params: n must be >= 1");
    if (!rand_numbers) return fail(NSMH_EINVAL, "set_params: rand_numbers is null");
    if (c->mg) return fail(NSMH_ESTATE, "set_params: not allowed once nsmh_mg_init has been called");
    NSMH_CK(cudaStreamSynchronize(c->stream));
    NSMH_CK(cudaStreamSynchronize(c->copy_stream));   // a table pre-clear of the old geometry may be in flight
    // the loaded reads stay; everything derived from the old parameters goes
    c->sketched = false;
    c->tables.built = false;
    c->bulk_valid = false;
    c->table_sketches = nullptr;
    c->table_reads = 0;
    c->id_base = 0;
    c->precleared_rows = 0;
    c->precleared_ptr = nullptr;
    c->k = k;
    c->n = n;
    c->thr = thr;
    c->rand.assign(rand_numbers, rand_numbers + n);
    NSMH_TRY(c->d_rand.ensure(n * sizeof(uint64_t), c->stream));
    NSMH_CK(cudaMemcpyAsync(c->d_rand.p, c->rand.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    NSMH_TRY(build_filter_tables(c));
    return NSMH_OK;
}

int nsmh_destroy(nsmh_handle h) {
    if (!h) return NSMH_OK;
    std::string keep = g_err;
    {
        DeviceGuard guard(h->device);
        if (h->stream) cudaStreamSynchronize(h->stream);
        mg_destroy(h);
        for (QueryWs *ws : h->pool) {
            if (ws->stream) cudaStreamSynchronize(ws->stream);
            free_ws(*ws, ws->stream);
            if (ws->stream) cudaStreamDestroy(ws->stream);
            delete ws;
        }
        h->pool.clear();
        cudaStream_t s = h->stream;
        free_ws(h->bulk, s);
        DevBuf *bufs[] = {&h->d_rand, &h->d_ftab_first, &h->d_ftab_next, &h->d_ftab_hit3, &h->sketches,
                          &h->tile_start, &h->read_flags, &h->counters, &h->build_multi, &h->build_tmp,
                          &h->tables.slots, &h->tables.ids, &h->defer.buf};
        if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);      // a deferred fix-up may still read its buffers
        for (DevBuf *b : bufs) b->release(s);
        if (h->reads.external_offsets) { h->reads.offsets.p = nullptr; h->reads.offsets.cap = 0; }
        h->reads.release(s);
        if (s) cudaStreamSynchronize(s);
        if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
        for (auto &ev : h->ev) if (ev) cudaEventDestroy(ev);
        if (h->ev_order) cudaEventDestroy(h->ev_order);
        if (h->ev_cleared) cudaEventDestroy(h->ev_cleared);
        if (h->defer.filtered) cudaEventDestroy(h->defer.filtered);
        if (h->defer.fixed) cudaEventDestroy(h->defer.fixed);
        if (h->stream && h->owns_stream) cudaStreamDestroy(h->stream);
        if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
        cudaGetLastError();
    }
    delete h;
    g_err = keep;
    return NSMH_OK;
}

// -------------------------------------------------------------------- reads --
static int set_offsets_host(nsmh_ctx *c, const uint64_t *offsets, uint32_t num_reads) {
    ReadSet &rs = c->reads;
    if (rs.external_offsets) { rs.offsets.p = nullptr; rs.offsets.cap = 0; rs.external_offsets = false; }
    for (uint32_t i = 0; i < num_reads; ++i)
        if (offsets[i + 1] < offsets[i]) return fail(NSMH_EINVAL, "load_reads: offsets must be non-decreasing");
    if (offsets[0] != 0) return fail(NSMH_EINVAL, "load_reads: offsets[0] must be 0");
    rs.num_reads = num_reads;
    NSMH_TRY(rs.offsets.ensure(((size_t)num_reads + 1) * sizeof(uint64_t), c->stream));
    NSMH_CK(cudaMemcpyAsync(rs.offsets.p, offsets, ((size_t)num_reads + 1) * sizeof(uint64_t),
                            cudaMemcpyHostToDevice, c->stream));
    return NSMH_OK;
}

static void invalidate(nsmh_ctx *c) {
    c->reads_loaded = false;
    c->load_timed = false;
    c->flags_valid = false;
    c->sketched = false;
    c->tables.built = false;
    c->bulk_valid = false;
    c->table_sketches = nullptr;
    c->table_reads = 0;
    c->id_base = 0;
}

// chunk size of the host loaders; NSMH_LOAD_CHUNK_BYTES lets the tests cut small inputs into many chunks
static uint64_t load_chunk_bytes(uint64_t dflt) {
    const char *e = getenv("NSMH_LOAD_CHUNK_BYTES");
    if (!e || !*e) return dflt;
    const long long v = atoll(e);
    return v < 16 ? 16 : (uint64_t)v;
}

// sketch pass: what comes before the first and after the last sketch_reads call
static int sketch_begin(nsmh_ctx *c, bool pipelined = false) {
    c->sketched = false;
    c->tables.built = false;
    c->bulk_valid = false;
    NSMH_TRY(c->sketches.ensure(std::max<size_t>((size_t)c->reads.num_reads * c->n, 1) * sizeof(uint64_t), c->stream));
    // the tables that the next build fills are cleared on the copy stream while the reads are sketched
    // (the pipelined loaders: on c->stream itself, which waits for the first chunk anyway)
    if (c->mg && c->mg_sub()) NSMH_TRY(preclear_tables(c->mg_sub(), c->mg_total_rows(), pipelined));
    else NSMH_TRY(preclear_tables(c, c->reads.num_reads, pipelined));
    NSMH_CK(cudaEventRecord(c->ev[2], c->stream));
    return NSMH_OK;
}

static int sketch_end(nsmh_ctx *c) {
    NSMH_CK(cudaEventRecord(c->ev[3], c->stream));
    c->sketched = true;
    c->table_sketches = c->sketches.as<uint64_t>();
    c->table_reads = c->reads.num_reads;
    c->id_base = 0;
    return NSMH_OK;
}

// reads [r0, r1) are complete in the packed stream: sketch them behind whatever c->stream holds
static int sketch_range(nsmh_ctx *c, uint32_t r0, uint32_t r1) {
    const bool last = r1 == c->reads.num_reads;     // the main-kernel events time the last range
    return sketch_reads(c, c->reads, c->sketches.as<uint64_t>(), c->tile_start, c->build_tmp, c->sketch_mode,
                        c->stream, &c->launches, last ? c->ev[4] : nullptr, last ? c->ev[5] : nullptr, r0, r1);
}

static int build_enqueue(nsmh_ctx *c, SketchDeferred *defer = nullptr) {
    NSMH_CK(cudaEventRecord(c->ev[6], c->stream));
    NSMH_TRY(build_tables(c, defer));
    NSMH_CK(cudaEventRecord(c->ev[7], c->stream));
    c->build_timed = true;          // no host round trip here: nsmh_get_stats reads the events
    ++c->build_epoch;               // query workspaces order their streams behind ev[7] (order_after_build)
    return NSMH_OK;
}

// Host ASCII -> packed stream.  then_sketch: the reads that a chunk completes are sketched right behind its
// pack, i.e. while the next chunks are still crossing PCIe, and the tables are built behind the last one;
// the call returns when the last byte has LEFT the host buffers, not when the device is done.
static int load_ascii(nsmh_ctx *c, const char *bases, const uint64_t *offsets, uint32_t num_reads, int stages) {
    const bool then_sketch = stages >= 1;       // stages: 0 load, 1 + sketch, 2 + build
    if (!offsets) return fail(NSMH_EINVAL, "load_reads_ascii: null offsets");
    const uint64_t total = offsets[num_reads];
    if (total && !bases) return fail(NSMH_EINVAL, "load_reads_ascii: null bases");
    invalidate(c);
    NSMH_CK(cudaEventRecord(c->ev[0], c->stream));
    NSMH_TRY(set_offsets_host(c, offsets, num_reads));
    NSMH_TRY(alloc_packed(c->reads, total, c->stream));
    if (then_sketch) {
        c->reads_loaded = true;
        const int rc0 = sketch_begin(c, true);
        if (rc0) { invalidate(c); return rc0; }
    }
    // Double-buffered chunks: H2D of chunk i+1 on the copy stream overlaps the pack of chunk i.
    const uint64_t chunk = load_chunk_bytes(64ULL << 20) + 15 & ~15ULL;   // bases per chunk, multiple of 16
    DevBuf stage[2];
    cudaEvent_t copied[2] = {nullptr, nullptr}, packed[2] = {nullptr, nullptr};
    int rc = NSMH_OK;
    for (int i = 0; i < 2 && !rc; ++i) {
        cudaError_t ee;
        if ((ee = cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming)) != cudaSuccess) rc = cuda_fail(ee, "event", __FILE__, __LINE__);
        else if ((ee = cudaEventCreateWithFlags(&packed[i], cudaEventDisableTiming)) != cudaSuccess) rc = cuda_fail(ee, "event", __FILE__, __LINE__);
    }
    const size_t stage_bytes = (size_t)std::min<uint64_t>(chunk, total ? total : 1) + 16;
    for (int i = 0; i < 2 && !rc; ++i) rc = stage[i].ensure(stage_bytes, c->stream);
    if (!rc && cudaEventRecord(packed[0], c->stream) != cudaSuccess) rc = NSMH_ECUDA;
    if (!rc && cudaEventRecord(packed[1], c->stream) != cudaSuccess) rc = NSMH_ECUDA;
    int bi = 0, last_bi = -1;
    uint32_t r_done = 0;
    for (uint64_t b0 = 0; b0 < total && !rc; b0 += chunk, bi ^= 1) {
        const uint64_t nb = std::min<uint64_t>(chunk, total - b0);
        cudaError_t e;
        if ((e = cudaStreamWaitEvent(c->copy_stream, packed[bi], 0)) != cudaSuccess) { rc = cuda_fail(e, "wait", __FILE__, __LINE__); break; }
        if ((e = cudaMemcpyAsync(stage[bi].p, bases + b0, nb, cudaMemcpyHostToDevice, c->copy_stream)) != cudaSuccess) { rc = cuda_fail(e, "h2d", __FILE__, __LINE__); break; }
        if ((e = cudaEventRecord(copied[bi], c->copy_stream)) != cudaSuccess) { rc = cuda_fail(e, "record", __FILE__, __LINE__); break; }
        if ((e = cudaStreamWaitEvent(c->stream, copied[bi], 0)) != cudaSuccess) { rc = cuda_fail(e, "wait", __FILE__, __LINE__); break; }
        rc = pack_ascii(c->reads, stage[bi].as<char>(), b0, nb, c->stream, &c->launches);
        if (!rc && (e = cudaEventRecord(packed[bi], c->stream)) != cudaSuccess) { rc = cuda_fail(e, "record", __FILE__, __LINE__); break; }
        last_bi = bi;
        if (!rc && then_sketch) {
            // reads that end at or before b0 + nb are whole now
            const uint32_t r = (uint32_t)(std::upper_bound(offsets + r_done, offsets + num_reads + 1, b0 + nb) - offsets) - 1;
            if (r > r_done) rc = sketch_range(c, r_done, r);
            r_done = r;
        }
    }
    if (!rc) cudaEventRecord(c->ev[1], c->stream);
    if (!rc && then_sketch) {
        if (r_done < num_reads) rc = sketch_range(c, r_done, num_reads);    // empty reads at the very end
        if (!rc) rc = sketch_end(c);
        if (!rc && stages >= 2) rc = build_enqueue(c);
        // the caller may reuse its buffers as soon as the last copy is through
        if (!rc && last_bi >= 0) {
            cudaError_t e = cudaEventSynchronize(copied[last_bi]);
            if (e != cudaSuccess) rc = cuda_fail(e, "sync", __FILE__, __LINE__);
        }
    } else if (!rc) {
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "sync", __FILE__, __LINE__);
    }
    if (rc) {   // a copy may still be writing into a staging buffer: let it finish before the buffers go
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamSynchronize(c->stream);
        cudaGetLastError();
    }
    for (int i = 0; i < 2; ++i) {
        stage[i].release(c->stream);
        if (copied[i]) cudaEventDestroy(copied[i]);
        if (packed[i]) cudaEventDestroy(packed[i]);
    }
    if (rc) { invalidate(c); return rc; }
    c->load_timed = then_sketch;       // nsmh_get_stats reads the events once the stream has drained
    if (!then_sketch) c->stats.h2d_pack_ms = elapsed(c->ev[0], c->ev[1]);
    c->reads_loaded = true;
    return NSMH_OK;
}

int nsmh_load_reads_ascii(nsmh_handle c, const char *bases, const uint64_t *offsets, uint32_t num_reads) {
    CTX_GUARD(c);
    return load_ascii(c, bases, offsets, num_reads, 0);
}

int nsmh_initialize_ascii(nsmh_handle c, const char *bases, const uint64_t *offsets, uint32_t num_reads) {
    CTX_GUARD(c);
    return load_ascii(c, bases, offsets, num_reads, 2);
}

int nsmh_load_reads_ascii_device(nsmh_handle c, const char *d_bases, const uint64_t *d_offsets,
                                 uint32_t num_reads, uint64_t total_bases) {
    CTX_GUARD(c);
    if (!d_offsets) return fail(NSMH_EINVAL, "load_reads_ascii_device: null offsets");
    if (total_bases && !d_bases) return fail(NSMH_EINVAL, "load_reads_ascii_device: null bases");
    invalidate(c);
    ReadSet &rs = c->reads;
    if (!rs.external_offsets) rs.offsets.release(c->stream);
    rs.offsets.p = const_cast<uint64_t *>(d_offsets);   // borrowed, never freed
    rs.offsets.cap = ((size_t)num_reads + 1) * sizeof(uint64_t);
    rs.external_offsets = true;
    rs.num_reads = num_reads;
    NSMH_CK(cudaEventRecord(c->ev[0], c->stream));
    NSMH_TRY(alloc_packed(rs, total_bases, c->stream));
    NSMH_TRY(pack_ascii(rs, d_bases, 0, total_bases, c->stream, &c->launches));
    NSMH_CK(cudaEventRecord(c->ev[1], c->stream));
    c->reads_loaded = true;
    return NSMH_OK;
}

// Host DnaBitset bytes -> packed stream, in chunks of whole reads: the copy of chunk i+1 runs beside the
// re-layout (and, then_sketch, the sketch) of chunk i.  See load_ascii for what then_sketch returns on.
static int load_dnabitset(nsmh_ctx *c, const uint8_t *packed, const uint32_t *lengths, uint32_t num_reads, int stages) {
    const bool then_sketch = stages >= 1;
    if (num_reads && (!packed || !lengths)) return fail(NSMH_EINVAL, "load_reads_dnabitset: null input");
    invalidate(c);
    std::vector<uint64_t> off((size_t)num_reads + 1, 0), boff((size_t)num_reads + 1, 0);
    for (uint32_t i = 0; i < num_reads; ++i) {
        off[i + 1] = off[i] + lengths[i];
        boff[i + 1] = boff[i] + (lengths[i] + 3) / 4;
    }
    NSMH_CK(cudaEventRecord(c->ev[0], c->stream));
    // chunks of whole reads, about `chunk` source bytes each
    const uint64_t chunk = load_chunk_bytes(16ULL << 20);      // (64 M bases)
    std::vector<uint32_t> cut(1, 0);
    while (cut.back() < num_reads) {
        const uint32_t r0 = cut.back();
        uint32_t r1 = (uint32_t)(std::lower_bound(boff.begin() + r0, boff.end(), boff[r0] + chunk) - boff.begin());
        if (r1 <= r0) r1 = r0 + 1;
        cut.push_back(std::min(r1, num_reads));
    }
    const size_t nchunks = cut.size() - 1;
    DevBuf d_src, d_boff;
    cudaEvent_t ready = nullptr;
    std::vector<cudaEvent_t> copied(nchunks, nullptr);
    int rc = d_src.ensure(boff[num_reads] + 16, c->stream);
    cudaError_t e = cudaSuccess;
    if (!rc && (e = cudaEventCreateWithFlags(&ready, cudaEventDisableTiming)) != cudaSuccess) rc = cuda_fail(e, "event", __FILE__, __LINE__);
    for (size_t i = 0; i < nchunks && !rc; ++i)
        if ((e = cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming)) != cudaSuccess) rc = cuda_fail(e, "event", __FILE__, __LINE__);
    // The bus is the bottleneck: the two small tables go first (pageable memory: such a copy holds the host until
    // it is through, and behind 250 MB of queued chunks that took a millisecond), then every chunk is queued
    // before anything else is set up.  (The staging buffer was allocated in c->stream's order, the copy stream
    // may touch it after `ready`.)
    if (!rc) rc = set_offsets_host(c, off.data(), num_reads);
    if (!rc) rc = d_boff.ensure(boff.size() * sizeof(uint64_t), c->stream);
    if (!rc && (e = cudaMemcpyAsync(d_boff.p, boff.data(), boff.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) rc = cuda_fail(e, "h2d", __FILE__, __LINE__);
    if (!rc && (e = cudaEventRecord(ready, c->stream)) != cudaSuccess) rc = cuda_fail(e, "record", __FILE__, __LINE__);
    if (!rc && (e = cudaStreamWaitEvent(c->copy_stream, ready, 0)) != cudaSuccess) rc = cuda_fail(e, "wait", __FILE__, __LINE__);
    int last_copy = -1;
    for (size_t i = 0; i < nchunks && !rc; ++i) {
        const uint64_t b0 = boff[cut[i]], nbytes = boff[cut[i + 1]] - b0;
        if (!nbytes) continue;
        if ((e = cudaMemcpyAsync(static_cast<uint8_t *>(d_src.p) + b0, packed + b0, nbytes, cudaMemcpyHostToDevice, c->copy_stream)) != cudaSuccess) { rc = cuda_fail(e, "h2d", __FILE__, __LINE__); break; }
        if ((e = cudaEventRecord(copied[i], c->copy_stream)) != cudaSuccess) { rc = cuda_fail(e, "record", __FILE__, __LINE__); break; }
        last_copy = (int)i;
    }
    if (!rc) rc = alloc_packed(c->reads, off[num_reads], c->stream);
    if (!rc && then_sketch) {
        c->reads_loaded = true;
        rc = sketch_begin(c, true);
    }
    for (size_t i = 0; i < nchunks && !rc; ++i) {
        const uint32_t r0 = cut[i], r1 = cut[i + 1];
        if (boff[r1] > boff[r0]) {
            if ((e = cudaStreamWaitEvent(c->stream, copied[i], 0)) != cudaSuccess) { rc = cuda_fail(e, "wait", __FILE__, __LINE__); break; }
            // the word that holds the first base of read r1 is written again, whole, by the next chunk
            rc = pack_from_dnabitset(c->reads, d_src.as<uint8_t>(), d_boff.as<uint64_t>(), c->stream, &c->launches,
                                     off[r0] / 16, (off[r1] + 15) / 16);
        }
        if (!rc && then_sketch) rc = sketch_range(c, r0, r1);
    }
    if (!rc) cudaEventRecord(c->ev[1], c->stream);
    if (!rc && then_sketch) {
        rc = sketch_end(c);
        if (!rc && stages >= 2) rc = build_enqueue(c);
        // the caller may reuse its buffers as soon as the last copy is through
        if (!rc && last_copy >= 0 && (e = cudaEventSynchronize(copied[last_copy])) != cudaSuccess) rc = cuda_fail(e, "sync", __FILE__, __LINE__);
    } else if (!rc && (e = cudaStreamSynchronize(c->stream)) != cudaSuccess) rc = cuda_fail(e, "sync", __FILE__, __LINE__);
    if (rc) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamSynchronize(c->stream);
        cudaGetLastError();
    }
    d_src.release(c->stream);
    d_boff.release(c->stream);
    if (ready) cudaEventDestroy(ready);
    for (cudaEvent_t ev : copied)
        if (ev) cudaEventDestroy(ev);
    if (rc) { invalidate(c); return rc; }
    c->load_timed = true;
    c->reads_loaded = true;
    return NSMH_OK;
}

int nsmh_load_reads_dnabitset(nsmh_handle c, const uint8_t *packed, const uint32_t *lengths,
                              uint32_t num_reads) {
    CTX_GUARD(c);
    return load_dnabitset(c, packed, lengths, num_reads, 0);
}

int nsmh_load_sketch_ascii(nsmh_handle c, const char *bases, const uint64_t *offsets, uint32_t num_reads) {
    CTX_GUARD(c);
    return load_ascii(c, bases, offsets, num_reads, 1);
}

int nsmh_load_sketch_dnabitset(nsmh_handle c, const uint8_t *packed, const uint32_t *lengths, uint32_t num_reads) {
    CTX_GUARD(c);
    return load_dnabitset(c, packed, lengths, num_reads, 1);
}

int nsmh_initialize_dnabitset(nsmh_handle c, const uint8_t *packed, const uint32_t *lengths,
                              uint32_t num_reads) {
    CTX_GUARD(c);
    return load_dnabitset(c, packed, lengths, num_reads, 2);
}

// ------------------------------------------------------------ FASTQ ingest --
// (kernels in fastq.cu)
static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int nsmh_load_fastq_device(nsmh_handle c, const char *d_text, size_t bytes) {
    CTX_GUARD(c);
    if (bytes && !d_text) return fail(NSMH_EINVAL, "load_fastq_device: null text");
    invalidate(c);
    uint8_t last = 0;
    if (bytes) {
        NSMH_CK(cudaMemcpyAsync(&last, d_text + bytes - 1, 1, cudaMemcpyDeviceToHost, c->stream));
        NSMH_CK(cudaStreamSynchronize(c->stream));
    }
    NSMH_TRY(parse_fastq_device(c, reinterpret_cast<const uint8_t *>(d_text), bytes, bytes, last));
    c->reads_loaded = true;
    return NSMH_OK;
}

int nsmh_load_fastq(nsmh_handle c, const char *text, size_t bytes) {
    CTX_GUARD(c);
    if (bytes && !text) return fail(NSMH_EINVAL, "load_fastq: null text");
    const double t0 = now_ms();
    invalidate(c);
    DevBuf d_text;
    NSMH_TRY(d_text.ensure(bytes + 64, c->stream));
    int rc = NSMH_OK;
    cudaError_t e = cudaSuccess;
    if (bytes && (e = cudaMemcpyAsync(d_text.p, text, bytes, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess)
        rc = cuda_fail(e, "h2d", __FILE__, __LINE__);
    if (!rc) rc = parse_fastq_device(c, d_text.as<uint8_t>(), bytes, d_text.cap, bytes ? (uint8_t)text[bytes - 1] : 0);
    d_text.release(c->stream);
    if (rc) return rc;
    c->stats.fastq_load_ms = (float)(now_ms() - t0);
    c->reads_loaded = true;
    return NSMH_OK;
}

namespace {
// Sequential producer of the (inflated) file contents.
struct TextSource {
    FILE *f = nullptr;
    bool gz = false, eof_in = false, done = false, at_boundary = true, z_init = false;
    z_stream zs;
    std::vector<uint8_t> in;
    ~TextSource() {
        if (z_init) inflateEnd(&zs);
        if (f) fclose(f);
    }
    int open(const char *path, int gzip) {
        f = fopen(path, "rb");
        if (!f) return fail(NSMH_EINVAL, std::string("load_fastq_file: cannot open ") + path);
        gz = gzip != 0;
        if (gz) {
            memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, 15 + 32) != Z_OK) return fail(NSMH_ENOMEM, "load_fastq_file: inflateInit2 failed");
            z_init = true;
            in.resize(1 << 20);
        }
        return NSMH_OK;
    }
    // bytes produced into out (0 = end of data), -1 = corrupt / truncated input
    long fill(uint8_t *out, size_t cap) {
        if (!gz) return (long)fread(out, 1, cap, f);
        size_t produced = 0;
        while (produced < cap && !done) {
            if (zs.avail_in == 0 && !eof_in) {
                const size_t got = fread(in.data(), 1, in.size(), f);
                if (got == 0) eof_in = true;
                zs.next_in = in.data();
                zs.avail_in = (uInt)got;
            }
            if (zs.avail_in == 0 && eof_in) {
                if (!at_boundary) return -1;       // the file ends inside a gzip member
                done = true;
                break;
            }
            zs.next_out = out + produced;
            zs.avail_out = (uInt)std::min<size_t>(cap - produced, (size_t)INT_MAX);
            const uInt before = zs.avail_out;
            at_boundary = false;
            const int zr = inflate(&zs, Z_NO_FLUSH);
            produced += before - zs.avail_out;
            if (zr == Z_STREAM_END) {
                inflateReset(&zs);                 // a further gzip member may follow
                at_boundary = true;
            } else if (zr != Z_OK && zr != Z_BUF_ERROR) {
                return -1;
            }
        }
        return (long)produced;
    }
};
}  // namespace

int nsmh_load_fastq_file(nsmh_handle c, const char *path, int gzip) {
    CTX_GUARD(c);
    if (!path) return fail(NSMH_EINVAL, "load_fastq_file: null path");
    const double t0 = now_ms();
    TextSource src;
    NSMH_TRY(src.open(path, gzip));
    uint64_t file_size = 0;
    if (fseek(src.f, 0, SEEK_END) == 0) {
        const long sz = ftell(src.f);
        if (sz > 0) file_size = (uint64_t)sz;
        fseek(src.f, 0, SEEK_SET);
    }
    invalidate(c);
    const size_t chunk = 32u << 20;
    uint8_t *pin[2] = {nullptr, nullptr};
    cudaEvent_t copied[2] = {nullptr, nullptr};
    DevBuf d_text;
    int rc = NSMH_OK;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && !rc; ++i) {
        if ((e = cudaMallocHost(reinterpret_cast<void **>(&pin[i]), chunk)) != cudaSuccess) rc = cuda_fail(e, "pinned", __FILE__, __LINE__);
        if (!rc && (e = cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming)) != cudaSuccess) rc = cuda_fail(e, "event", __FILE__, __LINE__);
    }
    if (!rc) rc = d_text.ensure((gzip ? file_size * 4 : file_size) + chunk + 64, c->stream);
    uint64_t used = 0;
    int last = 0, bi = 0;
    bool in_flight[2] = {false, false};
    while (!rc) {
        if (in_flight[bi] && (e = cudaEventSynchronize(copied[bi])) != cudaSuccess) { rc = cuda_fail(e, "sync", __FILE__, __LINE__); break; }
        const long got = src.fill(pin[bi], chunk);     // overlaps the H2D copy of the other buffer
        if (got < 0) { rc = fail(NSMH_EINVAL, std::string("load_fastq_file: corrupt or truncated gzip data in ") + path); break; }
        if (got == 0) break;
        if (used + (uint64_t)got + 64 > d_text.cap)
            rc = d_text.ensure(std::max<uint64_t>(2 * d_text.cap, used + (uint64_t)got + 64), c->stream, used);
        if (rc) break;
        if ((e = cudaMemcpyAsync(d_text.as<uint8_t>() + used, pin[bi], (size_t)got, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) { rc = cuda_fail(e, "h2d", __FILE__, __LINE__); break; }
        if ((e = cudaEventRecord(copied[bi], c->stream)) != cudaSuccess) { rc = cuda_fail(e, "record", __FILE__, __LINE__); break; }
        in_flight[bi] = true;
        last = pin[bi][got - 1];
        used += (uint64_t)got;
        bi ^= 1;
    }
    if (!rc) rc = parse_fastq_device(c, d_text.as<uint8_t>(), used, d_text.cap, last);
    cudaStreamSynchronize(c->stream);
    d_text.release(c->stream);
    for (int i = 0; i < 2; ++i) {
        if (pin[i]) cudaFreeHost(pin[i]);
        if (copied[i]) cudaEventDestroy(copied[i]);
    }
    if (rc) return rc;
    c->stats.fastq_load_ms = (float)(now_ms() - t0);
    c->reads_loaded = true;
    return NSMH_OK;
}

int nsmh_read_offsets(nsmh_handle c, uint64_t *offsets) {
    CTX_GUARD(c);
    if (!c->reads_loaded) return fail(NSMH_ESTATE, "read_offsets: no reads loaded");
    if (!offsets) return fail(NSMH_EINVAL, "read_offsets: null output");
    NSMH_CK(cudaMemcpyAsync(offsets, c->reads.offsets.p, ((size_t)c->reads.num_reads + 1) * sizeof(uint64_t),
                            cudaMemcpyDeviceToHost, c->stream));
    NSMH_CK(cudaStreamSynchronize(c->stream));
    return NSMH_OK;
}

int nsmh_get_reads_ascii(nsmh_handle c, uint32_t first, uint32_t count, char *out) {
    CTX_GUARD(c);
    if (!c->reads_loaded) return fail(NSMH_ESTATE, "get_reads_ascii: no reads loaded");
    if ((uint64_t)first + count > c->reads.num_reads) return fail(NSMH_EINVAL, "get_reads_ascii: read range out of bounds");
    if (!count) return NSMH_OK;
    uint64_t be[2] = {0, 0};
    NSMH_CK(cudaMemcpyAsync(&be[0], c->reads.d_offsets() + first, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    NSMH_CK(cudaMemcpyAsync(&be[1], c->reads.d_offsets() + first + count, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    NSMH_CK(cudaStreamSynchronize(c->stream));
    const uint64_t nb = be[1] - be[0];
    if (!nb) return NSMH_OK;
    if (!out) return fail(NSMH_EINVAL, "get_reads_ascii: null output");
    DevBuf d;
    NSMH_TRY(d.ensure(nb + 16, c->stream));
    int rc = unpack_ascii_device(c, be[0], nb, d.as<uint8_t>(), c->stream);
    cudaError_t e = cudaSuccess;
    if (!rc && (e = cudaMemcpyAsync(out, d.p, nb, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) rc = cuda_fail(e, "d2h", __FILE__, __LINE__);
    if (!rc && (e = cudaStreamSynchronize(c->stream)) != cudaSuccess) rc = cuda_fail(e, "sync", __FILE__, __LINE__);
    d.release(c->stream);
    return rc;
}

int nsmh_num_reads(nsmh_handle c, uint32_t *num_reads, uint64_t *total_bases) {
    if (!c) return fail(NSMH_EINVAL, "null handle");
    if (num_reads) *num_reads = c->reads.num_reads;
    if (total_bases) *total_bases = c->reads.total_bases;
    return NSMH_OK;
}

// ------------------------------------------------------------- initialize() --
int nsmh_set_sketch_mode(nsmh_handle c, int mode) {
    if (!c) return fail(NSMH_EINVAL, "null handle");
    if (mode != 0 && mode != 1) return fail(NSMH_EINVAL, "set_sketch_mode: mode must be 0 or 1");
    c->sketch_mode = mode;
    return NSMH_OK;
}

int nsmh_sketch(nsmh_handle c) {
    CTX_GUARD(c);
    if (!c->reads_loaded) return fail(NSMH_ESTATE, "sketch: no reads loaded");
    NSMH_TRY(sketch_begin(c));
    NSMH_TRY(sketch_range(c, 0, c->reads.num_reads));
    return sketch_end(c);
}

// sketch + build with the sketch's exact fix-up (sketch_missing / sketch_fixup kernels, ALU-bound) on the second
// stream beside the table insert (bound by the L2 atomic rate): the two do not compete for the same unit, and the
// handful of entries the fix-up produces are inserted from a list afterwards.  When the call returns everything is
// queued on the context's stream in order, so any later call sees final sketches and tables.
int nsmh_sketch_build(nsmh_handle c) {
    CTX_GUARD(c);
    if (!c->reads_loaded) return fail(NSMH_ESTATE, "sketch_build: no reads loaded");
    const bool overlap = c->sketch_mode == 0 && c->n <= 255 && !c->mg;      // the brute-force kernel has no fix-up
    NSMH_TRY(sketch_begin(c));
    SketchDeferred *d = nullptr;
    if (overlap) {
        d = &c->defer;
        d->aux = c->copy_stream;
        if (!d->filtered) NSMH_CK(cudaEventCreateWithFlags(&d->filtered, cudaEventDisableTiming));
        if (!d->fixed) NSMH_CK(cudaEventCreateWithFlags(&d->fixed, cudaEventDisableTiming));
        d->pending = false;
    }
    NSMH_TRY(sketch_reads(c, c->reads, c->sketches.as<uint64_t>(), c->tile_start, c->build_tmp, c->sketch_mode, c->stream,
                          &c->launches, c->ev[4], c->ev[5], 0, ~0u, d));
    NSMH_TRY(sketch_end(c));
    const int rc = build_enqueue(c, d);
    if (rc && d && d->pending) {        // the fix-up never ran or its values were not stored: the sketches are not final
        cudaStreamSynchronize(c->copy_stream);
        c->sketched = false;
        d->pending = false;
    }
    return rc;
}

// The same idea for the multi-GPU flow: nsmh_sketch + nsmh_mg_run as one call, the fix-up pass beside the column
// scatter (which sends all-ones for the entries concerned; their values follow in mg_scatter_list_kernel).
int nsmh_mg_sketch_run(nsmh_handle c, uint64_t *total_ids) {
    CTX_GUARD(c);
    if (!c->reads_loaded) return fail(NSMH_ESTATE, "mg_sketch_run: no reads loaded");
    if (!c->mg) return fail(NSMH_ESTATE, "mg_sketch_run: call nsmh_mg_init and nsmh_mg_connect first");
    const bool overlap = c->sketch_mode == 0 && c->n <= 255 && c->reads.num_reads > 0;
    NSMH_TRY(sketch_begin(c));
    SketchDeferred *d = nullptr;
    if (overlap) {
        d = &c->defer;
        d->aux = c->copy_stream;
        if (!d->filtered) NSMH_CK(cudaEventCreateWithFlags(&d->filtered, cudaEventDisableTiming));
        if (!d->fixed) NSMH_CK(cudaEventCreateWithFlags(&d->fixed, cudaEventDisableTiming));
        d->pending = false;
    }
    NSMH_TRY(sketch_reads(c, c->reads, c->sketches.as<uint64_t>(), c->tile_start, c->build_tmp, c->sketch_mode, c->stream,
                          &c->launches, c->ev[4], c->ev[5], 0, ~0u, d));
    NSMH_TRY(sketch_end(c));
    const int rc = mg_run_impl(c, total_ids, d);
    if (d && d->pending) {              // the run failed before the fix-up was queued: the sketches are not final
        cudaStreamSynchronize(c->copy_stream);
        c->sketched = false;
        d->pending = false;
    }
    return rc;
}

int nsmh_get_sketches(nsmh_handle c, uint64_t *out) {
    CTX_GUARD(c);
    if (!c->sketched) return fail(NSMH_ESTATE, "get_sketches: call nsmh_sketch first");
    size_t bytes = (size_t)c->reads.num_reads * c->n * sizeof(uint64_t);
    if (bytes && !out) return fail(NSMH_EINVAL, "get_sketches: null output");
    if (bytes) NSMH_CK(cudaMemcpyAsync(out, c->sketches.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    NSMH_CK(cudaStreamSynchronize(c->stream));
    return NSMH_OK;
}

int nsmh_sketches_device_ptr(nsmh_handle c, uint64_t **d_ptr) {
    if (!c || !d_ptr) return fail(NSMH_EINVAL, "sketches_device_ptr: null argument");
    if (!c->sketched) return fail(NSMH_ESTATE, "sketches_device_ptr: call nsmh_sketch first");
    *d_ptr = c->sketches.as<uint64_t>();
    return NSMH_OK;
}

int nsmh_set_table_sketches(nsmh_handle c, const uint64_t *d_sketches, uint32_t table_reads,
                            uint32_t id_base) {
    if (!c) return fail(NSMH_EINVAL, "null handle");
    if (table_reads && !d_sketches) return fail(NSMH_EINVAL, "set_table_sketches: null pointer");
    c->table_sketches = d_sketches;
    c->table_reads = table_reads;
    c->id_base = id_base;
    c->tables.built = false;
    c->bulk_valid = false;
    return NSMH_OK;
}

int nsmh_build(nsmh_handle c) {
    CTX_GUARD(c);
    if (!c->table_sketches && c->table_reads)
        return fail(NSMH_ESTATE, "build: no sketches (call nsmh_sketch or nsmh_set_table_sketches)");
    if (!c->sketched && c->table_sketches == nullptr)
        return fail(NSMH_ESTATE, "build: no sketches (call nsmh_sketch or nsmh_set_table_sketches)");
    return build_enqueue(c);
}

int nsmh_table_num_keys(nsmh_handle c, uint32_t j, uint32_t *num_keys) {
    CTX_GUARD(c);
    if (!num_keys) return fail(NSMH_EINVAL, "table_num_keys: null output");
    return table_num_keys(c, j, num_keys);
}

// --------------------------------------------------------------- bulk query --
int nsmh_query_all(nsmh_handle c, int rc_mode, uint64_t *total_ids) {
    CTX_GUARD(c);
    if (!c->sketched) return fail(NSMH_ESTATE, "query_all: call nsmh_sketch first");
    if (!c->tables.built) return fail(NSMH_ESTATE, "query_all: call nsmh_build first");
    QueryWs &ws = c->bulk;
    ws.stream = c->stream;
    c->bulk_valid = false;
    cudaEvent_t e0 = c->ev[8], e1 = c->ev[9];
    int rc = NSMH_OK;
    cudaEventRecord(e0, c->stream);
    const uint64_t *q = c->sketches.as<uint64_t>();
    if (rc_mode) {
        // sketch the reverse-complement strings, as the caller's second query does
        // (Consensus.cpp:181-191 via the public string overload, ReadFilter.cpp:85-97)
        ws.str_reads.offsets.p = c->reads.offsets.p;   // same boundaries, borrowed
        ws.str_reads.offsets.cap = c->reads.offsets.cap;
        rc = pack_reverse_complement(c->reads, ws.str_reads, c->stream, &ws.launches);
        if (!rc) rc = ws.qsketch.ensure(std::max<size_t>((size_t)c->reads.num_reads * c->n, 1) * sizeof(uint64_t), c->stream);
        if (!rc) rc = sketch_reads(c, ws.str_reads, ws.qsketch.as<uint64_t>(), ws.tile_start, ws.cub_tmp,
                                   c->sketch_mode, c->stream, &ws.launches, nullptr, nullptr);
        ws.str_reads.offsets.p = nullptr;
        ws.str_reads.offsets.cap = 0;
        q = ws.qsketch.as<uint64_t>();
    }
    if (!rc) rc = query_sketches_device(c, ws, q, c->reads.num_reads, c->stream);
    cudaEventRecord(e1, c->stream);
    cudaStreamSynchronize(c->stream);       // e1 only: the lookup itself ended with its one read-back
    if (!rc) {
        c->stats.query_ms = elapsed(e0, e1);
        c->stats.query_pairs = ws.last_pairs;
        c->stats.query_heavy = ws.last_heavy;
        c->stats.query_sorted = ws.last_sorted;
        c->bulk_valid = true;
        if (total_ids) *total_ids = ws.last_total;
    }
    return rc;
}

int nsmh_query_all_result(nsmh_handle c, uint64_t *offsets, uint32_t *ids) {
    CTX_GUARD(c);
    if (!c->bulk_valid) return fail(NSMH_ESTATE, "query_all_result: no bulk query result");
    QueryWs &ws = c->bulk;
    if (offsets)
        NSMH_CK(cudaMemcpyAsync(offsets, ws.out_off.p, ((size_t)ws.last_nq + 1) * sizeof(uint64_t),
                                cudaMemcpyDeviceToHost, c->stream));
    if (ids && ws.last_total)
        NSMH_CK(cudaMemcpyAsync(ids, ws.out_ids.p, ws.last_total * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                c->stream));
    NSMH_CK(cudaStreamSynchronize(c->stream));
    return NSMH_OK;
}

int nsmh_query_all_device_ptrs(nsmh_handle c, uint64_t **d_offsets, uint32_t **d_ids) {
    if (!c) return fail(NSMH_EINVAL, "null handle");
    if (!c->bulk_valid) return fail(NSMH_ESTATE, "query_all_device_ptrs: no bulk query result");
    if (d_offsets) *d_offsets = c->bulk.out_off.as<uint64_t>();
    if (d_ids) *d_ids = c->bulk.out_ids.as<uint32_t>();
    return NSMH_OK;
}

// ------------------------------------------------- caller-side pre-filters --
static int ensure_flags(nsmh_ctx *c) {
    if (!c->reads_loaded) return fail(NSMH_ESTATE, "read_flags: no reads loaded");
    if (c->flags_valid) return NSMH_OK;
    NSMH_TRY(compute_read_flags(c));
    c->flags_valid = true;
    return NSMH_OK;
}

int nsmh_read_flags(nsmh_handle c, uint8_t *out) {
    CTX_GUARD(c);
    NSMH_TRY(ensure_flags(c));
    if (out && c->reads.num_reads)
        NSMH_CK(cudaMemcpyAsync(out, c->read_flags.p, c->reads.num_reads, cudaMemcpyDeviceToHost, c->stream));
    NSMH_CK(cudaStreamSynchronize(c->stream));
    return NSMH_OK;
}

int nsmh_read_flags_device_ptr(nsmh_handle c, uint8_t **d_flags) {
    CTX_GUARD(c);
    if (!d_flags) return fail(NSMH_EINVAL, "read_flags_device_ptr: null output");
    NSMH_TRY(ensure_flags(c));
    *d_flags = c->read_flags.as<uint8_t>();
    return NSMH_OK;
}

int nsmh_query_all_drop(nsmh_handle c, uint32_t drop_mask, uint64_t *total_ids) {
    CTX_GUARD(c);
    if (!c->bulk_valid) return fail(NSMH_ESTATE, "query_all_drop: no bulk query result");
    if (c->mg || c->id_base != 0 || c->table_reads != c->reads.num_reads)
        return fail(NSMH_ESTATE, "query_all_drop: candidate ids are not ids of the loaded reads (multi-GPU result)");
    NSMH_TRY(ensure_flags(c));
    if (drop_mask) NSMH_TRY(drop_flagged_candidates(c, c->bulk, drop_mask, c->stream));
    if (total_ids) *total_ids = c->bulk.last_total;
    return NSMH_OK;
}

int nsmh_probe_lists(nsmh_handle c, const uint64_t *d_sketches, uint32_t num_queries, uint64_t *total_ids) {
    CTX_GUARD(c);
    if (num_queries && !d_sketches) return fail(NSMH_EINVAL, "probe_lists: null sketches");
    if (!c->tables.built) return fail(NSMH_ESTATE, "probe_lists: call nsmh_build first");
    c->bulk.stream = c->stream;
    c->bulk_valid = false;
    NSMH_TRY(probe_lists_device(c, c->bulk, d_sketches, num_queries, c->stream));
    c->bulk_valid = true;
    if (total_ids) *total_ids = c->bulk.last_total;
    return NSMH_OK;
}

int nsmh_count_lists(nsmh_handle c, uint32_t num_queries, uint32_t parts, const uint64_t *const *d_offsets,
                     const uint32_t *const *d_ids, uint64_t *total_ids) {
    CTX_GUARD(c);
    if (!d_offsets || !d_ids) return fail(NSMH_EINVAL, "count_lists: null pointer arrays");
    c->bulk.stream = c->stream;
    c->bulk_valid = false;
    NSMH_TRY(count_lists_device(c, c->bulk, num_queries, parts, d_offsets, d_ids, c->stream));
    c->bulk_valid = true;
    c->stats.query_pairs = c->bulk.last_pairs;
    c->stats.query_heavy = c->bulk.last_heavy;
    c->stats.query_sorted = c->bulk.last_sorted;
    if (total_ids) *total_ids = c->bulk.last_total;
    return NSMH_OK;
}

// ------------------------------------------------------------- online query --
static QueryWs *acquire_ws(nsmh_ctx *c) {
    {
        std::lock_guard<std::mutex> lk(c->pool_mu);
        if (!c->pool.empty()) {
            QueryWs *ws = c->pool.back();
            c->pool.pop_back();
            return ws;
        }
    }
    QueryWs *ws = new (std::nothrow) QueryWs();
    if (!ws) return nullptr;
    if (cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ws;
        return nullptr;
    }
    return ws;
}

// nsmh_build / nsmh_initialize_* / nsmh_sketch_build only QUEUE the build on the context's stream; a query
// workspace has a stream of its own, so before its first query against new tables it waits for the event that
// marks their completion (once per build and workspace: the online path stays at one launch + one wait per call).
static int order_after_build(nsmh_ctx *c, QueryWs &ws) {
    if (ws.stream != c->stream && ws.seen_build_epoch != c->build_epoch) {
        NSMH_CK(cudaStreamWaitEvent(ws.stream, c->ev[7], 0));
        ws.seen_build_epoch = c->build_epoch;
    }
    return NSMH_OK;
}

static void release_ws(nsmh_ctx *c, QueryWs *ws) {
    std::lock_guard<std::mutex> lk(c->pool_mu);
    c->launches += ws->launches;
    ws->launches = 0;
    c->pool.push_back(ws);
}

static int copy_csr_out(QueryWs &ws, uint32_t nq, uint64_t *offsets_out, uint32_t *ids_out, size_t cap) {
    cudaStream_t s = ws.stream;
    const size_t ncopy = std::min<size_t>(cap, ws.last_total);
    if (offsets_out)
        NSMH_CK(cudaMemcpyAsync(offsets_out, ws.out_off.p, ((size_t)nq + 1) * sizeof(uint64_t),
                                cudaMemcpyDeviceToHost, s));
    if (ncopy && ids_out)
        NSMH_CK(cudaMemcpyAsync(ids_out, ws.out_ids.p, ncopy * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    NSMH_CK(cudaStreamSynchronize(s));
    if (ws.last_total > cap) {
        char buf[128];
        snprintf(buf, sizeof buf, "query: output needs %llu ids, capacity is %zu",
                 (unsigned long long)ws.last_total, cap);
        return fail(NSMH_ERANGE, buf);
    }
    return NSMH_OK;
}

static int query_strings_impl(nsmh_ctx *c, QueryWs &ws, const char *bases, const uint64_t *offsets,
                              uint32_t nq, uint64_t *offsets_out, uint32_t *ids_out, size_t cap) {
    cudaStream_t s = ws.stream;
    const uint64_t total = offsets[nq];
    for (uint32_t i = 0; i < nq; ++i)
        if (offsets[i + 1] < offsets[i]) return fail(NSMH_EINVAL, "query_strings: offsets must be non-decreasing");
    if (offsets[0] != 0) return fail(NSMH_EINVAL, "query_strings: offsets[0] must be 0");
    {
        const int rc = online_query(c, ws, bases, offsets, nq, offsets_out, ids_out, cap);
        if (rc != 1) return rc;             // answered (or failed for good); 1 = take the general path below
    }
    // stage strings + offsets through pinned memory so concurrent callers never block each other
    const size_t off_bytes = ((size_t)nq + 1) * sizeof(uint64_t);
    NSMH_TRY(ensure_pinned(ws, off_bytes + total + 16));
    memcpy(ws.h_pinned, offsets, off_bytes);
    if (total) memcpy(static_cast<char *>(ws.h_pinned) + off_bytes, bases, total);
    ReadSet &rs = ws.str_reads;
    rs.num_reads = nq;
    NSMH_TRY(rs.offsets.ensure(off_bytes, s));
    NSMH_TRY(ws.str_bases.ensure(total + 16, s));
    NSMH_CK(cudaMemcpyAsync(rs.offsets.p, ws.h_pinned, off_bytes, cudaMemcpyHostToDevice, s));
    if (total)
        NSMH_CK(cudaMemcpyAsync(ws.str_bases.p, static_cast<char *>(ws.h_pinned) + off_bytes, total,
                                cudaMemcpyHostToDevice, s));
    NSMH_TRY(alloc_packed(rs, total, s));
    NSMH_TRY(pack_ascii(rs, ws.str_bases.as<char>(), 0, total, s, &ws.launches));
    NSMH_TRY(ws.qsketch.ensure(std::max<size_t>((size_t)nq * c->n, 1) * sizeof(uint64_t), s));
    NSMH_TRY(sketch_reads(c, rs, ws.qsketch.as<uint64_t>(), ws.tile_start, ws.cub_tmp, c->sketch_mode, s,
                          &ws.launches, nullptr, nullptr));
    NSMH_TRY(query_sketches_device(c, ws, ws.qsketch.as<uint64_t>(), nq, s));
    return copy_csr_out(ws, nq, offsets_out, ids_out, cap);
}

int nsmh_query_strings(nsmh_handle c, const char *bases, const uint64_t *offsets, uint32_t num_strings,
                       uint64_t *offsets_out, uint32_t *ids_out, size_t cap) {
    CTX_GUARD(c);
    if (!offsets) return fail(NSMH_EINVAL, "query_strings: null offsets");
    if (offsets[num_strings] && !bases) return fail(NSMH_EINVAL, "query_strings: null bases");
    if (!c->tables.built) return fail(NSMH_ESTATE, "query_strings: call nsmh_build first");
    QueryWs *ws = acquire_ws(c);
    if (!ws) return fail(NSMH_ENOMEM, "query_strings: cannot create a query workspace");
    int rc = order_after_build(c, *ws);
    if (!rc) rc = query_strings_impl(c, *ws, bases, offsets, num_strings, offsets_out, ids_out, cap);
    release_ws(c, ws);
    return rc;
}

int nsmh_query_string(nsmh_handle c, const char *s, size_t len, uint32_t *out, size_t cap, size_t *count) {
    uint64_t offsets[2] = {0, (uint64_t)len};
    uint64_t off_out[2] = {0, 0};
    int rc = nsmh_query_strings(c, s, offsets, 1, off_out, out, cap);
    if (count && (rc == NSMH_OK || rc == NSMH_ERANGE)) *count = (size_t)off_out[1];
    return rc;
}

int nsmh_query_sketches(nsmh_handle c, const uint64_t *sketches, uint32_t num_queries,
                        uint64_t *offsets_out, uint32_t *ids_out, size_t cap) {
    CTX_GUARD(c);
    if (num_queries && !sketches) return fail(NSMH_EINVAL, "query_sketches: null sketches");
    if (!c->tables.built) return fail(NSMH_ESTATE, "query_sketches: call nsmh_build first");
    QueryWs *ws = acquire_ws(c);
    if (!ws) return fail(NSMH_ENOMEM, "query_sketches: cannot create a query workspace");
    int rc = NSMH_OK;
    const size_t bytes = (size_t)num_queries * c->n * sizeof(uint64_t);
    do {
        if ((rc = order_after_build(c, *ws))) break;
        if ((rc = ws->qsketch.ensure(std::max<size_t>(bytes, 8), ws->stream))) break;
        if ((rc = ensure_pinned(*ws, bytes + 16))) break;
        memcpy(ws->h_pinned, sketches, bytes);
        if (bytes) {
            cudaError_t e = cudaMemcpyAsync(ws->qsketch.p, ws->h_pinned, bytes, cudaMemcpyHostToDevice, ws->stream);
            if (e != cudaSuccess) { rc = cuda_fail(e, "h2d", __FILE__, __LINE__); break; }
        }
        if ((rc = query_sketches_device(c, *ws, ws->qsketch.as<uint64_t>(), num_queries, ws->stream))) break;
        rc = copy_csr_out(*ws, num_queries, offsets_out, ids_out, cap);
    } while (0);
    release_ws(c, ws);
    return rc;
}

// ----------------------------------------------------------- instrumentation --
int nsmh_get_stats(nsmh_handle c, nsmh_stats *out) {
    CTX_GUARD(c);
    if (!out) return fail(NSMH_EINVAL, "get_stats: null output");
    NSMH_CK(cudaStreamSynchronize(c->stream));
    // device-resident loads return before the pack kernel has run: its events are read here
    if (c->reads.external_offsets && c->reads_loaded) c->stats.pack_ms = elapsed(c->ev[0], c->ev[1]);
    if (c->load_timed) c->stats.h2d_pack_ms = elapsed(c->ev[0], c->ev[1]);     // pipelined: includes the sketches of all but the last chunk
    if (c->build_timed) c->stats.build_ms = elapsed(c->ev[6], c->ev[7]);
    if (c->sketched) {
        c->stats.sketch_ms = elapsed(c->ev[2], c->ev[3]);
        c->stats.sketch_main_ms = elapsed(c->ev[4], c->ev[5]);
    }
    unsigned long long fix = 0;
    NSMH_CK(cudaMemcpy(&fix, c->counters.p, sizeof fix, cudaMemcpyDeviceToHost));
    c->stats.sketch_fixups = fix;
    c->stats.kernel_launches = c->launches + c->bulk.launches;
    *out = c->stats;
    return NSMH_OK;
}

int nsmh_stream(nsmh_handle c, void **stream) {
    if (!c || !stream) return fail(NSMH_EINVAL, "stream: null argument");
    *stream = c->stream;
    return NSMH_OK;
}

int nsmh_synchronize(nsmh_handle c) {
    CTX_GUARD(c);
    NSMH_CK(cudaStreamSynchronize(c->stream));
    return NSMH_OK;
}

} // extern "C"

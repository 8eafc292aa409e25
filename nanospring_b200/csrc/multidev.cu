// One process, several GPUs: the drop-in for the reference's own process model.
//
// NanoSpring is ONE process with OpenMP threads (src/main.cpp:35, src/Compressor.cpp:55): initialize() is
// called once, getFilteredReads() concurrently from every thread (src/Consensus.cpp:29,189).  nsmh_mg_* needs
// one process per GPU; this layer gives the same caller all GPUs of the box behind ONE handle:
//
//   load     the reads are split into `ndev` contiguous shards with equal numbers of BASES,
//            one shard per device, host -> device copies of all shards in flight together
//   sketch   every device sketches its shard (94 % of the reference's initialize() time)
//   build    every device pulls the other shards' sketch rows over NVLink (cudaMemcpyPeerAsync: 8 n bytes
//            per read) and builds the FULL tables, ids = global read ids.  The replicated build costs
//            (total reads) x n inserts per device, a few ms per million reads - paid once, and what it buys is
//            that every device can answer ANY query alone:
//   query    getFilteredReads(string): the calling thread is sent to one device (round robin), whose tables
//            hold all reads - no cross-device traffic on the latency-critical online path, and `ndev` times
//            the online query throughput;  bulk query: every device queries its own shard, the CSRs are
//            concatenated in shard order = the single-GPU result.
//
// Everything goes through the public single-device entry points of api.cu; no kernel is specific to this file.
#include <atomic>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "nsmh_internal.cuh"

using namespace nsmh;

struct nsmh_multi {
    std::vector<nsmh_handle> sub;
    std::vector<int> devices;
    std::vector<uint32_t> first_read;          // [ndev + 1] shard boundaries (read ids)
    std::vector<uint64_t *> gathered;          // per device: [total_reads][n] sketch matrix (cudaMalloc)
    uint32_t k = 0, n = 0, thr = 0, num_reads = 0;
    uint64_t total_bases = 0;
    std::atomic<uint32_t> next{0};
    bool loaded = false, sketched = false, built = false;
    std::vector<uint64_t> bulk_total;          // per device: ids of the last bulk query
};

namespace {

// run f(d) for every device on its own host thread (the calls block on copies / synchronisations)
template <typename F>
int for_each_device(nsmh_multi *m, F f) {
    const size_t nd = m->sub.size();
    std::vector<int> rc(nd, NSMH_OK);
    std::vector<std::string> err(nd);
    std::vector<std::thread> th;
    for (size_t d = 0; d < nd; ++d)
        th.emplace_back([&, d] {
            rc[d] = f((uint32_t)d);
            if (rc[d]) err[d] = nsmh_last_error();      // thread-local in api.cu: carry it over
        });
    for (auto &t : th) t.join();
    for (size_t d = 0; d < nd; ++d)
        if (rc[d]) return fail(rc[d], "device " + std::to_string(m->devices[d]) + ": " + err[d]);
    return NSMH_OK;
}

void free_gathered(nsmh_multi *m) {
    for (size_t d = 0; d < m->gathered.size(); ++d)
        if (m->gathered[d]) {
            cudaSetDevice(m->devices[d]);
            cudaFree(m->gathered[d]);
            m->gathered[d] = nullptr;
        }
}

// shard boundaries on the prefix sum of read lengths: equal bases, not equal read counts
void split_by_bases(const uint64_t *offsets, uint32_t num_reads, uint32_t nd, std::vector<uint32_t> &first) {
    first.assign(nd + 1, num_reads);
    first[0] = 0;
    const uint64_t total = offsets[num_reads];
    uint32_t i = 0;
    for (uint32_t d = 1; d < nd; ++d) {
        const uint64_t target = total / nd * d + total % nd * d / nd;
        while (i < num_reads && offsets[i] < target) ++i;
        first[d] = i;
    }
}

}  // namespace

extern "C" {

int nsmh_multi_create(uint32_t k, uint32_t n, uint32_t overlap_sketch_thr, const uint64_t *rand_numbers,
                      const int *devices, int ndev, nsmh_multi_handle *out) {
    if (!out) return fail(NSMH_EINVAL, "multi_create: null output");
    *out = nullptr;
    if (!devices || ndev < 1 || ndev > NSMH_MG_MAX_RANKS) return fail(NSMH_EINVAL, "multi_create: 1..16 devices expected");
    nsmh_multi *m = new (std::nothrow) nsmh_multi();
    if (!m) return fail(NSMH_ENOMEM, "multi_create: out of host memory");
    m->k = k;
    m->n = n;
    m->thr = overlap_sketch_thr;
    int rc = NSMH_OK;
    for (int d = 0; d < ndev && !rc; ++d) {
        nsmh_handle h = nullptr;
        rc = nsmh_create(k, n, overlap_sketch_thr, rand_numbers, devices[d], &h);
        if (!rc) {
            m->sub.push_back(h);
            m->devices.push_back(devices[d]);
        }
    }
    // the sketch rows travel device to device when the tables are built
    for (int a = 0; a < ndev && !rc; ++a)
        for (int b = 0; b < ndev; ++b) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            if (!can) continue;                         // cudaMemcpyPeerAsync then stages through the host
            cudaSetDevice(devices[a]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
            cudaGetLastError();
        }
    if (rc) {
        std::string keep = nsmh_last_error();
        for (auto h : m->sub) nsmh_destroy(h);
        delete m;
        set_error(keep);
        return rc;
    }
    m->gathered.assign(ndev, nullptr);
    m->bulk_total.assign(ndev, 0);
    *out = m;
    return NSMH_OK;
}

int nsmh_multi_destroy(nsmh_multi_handle m) {
    if (!m) return NSMH_OK;
    free_gathered(m);
    for (auto h : m->sub) nsmh_destroy(h);
    delete m;
    return NSMH_OK;
}

int nsmh_multi_num_devices(nsmh_multi_handle m, int *ndev) {
    if (!m || !ndev) return fail(NSMH_EINVAL, "multi_num_devices: null argument");
    *ndev = (int)m->sub.size();
    return NSMH_OK;
}

int nsmh_multi_shards(nsmh_multi_handle m, uint32_t *first_read /* [ndev + 1] */) {
    if (!m || !first_read) return fail(NSMH_EINVAL, "multi_shards: null argument");
    if (!m->loaded) return fail(NSMH_ESTATE, "multi_shards: no reads loaded");
    std::memcpy(first_read, m->first_read.data(), m->first_read.size() * sizeof(uint32_t));
    return NSMH_OK;
}

int nsmh_multi_load_reads_ascii(nsmh_multi_handle m, const char *bases, const uint64_t *offsets, uint32_t num_reads) {
    if (!m) return fail(NSMH_EINVAL, "null handle");
    if (num_reads && (!bases || !offsets)) return fail(NSMH_EINVAL, "multi_load_reads_ascii: null input");
    m->loaded = m->sketched = m->built = false;
    const uint64_t zero = 0;
    if (!num_reads) offsets = &zero;
    split_by_bases(offsets, num_reads, (uint32_t)m->sub.size(), m->first_read);
    m->num_reads = num_reads;
    m->total_bases = offsets[num_reads];
    NSMH_TRY(for_each_device(m, [&](uint32_t d) {
        const uint32_t lo = m->first_read[d], hi = m->first_read[d + 1];
        std::vector<uint64_t> off((size_t)(hi - lo) + 1);
        for (uint32_t i = lo; i <= hi; ++i) off[i - lo] = offsets[i] - offsets[lo];
        return nsmh_load_reads_ascii(m->sub[d], bases + offsets[lo], off.data(), hi - lo);
    }));
    m->loaded = true;
    return NSMH_OK;
}

int nsmh_multi_load_reads_dnabitset(nsmh_multi_handle m, const uint8_t *packed, const uint32_t *lengths, uint32_t num_reads) {
    if (!m) return fail(NSMH_EINVAL, "null handle");
    if (num_reads && (!packed || !lengths)) return fail(NSMH_EINVAL, "multi_load_reads_dnabitset: null input");
    m->loaded = m->sketched = m->built = false;
    std::vector<uint64_t> off((size_t)num_reads + 1, 0), boff((size_t)num_reads + 1, 0);
    for (uint32_t i = 0; i < num_reads; ++i) {
        off[i + 1] = off[i] + lengths[i];
        boff[i + 1] = boff[i] + (lengths[i] + 3) / 4;
    }
    split_by_bases(off.data(), num_reads, (uint32_t)m->sub.size(), m->first_read);
    m->num_reads = num_reads;
    m->total_bases = off[num_reads];
    NSMH_TRY(for_each_device(m, [&](uint32_t d) {
        const uint32_t lo = m->first_read[d], hi = m->first_read[d + 1];
        return nsmh_load_reads_dnabitset(m->sub[d], packed + boff[lo], lengths + lo, hi - lo);
    }));
    m->loaded = true;
    return NSMH_OK;
}

int nsmh_multi_num_reads(nsmh_multi_handle m, uint32_t *num_reads, uint64_t *total_bases) {
    if (!m) return fail(NSMH_EINVAL, "null handle");
    if (num_reads) *num_reads = m->num_reads;
    if (total_bases) *total_bases = m->total_bases;
    return NSMH_OK;
}

int nsmh_multi_sketch(nsmh_multi_handle m) {
    if (!m) return fail(NSMH_EINVAL, "null handle");
    if (!m->loaded) return fail(NSMH_ESTATE, "multi_sketch: no reads loaded");
    m->sketched = m->built = false;
    // nsmh_sketch only queues work: the devices sketch side by side
    for (auto h : m->sub) NSMH_TRY(nsmh_sketch(h));
    m->sketched = true;
    return NSMH_OK;
}

int nsmh_multi_get_sketches(nsmh_multi_handle m, uint64_t *out) {
    if (!m) return fail(NSMH_EINVAL, "null handle");
    if (!m->sketched) return fail(NSMH_ESTATE, "multi_get_sketches: call nsmh_multi_sketch first");
    return for_each_device(m, [&](uint32_t d) { return nsmh_get_sketches(m->sub[d], out + (size_t)m->first_read[d] * m->n); });
}

int nsmh_multi_build(nsmh_multi_handle m) {
    if (!m) return fail(NSMH_EINVAL, "null handle");
    if (!m->sketched) return fail(NSMH_ESTATE, "multi_build: call nsmh_multi_sketch first");
    m->built = false;
    const size_t nd = m->sub.size();
    const size_t row_bytes = (size_t)m->n * sizeof(uint64_t);
    free_gathered(m);
    if (nd == 1) {
        NSMH_TRY(nsmh_build(m->sub[0]));
        m->built = true;
        return NSMH_OK;
    }
    for (auto h : m->sub) NSMH_TRY(nsmh_synchronize(h));            // every shard's sketch rows are complete
    NSMH_TRY(for_each_device(m, [&](uint32_t d) {
        NSMH_CK(cudaSetDevice(m->devices[d]));
        NSMH_CK(cudaMalloc(reinterpret_cast<void **>(&m->gathered[d]), std::max<size_t>((size_t)m->num_reads * row_bytes, 16)));
        void *sv = nullptr;
        NSMH_TRY(nsmh_stream(m->sub[d], &sv));
        cudaStream_t s = static_cast<cudaStream_t>(sv);
        for (size_t o = 0; o < nd; ++o) {
            const uint32_t lo = m->first_read[o], rows = m->first_read[o + 1] - lo;
            if (!rows) continue;
            uint64_t *src = nullptr;
            NSMH_TRY(nsmh_sketches_device_ptr(m->sub[o], &src));
            NSMH_CK(cudaMemcpyPeerAsync(m->gathered[d] + (size_t)lo * m->n, m->devices[d], src, m->devices[o],
                                        (size_t)rows * row_bytes, s));
        }
        NSMH_TRY(nsmh_set_table_sketches(m->sub[d], m->gathered[d], m->num_reads, m->first_read[d]));
        NSMH_TRY(nsmh_build(m->sub[d]));
        return nsmh_synchronize(m->sub[d]);
    }));
    m->built = true;
    return NSMH_OK;
}

int nsmh_multi_query_string(nsmh_multi_handle m, const char *s, size_t len, uint32_t *out, size_t cap, size_t *count) {
    if (!m) return fail(NSMH_EINVAL, "null handle");
    if (!m->built) return fail(NSMH_ESTATE, "multi_query_string: call nsmh_multi_build first");
    // every device holds the tables of ALL reads: any one of them answers; callers are spread round robin
    static thread_local uint32_t mine = 0xFFFFFFFFu;
    if (mine == 0xFFFFFFFFu) mine = m->next.fetch_add(1, std::memory_order_relaxed);
    return nsmh_query_string(m->sub[mine % m->sub.size()], s, len, out, cap, count);
}

int nsmh_multi_query_all(nsmh_multi_handle m, int rc_mode, uint64_t *total_ids) {
    if (!m) return fail(NSMH_EINVAL, "null handle");
    if (!m->built) return fail(NSMH_ESTATE, "multi_query_all: call nsmh_multi_build first");
    NSMH_TRY(for_each_device(m, [&](uint32_t d) { return nsmh_query_all(m->sub[d], rc_mode, &m->bulk_total[d]); }));
    uint64_t t = 0;
    for (auto x : m->bulk_total) t += x;
    if (total_ids) *total_ids = t;
    return NSMH_OK;
}

int nsmh_multi_query_all_result(nsmh_multi_handle m, uint64_t *offsets, uint32_t *ids) {
    if (!m) return fail(NSMH_EINVAL, "null handle");
    if (!m->built) return fail(NSMH_ESTATE, "multi_query_all_result: no bulk query result");
    std::vector<uint64_t> base(m->sub.size() + 1, 0);
    for (size_t d = 0; d < m->sub.size(); ++d) base[d + 1] = base[d] + m->bulk_total[d];
    NSMH_TRY(for_each_device(m, [&](uint32_t d) {
        const uint32_t lo = m->first_read[d], hi = m->first_read[d + 1];
        std::vector<uint64_t> off((size_t)(hi - lo) + 1, 0);
        NSMH_TRY(nsmh_query_all_result(m->sub[d], offsets ? off.data() : nullptr, ids ? ids + base[d] : nullptr));
        if (offsets)
            for (uint32_t i = lo; i < hi; ++i) offsets[i] = off[i - lo] + base[d];
        return NSMH_OK;
    }));
    if (offsets) offsets[m->num_reads] = base[m->sub.size()];
    return NSMH_OK;
}

} // extern "C"

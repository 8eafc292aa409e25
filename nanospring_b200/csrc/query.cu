// Candidate lookup on the device.  Replaces
//   MinHashReadFilter::getFilteredReads(kMer_t sketch[], results)   (src/ReadFilter.cpp:65-83)
//   BBHashMap::pushMatchesInVector                                  (src/BBHashMap.cpp:101-120)
// for a whole batch of query sketches at once: n exact probes per query, the id
// lists of all probes are gathered as (query, id) pairs, sorted, and an id is
// emitted when it occurs at least overlapSketchThreshold times.  Output per query
// is ascending and contains the query read itself, exactly like the reference's
// std::sort + upper_bound run counting.
#include "nsmh_internal.cuh"

namespace nsmh {

__device__ __forceinline__ uint64_t slot_hash_q(uint64_t key, uint32_t log2cap) {
    return (key * 0x9E3779B97F4A7C15ULL) >> (64 - log2cap);
}

// one thread per (query, hash): exact probe -> [begin, begin+cnt) in ids
__global__ void __launch_bounds__(256)
probe_kernel(const uint64_t *__restrict__ qsk, uint64_t items, uint32_t n, uint64_t cap,
             uint32_t log2cap, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ cnt,
             const uint32_t *__restrict__ begin, uint32_t *__restrict__ pbegin,
             uint32_t *__restrict__ pcnt) {
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < items;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t l = (uint32_t)(t % n);
        const uint64_t key = qsk[t];
        const uint64_t base = (uint64_t)l * (cap + 1);
        uint64_t s = base + cap;
        bool found = true;
        if (key != kEmptyKey) {
            uint64_t h = slot_hash_q(key, log2cap);
            for (;;) {
                s = base + h;
                uint64_t kk = keys[s];
                if (kk == key) break;
                if (kk == kEmptyKey) { found = false; break; }
                h = (h + 1) & (cap - 1);
            }
        }
        uint32_t c = found ? cnt[s] : 0u;
        pcnt[t] = c;
        pbegin[t] = c ? begin[s] : 0u;
    }
}

// one thread per (query, hash): copy the group's ids as (local query << 32 | id)
__global__ void __launch_bounds__(256)
gather_pairs_kernel(uint64_t item0, uint64_t items, uint32_t n, uint32_t q0,
                    const uint32_t *__restrict__ pbegin, const uint32_t *__restrict__ pcnt,
                    const uint64_t *__restrict__ poff, uint64_t pair0,
                    const uint32_t *__restrict__ ids, uint32_t id_base, uint64_t *__restrict__ pairs) {
    for (uint64_t t = item0 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < item0 + items;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = pcnt[t];
        if (!c) continue;
        const uint64_t q = t / n - q0;
        const uint32_t *src = ids + pbegin[t];
        uint64_t *dst = pairs + (poff[t] - pair0);
        for (uint32_t r = 0; r < c; ++r) dst[r] = (q << 32) | (uint64_t)(src[r] + id_base);
    }
}

// sorted pairs -> flag the first element of every run of length >= thr
__global__ void __launch_bounds__(256)
flag_runs_kernel(const uint64_t *__restrict__ pairs, uint64_t T, uint32_t thr, uint32_t q0,
                 uint8_t *__restrict__ flags, uint32_t *__restrict__ qcount) {
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < T;
         p += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = pairs[p];
        bool head = p == 0 || pairs[p - 1] != v;
        bool ok = head && (thr <= 1 || (p + thr - 1 < T && pairs[p + thr - 1] == v));
        flags[p] = ok;
        if (ok) atomicAdd(qcount + q0 + (uint32_t)(v >> 32), 1u);
    }
}

static int grid_for(uint64_t items, int sms) {
    uint64_t b = (items + 255) / 256;
    uint64_t cap = (uint64_t)sms * 16;
    return (int)(b < cap ? (b ? b : 1) : cap);
}

// Query nq device-resident sketches [nq][n] against the tables.  Result CSR in
// ws.out_off (u64 [nq+1]) / ws.out_ids (u32 [ws.last_total]).  Synchronises `s`.
int query_sketches_device(nsmh_ctx *c, QueryWs &ws, const uint64_t *d_qsketch, uint32_t nq,
                          cudaStream_t s) {
    Tables &T = c->tables;
    if (!T.built) return fail(NSMH_ESTATE, "query: tables not built (call nsmh_build)");
    const uint32_t n = c->n;
    const uint64_t items = (uint64_t)nq * n;
    ws.last_nq = nq;
    ws.last_total = 0;
    ws.last_pairs = 0;
    NSMH_TRY(ws.out_off.ensure(((size_t)nq + 1) * sizeof(uint64_t), s));
    NSMH_TRY(ws.qcount.ensure(((size_t)nq + 1) * sizeof(uint32_t), s));
    NSMH_CK(cudaMemsetAsync(ws.qcount.p, 0, ((size_t)nq + 1) * sizeof(uint32_t), s));
    if (nq == 0) {
        NSMH_CK(cudaMemsetAsync(ws.out_off.p, 0, sizeof(uint64_t), s));
        NSMH_CK(cudaStreamSynchronize(s));
        return NSMH_OK;
    }
    NSMH_TRY(ws.pbegin.ensure(items * sizeof(uint32_t), s));
    NSMH_TRY(ws.pcnt.ensure((items + 1) * sizeof(uint32_t), s));
    NSMH_TRY(ws.poff.ensure((items + 1) * sizeof(uint64_t), s));
    NSMH_TRY(ws.nsel.ensure(4 * sizeof(uint64_t), s));
    NSMH_CK(cudaMemsetAsync(ws.pcnt.as<uint32_t>() + items, 0, sizeof(uint32_t), s));

    probe_kernel<<<grid_for(items, c->num_sms), 256, 0, s>>>(
        d_qsketch, items, n, T.cap, T.log2cap, T.keys.as<uint64_t>(), T.cnt.as<uint32_t>(),
        T.begin.as<uint32_t>(), ws.pbegin.as<uint32_t>(), ws.pcnt.as<uint32_t>());
    ++ws.launches;
    NSMH_CK(cudaGetLastError());
    size_t tmp_bytes = 0;
    NSMH_CK(cub_exclusive_sum_u32_to_u64(nullptr, tmp_bytes, ws.pcnt.as<uint32_t>(),
                                         ws.poff.as<uint64_t>(), items + 1, s));
    NSMH_TRY(ws.cub_tmp.ensure(tmp_bytes, s));
    NSMH_CK(cub_exclusive_sum_u32_to_u64(ws.cub_tmp.p, tmp_bytes, ws.pcnt.as<uint32_t>(),
                                         ws.poff.as<uint64_t>(), items + 1, s));
    ws.launches += 2;
    uint64_t total_pairs = 0;
    NSMH_CK(cudaMemcpyAsync(&total_pairs, ws.poff.as<uint64_t>() + items, sizeof(uint64_t),
                            cudaMemcpyDeviceToHost, s));
    NSMH_CK(cudaStreamSynchronize(s));
    ws.last_pairs = total_pairs;

    // Batches of whole queries whose pairs fit the scratch budget.
    size_t free_b = 0, total_b = 0;
    NSMH_CK(cudaMemGetInfo(&free_b, &total_b));
    uint64_t budget_pairs = (uint64_t)((free_b + ws.pairs.cap + ws.pairs_alt.cap + ws.flags.cap) * 0.6 / 21.0);
    if (budget_pairs < (1u << 20)) budget_pairs = 1u << 20;
    std::vector<uint64_t> qstart;   // pair offset of every query (only needed when batching)
    std::vector<uint32_t> cuts{0, nq};
    if (total_pairs > budget_pairs) {
        qstart.resize((size_t)nq + 1);
        NSMH_CK(cudaMemcpy2DAsync(qstart.data(), sizeof(uint64_t), ws.poff.p, (size_t)n * sizeof(uint64_t),
                                  sizeof(uint64_t), nq, cudaMemcpyDeviceToHost, s));
        NSMH_CK(cudaStreamSynchronize(s));
        qstart[nq] = total_pairs;
        cuts.assign(1, 0);
        uint32_t q = 0;
        while (q < nq) {
            uint32_t e = q + 1;   // at least one query per batch
            while (e < nq && qstart[e + 1] - qstart[q] <= budget_pairs) ++e;
            cuts.push_back(e);
            q = e;
        }
    }

    // pass A: per batch gather + sort + flag (counts per query); results appended to out_ids
    // in query order because batches are processed in order and sorted by (query, id).
    // We do not know the output size in advance; select writes at most T entries per batch.
    // Two-step per batch: flag (gives exact per-query counts), then select into place.
    uint64_t out_total = 0;
    for (size_t bi = 0; bi + 1 < cuts.size(); ++bi) {
        const uint32_t q0 = cuts[bi], q1 = cuts[bi + 1];
        const uint64_t item0 = (uint64_t)q0 * n, bitems = (uint64_t)(q1 - q0) * n;
        uint64_t pair0, pair1;
        if (qstart.empty()) { pair0 = 0; pair1 = total_pairs; }
        else { pair0 = qstart[q0]; pair1 = qstart[q1]; }
        const uint64_t Tn = pair1 - pair0;
        if (Tn == 0) continue;
        NSMH_TRY(ws.pairs.ensure(Tn * sizeof(uint64_t), s));
        NSMH_TRY(ws.pairs_alt.ensure(Tn * sizeof(uint64_t), s));
        NSMH_TRY(ws.flags.ensure(Tn, s));
        gather_pairs_kernel<<<grid_for(bitems, c->num_sms), 256, 0, s>>>(
            item0, bitems, n, q0, ws.pbegin.as<uint32_t>(), ws.pcnt.as<uint32_t>(),
            ws.poff.as<uint64_t>(), pair0, T.ids.as<uint32_t>(), 0u, ws.pairs.as<uint64_t>());
        ++ws.launches;
        NSMH_CK(cudaGetLastError());
        int qbits = 1;
        while ((1ULL << qbits) < (uint64_t)(q1 - q0)) ++qbits;
        bool in_alt = false;
        tmp_bytes = 0;
        NSMH_CK(cub_sort_keys_u64(nullptr, tmp_bytes, ws.pairs.as<uint64_t>(), ws.pairs_alt.as<uint64_t>(),
                                  Tn, 0, 32 + qbits, in_alt, s));
        NSMH_TRY(ws.cub_tmp.ensure(tmp_bytes, s));
        NSMH_CK(cub_sort_keys_u64(ws.cub_tmp.p, tmp_bytes, ws.pairs.as<uint64_t>(),
                                  ws.pairs_alt.as<uint64_t>(), Tn, 0, 32 + qbits, in_alt, s));
        ws.launches += 2 + (32 + qbits + 7) / 8;   // histogram + onesweep passes (approximate)
        const uint64_t *sorted = in_alt ? ws.pairs_alt.as<uint64_t>() : ws.pairs.as<uint64_t>();
        flag_runs_kernel<<<grid_for(Tn, c->num_sms), 256, 0, s>>>(sorted, Tn, c->thr, q0,
                                                                  ws.flags.as<uint8_t>(),
                                                                  ws.qcount.as<uint32_t>());
        ++ws.launches;
        NSMH_CK(cudaGetLastError());
        // worst case every pair is selected
        NSMH_TRY(ws.out_ids.ensure((out_total + Tn) * sizeof(uint32_t), s, out_total * sizeof(uint32_t)));
        tmp_bytes = 0;
        NSMH_CK(cub_select_low32_flagged(nullptr, tmp_bytes, sorted, ws.flags.as<uint8_t>(),
                                         ws.out_ids.as<uint32_t>() + out_total, ws.nsel.as<uint64_t>(), Tn, s));
        NSMH_TRY(ws.cub_tmp.ensure(tmp_bytes, s));
        NSMH_CK(cub_select_low32_flagged(ws.cub_tmp.p, tmp_bytes, sorted, ws.flags.as<uint8_t>(),
                                         ws.out_ids.as<uint32_t>() + out_total, ws.nsel.as<uint64_t>(), Tn, s));
        ws.launches += 2;
        uint64_t nsel = 0;
        NSMH_CK(cudaMemcpyAsync(&nsel, ws.nsel.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        NSMH_CK(cudaStreamSynchronize(s));
        out_total += nsel;
    }
    tmp_bytes = 0;
    NSMH_CK(cub_exclusive_sum_u32_to_u64(nullptr, tmp_bytes, ws.qcount.as<uint32_t>(),
                                         ws.out_off.as<uint64_t>(), (size_t)nq + 1, s));
    NSMH_TRY(ws.cub_tmp.ensure(tmp_bytes, s));
    NSMH_CK(cub_exclusive_sum_u32_to_u64(ws.cub_tmp.p, tmp_bytes, ws.qcount.as<uint32_t>(),
                                         ws.out_off.as<uint64_t>(), (size_t)nq + 1, s));
    ws.launches += 2;
    NSMH_CK(cudaStreamSynchronize(s));
    ws.last_total = out_total;
    return NSMH_OK;
}

} // namespace nsmh

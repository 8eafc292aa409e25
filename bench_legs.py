"""Side measurements of bench.py: the other BASELINE.json configs and the parity checks that make a
multi-GPU number trustworthy.  Nothing here enters `value`; every leg reports its own time, its phase
split and a parity flag.

  parity   (a) sketch rows of a sample of reads against the CPU oracle (oracle/minhash_oracle.c);
           (b) candidate lists of a sample of queries against the DEFINITION evaluated with torch on the
               device-resident sketch matrix of ALL reads (ids whose rows agree with the query row in at
               least overlap-sketch-thr columns, ReadFilter.cpp:65-83) - independent of the tables;
           (c) N > 1: the CSR of the peer-memory path against the CSR of the NCCL all-gather path
               (full tables on every rank, what one GPU computes for the concatenated reads), element-wise
               on every rank.
  legs     rich          the headline reads at 2 % error: overlapping reads share sketch entries, the
                         lookup gathers real candidate lists (the headline's 10 % error leaves every read
                         alone with itself)
           c4_k15_n120   BASELINE configs[3] corner: 1 M reads, k=15, n=120 - the hash-table-bound regime
           c3            BASELINE configs[2]: 1 M reads (~10 Gbases) k=23 n=60, at N > 1 split by bases
           c5            BASELINE configs[4]: ultra-long reads (100 kb mean, 5 Gbases), split by bases
"""
import ctypes as C
import hashlib
import time

import numpy as np


def digest(*arrays):
    h = hashlib.blake2b(digest_size=8)
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def sample_rows(n_rows, want, seed):
    if n_rows == 0:
        return np.zeros(0, dtype=np.int64)
    rng = np.random.default_rng(seed)
    return np.unique(np.concatenate([[0, n_rows - 1], rng.integers(0, n_rows, size=min(want, n_rows))])).astype(np.int64)


def check_sketch_rows(f, d_bases, offsets, rows, k, n, rnd):
    """GPU sketch rows of `rows` (local read indices) against the CPU oracle on the same bases."""
    from oracle.oracle import Oracle
    from nanospring_b200 import shard
    import torch
    if rows.size == 0:
        return True, 0
    S = shard.sketches_as_tensor(f, offsets.size - 1)
    got = S[torch.from_numpy(rows).to(S.device)].cpu().numpy().view(np.uint64)
    pieces, offs = [], [0]
    for r in rows:
        lo, hi = int(offsets[r]), int(offsets[r + 1])
        pieces.append(d_bases[lo:hi].cpu().numpy())
        offs.append(offs[-1] + hi - lo)
    bases = np.concatenate(pieces + [np.zeros(1, np.uint8)])
    want = Oracle.get().sketch_all(bases, np.asarray(offs, dtype=np.uint64), k, n, rnd)
    return bool((got == want).all()), int(rows.size)


def check_candidates_by_definition(S_all, row_base, rows, thr, off, ids):
    """Candidate lists of the local rows `rows` (CSR off/ids, global ids) against the definition evaluated
    with torch on S_all [all reads][n] (int64 view of the u64 sketches)."""
    import torch
    thr = max(int(thr), 1)
    ok = True
    for r in rows:
        g = row_base + int(r)
        cnt = (S_all == S_all[g]).sum(dim=1)
        want = torch.nonzero(cnt >= thr).flatten().cpu().numpy().astype(np.uint32)
        got = ids[int(off[r]):int(off[r + 1])]
        if got.size != want.size or not (got == want).all():
            ok = False
            break
    return ok, int(len(rows))


class Leg:
    """One workload: synthetic reads of this rank's shard on the device + a filter (+ the peer-memory
    path when world > 1)."""

    def __init__(self, name, what, total_reads, mean_len, k, n, thr, rank, world, local_rank, rand_seed,
                 length_seed, synth_kw=None, dist_name="mixgamma"):
        import torch
        import nanospring_b200 as ns
        from nanospring_b200 import shard
        from nanospring_b200._lib import check, lib
        self.name, self.what = name, what
        self.k, self.n, self.thr = k, n, thr
        self.rank, self.world, self.local_rank = rank, world, local_rank
        lengths_all = ns.synth_lengths(total_reads, mean_len, seed=length_seed, dist=dist_name)
        off_all = np.zeros(total_reads + 1, dtype=np.uint64)
        off_all[1:] = np.cumsum(lengths_all, dtype=np.uint64)
        self.bounds = shard.shard_bounds_by_bases(off_all, world)       # equal BASES, not reads
        lo, hi = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.row_base = lo
        self.rows_per_rank = [int(x) for x in np.diff(self.bounds)]
        self.bases_per_rank = [int(off_all[self.bounds[r + 1]] - off_all[self.bounds[r]]) for r in range(world)]
        self.total_reads, self.total_bases_all = total_reads, int(off_all[-1])
        self.max_read = int(lengths_all.max())
        self.offsets = shard.local_offsets(off_all, lo, hi)
        self.bases = int(self.offsets[-1])
        self.params = ns.synth_params(**(synth_kw or {}))
        self.d_off = torch.from_numpy(self.offsets.astype(np.int64)).cuda()
        self.d_bases = torch.empty(self.bases + 64, dtype=torch.uint8, device="cuda")
        if hi > lo:
            check(lib().nsmh_synth_reads_device(local_rank, C.byref(self.params), lo, hi - lo,
                                                self.d_off.data_ptr(), self.d_bases.data_ptr()))
        self.rnd = ns.rand_from_seed(rand_seed, n)
        self.f = ns.MinHashReadFilter(device=local_rank)
        self.f.k, self.f.n, self.f.overlapSketchThreshold = k, n, thr
        self.f.randNumbers = self.rnd
        self.f._create()
        self.pf = shard.PeerPartitionedFilter(self.f, rank, world, self.rows_per_rank) if world > 1 else None
        torch.cuda.synchronize()

    def step(self):
        f = self.f
        f.load_device(self.d_bases.data_ptr(), self.d_off.data_ptr(), self.offsets.size - 1, self.bases)
        if self.pf is None:
            f.sketch()
            f.build()
            return f.queryAll(False, fetch=False)
        return self.pf.run(sketch=True)

    def csr(self, total):
        if self.pf is not None:
            return self.pf.result(self.offsets.size - 1, total)
        from nanospring_b200._lib import check, lib, u32p, u64p
        off = np.zeros(self.offsets.size, dtype=np.uint64)
        ids = np.zeros(max(int(total), 1), dtype=np.uint32)
        check(lib().nsmh_query_all_result(self.f._h, off.ctypes.data_as(u64p), ids.ctypes.data_as(u32p)))
        return off, ids[:int(total)]

    def close(self):
        if self.pf is not None:
            self.pf.shutdown()
        self.f.close()
        self.d_bases = self.d_off = None


def all_sketches(leg, dist):
    """[all reads][n] int64 on this rank's device (NCCL all-gather at N > 1)."""
    from nanospring_b200 import shard
    local = shard.sketches_as_tensor(leg.f, leg.offsets.size - 1)
    if leg.world == 1:
        return local
    return shard.all_gather_rows(local, leg.rows_per_rank)


def parity_of(leg, total, dist, sample=48, replicated_check=True):
    """The three checks of the module docstring; every rank takes part, the verdict is the AND over ranks."""
    import torch
    import nanospring_b200 as ns
    from nanospring_b200 import shard
    n_local = leg.offsets.size - 1
    out = {}
    off, ids = leg.csr(total)
    rows = sample_rows(n_local, sample, seed=7 + leg.rank)
    ok_s, m_s = check_sketch_rows(leg.f, leg.d_bases, leg.offsets, rows[:16], leg.k, leg.n, leg.rnd)
    leg.f.synchronize()
    S_all = all_sketches(leg, dist)
    ok_c, m_c = check_candidates_by_definition(S_all, leg.row_base, rows, leg.thr, off, ids)
    ok_r = True
    if leg.world > 1 and replicated_check:
        f2 = ns.MinHashReadFilter(device=leg.local_rank)
        f2.k, f2.n, f2.overlapSketchThreshold = leg.k, leg.n, leg.thr
        f2.randNumbers = leg.rnd
        f2._create()
        f2.load_device(leg.d_bases.data_ptr(), leg.d_off.data_ptr(), n_local, leg.bases)
        f2.sketch()
        f2.synchronize()
        keep = shard.gather_and_build(f2, n_local, leg.rows_per_rank, leg.rank)
        off2, ids2 = f2.queryAll(False, fetch=True)
        ok_r = bool(off.size == off2.size and (off == off2).all() and ids.size == ids2.size and (ids == ids2).all())
        del keep
        f2.close()
    del S_all
    ok = ok_s and ok_c and ok_r
    mine = {"rank": leg.rank, "csr_blake2b64": digest(off, ids), "ids": int(ids.size), "ok": bool(ok)}
    per_rank = [mine]
    if dist is not None:
        t = torch.tensor([1 if ok else 0], device="cuda", dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(int(t.item()))
        per_rank = [None] * leg.world
        dist.all_gather_object(per_rank, mine)
    out["ok"] = ok
    out["checks"] = {"sketch_rows_vs_cpu_oracle": m_s, "candidate_lists_vs_definition_on_all_sketches": m_c,
                     **({"csr_vs_nccl_all_gather_path": "element-wise, every rank"} if leg.world > 1 and replicated_check else {})}
    out["per_rank"] = per_rank
    return out


def run_leg(leg, steps, dist, with_parity=True):
    import torch
    ext = torch.cuda.ExternalStream(leg.f.stream(), device=leg.local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    total = 0
    for _ in range(2):
        total = leg.step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(steps):
        total = leg.step()
    e1.record(ext)
    barrier()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    st = leg.f.stats()
    all_ids = int(total)
    if dist is not None:
        t = torch.tensor([all_ids], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        all_ids = int(t.item())
    res = {"workload": leg.what, "k": leg.k, "num_hash": leg.n, "overlap_sketch_thr": leg.thr,
           "reads": leg.total_reads, "bases": leg.total_bases_all, "max_read_len": leg.max_read,
           "ms_per_step": ms / steps, "gbases_per_s": leg.total_bases_all * steps / (ms * 1e-3) / 1e9,
           "phases_last_step_rank0": {"pack_ms": st["pack_ms"], "sketch_ms": st["sketch_ms"], "build_ms": st["build_ms"],
                                      "lookup_ms": st["query_ms"]},
           "candidate_ids": all_ids, "ids_per_query": all_ids / max(leg.total_reads, 1),
           "gathered_ids_rank0": st["query_pairs"], "queries_beyond_warp_buffer_rank0": st["query_heavy"],
           "queries_global_sort_rank0": st["query_sorted"]}
    if leg.world > 1:
        mean = sum(leg.bases_per_rank) / leg.world
        res["shards"] = {"split": "prefix sum of read lengths (equal bases)", "reads_per_rank": leg.rows_per_rank,
                         "bases_per_rank": leg.bases_per_rank,
                         "base_imbalance_max_over_mean": max(leg.bases_per_rank) / mean if mean else None}
        if leg.pf is not None:
            res["stage_ms_rank0"] = {k2: round(v, 3) for k2, v in leg.pf.last_ms.items()}
    if with_parity:
        t0 = time.perf_counter()
        res["parity"] = parity_of(leg, total, dist)
        res["parity"]["seconds"] = round(time.perf_counter() - t0, 2)
    return res


# ------------------------------------------------------------------------------------------
def online_leg(f, host_rd, n_windows=256, window_len=10_000, thread_counts=(1, 8, 32), ref_filter=None):
    """the production query, one consensus window per call (Consensus.cpp:168-191), through
    nsmh_query_string from 1 / 8 / 32 host threads; latency percentiles per call and queries per second.
    ref_filter: the reference's own MinHashReadFilter (oracle/_ref) for the same windows on one host core."""
    import threading
    import torch
    from nanospring_b200 import shard
    from nanospring_b200._lib import lib, u32p
    from oracle.oracle import Oracle
    L = lib()
    rng = np.random.default_rng(5)
    lens = np.diff(host_rd.offsets.astype(np.int64))
    long_reads = np.flatnonzero(lens >= window_len)
    pick = rng.choice(long_reads, size=min(n_windows, long_reads.size), replace=False)
    windows = []
    for r in pick:
        s0 = int(host_rd.offsets[r]) + int(rng.integers(0, lens[r] - window_len + 1))
        windows.append(host_rd.bases[s0:s0 + window_len].tobytes())

    def run_thread(idx, reps, out_lat):
        buf = np.empty(4096, dtype=np.uint32)
        cnt = C.c_size_t(0)
        p = buf.ctypes.data_as(u32p)
        for _ in range(reps):
            for i in idx:
                w = windows[i]
                t0 = time.perf_counter_ns()
                rc = L.nsmh_query_string(f._h, w, len(w), p, 4096, C.byref(cnt))
                out_lat.append(time.perf_counter_ns() - t0)
                if rc != 0:
                    raise RuntimeError("nsmh_query_string failed")

    res = {"what": f"{len(windows)} windows of {window_len} bases cut from the reads, nsmh_query_string per window "
                   f"(= ReadFilter::getFilteredReads), tables over all {host_rd.numReads} reads",
           "window_len": window_len, "threads": {}}
    for nt in thread_counts:
        # warm-up with the same number of threads: every concurrent caller gets its own workspace (stream, mapped
        # buffers) the first time the pool runs dry
        ths = [threading.Thread(target=run_thread, args=(range(t, len(windows), nt), 1, [])) for t in range(nt)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        lats = [[] for _ in range(nt)]
        ths = [threading.Thread(target=run_thread, args=(range(t, len(windows), nt), 2, lats[t])) for t in range(nt)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        wall = time.perf_counter() - t0
        lat = np.concatenate([np.asarray(x, dtype=np.float64) for x in lats]) / 1e3
        res["threads"][str(nt)] = {"p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99)),
                                   "queries_per_s": lat.size / wall}
    res["note"] = ("host threads are Python threads here (ctypes releases the GIL inside the call, the bookkeeping between calls "
                   "is serialised): the C++ drop-in test under OpenMP reports the same per-call latency, profiles/r2_dropin_latency_*.log")
    # parity of the fused online kernel: the candidates of a few windows against the definition evaluated on the
    # sketches of all reads, with the window's sketch computed by the CPU oracle
    S_all = shard.sketches_as_tensor(f, host_rd.numReads)
    orc = Oracle.get()
    ok = True
    for w in windows[:12]:
        wb = np.frombuffer(w, dtype=np.uint8)
        sk = orc.sketch_all(np.concatenate([wb, np.zeros(1, np.uint8)]), np.asarray([0, wb.size], dtype=np.uint64), f.k, f.n,
                            np.asarray(f.randNumbers, dtype=np.uint64))
        q = torch.from_numpy(sk.view(np.int64)).to(S_all.device)
        want = torch.nonzero((S_all == q[0]).sum(dim=1) >= max(int(f.overlapSketchThreshold), 1)).flatten().cpu().numpy().astype(np.uint32)
        got = f.getFilteredReads(w)
        ok = ok and got.size == want.size and bool((got == want).all())
    res["parity"] = {"ok": bool(ok), "windows_checked": min(12, len(windows)),
                     "against": "definition on the sketches of all reads, window sketch by the CPU oracle"}
    if ref_filter is not None:
        lat = []
        out = np.zeros(ref_filter.N + 1, dtype=np.uint32)
        po = out.ctypes.data_as(u32p)
        for w in windows[:64]:
            t0 = time.perf_counter_ns()
            ref_filter.ref.lib.nsref_query_string(ref_filter.h, w, len(w), po, out.size)
            lat.append(time.perf_counter_ns() - t0)
        lat = np.asarray(lat, dtype=np.float64) / 1e3
        res["cpu_reference"] = {"p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99)),
                                "what": f"the reference's getFilteredReads(string) on 64 of the windows, one host core, "
                                        f"tables over {ref_filter.N} reads (the window's sketch dominates its cost)"}
    return res

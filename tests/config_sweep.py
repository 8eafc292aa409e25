#!/usr/bin/env python
"""BASELINE.json configs at FULL size on one B200, with size-independent parity checks.

    python tests/config_sweep.py [--configs c2,c3,c4,c5] [--out profiles/x.jsonl]

  c2  100 000 reads, ~10 kb mean, 10 % error, 50 Mb genome, k=23 n=60 thr=6        (~1 Gbase)
  c2lo  the same at 2 % error: rich candidate sets for the lookup checks
  c3  1 000 000 reads (~10 Gbases), k=23 n=60 thr=6
  c4  parameter sweep n in {30,60,120} x k in {15,23,31} (thr = n/10) on the c3 reads
  c5  ultra-long reads: 50 000 reads, ~100 kb mean (~5 Gbases), k=23 n=60 thr=6

Reads are generated on the device (nsmh_synth_reads_device); one line of JSON per run: the
device-resident rate (2-bit pack -> sketch -> build -> bulk lookup, CUDA events on the engine's
stream), the per-stage times, and the outcome of the checks.  Checks (no CPU pass over 10 Gbases
is needed for any of them):
  * the brute-force kernel (every k-mer x every hash, the reference's operation count) and the
    filter kernel give bit-identical sketch matrices;
  * the sketch rows of a random sample of reads equal the CPU oracle's (oracle/minhash_oracle.c);
  * the candidate lists of a random sample of reads equal an independent evaluation of the
    reference's definition (ReadFilter.cpp:65-83) with torch integer ops on the full sketch
    matrix: {j : #{l : S[j][l] == S[i][l]} >= thr}, ascending;
  * every read is its own candidate, lists ascend strictly, the forward relation is symmetric.
tests/test_gpu_fullsize.py runs the same function with asserts."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

RAND_SEED = 20261017


class DeviceReads:
    """A synthetic read set resident on the device (ASCII bases + u64 offsets)."""

    def __init__(self, n_reads, mean_len, p_err=0.10, genome_len=50_000_000, seed=1, device=0):
        import torch
        import nanospring_b200 as ns
        from nanospring_b200._lib import check, lib
        self.lengths = ns.synth_lengths(n_reads, mean_len, seed=seed)
        self.offsets = np.zeros(n_reads + 1, dtype=np.uint64)
        self.offsets[1:] = np.cumsum(self.lengths, dtype=np.uint64)
        self.total = int(self.offsets[-1])
        self.n_reads = n_reads
        self.params = ns.synth_params(genome_len=genome_len, p_ins=0.3 * p_err, p_del=0.3 * p_err, p_sub=0.4 * p_err)
        self.d_off = torch.from_numpy(self.offsets.astype(np.int64)).cuda(device)
        self.d_bases = torch.empty(self.total + 64, dtype=torch.uint8, device=f"cuda:{device}")
        t0 = time.perf_counter()
        check(lib().nsmh_synth_reads_device(device, C.byref(self.params), 0, n_reads, self.d_off.data_ptr(),
                                            self.d_bases.data_ptr()))
        self.synth_s = time.perf_counter() - t0

    def host_read(self, i):
        a, b = int(self.offsets[i]), int(self.offsets[i + 1])
        return self.d_bases[a:b].cpu().numpy()


def csr_tensors(f, n_reads, total):
    import torch
    from nanospring_b200 import shard
    from nanospring_b200._lib import check, lib
    p_off, p_ids = C.c_void_p(), C.c_void_p()
    check(lib().nsmh_query_all_device_ptrs(f._h, C.byref(p_off), C.byref(p_ids)))
    dev = f"cuda:{f.device}"
    off = torch.as_tensor(shard.DeviceAlias(p_off.value, n_reads + 1), device=dev)
    ids = torch.as_tensor(shard.DeviceAlias(p_ids.value, max(total, 1), "<i4"), device=dev)[:total]
    return off, ids


def run_config(name, reads, k, n, thr, steps=3, warmup=2, sample=64, check_brute=True, log=print):
    """Times the device-resident step on `reads` and runs the parity checks.  Returns a dict;
    result["ok"] is the conjunction of all checks, result["checks"] names each of them."""
    import torch
    import nanospring_b200 as ns
    from nanospring_b200 import shard
    from oracle.oracle import Oracle
    dev = reads.d_bases.device
    rnd = ns.rand_from_seed(RAND_SEED, n)
    f = ns.MinHashReadFilter(device=dev.index or 0)
    f.k, f.n, f.overlapSketchThreshold, f.randNumbers = k, n, thr, rnd
    f._create()
    ext = torch.cuda.ExternalStream(f.stream(), device=dev)

    def step():
        f.load_device(reads.d_bases.data_ptr(), reads.d_off.data_ptr(), reads.n_reads, reads.total)
        f.sketch()
        f.build()
        return f.queryAll(False, fetch=False)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(steps):
        total = step()
    e1.record(ext)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = f.stats()
    N = reads.n_reads
    res = {"config": name, "reads": N, "bases": reads.total, "max_read": int(reads.lengths.max()), "k": k, "num_hash": n,
           "overlap_sketch_thr": thr, "ms_per_step": ms, "gbases_per_s": reads.total / (ms * 1e-3) / 1e9,
           "pack_ms": st["pack_ms"], "sketch_ms": st["sketch_ms"], "sketch_main_kernel_ms": st["sketch_main_ms"],
           "build_ms": st["build_ms"], "query_ms": st["query_ms"], "candidate_ids": int(total),
           "fixups_per_step": st["sketch_fixups"] / (steps + warmup), "checks": {}}
    ck = res["checks"]

    S = shard.sketches_as_tensor(f, N)                                     # int64 [N, n] on the device
    # -- brute-force kernel == filter kernel, bit for bit
    if check_brute:
        g = ns.MinHashReadFilter(device=dev.index or 0)
        g.k, g.n, g.overlapSketchThreshold, g.randNumbers, g.sketchMode = k, n, thr, rnd, 1
        g._create()
        g.load_device(reads.d_bases.data_ptr(), reads.d_off.data_ptr(), N, reads.total)
        t0 = time.perf_counter()
        g.sketch()
        g.synchronize()
        res["brute_sketch_ms"] = 1e3 * (time.perf_counter() - t0)
        ck["brute_equals_filter"] = bool(torch.equal(shard.sketches_as_tensor(g, N), S))
        g.close()
    # -- sampled sketch rows == CPU oracle
    rng = np.random.default_rng(7)
    order = np.argsort(reads.lengths)
    pick = np.unique(np.concatenate([rng.choice(N, size=min(sample, N), replace=False), order[:4], order[-2:]]))
    orc = Oracle.get()
    ok = True
    for i in pick.tolist():
        s = reads.host_read(i)
        want = orc.sketch_all(s, np.array([0, s.size], dtype=np.uint64), k, n, rnd)[0]
        ok &= bool((S[i].cpu().numpy().view(np.uint64) == want).all())
    ck["sampled_sketches_equal_oracle"] = ok
    # -- CSR properties on the device
    off, ids = csr_tensors(f, N, int(total))
    counts = off[1:] - off[:-1]
    owner = torch.repeat_interleave(torch.arange(N, device=dev), counts)
    ids64 = ids.to(torch.int64)
    ck["self_is_candidate_once"] = bool((torch.bincount(owner[owner == ids64], minlength=N) == 1).all())
    if ids64.numel() > 1:
        d = ids64[1:] - ids64[:-1]
        same_row = owner[1:] == owner[:-1]
        ck["lists_strictly_ascending"] = bool((d[same_row] > 0).all())
    else:
        ck["lists_strictly_ascending"] = True
    fwd = torch.sort(owner * N + ids64).values
    bwd = torch.sort(ids64 * N + owner).values
    ck["relation_symmetric"] = bool(torch.equal(fwd, bwd))
    # -- sampled candidate lists == the definition evaluated with torch on the full matrix
    ok = True
    widest = int(torch.argmax(counts).item())
    for i in np.unique(np.append(pick, widest)).tolist():
        shared = (S == S[i]).sum(dim=1)
        want = torch.nonzero(shared >= thr).flatten()
        got = ids64[int(off[i].item()):int(off[i + 1].item())]
        ok &= bool(want.numel() == got.numel() and torch.equal(want, got))
    ck["sampled_candidates_equal_definition"] = ok
    res["max_candidates_per_read"] = int(counts.max().item())
    res["ok"] = all(ck.values())
    f.close()
    log(json.dumps(res))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c2,c2lo,c3,c4,c5")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    want = set(args.configs.split(","))
    out = open(args.out, "w") if args.out else None

    def log(line):
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()

    bad = 0
    if "c2" in want:
        r = DeviceReads(100_000, 10_000, 0.10, seed=1000)
        bad += not run_config("c2: 100k reads, 10 kb mean, 10% error", r, 23, 60, 6, log=log)["ok"]
        del r
    if "c2lo" in want:
        r = DeviceReads(100_000, 10_000, 0.02, seed=1000)
        bad += not run_config("c2lo: 100k reads, 10 kb mean, 2% error (rich candidate sets)", r, 23, 60, 6, log=log)["ok"]
        del r
    if "c3" in want or "c4" in want:
        r = DeviceReads(1_000_000, 10_000, 0.10, seed=2000)
        if "c3" in want:
            bad += not run_config("c3: 1M reads, 10 kb mean, 10% error", r, 23, 60, 6, log=log)["ok"]
        if "c4" in want:
            for n in (30, 60, 120):
                for k in (15, 23, 31):
                    bad += not run_config(f"c4: 1M reads, sweep n={n} k={k}", r, k, n, n // 10, sample=24,
                                          check_brute=(k == 15 or n == 120), log=log)["ok"]
        del r
    if "c5" in want:
        r = DeviceReads(50_000, 100_000, 0.10, seed=3000)
        bad += not run_config("c5: 50k ultra-long reads, 100 kb mean, 10% error", r, 23, 60, 6, log=log)["ok"]
        del r
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()

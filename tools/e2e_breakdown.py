#!/usr/bin/env python
"""Where the end-to-end step goes: host wall time of every call of the host-buffer path (each followed by a
stream synchronize, so the pieces add up to more than the pipelined call), next to the pipelined
nsmh_initialize_* call that bench.py's e2e legs time.  bench.py's workload, one GPU.

    python tools/e2e_breakdown.py [--reads 100000] [--reps 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=100000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    sys.argv = sys.argv[:1]
    import ctypes as C

    import torch

    import bench
    import nanospring_b200 as ns
    from nanospring_b200._lib import check, lib

    lengths = bench.shard_lengths(0)[:args.reads]
    offsets = np.zeros(lengths.size + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lengths, dtype=np.uint64)
    total = int(offsets[-1])
    params = ns.synth_params(genome_len=bench.GENOME_LEN)
    d_off = torch.from_numpy(offsets.astype(np.int64)).cuda()
    d_bases = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
    check(lib().nsmh_synth_reads_device(0, C.byref(params), 0, lengths.size, d_off.data_ptr(), d_bases.data_ptr()))
    h_bases = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    h_bases.copy_(d_bases[:total])
    torch.cuda.synchronize()
    rd = ns.ReadData(h_bases.numpy(), offsets)
    h_packed, len32 = bench.dnabitset_host(d_bases, offsets)
    del d_bases

    f = ns.MinHashReadFilter(device=0)
    f.k, f.n, f.overlapSketchThreshold = bench.K, bench.NHASH, bench.THR
    f.randNumbers = ns.rand_from_seed(bench.RAND_SEED, bench.NHASH)

    def timed(fn):
        t = time.perf_counter()
        out = fn()
        f.synchronize()
        return (time.perf_counter() - t) * 1e3, out

    def median(rows):
        return {k: round(float(np.median([r[k] for r in rows])), 3) for k in rows[0]}

    out = {"reads": int(lengths.size), "gbases": total / 1e9}
    for name, load, init in (("ascii", lambda: f.load(rd), lambda: f.initialize(rd)),
                             ("dnabitset", lambda: f.load_dnabitset(h_packed, len32), lambda: f.initialize_dnabitset(h_packed, len32))):
        rows, piped = [], []
        for rep in range(args.reps + 2):
            r = {}
            r["load_ms"], _ = timed(load)
            r["sketch_ms"], _ = timed(f.sketch)
            r["build_ms"], _ = timed(f.build)
            r["query_fetch_ms"], _ = timed(lambda: f.queryAll(False, fetch=True))
            r["sum_ms"] = sum(r.values())
            t = time.perf_counter()
            init()
            t1 = time.perf_counter()
            f.queryAll(False, fetch=True)
            t2 = time.perf_counter()
            p = {"initialize_returns_ms": (t1 - t) * 1e3, "query_fetch_ms": (t2 - t1) * 1e3, "total_ms": (t2 - t) * 1e3}
            if rep >= 2:
                rows.append(r)
                piped.append(p)
        h2d = total if name == "ascii" else int(h_packed.nbytes)
        out[name] = {"separate_calls": median(rows), "pipelined": median(piped), "h2d_bytes": h2d,
                     "h2d_alone_ms_at_measured_rate": None}
        # the copy alone, same chunks: what PCIe allows
        src = h_bases if name == "ascii" else torch.from_numpy(h_packed)
        dst = torch.empty(h2d, dtype=torch.uint8, device="cuda")
        best = 1e9
        for _ in range(4):
            torch.cuda.synchronize()
            t = time.perf_counter()
            dst.copy_(src[:h2d], non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t) * 1e3)
        out[name]["h2d_alone_ms_at_measured_rate"] = round(best, 3)
        out[name]["h2d_gb_per_s"] = round(h2d / best / 1e6, 2)
        del dst
    print(json.dumps(out))


if __name__ == "__main__":
    main()

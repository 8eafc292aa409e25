#!/usr/bin/env python
"""A/B of the candidate lookup on BASELINE config C4's small-k corner (1 M reads, ~10 Gbases, k=15):
bulk forward lookup with the counting-filter tier (csrc/query_kernels.cuh) and with every overflowing
query sent to the global radix sort (NSMH_MID_TIER=0).  Prints one JSON line per run; the CSRs of
the two runs must be identical (same offsets, same ids).

    python tools/mid_tier_ab.py [--reads 1000000] [--cfg 15,120,12 15,60,6]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))       # config_sweep.py: the device-resident read generator


def main():
    import torch
    import nanospring_b200 as ns
    from config_sweep import RAND_SEED, DeviceReads, csr_tensors
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--cfg", nargs="*", default=["15,120,12", "15,60,6"])
    args = ap.parse_args()
    reads = DeviceReads(args.reads, 10_000)
    for cfg in args.cfg:
        k, n, thr = (int(x) for x in cfg.split(","))
        keep = {}
        for tier in ("1", "0"):
            os.environ["NSMH_MID_TIER"] = tier
            f = ns.MinHashReadFilter(device=0)
            f.k, f.n, f.overlapSketchThreshold, f.randNumbers = k, n, thr, ns.rand_from_seed(RAND_SEED, n)
            f._create()
            f.load_device(reads.d_bases.data_ptr(), reads.d_off.data_ptr(), reads.n_reads, reads.total)
            f.sketch()
            f.build()
            q_ms = []
            for _ in range(3):
                total = f.queryAll(False, fetch=False)
                q_ms.append(f.stats()["query_ms"])
            st = f.stats()
            off, ids = csr_tensors(f, reads.n_reads, int(total))
            keep[tier] = (off.clone(), ids.clone())
            print(json.dumps({"k": k, "num_hash": n, "thr": thr, "reads": reads.n_reads, "bases": reads.total,
                              "counting_filter_tier": tier == "1", "query_ms": round(min(q_ms), 3),
                              "query_ms_all": [round(x, 3) for x in q_ms], "gathered_ids": int(st["query_pairs"]),
                              "queries_over_warp_buffer": int(st["query_heavy"]),
                              "queries_global_sort": int(st["query_sorted"]), "candidate_ids": int(total)}), flush=True)
            f.close()
        same = bool(torch.equal(keep["1"][0], keep["0"][0]) and torch.equal(keep["1"][1], keep["0"][1]))
        print(json.dumps({"k": k, "num_hash": n, "thr": thr, "identical_csr": same}), flush=True)
        del keep
    os.environ.pop("NSMH_MID_TIER", None)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Every kernel family of the library once, on small inputs, for a run under compute-sanitizer:

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize_smoke.py
    compute-sanitizer --tool synccheck --error-exitcode 1 python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_smoke.py

(racecheck reports the shared-memory hazards of sketch_filter_kernel's phase 2 by design: smaller minima are
written with plain stores, re-read after a warp barrier and repaired with an atomic - sketch_kernels.cuh.)
numpy + ctypes only (no torch: the sanitizer instruments every kernel of the process).  Results are compared
with the oracle, so a pass is also a parity run."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import nanospring_b200 as ns
    from conftest import pack_dnabitset
    from oracle.oracle import Oracle

    orc = Oracle.get()
    os.environ["NSMH_LOAD_CHUNK_BYTES"] = "40000"           # several chunks: the pipelined loaders' range calls
    for k, n, thr, reads, mean, glen in ((23, 60, 6, 300, 2500, 60_000), (15, 120, 12, 1500, 1200, 40_000), (31, 7, 1, 60, 900, 9_000)):
        lengths = ns.synth_lengths(reads, mean, seed=k)
        lengths[:6] = [0, 3, k - 1, k, 40000, 17]
        rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=glen, genome_seed=k, read_seed=n, p_ins=0.01, p_del=0.01, p_sub=0.02))
        rnd = ns.rand_from_seed(7 * k, n)
        want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
        T = orc.build_tables(want)
        fwd = T.query_all(rd.bases, rd.offsets, want, k, rnd, thr, 0)
        rc = T.query_all(rd.bases, rd.offsets, want, k, rnd, thr, 1)
        f = ns.MinHashReadFilter(device=0)
        f.k, f.n, f.overlapSketchThreshold, f.randNumbers = k, n, thr, rnd
        f.initialize(rd)                                     # pack + filter/fix-up kernels + tables, pipelined
        assert (f.sketches() == want).all()
        for got, ref in ((f.queryAll(False), fwd), (f.queryAll(True), rc)):
            assert (got[0] == ref[0]).all() and (got[1] == ref[1]).all()
        f.readFlags()
        f.queryAll(False, fetch=False)
        f.queryAllDrop(3)
        for i in (4, 7, 11):                                 # the fused online kernel and the general path
            s = rd.getRead(i)
            for q in (s, s[: len(s) // 2], ns.reverse_complement(s)):
                assert (f.getFilteredReads(q) == T.query_string(q, k, rnd, thr)).all()
        f.getFilteredReadsBatch([rd.getRead(8), rd.getRead(9)[:100], b""])
        packed, len32 = pack_dnabitset(rd.bases, rd.offsets)
        f.initialize_dnabitset(packed, len32)                # the DnaBitset re-layout kernel
        assert (f.sketches() == want).all()
        f.sketchMode = 1                                     # brute-force kernel
        f.load(rd)
        f.sketch()
        assert (f.sketches() == want).all()
        f.close()
        print(f"k={k} n={n}: ok ({rd.numReads} reads, {int(fwd[0][-1])} candidates)", flush=True)
    # FASTQ ingest
    recs = []
    rng = np.random.default_rng(1)
    for i in range(200):
        s = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=int(rng.integers(0, 700))))
        recs.append(b"@r%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")
    g = ns.GpuReadData(0)
    g.loadFromText(b"".join(recs))
    assert g.getNumReads() == 200
    g.getRead(5)
    g.close()
    print("fastq: ok", flush=True)


if __name__ == "__main__":
    main()

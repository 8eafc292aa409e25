"""Generates the committed fixtures in tests/golden/ from the UNMODIFIED reference
(oracle/_ref/libnsref.so, built by oracle/Makefile from /root/reference).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

Outputs
  c1_reads.npz      the reference's CI input util/test_file.fastq.gz, bases only,
                    2-bit packed exactly like DnaBitset (dnaToBits.cpp:11-36:
                    4 bases/byte, first base in bits 7..6, per-read byte aligned)
                    + read lengths.  Alphabet of that file is exactly {A,C,G,T},
                    so the packing is lossless.
  c1_golden.json    checksums of the reference's sketches / candidate sets on it
                    for three (seed,k,n,thr) settings.
  edge_golden.npz   a small hand-made read set (empty reads, len k-2..k+1, N and
                    lower-case bytes, duplicates, one long read) with the
                    reference's full outputs for several settings.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle, RefLib, load_fastq, reads_to_buffers  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SETTINGS = [(20261017, 23, 60, 6), (1, 15, 30, 3), (7, 31, 120, 12)]


def pack_dnabitset(reads):
    chunks = []
    for r in reads:
        a = np.frombuffer(r, dtype=np.uint8)
        code = (a & 2) | ((a & 4) >> 2)
        pad = (-len(code)) % 4
        code = np.concatenate([code, np.zeros(pad, np.uint8)]).reshape(-1, 4)
        chunks.append((code[:, 0] << 6 | code[:, 1] << 4 | code[:, 2] << 2 | code[:, 3]).astype(np.uint8))
    return np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)


def edge_reads(rng):
    def rnd(L):
        return bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=L).tobytes())
    reads = [b"", b"A", b"", b"AC"]
    for L in range(0, 40):
        reads.append(rnd(L))
    genome = rnd(30000)
    for i in range(120):                      # overlapping error-free reads -> real candidates
        s = int(rng.integers(0, 27000))
        reads.append(genome[s:s + int(rng.integers(200, 3000))])
    reads.append(genome[:20000])              # one long read
    reads.append(genome[100:900])
    reads.append(genome[100:900])             # exact duplicate
    reads.append(b"ACGTNNNNACGTacgtnnnnRYKM" * 20)   # non-ACGT bytes go through the bit trick
    reads.append(b"A" * 500)
    reads.append(b"AT" * 300)
    reads.append(b"N" * 64)
    reads.append(rnd(22) + b"\x00\xff\x7f")
    for i in range(30):
        reads.append(rnd(int(rng.integers(14, 64))))
    return reads


def main():
    orc, ref = Oracle.get(), RefLib.get()
    # ---- C1 ----------------------------------------------------------------
    reads = load_fastq("/root/reference/util/test_file.fastq.gz")
    alphabet = set(b"".join(reads))
    assert alphabet == set(b"ACGT"), alphabet
    lengths = np.array([len(r) for r in reads], dtype=np.uint32)
    np.savez_compressed(os.path.join(OUT, "c1_reads.npz"), packed=pack_dnabitset(reads),
                        lengths=lengths)
    bases, offsets = reads_to_buffers(reads)
    gold = {"source": "util/test_file.fastq.gz", "num_reads": len(reads),
            "num_bases": int(lengths.sum()), "settings": []}
    for seed, k, n, thr in SETTINGS:
        rnd = ref.rand_from_seed(seed, n)
        rf = ref.create(bases, offsets, k, n, thr, rnd)
        sk = rf.sketches()
        off, ids = rf.query_all(0)
        offr, idsr = rf.query_all(1)
        d = np.diff(off.astype(np.int64))
        gold["settings"].append({
            "seed": seed, "k": k, "n": n, "thr": thr,
            "rand_first": "%016x" % rnd[0], "rand_last": "%016x" % rnd[-1],
            "fnv_sketches": "%016x" % orc.fnv_u64(sk.ravel()),
            "sketch_read4_first3": ["%016x" % x for x in sk[4][:3]],
            "fwd_total": int(off[-1]), "fwd_fnv": "%016x" % orc.fnv_csr(off, ids),
            "fwd_cands_read4": [int(x) for x in ids[off[4]:off[5]]],
            "fwd_max": int(d.max()), "fwd_argmax": int(d.argmax()),
            "rc_total": int(offr[-1]), "rc_fnv": "%016x" % orc.fnv_csr(offr, idsr),
        })
        rf.close()
        print(gold["settings"][-1])
    with open(os.path.join(OUT, "c1_golden.json"), "w") as f:
        json.dump(gold, f, indent=1)

    # ---- edge set ----------------------------------------------------------
    rng = np.random.default_rng(12345)
    er = edge_reads(rng)
    eb, eo = reads_to_buffers(er)
    out = {"bases": eb, "offsets": eo}
    cfgs = [(11, 23, 60, 6), (12, 15, 30, 3), (13, 31, 120, 12), (14, 4, 8, 1), (15, 16, 33, 2),
            (16, 17, 64, 60), (17, 1, 5, 5)]
    out["cfgs"] = np.array(cfgs, dtype=np.int64)
    for ci, (seed, k, n, thr) in enumerate(cfgs):
        rnd = ref.rand_from_seed(seed, n)
        rf = ref.create(eb, eo, k, n, thr, rnd)
        out[f"rand_{ci}"] = rnd
        out[f"sketches_{ci}"] = rf.sketches()
        out[f"fwd_off_{ci}"], out[f"fwd_ids_{ci}"] = rf.query_all(0)
        out[f"rc_off_{ci}"], out[f"rc_ids_{ci}"] = rf.query_all(1)
        rf.close()
    np.savez_compressed(os.path.join(OUT, "edge_golden.npz"), **out)
    print("edge set:", len(er), "reads", eb.size, "bases")


if __name__ == "__main__":
    main()

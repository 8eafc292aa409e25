"""Generates tests/golden/prefilter_golden.npz from the UNMODIFIED reference consensus
translation unit (oracle/_ref/libnsref_consensus.so = /root/reference/src/Consensus.cpp compiled
by oracle/Makefile): isRepetitive[] of Consensus::initialize (Consensus.cpp:426-442, i.e.
checkRepetitive, :405-424) for a hand-made read set that brackets the 0.7 threshold.

Run in the build container only (needs /root/reference):
    python tests/golden/make_prefilter_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import RefConsensus, reads_to_buffers  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def prefilter_reads(seed=11):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)

    def rnd(L):
        return rng.choice(acgt, size=L).tobytes()

    def tandem(unit, L, p_mut):
        """Tandem repeat of `unit` to length L with a fraction p_mut of substituted bases."""
        a = np.frombuffer((unit * (L // len(unit) + 1))[:L], np.uint8).copy()
        hit = rng.random(L) < p_mut
        a[hit] = rng.choice(acgt, size=int(hit.sum()))
        return a.tobytes()

    reads = [b"", b"A", b"AC", b"ACG", b"ACGT", b"ACGTA", b"ACGTAC", b"ACGTACG", b"AAAAAA", b"AAAAAAA"]
    for L in range(0, 72):                                   # every length around the wrap / word edges
        reads.append(rnd(L))
        reads.append(tandem(b"ACG", L, 0.0))
    for period in range(1, 9):                               # periods 1..6 are caught, 7 and 8 are not
        unit = rnd(period)
        for L in (31, 32, 33, 100, 511, 512, 513, 517, 518, 519, 1030, 5000):
            for p in (0.0, 0.1, 0.17, 0.2, 0.22, 0.3, 0.5):  # ~ (1-p)^2 + p/4-ish agreement: brackets 0.7
                reads.append(tandem(unit, L, p))
    for L in (8191, 8192, 8193, 8200, 20000, 70000):         # longer than one warp visit of the kernel
        reads.append(rnd(L))
        reads.append(tandem(b"AT", L, 0.19))
        reads.append(tandem(b"ACGTT", L, 0.21))
    reads.append(b"ACGTNNNNACGTacgtnnnnRYKM" * 20)            # non-ACGT bytes go through the 2-bit store
    reads.append(b"N" * 64)
    reads += [b"", b"", b"T", b""]                            # empty reads between others
    order = rng.permutation(len(reads))                       # short and long reads interleaved
    return [reads[i] for i in order]


def main():
    reads = prefilter_reads()
    bases, offsets = reads_to_buffers(reads)
    rep = RefConsensus.get().is_repetitive(bases, offsets, threads=8)
    np.savez_compressed(os.path.join(OUT, "prefilter_golden.npz"), bases=bases, offsets=offsets, repetitive=rep)
    print(f"{len(reads)} reads, {int(offsets[-1])} bases, {int(rep.sum())} repetitive")


if __name__ == "__main__":
    main()

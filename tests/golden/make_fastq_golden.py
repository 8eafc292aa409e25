"""Generates tests/golden/fastq_golden.npz from the UNMODIFIED reference loader
(oracle/_ref/libnsref_readdata.so = /root/reference/src/ReadData.cpp + dnaToBits.cpp, built by
`make -C oracle readdata`): for every text of tests/fastq_cases.py (hand-made corner cases +
seeded random FASTQ) the reads ReadData::loadFromFile + getRead give back, in the CLI's low-memory
mode (main.cpp:40), from a plain file and from a gzip file; plus checksums of the reference's CI
input util/test_file.fastq.gz as loaded by the reference itself.

Run in the build container only (needs /root/reference):
    python tests/golden/make_fastq_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fastq_cases import EDGE_TEXTS, random_fastq  # noqa: E402
from oracle.oracle import Oracle, RefReadData  # noqa: E402


def golden_texts():
    rng = np.random.default_rng(20261017)
    texts = list(EDGE_TEXTS)
    ends = ("\n", "", "\n@tail", "\n@tail\n", "\n@t\nACGT", "\n\n", "\n", "")
    for i in range(8):
        texts.append(random_fastq(rng, int(rng.integers(1, 400)), int(rng.integers(20, 3000)), crlf=(i % 3 == 2),
                                  end=ends[i]))
    return texts


def main():
    ref = RefReadData.get()
    orc = Oracle.get()
    texts = golden_texts()
    t_cat, t_off, b_cat, b_off, o_cat, o_off = [], [0], [], [0], [], [0]
    for t in texts:
        if len(t) == 0:
            # the reference asserts numReads != 0 (ReadData.cpp:199): no golden, 0 reads by definition
            bases, offsets = np.zeros(0, np.uint8), np.zeros(1, np.uint64)
        else:
            bases, offsets, avg, mx = ref.load_text(t, gzip_flag=False, low_mem=True)
            gb, go, _, _ = ref.load_text(t, gzip_flag=True, low_mem=True)
            assert (gb == bases).all() and (go == offsets).all(), "gzip and plain loads differ"
            lens = np.diff(offsets.astype(np.int64))
            assert avg == int(lens.sum() // lens.size) and mx == int(lens.max())
        t_cat.append(np.frombuffer(t, np.uint8))
        t_off.append(t_off[-1] + len(t))
        b_cat.append(bases)
        b_off.append(b_off[-1] + bases.size)
        o_cat.append(offsets)
        o_off.append(o_off[-1] + offsets.size)
    # the reference's CI file through the reference's loader
    c1 = "/root/reference/util/test_file.fastq.gz"
    bases, offsets, avg, mx = ref.load(c1, True, True)
    hb, ho, _, _ = ref.load(c1, True, False)
    assert (hb == bases).all() and (ho == offsets).all()
    pad = (-bases.size) % 8
    c1_fnv_bases = orc.fnv_u64(np.concatenate([bases, np.zeros(pad, np.uint8)]).view(np.uint64))
    c1_fnv_offsets = orc.fnv_u64(offsets)
    np.savez_compressed(
        os.path.join(HERE, "fastq_golden.npz"),
        texts=np.concatenate(t_cat), text_off=np.asarray(t_off, np.int64),
        bases=np.concatenate(b_cat), bases_off=np.asarray(b_off, np.int64),
        offsets=np.concatenate(o_cat), offsets_off=np.asarray(o_off, np.int64),
        c1=np.asarray([offsets.size - 1, int(offsets[-1]), avg, mx, c1_fnv_bases, c1_fnv_offsets], dtype=np.uint64))
    print(f"{len(texts)} texts, {sum(len(t) for t in texts)} bytes; C1: {offsets.size - 1} reads, {int(offsets[-1])} bases, "
          f"avg {avg}, max {mx}, fnv bases {c1_fnv_bases:016x}, fnv offsets {c1_fnv_offsets:016x}")


if __name__ == "__main__":
    main()

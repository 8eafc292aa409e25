"""The FASTQ-ingest kernels (nanospring_b200/csrc/fastq_kernels.cuh) compiled for the HOST and run in
lock step (tests/cpp/cuda_host_shim.h), compared with the oracle.  A logic check of the device code
for the container without a GPU; the GPU parity proper is tests/test_gpu_fastq.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from fastq_cases import EDGE_TEXTS, expected_packed, random_fastq
from oracle.oracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "libfastq_emul.so")


@pytest.fixture(scope="module")
def emul():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "emul"])
    L = C.CDLL(SO)
    u64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    L.fq_emul_parse.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint, C.c_uint32, u64p, C.c_uint64, u32p, C.c_uint64,
                                u32p, u64p]
    L.fq_emul_unpack.argtypes = [u32p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint]
    return L


def run_emul(L, text, misalign=0, grid=2, slack=64, pack_iters=0):
    """misalign: byte offset of the text inside a 64-byte aligned buffer; slack: readable bytes behind it;
    pack_iters: words per lane and chunk of the pack kernel (0 = the library's default)."""
    raw = np.zeros(len(text) + misalign + slack + 64, dtype=np.uint8)
    base = (-raw.ctypes.data) % 64 + misalign
    buf = raw[base:base + len(text) + slack]
    buf[:len(text)] = np.frombuffer(text, np.uint8)
    buf[len(text):] = 0x0A                     # anything read past the end must not matter
    nl_cap = text.count(b"\n") // 4 + 3
    offsets = np.zeros(nl_cap, dtype=np.uint64)
    words = np.full(len(text) // 16 + 2 + 8, 0, dtype=np.uint32)
    n, nls = C.c_uint32(0), C.c_uint64(0)
    rc = L.fq_emul_parse(buf.ctypes.data, len(text), len(text) + slack, grid, pack_iters,
                         offsets.ctypes.data_as(C.POINTER(C.c_uint64)), offsets.size,
                         words.ctypes.data_as(C.POINTER(C.c_uint32)), words.size - 8, C.byref(n), C.byref(nls))
    assert rc == 0
    assert nls.value == text.count(b"\n")
    off = offsets[:n.value + 1].copy()
    return off, words


def check(L, orc, text, **kw):
    off, words = run_emul(L, text, **kw)
    bases, want_off = orc.fastq_reads(text)
    assert off.size == want_off.size and (off == want_off).all()
    want = expected_packed(bases)
    assert (words[:want.size] == want).all()
    total = int(want_off[-1])
    if total:
        # unpack an unaligned sub-range and the whole stream
        for b0, nb in ((0, total), (min(7, total - 1), max(0, total - min(7, total - 1) - 3))):
            out = np.zeros(nb + 16, dtype=np.uint8)
            L.fq_emul_unpack(words.ctypes.data_as(C.POINTER(C.c_uint32)), b0, nb, out.ctypes.data, 2)
            assert (out[:nb] == orc.store_roundtrip(bases)[b0:b0 + nb]).all()
            assert (out[nb:] == 0).all()


def test_emulated_kernels_edge_texts(emul):
    orc = Oracle.get()
    for t in EDGE_TEXTS:
        for mis in (0, 1, 4, 6):
            check(emul, orc, t, misalign=mis)
        check(emul, orc, t, slack=0)            # no readable byte behind the text


def test_emulated_kernels_random_texts(emul):
    orc = Oracle.get()
    rng = np.random.default_rng(5)
    for i in range(6):
        t = random_fastq(rng, int(rng.integers(1, 40)), int(rng.integers(20, 1500)), crlf=bool(i & 1),
                         end=("\n", "", "\n@tail", "\n@tail\n", "\n@t\nACGT", "\n\n")[i])
        check(emul, orc, t, misalign=int(rng.integers(0, 16)), grid=int(rng.integers(1, 4)))


def test_emulated_kernels_many_short_reads(emul):
    """chunks of the pack kernel that span hundreds of reads, empty reads in a row"""
    orc = Oracle.get()
    rng = np.random.default_rng(9)
    recs = []
    for i in range(700):
        L = int(rng.integers(0, 6)) if i % 5 else 0
        recs.append(b"@\n" + rng.choice(np.frombuffer(b"ACGT", np.uint8), size=L).tobytes() + b"\n+\n" + b"I" * L + b"\n")
    check(emul, orc, b"".join(recs))


def test_emulated_kernels_long_reads_many_steps(emul):
    """reads longer than a pack chunk (16 k bases) and a text of several hundred tiles: every warp
    runs several steps of the 4-tile count loop, the 32-tile newline loop and the unrolled pack loop"""
    orc = Oracle.get()
    rng = np.random.default_rng(21)
    recs = []
    for L in (40000, 3, 17001, 0, 65, 33000, 16, 9000):
        seq = rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=L, p=[.24, .24, .24, .24, .04]).tobytes()
        recs.append(b"@r\n" + seq + b"\n+\n" + b"5" * L + b"\n")
    text = b"".join(recs)
    for mis, grid, iters in ((0, 1, 0), (3, 2, 4), (8, 3, 16), (5, 1, 64)):
        check(emul, orc, text, misalign=mis, grid=grid, pack_iters=iters)

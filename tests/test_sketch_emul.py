"""The sketch kernels (nanospring_b200/csrc/sketch_kernels.cuh: row init, tile map, the filter kernel
with its exact fix-up, the brute-force kernel) compiled for the HOST and run in lock step
(tests/cpp/cuda_host_shim.h), compared with the oracle's sketch matrix.  A logic check of the device
code for the container without a GPU; the GPU parity proper is tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import nanospring_b200 as ns
from fastq_cases import expected_packed
from oracle.oracle import reads_to_buffers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "libsketch_emul.so")
u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)


@pytest.fixture(scope="module")
def emul():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "emul"])
    L = C.CDLL(SO)
    L.sketch_emul_run.argtypes = [u32p, u64p, C.c_uint32, C.c_uint32, C.c_uint32, u64p, C.c_int, C.c_int, C.c_uint32,
                                  C.c_uint, u64p, C.POINTER(C.c_ulonglong)]
    return L


def sketch(L, bases, offsets, k, n, rnd, mode=0, lam=2, tile_words=640, grid=2, deferred=False):
    L.sketch_emul_set_deferred.argtypes = [C.c_int]
    L.sketch_emul_set_deferred(1 if deferred else 0)
    W = np.concatenate([expected_packed(bases), np.zeros(8, np.uint32)])
    N = offsets.size - 1
    sk = np.full((max(N, 1), n), 0x5555555555555555, dtype=np.uint64)
    fix = C.c_ulonglong(0)
    rnd = np.ascontiguousarray(rnd, dtype=np.uint64)
    rc = L.sketch_emul_run(W.ctypes.data_as(u32p), np.ascontiguousarray(offsets).ctypes.data_as(u64p), N, k, n,
                           rnd.ctypes.data_as(u64p), mode, lam, tile_words, grid, sk.ctypes.data_as(u64p), C.byref(fix))
    assert rc == 0
    return sk[:N], fix.value


def read_set(rng, k):
    acgt = np.frombuffer(b"ACGT", np.uint8)
    reads = [b"", b"A" * (k - 2) if k >= 2 else b"", b"C" * (k - 1), rng.choice(acgt, size=k).tobytes(),
             rng.choice(acgt, size=k + 1).tobytes(), b"A" * 700, b"AC" * 400, b"ACGTNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNACGT" * 9,
             rng.integers(0, 256, size=300, dtype=np.uint8).tobytes()]
    reads += [rng.choice(acgt, size=int(l)).tobytes() for l in (15, 16, 17, 31, 33, 63, 64, 65, 1000, 5000, 23000)]
    order = rng.permutation(len(reads))
    return reads_to_buffers([reads[i] for i in order])


@pytest.mark.parametrize("k,n", [(23, 60), (15, 30), (31, 120), (16, 33), (17, 7), (8, 60), (1, 5)])
def test_filter_and_brute_kernels_equal_oracle(emul, orc, k, n):
    rng = np.random.default_rng(k * 1000 + n)
    bases, offsets = read_set(rng, k)
    rnd = ns.rand_from_seed(k + n, n)
    want = orc.sketch_all(bases, offsets, k, n, rnd)
    got, fixups = sketch(emul, bases, offsets, k, n, rnd, mode=0)
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"filter kernel: first difference at (read, hash) {bad[:3].tolist()}"
    # the fix-up with its values in a buffer of their own (nsmh_sketch_build), stored afterwards
    got_d, fixups_d = sketch(emul, bases, offsets, k, n, rnd, mode=0, deferred=True)
    assert (got_d == want).all() and fixups_d == fixups
    if k > 2:
        assert fixups > 0, "the homopolymer / dinucleotide reads need the fix-up pass"
    got, _ = sketch(emul, bases, offsets, k, n, rnd, mode=1)
    assert (got == want).all(), "brute-force kernel"


# (640, 8), (64, 6): nearly every position is a hit, so a step holds more hits than the list (the lane-by-lane
# walk with atomics) and a tile needs several rounds of phase 2
@pytest.mark.parametrize("tile_words,lam", [(64, 2), (128, 0), (4096, 5), (640, 8), (100, 3), (64, 6), (192, 5)])
def test_tile_sizes_and_filter_density(emul, orc, tile_words, lam):
    k, n = 23, 60
    rng = np.random.default_rng(tile_words + lam)
    bases, offsets = read_set(rng, k)
    rnd = ns.rand_from_seed(20261017, n)
    want = orc.sketch_all(bases, offsets, k, n, rnd)
    got, _ = sketch(emul, bases, offsets, k, n, rnd, mode=0, lam=lam, tile_words=tile_words, grid=3)
    assert (got == want).all()


def test_reference_static_kats_through_the_kernels(emul, orc):
    """string2KMers("ACGTTGCAAC", 4) = 45 181 215 94 120 224 130: with rand = 0 the sketch
    is the smallest k-mer; with rand = all-ones it is the complement of the largest."""
    bases, offsets = reads_to_buffers([b"ACGTTGCAAC"])
    for mode in (0, 1):
        got, _ = sketch(emul, bases, offsets, 4, 2, np.array([0, 0xFF], dtype=np.uint64), mode=mode)
        assert got[0, 0] == 45 and got[0, 1] == (224 ^ 0xFF)


@pytest.mark.parametrize("k,n", [(23, 24), (15, 9)])
def test_read_ranges_as_the_pipelined_loaders_sketch_them(emul, orc, k, n):
    """sketch_reads(r0, r1): the kernels see a range of the reads as a read set of its own - the offsets stay
    absolute positions in the one packed stream (so off[0] != 0), the rows are those of the range.  Every
    range, stitched together, must give the sketch matrix of the whole set (nsmh_initialize_* / nsmh_load_sketch_*)."""
    rng = np.random.default_rng(k)
    bases, offsets = read_set(rng, k)
    rnd = ns.rand_from_seed(k * n, n)
    want = orc.sketch_all(bases, offsets, k, n, rnd)
    N = offsets.size - 1
    W = np.concatenate([expected_packed(bases), np.zeros(8, np.uint32)])
    emul.sketch_emul_set_deferred.argtypes = [C.c_int]
    emul.sketch_emul_set_deferred(0)
    rnd_c = np.ascontiguousarray(rnd, dtype=np.uint64)
    cuts = [0, 1, 5, 6, N // 2, N - 1, N]
    got = np.full((N, n), 0x5555555555555555, dtype=np.uint64)
    for r0, r1 in zip(cuts[:-1], cuts[1:]):
        if r1 <= r0:
            continue
        off = np.ascontiguousarray(offsets[r0:r1 + 1])
        sk = np.full((r1 - r0, n), 0x5555555555555555, dtype=np.uint64)
        fix = C.c_ulonglong(0)
        rc = emul.sketch_emul_run(W.ctypes.data_as(u32p), off.ctypes.data_as(u64p), r1 - r0, k, n, rnd_c.ctypes.data_as(u64p),
                                  0, 2, 640, 2, sk.ctypes.data_as(u64p), C.byref(fix))
        assert rc == 0
        got[r0:r1] = sk
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"first difference at (read, hash) {bad[:3].tolist()}"

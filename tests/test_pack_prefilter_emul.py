"""The 2-bit packers (csrc/pack_kernels.cuh) and the candidate pre-filter kernels
(csrc/prefilter_kernels.cuh) compiled for the HOST and run in lock step (tests/cpp/cuda_host_shim.h),
compared with the oracle and with the reference's goldens.  A logic check of the device code for the
container without a GPU; the GPU parity proper is tests/test_gpu_parity.py / test_gpu_prefilter.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN
from fastq_cases import expected_packed
from oracle.oracle import reads_to_buffers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "libpack_prefilter_emul.so")
u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
PAD = 8          # kPackPadWords


@pytest.fixture(scope="module")
def emul():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "emul"])
    L = C.CDLL(SO)
    L.pp_emul_pack_ascii.argtypes = [C.c_void_p, C.c_uint64, u32p, C.c_uint]
    L.pp_emul_pack_rc.argtypes = [u64p, C.c_uint32, u32p, u32p, C.c_uint]
    L.pp_emul_pack_dnabitset.argtypes = [u64p, C.c_uint32, u8p, u64p, u32p, C.c_uint]
    L.pp_emul_read_flags.argtypes = [u64p, C.c_uint32, u32p, u8p, C.c_uint]
    L.pp_emul_csr_drop.argtypes = [u64p, u32p, C.c_uint32, u8p, C.c_uint32, C.c_uint32, u64p, u32p, C.c_uint]
    L.pp_emul_csr_drop.restype = C.c_uint64
    for f in (L.pp_emul_pack_ascii, L.pp_emul_pack_rc, L.pp_emul_pack_dnabitset, L.pp_emul_read_flags):
        f.restype = None
    return L


def pack_ascii(L, bases, misalign=0, grid=2):
    raw = np.zeros(bases.size + 64 + misalign, dtype=np.uint8)
    base = (-raw.ctypes.data) % 16 + misalign
    buf = raw[base:base + bases.size]
    buf[:] = bases
    W = np.zeros((bases.size + 15) // 16 + PAD, dtype=np.uint32)
    L.pp_emul_pack_ascii(buf.ctypes.data, bases.size, W.ctypes.data_as(u32p), grid)
    return W


def read_sets(rng):
    """(bases, offsets) pairs: empty set, empty reads in a row, reads shorter than a word, reads that
    share words, any byte values, one long read."""
    acgt = np.frombuffer(b"ACGT", np.uint8)
    sets = [reads_to_buffers([]), reads_to_buffers([b""]), reads_to_buffers([b"", b"", b"A", b"", b"CG"])]
    sets.append(reads_to_buffers([rng.choice(acgt, size=int(l)).tobytes() for l in rng.integers(0, 40, size=60)]))
    sets.append(reads_to_buffers([rng.integers(0, 256, size=int(l), dtype=np.uint8).tobytes()
                                  for l in (15, 16, 17, 1, 33, 700)]))
    sets.append(reads_to_buffers([rng.choice(acgt, size=9000).tobytes(), b"ACG" * 400, b"AT" * 10, b"G" * 1300,
                                  rng.choice(acgt, size=31).tobytes(), rng.choice(acgt, size=32).tobytes()]))
    return sets


def test_pack_ascii_any_alignment(emul):
    rng = np.random.default_rng(1)
    for bases, _ in read_sets(rng):
        for mis in (0, 1, 8):
            W = pack_ascii(emul, bases, misalign=mis)
            want = expected_packed(bases)
            assert (W[:want.size] == want).all() and (W[want.size:] == 0).all()


def test_reverse_complement_and_dnabitset_layout(emul, orc):
    rng = np.random.default_rng(2)
    for bases, offsets in read_sets(rng):
        n = offsets.size - 1
        W = pack_ascii(emul, bases)
        # reverse complement of every read, on the codes (pack.cu: code ^ 1 of the mirrored position)
        Wrc = np.zeros_like(W)
        emul.pp_emul_pack_rc(offsets.ctypes.data_as(u64p), n, W.ctypes.data_as(u32p), Wrc.ctypes.data_as(u32p), 2)
        stored = orc.store_roundtrip(bases)                       # what getRead hands to the caller
        rc = [orc.reverse_complement(stored[int(offsets[i]):int(offsets[i + 1])].tobytes()) for i in range(n)]
        want = expected_packed(np.frombuffer(b"".join(rc), np.uint8))
        assert (Wrc[:want.size] == want).all()
        # the reference's DnaBitset bytes (4 bases per byte, first base in bits 7..6, per-read alignment)
        lens = np.diff(offsets.astype(np.int64))
        boff = np.zeros(n + 1, dtype=np.uint64)
        boff[1:] = np.cumsum((lens + 3) // 4)
        src = np.zeros(int(boff[-1]) + 16, dtype=np.uint8)
        code = ((bases & 2) | ((bases & 4) >> 2)).astype(np.uint8)
        for i in range(n):
            for j in range(int(lens[i])):
                src[int(boff[i]) + j // 4] |= code[int(offsets[i]) + j] << (6 - 2 * (j % 4))
        Wd = np.zeros_like(W)
        emul.pp_emul_pack_dnabitset(offsets.ctypes.data_as(u64p), n, src.ctypes.data_as(u8p), boff.ctypes.data_as(u64p),
                                    Wd.ctypes.data_as(u32p), 3)
        want = expected_packed(bases)
        assert (Wd[:want.size] == want).all()


def flags_of(L, bases, offsets, grid=2):
    W = pack_ascii(L, bases)
    n = offsets.size - 1
    flags = np.full(max(n, 1), 0xEE, dtype=np.uint8)
    L.pp_emul_read_flags(np.ascontiguousarray(offsets).ctypes.data_as(u64p), n, W.ctypes.data_as(u32p),
                         flags.ctypes.data_as(u8p), grid)
    return flags[:n]


def test_read_flags_equal_oracle_and_reference_golden(emul, orc):
    rng = np.random.default_rng(3)
    for bases, offsets in read_sets(rng):
        assert (flags_of(emul, bases, offsets) == orc.read_flags(bases, offsets)).all()
    pre = np.load(os.path.join(GOLDEN, "prefilter_golden.npz"))
    keep = min(pre["offsets"].size - 1, 400)                      # the emulation is slow: a prefix of the set
    offsets = pre["offsets"][:keep + 1].astype(np.uint64)
    bases = pre["bases"][:int(offsets[-1])]
    got = flags_of(emul, bases, offsets, grid=3)
    assert ((got & 1) == pre["repetitive"][:keep]).all()
    assert (got == orc.read_flags(bases, offsets)).all()


def test_csr_drop_keeps_order_and_foreign_ids(emul):
    rng = np.random.default_rng(4)
    n_flags = 300
    flags = rng.integers(0, 4, size=n_flags).astype(np.uint8)
    rows = [rng.integers(0, n_flags + 50, size=int(l)).astype(np.uint32) for l in (0, 1, 31, 32, 33, 100, 0, 64)]
    off = np.zeros(len(rows) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    ids = np.concatenate(rows)
    for drop in (1, 2, 3):
        new_off = np.zeros(len(rows) + 1, dtype=np.uint64)
        out = np.full(ids.size + 1, 0xDEADBEEF, dtype=np.uint32)
        total = emul.pp_emul_csr_drop(off.ctypes.data_as(u64p), ids.ctypes.data_as(u32p), len(rows), flags.ctypes.data_as(u8p),
                                      n_flags, drop, new_off.ctypes.data_as(u64p), out.ctypes.data_as(u32p), 2)
        want = [r[~((r < n_flags) & ((flags[np.minimum(r, n_flags - 1)] & drop) != 0))] for r in rows]
        assert total == sum(w.size for w in want)
        for q, w in enumerate(want):
            assert (out[int(new_off[q]):int(new_off[q + 1])] == w).all()
        assert out[total] == 0xDEADBEEF

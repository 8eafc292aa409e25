import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def unpack_dnabitset(packed, lengths):
    """DnaBitset bytes (4 bases/byte, first base in bits 7..6, reads byte aligned) -> ASCII
    bases + offsets; inverse table A,T,C,G as in dnaToBits.cpp:82."""
    lengths = np.asarray(lengths, dtype=np.int64)
    nbytes = (lengths + 3) // 4
    codes = np.stack([(packed >> s) & 3 for s in (6, 4, 2, 0)], axis=1).reshape(-1)
    # drop the padding codes at the end of every read
    byte_off = np.concatenate([[0], np.cumsum(nbytes)])
    keep = np.ones(codes.size, dtype=bool)
    pad = nbytes * 4 - lengths
    ends = byte_off[1:] * 4
    for p in (1, 2, 3):
        idx = ends[pad >= p] - p
        keep[idx] = False
    codes = codes[keep]
    bases = np.frombuffer(b"ATCG", dtype=np.uint8)[codes]
    offsets = np.zeros(lengths.size + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lengths)
    return np.ascontiguousarray(bases), offsets


def pack_dnabitset(bases, offsets):
    """ASCII bases + offsets -> DnaBitset bytes + u32 lengths (dnaToBits.cpp:11-36: code (c&2)|((c&4)>>2),
    4 bases per byte, first base in bits 7..6, every read starts on a byte)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    lengths = np.diff(offsets)
    nbytes = (lengths + 3) // 4
    boff = np.concatenate([[0], np.cumsum(nbytes)])
    c = np.asarray(bases[:offsets[-1]], dtype=np.uint8)
    code = ((c & 2) | ((c & 4) >> 2)).astype(np.uint8)
    padded = np.zeros(4 * int(boff[-1]), dtype=np.uint8)
    if code.size:
        pos = np.arange(code.size, dtype=np.int64) + np.repeat(4 * boff[:-1] - offsets[:-1], lengths)
        padded[pos] = code
    q = padded.reshape(-1, 4)
    packed = ((q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]).astype(np.uint8)
    return np.ascontiguousarray(packed), lengths.astype(np.uint32)


@pytest.fixture(scope="session")
def c1_raw():
    z = np.load(os.path.join(GOLDEN, "c1_reads.npz"))
    return z["packed"], z["lengths"]


@pytest.fixture(scope="session")
def c1_reads(c1_raw):
    """The reference's CI input (util/test_file.fastq.gz): (bases u8, offsets u64)."""
    return unpack_dnabitset(*c1_raw)


@pytest.fixture(scope="session")
def c1_golden():
    with open(os.path.join(GOLDEN, "c1_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def edge():
    return np.load(os.path.join(GOLDEN, "edge_golden.npz"))


@pytest.fixture(scope="session")
def orc():
    from oracle.oracle import Oracle
    return Oracle.get()

"""CPU: the restatement of the consensus builder's candidate pre-filters (oracle.read_flags,
Consensus.cpp:405-442 and :213) against the reference's own Consensus translation unit and the
fixture generated from it (tests/golden/make_prefilter_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.oracle import RefConsensus


@pytest.fixture(scope="module")
def pre():
    return np.load(os.path.join(GOLDEN, "prefilter_golden.npz"))


def test_oracle_flags_equal_reference_golden(orc, pre):
    flags = orc.read_flags(pre["bases"], pre["offsets"])
    assert ((flags & 1) == pre["repetitive"]).all()
    lens = np.diff(pre["offsets"].astype(np.int64))
    assert (((flags >> 1) & 1) == (lens < 32)).all()
    assert 100 < int(pre["repetitive"].sum()) < lens.size - 100       # the set brackets the threshold


def test_known_cases(orc):
    def f(s):
        return int(orc.read_flags(np.frombuffer(s, dtype=np.uint8), np.array([0, len(s)], dtype=np.uint64))[0])
    assert f(b"") == 2                      # 0 > 0.7*0 is false (Consensus.cpp:420)
    assert f(b"A") == 3                     # (j+i) % 1 == j: always equal
    assert f(b"ACGTAC") == 3                # shift 6 on length 6 is the identity
    assert f(b"ACG" * 12) == 1              # period 3, length 36
    assert f(b"ACGTTGCAAC" * 5) == 0
    assert f(b"ACGTNNNNACGT") == 2          # N reads back as G (dnaToBits.cpp:81-98)


@pytest.mark.skipif(not RefConsensus.available(), reason="oracle/_ref/libnsref_consensus.so not built")
def test_oracle_equals_reference_code(orc, edge, c1_reads):
    ref = RefConsensus.get()
    for bases, offsets in ((edge["bases"], edge["offsets"]), c1_reads):
        want = ref.is_repetitive(bases, offsets, threads=8)
        got = orc.read_flags(bases, offsets)
        assert ((got & 1) == want).all()
    for s in (b"", b"A", b"AT" * 40, b"ACGTACGTAC", b"GGGGGGG"):
        assert ref.check_repetitive(s) == bool(orc.read_flags(np.frombuffer(s, dtype=np.uint8),
                                                              np.array([0, len(s)], dtype=np.uint64))[0] & 1)

"""GPU: candidate pre-filters on the device (csrc/prefilter.cu) against the oracle, the fixture made
by the reference's own Consensus::initialize, and - when it travelled - the reference library."""
import os

import numpy as np
import pytest

import nanospring_b200 as ns
from conftest import GOLDEN
from nanospring_b200 import ReadData
from oracle.oracle import RefConsensus

pytestmark = pytest.mark.gpu


def flags_of(bases, offsets, k=23, n=8):
    f = ns.MinHashReadFilter(device=0)
    f.k, f.n, f.overlapSketchThreshold, f.randNumbers = k, n, 2, ns.rand_from_seed(3, n)
    f._create()
    f.load(ReadData(bases, offsets))
    flags = f.readFlags()
    again = f.readFlags()                                   # cached
    f.close()
    assert (flags == again).all()
    return flags


def test_flags_equal_reference_golden_and_oracle(orc):
    pre = np.load(os.path.join(GOLDEN, "prefilter_golden.npz"))
    flags = flags_of(pre["bases"], pre["offsets"])
    assert ((flags & 1) == pre["repetitive"]).all()
    assert (flags == orc.read_flags(pre["bases"], pre["offsets"])).all()


def test_flags_on_edge_set_and_ci_file(orc, edge, c1_reads):
    for bases, offsets in ((edge["bases"], edge["offsets"]), c1_reads):
        got = flags_of(bases, offsets)
        assert (got == orc.read_flags(bases, offsets)).all()
        if RefConsensus.available():
            assert ((got & 1) == RefConsensus.get().is_repetitive(bases, offsets, threads=8)).all()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_flags_random_layouts(orc, seed):
    """Many tiny reads per packed word, reads ending at every offset of a word / a 512-base group /
    an 8192-base warp visit, low-complexity and random content."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    reads = []
    for _ in range(3000):
        kind = rng.integers(0, 5)
        L = int(rng.choice([rng.integers(0, 12), rng.integers(0, 80), rng.integers(400, 700), rng.integers(8000, 8500),
                            rng.integers(0, 30000)], p=[0.4, 0.3, 0.15, 0.05, 0.1]))
        if kind == 0:
            a = rng.choice(acgt, size=L)
        else:
            unit = rng.choice(acgt, size=int(rng.integers(1, 8)))
            a = np.resize(unit, L).copy() if L else np.zeros(0, np.uint8)
            hit = rng.random(L) < rng.choice([0.0, 0.15, 0.2, 0.25, 0.4])
            a[hit] = rng.choice(acgt, size=int(hit.sum()))
        reads.append(a.astype(np.uint8).tobytes())
    rd = ReadData.from_reads(reads)
    got = flags_of(rd.bases, rd.offsets)
    want = orc.read_flags(rd.bases, rd.offsets)
    assert (got == want).all(), np.flatnonzero(got != want)[:10]
    assert 0 < int((want & 1).sum()) < len(reads)


def test_drop_flagged_candidates_from_the_bulk_csr(orc, edge):
    k, n, thr = 23, 60, 6
    rnd = ns.rand_from_seed(20261017, n)
    f = ns.MinHashReadFilter(device=0)
    f.k, f.n, f.overlapSketchThreshold, f.randNumbers = k, n, thr, rnd
    f.initialize(ReadData(edge["bases"], edge["offsets"]))
    off, ids = f.queryAll(False)
    flags = f.readFlags()
    assert (flags == orc.read_flags(edge["bases"], edge["offsets"])).all()
    for mask in (0, 1, 2, 3):
        off0, ids0 = f.queryAll(False)
        assert (off0 == off).all() and (ids0 == ids).all()
        doff, dids = f.queryAllDrop(mask)
        keep = (flags[ids] & mask) == 0
        owner = np.repeat(np.arange(off.size - 1), np.diff(off.astype(np.int64)))
        want_counts = np.bincount(owner[keep], minlength=off.size - 1)
        assert (np.diff(doff.astype(np.int64)) == want_counts).all()
        assert (dids == ids[keep]).all()
    assert int(((flags[ids] & 3) != 0).sum()) > 0          # the edge set has short reads that collide
    f.close()

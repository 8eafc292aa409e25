"""One process, several GPUs (include/nsmh.h nsmh_multi_*, csrc/multidev.cu): the reads are split by bases,
every device sketches its shard and builds the tables of ALL reads; sketches, the bulk CSR (forward and
reverse complement) and online queries from many host threads equal the oracle / the single-device filter.
On a box with one GPU the same device is named several times (the logic is the same, the copies are local)."""
import threading

import numpy as np
import pytest

import nanospring_b200 as ns

pytestmark = pytest.mark.gpu


def devices(n):
    import torch
    have = torch.cuda.device_count()
    return [i % have for i in range(n)]


@pytest.mark.parametrize("ndev", [1, 2, 3, 8])
def test_multi_device_filter_equals_oracle(orc, ndev):
    k, n, thr = 23, 60, 4
    rnd = ns.rand_from_seed(77, n)
    lengths = ns.synth_lengths(3000, 2500, seed=5)
    lengths[:6] = [0, 1, k - 1, k, 60000, 33]
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=250_000, genome_seed=3, read_seed=4,
                                                      p_ins=0.01, p_del=0.01, p_sub=0.01))
    f = ns.MultiGpuMinHashReadFilter(devices(ndev))
    f.k, f.n, f.overlapSketchThreshold, f.randNumbers = k, n, thr, rnd
    f.initialize(rd)
    sh = f.shards()
    assert sh[0] == 0 and sh[-1] == rd.numReads and (np.diff(sh.astype(np.int64)) >= 0).all()
    if ndev > 1:
        b = np.diff(rd.offsets[sh].astype(np.int64))
        assert b.max() - b.min() <= int(lengths.max()) + 1, "shards hold equal numbers of bases, up to one read"
    want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
    assert (f.sketches() == want).all()
    T = orc.build_tables(want)
    for rc in (False, True):
        off, ids = f.queryAll(rc)
        woff, wids = T.query_all(rd.bases, rd.offsets, want, k, rnd, thr, int(rc))
        assert (off == woff).all() and (ids == wids).all(), f"bulk query, rc {rc}, {ndev} devices"
    # online queries from many threads: every device must give the same answer
    errors = []

    def worker(t):
        try:
            for i in range(t, rd.numReads, 97):
                q = rd.getRead(i)[:3000]
                got = f.getFilteredReads(q)
                if not (got == T.query_string(q, k, rnd, thr)).all():
                    errors.append(i)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(2 * ndev + 1)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors[:5]
    f.close()


def test_multi_device_dnabitset_input(orc):
    """The reference's 2-bit store as input (nsmh_multi_load_reads_dnabitset)."""
    from conftest import pack_dnabitset
    k, n, thr = 15, 30, 3
    rnd = ns.rand_from_seed(3, n)
    rd = ns.synth_reads_host(ns.synth_lengths(800, 900, seed=9), ns.synth_params(genome_len=60_000))
    packed, lens = pack_dnabitset(rd.bases, rd.offsets)
    f = ns.MultiGpuMinHashReadFilter(devices(2))
    f.k, f.n, f.overlapSketchThreshold, f.randNumbers = k, n, thr, rnd
    f.initialize(None, packed=(packed, lens))
    want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
    assert (f.sketches() == want).all()
    off, ids = f.queryAll(False)
    woff, wids = orc.build_tables(want).query_all(rd.bases, rd.offsets, want, k, rnd, thr, 0)
    assert (off == woff).all() and (ids == wids).all()
    f.close()

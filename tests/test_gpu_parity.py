"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI
(include/nsmh.h via nanospring_b200.filter), against
  * the committed outputs of the UNMODIFIED reference (tests/golden/*), and
  * the CPU oracle (oracle/minhash_oracle.c) on the same seeded inputs.
Bit-exact everywhere: sketches (u64) and candidate sets (ascending u32 ids per read).
"""
import threading

import numpy as np
import pytest

import nanospring_b200 as ns
from nanospring_b200.filter import MinHashReadFilter, ReadData

pytestmark = pytest.mark.gpu


def make_filter(k, n, thr, rnd, mode=0):
    f = MinHashReadFilter()
    f.k, f.n, f.overlapSketchThreshold = k, n, thr
    f.randNumbers = np.asarray(rnd, dtype=np.uint64)
    f.sketchMode = mode
    return f


def assert_csr_equal(got, want, what):
    off_g, ids_g = got
    off_w, ids_w = want
    assert off_g.shape == off_w.shape, what
    bad = np.nonzero(off_g != off_w)[0]
    assert bad.size == 0, f"{what}: offsets first differ at read {bad[:1]}"
    assert ids_g.size == ids_w.size and (ids_g == ids_w).all(), f"{what}: ids differ"


@pytest.mark.parametrize("mode", [0, 1])
def test_edge_set_all_configs(edge, mode):
    """Empty reads, len k-2..k+1, non-ACGT bytes, duplicates, homopolymers, one long read:
    the reference's full outputs for seven (k, n, thr) settings."""
    rd = ReadData(edge["bases"], edge["offsets"])
    for ci, (seed, k, n, thr) in enumerate(edge["cfgs"]):
        k, n, thr = int(k), int(n), int(thr)
        f = make_filter(k, n, thr, edge[f"rand_{ci}"], mode)
        f.initialize(rd)
        sk = f.sketches()
        want = edge[f"sketches_{ci}"]
        bad = np.argwhere(sk != want)
        assert bad.size == 0, f"cfg {ci} mode {mode}: sketch differs at (read,hash) {bad[:3].tolist()}"
        assert_csr_equal(f.queryAll(False), (edge[f"fwd_off_{ci}"], edge[f"fwd_ids_{ci}"]), f"cfg {ci} fwd")
        assert_csr_equal(f.queryAll(True), (edge[f"rc_off_{ci}"], edge[f"rc_ids_{ci}"]), f"cfg {ci} rc")
        f.close()


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_c1_reference_ci_file(orc, c1_reads, c1_golden, idx):
    """BASELINE config #1: util/test_file.fastq.gz, checked against the golden checksums of
    the reference AND element-wise against the oracle."""
    bases, offsets = c1_reads
    g = c1_golden["settings"][idx]
    rnd = ns.rand_from_seed(g["seed"], g["n"])
    assert "%016x" % rnd[0] == g["rand_first"] and "%016x" % rnd[-1] == g["rand_last"]
    f = make_filter(g["k"], g["n"], g["thr"], rnd)
    f.initialize(ReadData(bases, offsets))
    sk = f.sketches()
    assert "%016x" % orc.fnv_u64(sk.ravel()) == g["fnv_sketches"]
    want = orc.sketch_all(bases, offsets, g["k"], g["n"], rnd)
    assert (sk == want).all()
    off, ids = f.queryAll(False)
    assert int(off[-1]) == g["fwd_total"]
    assert "%016x" % orc.fnv_csr(off, ids) == g["fwd_fnv"]
    assert [int(x) for x in ids[int(off[4]):int(off[5])]] == g["fwd_cands_read4"]
    offr, idsr = f.queryAll(True)
    assert int(offr[-1]) == g["rc_total"]
    assert "%016x" % orc.fnv_csr(offr, idsr) == g["rc_fnv"]
    T = orc.build_tables(want)
    assert_csr_equal((off, ids), T.query_all(bases, offsets, want, g["k"], rnd, g["thr"], 0), "fwd")
    assert_csr_equal((offr, idsr), T.query_all(bases, offsets, want, g["k"], rnd, g["thr"], 1), "rc")
    # distinct keys per table == BBHashMap::numKeys
    for j in (0, g["n"] // 2, g["n"] - 1):
        assert f.tableNumKeys(j) == T.num_keys(j)
    f.close()


def test_c1_brute_force_kernel_equals_filter_kernel(c1_reads, c1_golden, orc):
    bases, offsets = c1_reads
    g = c1_golden["settings"][0]
    rnd = ns.rand_from_seed(g["seed"], g["n"])
    f = make_filter(g["k"], g["n"], g["thr"], rnd, mode=1)
    f.load(ReadData(bases, offsets))
    f.sketch()
    assert "%016x" % orc.fnv_u64(f.sketches().ravel()) == g["fnv_sketches"]
    f.close()


def test_c1_dnabitset_loader(c1_raw, c1_golden, orc):
    """Reads handed over in the reference's own 2-bit layout (DnaBitset)."""
    packed, lengths = c1_raw
    g = c1_golden["settings"][0]
    f = make_filter(g["k"], g["n"], g["thr"], ns.rand_from_seed(g["seed"], g["n"]))
    f.load_dnabitset(packed, lengths)
    f.sketch()
    assert "%016x" % orc.fnv_u64(f.sketches().ravel()) == g["fnv_sketches"]
    f.close()


@pytest.mark.parametrize("k,n,thr", [(23, 60, 6), (15, 30, 3), (31, 120, 12), (9, 33, 2)])
def test_sketch_build_overlap_equals_separate_calls(orc, k, n, thr):
    """nsmh_sketch_build: the sketch's fix-up pass runs on the second stream beside the table insert, which leaves
    the all-ones entries to table_insert_list_kernel.  Sketches, distinct keys per table and both bulk lookups must
    be those of nsmh_sketch + nsmh_build; reads of k-1 bases (legitimate all-ones rows), empty reads, reads shorter
    than the filter prefix and low-complexity reads (many fix-ups) are in the set."""
    lengths = ns.synth_lengths(900, 2500, seed=k)
    lengths[:8] = [0, 1, k - 2 if k > 2 else 0, k - 1, k - 1, k, k + 3, 70000]
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=150_000, genome_seed=k, read_seed=n))
    b = rd.bases.copy()
    o = rd.offsets
    b[int(o[20]):int(o[21])] = ord("A")                      # homopolymer / dinucleotide reads: hardly any distinct k-mer
    b[int(o[21]):int(o[22]):2] = ord("C")
    rd = ReadData(b, o)
    rnd = ns.rand_from_seed(5 * k + n, n)
    want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
    ref = make_filter(k, n, thr, rnd)
    ref.load(rd)
    ref.sketch()
    ref.build()
    keys = [ref.tableNumKeys(j) for j in range(0, n, 7)]
    fwd, rc = ref.queryAll(False), ref.queryAll(True)
    assert (ref.sketches() == want).all()
    ref.close()
    f = make_filter(k, n, thr, rnd)
    for rep in range(2):                                        # twice on one handle: buffers and events are reused
        f.load(rd)
        f.sketch_build()
        assert (f.sketches() == want).all(), "sketches"
        assert [f.tableNumKeys(j) for j in range(0, n, 7)] == keys
        assert_csr_equal(f.queryAll(False), fwd, "fwd")
        assert_csr_equal(f.queryAll(True), rc, "rc")
    assert f.stats()["sketch_fixups"] > 0
    f.close()


@pytest.mark.parametrize("chunk", [64, 4000, 70001, 1 << 30])
def test_pipelined_initialize_equals_separate_calls(orc, monkeypatch, chunk):
    """nsmh_initialize_ascii / _dnabitset (load + sketch + build pipelined: a chunk's reads are sketched while
    the next chunk is copied) against the three separate calls and the oracle, with the chunk size forced
    down so that reads straddle chunks, chunks hold no whole read, and empty reads sit on chunk borders."""
    from conftest import pack_dnabitset
    k, n, thr = 21, 24, 3
    lengths = ns.synth_lengths(400, 1500, seed=8)
    lengths[:10] = [0, 0, 5, k - 1, k, 16, 17, 30000, 0, 64]
    lengths[-3:] = [0, 7, 0]
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=60_000, genome_seed=11, read_seed=12))
    rnd = ns.rand_from_seed(99, n)
    want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
    ref = make_filter(k, n, thr, rnd)
    ref.load(rd)
    ref.sketch()
    ref.build()
    assert (ref.sketches() == want).all()
    want_csr = ref.queryAll(False)
    ref.close()
    monkeypatch.setenv("NSMH_LOAD_CHUNK_BYTES", str(chunk))
    packed, len32 = pack_dnabitset(rd.bases, rd.offsets)
    for how in ("ascii", "dnabitset", "dnabitset_load_only", "ascii_load_sketch", "dnabitset_load_sketch"):
        f = make_filter(k, n, thr, rnd)
        if how == "ascii":
            f.initialize(rd)
        elif how == "dnabitset":
            f.initialize_dnabitset(packed, len32)
        elif how.endswith("load_sketch"):
            f.load_sketch(rd) if how.startswith("ascii") else f.load_sketch(packed=(packed, len32))
            f.build()
        else:
            f.load_dnabitset(packed, len32)
            f.sketch()
            f.build()
        assert (f.sketches() == want).all(), how
        assert_csr_equal(f.queryAll(False), want_csr, how)
        # a second initialize on the same handle (the wrapper keeps it) starts from a clean state
        if how == "ascii":
            f.initialize(rd)
            assert_csr_equal(f.queryAll(False), want_csr, how + " again")
        f.close()


@pytest.mark.parametrize("k,n,thr", [(23, 60, 6), (15, 30, 3), (31, 120, 12), (16, 33, 4), (17, 7, 1)])
def test_synthetic_reads_vs_oracle(orc, k, n, thr):
    """Nanopore-like synthetic reads (createData.py recipe, 10% error, 20x coverage of a small
    genome so that real overlaps exist), ragged lengths incl. tile-boundary cases."""
    lengths = ns.synth_lengths(1500, 3000, seed=3)
    lengths[:12] = [0, 1, k - 2, k - 1, k, k + 1, 2047 + k, 2048 + k, 2049 + k, 4096 + k - 1, 150000, 33]
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=200_000, genome_seed=5, read_seed=6))
    rnd = ns.rand_from_seed(20261017, n)
    f = make_filter(k, n, thr, rnd)
    f.initialize(rd)
    want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
    sk = f.sketches()
    bad = np.argwhere(sk != want)
    assert bad.size == 0, f"sketch differs at {bad[:3].tolist()}"
    T = orc.build_tables(want)
    fwd = f.queryAll(False)
    assert_csr_equal(fwd, T.query_all(rd.bases, rd.offsets, want, k, rnd, thr, 0), "fwd")
    assert int(fwd[0][-1]) > rd.numReads          # real overlaps were found, not only self hits
    assert_csr_equal(f.queryAll(True), T.query_all(rd.bases, rd.offsets, want, k, rnd, thr, 1), "rc")
    f.close()


def test_device_synth_equals_host_synth():
    import torch
    lengths = ns.synth_lengths(400, 2000, seed=9)
    p = ns.synth_params(genome_len=100_000, genome_seed=3, read_seed=4)
    rd = ns.synth_reads_host(lengths, p, first_read=17)
    d_off = torch.from_numpy(rd.offsets.astype(np.int64)).cuda()
    d_bases = torch.zeros(int(rd.offsets[-1]) + 16, dtype=torch.uint8, device="cuda")
    import ctypes as C
    ns._lib.check(ns.lib().nsmh_synth_reads_device(0, C.byref(p), 17, lengths.size, d_off.data_ptr(),
                                                   d_bases.data_ptr()))
    got = d_bases[:int(rd.offsets[-1])].cpu().numpy()
    assert (got == rd.bases).all()


def test_online_query_matches_reference_semantics(orc, edge):
    """getFilteredReads(string): forward windows, reverse complements, foreign strings, short
    strings; also from many host threads at once (Consensus.cpp:29,189)."""
    bases, offsets = edge["bases"], edge["offsets"]
    seed, k, n, thr = (int(v) for v in edge["cfgs"][0])
    rnd = edge["rand_0"]
    f = make_filter(k, n, thr, rnd)
    f.initialize(ReadData(bases, offsets))
    T = orc.build_tables(edge["sketches_0"])
    rd = ReadData(bases, offsets)
    queries = [rd.getRead(i) for i in (0, 1, 3, 20, 27, 30, 60, 100, 164, 165, 166, 167, 168)]
    queries += [rd.getRead(164)[500:4500], ns.reverse_complement(rd.getRead(164)[1000:3000]),
                b"ACGT" * 100, b"", b"A" * 22, b"A" * 21, b"N" * 50]
    want = [T.query_string(q, k, rnd, thr) for q in queries]
    for q, w in zip(queries, want):
        got = f.getFilteredReads(q)
        assert got.dtype == np.uint32 and (got == w).all()
    res = []
    f.getFilteredReads(queries[0], res)
    assert res == [int(x) for x in want[0]]
    for got, w in zip(f.getFilteredReadsBatch(queries), want):
        assert (got == w).all()
    # concurrent callers
    errors = []

    def worker(tid):
        try:
            for rep in range(5):
                for q, w in list(zip(queries, want))[tid % 3::3]:
                    got = f.getFilteredReads(q)
                    if got.size != w.size or (got != w).any():
                        errors.append((tid, len(q)))
        except Exception as e:  # noqa: BLE001
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    # raw sketches (private overload, ReadFilter.cpp:65-83)
    off, ids = f.querySketches(edge["sketches_0"][:50])
    assert_csr_equal((off, ids), (edge["fwd_off_0"][:51], edge["fwd_ids_0"][:int(edge["fwd_off_0"][50])]), "sk")
    f.close()


def test_error_behaviour():
    f = MinHashReadFilter()
    f.k = 32                                   # UB in the reference (1ull << 64); rejected here
    with pytest.raises(ns.NsmhError):
        f.initialize([b"ACGT"])
    f = MinHashReadFilter()
    f.k, f.n = 5, 4
    f.randNumbers = ns.rand_from_seed(1, 4)
    f._create()
    with pytest.raises(ns.NsmhError):
        f.sketch()                             # no reads loaded
    f.load([b"ACGTACGT", b"ACGTAC"])
    with pytest.raises(ns.NsmhError):
        f.queryAll()                           # not sketched / built
    f.sketch()
    with pytest.raises(ns.NsmhError):
        f.getFilteredReads(b"ACGTACG")         # tables not built
    f.build()
    assert f.getFilteredReads(b"ACGTACGT").size == 0     # thr 6 > n 4: nothing can qualify
    f.close()


def test_threshold_extremes(orc):
    """thr = 0 / 1 (every id that shares a slot), thr = n (identical sketches only), thr > n."""
    lengths = ns.synth_lengths(300, 800, seed=5)
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=20_000, genome_seed=8, read_seed=9,
                                                      p_ins=0.0, p_del=0.0, p_sub=0.01))
    k, n = 15, 16
    rnd = ns.rand_from_seed(77, n)
    want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
    T = orc.build_tables(want)
    for thr in (0, 1, 2, n, n + 1):
        f = make_filter(k, n, thr, rnd)
        f.initialize(rd)
        assert_csr_equal(f.queryAll(False), T.query_all(rd.bases, rd.offsets, want, k, rnd, thr, 0), f"thr {thr}")
        f.close()


def test_large_roundtrip_properties():
    """BASELINE-size-independent properties on a larger synthetic set (no oracle needed):
    brute-force and filter kernels agree bit for bit; every read finds itself; candidate lists
    ascend; the relation is symmetric for the forward self-query."""
    lengths = ns.synth_lengths(20000, 8000, seed=12)
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=5_000_000, genome_seed=21, read_seed=22))
    rnd = ns.rand_from_seed(20261017, 60)
    f = make_filter(23, 60, 6, rnd, mode=0)
    f.initialize(rd)
    sk0 = f.sketches()
    off, ids = f.queryAll(False)
    b = make_filter(23, 60, 6, rnd, mode=1)
    b.load(rd)
    b.sketch()
    assert (b.sketches() == sk0).all()
    b.close()
    N = rd.numReads
    owner = np.repeat(np.arange(N, dtype=np.int64), np.diff(off.astype(np.int64)))
    assert (np.bincount(owner[owner == ids], minlength=N) == 1).all()  # self is a candidate
    d = np.diff(ids.astype(np.int64))
    boundaries = off[1:-1].astype(np.int64) - 1
    inner = np.ones(d.size, dtype=bool)
    inner[boundaries[(boundaries >= 0) & (boundaries < d.size)]] = False
    assert (d[inner] > 0).all()                                        # ascending, no duplicates
    fwd = set(zip(owner.tolist(), ids.tolist()))
    assert all((j, i) in fwd for (i, j) in fwd)                        # symmetric
    f.close()


# ---------------------------------------------------------------------------------------------
# Tests aimed at the internals of the sm_100a kernels (tile staging, filter levels, fix-up,
# bucketed tables, the three lookup paths).  All of them compare with the oracle bit for bit.
def _mk_reads(seqs):
    return ReadData.from_reads(seqs)


def _random_seq(rng, n):
    return rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n).tobytes()


@pytest.mark.parametrize("tile_words", [64, 128, 640, 4096])
def test_tile_sizes_and_alignment(orc, monkeypatch, tile_words):
    """Tiles start on 16-byte boundaries of the packed stream and are staged with one bulk copy:
    reads that start at every word offset, lengths around the tile size, many tiles per read."""
    monkeypatch.setenv("NSMH_TILE_WORDS", str(tile_words))
    k, n, thr = 23, 60, 6
    rng = np.random.default_rng(11)
    tile_bases = tile_words * 16
    lens = [1, 5, 17, 40, 63, 64, 65, 100]                     # shift the following reads through all alignments
    for d in (-65, -17, -1, 0, 1, 15, 16, 17, 63, 64):
        lens += [tile_bases + k - 1 + d, 3, 2 * tile_bases + d, 29]
    lens += [5 * tile_bases + 7, 12345]
    seqs = [_random_seq(rng, L) for L in lens]
    rd = _mk_reads(seqs)
    rnd = ns.rand_from_seed(5, n)
    f = make_filter(k, n, thr, rnd)
    f.initialize(rd)
    want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
    bad = np.argwhere(f.sketches() != want)
    assert bad.size == 0, f"tile_words {tile_words}: sketch differs at {bad[:3].tolist()}"
    f.close()


@pytest.mark.parametrize("lam", [0, 1, 2, 3, 5, 8])
def test_filter_level_does_not_change_results(orc, monkeypatch, lam):
    """The prefix filter is an exact shortcut at every level: few expected k-mers per bucket
    (lam 0: 1..2, a third of the (read, hash) pairs go through the fix-up kernels) or many."""
    monkeypatch.setenv("NSMH_LAMBDA_LOG2", str(lam))
    k, n = 21, 64
    lengths = ns.synth_lengths(400, 4000, seed=8)
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=300_000, genome_seed=2, read_seed=3))
    rnd = ns.rand_from_seed(99, n)
    f = make_filter(k, n, 6, rnd)
    f.load(rd)
    f.sketch()
    want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
    assert (f.sketches() == want).all()
    if lam == 0:
        assert f.stats()["sketch_fixups"] > 1000          # the fix-up path really ran
    f.close()


@pytest.mark.parametrize("k", [9, 16, 17, 31])
def test_fixup_with_repeats_and_ties(orc, monkeypatch, k):
    """Fix-up scan on 32-bit prefixes: reads made of one unit repeated at a period of 32 packed
    words put equal k-mers into different words of the SAME lane (the tie fallback), plus
    homopolymers / dinucleotide repeats where no bucket of the filter is hit."""
    monkeypatch.setenv("NSMH_LAMBDA_LOG2", "0")
    rng = np.random.default_rng(k)
    unit = _random_seq(rng, 512)
    seqs = [unit * 4, unit * 7 + b"ACGT", b"G" + unit * 3, b"A" * 3000, b"AC" * 2000, b"ACG" * 1500,
            _random_seq(rng, 5000), unit[:300] * 9]
    rd = _mk_reads(seqs)
    n = 40
    rnd = ns.rand_from_seed(1234, n)
    f = make_filter(k, n, 3, rnd)
    f.initialize(rd)
    want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
    bad = np.argwhere(f.sketches() != want)
    assert bad.size == 0, f"k {k}: sketch differs at {bad[:3].tolist()}"
    T = orc.build_tables(want)
    assert_csr_equal(f.queryAll(False), T.query_all(rd.bases, rd.offsets, want, k, rnd, 3, 0), "fwd")
    f.close()


@pytest.mark.parametrize("thr", [1, 6, 60])
def test_lookup_paths_small_sort_heavy(orc, thr):
    """Families of identical reads of size 1, 3 (counting table in shared memory), 10 (bitonic
    sort path: 600 gathered ids) and 30 (global path: 1800 ids), plus near-duplicates that share
    only some slots; every table group of two or more members goes through the member lists."""
    k, n = 23, 60
    rng = np.random.default_rng(3)
    seqs = []
    for copies in (1, 3, 10, 30, 1, 3):
        s = _random_seq(rng, 3000)
        seqs += [s] * copies
        m = bytearray(s)
        for p in range(0, 3000, 150):                      # a mutated relative: shares some minima only
            m[p] = ord("A") if m[p] != ord("A") else ord("C")
        seqs.append(bytes(m))
    order = rng.permutation(len(seqs))
    seqs = [seqs[i] for i in order]
    rd = _mk_reads(seqs)
    rnd = ns.rand_from_seed(42, n)
    f = make_filter(k, n, thr, rnd)
    f.initialize(rd)
    want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
    assert (f.sketches() == want).all()
    T = orc.build_tables(want)
    for j in (0, n - 1):
        assert f.tableNumKeys(j) == T.num_keys(j)
    got = f.queryAll(False)
    assert_csr_equal(got, T.query_all(rd.bases, rd.offsets, want, k, rnd, thr, 0), f"thr {thr}")
    assert int(np.diff(got[0].astype(np.int64)).max()) >= 30
    f.close()


def test_many_results_per_query(orc):
    """thr = 1 at high coverage: dozens of result ids per query (shuffle sort of <= 32 results,
    shared-memory sort beyond), gathered-id counts on both sides of the counting-table limit."""
    lengths = ns.synth_lengths(1200, 900, seed=15)
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=12_000, genome_seed=4, read_seed=5,
                                                      p_ins=0.0, p_del=0.0, p_sub=0.005))
    k, n = 15, 24
    rnd = ns.rand_from_seed(7, n)
    want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
    T = orc.build_tables(want)
    for thr in (1, 2, 5):
        f = make_filter(k, n, thr, rnd)
        f.initialize(rd)
        got = f.queryAll(False)
        assert_csr_equal(got, T.query_all(rd.bases, rd.offsets, want, k, rnd, thr, 0), f"thr {thr}")
        if thr == 1:
            assert int(np.diff(got[0].astype(np.int64)).max()) > 64
        f.close()


def test_rebuild_and_requery_same_handle(orc):
    """A handle is reused for a second, different read set (tables are pre-cleared on the copy
    stream while the new reads are sketched): no state of the first set may leak."""
    k, n, thr = 23, 60, 6
    rnd = ns.rand_from_seed(20261017, n)
    f = make_filter(k, n, thr, rnd)
    for seed, count in ((1, 900), (2, 300), (3, 1500)):
        lengths = ns.synth_lengths(count, 2500, seed=seed)
        rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=150_000, genome_seed=seed, read_seed=seed + 10))
        f.initialize(rd)
        want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
        assert (f.sketches() == want).all()
        T = orc.build_tables(want)
        assert_csr_equal(f.queryAll(False), T.query_all(rd.bases, rd.offsets, want, k, rnd, thr, 0), f"set {seed}")
    f.close()


@pytest.mark.parametrize("thr", [1, 6, 20])
def test_counting_filter_tier_small_k(orc, monkeypatch, thr):
    """k = 8: every table group holds tens of reads that share a sketch value by chance, so most
    queries gather 1000..3000 ids (more than the warp's sort buffer) of which a handful reach the
    threshold.  The counting-filter tier (csrc/query_kernels.cuh) resolves them without the global sort;
    with thr = 1 every id survives the filter and the queries are handed on.  Same CSR as the oracle
    and as the run with the tier switched off."""
    k, n = 8, 60
    rnd = ns.rand_from_seed(11, n)
    lengths = ns.synth_lengths(3000, 1500, seed=21)
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=300_000, genome_seed=5, read_seed=6))
    want_sk = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
    T = orc.build_tables(want_sk)
    results = {}
    for tier in ("1", "0"):
        monkeypatch.setenv("NSMH_MID_TIER", tier)
        f = make_filter(k, n, thr, rnd)
        f.initialize(rd)
        assert (f.sketches() == want_sk).all()
        for rc in (False, True):
            got = f.queryAll(rc)
            st = f.stats()
            assert_csr_equal(got, T.query_all(rd.bases, rd.offsets, want_sk, k, rnd, thr, int(rc)), f"tier {tier} rc {rc}")
            results[(tier, rc)] = (st["query_heavy"], st["query_sorted"])
        f.close()
    for rc in (False, True):
        heavy_on, sorted_on = results[("1", rc)]
        heavy_off, sorted_off = results[("0", rc)]
        assert heavy_on == heavy_off and heavy_on > 1000, "most queries overflow the warp buffer on this input"
        assert sorted_off == heavy_off, "tier off: every heavy query goes through the global sort"
        if thr == 1:
            assert sorted_on == heavy_on, "thr 1: nothing can be filtered out"
        else:
            assert sorted_on < heavy_on // 10, "the tier resolves nearly all heavy queries"


@pytest.mark.parametrize("tile_words,lam", [("64", "6"), ("640", "8"), ("192", "5"), ("640", "2")])
@pytest.mark.parametrize("k,n", [(23, 60), (15, 30), (31, 120), (9, 33)])
def test_filter_kernel_list_rounds_and_overflow(orc, edge, monkeypatch, k, n, tile_words, lam):
    """sketch_filter_kernel's hit list: dense filters (NSMH_LAMBDA_LOG2 6 / 8: nearly every position is a hit)
    and small tiles (NSMH_TILE_WORDS) force several rounds of phase 2 per tile and the lane-by-lane walk of a
    step that holds more hits than the list; same sketch matrix as the oracle in every setting."""
    monkeypatch.setenv("NSMH_TILE_WORDS", tile_words)
    monkeypatch.setenv("NSMH_LAMBDA_LOG2", lam)
    rnd = ns.rand_from_seed(k * n, n)
    lengths = ns.synth_lengths(1500, 3000, seed=k)
    lengths[:8] = [0, 1, k - 1, k, k + 1, 40, 70000, 33]
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=200_000, genome_seed=2, read_seed=3))
    for s in (rd, ReadData(edge["bases"], edge["offsets"])):
        want = orc.sketch_all(s.bases, s.offsets, k, n, rnd)
        f = make_filter(k, n, 3, rnd)
        f.initialize(s)
        assert (f.sketches() == want).all()
        f.close()


@pytest.mark.skipif(not __import__("os").environ.get("NSMH_TEST_EXPERIMENTS"),
                    reason="experimental host flow: set NSMH_TEST_EXPERIMENTS=1 (tools/gpu_session.sh does)")
def test_experiment_speculative_placement_of_the_lookup(orc, edge, monkeypatch):
    """NSMH_LOOKUP_SPECULATE=1 (query.cu: prefix sum + guarded placement queued behind the counting kernel,
    one host round trip when no query overflows): same CSR as the oracle on inputs that take the fast
    exit, the result-buffer overflow (dozens of results per query), the counting-filter tier and the
    global sort; bulk and online queries."""
    monkeypatch.setenv("NSMH_LOOKUP_SPECULATE", "1")
    cases = []
    lengths = ns.synth_lengths(1200, 900, seed=15)
    cases.append((ns.synth_reads_host(lengths, ns.synth_params(genome_len=12_000, genome_seed=4, read_seed=5,
                                                               p_ins=0.0, p_del=0.0, p_sub=0.005)), 15, 24, (1, 2, 5)))
    lengths = ns.synth_lengths(3000, 1500, seed=21)
    cases.append((ns.synth_reads_host(lengths, ns.synth_params(genome_len=300_000, genome_seed=5, read_seed=6)), 8, 60, (1, 6)))
    cases.append((ReadData(edge["bases"], edge["offsets"]), 23, 60, (6,)))
    lengths = ns.synth_lengths(2000, 2500, seed=1)
    cases.append((ns.synth_reads_host(lengths, ns.synth_params(genome_len=150_000)), 23, 60, (6,)))     # fast exit
    for rd, k, n, thrs in cases:
        rnd = ns.rand_from_seed(k + n, n)
        want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
        T = orc.build_tables(want)
        for thr in thrs:
            f = make_filter(k, n, thr, rnd)
            f.initialize(rd)
            for rc in (False, True):
                assert_csr_equal(f.queryAll(rc), T.query_all(rd.bases, rd.offsets, want, k, rnd, thr, int(rc)),
                                 f"k {k} n {n} thr {thr} rc {rc}")
            for i in (0, 7, rd.numReads - 1):
                q = rd.getRead(i)
                assert (f.getFilteredReads(q) == T.query_string(q, k, rnd, thr)).all()
            f.close()

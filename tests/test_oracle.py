"""CPU tests: pin the oracle (oracle/minhash_oracle.c) against
  * the reference's static known answers,
  * the golden checksums produced by the UNMODIFIED reference code on its own CI file
    (tests/golden/c1_golden.json, made by tests/golden/make_golden.py),
  * the reference's full outputs on the hand-made edge set (tests/golden/edge_golden.npz),
  * the reference build itself (oracle/_ref/libnsref.so) wherever it is present.
"""
import numpy as np
import pytest

from oracle.oracle import Oracle, RefLib, reads_to_buffers


def test_known_answers(orc):
    # MinHashReadFilter::kMerToInt / string2KMers (ReadFilter.cpp:101-109, 138-152)
    assert orc.kmer_to_int("ATCG") == 27
    assert orc.kmer_to_int("GGGG") == 255
    assert list(orc.string2kmers("ACGTTGCAAC", 4)) == [45, 181, 215, 94, 120, 224, 130]
    assert orc.string2kmers("ACG", 4).size == 0
    # base code for arbitrary bytes: N -> 3, lower case like upper case
    assert [orc.kmer_to_int(c) for c in "ATCGNatcgn"] == [0, 1, 2, 3, 3, 0, 1, 2, 3, 3]


def test_rand_stream_matches_mt19937_64(orc, c1_golden):
    for s in c1_golden["settings"]:
        r = orc.rand_from_seed(s["seed"], s["n"])
        assert "%016x" % r[0] == s["rand_first"]
        assert "%016x" % r[-1] == s["rand_last"]
    # first outputs of std::mt19937_64 with the default seed 5489 (C++11 26.5.5: 10000th is
    # 9981545732273789042)
    r = orc.rand_from_seed(5489, 10000)
    assert int(r[9999]) == 9981545732273789042


def test_short_read_semantics(orc):
    rnd = orc.rand_from_seed(3, 8)
    k = 5
    assert (orc.string2sketch("ACG", k, 8, rnd) == 0).all()            # len < k-1: untouched
    assert (orc.string2sketch("ACGT", k, 8, rnd) == 2**64 - 1).all()   # len == k-1: all ones
    sk = orc.string2sketch("ACGTA", k, 8, rnd)                         # exactly one k-mer
    x = orc.kmer_to_int("ACGTA")
    assert (sk == (np.uint64(x) ^ rnd)).all()


def test_reverse_complement(orc):
    assert orc.reverse_complement(b"AACGTN") == b"NACGTT"
    assert orc.reverse_complement(b"acgt") == b"tgca"     # lower case is NOT complemented
    assert orc.reverse_complement(b"") == b""


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_c1_golden_checksums(orc, c1_reads, c1_golden, idx):
    """Sketches and candidate sets of the reference's CI file, three parameter sets."""
    bases, offsets = c1_reads
    g = c1_golden["settings"][idx]
    assert offsets.size - 1 == c1_golden["num_reads"] and int(offsets[-1]) == c1_golden["num_bases"]
    rnd = orc.rand_from_seed(g["seed"], g["n"])
    sk = orc.sketch_all(bases, offsets, g["k"], g["n"], rnd)
    assert "%016x" % orc.fnv_u64(sk.ravel()) == g["fnv_sketches"]
    assert ["%016x" % x for x in sk[4][:3]] == g["sketch_read4_first3"]
    T = orc.build_tables(sk)
    off, ids = T.query_all(bases, offsets, sk, g["k"], rnd, g["thr"], 0)
    assert int(off[-1]) == g["fwd_total"]
    assert "%016x" % orc.fnv_csr(off, ids) == g["fwd_fnv"]
    assert [int(x) for x in ids[int(off[4]):int(off[5])]] == g["fwd_cands_read4"]
    d = np.diff(off.astype(np.int64))
    assert int(d.max()) == g["fwd_max"] and int(d.argmax()) == g["fwd_argmax"]
    offr, idsr = T.query_all(bases, offsets, sk, g["k"], rnd, g["thr"], 1)
    assert int(offr[-1]) == g["rc_total"]
    assert "%016x" % orc.fnv_csr(offr, idsr) == g["rc_fnv"]


def test_edge_set_against_reference_outputs(orc, edge):
    """Full arrays from the reference on empty/short/non-ACGT/duplicate/long reads."""
    bases, offsets = edge["bases"], edge["offsets"]
    for ci, (seed, k, n, thr) in enumerate(edge["cfgs"]):
        k, n, thr = int(k), int(n), int(thr)
        rnd = edge[f"rand_{ci}"]
        assert (orc.rand_from_seed(int(seed), n) == rnd).all()
        sk = orc.sketch_all(bases, offsets, k, n, rnd)
        assert (sk == edge[f"sketches_{ci}"]).all(), f"sketches differ for cfg {ci}"
        T = orc.build_tables(sk)
        for rc, tag in ((0, "fwd"), (1, "rc")):
            off, ids = T.query_all(bases, offsets, sk, k, rnd, thr, rc)
            assert (off == edge[f"{tag}_off_{ci}"]).all(), f"{tag} offsets differ for cfg {ci}"
            assert (ids == edge[f"{tag}_ids_{ci}"]).all(), f"{tag} ids differ for cfg {ci}"


def test_query_string_matches_bulk(orc, edge):
    bases, offsets = edge["bases"], edge["offsets"]
    seed, k, n, thr = (int(v) for v in edge["cfgs"][0])
    rnd = edge["rand_0"]
    sk = edge["sketches_0"]
    T = orc.build_tables(sk)
    off, ids = edge["fwd_off_0"], edge["fwd_ids_0"]
    for i in (0, 4, 50, 100, 164, 170):
        s = bases[int(offsets[i]):int(offsets[i + 1])].tobytes()
        got = T.query_string(s, k, rnd, thr)
        assert (got == ids[int(off[i]):int(off[i + 1])]).all()


needs_ref = pytest.mark.skipif(not RefLib.available(), reason="oracle/_ref/libnsref.so not built")


@needs_ref
def test_reference_static_helpers():
    ref = RefLib.get()
    assert ref.kmer_to_int("ATCG") == 27 and ref.kmer_to_int("GGGG") == 255
    assert list(ref.string2kmers("ACGTTGCAAC", 4)) == [45, 181, 215, 94, 120, 224, 130]


@needs_ref
def test_oracle_equals_reference_on_random_reads(orc):
    """Seeded random read sets through the reference's own code vs the restatement."""
    ref = RefLib.get()
    rng = np.random.default_rng(7)
    genome = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=20000)
    reads = []
    for _ in range(300):
        s = int(rng.integers(0, 19000))
        L = int(rng.integers(0, 1500))
        r = genome[s:s + L].copy()
        flips = rng.random(r.size) < 0.05
        r[flips] = rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=int(flips.sum()))
        reads.append(r.tobytes())
    bases, offsets = reads_to_buffers(reads)
    for seed, k, n, thr in [(5, 23, 60, 6), (6, 12, 20, 2), (8, 31, 16, 1)]:
        rnd = ref.rand_from_seed(seed, n)
        rf = ref.create(bases, offsets, k, n, thr, rnd, threads=2)
        sk = orc.sketch_all(bases, offsets, k, n, rnd)
        assert (sk == rf.sketches()).all()
        T = orc.build_tables(sk)
        for mode in (0, 1):
            off_r, ids_r = rf.query_all(mode, threads=2)
            off_o, ids_o = T.query_all(bases, offsets, sk, k, rnd, thr, mode)
            assert (off_r == off_o).all() and (ids_r == ids_o).all()
        # forward string overload == private sketch overload (ReadFilter.cpp:85-97 vs 65-83)
        off_s, ids_s = rf.query_all(2, threads=2)
        off_o, ids_o = T.query_all(bases, offsets, sk, k, rnd, thr, 0)
        assert (off_s == off_o).all() and (ids_s == ids_o).all()
        rf.close()


@needs_ref
def test_verbatim_initialize_agrees_with_replay(orc):
    """Run the reference's initialize() untouched (random_device), read back the numbers it
    drew, and check the restatement reproduces its answers: the seeded replay loses nothing."""
    ref = RefLib.get()
    rng = np.random.default_rng(11)
    reads = [rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(rng.integers(0, 400))).tobytes()
             for _ in range(120)]
    reads += [reads[3], reads[3][5:], b"", b"ACGTACGTACGTACGTACGTAC"]
    bases, offsets = reads_to_buffers(reads)
    k, n, thr = 23, 60, 6
    rf = ref.create(bases, offsets, k, n, thr, None, threads=2, verbatim=True)
    rnd = rf.rand
    sk = orc.sketch_all(bases, offsets, k, n, rnd)
    T = orc.build_tables(sk)
    off_s, ids_s = rf.query_all(2, threads=2)     # public string overload on every read
    off_o, ids_o = T.query_all(bases, offsets, sk, k, rnd, thr, 0)
    assert (off_s == off_o).all() and (ids_s == ids_o).all()
    off_s, ids_s = rf.query_all(1, threads=2)
    off_o, ids_o = T.query_all(bases, offsets, sk, k, rnd, thr, 1)
    assert (off_s == off_o).all() and (ids_s == ids_o).all()
    rf.close()

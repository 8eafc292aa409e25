"""Property tests (hypothesis, derandomised) of the emulated device code against the oracle: random small
read sets with arbitrary bytes, every k in 1..31, any n up to 70, any filter density and tile size for the
sketch kernels (filter + fix-up, brute force); random id lists and thresholds for the
lookup body.  The oracle is pinned to the unmodified reference, so a counter-example is a kernel bug."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import nanospring_b200 as ns
from oracle.oracle import reads_to_buffers
from test_query_emul import expected, run_count
from test_sketch_emul import sketch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)


@pytest.fixture(scope="module")
def sketch_emul():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "emul"])
    L = C.CDLL(os.path.join(ROOT, "oracle", "libsketch_emul.so"))
    L.sketch_emul_run.argtypes = [u32p, u64p, C.c_uint32, C.c_uint32, C.c_uint32, u64p, C.c_int, C.c_int, C.c_uint32,
                                  C.c_uint, u64p, C.POINTER(C.c_ulonglong)]
    return L


@pytest.fixture(scope="module")
def query_emul():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "emul"])
    return C.CDLL(os.path.join(ROOT, "oracle", "libquery_emul.so"))


COMMON = dict(deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow,
                                                                     HealthCheck.data_too_large])

reads_strategy = st.lists(
    st.one_of(st.binary(min_size=0, max_size=90),                                   # any bytes, around the k-mer length
              st.text(alphabet="ACGT", min_size=0, max_size=400).map(str.encode),
              st.sampled_from([b"A" * 300, b"AC" * 200, b"ACGT" * 20, b"T" * 31, b"G" * 32, b"C" * 33])),
    min_size=1, max_size=8)


@settings(max_examples=120, **COMMON)
@given(reads=reads_strategy, k=st.integers(1, 31), n=st.integers(1, 70), lam=st.integers(0, 8),
       tile_words=st.sampled_from([64, 128, 640, 1024]), seed=st.integers(0, 2**32 - 1), mode=st.sampled_from([0, 1]))
def test_sketch_kernels_any_input(sketch_emul, orc, reads, k, n, lam, tile_words, seed, mode):
    bases, offsets = reads_to_buffers(reads)
    rnd = ns.rand_from_seed(seed, n)
    want = orc.sketch_all(bases, offsets, k, n, rnd)
    got, _ = sketch(sketch_emul, bases, offsets, k, n, rnd, mode=mode, lam=lam, tile_words=tile_words, grid=1)
    assert (got == want).all()


lists_strategy = st.lists(st.lists(st.integers(0, 40), min_size=0, max_size=30), min_size=1, max_size=24)


@settings(max_examples=120, **COMMON)
@given(lists=lists_strategy, thr=st.integers(0, 12), spread=st.sampled_from([1, 3, 50, 100000]))
def test_lookup_body_any_lists(query_emul, lists, thr, spread):
    subs = len(lists)
    qs = [[np.asarray([x * spread for x in l], dtype=np.uint32) for l in lists],
          [np.asarray(l[::-1], dtype=np.uint32) for l in lists]]
    qcount, qpos, tmp, heavy, counters = run_count(query_emul, qs, subs, thr)
    for q, ls in enumerate(qs):
        total = sum(len(l) for l in ls)
        if total > 1024:
            continue
        want = expected(ls, thr)
        got = tmp[int(qpos[q]):int(qpos[q]) + int(qcount[q])]
        assert got.size == want.size and (got == want).all()


# ---------------------------------------------------------------- FASTQ ingest, tables ----
from fastq_cases import expected_packed  # noqa: E402
from test_fastq_emul import run_emul  # noqa: E402
from test_table_emul import query_all  # noqa: E402


@pytest.fixture(scope="module")
def fastq_emul():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "emul"])
    L = C.CDLL(os.path.join(ROOT, "oracle", "libfastq_emul.so"))
    L.fq_emul_parse.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint, C.c_uint32, u64p, C.c_uint64, u32p, C.c_uint64,
                                u32p, u64p]
    return L


@pytest.fixture(scope="module")
def table_emul():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "emul"])
    L = C.CDLL(os.path.join(ROOT, "oracle", "libtable_emul.so"))
    L.table_emul_build.argtypes = [u64p, C.c_uint32, C.c_uint32, C.c_uint]
    L.table_emul_num_keys.argtypes = [C.c_uint32]
    L.table_emul_num_keys.restype = C.c_uint32
    L.table_emul_query.argtypes = [u64p, C.c_uint32, C.c_uint32, C.c_uint, u32p, u64p, u32p, C.c_uint64, u32p,
                                   C.POINTER(C.c_ulonglong)]
    L.table_emul_query.restype = None
    return L


line = st.one_of(st.binary(min_size=0, max_size=40).map(lambda b: b.replace(b"\n", b"N")),
                 st.text(alphabet="ACGTN\r@+", min_size=0, max_size=600).map(str.encode))


@settings(max_examples=80, **COMMON)
@given(lines=st.lists(line, min_size=0, max_size=14), trailing=st.booleans(), misalign=st.integers(0, 15),
       pack_iters=st.sampled_from([0, 4, 16, 64]))
def test_fastq_ingest_any_text(fastq_emul, orc, lines, trailing, misalign, pack_iters):
    """any sequence of lines (any bytes but the newline itself), with or without a final newline, at any
    alignment: read table and packed words as the reference's record loop gives them"""
    text = b"\n".join(lines) + (b"\n" if trailing and lines else b"")
    off, words = run_emul(fastq_emul, text, misalign=misalign, grid=1, pack_iters=pack_iters)
    bases, want_off = orc.fastq_reads(text)
    assert off.size == want_off.size and (off == want_off).all()
    want = expected_packed(bases)
    assert (words[:want.size] == want).all()


@settings(max_examples=25, **COMMON)
@given(rows=st.integers(1, 90), n=st.integers(1, 9), keyspace=st.sampled_from([1, 3, 40, 2**40]), thr=st.integers(1, 4),
       seed=st.integers(0, 2**31))
def test_tables_any_key_multiplicity(table_emul, orc, rows, n, keyspace, thr, seed):
    """sketch matrices whose columns hold one key only, a few keys (large groups), mostly distinct keys, the
    empty marker ~0 and 0: distinct keys per table and the candidate list of every row equal the oracle"""
    rng = np.random.default_rng(seed)
    S = rng.integers(0, keyspace, size=(rows, n), dtype=np.uint64)
    S[rng.random((rows, n)) < 0.1] = np.uint64(0xFFFFFFFFFFFFFFFF)
    S = np.ascontiguousarray(S)
    assert table_emul.table_emul_build(S.ctypes.data_as(u64p), rows, n, 2) == 0
    T = orc.build_tables(S)
    for j in range(n):
        assert table_emul.table_emul_num_keys(j) == T.num_keys(j)
    got = query_all(table_emul, S, thr, grid=1)
    for q in range(rows):
        if got[q] is not None:
            want = T.query_sketch(S[q], thr)
            assert got[q].size == want.size and (got[q] == want).all()


@settings(max_examples=15, **COMMON)
@given(rows=st.integers(1, 300), n=st.integers(1, 9), keyspace=st.sampled_from([1, 3, 40, 2**40]), thr=st.integers(1, 4),
       frac=st.sampled_from([0.0, 0.05, 0.5, 1.0]), seed=st.integers(0, 2**31))
def test_tables_with_deferred_entries(table_emul, orc, rows, n, keyspace, thr, frac, seed):
    """nsmh_sketch_build's table build: any subset of the entries is still all-ones when the insert kernel runs
    and arrives from a list afterwards (table_insert_list_kernel), rows beyond one 256-row work unit included;
    tables and matrix must be those of the plain build"""
    rng = np.random.default_rng(seed)
    S = rng.integers(0, keyspace, size=(rows, n), dtype=np.uint64)
    ones = np.uint64(0xFFFFFFFFFFFFFFFF)
    S[rng.random((rows, n)) < 0.1] = ones
    S = np.ascontiguousarray(S)
    work = S.copy()
    work[rng.random((rows, n)) < frac] = ones
    lst = np.flatnonzero(work.ravel() == ones).astype(np.uint32)
    lst = np.ascontiguousarray(np.concatenate([lst[rng.permutation(lst.size)], np.zeros(1, np.uint32)]))
    vals = np.ascontiguousarray(S.ravel()[lst])
    table_emul.table_emul_build_deferred.argtypes = [u64p, C.c_uint32, C.c_uint32, C.c_uint, u32p, C.c_uint, u64p]
    assert table_emul.table_emul_build_deferred(work.ctypes.data_as(u64p), rows, n, 2, lst.ctypes.data_as(u32p), lst.size - 1,
                                                vals.ctypes.data_as(u64p)) == 0
    assert (work == S).all()
    T = orc.build_tables(S)
    for j in range(n):
        assert table_emul.table_emul_num_keys(j) == T.num_keys(j)
    got = query_all(table_emul, S, thr, grid=1)
    for q in range(rows):
        if got[q] is not None:
            want = T.query_sketch(S[q], thr)
            assert got[q].size == want.size and (got[q] == want).all()

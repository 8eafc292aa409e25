"""Property tests (hypothesis, derandomised) of the emulated device code against the oracle: random small
read sets with arbitrary bytes, every k in 1..31, any n up to 70, any filter density and tile size for the
sketch kernels (filter + fix-up, balanced variant, brute force); random id lists and thresholds for the
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
       tile_words=st.sampled_from([64, 128, 640, 1024]), seed=st.integers(0, 2**32 - 1), mode=st.sampled_from([0, 1, 2]))
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

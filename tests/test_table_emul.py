"""The table build (nanospring_b200/csrc/table_kernels.cuh: insert with one 128-bit compare-and-swap per
new key, group ranges, member fill) and the probing lookup (csrc/query_kernels.cuh: count_body<ProbeSrc>)
compiled for the HOST and run as real concurrent threads (tests/cpp/cuda_host_shim.h, whole 256-thread
blocks), compared with the oracle's dictionary (BBHashMap semantics: exact key -> list of read ids,
BBHashMap.cpp:10-120) and candidate lists (ReadFilter.cpp:65-83).  A logic check of the device code for
the container without a GPU; the GPU parity proper is tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import nanospring_b200 as ns

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "libtable_emul.so")
u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)


@pytest.fixture(scope="module")
def emul():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "emul"])
    L = C.CDLL(SO)
    L.table_emul_build.argtypes = [u64p, C.c_uint32, C.c_uint32, C.c_uint]
    L.table_emul_num_keys.argtypes = [C.c_uint32]
    L.table_emul_num_keys.restype = C.c_uint32
    L.table_emul_query.argtypes = [u64p, C.c_uint32, C.c_uint32, C.c_uint, u32p, u64p, u32p, C.c_uint64, u32p,
                                   C.POINTER(C.c_ulonglong)]
    L.table_emul_query.restype = None
    return L


def query_all(L, qsk, thr, grid=2):
    nq, n = qsk.shape
    cap = 16 * nq + max(nq * 64, nq * nq) + 1024   # fixed places + every row can match every row of a small table
    qcount = np.zeros(nq + 1, dtype=np.uint32)
    qpos = np.zeros(nq, dtype=np.uint64)
    tmp = np.zeros(cap, dtype=np.uint32)
    heavy = np.zeros(nq + 1, dtype=np.uint32)
    counters = (C.c_ulonglong * 4)(0, 0, 0, 0)
    L.table_emul_query(np.ascontiguousarray(qsk).ctypes.data_as(u64p), nq, thr, grid, qcount.ctypes.data_as(u32p),
                       qpos.ctypes.data_as(u64p), tmp.ctypes.data_as(u32p), cap, heavy.ctypes.data_as(u32p), counters)
    assert 16 * nq + counters[3] <= cap         # the fixed places + the larger result lists
    out = []
    handed_on = set(int(x) for x in heavy[:counters[0]])
    for q in range(nq):
        out.append(None if q in handed_on else tmp[int(qpos[q]):int(qpos[q]) + int(qcount[q])].copy())
    return out


def sketch_matrix(orc, k, n, seed, n_reads=200, mean=1200, dup=True):
    """sketches of synthetic reads at low error (real overlaps) plus duplicated reads (groups of many
    sizes in every table), short reads (the all-zero / all-ones rows of SURVEY S5) and the key ~0"""
    rnd = ns.rand_from_seed(seed, n)
    lengths = ns.synth_lengths(n_reads, mean, seed=seed)
    lengths[:6] = [0, 1, k - 2 if k > 2 else 0, k - 1, k - 1, k]
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=60_000, genome_seed=seed, read_seed=seed + 1,
                                                      p_ins=0.01, p_del=0.01, p_sub=0.01))
    sk = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
    if dup:
        reps = np.concatenate([np.repeat(np.arange(10, 20), 3), np.repeat(np.arange(40, 43), 24), np.arange(n_reads)])
        sk = sk[np.random.default_rng(seed).permutation(reps)]
    return np.ascontiguousarray(sk)


@pytest.mark.parametrize("k,n,thr,blocks", [(23, 60, 6, 5), (15, 30, 3, 1), (31, 7, 2, 64), (9, 33, 4, 3)])
def test_tables_and_lookup_equal_oracle(emul, orc, k, n, thr, blocks):
    sk = sketch_matrix(orc, k, n, seed=k + n)
    rows = sk.shape[0]
    assert emul.table_emul_build(sk.ctypes.data_as(u64p), rows, n, blocks) == 0
    T = orc.build_tables(sk)
    for j in range(n):
        assert emul.table_emul_num_keys(j) == T.num_keys(j), f"distinct keys of table {j}"
    got = query_all(emul, sk, thr)
    resolved = 0
    for q in range(rows):
        if got[q] is None:                                  # more than 1024 gathered ids: the global path's job
            continue
        want = T.query_sketch(sk[q], thr)
        assert got[q].size == want.size and (got[q] == want).all(), f"query {q}"
        resolved += 1
    assert resolved > rows // 2
    # foreign sketches: rows that are not in the tables, incl. keys that collide with nothing
    rng = np.random.default_rng(1)
    foreign = sk[rng.integers(0, rows, size=40)].copy()
    foreign[:, ::2] ^= np.uint64(0x9E3779B97F4A7C15)
    foreign[0, :] = np.uint64(0xFFFFFFFFFFFFFFFF)            # probes the extra slot of every table
    foreign[1, :] = 0
    for q, g in enumerate(query_all(emul, foreign, max(thr // 2, 1), grid=1)):
        if g is not None:
            want = T.query_sketch(foreign[q], max(thr // 2, 1))
            assert g.size == want.size and (g == want).all(), f"foreign query {q}"


@pytest.mark.parametrize("k,n,thr,blocks,frac", [(23, 60, 6, 3, 0.02), (15, 30, 3, 2, 0.3), (31, 7, 2, 8, 1.0)])
def test_build_with_deferred_entries_equals_plain_build(emul, orc, k, n, thr, blocks, frac):
    """nsmh_sketch_build: the insert kernel leaves the all-ones entries out (skip_empty), table_insert_list_kernel
    stores and inserts them from a list afterwards - pending sketch values and the legitimate all-ones entries of
    the len == k-1 reads alike.  Same distinct keys, same answers as the plain build of the final matrix."""
    sk = sketch_matrix(orc, k, n, seed=k + n + 1)
    rows = sk.shape[0]
    rng = np.random.default_rng(k)
    ones = np.uint64(0xFFFFFFFFFFFFFFFF)
    pending = rng.random(sk.shape) < frac
    assert (sk == ones).any(), "the set must hold legitimate all-ones entries too"
    work = sk.copy()
    work[pending] = ones
    lst = np.flatnonzero(work.ravel() == ones).astype(np.uint32)           # what sketch_missing_kernel lists
    lst = lst[rng.permutation(lst.size)]
    vals = np.ascontiguousarray(sk.ravel()[lst])
    L = emul
    L.table_emul_build_deferred.argtypes = [u64p, C.c_uint32, C.c_uint32, C.c_uint, u32p, C.c_uint, u64p]
    assert L.table_emul_build_deferred(work.ctypes.data_as(u64p), rows, n, blocks, lst.ctypes.data_as(u32p), lst.size,
                                       vals.ctypes.data_as(u64p)) == 0
    assert (work == sk).all(), "the list kernel stores the values into the sketch matrix"
    T = orc.build_tables(sk)
    for j in range(n):
        assert L.table_emul_num_keys(j) == T.num_keys(j), f"distinct keys of table {j}"
    got = query_all(L, sk, thr)
    resolved = 0
    for q in range(rows):
        if got[q] is None:
            continue
        want = T.query_sketch(sk[q], thr)
        assert got[q].size == want.size and (got[q] == want).all(), f"query {q}"
        resolved += 1
    assert resolved > rows // 2


def test_empty_and_tiny_tables(emul, orc):
    n = 8
    sk = np.zeros((0, n), dtype=np.uint64)
    assert emul.table_emul_build(sk.ctypes.data_as(u64p), 0, n, 4) == 0
    assert emul.table_emul_num_keys(0) == 0
    q = np.zeros((2, n), dtype=np.uint64)
    assert all(g is not None and g.size == 0 for g in query_all(emul, q, 1))
    sk = np.array([[5] * n, [5] * n, [np.uint64(0xFFFFFFFFFFFFFFFF)] * n], dtype=np.uint64)
    assert emul.table_emul_build(sk.ctypes.data_as(u64p), 3, n, 4) == 0
    assert emul.table_emul_num_keys(3) == 2
    got = query_all(emul, sk, n)
    assert got[0].tolist() == [0, 1] and got[1].tolist() == [0, 1] and got[2].tolist() == [2]

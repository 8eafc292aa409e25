"""The counting-filter tier of the candidate lookup (nanospring_b200/csrc/query_kernels.cuh) compiled for
the HOST and run in lock step (tests/cpp/cuda_host_shim.h), compared with a plain sort-and-count of
the gathered ids - the definition in ReadFilter.cpp:65-83.  A logic check of the device code for the
container without a GPU; the GPU parity proper is tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "libquery_emul.so")
u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)


@pytest.fixture(scope="module")
def emul():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "emul"])
    L = C.CDLL(SO)
    L.mid_emul_run.argtypes = [u64p, u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint, u32p, u64p, u32p, C.c_uint64,
                               u32p, C.POINTER(C.c_ulonglong)]
    L.mid_emul_run.restype = None
    b, cap, mx = C.c_uint32(), C.c_uint32(), C.c_uint32()
    L.mid_emul_constants(C.byref(b), C.byref(cap), C.byref(mx))
    L.consts = (b.value, cap.value, mx.value)
    return L


def run(L, queries, subs, thr, grid=2):
    """queries: list of lists of id arrays (subs lists per query).  Returns per query either the
    emitted ids (np.uint32, in the order stored) or None when the tier handed the query on."""
    nq = len(queries)
    flat, off = [], [0]
    for lists in queries:
        assert len(lists) == subs
        for l in lists:
            flat.append(np.asarray(l, dtype=np.uint32))
            off.append(off[-1] + len(l))
    ids = np.concatenate(flat + [np.zeros(1, np.uint32)])
    off = np.asarray(off, dtype=np.uint64)
    total = int(off[-1])
    qcount = np.zeros(nq, dtype=np.uint32)
    qpos = np.full(nq, np.iinfo(np.uint64).max, dtype=np.uint64)
    mid_cap = total // max(thr, 1) + 1
    mid_ids = np.full(mid_cap + 8, 0xDEADBEEF, dtype=np.uint32)
    unresolved = np.full(nq + 1, 0xFFFFFFFF, dtype=np.uint32)
    counters = (C.c_ulonglong * 3)(0, 0, 0)
    L.mid_emul_run(off.ctypes.data_as(u64p), ids.ctypes.data_as(u32p), nq, subs, thr, grid,
                   qcount.ctypes.data_as(u32p), qpos.ctypes.data_as(u64p), mid_ids.ctypes.data_as(u32p), mid_cap,
                   unresolved.ctypes.data_as(u32p), counters)
    assert counters[2] == 0, "mid_ids overflowed although its size is the proven bound"
    assert (mid_ids[mid_cap:] == 0xDEADBEEF).all()
    handed_on = set(int(x) for x in unresolved[:counters[0]])
    assert len(handed_on) == counters[0] and (unresolved[counters[0]:] == 0xFFFFFFFF).all()
    out, used = [], 0
    for q in range(nq):
        if q in handed_on:
            assert qcount[q] == 0 and qpos[q] == np.iinfo(np.uint64).max     # untouched for the global path
            out.append(None)
            continue
        pos = int(qpos[q])
        assert pos >> 63 == 1, "resolved queries carry the mid flag"
        pos &= (1 << 63) - 1
        out.append(mid_ids[pos:pos + int(qcount[q])].copy())
        used += int(qcount[q])
    assert used == counters[1]
    return out


def expected(lists, thr):
    allids = np.concatenate([np.asarray(l, dtype=np.uint32) for l in lists] + [np.zeros(0, np.uint32)])
    v, c = np.unique(allids, return_counts=True)
    return v[c >= max(thr, 1)].astype(np.uint32)


def must_hand_on(L, lists, thr):
    """exactly the tier's own rule: too many ids, or too many ids in buckets that reach the threshold"""
    buckets, cap, max_ids = L.consts
    allids = np.concatenate([np.asarray(l, dtype=np.uint32) for l in lists] + [np.zeros(0, np.uint32)])
    if allids.size > max_ids:
        return True
    b = ((allids.astype(np.uint64) * 0x9E3779B1) & 0xFFFFFFFF) >> 20
    assert buckets == 4096
    cnt = np.bincount(b.astype(np.int64), minlength=buckets)
    return int((cnt[b.astype(np.int64)] >= max(thr, 1)).sum()) > cap


def check(L, queries, subs, thr, grid=2):
    got = run(L, queries, subs, thr, grid)
    for q, lists in enumerate(queries):
        if must_hand_on(L, lists, thr):
            assert got[q] is None, f"query {q} should have been handed to the global path"
        else:
            assert got[q] is not None, f"query {q} should have been resolved"
            want = expected(lists, thr)
            assert got[q].size == want.size and (got[q] == want).all(), f"query {q}"


def chance_query(rng, subs, list_len, true_ids, hits_per_true):
    """lists of unrelated ids (chance collisions) + a few ids present in hits_per_true of the lists"""
    lists = [list(rng.integers(0, 2_000_000, size=int(rng.integers(0, list_len + 1)))) for _ in range(subs)]
    for t in true_ids:
        for j in rng.choice(subs, size=min(hits_per_true, subs), replace=False):
            lists[j].append(t)
    return [np.asarray(rng.permutation(l), dtype=np.uint32) for l in lists]


@pytest.mark.parametrize("subs,thr", [(60, 6), (120, 12), (30, 3), (7, 1), (33, 2)])
def test_chance_collisions_and_true_candidates(emul, subs, thr):
    rng = np.random.default_rng(subs * 100 + thr)
    qs = [chance_query(rng, subs, L, rng.integers(0, 4_000_000_000, size=t), h)
          for L, t, h in ((40, 3, subs), (120, 20, thr), (300, 0, 0), (3, 1, thr - 1 if thr > 1 else 1), (0, 2, subs),
                          (1, 5, thr + 1))]
    check(emul, qs, subs, thr, grid=int(rng.integers(1, 4)))


def test_empty_singleton_and_sentinel_like_ids(emul):
    subs = 12
    big = 0xFFFFFFFE                                    # largest legal read id
    qs = [[np.zeros(0, np.uint32)] * subs,
          [np.asarray([5], np.uint32)] * subs,          # singletons in both representations
          [np.asarray([big, 0], np.uint32)] * subs,
          [np.asarray([7] * 9, np.uint32)] + [np.zeros(0, np.uint32)] * (subs - 1)]     # one list repeats an id
    for thr in (1, 2, 9, 12, 13):
        check(emul, qs, subs, thr, grid=1)


def test_hand_on_rules(emul):
    """more than 65535 ids, or more than kMidCap ids above the threshold, go to the global path; the
    query next to them is still resolved"""
    buckets, cap, max_ids = emul.consts
    rng = np.random.default_rng(3)
    subs = 40
    too_many = [rng.integers(0, 1 << 31, size=max_ids // subs + 2).astype(np.uint32) for _ in range(subs)]
    crowd = [np.arange(1000, 1000 + 400, dtype=np.uint32)] * subs        # 400 ids x 40 lists, all above thr
    fits = [np.arange(50, 50 + cap // subs, dtype=np.uint32)] * subs      # exactly kMidCap-ish survivors
    normal = chance_query(rng, subs, 60, [11, 12, 13], 10)
    qs = [too_many, normal, crowd, fits]
    assert must_hand_on(emul, too_many, 5) and must_hand_on(emul, crowd, 5) and not must_hand_on(emul, fits, 5)
    check(emul, qs, subs, 5, grid=2)


@pytest.mark.parametrize("thr", [1, 2, 6, 12, 60])
def test_in_place_filter_of_the_warp_sort_path(emul, thr):
    """count_kernel's sort path (up to 1024 gathered ids in the warp buffer): with the counting
    filter in front of the bitonic sort the emitted ids are the same as without it, and equal to
    sort-and-count; on chance-collision inputs the filter leaves a small fraction to sort."""
    emul.sortpath_emul_run.argtypes = [u32p, C.c_uint32, C.c_uint32, C.c_int, u32p, u32p, u32p]
    emul.sortpath_emul_run.restype = None
    rng = np.random.default_rng(100 + thr)
    cases = []
    for T in (0, 1, 31, 64, 65, 257, 600, 1000, 1024):
        ids = rng.integers(0, 3_000_000, size=T).astype(np.uint32)          # chance collisions only
        cases.append(ids)
        if T >= 65:
            true = rng.integers(0, 0xFFFFFFFE, size=4, dtype=np.uint64).astype(np.uint32)
            reps = np.repeat(true, [max(thr, 1), max(thr - 1, 1), min(3 * max(thr, 1), 60), 1])[:T // 2]
            mixed = np.concatenate([ids[:T - reps.size], reps])
            cases.append(rng.permutation(mixed).astype(np.uint32))
    cases.append(np.full(1024, 77, dtype=np.uint32))                        # one id only
    cases.append(np.repeat(np.arange(64, dtype=np.uint32), 16))             # everything qualifies for thr <= 16
    for ids in cases:
        v, c = np.unique(ids, return_counts=True)
        want = v[c >= max(thr, 1)].astype(np.uint32)
        for use_filter in (1, 0):
            out = np.zeros(1024 + 8, dtype=np.uint32)
            R, S = C.c_uint32(0), C.c_uint32(0)
            emul.sortpath_emul_run(np.ascontiguousarray(ids).ctypes.data_as(u32p), ids.size, thr, use_filter,
                                   out.ctypes.data_as(u32p), C.byref(R), C.byref(S))
            assert R.value == want.size and (out[:R.value] == want).all(), (ids.size, thr, use_filter)
            if use_filter and thr > 1 and ids.size > 64:
                b = (((ids.astype(np.uint64) * 0x9E3779B1) & 0xFFFFFFFF) >> 23).astype(np.int64)
                kept = int((np.bincount(b, minlength=512)[b] >= thr).sum())
                assert S.value == kept, "the filter keeps exactly the ids of the buckets that reach the threshold"
                if thr >= 6 and c.max() == 1 and ids.size >= 600:
                    assert kept <= ids.size // 4, "chance collisions are mostly filtered out"
            else:
                assert S.value == ids.size


# ------------------------------------------------------------------ the lookup kernel's body --
def run_count(L, queries, subs, thr, grid=1, tmp_cap=None, wide=False):
    L.count_emul_set_wide.argtypes = [C.c_int]
    L.count_emul_set_wide(1 if wide else 0)
    L.count_emul_run.argtypes = [u64p, u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint, u32p, u64p, u32p, C.c_uint64,
                                 u32p, C.POINTER(C.c_ulonglong)]
    L.count_emul_run.restype = None
    nq = len(queries)
    flat, off = [], [0]
    for lists in queries:
        assert len(lists) == subs
        for l in lists:
            flat.append(np.asarray(l, dtype=np.uint32))
            off.append(off[-1] + len(l))
    ids = np.concatenate(flat + [np.zeros(1, np.uint32)])
    off = np.asarray(off, dtype=np.uint64)
    L.lookup_fixed_ids.restype = C.c_uint32
    fixed = nq * L.lookup_fixed_ids()       # every query owns kFixedIds entries, larger result lists follow
    cap = fixed + (int(off[-1]) + 1 if tmp_cap is None else tmp_cap)
    qcount = np.full(nq + 1, 0xEEEEEEEE, dtype=np.uint32)
    qpos = np.full(nq, 0x1234, dtype=np.uint64)
    tmp = np.full(cap + 8, 0xDEADBEEF, dtype=np.uint32)
    heavy = np.full(nq + 1, 0xFFFFFFFF, dtype=np.uint32)
    counters = (C.c_ulonglong * 4)(0, 0, 0, 0)
    L.count_emul_run(off.ctypes.data_as(u64p), ids.ctypes.data_as(u32p), nq, subs, thr, grid, qcount.ctypes.data_as(u32p),
                     qpos.ctypes.data_as(u64p), tmp.ctypes.data_as(u32p), cap, heavy.ctypes.data_as(u32p), counters)
    assert (tmp[cap:] == 0xDEADBEEF).all(), "wrote past tmp_cap"
    return qcount, qpos, tmp, heavy, counters


@pytest.mark.parametrize("subs,thr", [(60, 6), (60, 1), (24, 2), (120, 12), (7, 7), (33, 0), (128, 3), (200, 4)])
def test_lookup_kernel_body_all_paths(emul, subs, thr):
    """count_body: counting in registers with warp votes (<= 128 gathered ids: straight from the probes when
    every list is a single id, through the buffer otherwise), filter + sort path (<= 1024), hand-over of
    larger queries, more than 128 lists per query (nothing kept in registers); results per query ascending
    and equal to sort-and-count; the counters add up."""
    rng = np.random.default_rng(subs * 7 + thr)
    specs = [(0, 0, 0), (1, 1, subs), (3, 4, max(thr, 1)), (4, 40, max(thr, 1) + 1), (8, 3, subs // 2 + 1), (12, 0, 0),
             (16, 6, max(thr, 1)), (30, 2, subs), (90, 1, 1), (2, 45, subs), (1, 70, 3)]
    qs = [chance_query(rng, subs, L, rng.integers(0, 0xFFFFFFFE, size=t, dtype=np.uint64), min(h, subs)) for L, t, h in specs]
    qs.append([np.asarray([9], np.uint32)] * subs)                          # singletons only (inlined and pointed-to)
    qs.append([np.arange(20, dtype=np.uint32)] * subs if subs * 20 <= 1024 else [np.arange(5, dtype=np.uint32)] * subs)
    qcount, qpos, tmp, heavy, counters = run_count(emul, qs, subs, thr, grid=int(rng.integers(1, 3)))
    T = [sum(len(l) for l in lists) for lists in qs]
    want_heavy = sorted(q for q, t in enumerate(T) if t > 1024)
    assert sorted(int(x) for x in heavy[:counters[0]]) == want_heavy
    assert counters[1] == sum(T)
    paths = set()
    emitted = 0
    for q, lists in enumerate(qs):
        if T[q] > 1024:
            assert qcount[q] == 0 and qpos[q] == np.iinfo(np.uint64).max
            paths.add("handed on")
            continue
        want = expected(lists, thr)
        got = tmp[int(qpos[q]):int(qpos[q]) + int(qcount[q])]
        assert got.size == want.size and (got == want).all(), f"query {q} (T = {T[q]})"
        emitted += want.size
        paths.add("registers" if T[q] <= 128 else "sort")
    assert counters[2] == emitted
    assert {"registers", "sort"} <= paths


def test_lookup_kernel_body_wide_buffer(emul):
    """The n > 64 instantiation (count_body<Src, 4, kLookupCapWide>): queries with 1024 < T <= 2048 gathered ids
    are resolved by the sort path instead of being handed on; beyond 2048 they still are."""
    rng = np.random.default_rng(77)
    subs, thr = 120, 12
    qs = [chance_query(rng, subs, 20, rng.integers(0, 0xFFFFFFFE, size=3, dtype=np.uint64), subs),       # ~1200 + 360
          chance_query(rng, subs, 25, rng.integers(0, 0xFFFFFFFE, size=2, dtype=np.uint64), thr),        # ~1500
          chance_query(rng, subs, 30, [], 0),                                                            # ~1800
          chance_query(rng, subs, 45, rng.integers(0, 0xFFFFFFFE, size=1, dtype=np.uint64), subs),       # > 2048
          chance_query(rng, subs, 3, rng.integers(0, 0xFFFFFFFE, size=5, dtype=np.uint64), thr + 1),
          [np.asarray([4], np.uint32)] * subs]
    T = [sum(len(l) for l in lists) for lists in qs]
    assert sum(1024 < t <= 2048 for t in T) >= 3 and any(t > 2048 for t in T), T
    qcount, qpos, tmp, heavy, counters = run_count(emul, qs, subs, thr, grid=2, wide=True)
    assert sorted(int(x) for x in heavy[:counters[0]]) == [q for q, t in enumerate(T) if t > 2048]
    assert counters[1] == sum(T)
    for q, lists in enumerate(qs):
        if T[q] > 2048:
            assert qcount[q] == 0 and qpos[q] == np.iinfo(np.uint64).max
            continue
        want = expected(lists, thr)
        got = tmp[int(qpos[q]):int(qpos[q]) + int(qcount[q])]
        assert got.size == want.size and (got == want).all(), f"query {q} (T = {T[q]})"
    # the narrow instantiation hands the same three queries on
    qcount, qpos, tmp, heavy, counters = run_count(emul, qs, subs, thr, grid=2, wide=False)
    assert sorted(int(x) for x in heavy[:counters[0]]) == [q for q, t in enumerate(T) if t > 1024]


def test_lookup_kernel_body_result_buffer_too_small(emul):
    """The result cursor keeps counting past tmp_cap and nothing is written there: the host retries once
    with the exact size (count_and_emit in query.cu)."""
    rng = np.random.default_rng(8)
    subs, thr = 20, 1
    qs = [chance_query(rng, subs, 10, [], 0) for _ in range(6)]
    total = sum(expected(l, thr).size for l in qs)
    big = sum(expected(l, thr).size for l in qs if expected(l, thr).size > emul.lookup_fixed_ids())
    assert big > 0, "the case needs result lists beyond the fixed places"
    qcount, qpos, tmp, heavy, counters = run_count(emul, qs, subs, thr, tmp_cap=big // 3)
    assert counters[2] == total and counters[3] == big and counters[0] == 0
    qcount, qpos, tmp, heavy, counters = run_count(emul, qs, subs, thr, tmp_cap=big)
    for q, lists in enumerate(qs):
        assert (tmp[int(qpos[q]):int(qpos[q]) + int(qcount[q])] == expected(lists, thr)).all()

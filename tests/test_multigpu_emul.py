"""A whole world of ranks of the peer-memory multi-GPU path (csrc/multigpu.cu: tables partitioned by hash
function, sketch columns scattered to the table owners, probe results and small groups stored into the
read owners' arenas, larger groups read from the owners) run on the HOST with the device code compiled
for it (tests/cpp/multigpu_host_emul.cpp): every rank's candidate lists must equal the oracle's for the
global read set, for any number of ranks, uneven column and row splits, and when the inboxes overflow.
A logic check for the container without a GPU; on GPUs: tests/test_gpu_peer.py, tests/test_gpu_multi.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_table_emul import sketch_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "libmultigpu_emul.so")
u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)


@pytest.fixture(scope="module")
def emul():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "emul"])
    L = C.CDLL(SO)
    L.mg_emul_run.argtypes = [u64p, u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_longlong, C.c_uint, u32p, u64p, u32p,
                              C.c_uint64, u32p, C.POINTER(C.c_ulonglong), u32p]
    return L


def run_world(L, S, rows, thr, inbox_cap=0, grid=2):
    world = len(rows)
    total, n = S.shape
    assert sum(rows) == total
    rows_a = np.asarray(rows, dtype=np.uint32)
    tmp_cap = total * 64 + 1024
    qcount = np.zeros(total + world, dtype=np.uint32)
    qpos = np.zeros(max(total, 1), dtype=np.uint64)
    tmp = np.zeros(world * tmp_cap, dtype=np.uint32)
    heavy = np.zeros(total + world, dtype=np.uint32)
    counters = (C.c_ulonglong * (4 * world))()
    col_end = np.zeros(16, dtype=np.uint32)
    rc = L.mg_emul_run(np.ascontiguousarray(S).ctypes.data_as(u64p), rows_a.ctypes.data_as(u32p), world, n, thr, inbox_cap,
                       grid, qcount.ctypes.data_as(u32p), qpos.ctypes.data_as(u64p), tmp.ctypes.data_as(u32p), tmp_cap,
                       heavy.ctypes.data_as(u32p), counters, col_end.ctypes.data_as(u32p))
    if rc:
        return rc, None, None
    out, row0, qc = [], 0, 0
    for r in range(world):
        c = counters[4 * r:4 * r + 4]
        assert 16 * rows[r] + c[3] <= tmp_cap        # the fixed places + the larger result lists
        handed_on = set(int(x) for x in heavy[qc:qc + c[0]])
        for q in range(rows[r]):
            if q in handed_on:
                out.append(None)
            else:
                p = r * tmp_cap + int(qpos[row0 + q])
                out.append(tmp[p:p + int(qcount[qc + q])].copy())
        row0 += rows[r]
        qc += rows[r] + 1
    return 0, out, col_end[:world].tolist()


def check_world(L, orc, S, rows, thr, **kw):
    rc, got, col_end = run_world(L, S, rows, thr, **kw)
    assert rc == 0
    T = orc.build_tables(S)
    resolved = 0
    for q in range(S.shape[0]):
        if got[q] is None:
            continue
        want = T.query_sketch(S[q], thr)
        assert got[q].size == want.size and (got[q] == want).all(), f"global row {q}, ranks {rows}"
        resolved += 1
    assert resolved > S.shape[0] // 2
    return col_end


@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
def test_world_of_ranks_equals_oracle(emul, orc, world):
    n, thr = 60, 6
    S = sketch_matrix(orc, 23, n, seed=world + 40, n_reads=180)
    total = S.shape[0]
    rng = np.random.default_rng(world)
    cuts = np.sort(rng.choice(np.arange(1, total), size=world - 1, replace=False)) if world > 1 else np.zeros(0, int)
    rows = np.diff(np.concatenate([[0], cuts, [total]])).astype(int).tolist()
    col_end = check_world(emul, orc, S, rows, thr)
    widths = np.diff([0] + col_end)
    assert col_end[-1] == n and widths.max() - widths.min() <= 1            # as even as n allows (8/8/8/8/7/7/7/7)
    if world in (8,):
        assert len(set(widths.tolist())) == 2


def test_odd_hash_count_empty_rank_and_inbox_overflow(emul, orc):
    """n = 30 (not a multiple of 4: scalar key loads, odd column blocks), a rank without
    reads, and inboxes far too small for the groups (they are then read from the owners)."""
    S = sketch_matrix(orc, 15, 30, seed=9, n_reads=150)
    total = S.shape[0]
    check_world(emul, orc, S, [total // 3, 0, total - total // 3], 3)
    check_world(emul, orc, S, [total // 2, total - total // 2], 3, inbox_cap=48)
    check_world(emul, orc, S, [7, total - 7], 3, inbox_cap=1)


def test_more_ranks_than_hash_functions_is_refused(emul):
    S = np.zeros((4, 3), dtype=np.uint64)
    rc, _, _ = run_world(emul, S, [1, 1, 1, 1], 1)
    assert rc == -1

"""Multi-rank path over peer memory (csrc/multigpu.cu, nsmh_mg_*) on a single-GPU box:
world = 1 in-process (all kernels of the path, arena = local memory), and two ranks as two
processes sharing device 0 (real cudaIpc mappings + device-side flag barriers)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("k,n,thr", [(23, 60, 6), (15, 30, 3), (31, 7, 2)])
def test_world_of_one_equals_plain_path_and_oracle(k, n, thr):
    import nanospring_b200 as ns
    from nanospring_b200 import shard
    from oracle.oracle import Oracle
    rnd = ns.rand_from_seed(5, n)
    lengths = ns.synth_lengths(1500, 2000, seed=3)
    lengths[:6] = [0, 1, k - 2, k - 1, k, 30000]
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=200_000, p_ins=0.01, p_del=0.01, p_sub=0.01))
    f = ns.MinHashReadFilter(device=0)
    f.k, f.n, f.overlapSketchThreshold, f.randNumbers = k, n, thr, rnd
    f.initialize(rd)
    off0, ids0 = f.queryAll(False)
    peer = shard.PeerPartitionedFilter(f, 0, 1, [rd.numReads])
    for _ in range(2):
        f.sketch()
        total = peer.run()
        off, ids = peer.result(rd.numReads, total)
        assert (off == off0).all() and (ids == ids0).all()
    assert set(peer.last_ms) == {"scatter_columns", "barrier_1", "build_owned_tables", "probe_to_peers",
                                 "barrier_2", "count"}
    peer.shutdown()
    # the plain path still works on the same handle afterwards
    f.sketch()
    f.build()
    off1, ids1 = f.queryAll(False)
    assert (off1 == off0).all() and (ids1 == ids0).all()
    f.close()
    orc = Oracle.get()
    want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
    woff, wids = orc.build_tables(want).query_all(rd.bases, rd.offsets, want, k, rnd, thr, 0)
    assert (off0 == woff).all() and (ids0 == wids).all()


@pytest.mark.parametrize("inbox_cap", [None, "48"])
def test_groups_of_every_size_inbox_and_remote_paths(inbox_cap, monkeypatch):
    """Duplicated reads give groups of 2..100 members in every table: small groups travel through
    the inbox, groups above kInboxMaxGroup (32) and everything beyond the inbox capacity are read
    from the owner's ids array; all of it must agree with the plain single-GPU path."""
    import nanospring_b200 as ns
    from nanospring_b200 import shard
    if inbox_cap:
        monkeypatch.setenv("NSMH_MG_INBOX_CAP", inbox_cap)
    k, n, thr = 23, 60, 6
    rnd = ns.rand_from_seed(9, n)
    base = ns.synth_reads_host(ns.synth_lengths(40, 1500, seed=8), ns.synth_params(genome_len=100_000))
    reads = []
    for i in range(40):
        reads += [base.getRead(i)] * (1 + (i * 7) % 100 if i % 3 == 0 else 1 + i % 4)
    rd = ns.ReadData.from_reads(reads)
    f = ns.MinHashReadFilter(device=0)
    f.k, f.n, f.overlapSketchThreshold, f.randNumbers = k, n, thr, rnd
    f.initialize(rd)
    off0, ids0 = f.queryAll(False)
    assert int(np.diff(off0.astype(np.int64)).max()) > 32
    peer = shard.PeerPartitionedFilter(f, 0, 1, [rd.numReads])
    f.sketch()
    total = peer.run()
    off, ids = peer.result(rd.numReads, total)
    assert (off == off0).all() and (ids == ids0).all()
    peer.shutdown()
    f.close()


def test_mg_argument_checks():
    import ctypes as C
    import nanospring_b200 as ns
    from nanospring_b200._lib import MG_TOKEN_BYTES, NSMH_EINVAL, NSMH_ESTATE, lib, u32p
    f = ns.MinHashReadFilter(device=0)
    f.k, f.n, f.overlapSketchThreshold, f.randNumbers = 23, 8, 2, ns.rand_from_seed(1, 8)
    f._create()
    tok = (C.c_uint8 * MG_TOKEN_BYTES)()
    rows = np.array([4] * 16, dtype=np.uint32)
    assert lib().nsmh_mg_run(f._h, None) == NSMH_ESTATE                      # not initialised
    assert lib().nsmh_mg_init(f._h, 0, 17, rows.ctypes.data_as(u32p), tok) == NSMH_EINVAL
    assert lib().nsmh_mg_init(f._h, 3, 2, rows.ctypes.data_as(u32p), tok) == NSMH_EINVAL
    assert lib().nsmh_mg_init(f._h, 0, 16, rows.ctypes.data_as(u32p), tok) == NSMH_EINVAL   # 16 ranks, 8 hashes
    assert lib().nsmh_mg_init(f._h, 0, 1, rows.ctypes.data_as(u32p), tok) == 0
    assert lib().nsmh_mg_connect(f._h, bytes(MG_TOKEN_BYTES)) == NSMH_EINVAL      # garbage token
    assert lib().nsmh_mg_connect(f._h, bytes(tok)) == 0
    assert lib().nsmh_mg_run(f._h, None) == NSMH_ESTATE                      # nothing sketched
    f.close()


@pytest.mark.parametrize("world,n", [(2, 60), (2, 30), (4, 60)])
def test_ranks_sharing_one_device(world, n):
    """world processes, each with its own CUDA context on device 0: real cudaIpc mappings and the
    device-side flag barriers; with 4 ranks the 15 column units of n = 60 split unevenly (4, 4, 4, 3).
    (8 ranks: profiles/r1_peer_same_device_w8_s8.log, tools/gpu_session.sh peer.)"""
    env = dict(os.environ, NSMH_TEST_N=str(n), NSMH_MG_TIMEOUT_MS="60000")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29541 + n + world), os.path.join(ROOT, "tools", "peer_same_device.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "OK" in r.stdout

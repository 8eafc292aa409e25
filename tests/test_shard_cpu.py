"""CPU tests of the multi-GPU plumbing (nanospring_b200/shard.py) with the gloo backend,
world_size 2: shard boundaries on the prefix sum of bases, ragged all-gather of sketch rows in
rank order (global read id = shard base + local id), and the identity
    concat_r query(shard_r sketches vs. tables built from ALL sketches) == single-process query
checked with the CPU oracle standing in for the device (it is the checker here, the product
path never runs on the CPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nanospring_b200 import shard


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def make_reads(seed=3, n_reads=90):
    rng = np.random.default_rng(seed)
    genome = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=6000)
    reads = []
    for i in range(n_reads):
        s = int(rng.integers(0, 5000))
        L = int(rng.integers(0, 900)) if i % 7 else int(rng.integers(0, 30))
        reads.append(genome[s:s + L].tobytes())
    return reads


def test_shard_bounds_balance_bases_not_reads():
    lens = np.array([10] * 50 + [1000] * 5 + [10] * 45, dtype=np.uint64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    b = shard.shard_bounds_by_bases(off, 4)
    assert b[0] == 0 and b[-1] == 100 and (np.diff(b) >= 0).all()
    per = [int(off[b[r + 1]] - off[b[r]]) for r in range(4)]
    assert max(per) - min(per) <= 1000          # within one read of each other
    assert len(set(np.diff(b).tolist())) > 1      # read counts differ: balance is on bases
    # degenerate: more ranks than reads, empty input
    assert list(shard.shard_bounds_by_bases(np.array([0, 5], dtype=np.uint64), 3))[-1] == 1
    assert list(shard.shard_bounds_by_bases(np.array([0], dtype=np.uint64), 2)) == [0, 0, 0]
    lo = shard.local_offsets(off, int(b[1]), int(b[2]))
    assert lo[0] == 0 and lo[-1] == off[b[2]] - off[b[1]]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle, reads_to_buffers
    orc = Oracle.get()
    orc.set_num_threads(1)
    k, n, thr = 15, 12, 2
    rnd = orc.rand_from_seed(99, n)
    bases, offsets = reads_to_buffers(make_reads())
    bounds = shard.shard_bounds_by_bases(offsets, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    loff = shard.local_offsets(offsets, lo, hi)
    lbases = bases[int(offsets[lo]):int(offsets[hi])]
    sk_local = orc.sketch_all(lbases, loff, k, n, rnd)                      # "sketch my shard"
    rows = [int(bounds[r + 1] - bounds[r]) for r in range(world)]
    full = shard.all_gather_rows(torch.from_numpy(sk_local.view(np.int64)), rows)   # the one exchange
    full = full.numpy().view(np.uint64)
    T = orc.build_tables(full)                                              # every rank: full tables
    res = [T.query_sketch(sk_local[i], thr) for i in range(hi - lo)]        # query my shard only
    off_l = np.concatenate([[0], np.cumsum([r.size for r in res])]).astype(np.uint64)
    ids_l = np.concatenate(res) if res else np.zeros(0, np.uint32)
    ret[rank] = (full.copy(), off_l, ids_l, lo, hi)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_pipeline_equals_single_process():
    world = 2
    port = free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    from oracle.oracle import Oracle, reads_to_buffers
    orc = Oracle.get()
    k, n, thr = 15, 12, 2
    rnd = orc.rand_from_seed(99, n)
    bases, offsets = reads_to_buffers(make_reads())
    sk = orc.sketch_all(bases, offsets, k, n, rnd)
    woff, wids = orc.build_tables(sk).query_all(bases, offsets, sk, k, rnd, thr, 0)
    # gathered matrix is the single-process matrix on every rank
    for r in range(world):
        assert (ret[r][0] == sk).all()
    # concatenating the shards' CSR in rank order gives the single-process CSR (ids are global)
    ids = np.concatenate([ret[r][2] for r in range(world)])
    counts = np.concatenate([np.diff(ret[r][1].astype(np.int64)) for r in range(world)])
    assert (counts == np.diff(woff.astype(np.int64))).all()
    assert (ids == wids).all()
    assert ret[0][4] == ret[1][3]     # contiguous shards

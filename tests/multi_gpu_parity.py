"""Multi-GPU parity check, launched with torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_parity.py

Every rank sketches its shard (boundaries on the prefix sum of bases), the sketch rows are
all-gathered over NCCL, every rank builds the full tables and queries its own shard.  Rank 0
gathers the CSR pieces, concatenates them in rank order and compares them - and the gathered
sketch matrix - bit for bit with a single-GPU run and with the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nanospring_b200 as ns  # noqa: E402
from nanospring_b200 import shard  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    k, n, thr = 23, 60, 6
    rnd = ns.rand_from_seed(20261017, n)
    lengths = ns.synth_lengths(6000, 4000, seed=5)
    lengths[:5] = [0, 10, 22, 23, 120000]          # short reads + one long read: ragged shards
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=1_000_000, genome_seed=7, read_seed=8,
                                                      p_ins=0.01, p_del=0.01, p_sub=0.01))
    bounds = shard.shard_bounds_by_bases(rd.offsets, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    rows = [int(bounds[r + 1] - bounds[r]) for r in range(world)]
    local = ns.ReadData(rd.bases[int(rd.offsets[lo]):int(rd.offsets[hi])], shard.local_offsets(rd.offsets, lo, hi))

    f = ns.MinHashReadFilter(device=local_rank)
    f.k, f.n, f.overlapSketchThreshold, f.randNumbers = k, n, thr, rnd
    f._create()
    f.load(local)
    f.sketch()
    full = shard.gather_and_build(f, hi - lo, rows, rank)
    off, ids = f.queryAll(False)
    pieces = [None] * world
    dist.all_gather_object(pieces, (np.diff(off.astype(np.int64)), ids))
    # second strategy: tables partitioned by hash function (two all-to-alls, no replicated build)
    pf = shard.PartitionedFilter(f, rank, world)
    ptotal = pf.run(hi - lo, rows)
    poff, pids = pf.result(hi - lo, ptotal)
    same = bool((poff == off).all() and pids.size == ids.size and (pids == ids).all())
    # third strategy (the default): the same partitioning, exchanges over NVLink peer memory
    # inside the kernels, device-side flag barriers (csrc/multigpu.cu); run twice (epochs)
    peer = shard.PeerPartitionedFilter(f, rank, world, rows)
    same_peer = True
    for _ in range(2):
        f.sketch()
        qtotal = peer.run()
        qoff, qids = peer.result(hi - lo, qtotal)
        same_peer &= bool((qoff == off).all() and qids.size == ids.size and (qids == ids).all())
    dist.barrier()
    peer.shutdown()
    flags = [None] * world
    dist.all_gather_object(flags, (same, same_peer))
    ok = all(a and b for a, b in flags)
    if rank == 0:
        print(f"partitioned tables (NCCL) == replicated tables on every rank: {all(a for a, _ in flags)}", flush=True)
        print(f"partitioned tables (peer memory) == replicated tables on every rank: {all(b for _, b in flags)} "
              f"stages ms {dict((k2, round(v, 3)) for k2, v in peer.last_ms.items())}", flush=True)
    if rank == 0:
        counts = np.concatenate([p[0] for p in pieces])
        all_ids = np.concatenate([p[1] for p in pieces])
        g = ns.MinHashReadFilter(device=local_rank)
        g.k, g.n, g.overlapSketchThreshold, g.randNumbers = k, n, thr, rnd
        g.initialize(rd)
        soff, sids = g.queryAll(False)
        sk_single = g.sketches()
        sk_gathered = full.cpu().numpy().view(np.uint64)
        ok &= bool((sk_gathered == sk_single).all())
        ok &= bool((counts == np.diff(soff.astype(np.int64))).all())
        ok &= bool(all_ids.size == sids.size and (all_ids == sids).all())
        from oracle.oracle import Oracle
        orc = Oracle.get()
        want = orc.sketch_all(rd.bases, rd.offsets, k, n, rnd)
        woff, wids = orc.build_tables(want).query_all(rd.bases, rd.offsets, want, k, rnd, thr, 0)
        ok &= bool((sk_single == want).all() and (soff == woff).all() and (sids == wids).all())
        print(f"multi-GPU parity world={world}: shards {rows}, {int(soff[-1])} candidate ids: {'OK' if ok else 'MISMATCH'}",
              flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()

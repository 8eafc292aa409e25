"""Peer-memory multi-rank path with every rank on ONE GPU (the round-end box has a single B200):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29541 tools/peer_same_device.py

Rank r is its own process with its own CUDA context on device 0; the arenas are shared through
cudaIpc handles exactly as between two GPUs (csrc/multigpu.cu), the kernels of the processes
time-slice on the device.  torch.distributed (gloo) only carries the setup tokens and the
verdict.  Every rank compares its shard's CSR with a single-process run over all reads."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nanospring_b200 as ns  # noqa: E402
from nanospring_b200 import shard  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    k, n, thr = 23, int(os.environ.get("NSMH_TEST_N", "60")), 4
    rnd = ns.rand_from_seed(77, n)
    lengths = ns.synth_lengths(3000, 2500, seed=11)
    lengths[:5] = [0, 10, 22, 23, 60000]
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=400_000, genome_seed=3, read_seed=4,
                                                      p_ins=0.01, p_del=0.01, p_sub=0.01))
    bounds = shard.shard_bounds_by_bases(rd.offsets, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    rows = [int(bounds[r + 1] - bounds[r]) for r in range(world)]
    local = ns.ReadData(rd.bases[int(rd.offsets[lo]):int(rd.offsets[hi])], shard.local_offsets(rd.offsets, lo, hi))

    g = ns.MinHashReadFilter(device=0)
    g.k, g.n, g.overlapSketchThreshold, g.randNumbers = k, n, thr, rnd
    g.initialize(rd)
    soff, sids = g.queryAll(False)
    g.close()
    want_counts = np.diff(soff.astype(np.int64))[lo:hi]
    want_ids = sids[int(soff[lo]):int(soff[hi])]

    f = ns.MinHashReadFilter(device=0)
    f.k, f.n, f.overlapSketchThreshold, f.randNumbers = k, n, thr, rnd
    f._create()
    f.load(local)
    peer = shard.PeerPartitionedFilter(f, rank, world, rows)
    ok = True
    for _ in range(2):
        f.sketch()
        total = peer.run()
        off, ids = peer.result(hi - lo, total)
        ok &= bool((np.diff(off.astype(np.int64)) == want_counts).all() and ids.size == want_ids.size
                   and (ids == want_ids).all())
    dist.barrier()
    peer.shutdown()
    f.close()
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        print(f"peer-memory path, {world} ranks on one device, shards {rows}, {int(soff[-1])} candidate ids: "
              f"{'OK' if all(flags) else 'MISMATCH'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()

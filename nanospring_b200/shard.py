"""Multi-GPU plumbing for the read-overlap stage: one process per GPU, reads sharded by
bases, sketches all-gathered (NCCL over NVLink; gloo on CPU for the tests), every rank
builds the full tables and queries its own shard.

The reference has no distributed path (single process, OpenMP: ReadFilter.cpp:31-44 is a
loop over reads, ReadFilter.cpp:163-165 a loop over tables), so this module defines the
only exchange step the path has: one all-gather of [reads][n] u64 sketch rows.
"""
import numpy as np


def shard_bounds_by_bases(offsets, world_size):
    """Contiguous read ranges with (nearly) equal numbers of BASES, not reads: boundaries on
    the prefix sum of read lengths.  Returns int64[world_size+1] read indices."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n_reads = offsets.size - 1
    total = int(offsets[-1])
    bounds = np.zeros(world_size + 1, dtype=np.int64)
    for r in range(1, world_size):
        target = total * r // world_size
        bounds[r] = int(np.searchsorted(offsets, np.uint64(target), side="left"))
    bounds[world_size] = n_reads
    bounds = np.maximum.accumulate(np.clip(bounds, 0, n_reads))
    return bounds


def local_offsets(offsets, lo, hi):
    """Offsets of reads [lo, hi) rebased to start at 0."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    return (offsets[lo:hi + 1] - offsets[lo]).astype(np.uint64)


def all_gather_rows(local_rows, rows_per_rank, group=None):
    """All-gather a ragged row-sharded matrix.  local_rows: torch tensor [rows_r, n] (int64 view
    of the u64 sketches); rows_per_rank: list of row counts.  Returns [sum(rows), n] in rank
    order, identical on every rank, so global read id = shard base + local id."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    n = local_rows.shape[1]
    max_rows = int(max(rows_per_rank))
    if all(int(r) == max_rows for r in rows_per_rank):
        out = torch.empty((world * max_rows, n), dtype=local_rows.dtype, device=local_rows.device)
        dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group)
        return out
    padded = torch.zeros((max_rows, n), dtype=local_rows.dtype, device=local_rows.device)
    padded[:local_rows.shape[0]] = local_rows
    buf = torch.empty((world * max_rows, n), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    parts = [buf[r * max_rows:r * max_rows + int(rows_per_rank[r])] for r in range(world)]
    return torch.cat(parts, dim=0).contiguous()


class DeviceAlias:
    """Zero-copy torch view of a device buffer owned by libnsmh.so."""

    def __init__(self, ptr, count, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3}


def sketches_as_tensor(filt, n_reads):
    """The filter's device-resident sketch matrix as an int64 torch tensor [n_reads, n]."""
    import torch
    if n_reads == 0:
        return torch.empty((0, filt.n), dtype=torch.int64, device=f"cuda:{filt.device}")
    alias = DeviceAlias(filt.sketchesDevicePtr(), n_reads * filt.n)
    return torch.as_tensor(alias, device=f"cuda:{filt.device}").view(n_reads, filt.n)


def gather_and_build(filt, n_local, rows_per_rank, rank, group=None):
    """sketch (already done) -> all-gather -> build full tables on this rank.
    Returns the gathered tensor (must stay alive until build has finished; it has)."""
    import torch
    filt.synchronize()
    local = sketches_as_tensor(filt, n_local)
    full = all_gather_rows(local, rows_per_rank, group)
    torch.cuda.synchronize()
    base = int(sum(int(r) for r in rows_per_rank[:rank]))
    filt.setTableSketches(full.data_ptr(), full.shape[0], base)
    filt.build()
    return full

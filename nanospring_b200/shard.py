"""Multi-GPU plumbing for the read-overlap stage: one process per GPU, reads sharded by
bases.  Three strategies, all bit-identical to one GPU:

  PeerPartitionedFilter  (default) tables partitioned by hash function, both exchanges done by
                         the producing kernels as stores into NVLink peer memory, flag barriers
                         on the device (csrc/multigpu.cu, nsmh_mg_*); torch.distributed only
                         carries the 256-byte setup tokens.
  PartitionedFilter      the same partitioning with NCCL all-to-alls driven from the host.
  gather_and_build       SURVEY's baseline: NCCL all-gather of the sketch rows, every rank
                         builds the full tables (cost grows with the global read count).

The reference has no distributed path (single process, OpenMP: ReadFilter.cpp:31-44 is a
loop over reads, ReadFilter.cpp:163-165 a loop over tables).
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


# ------------------------------------------------------------------------------------------
# Tables partitioned by hash function: constant work per GPU under weak scaling.
#
# The replicated build above grows with the GLOBAL read count on every rank.  Here rank g owns
# the tables of the hash functions cols[g] (a contiguous block of the n columns) for ALL reads:
#   1. all-to-all of sketch COLUMNS: rank r sends S_r[:, cols[g]] to rank g, which receives the
#      [total_reads][n_g] matrix (rows in global read order);
#   2. rank g builds its n_g tables from it and probes them for every read (nsmh_probe_lists):
#      one id list per (read, owned hash), concatenated per read;
#   3. all-to-all of those per-read lists to the rank that owns the read;
#   4. the owner thresholds the union of the `world` partial lists (nsmh_count_lists).
# Every step handles reads_per_rank * n items per rank, whatever the world size; the two
# exchanges move ~8 B and ~4 B per (read, hash).  Results are bit-identical to the single-GPU
# run because the multiset of (read, hash) -> id-list is the same, only computed elsewhere.
class PartitionedFilter:
    def __init__(self, local_filter, rank, world, group=None):
        import ctypes as C
        from .filter import MinHashReadFilter
        self.C = C
        self.f = local_filter                       # n hashes: pack, sketch, final count
        self.rank, self.world, self.group = rank, world, group
        n = local_filter.n
        self.cols = [c for c in np.array_split(np.arange(n), world)]
        if any(c.size == 0 for c in self.cols):
            raise ValueError("more ranks than hash functions")
        mine = self.cols[rank]
        self.sub = MinHashReadFilter(device=local_filter.device)
        self.sub.k, self.sub.n = local_filter.k, int(mine.size)
        self.sub.overlapSketchThreshold = local_filter.overlapSketchThreshold
        self.sub.randNumbers = np.ascontiguousarray(np.asarray(local_filter.randNumbers, dtype=np.uint64)[mine])
        self.sub._create()
        self.keep = []

    def run(self, n_local, rows_per_rank):
        """After self.f.sketch(): exchange, build, probe, exchange, count.  Returns total ids;
        the CSR of the local reads is then in self.f (queryAll-style accessors)."""
        import torch
        import torch.distributed as dist
        from ._lib import check, lib
        C = self.C
        f, sub, world, rank = self.f, self.sub, self.world, self.rank
        dev = f"cuda:{f.device}"
        import time
        rows = [int(r) for r in rows_per_rank]
        total_rows = sum(rows)
        bounds = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
        n_mine = sub.n
        f.synchronize()
        t0 = time.perf_counter()
        S = sketches_as_tensor(f, n_local)
        # 1. sketch columns to their table owners
        send = torch.cat([S[:, int(c[0]):int(c[-1]) + 1].reshape(-1) for c in self.cols])
        in_splits = [n_local * int(c.size) for c in self.cols]
        out_splits = [r * n_mine for r in rows]
        M = torch.empty(total_rows * n_mine, dtype=torch.int64, device=dev)
        dist.all_to_all_single(M, send, out_splits, in_splits, group=self.group)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        # 2. build the owned tables over all reads, probe them for all reads
        sub.setTableSketches(M.data_ptr(), total_rows, 0)
        sub.build()
        t2 = time.perf_counter()
        tot = C.c_uint64(0)
        check(lib().nsmh_probe_lists(sub._h, M.data_ptr(), total_rows, C.byref(tot)))
        t3 = time.perf_counter()
        p_off, p_ids = C.c_void_p(), C.c_void_p()
        check(lib().nsmh_query_all_device_ptrs(sub._h, C.byref(p_off), C.byref(p_ids)))
        offs = torch.as_tensor(DeviceAlias(p_off.value, total_rows + 1), device=dev)
        ids = (torch.as_tensor(DeviceAlias(p_ids.value, max(tot.value, 1), "<i4"), device=dev)[:tot.value])
        # 3. per-read lists to the read owners.  The per-read counts travel first, with the
        #    segment's total id count piggybacked as one extra element, so one host read gives
        #    both the send and the receive sizes of the id exchange.
        b = self._bounds_dev(bounds, dev)
        sizes = offs[b[1:]] - offs[b[:-1]]
        counts = (offs[1:] - offs[:-1])
        pieces = []
        for r in range(world):
            pieces.append(counts[int(bounds[r]):int(bounds[r + 1])])
            pieces.append(sizes[r:r + 1])
        # int64 on the wire: a segment's total can pass 2^31 ids long before a single read's count does
        # (k = 15, n = 120 gathers ~800 ids per read), and a wrapped total would give the ranks
        # disagreeing split sizes for the id exchange below
        send_cnt = torch.cat(pieces).to(torch.int64)
        cnt_recv = torch.empty(world * (n_local + 1), dtype=torch.int64, device=dev)
        dist.all_to_all_single(cnt_recv, send_cnt, [n_local + 1] * world, [r + 1 for r in rows], group=self.group)
        cnt_recv = cnt_recv.view(world, n_local + 1)
        both = torch.cat([sizes.to(torch.int64), cnt_recv[:, n_local]]).cpu().tolist()
        send_sizes, recv_sizes = both[:world], both[world:]
        part_offs = torch.zeros((world, n_local + 1), dtype=torch.int64, device=dev)
        torch.cumsum(cnt_recv[:, :n_local], dim=1, out=part_offs[:, 1:])
        if max([0] + [int(x) for x in send_sizes + recv_sizes]) >= 2**31:
            raise ValueError("id exchange: more than 2^31-1 ids between one pair of ranks; use more ranks or the peer-memory path")
        ids_recv = torch.empty(max(int(sum(recv_sizes)), 1), dtype=torch.int32, device=dev)
        dist.all_to_all_single(ids_recv[:int(sum(recv_sizes))], ids, [int(x) for x in recv_sizes],
                               [int(x) for x in send_sizes], group=self.group)
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        # 4. threshold the union of the partial lists of every local read
        starts = np.concatenate([[0], np.cumsum(recv_sizes)]).astype(np.int64)
        off_ptrs = (C.c_void_p * world)(*[part_offs[p].data_ptr() for p in range(world)])
        id_ptrs = (C.c_void_p * world)(*[ids_recv.data_ptr() + 4 * int(starts[p]) for p in range(world)])
        out_total = C.c_uint64(0)
        check(lib().nsmh_count_lists(f._h, n_local, world, off_ptrs, id_ptrs, C.byref(out_total)))
        self.keep = [M, part_offs, ids_recv]
        t5 = time.perf_counter()
        # host wall time of the five stages of the last run, ms (every stage ends synchronised)
        self.last_ms = {"exchange_columns": 1e3 * (t1 - t0), "build_owned_tables": 1e3 * (t2 - t1),
                        "probe_lists": 1e3 * (t3 - t2), "exchange_lists": 1e3 * (t4 - t3),
                        "count_lists": 1e3 * (t5 - t4)}
        return out_total.value

    def _bounds_dev(self, bounds, dev):
        import torch
        key = tuple(int(x) for x in bounds)
        if getattr(self, "_bkey", None) != key:
            self._bkey, self._bdev = key, torch.from_numpy(np.asarray(bounds, dtype=np.int64)).to(dev)
        return self._bdev

    def result(self, n_local, total):
        """CSR (offsets u64[n_local+1], ids u32[total]) of the local reads, global read ids."""
        from ._lib import check, lib, u32p, u64p
        off = np.zeros(n_local + 1, dtype=np.uint64)
        ids = np.zeros(max(total, 1), dtype=np.uint32)
        check(lib().nsmh_query_all_result(self.f._h, off.ctypes.data_as(u64p), ids.ctypes.data_as(u32p)))
        return off, ids[:total]


# ------------------------------------------------------------------------------------------
class PeerPartitionedFilter:
    """Tables partitioned by hash function; exchanges over NVLink peer memory inside the kernels
    (include/nsmh.h, nsmh_mg_*).  `group` is only used once, to exchange the setup tokens."""

    def __init__(self, local_filter, rank, world, rows_per_rank, group=None):
        import ctypes as C
        from ._lib import MG_TOKEN_BYTES, check, lib, u32p
        self.C = C
        self.f = local_filter
        self.rank, self.world = rank, world
        rows = np.ascontiguousarray(np.asarray(rows_per_rank, dtype=np.uint32))
        if rows.size != world:
            raise ValueError("rows_per_rank must have one entry per rank")
        token = (C.c_uint8 * MG_TOKEN_BYTES)()
        # A rank whose arena could not be set up still takes part in the token exchange (with an
        # empty token), so that the ranks never sit in different collectives; all of them then raise.
        init_error = None
        try:
            check(lib().nsmh_mg_init(self.f._h, rank, world, rows.ctypes.data_as(u32p), token))
        except Exception as e:  # noqa: BLE001
            init_error = e
        mine = b"" if init_error is not None else bytes(token)
        tokens = [None] * world
        if world > 1:
            import torch.distributed as dist
            dist.all_gather_object(tokens, mine, group=group)
        else:
            tokens[0] = mine
        if init_error is not None:
            raise init_error
        bad = [r for r, t in enumerate(tokens) if len(t) != MG_TOKEN_BYTES]
        if bad:
            lib().nsmh_mg_shutdown(self.f._h)
            raise RuntimeError(f"nsmh_mg_init failed on rank(s) {bad}")
        blob = b"".join(tokens)
        check(lib().nsmh_mg_connect(self.f._h, blob))
        self.last_ms = {}

    def run(self, n_local=None, rows_per_rank=None, sketch=False):
        """After self.f.sketch() - or, sketch=True, including it (nsmh_mg_sketch_run: the sketch's fix-up pass runs
        beside the column scatter): scatter columns, build owned tables, probe, count.  Returns the number of
        candidate ids of the local reads; the CSR stays on the device."""
        from ._lib import check, lib
        C = self.C
        total = C.c_uint64(0)
        check((lib().nsmh_mg_sketch_run if sketch else lib().nsmh_mg_run)(self.f._h, C.byref(total)))
        ms = (C.c_float * 6)()
        check(lib().nsmh_mg_stage_ms(self.f._h, ms))
        self.last_ms = dict(zip(("scatter_columns", "barrier_1", "build_owned_tables", "probe_to_peers",
                                 "barrier_2", "count"), [float(x) for x in ms]))
        return total.value

    def result(self, n_local, total):
        """CSR (offsets u64[n_local+1], ids u32[total]) of the local reads, global read ids."""
        from ._lib import check, lib, u32p, u64p
        off = np.zeros(n_local + 1, dtype=np.uint64)
        ids = np.zeros(max(total, 1), dtype=np.uint32)
        check(lib().nsmh_query_all_result(self.f._h, off.ctypes.data_as(u64p), ids.ctypes.data_as(u32p)))
        return off, ids[:total]

    def shutdown(self):
        from ._lib import check, lib
        check(lib().nsmh_mg_shutdown(self.f._h))

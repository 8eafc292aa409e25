#!/usr/bin/env python
"""Benchmark of the MinHash read-overlap stage (BASELINE.json metric: Gbases/s sketch+lookup).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic reads:
    2-bit pack -> sketch (n minima per read) -> build n hash tables -> bulk lookup (forward
    self-query of every read, CSR of candidate ids left on the device).
Workload at N=1 = BASELINE.json configs[1]: 100 000 synthetic nanopore-like reads, ~10 kb
mean (mixed-gamma lengths), 10 % error (3 % ins, 3 % del, 4 % sub), 50 % reverse strand,
random 50 Mb genome, k=23, n=60, overlap-sketch-thr=6.  At N>1 every rank owns one such
shard (weak scaling); the tables are partitioned by hash function across the ranks and the
two exchanges (sketch columns to the table owners, probe results back to the read owners) are
stores into NVLink peer memory issued by the producing kernels (csrc/multigpu.cu).

`value`  : device-resident Gbases/s (ASCII reads already in HBM when the clock starts).
`e2e`    : the same metric through the public API with HOST buffers: pinned ASCII reads are
           copied host->device and the candidate CSR is copied back inside the timed region.
`--impl reference` times the reference's own CPU implementation (oracle/_ref, the unmodified
ReadFilter.cpp/BBHashMap.cpp; falls back to the C port in oracle/) on all host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, NHASH, THR = 23, 60, 6
READS_PER_GPU = 100_000
MEAN_LEN = 10_000
RAND_SEED = 20261017
GENOME_LEN = 50_000_000
# Counters of ONE launch of the dominant kernel on exactly this workload (rank-0 shard, k=23 n=60) from an
# `ncu --set full` capture: recorded figures, not measured by this run.  They live in a committed file that
# names the commit and the capture they come from (profiles/ncu_counters.json); the kernel is deterministic
# on this workload, so with the live kernel time they give utilisations.
def recorded_counters():
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_counters.json")) as fh:
            return json.load(fh)
    except (OSError, ValueError):
        return {}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_name(n_gpus):
    return (f"synthetic {READS_PER_GPU * n_gpus} nanopore reads (~10 kb mean mixed-gamma, 10% error, "
            f"50 Mb random genome), k={K} n={NHASH} thr={THR}" + (f", {n_gpus} shards of {READS_PER_GPU}" if n_gpus > 1 else ""))


def shard_lengths(rank):
    import nanospring_b200 as ns
    return ns.synth_lengths(READS_PER_GPU, MEAN_LEN, seed=1000 + rank)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            inside = t0 - 0.05 <= ts <= t1 + 0.15
            try:
                if inside:
                    sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            if inside:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
def reference_arm(args, rank):
    """The reference's own CPU implementation on a bounded sample of the same workload."""
    if rank != 0:
        return
    import nanospring_b200 as ns
    from oracle.oracle import Oracle, RefLib
    cores = os.cpu_count() or 1
    sample_reads = 10_000
    lengths = shard_lengths(0)[:sample_reads]
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=GENOME_LEN))
    rnd = ns.rand_from_seed(RAND_SEED, NHASH)
    bases_total = int(rd.offsets[-1])
    use_ref = RefLib.available()

    def one_step():
        t0 = time.perf_counter()
        if use_ref:
            rf = RefLib.get().create(rd.bases, rd.offsets, K, NHASH, THR, rnd, threads=cores)
            rf.query_all(0, threads=cores)
            rf.close()
        else:
            orc = Oracle.get()
            orc.set_num_threads(cores)
            sk = orc.sketch_all(rd.bases, rd.offsets, K, NHASH, rnd)
            T = orc.build_tables(sk)
            T.query_all(rd.bases, rd.offsets, sk, K, rnd, THR, 0)
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    ts = [one_step() for _ in range(args.steps)]
    total = sum(ts)
    value = bases_total * args.steps / total / 1e9
    sample = (f"first {sample_reads} reads of shard 0 ({bases_total / 1e9:.3f} Gbases): sketch + "
              f"populateHashTables + getFilteredReads(sketch) for every read, {cores} OpenMP threads")
    line = {
        "impl": "reference", "metric": "Gbases/s MinHash sketch+lookup", "value": value, "unit": "Gbases/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Gbases/s", "cores": cores,
                         "kind": "reference" if use_ref else "port", "sample": sample},
        "e2e": {"value": value, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def fastq_text_device(d_bases, offsets, max_reads):
    """FASTQ text ("@\n" seq "\n+\n" qual "\n" per record) of the first max_reads reads, built on the device.
    Returns (text u8 tensor with 64 spare bytes, text bytes, bases, reads, record starts int64[reads+1])."""
    import torch
    n = min(max_reads, offsets.size - 1)
    L = torch.from_numpy(np.diff(offsets[:n + 1].astype(np.int64))).cuda()
    nb = int(offsets[n])
    rec = 2 * L + 6
    start = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    torch.cumsum(rec, 0, out=start[1:])
    nbytes = int(start[-1].item())
    text = torch.full((nbytes + 64,), ord("I"), dtype=torch.uint8, device="cuda")
    s0 = start[:-1]
    seq0 = s0 + 2
    text[s0] = ord("@")
    text[s0 + 1] = 10
    off_d = torch.from_numpy(offsets[:n].astype(np.int64)).cuda()
    idx = torch.repeat_interleave(seq0 - off_d, L) + torch.arange(nb, dtype=torch.int64, device="cuda")
    text[idx] = d_bases[:nb]
    del idx
    text[seq0 + L] = 10
    text[seq0 + L + 1] = ord("+")
    text[seq0 + L + 2] = 10
    text[start[1:] - 1] = 10
    torch.cuda.synchronize()
    return text, nbytes, nb, n, start


def ingest_leg(local_rank, d_bases, offsets, steps, hbm_peak):
    """SURVEY 8(f) N2, reported beside the headline: FASTQ text -> packed reads on the device
    (csrc/fastq.cu) for the first INGEST_READS reads of the workload, (a) text already in HBM,
    (b) text in pinned host memory (H2D inside the timed region), and the reference's own loader
    (oracle/_ref/libnsref_readdata.so, ReadData::loadFromFile low_mem = true) on a bounded sample."""
    import tempfile
    import torch
    import nanospring_b200 as ns
    text, nbytes, nb, n, start = fastq_text_device(d_bases, offsets, 25_000)
    rd = ns.GpuReadData(device=local_rank)
    for _ in range(2):
        rd.loadFromDeviceText(text.data_ptr(), nbytes)
    assert rd.getNumReads() == n and int(rd.offsets[-1]) == nb, "ingest: wrong read table"
    ms, pk = [], []
    for _ in range(steps):
        rd.loadFromDeviceText(text.data_ptr(), nbytes)
        st = rd.stats()
        ms.append(st["fastq_parse_ms"])
        pk.append(st["fastq_pack_ms"])
    dev_ms = float(np.mean(ms))
    alg_bytes = nbytes + nb + nb / 4 + 8 * n             # text once, read lines again, packed words + offsets out
    h_text = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_text.copy_(text[:nbytes])
    torch.cuda.synchronize()
    rd.loadFromText(h_text.numpy())
    t0 = time.perf_counter()
    for _ in range(steps):
        rd.loadFromText(h_text.numpy())
    e2e_ms = 1e3 * (time.perf_counter() - t0) / steps
    out = {"workload": f"FASTQ text of the first {n} reads ({nb / 1e9:.3f} Gbases, {nbytes / 1e9:.3f} GB of text)",
           "device_ms": dev_ms, "device_gbases_per_s": nb / (dev_ms * 1e-3) / 1e9, "pack_kernel_ms": float(np.mean(pk)),
           "roofline": {"bound": "hbm", "achieved": alg_bytes / (dev_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / (dev_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": alg_bytes},
           "e2e_host_text_ms": e2e_ms, "e2e_gbases_per_s": nb / (e2e_ms * 1e-3) / 1e9, "h2d_bytes": nbytes}
    rd.close()
    try:
        from oracle.oracle import RefReadData
        if RefReadData.available():
            m = min(n, 2_000)
            mb = int(start[m].item())
            with tempfile.TemporaryDirectory() as td:
                pth = os.path.join(td, "sample.fastq")
                with open(pth, "wb") as f:
                    f.write(h_text.numpy()[:mb].tobytes())
                t0 = time.perf_counter()
                _, roff, _, _ = RefReadData.get().load(pth, False, True)
                dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": int(roff[-1]) / dt / 1e9, "unit": "Gbases/s", "cores": 1, "kind": "reference",
                                   "sample": f"first {m} records ({mb / 1e6:.0f} MB of text): ReadData::loadFromFile "
                                             f"(FASTQ, low_mem) + getRead of every read", "seconds": dt}
    except Exception as e:  # noqa: BLE001
        out["cpu_baseline"] = {"error": str(e)[:200]}
    del text
    return out


def dnabitset_host(d_bases, offsets, chunk_reads=8192):
    """The reads in the reference's DnaBitset form (dnaToBits.cpp:11-36: code (c&2)|((c&4)>>2), 4 bases per byte,
    first base in bits 7..6, every read starts on a byte), built on the device and copied to pinned host memory.
    Returns (uint8 numpy view of the pinned buffer, uint32 lengths)."""
    import torch
    n = offsets.size - 1
    lens = np.diff(offsets.astype(np.int64))
    nbytes = (lens + 3) // 4
    boff = np.zeros(n + 1, dtype=np.int64)
    boff[1:] = np.cumsum(nbytes)
    out = torch.zeros(int(boff[-1]) + 4, dtype=torch.uint8, device="cuda")
    for r0 in range(0, n, chunk_reads):
        r1 = min(n, r0 + chunk_reads)
        b0, b1 = int(offsets[r0]), int(offsets[r1])
        if b1 == b0:
            continue
        c = d_bases[b0:b1].to(torch.int32)
        code = (c & 2) | ((c & 4) >> 2)
        L = torch.from_numpy(lens[r0:r1]).cuda()
        # position of every base in the padded layout (4 * byte offset of its read + index inside the read)
        shift = torch.from_numpy(4 * (boff[r0:r1] - boff[r0]) - (offsets[r0:r1].astype(np.int64) - b0)).cuda()
        pos = torch.arange(b1 - b0, dtype=torch.int64, device="cuda") + torch.repeat_interleave(shift, L)
        padded = torch.zeros(4 * int(boff[r1] - boff[r0]), dtype=torch.int32, device="cuda")
        padded[pos] = code
        q = padded.view(-1, 4)
        out[int(boff[r0]):int(boff[r1])] = ((q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]).to(torch.uint8)
        del c, code, pos, padded, q
    h = torch.empty(int(boff[-1]), dtype=torch.uint8, pin_memory=True)
    h.copy_(out[:int(boff[-1])])
    torch.cuda.synchronize()
    return h.numpy(), lens.astype(np.uint32)


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="nsmh", choices=["nsmh", "reference"])
    ap.add_argument("--sketch-mode", type=int, default=0, help="0 filtered kernel, 1 brute force")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the host-buffer leg")
    ap.add_argument("--no-ingest", action="store_true", help="skip the FASTQ-ingest side measurement (N=1 only)")
    ap.add_argument("--separate-build", action="store_true", help="nsmh_sketch then nsmh_build instead of nsmh_sketch_build (A/B)")
    ap.add_argument("--no-legs", action="store_true", help="skip the side legs (other BASELINE configs, brute-force roofline)")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the parity checks of the multi-GPU result")
    ap.add_argument("--multi", default="auto", choices=["auto", "peer", "replicated", "partitioned"],
                    help="N>1: tables partitioned by hash function with exchanges over NVLink peer memory inside the "
                         "kernels (peer, default), the same with NCCL all-to-alls (partitioned), or NCCL all-gather "
                         "of the sketches + full tables on every rank (replicated)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "nsmh" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import nanospring_b200 as ns
    from nanospring_b200 import shard
    from nanospring_b200._lib import check, lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # host side of the e2e legs: this rank's threads and its pinned buffers live on the NUMA node its GPU hangs
    # off (first touch), so that eight ranks loading at once do not meet on the socket interconnect
    all_cpus = os.sched_getaffinity(0)
    numa = C.c_int(-1)
    check(lib().nsmh_bind_thread_near(local_rank, C.byref(numa)))
    host_numa = {"gpu_numa_node": int(numa.value), "cpus_bound": len(os.sched_getaffinity(0)), "cpus_all": len(all_cpus)}
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")     # keep NCCL's version banner off stdout
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    n_gpus = world

    # ---- synthetic shard, generated on the device; a pinned host copy feeds the e2e leg ----
    lengths = shard_lengths(rank)
    offsets = np.zeros(lengths.size + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lengths, dtype=np.uint64)
    total_bases = int(offsets[-1])
    params = ns.synth_params(genome_len=GENOME_LEN)
    d_off = torch.from_numpy(offsets.astype(np.int64)).cuda()
    d_bases = torch.empty(total_bases + 64, dtype=torch.uint8, device="cuda")
    check(lib().nsmh_synth_reads_device(local_rank, C.byref(params), rank * READS_PER_GPU, lengths.size,
                                        d_off.data_ptr(), d_bases.data_ptr()))
    h_bases = torch.empty(total_bases, dtype=torch.uint8, pin_memory=True)
    h_bases.copy_(d_bases[:total_bases])
    torch.cuda.synchronize()
    host_rd = ns.ReadData(h_bases.numpy(), offsets)
    log(f"[rank {rank}] shard: {lengths.size} reads, {total_bases / 1e9:.3f} Gbases, max read {int(lengths.max())}")

    f = ns.MinHashReadFilter(device=local_rank)
    f.k, f.n, f.overlapSketchThreshold = K, NHASH, THR
    f.randNumbers = ns.rand_from_seed(RAND_SEED, NHASH)
    f.sketchMode = args.sketch_mode
    f._create()
    rows_per_rank = [READS_PER_GPU] * world
    multi = args.multi if args.multi != "auto" else "peer"
    pf = None
    multi_note = None
    if world > 1 and multi == "peer":
        # Peer-memory path (cudaIpc arenas + device-side flag barriers).  In auto mode one trial step
        # runs first and all ranks agree on the outcome; if any rank cannot set it up (no P2P access,
        # IPC refused, a barrier timing out), every rank switches to the NCCL all-gather strategy.
        ok, why = 1, ""
        try:
            pf = shard.PeerPartitionedFilter(f, rank, world, rows_per_rank)
            f.load_device(d_bases.data_ptr(), d_off.data_ptr(), lengths.size, total_bases)
            f.sketch()
            pf.run(lengths.size, rows_per_rank)
        except Exception as e:  # noqa: BLE001
            if args.multi != "auto":
                raise
            ok, why = 0, f"{type(e).__name__}: {e}"[:200]
        if args.multi == "auto":
            t = torch.tensor([ok], device="cuda", dtype=torch.int32)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if int(t.item()) == 0:
                log(f"[rank {rank}] peer-memory path unavailable ({why or 'another rank failed'}); using NCCL all-gather")
                multi, multi_note = "replicated", "auto: peer-memory path failed its trial step, fell back to NCCL all-gather"
                try:
                    if pf is not None:
                        pf.shutdown()
                except Exception:  # noqa: BLE001
                    pass
                pf = None
                f.close()
                f = ns.MinHashReadFilter(device=local_rank)
                f.k, f.n, f.overlapSketchThreshold = K, NHASH, THR
                f.randNumbers = ns.rand_from_seed(RAND_SEED, NHASH)
                f.sketchMode = args.sketch_mode
                f._create()
    elif world > 1 and multi == "partitioned":
        pf = shard.PartitionedFilter(f, rank, world)
    ext = torch.cuda.ExternalStream(f.stream(), device=local_rank)
    keep = {}

    def device_step():
        f.load_device(d_bases.data_ptr(), d_off.data_ptr(), lengths.size, total_bases)
        if world == 1 and not args.separate_build:
            f.sketch_build()        # nsmh_sketch_build: the sketch's fix-up pass beside the table insert
            return f.queryAll(False, fetch=False)
        if pf is not None:
            if args.separate_build:
                f.sketch()
            return pf.run(lengths.size, rows_per_rank, sketch=not args.separate_build)
        f.sketch()
        if world > 1:
            keep["g"] = shard.gather_and_build(f, lengths.size, rows_per_rank, rank)
        else:
            f.build()
        return f.queryAll(False, fetch=False)

    def e2e_step():
        if world == 1:
            # the reference's own call: initialize(ReadData&) = load + sketch + build (pipelined inside the library)
            f.initialize(host_rd)
            return f.queryAll(False, fetch=True)
        f.load_sketch(host_rd)
        if pf is not None:
            return pf.result(lengths.size, pf.run(lengths.size, rows_per_rank))
        if world > 1:
            keep["g"] = shard.gather_and_build(f, lengths.size, rows_per_rank, rank)
        else:
            f.build()
        off, ids = f.queryAll(False, fetch=True)
        return off, ids

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(ext)
        out = None
        for _ in range(steps):
            out = fn()
        e1.record(ext)
        barrier()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out, t0, t1

    for _ in range(args.warmup):
        device_step()
    launches0 = f.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    ms_dev, total_ids, t0, t1 = timed(device_step, args.steps)
    clocks = sampler.stop(t0, t1)
    st = f.stats()
    launches = st["kernel_launches"] - launches0
    all_bases = total_bases
    if dist is not None:
        t = torch.tensor([total_bases], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        all_bases = int(t.item())
    value = all_bases * args.steps / (ms_dev * 1e-3) / 1e9

    # per-phase device times of the last step (CUDA events inside the library)
    phases = {"pack_ms": st["pack_ms"], "sketch_ms": st["sketch_ms"], "sketch_main_kernel_ms": st["sketch_main_ms"],
              "build_ms": st["build_ms"], "query_ms": st["query_ms"], "sketch_fixups_total": st["sketch_fixups"],
              "query_pairs": st["query_pairs"], "candidate_ids": int(total_ids)}

    if pf is not None:
        phases["partitioned_stage_ms"] = {k2: round(v, 3) for k2, v in pf.last_ms.items()}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, [round(v, 3) for v in pf.last_ms.values()])
        if rank == 0:
            log("stage ms per rank " + " ".join(pf.last_ms.keys()))
            for r, v in enumerate(per_rank):
                log(f"  rank {r}: {v}")

    # ---- e2e: host buffers in, CSR out ----
    for _ in range(0 if args.no_e2e else 2):
        e2e_step()
    ms_e2e, (off, ids), _, _ = timed(e2e_step, 1 if args.no_e2e else args.steps)
    if args.no_e2e:
        ms_e2e *= args.steps
    e2e_value = all_bases * args.steps / (ms_e2e * 1e-3) / 1e9
    h2d = total_bases + offsets.nbytes
    d2h = off.nbytes + ids.nbytes

    # ---- e2e with the reads 2-bit packed on the host, the reference's own in-memory / temp-file form
    #      (DnaBitset, dnaToBits.cpp:11-36: 4 bases per byte, first base in bits 7..6, every read byte-aligned):
    #      what the real caller's ReadData holds (ReadData.cpp:156-235), 4x fewer bytes over PCIe ----
    e2e_packed = None
    prep_ok, h_packed = 0, None
    if not args.no_e2e:
        try:
            h_packed, len32 = dnabitset_host(d_bases, offsets)
            prep_ok = 1
        except Exception as e:  # noqa: BLE001
            e2e_packed = {"error": f"{type(e).__name__}: {e}"[:300]}
        if dist is not None:            # the leg is collective: every rank runs it or none does
            t = torch.tensor([prep_ok], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            prep_ok = int(t.item())
    if prep_ok:
        try:

            def e2e_packed_step():
                if world == 1:
                    f.initialize_dnabitset(h_packed, len32)
                    return f.queryAll(False, fetch=True)
                f.load_sketch(packed=(h_packed, len32))
                if pf is not None:
                    return pf.result(lengths.size, pf.run(lengths.size, rows_per_rank))
                keep["g"] = shard.gather_and_build(f, lengths.size, rows_per_rank, rank)
                return f.queryAll(False, fetch=True)

            for _ in range(2):
                e2e_packed_step()
            ms_p, (off_p, ids_p), _, _ = timed(e2e_packed_step, args.steps)
            e2e_packed = {"value": all_bases * args.steps / (ms_p * 1e-3) / 1e9, "unit": "Gbases/s",
                          "h2d_bytes_per_step": int(h_packed.nbytes + len32.nbytes), "d2h_bytes_per_step": int(off_p.nbytes + ids_p.nbytes),
                          "ms_per_step": ms_p / args.steps, "same_csr_as_ascii_path": bool((off_p == off).all() and (ids_p == ids).all()),
                          "input": "DnaBitset bytes + u32 lengths in pinned host memory, the reference's own 2-bit store (" + ("nsmh_initialize_dnabitset" if world == 1 else "nsmh_load_sketch_dnabitset + multi-GPU build") + ")"}
            del h_packed
        except Exception as e:  # noqa: BLE001
            e2e_packed = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- N > 1: is the multi-GPU result the single-GPU result?  (bench_legs.parity_of) ----
    parity = None
    if world > 1 and not args.no_parity:
        import types
        import bench_legs
        total_now = device_step()
        view = types.SimpleNamespace(f=f, pf=pf, d_bases=d_bases, d_off=d_off, offsets=offsets, bases=total_bases, k=K, n=NHASH,
                                     thr=THR, rnd=f.randNumbers, rank=rank, world=world, local_rank=local_rank,
                                     row_base=rank * READS_PER_GPU, rows_per_rank=rows_per_rank)
        if pf is not None:
            view.csr = lambda total: pf.result(lengths.size, total)
        else:
            view.csr = lambda total: f.queryAll(False, fetch=True)
        try:
            parity = bench_legs.parity_of(view, total_now, dist, replicated_check=pf is not None)
        except Exception as e:  # noqa: BLE001
            parity = {"ok": False, "error": f"{type(e).__name__}: {e}"[:300]}

    # ---- side legs: the other BASELINE configs (never part of `value`) ----
    legs = {}
    if not args.no_legs and not args.no_e2e:
        import bench_legs
        low_err = dict(genome_len=GENOME_LEN, p_ins=0.006, p_del=0.006, p_sub=0.008)
        plan = []
        if world == 1:
            plan.append(("rich", "the headline's 100 000 reads at 2 % error (0.6 % ins, 0.6 % del, 0.8 % sub): real candidate lists",
                         READS_PER_GPU, MEAN_LEN, K, NHASH, THR, 1000, low_err))
            plan.append(("c4_k15_n120", "BASELINE configs[3] corner: synthetic 1M reads (~10 kb mean), k=15 n=120 thr=12, 1 GPU",
                         1_000_000, MEAN_LEN, 15, 120, 12, 31, dict(genome_len=GENOME_LEN)))
        else:
            plan.append(("c3", f"BASELINE configs[2]: synthetic 1M reads (~10 Gbases), k=23 n=60 thr=6, split by bases over {world} GPUs",
                         1_000_000, MEAN_LEN, K, NHASH, THR, 31, dict(genome_len=GENOME_LEN)))
        plan.append(("c5", f"BASELINE configs[4]: synthetic ultra-long reads (100 kb mean, ~5 Gbases), k=23 n=60 thr=6, "
                           f"{'1 GPU' if world == 1 else f'split by bases over {world} GPUs'}",
                     50_000, 100_000, K, NHASH, THR, 41, dict(genome_len=GENOME_LEN)))
        for name, what, nreads, mean, k_, n_, thr_, lseed, skw in plan:
            leg = None
            try:
                leg = bench_legs.Leg(name, what, nreads, mean, k_, n_, thr_, rank, world, local_rank, RAND_SEED, lseed, skw)
                legs[name] = bench_legs.run_leg(leg, 3, dist, with_parity=not args.no_parity)
            except Exception as e:  # noqa: BLE001
                legs[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            finally:
                if leg is not None:
                    try:
                        leg.close()
                    except Exception:  # noqa: BLE001
                        pass
                torch.cuda.empty_cache()

    # ---- the kernel the INT32 roof describes: brute force, every k-mer against every hash (N=1 only) ----
    roofline_brute = None
    if world == 1 and not args.no_legs and not args.no_e2e and args.sketch_mode == 0:
        try:
            f.sketchMode = 1
            check(lib().nsmh_set_sketch_mode(f._h, 1))
            for _ in range(2):
                f.load_device(d_bases.data_ptr(), d_off.data_ptr(), lengths.size, total_bases)
                f.sketch()
            bms = f.stats()["sketch_main_ms"]
            f.sketchMode = 0
            check(lib().nsmh_set_sketch_mode(f._h, 0))
            f.load_device(d_bases.data_ptr(), d_off.data_ptr(), lengths.size, total_bases)
            f.sketch()
            f.build()
            roofline_brute = {"kernel": "sketch_brute_kernel", "kernel_ms": bms}
        except Exception as e:  # noqa: BLE001
            roofline_brute = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- roofline of the dominant kernel (sketch_filter_kernel / sketch_brute_kernel) ----
    hbm_peak, peak_src, sm_max = peaks()
    mean_len = total_bases / lengths.size
    alg_bytes = total_bases * 0.25 + lengths.size * NHASH * 8          # packed bases in + sketch rows out
    sk_ms = st["sketch_main_ms"]
    achieved = alg_bytes / (sk_ms * 1e-3) / 1e9 if sk_ms > 0 else 0.0
    # the reference's operation count on the INT32 pipe: (6n+6) lane-ops per base
    int_ops = total_bases * (6 * NHASH + 6)
    int_peak = 148 * 64 * sm_max * 1e6          # ALU pipe: 16 lanes/clk/SMSP (B300_MICROARCH rt_SMSP=2)
    int_measured = None                         # tools/micro/int_roof on a B200, when a copy has been committed
    try:
        with open(os.path.join(ROOT, "profiles", "int_roof.json")) as fh:
            int_measured = json.loads(fh.read().strip().splitlines()[-1])
    except (OSError, ValueError, IndexError):
        pass
    dominant = "sketch_filter_kernel" if args.sketch_mode == 0 else "sketch_brute_kernel"
    rec = recorded_counters().get(dominant, {}) if rank == 0 else {}
    cyc_per_s = 148 * sm_max * 1e6
    int_xormin = (int_measured or {}).get("xormin64_lane_tops_survey_count")
    achieved_tops = int_ops / (sk_ms * 1e-3) / 1e12 if sk_ms > 0 else None
    roofline = {"bound": "hbm", "kernel": dominant,
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "peak_source": peak_src,
                "traffic": rec.get("dram_bytes_read", 0) + rec.get("dram_bytes_write", 0) if rec else None,
                "traffic_source": rec.get("source"),
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": sk_ms,
                "kernel_gbases_per_s": total_bases / (sk_ms * 1e-3) / 1e9 if sk_ms > 0 else None,
                # what actually binds the filter kernel (ncu): the shared-memory data pipe, then instruction issue
                "binding_resource": ({"name": "shared-memory data pipe (l1tex wavefronts, 1 per SM and clock)",
                                      "wavefronts_per_launch": rec["smem_wavefronts"],
                                      "frac": rec["smem_wavefronts"] / (sk_ms * 1e-3) / cyc_per_s,
                                      "ncu_pct_of_peak_at_capture": rec.get("l1tex_data_pipe_pct")}
                                     if rec.get("smem_wavefronts") and sk_ms > 0 else None),
                "issue_slots": ({"warp_instructions": rec["warp_instructions"], "peak_per_s": 4 * cyc_per_s,
                                 "frac": rec["warp_instructions"] / (sk_ms * 1e-3) / (4 * cyc_per_s)}
                                if rec.get("warp_instructions") and sk_ms > 0 else None),
                # the reference's operation count, (6n+6) INT32 lane-ops per base, against the
                # MEASURED rate of the 64-bit XOR+MIN pair (tools/micro/int_roof.cu on a B200)
                "int32_reference_count": {"lane_ops_per_base": 6 * NHASH + 6, "achieved_tops": achieved_tops,
                                          "nominal_alu_pipe_tops": int_peak / 1e12,
                                          "measured_xormin64_lane_tops": int_xormin,
                                          "frac_of_measured": achieved_tops / int_xormin if int_xormin and achieved_tops else None,
                                          "note": ("brute-force kernel: this is its utilisation of the integer roof" if args.sketch_mode else
                                                   "the filter kernel skips most (k-mer, hash) pairs exactly: > 1 means faster than any "
                                                   "brute-force kernel could be; see roofline_brute for the kernel this roof describes")}}

    if roofline_brute and roofline_brute.get("kernel_ms"):
        bt = int_ops / (roofline_brute["kernel_ms"] * 1e-3) / 1e12
        roofline_brute.update({"bound": "int32", "achieved": bt, "unit": "T lane-ops/s (6n+6 per base, SURVEY 8(d))",
                               "peak": int_xormin, "peak_source": "profiles/int_roof.json: tools/micro/int_roof.cu on a B200, "
                               "the 64-bit XOR + unsigned MIN pair at 6 lane-ops (2 399 G pairs/s)",
                               "frac": bt / int_xormin if int_xormin else None,
                               "gbases_per_s": total_bases / (roofline_brute["kernel_ms"] * 1e-3) / 1e9})
    if dist is not None:
        dist.barrier()
    if isinstance(pf, shard.PeerPartitionedFilter):
        pf.shutdown()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    line = {
        "metric": "Gbases/s MinHash sketch+lookup", "value": value, "unit": "Gbases/s", "n_gpus": n_gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(n_gpus), "k": K, "num_hash": NHASH, "overlap_sketch_thr": THR,
                   "reads_per_gpu": READS_PER_GPU, "bases_per_gpu": total_bases, "mean_read_len": mean_len,
                   "sketch_mode": "filter" if args.sketch_mode == 0 else "brute",
                   "multi_gpu": ("n/a" if world == 1 else multi), **({"multi_gpu_note": multi_note} if multi_note else {}),
                   "l2": "inputs larger than L2 (1 GB ASCII + 0.25 GB packed per step vs 126 MB L2)",
                   "host": host_numa,
                   "step": "pack + sketch + build tables + bulk forward lookup, CSR left on device"},
        "phases_last_step": phases,
        "e2e": {"value": e2e_value, "unit": "Gbases/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e2e / args.steps,
                "call": ("MinHashReadFilter.initialize(ReadData) [nsmh_initialize_ascii: load + sketch + build, pipelined] + queryAll, "
                         "pinned host ASCII in, CSR in host memory out" if world == 1 else
                         "nsmh_load_sketch_ascii (pipelined) + multi-GPU build + queryAll, pinned host ASCII in, CSR in host memory out")},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "clocks": clocks,
    }
    if e2e_packed is not None:
        line["e2e_packed"] = e2e_packed
    if parity is not None:
        line["parity"] = parity
    if roofline_brute is not None:
        line["roofline_brute"] = roofline_brute
    if legs:
        line["legs"] = legs

    # ---- FASTQ ingest: a side measurement, never part of `value` ----
    if n_gpus == 1 and not args.no_ingest and not args.no_e2e:
        try:
            line["ingest"] = ingest_leg(local_rank, d_bases, offsets, args.steps, hbm_peak)
        except Exception as e:  # noqa: BLE001
            line["ingest"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- reverse-complement lookup: the production caller queries every window forward AND
    #      reverse-complemented (Consensus.cpp:181-191); the second bulk pass is reported beside the step,
    #      never inside `value`.  It packs the reverse complements, sketches them and probes the same tables.
    if n_gpus == 1 and not args.no_e2e:
        try:
            f.queryAll(True, fetch=False)
            rc_ms = []
            for _ in range(max(1, min(args.steps, 3))):
                rc_total = f.queryAll(True, fetch=False)
                rc_ms.append(f.stats()["query_ms"])
            line["rc_query"] = {"ms": float(np.mean(rc_ms)), "candidate_ids": int(rc_total),
                                "gbases_per_s": total_bases / (float(np.mean(rc_ms)) * 1e-3) / 1e9,
                                "what": "nsmh_query_all(rc=1): reverse-complement pack + sketch + bulk lookup of all reads"}
        except Exception as e:  # noqa: BLE001
            line["rc_query"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- CPU baseline: the reference's own code on this box's host cores, bounded sample ----
    ref_for_online = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        from oracle.oracle import Oracle, RefLib
        os.sched_setaffinity(0, all_cpus)           # the CPU arm gets every core of the box
        cores = os.cpu_count() or 1
        sample_reads = 20_000
        sub = ns.ReadData(host_rd.bases[:int(offsets[sample_reads])].copy(), offsets[:sample_reads + 1].copy())
        rnd = ns.rand_from_seed(RAND_SEED, NHASH)
        t0 = time.perf_counter()
        if RefLib.available():
            rf = RefLib.get().create(sub.bases, sub.offsets, K, NHASH, THR, rnd, threads=cores)
            c_off, c_ids = rf.query_all(0, threads=cores)
            detail = {"sketch_ms": rf.sketch_ms, "build_ms": rf.build_ms, "query_ms": rf.query_ms}
            ref_for_online = rf
            kind = "reference"
        else:
            orc = Oracle.get()
            orc.set_num_threads(cores)
            sk = orc.sketch_all(sub.bases, sub.offsets, K, NHASH, rnd)
            T = orc.build_tables(sk)
            c_off, c_ids = T.query_all(sub.bases, sub.offsets, sk, K, rnd, THR, 0)
            detail = {}
            kind = "port"
        dt = time.perf_counter() - t0
        sb = int(sub.offsets[-1])
        line["cpu_baseline"] = {"value": sb / dt / 1e9, "unit": "Gbases/s", "cores": cores, "kind": kind,
                                "sample": f"first {sample_reads} reads of the workload ({sb / 1e9:.3f} Gbases), "
                                          f"sketch + tables + forward lookup, {cores} OpenMP threads",
                                "seconds": dt, **detail}
    # ---- online query latency: one window per call, as the consensus builder asks ----
    if n_gpus == 1 and not args.no_legs and not args.no_e2e:
        try:
            import bench_legs
            f.load_device(d_bases.data_ptr(), d_off.data_ptr(), lengths.size, total_bases)
            f.sketch()
            f.build()
            f.synchronize()
            line["online_query"] = bench_legs.online_leg(f, host_rd, ref_filter=ref_for_online)
        except Exception as e:  # noqa: BLE001
            line["online_query"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if ref_for_online is not None:
        ref_for_online.close()
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

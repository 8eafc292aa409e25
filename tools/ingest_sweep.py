#!/usr/bin/env python
"""FASTQ-ingest kernels (csrc/fastq.cu) on the bench workload's text, timed with the library's own
CUDA events for several chunk sizes of the pack kernel (NSMH_FQ_PACK_ITERS, read at every call).
Each setting is also checked: same read table and the same packed words as the ASCII load path.

    python tools/ingest_sweep.py [--iters 16 32 ...] [--reps 5] [--warmup 2]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nanospring_b200 as ns  # noqa: E402
from nanospring_b200._lib import check, lib  # noqa: E402


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, nargs="*", default=[8, 16, 32, 64, 128])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    lengths = bench.shard_lengths(0)
    offsets = np.zeros(lengths.size + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lengths, dtype=np.uint64)
    total = int(offsets[-1])
    params = ns.synth_params(genome_len=bench.GENOME_LEN)
    d_off = torch.from_numpy(offsets.astype(np.int64)).cuda()
    d_bases = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
    check(lib().nsmh_synth_reads_device(0, C.byref(params), 0, lengths.size, d_off.data_ptr(), d_bases.data_ptr()))
    text, nbytes, nb, n, _ = bench.fastq_text_device(d_bases, offsets, 25_000)
    want = d_bases[:nb].cpu().numpy()
    hbm_peak = bench.peaks()[0]
    alg_bytes = nbytes + nb + nb / 4 + 8 * n
    rd = ns.GpuReadData(device=0)
    for it in args.iters:
        os.environ["NSMH_FQ_PACK_ITERS"] = str(it)
        for _ in range(args.warmup):
            rd.loadFromDeviceText(text.data_ptr(), nbytes)
        ms, pk = [0.0], [0.0]
        for _ in range(args.reps):
            rd.loadFromDeviceText(text.data_ptr(), nbytes)
            st = rd.stats()
            ms.append(st["fastq_parse_ms"])
            pk.append(st["fastq_pack_ms"])
        ok = bool(rd.getNumReads() == n and (rd.offsets == offsets[:n + 1]).all())
        got = np.frombuffer(b"".join(rd.getRead(i) for i in (0, 1, n // 2, n - 1)), np.uint8)
        ref = np.concatenate([want[int(offsets[i]):int(offsets[i + 1])] for i in (0, 1, n // 2, n - 1)])
        ok = ok and got.size == ref.size and bool((got == ref).all())
        print(json.dumps({"pack_iters": it, "parse_ms": round(float(np.mean(ms)), 4), "pack_kernel_ms": round(float(np.mean(pk)), 4),
                          "gbases_per_s": round(nb / (np.mean(ms) * 1e-3) / 1e9, 1),
                          "roofline_frac": round(alg_bytes / (np.mean(ms) * 1e-3) / 1e9 / hbm_peak, 3), "ok": ok}), flush=True)
    os.environ.pop("NSMH_FQ_PACK_ITERS", None)
    rd.close()


if __name__ == "__main__":
    main()

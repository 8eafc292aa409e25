#!/usr/bin/env python
"""Static SASS evidence for profiles/: per-kernel counts of the mnemonics that matter on sm_100a
(UBLKCP = cp.async.bulk, SYNCS = mbarrier, UBLKPF = bulk L2 prefetch, ATOMS/ATOMG/REDG atomics, ...)
from `cuobjdump -sass` of the objects the Makefile built, and the full listing of one kernel.

    python tools/sass_summary.py > profiles/rN_sass_mnemonics.txt
    python tools/sass_summary.py --listing sketch sketch_filter_kernel > profiles/rN_sass_sketch_filter_kernel.txt
"""
import collections
import os
import re
import subprocess
import sys

BUILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "nanospring_b200", "csrc", "build")
OBJECTS = ["sketch", "query", "table", "multigpu", "pack", "fastq", "prefilter"]
FULL = ("UBLKCP", "UBLKPF", "SYNCS", "ATOMS", "ATOMG", "REDG", "FENCE")
BASE = ("LDS", "STS", "LDG", "STG", "SHFL", "VOTE", "MATCH", "REDUX", "LDGSTS", "UTMALDG", "MEMBAR", "WARPSYNC", "BAR")


def sass(obj):
    return subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj + ".o")], capture_output=True, text=True, check=True).stdout


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()


def summary():
    print("SASS mnemonic counts per kernel (static instruction counts, `cuobjdump -sass nanospring_b200/csrc/build/*.o`, sm_100a)")
    print("UBLKCP = cp.async.bulk (bulk copy through the TMA unit), SYNCS.* = mbarrier, UBLKPF = bulk L2 prefetch,")
    print("ATOMS.CAST.SPIN = shared-memory compare-and-swap loop (what a 64-bit shared atomicMin compiles to)\n")
    for obj in OBJECTS:
        fn, cnt = None, None
        rows = []
        for line in sass(obj).splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                if fn:
                    rows.append((fn, cnt))
                fn, cnt = demangle(m.group(1)), collections.Counter()
                continue
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and fn is not None:
                op = m.group(1)
                base = op.split(".")[0]
                cnt["total"] += 1
                if base in FULL:
                    cnt[op] += 1
                elif base in BASE:
                    cnt[base] += 1
        if fn:
            rows.append((fn, cnt))
        for fn, cnt in rows:
            if "cub::" in fn:
                continue
            print(f"{obj}.o  {fn[:140]}")
            print("    " + "  ".join(f"{k}={v}" for k, v in sorted(cnt.items())))


def listing(obj, kernel):
    on = False
    for line in sass(obj).splitlines():
        if "Function : " in line:
            on = kernel in line
        if on:
            print(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/$", "", line))


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--listing":
        listing(sys.argv[2], sys.argv[3])
    else:
        summary()

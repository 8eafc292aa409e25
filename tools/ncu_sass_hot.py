#!/usr/bin/env python
"""Print the SASS of an `ncu --page source --csv` export with executed-instruction counts and
stall samples, so hot regions can be read off (usage: ncu_sass_hot.py file.csv [min_exec])."""
import csv
import sys


def main(path, min_exec=0):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isrc, isamp, iex, ithr = (hdr.index(x) for x in
                                  ("Address", "Source", "# Samples", "Instructions Executed", "Avg. Threads Executed"))
    total = sum(int(r[iex]) for r in rows[2:] if len(r) > iex and r[iex].isdigit())
    tsamp = sum(int(r[isamp]) for r in rows[2:] if len(r) > isamp and r[isamp].isdigit())
    print(f"total executed warp instructions {total}, samples {tsamp}")
    for n, r in enumerate(rows[2:]):
        if len(r) <= iex or not r[iex].isdigit():
            continue
        ex = int(r[iex])
        if ex < min_exec:
            continue
        print(f"{n:5d} {ex:11d} {100.0 * ex / total:5.2f}% s={int(r[isamp]):6d} thr={r[ithr]:>5s}  {r[isrc].strip()}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)

#!/usr/bin/env python
"""Print the last N launches (kernel, microseconds) of an `ncu --metrics gpu__time_duration.sum --csv` log."""
import csv
import sys


def main(path, last=16):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    seq = [(r[ki][:72], float(r[vi].replace(',', '')) / 1e3) for r in data if len(r) > vi]
    for name, us in seq[-last:]:
        print(f"{us:10.1f} us  {name}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 16)

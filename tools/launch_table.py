#!/usr/bin/env python
"""Print an ncu `--metrics gpu__time_duration.sum --csv` launch list as a table; --last N keeps the last N launches,
   --every K splits the list into K equal steps and prints the last one (a target that runs K identical steps)."""
import csv
import re
import sys


def main():
    path = sys.argv[1]
    every = int(sys.argv[sys.argv.index("--every") + 1]) if "--every" in sys.argv else 1
    rows, hdr = [], None
    for r in csv.reader(open(path, errors="replace")):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        rows.append((re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "")[:100], d["Grid Size"], d["Block Size"], float(d["Metric Value"].replace(",", "")) / 1e3))
    n = len(rows) // every
    rows = rows[len(rows) - n:]
    total = sum(r[3] for r in rows)
    for name, grid, block, us in rows:
        print("%8.1f us  %-18s %-14s %s" % (us, grid, block, name))
    print("total %.1f us over %d launches" % (total, len(rows)))


if __name__ == "__main__":
    main()

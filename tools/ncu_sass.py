#!/usr/bin/env python
"""List the SASS of an `ncu --page source --print-source cuda,sass --csv` dump in address order with
executed counts, restricted to source lines [lo, hi] of one file (hot-loop inspection)."""
import csv
import sys


def main(path, fname, lo, hi, min_exec=1):
    rows = list(csv.reader(open(path)))
    cur = hdr = None
    line = None
    out = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            ai, si, ii = 2, hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr is None:
            continue
        if r[0] != "":
            line = int(r[0])
            continue
        if cur != fname or line is None or not (lo <= line <= hi):
            continue
        try:
            addr = int(r[ai], 16)
            n = int(r[ii] or 0)
            s = int(r[si] or 0)
        except ValueError:
            continue
        if n >= min_exec:
            out.append((addr, line, n, s, r[3].strip()))
    out.sort()
    tot = sum(o[2] for o in out)
    print(f"{len(out)} instructions, {tot} executed")
    for a, l, n, s, txt in out:
        print(f"{a & 0xfffff:06x} L{l:<4d} {n:>12d} {s:>7d}  {txt[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]) if len(sys.argv) > 5 else 1)

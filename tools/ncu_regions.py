#!/usr/bin/env python
"""Instruction / sample share per source file and per named line range of an ncu cuda,sass csv dump."""
import csv
import sys
from collections import defaultdict


def load(path):
    rows = list(csv.reader(open(path)))
    cur = hdr = key = None
    agg = defaultdict(lambda: [0, 0, 0, ""])
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
            si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            continue
        if hdr is None:
            continue
        if r[0] != "":
            key = (cur, int(r[0]))
            agg[key][3] = r[1]
            continue
        if key is None or len(r) <= ti:
            continue
        try:
            agg[key][0] += int(r[si] or 0)
            agg[key][1] += int(r[ii] or 0)
            agg[key][2] += int(r[ti] or 0)
        except ValueError:
            pass
    return agg


def main(path, src):
    agg = load(path)
    tot_i = sum(v[1] for v in agg.values())
    tot_s = sum(v[0] for v in agg.values())
    # regions = functions / phases found by scanning the source for marker lines
    lines = open(src).read().splitlines()
    marks = []
    for n, l in enumerate(lines, 1):
        ls = l.strip()
        if ls.startswith("// ----") and len(ls) > 12 and not ls.startswith("// -------------"):
            marks.append((n, ls[7:50]))
        elif ("__device__" in l or "__global__" in l) and "(" in l and not ls.startswith("//"):
            marks.append((n, ls[:60]))
    marks.append((len(lines) + 1, "END"))
    fname = src.split("/")[-1]
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    for (a, name), (b, _) in zip(marks[:-1], marks[1:]):
        i = sum(v[1] for k, v in agg.items() if k[0] == fname and a <= k[1] < b)
        s = sum(v[0] for k, v in agg.items() if k[0] == fname and a <= k[1] < b)
        t = sum(v[2] for k, v in agg.items() if k[0] == fname and a <= k[1] < b)
        if i:
            print(f"{fname}:{a:<4d} inst {100 * i / tot_i:5.1f}%  samples {100 * s / tot_s:5.1f}%  thr/inst {t / i:4.1f}  {name}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

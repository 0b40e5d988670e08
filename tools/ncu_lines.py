#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line."""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur_file, hdr = None, None
    agg = defaultdict(lambda: [0, 0, 0, ""])
    cur_key = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            si = hdr.index("# Samples")
            ii = hdr.index("Instructions Executed")
            ti = hdr.index("Thread Instructions Executed")
            continue
        if hdr is None:
            continue
        if r[0] != "":
            cur_key = (cur_file, int(r[0]))
            agg[cur_key][3] = r[1]
            continue
        if cur_key is None or len(r) <= ti:
            continue
        try:
            agg[cur_key][0] += int(r[si] or 0)
            agg[cur_key][1] += int(r[ii] or 0)
            agg[cur_key][2] += int(r[ti] or 0)
        except ValueError:
            pass
    tot_s = sum(v[0] for v in agg.values()) or 1
    tot_i = sum(v[1] for v in agg.values()) or 1
    print(f"total samples {tot_s}  total warp instructions {tot_i}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * v[0] / tot_s:5.1f}% samp {100 * v[1] / tot_i:5.1f}% inst  thr/inst {v[2] / max(v[1], 1):4.1f}  "
              f"{k[0]}:{k[1]:<4d} {v[3].strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)

#!/usr/bin/env python
"""Top source lines of an .ncu-rep by executed instructions / stall samples (needs -lineinfo)."""
import csv
import subprocess
import sys
from collections import defaultdict


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    agg = defaultdict(lambda: [0, 0, ""])
    cur_file, hdr = "", None
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ie, sm = r.index("Instructions Executed"), r.index("# Samples")
            continue
        if hdr is None or len(r) <= ie:
            continue
        if r[0]:  # a source line row (may carry its own totals) — remember text
            line = r[0]
            agg[(cur_file, line)][2] = r[1][:100]
            cur_line = line
        try:
            n, s = int(r[ie] or 0), int(r[sm] or 0)
        except ValueError:
            continue
        if r[2]:  # SASS row under the current source line
            agg[(cur_file, cur_line)][0] += n
            agg[(cur_file, cur_line)][1] += s
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    print(f"total warp-inst {tot}  samples {tots}")
    for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*v[0]/tot:5.1f}% inst {100*v[1]/tots:5.1f}% smp  {f}:{ln:>4s}  {v[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)

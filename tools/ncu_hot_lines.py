#!/usr/bin/env python
"""Hot CUDA source lines of a kernel from an .ncu-rep captured with --import-source on (-lineinfo build):
tools/ncu_hot_lines.py report.ncu-rep [top] -> per-file shares and the top lines by warp-stall samples / instructions."""
import csv
import subprocess
import sys
from collections import defaultdict


def main(path, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                         text=True).stdout
    cur, hdr, out = None, None, []
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur, hdr = r[1], None
        elif r[0] == "Line No":
            hdr = r
        elif hdr and cur and r[0] not in ("", "Function Name"):
            d = {}
            for k, v in zip(hdr, r):
                d.setdefault(k, v)
            try:
                out.append((int(d["# Samples"]), int(d["Instructions Executed"]), cur.split("/")[-1], d["Line No"], d["Source"][:100]))
            except (KeyError, ValueError):
                pass
    ts, ti = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
    files = defaultdict(lambda: [0, 0])
    for s, ie, fn, _, _ in out:
        files[fn][0] += s
        files[fn][1] += ie
    print(f"samples {ts}  warp instructions {ti}")
    for fn, (s, ie) in sorted(files.items(), key=lambda kv: -kv[1][0]):
        print(f"  {fn:28s} samples {100 * s / ts:5.1f} %  inst {100 * ie / ti:5.1f} %")
    for s, ie, fn, ln, src in sorted(out, reverse=True)[:top]:
        print(f"{100 * s / ts:5.1f} % smp {100 * ie / ti:5.1f} % inst  {fn}:{ln}  {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)

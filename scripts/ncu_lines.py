#!/usr/bin/env python
"""Per SOURCE LINE totals (samples, warp-instructions) of one .ncu-rep captured with --import-source on."""
import csv
import subprocess
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
    lines = [(r[0], r[1], num(r[iS]), num(r[iI])) for r in rows if len(r) > iI and r[0].isdigit()]
    ts, ti = sum(l[2] for l in lines) or 1, sum(l[3] for l in lines) or 1
    print(f"total samples {ts}  warp-instructions {ti}")
    key = 2 if (len(sys.argv) > 3 and sys.argv[3] == "smp") else 3
    for ln, src, s, i in sorted(lines, key=lambda l: -l[key])[:top]:
        print(f"{100.0 * i / ti:5.1f}% inst {100.0 * s / ts:5.1f}% smp  L{ln:>4}  {src.strip()[:110]}")


if __name__ == "__main__":
    main()

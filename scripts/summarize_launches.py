#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md
ncu times are cold-cache and serialised: compare SHARES with bench.py's live CUDA-event shares, not absolutes."""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"(?:stgcma::)?([\w:]+)(<[^(]*>)?\(", name)
    if not m:
        return name[:60]
    base, targs = m.group(1), m.group(2) or ""
    targs = targs.replace("__nv_bfloat16", "bf16").replace(" ", "")
    if base.startswith("at::") or "at::" in name[:20]:
        return "torch:" + base.split("::")[-1]
    return base + (targs if len(targs) < 40 else "")


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    header = next(rd)
    ik, iv, ig, ib = header.index("Kernel Name"), header.index("Metric Value"), header.index("Grid Size"), header.index("Block Size")
    for r in rd:
        try:
            rows.append((short(r[ik]), float(r[iv].replace(",", "")), r[ig], r[ib]))
        except (ValueError, IndexError):
            continue
    agg = OrderedDict()
    for k, ns, g, b in rows:
        d = agg.setdefault(k, dict(n=0, ns=0.0, mx=0.0))
        d["n"] += 1
        d["ns"] += ns
        d["mx"] = max(d["mx"], ns)
    tot = sum(d["ns"] for d in agg.values())
    print(f"launches: {len(rows)}   total device time (serialised, cold cache): {tot / 1e6:.3f} ms\n")
    print("| kernel | launches | total ms | share | avg us | max us |")
    print("|---|---:|---:|---:|---:|---:|")
    for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        print(f"| `{k}` | {d['n']} | {d['ns'] / 1e6:.3f} | {100 * d['ns'] / tot:.1f}% | {d['ns'] / d['n'] / 1e3:.1f} | {d['mx'] / 1e3:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])

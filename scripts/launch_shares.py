"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals and
shares.  Used to write profiles/*_launches.md next to the raw csv.

    python scripts/launch_shares.py gpurun_out/launches.csv [steps]
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    m = re.match(r"(?:void )?([A-Za-z0-9_]+)(<[^(]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def main(path, steps=1):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    ig, ib = hdr.index("Grid Size"), hdr.index("Block Size")
    for r in rd:
        if r[im] != "gpu__time_duration.sum":
            continue
        rows.append((short(r[ik]), float(r[iv].replace(",", "")), r[ig], r[ib]))
    agg = OrderedDict()
    for k, ns, g, b in rows:
        a = agg.setdefault(k, [0, 0.0, 0.0, g, b])
        a[0] += 1
        a[1] += ns
        a[2] = max(a[2], ns)
    total = sum(a[1] for a in agg.values())
    print(f"{len(rows)} launches, {total / 1e6:.3f} ms of kernel time under ncu (serialised, cold cache)"
          f"{'' if steps == 1 else f', {steps} steps'}\n")
    print("| kernel | launches | total ms | share | mean us | max us | grid x block (last) |")
    print("|---|---:|---:|---:|---:|---:|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / total:.1f} % | {a[1] / a[0] / 1e3:.1f} | "
              f"{a[2] / 1e3:.1f} | {a[3]} x {a[4]} |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)

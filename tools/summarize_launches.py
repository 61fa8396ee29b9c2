#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (name, grid)."""
import collections
import csv
import re
import sys


def main(path, title):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
        key = (name, row.get("Grid Size", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# {title}\n")
    print(f"Source: `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised launches: compare shares).\n")
    print("| kernel | grid | launches | total ms | avg us | share |\n|---|---|---:|---:|---:|---:|")
    for (k, g), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if a[1] / tot < 0.0005:
            continue
        print(f"| `{k}` | {g} | {a[0]} | {a[1] / 1e6:.3f} | {a[1] / a[0] / 1e3:.1f} | {100 * a[1] / tot:.1f}% |")
    print(f"\nTotal kernel time: {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches.")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "launch summary")

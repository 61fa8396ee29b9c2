#!/usr/bin/env python
"""Rebuild the table of profiles/r01_results.md from the bench JSON lines in profiles/ (python tools/results_table.py)."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fmt(v, nd=1):
    return "" if v is None else f"{v:.{nd}f}"


def main():
    print("| file | impl | workload | dtype | GPUs | value | unit | ms/step | e2e (host buffers) | e2e uint8 ext. | CPU oracle port | cores | "
          "dominant-kernel frac of measured tensor peak | kernel ms | optimizer tail frac of measured HBM peak |")
    print("|---|---|---|---|---:|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_*.json"))):
        try:
            d = json.loads(open(path).read().strip().splitlines()[-1])
        except (ValueError, IndexError):
            continue
        if not isinstance(d, dict) or "value" not in d:
            continue
        roof, cpu, tail = d.get("roofline") or {}, d.get("cpu_baseline") or {}, d.get("roofline_optimizer_tail") or {}
        print("| " + " | ".join([
            os.path.basename(path), d.get("impl", "b200"), d["config"].get("workload", ""), d.get("dtype", ""), str(d["n_gpus"]),
            fmt(d["value"]), d["unit"], fmt(d["ms_per_step"], 3), fmt(d["e2e"]["value"]), fmt((d.get("e2e_uint8") or {}).get("value")),
            fmt(cpu.get("value"), 2), str(cpu.get("cores", "")), fmt(roof.get("frac"), 3), fmt(roof.get("kernel_ms"), 4),
            fmt(tail.get("frac"), 3)]) + " |")


if __name__ == "__main__":
    main()

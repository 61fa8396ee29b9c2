#!/usr/bin/env python
"""Summarise `ncu --set full` reports (raw page) into a small JSON/markdown: duration, DRAM bytes, DRAM %, tensor-pipe %."""
import csv, io, json, subprocess, sys

KEYS = {"gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct", "launch__registers_per_thread": "regs",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct"}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    for row in r[2:]:
        d = {"kernel": row[hdr.index("Kernel Name")][:90], "grid": row[hdr.index("Grid Size")]}
        for k, name in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[name] = float(row[i].replace(",", "")) * UNIT.get(units[i], 1.0)
                except ValueError:
                    pass
        if "time_us" in d and "dram_read_bytes" in d:
            d["dram_gbs"] = (d["dram_read_bytes"] + d["dram_write_bytes"]) / d["time_us"] / 1e3
        yield d


if __name__ == "__main__":
    res = [d for p in sys.argv[1:] for d in rows(p)]
    print(json.dumps(res, indent=1))

"""Achieved bandwidth of the bandwidth-bound kernels (north star: "justified by ... achieved HBM GB/s against B200 peak").

Reads an ncu CSV taken with
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv ...
and prints, per kernel name: launches, mean duration, DRAM GB/s (read + write) and L2 GB/s, as a fraction of the measured
HBM peak (MEASURED_PEAKS.json).  Small kernels whose working set lives in L2 show up with a low DRAM figure and a high
L2 figure -- both are reported.

    python tools/bandwidth_report.py gpurun_out/r2_bw.csv > profiles/r2_bandwidth_kernels.md
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}.get(unit, 1)


def main():
    path = sys.argv[1]
    peak = 6453.1
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, idi, mi, ui, vi = h.index("Kernel Name"), h.index("ID"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
    per = collections.defaultdict(dict)
    name_of = {}
    for r in rows[1:]:
        name_of[r[idi]] = r[ki]
        per[r[idi]][r[mi]] = (r[vi], r[ui])
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i, m in per.items():
        if "gpu__time_duration.sum" not in m:
            continue
        us = to_us(*m["gpu__time_duration.sum"])
        dram = to_bytes(*m.get("dram__bytes_read.sum", ("0", "byte"))) + to_bytes(*m.get("dram__bytes_write.sum", ("0", "byte")))
        l2 = to_bytes(*m.get("lts__t_bytes.sum", ("0", "byte")))
        name = name_of[i].split("(")[0].replace("void ", "").replace("bb::<unnamed>::", "").replace("unnamed>::", "")
        a = agg[name]
        a[0] += 1
        a[1] += us
        a[2] += dram
        a[3] += l2
    print(f"| kernel | launches | mean us | DRAM GB/s | of HBM peak ({peak:.0f}) | L2 GB/s |")
    print("|---|---|---|---|---|---|")
    for name, (n, us, dram, l2) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = dram / (us * 1e-6) / 1e9 if us else 0
        l2gbs = l2 / (us * 1e-6) / 1e9 if us else 0
        print(f"| `{name}` | {n} | {us / n:.1f} | {gbs:.0f} | {gbs / peak:.2f} | {l2gbs:.0f} |")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Kernel name, duration and DRAM bytes per launch from `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`.
usage: python tools/ncu_kernel_times.py launches.csv"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ik, im, iv, iu, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
d = {}
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    if u in ("ns", "nsecond"): v /= 1e6
    elif u in ("us", "usecond"): v /= 1e3
    elif u in ("s", "second"): v *= 1e3
    elif u == "Kbyte": v *= 1e3
    elif u == "Mbyte": v *= 1e6
    elif u == "Gbyte": v *= 1e9
    d.setdefault((int(r[iid]), r[ik]), {})[r[im]] = v
for (i, k), m in sorted(d.items()):
    t = m.get("gpu__time_duration.sum", 0.0)
    b = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    print("%3d %-46s %8.3f ms  %6.2f GB dram  %6.0f GB/s" % (i, k[:46], t, b / 1e9, b / 1e6 / t if t else 0))

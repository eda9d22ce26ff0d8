#!/usr/bin/env python
"""Markdown table (kernel, launches, total ms, share) from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <cmd>`).
usage: python tools/ncu_launch_md.py launches.csv "<cmd>" > launches.md"""
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) if len(r) > 10]
h = rows[0]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(r[iu], 1.0)
    tot[r[ik]] += v
    cnt[r[ik]] += 1
allms = sum(tot.values())
print("# ncu launch list — `%s` (B200, --clock-control none)\n" % (sys.argv[2] if len(sys.argv) > 2 else "?"))
print("Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, v in tot.most_common():
    print("| `%s` | %d | %.3f | %.1f%% |" % (k[:70], cnt[k], v, 100 * v / allms))

#!/usr/bin/env python
"""Per-CUDA-source-line instruction and stall-sample totals from an .ncu-rep (needs -lineinfo).
usage: python tools/ncu_lines.py rep.ncu-rep [top_n]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = None; hdr = None
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter(), ""])
tot_i = tot_s = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples"); st = {n: i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n}; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != "" and r[2] == "-":  # a CUDA source line (aggregate over its SASS, inlined callees included)
        try: n = float(r[ci] or 0); s = float(r[cs] or 0)
        except ValueError: continue
        key = (cur_file, r[0]); a = agg[key]; a[0] += n; a[1] += s; a[3] = r[1].strip()[:100]
        for k, i in st.items(): a[2][k] += float(r[i] or 0)
        tot_i = max(tot_i, 0); tot_s = max(tot_s, 0)
tot_i = float(sys.argv[3]) if len(sys.argv) > 3 else max(a[0] for a in agg.values()); tot_s = max(a[1] for a in agg.values())
print("normalising to warp-inst %.4g samples %d (largest line)" % (tot_i, tot_s))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    top = " ".join("%s=%d" % (k[6:], v) for k, v in a[2].most_common(3) if v)
    print("%5.1f%%i %5.1f%%s %s:%s  %s  [%s]" % (100 * a[0] / tot_i, 100 * a[1] / max(tot_s, 1), key[0], key[1], a[3], top))

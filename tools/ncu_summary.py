#!/usr/bin/env python
"""Summarise an .ncu-rep (first kernel): headline metrics, opcode mix and stall reasons from the SASS page.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [voxel_views]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
vv = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct"]
for name in want:
    if name in h:
        i = h.index(name)
        print("%-70s %-10s %s" % (name, u[i], v[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hh = rows[1]
ci, cs, so = hh.index("Instructions Executed"), hh.index("# Samples"), hh.index("Source")
st = {n: hh.index(n) for n in hh if n.startswith("stall_") and "Not Issued" not in n}
byop, bys, stall = collections.Counter(), collections.Counter(), collections.Counter()
tot = tots = 0
for r in rows[2:]:
    try:
        n, s = float(r[ci]), float(r[cs])
    except (ValueError, IndexError):
        continue
    op = r[so].split()
    o = op[0] if not op[0].startswith("@") else op[1]
    parts = o.split(".")
    o = parts[0] + ("." + parts[1] if parts[0] in ("F2F", "F2I", "I2F", "LDS", "LDG", "STS", "STG", "MUFU", "FRND") and len(parts) > 1 else "")
    byop[o] += n
    bys[o] += s
    tot += n
    tots += s
    for k, vi in st.items():
        stall[k] += float(r[vi] or 0)
print("SASS lines %d, warp-inst %.4g%s" % (len(rows) - 2, tot, ", per 32 voxel-views %.1f" % (tot / (vv / 32)) if vv else ""))
for o, n in byop.most_common(24):
    print("  %-12s %5.1f%% inst %5.1f%% samples%s" % (o, 100 * n / tot, 100 * bys[o] / max(tots, 1), "  %.1f/32vv" % (n / (vv / 32)) if vv else ""))
print("stalls:", ", ".join("%s=%d" % (k[6:], x) for k, x in stall.most_common(9)))

#!/usr/bin/env python
"""Summarise an .ncu-rep (first kernel): key metrics, per-opcode executed
instructions per warp-surface and top stall sites.  usage:
    python scripts/ncu_summary.py rep.ncu-rep RAYS SURFACES [out.txt]"""
import csv, io, subprocess, sys
from collections import Counter
rep, rays, S = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
out = open(sys.argv[4], "w") if len(sys.argv) > 4 else sys.stdout
def run(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(run("raw")))); hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__shared_mem_per_block_dynamic','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active','smsp__inst_executed_op_tma_st.sum','sm__cycles_elapsed.avg.per_second','dram__cycles_elapsed.avg.per_second','lts__t_sectors_srcunit_ltcfabric.sum','lts__t_sectors_op_write.sum','lts__t_sectors_srcunit_tex_op_write.sum']
print("# %s  rays=%g S=%d" % (rep, rays, S), file=out)
for i, h in enumerate(hdr):
    if h in keys or 'pcsamp_warps_issue_stalled' in h and float(vals[i].replace(',','') or 0) > 500:
        print("%s [%s] = %s" % (h, units[i], vals[i]), file=out)
rows = list(csv.reader(io.StringIO(run("source")))); hdr = rows[1]; data = rows[2:]
isrc, iex, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
ws = rays/32*S
tot = sum(int(r[iex]) for r in data); tots = sum(int(r[ist]) for r in data)
print("instructions per warp(32 rays)-surface: %.1f   static instructions: %d" % (tot/ws, len(data)), file=out)
c = Counter(); s = Counter()
for r in data:
    t = r[isrc].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += int(r[iex])/ws; s[op] += 100*int(r[ist])/tots
print("opcode: executed per warp-surface | stall-sample %", file=out)
for op, v in c.most_common(28):
    print("  %-10s %7.2f | %5.1f" % (op, v, s[op]), file=out)
print("top stall sites:", file=out)
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for n in sorted(range(len(data)), key=lambda n: -int(data[n][ist]))[:16]:
    r = data[n]
    why = max(names, key=lambda k: int(r[hdr.index(k)] or 0))
    print("  %5.2f%%  %-22s %s" % (100*int(r[ist])/tots, why, r[isrc].strip()), file=out)

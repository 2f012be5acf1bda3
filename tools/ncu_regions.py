#!/usr/bin/env python3
"""Headline counters and a per-source-line table of one kernel of an ncu report (captured with --import-source on).
usage: ncu_regions.py report.ncu-rep kernel-regex [top N]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
hdr, units = rows[0], rows[1]
import re
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if not re.search(kern, d["Kernel Name"]):
        continue
    print("==", d["Kernel Name"])
    for w in WANT + [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]:
        if w in d:
            try:
                v = float(d[w].replace(",", ""))
            except ValueError:
                continue
            if "stalled" in w and v < 0.2:
                continue
            print("  %-88s %14.3f %s" % (w, v, units[hdr.index(w)]))
    break
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


cur, L = None, []
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0].isdigit() and len(r) > 8:
        L.append((cur, int(r[0]), r[1].strip(), num(r[7]), num(r[6]), num(r[8])))
ti = sum(l[3] for l in L) or 1
ts = sum(l[4] for l in L) or 1
print("source lines: %d warp instructions, %d samples" % (ti, ts))
for f, n, src, inst, samp, tinst in sorted(L, key=lambda l: -l[4])[:top]:
    print("%-22s %4d %6.2f%% inst %6.2f%% samp act=%4.1f  %s" % (f, n, 100.0 * inst / ti, 100.0 * samp / ts, tinst / max(inst, 1), src[:90]))

#!/usr/bin/env python3
"""Top source lines of a kernel by executed warp instructions / stall samples, from an ncu report captured with
--import-source on (the kernel must be compiled with -lineinfo).  usage: ncu_lines.py report.ncu-rep [top N] [kernel regex]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = ["-k", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
out = subprocess.run(["ncu", "-i", rep] + kern + ["--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines = []
cur_file = None
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_src, i_inst, i_samp = 1, r.index("Instructions Executed"), r.index("# Samples")
        continue
    if hdr and r and r[0] not in ("", "Line No") and r[0].isdigit() and len(r) > i_inst:
        try:
            lines.append((cur_file, int(r[0]), r[i_src].strip(), int(r[i_inst]), int(r[i_samp])))
        except ValueError:
            pass
ti = sum(l[3] for l in lines) or 1
ts = sum(l[4] for l in lines) or 1
print("total warp instructions %d, samples %d" % (ti, ts))
print("%-22s %5s %7s %7s  %s" % ("file", "line", "inst%", "samp%", "source"))
for f, n, src, inst, samp in sorted(lines, key=lambda l: -l[3])[:top]:
    print("%-22s %5d %6.2f%% %6.2f%%  %s" % (f, n, 100.0 * inst / ti, 100.0 * samp / ts, src[:110]))

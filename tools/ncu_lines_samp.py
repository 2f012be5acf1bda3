#!/usr/bin/env python3
"""Like ncu_lines.py but sorted by stall samples.  usage: ncu_lines_samp.py report.ncu-rep [top N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines, cur_file, hdr = [], None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; i_src, i_inst, i_samp = 1, r.index("Instructions Executed"), r.index("# Samples"); continue
    if hdr and r and r[0].isdigit() and len(r) > i_inst:
        try: lines.append((cur_file, int(r[0]), r[i_src].strip(), int(r[i_inst]), int(r[i_samp])))
        except ValueError: pass
ti = sum(l[3] for l in lines) or 1; ts = sum(l[4] for l in lines) or 1
print("total warp instructions %d, samples %d" % (ti, ts))
for f, n, src, inst, samp in sorted(lines, key=lambda l: -l[4])[:top]:
    print("%-22s %5d %6.2f%% inst %6.2f%% samp  %s" % (f, n, 100.0 * inst / ti, 100.0 * samp / ts, src[:100]))

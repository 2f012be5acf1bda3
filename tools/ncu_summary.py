#!/usr/bin/env python3
"""Summarise ncu captures brought back in gpurun_out/ into profiles/ (tracked).
usage: python tools/ncu_summary.py <round tag> <name>=<file.ncu-rep>:<log of the bench run under ncu> ...  [launches.csv]
The bench log's JSON line supplies the algorithmic bytes per kernel (roofline.kernels) of the very batch that was captured."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, zip(r, units))) for r in rows[2:]]


def main():
    tag = sys.argv[1]
    traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    lines = ["# ncu summaries, round %s\n" % tag,
             "Captured with `ncu --set full --clock-control none --import-source on -k regex:bro_` under gpurun",
             "(see tools/gpu_round.sh); per-launch values of every captured launch of each report. Times under ncu are",
             "cold-cache and serialised: use them for shares and counters, never as bench values.\n"]
    for arg in sys.argv[2:]:
        if "=" not in arg:
            continue
        name, rest = arg.split("=", 1)
        rep, benchlog = rest.rsplit(":", 1)
        line = [l for l in open(benchlog).read().splitlines() if l.startswith("{")][-1]
        kern_algo = {k.split(" ")[0].split("_kernel")[0]: v["algorithmic_bytes"] for k, v in json.loads(line)["roofline"]["kernels"].items()}
        launches = raw(rep)
        lines.append("## %s (`%s`, %d launch(es) captured)\n" % (name, os.path.basename(rep), len(launches)))
        per_kernel = {}
        for L in launches:
            kn = L["Kernel Name"][0]
            lines.append("kernel `%s`\n" % kn)
            lines.append("| metric | value | unit |\n|---|---|---|")
            for m in WANT:
                if m in L:
                    v, u = L[m]
                    lines.append("| %s | %s | %s |" % (m, v, u))
            lines.append("")

            def num(m):
                v, u = L[m]
                f = float(v.replace(",", ""))
                return f * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(u, 1)
            key = kn.replace("void ", "").split("<")[0].split("(")[0].split("_kernel")[0]
            algo = kern_algo.get(key)
            dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
            if algo:
                lines.append("algorithmic bytes of this launch: %.0f; DRAM traffic (read+write): %.0f; traffic / algorithmic = %.3f\n" % (algo, dram, dram / algo))
                per_kernel[key + "_kernel"] = {"dram_bytes_per_algorithmic_byte": dram / algo, "dram_read": num("dram__bytes_read.sum"),
                                               "dram_write": num("dram__bytes_write.sum"), "algorithmic_bytes": algo}
            else:
                lines.append("DRAM traffic (read+write): %.0f\n" % dram)
        traffic[name] = {"kernels": per_kernel, "round": tag, "report": os.path.basename(rep)}
    with open(os.path.join(ROOT, "profiles", "%s_ncu_summary.md" % tag), "w") as f:
        f.write("\n".join(lines))
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    for arg in sys.argv[2:]:
        if arg.endswith(".csv"):
            txt = open(arg).read()
            rows = list(csv.reader(io.StringIO(txt[txt.index('"ID"'):])))
            hdr = rows[0]
            ik, iv, ib, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Block Size"), hdr.index("Grid Size")
            out = ["id,kernel,block,grid,gpu__time_duration.sum_ns"]
            tot, mine = 0.0, 0.0
            for r in rows[1:]:
                k = r[ik].split("(")[0][:80].replace(",", ";")
                ns = float(r[iv].replace(",", ""))
                tot += ns
                if "bro_" in r[ik]:
                    mine += ns
                out.append("%s,%s,%s,%s,%.0f" % (r[0], k, r[ib].replace(",", " "), r[ig].replace(",", " "), ns))
            out.append("# total %.0f ns; bro_* kernels %.0f ns (%.1f%% of all launches listed, which include the harness's torch setup/verification kernels)" % (tot, mine, 100 * mine / tot))
            open(os.path.join(ROOT, "profiles", "%s_launches.csv" % tag), "w").write("\n".join(out) + "\n")
    print(open(os.path.join(ROOT, "profiles", "%s_ncu_summary.md" % tag)).read()[:6000])


if __name__ == "__main__":
    main()

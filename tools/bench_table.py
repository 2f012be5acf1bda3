#!/usr/bin/env python3
"""Collect the JSON lines of a GPU round (gpurun_out/bench_*.json, scale_*.json) into profiles/<tag>_bench_lines.jsonl,
profiles/<tag>_bench.md and profiles/<tag>_scaling.md.  usage: python tools/bench_table.py <tag>"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = os.path.join(ROOT, "gpurun_out")


def last_json(path):
    try:
        return json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])
    except Exception:
        return None


rows, lines = [], []
for f in sorted(glob.glob(os.path.join(out, "bench_*.json"))):
    j = last_json(f)
    if j:
        j["_file"] = os.path.basename(f)
        rows.append(j)
        lines.append(json.dumps(j))
open(os.path.join(ROOT, "profiles", "%s_bench_lines.jsonl" % tag), "w").write("\n".join(lines) + "\n")
md = ["# Bench lines, round %s (1 x B200, fresh gpurun box; commands in tools/gpu_round.sh)\n" % tag,
      "| workload | impl / path | value GB/s | ms/step | kernels (ms, HBM roofline frac of measured peak) | dominant kernel | DRAM traffic / algorithmic | e2e GB/s | cpu_baseline GB/s (cores) | gpu_launches | SM MHz |",
      "|---|---|---|---|---|---|---|---|---|---|---|"]
for j in rows:
    cfg = j["config"]
    if j.get("impl") == "reference":
        md.append("| %s | reference (oracle port, %d threads) | %.2f | %.1f | - | - | - | %.2f | %.2f (%d) | 0 | - |" % (
            cfg["workload"], j["cpu_baseline"]["cores"], j["value"], j["ms_per_step"], j["e2e"]["value"], j["cpu_baseline"]["value"], j["cpu_baseline"]["cores"]))
        continue
    r = j["roofline"]
    ks = "; ".join("%s %.3f%s" % (k.replace("bro_", "").replace("_kernel", "").split(" ")[0], v["ms"], (" (%.3f)" % v["frac"]) if v["frac"] else "")
                   for k, v in r["kernels"].items())
    tr = ("%.2f" % (r["traffic"] / r["algorithmic_bytes_per_launch"])) if r.get("traffic") else "-"
    cpu = j.get("cpu_baseline")
    md.append("| %s | b200 / %s | %.1f | %.3f | %s | %s (frac %.3f, %.0f%% of step) | %s | %s | %s | %d | %s |" % (
        cfg["workload"], cfg.get("mode", "auto"), j["value"], j["ms_per_step"], ks, r["kernel"], r["frac"], 100 * r["kernel_share_of_step"], tr,
        ("%.1f" % j["e2e"]["value"]) if j.get("e2e") else "-", ("%.2f (%d)" % (cpu["value"], cpu["cores"])) if cpu else "-",
        j["gpu_launches"], j["clocks"].get("sm_mhz")))
md.append("\n`value` is kernel-only (inputs and outputs resident in HBM), `e2e` goes through bro_batch_decode_host with pinned host "
          "buffers. The roofline peak is MEASURED_PEAKS.json's hbm_gbs. Full lines: `%s_bench_lines.jsonl`." % tag)
open(os.path.join(ROOT, "profiles", "%s_bench.md" % tag), "w").write("\n".join(md) + "\n")
md, all_sc = ["# Scaling, round %s (gpurun --gpus N, one box; tools/scale_round.sh; torchrun, NCCL for barrier + reductions only)\n" % tag], []
for mode, what in (("weak", "every rank decodes its own 100,000-stream shard (the job is N x 100,000 streams)"),
                   ("strong", "one 100,000-stream batch split over the ranks")):
    sc = [j for j in (last_json(f) for f in sorted(glob.glob(os.path.join(out, "scale_%s_n*.json" % mode)))) if j]
    if not sc:
        continue
    sc.sort(key=lambda j: j["n_gpus"])
    all_sc += sc
    base = sc[0]["value"] / sc[0]["n_gpus"]
    md += ["## %s scaling: %s\n" % (mode, what),
           "| N | value GB/s | ms/step | e2e GB/s | streams/rank | dominant kernel | value / (N=1 value) |", "|---|---|---|---|---|---|---|"]
    for j in sc:
        md.append("| %d | %.1f | %.3f | %s | %d | %s | %.2f |" % (j["n_gpus"], j["value"], j["ms_per_step"], ("%.1f" % j["e2e"]["value"]) if j.get("e2e") else "-",
                                                                 j["config"]["streams_per_rank"], j["roofline"]["kernel"], j["value"] / base))
    md.append("")
for f in sorted(glob.glob(os.path.join(out, "scale_ref_n*.json"))):
    j = last_json(f)
    if j:
        md.append("reference arm at N=%d (rank 0 only, oracle port on %d host threads): %.2f GB/s\n" % (j["n_gpus"], j["cpu_baseline"]["cores"], j["value"]))
if all_sc:
    open(os.path.join(ROOT, "profiles", "%s_scaling.md" % tag), "w").write("\n".join(md) + "\n")
    open(os.path.join(ROOT, "profiles", "%s_scaling_lines.jsonl" % tag), "w").write("\n".join(json.dumps(j) for j in all_sc) + "\n")
print(open(os.path.join(ROOT, "profiles", "%s_bench.md" % tag)).read())

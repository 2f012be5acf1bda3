#!/usr/bin/env python3
"""Kernel-only timing of every BASELINE workload for one or more builds of the library (tuning helper).
usage: python tools/quick_perf.py [lib.so ...]     (default: the product build)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sys.argv[1:] or [""]
WL = os.environ.get("BRO_WORKLOADS", "c4_highratio_w16,c2_quickfox_x10k,c3_corpus_x1000,c5_stored_10k,c5b_literals_10k").split(",")
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["BRO_B200_LIB"] = os.path.join(ROOT, "brotli_rs_b200", "lib", lib)
    for w in WL:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", w, "--steps", "5", "--warmup", "2",
                            "--no-e2e", "--no-cpu-baseline"] + (["--mode", os.environ["BRO_BENCH_MODE"]] if os.environ.get("BRO_BENCH_MODE") else []),
                           capture_output=True, text=True, env=env)
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            ks = "  ".join("%s %.3f ms%s" % (k.split("_kernel")[0].replace("bro_", ""), v["ms"], (" (%.2f)" % v["frac"]) if v["frac"] else "")
                           for k, v in j["roofline"]["kernels"].items())
            print("%-22s %-20s %9.1f GB/s  %8.3f ms  | %s | retried %d" % (lib or "product", w, j["value"], j["ms_per_step"], ks,
                                                                         j["roofline"]["batch_stats"]["retried_streams"]), flush=True)
        except Exception:
            print(lib, w, "FAILED", r.stderr[-400:], flush=True)

#!/usr/bin/env python3
"""bench.py -- the headline measurement (BASELINE.json): uncompressed GB/s of a many-stream Brotli decode batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4_highratio_w16] [--streams 100000]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)
    python bench.py --impl reference ...                                              (the CPU arm)

A "step" is one pass of the hot path (one bro_batch_decode call per rank: ordering, parse, copy and fused-retry
kernels for a large batch, the fused kernel alone for a small one) over the whole batch.  Streams are independent: the
job is sharded by stream with no data-path collective.  --scaling weak (default, what the contract asks of a path that
partitions): every rank decodes its own BASELINE-size batch (100,000 streams per GPU; the job is N such shards).
--scaling strong: one BASELINE-size batch split over the N ranks (profiles/r01b_scaling.md has both).

  value   whole-job uncompressed GB/s with inputs and outputs resident in HBM (CUDA events, max over ranks)
  e2e     the same metric through the reference-facing C ABI call bro_batch_decode_host with pinned HOST buffers:
          the H2D copy of the compressed batch and the D2H copy of every output slot are inside the timed region
  roofline  the dominant kernel of the step: its algorithmic bytes per launch / its mean duration, both measured live
          (CUDA events recorded by the library around each kernel); the other kernels are listed under "kernels"
  cpu_baseline  the oracle port of the reference algorithm on the host cores, on a bounded sample (N = 1 only)

Only the cpu_baseline / --impl reference legs execute anything under oracle/; the GPU path never does.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_UNIQUE = 1024          # distinct synthetic streams, tiled to the batch size


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4_highratio_w16")
    ap.add_argument("--streams", type=int, default=None, help="batch size (default: the workload's BASELINE size)")
    ap.add_argument("--e2e-steps", type=int, default=None, help="timed end-to-end steps (default: min(steps, 5))")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --streams per rank (the job is N shards of that size); strong: --streams split over the ranks")
    ap.add_argument("--mode", default="auto", choices=["auto", "warp", "twophase"],
                    help="decode path (bro_ctx_set_mode); auto = the library's default policy")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-streams", type=int, default=None)
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload, algorithmic_bytes, kernel):
    """dram bytes per launch of `kernel` from the committed ncu capture of this workload, scaled to this launch's size."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        rec = json.load(open(p)).get(workload)
        if not rec:
            return None
        if "kernels" in rec:
            rec = rec["kernels"].get(kernel.split(" ")[0])
            if not rec:
                return None
        elif not kernel.startswith("bro_decode_warp_kernel"):
            return None
        return rec["dram_bytes_per_algorithmic_byte"] * algorithmic_bytes
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].strip().lower() == "active":
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


DEFAULT_STREAMS = {"c4_highratio_w16": 100000, "c5_stored_10k": 100000, "c5b_literals_10k": 100000,
                   "c2_quickfox_x10k": 10000, "c3_corpus_x1000": 52000, "c1_alice29_single": 1}


def build_workload(name, n_streams):
    """-> dict(streams, raws, status, gidx, desc): the distinct streams, their expected bytes (None = invalid stream),
    expected status, and the tiling of the distinct streams into the batch."""
    from brotli_rs_b200 import workloads as w
    data = os.path.join(ROOT, "tests", "golden", "data")
    if name == "c1_alice29_single":
        # BASELINE configs[0]: one stream, the reference's own CPU-runnable case (plumbing and single-stream latency)
        _, streams, raws, status = w.corpus_workload(data, only={"alice29.txt.compressed"})
        desc = "data/alice29.txt.compressed, one stream (50,096 B -> 152,089 B)"
        gidx = np.zeros(n_streams, dtype=np.int64)
    elif name == "c2_quickfox_x10k":
        _, streams, raws, status = w.corpus_workload(data, only={"quickfox_repeated.compressed"})
        desc = "copies of data/quickfox_repeated.compressed (58 B -> 176,128 B, one overlapping copy at distance 43)"
        gidx = np.zeros(n_streams, dtype=np.int64)
    elif name == "c3_corpus_x1000":
        _, streams, raws, status = w.corpus_workload(data)
        desc = "every data/*compressed* stream (52, 9 invalid) x replicas, shuffled with default_rng(0)"
        gidx = np.tile(np.arange(len(streams)), (n_streams + len(streams) - 1) // len(streams))[:n_streams]
        np.random.default_rng(0).shuffle(gidx)
    else:
        streams, raws = w.make_unique_streams(name, N_UNIQUE)
        status = [0] * len(streams)
        desc = w.WORKLOADS[name][5]
        gidx = np.arange(n_streams) % N_UNIQUE
    return {"streams": streams, "raws": raws, "status": np.array(status, dtype=np.int32), "gidx": gidx, "desc": desc}


def cpu_baseline_run(wl, nthreads, sample_streams):
    """Oracle port on the host cores over the first `sample_streams` streams of the batch."""
    from brotli_rs_b200.batch import pack_streams, slot_offsets
    from oracle import oracle
    idx = wl["gidx"][:sample_streams]
    sel = [wl["streams"][i] for i in idx]
    in_buf, in_off = pack_streams(sel)
    out_off = slot_offsets([len(wl["raws"][i]) if wl["raws"][i] is not None else 70000 for i in idx])
    out = np.empty(int(out_off[-1]), dtype=np.uint8)
    t0 = time.perf_counter()
    _, out_len, status = oracle.decode_batch(in_buf, in_off, out_off, nthreads=nthreads, out=out)
    dt = time.perf_counter() - t0
    assert (status == wl["status"][idx]).all()
    return float(out_len[status == 0].sum()), dt


def cpu_sample_size(wl, cores, seconds, n_streams):
    per_stream = float(np.mean([len(r) for r in wl["raws"] if r is not None]))
    return int(max(cores, min(n_streams, 16384, seconds * 60e6 * cores / per_stream)))


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port of brotli-rs; the Rust crate cannot
    be built in this image) with all host threads, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_streams = (args.streams or DEFAULT_STREAMS[args.workload]) * (world if args.scaling == "weak" else 1)
    cores = min(os.cpu_count() or 1, n_streams)          # a stream is decoded by one thread
    wl = build_workload(args.workload, n_streams)
    desc = wl["desc"]
    sample = args.cpu_sample_streams or cpu_sample_size(wl, cores, 1.5, n_streams)
    cpu_baseline_run(wl, cores, sample)
    total_b, total_t = 0.0, 0.0
    for _ in range(args.steps):
        b, t = cpu_baseline_run(wl, cores, sample)
        total_b += b; total_t += t
    val = total_b / total_t / 1e9
    line = {
        "impl": "reference", "metric": "uncompressed GB/s (many-stream batch)", "value": val, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "streams": n_streams,
                   "sample": "%d streams per step" % sample},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": "%d of %d streams per step, oracle/brotli_oracle.c (C restatement of brotli-rs 0.3.23), %d threads"
                                   % (sample, n_streams, cores)},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from brotli_rs_b200 import BatchDecoder, shard_streams
    from brotli_rs_b200.batch import slot_offsets

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU decode path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- workload: N_UNIQUE distinct streams tiled to the job's batch; this rank's shard ----
    per_rank = args.streams or DEFAULT_STREAMS[args.workload]
    n_streams = per_rank * world if args.scaling == "weak" else per_rank
    wl = build_workload(args.workload, n_streams)
    streams, raws, desc, gidx = wl["streams"], wl["raws"], wl["desc"], wl["gidx"]
    ulen_in = np.array([len(s) for s in streams], dtype=np.int64)
    ulen_out = np.array([len(r) if r is not None else 0 for r in raws], dtype=np.int64)
    ucap = np.array([len(r) if r is not None else 70000 for r in raws], dtype=np.int64)
    if args.scaling == "weak":
        mine = np.arange(rank * per_rank, (rank + 1) * per_rank)      # every rank: one BASELINE-size shard of the job
    else:
        mine = shard_streams(ulen_in[gidx], ucap[gidx], world, rank)
    uidx = gidx[mine]
    n = len(uidx)
    in_off = np.concatenate([[0], np.cumsum(ulen_in[uidx])]).astype(np.int64)
    out_off = slot_offsets(ucap[uidx]).astype(np.int64)
    ubuf = torch.from_numpy(np.frombuffer(b"".join(streams), dtype=np.uint8).copy()).to(dev)
    uoff = np.concatenate([[0], np.cumsum(ulen_in)]).astype(np.int64)
    # gather the tiled compressed batch on the device
    delta = torch.from_numpy(uoff[uidx] - in_off[:-1]).to(dev)
    src_idx = torch.repeat_interleave(delta, torch.from_numpy(ulen_in[uidx]).to(dev)) + torch.arange(int(in_off[-1]), device=dev)
    d_in = ubuf[src_idx].contiguous()
    del src_idx, delta
    d_in_off = torch.from_numpy(in_off).to(dev)
    d_out_off = torch.from_numpy(out_off).to(dev)
    d_out = torch.empty(int(out_off[-1]), dtype=torch.uint8, device=dev)
    d_len = torch.empty(n, dtype=torch.int64, device=dev)
    d_st = torch.empty(n, dtype=torch.int32, device=dev)
    dec = BatchDecoder(local_rank, mode={"auto": None, "warp": BatchDecoder.MODE_WARP, "twophase": BatchDecoder.MODE_TWOPHASE}[args.mode])

    comp_bytes = float(in_off[-1])
    uncomp_bytes = float(ulen_out[uidx].sum())

    # ---- parity gate (a timing is only reported if the results are right) ----
    dec.decode_device(d_in, d_in_off, d_out, d_out_off, d_len, d_st)
    torch.cuda.synchronize()
    want_st = torch.from_numpy(wl["status"][uidx]).to(dev)
    assert bool((d_st == want_st).all()), "status mismatch for some stream"
    ok_mask = want_st == 0
    assert float(d_len[ok_mask].sum().item()) == uncomp_bytes
    d_raw = torch.from_numpy(np.frombuffer(b"".join(r for r in raws if r is not None), dtype=np.uint8).copy()).to(dev)
    roff = np.concatenate([[0], np.cumsum(ulen_out)]).astype(np.int64)
    check = np.unique(np.concatenate([np.arange(min(n, 2048)), np.random.default_rng(rank).integers(0, n, 2048)]))
    for k in check:
        u = int(uidx[k])
        if raws[u] is None:
            continue
        a = d_out[int(out_off[k]): int(out_off[k]) + int(ulen_out[u])]
        assert torch.equal(a, d_raw[int(roff[u]): int(roff[u + 1])]), "GPU output differs from the expected bytes (stream %d)" % k
    del d_raw

    # ---- timed region: K launches, CUDA events on the launching stream ----
    for _ in range(args.warmup):
        dec.decode_device(d_in, d_in_off, d_out, d_out_off, d_len, d_st)
    sampler = ClockSampler(local_rank)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    sampler.start()
    launches0 = dec.launch_count
    ev[0].record()
    for k in range(args.steps):
        dec.decode_device(d_in, d_in_off, d_out, d_out_off, d_len, d_st)
        ev[k + 1].record()
    barrier()
    clocks = sampler.stop()
    launches = dec.launch_count - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    assert bool((d_st == want_st).all())
    total_ms_max = max_over_ranks(total_ms)
    all_uncomp = sum_over_ranks(uncomp_bytes)
    all_comp = sum_over_ranks(comp_bytes)
    value = all_uncomp * args.steps / (total_ms_max * 1e-3) / 1e9

    # ---- per-kernel times (CUDA events recorded by the library around its kernels, on the launching stream) and the
    # roofline of the dominant kernel on this rank ----
    peak, peak_src = measured_peaks()
    dec.set_timing(True)
    ksum = {"order": 0.0, "parse": 0.0, "copy": 0.0, "fused": 0.0}
    for _ in range(args.steps):
        dec.decode_device(d_in, d_in_off, d_out, d_out_off, d_len, d_st)
        for k, v in dec.last_kernel_ms().items():
            ksum[k] += v
    dec.set_timing(False)
    stats = dec.last_batch_stats()
    kms = {k: v / args.steps for k, v in ksum.items()}
    # algorithmic bytes per launch: fused = compressed read + uncompressed written; copy = bytes
    # written by copy records + 16 B per record read; parse = compressed read + bytes it writes itself (literals,
    # dictionary words) + 16 B per record written
    two_phase = kms["parse"] > 0.0 and not stats.get("gated_to_fused")
    rec_bytes = 16.0 * stats["copy_records"]
    if two_phase:
        retried = stats["retried_streams"]
        algo = {"copy": stats["copy_bytes"] + rec_bytes,
                "parse": comp_bytes + max(0.0, uncomp_bytes - stats["copy_bytes"]) + rec_bytes if retried == 0 else comp_bytes + rec_bytes,
                "fused": (comp_bytes + uncomp_bytes) if retried == n else None, "order": 8.0 * (n + 1) * 2}
    else:
        algo = {"fused": comp_bytes + uncomp_bytes}
        kms = {k: (v if k == "fused" or v > 0.05 else 0.0) for k, v in kms.items()}   # gated batch: the other kernels returned at once
    names = {"order": "bro_order_*_kernel (3)", "parse": "bro_parse_kernel", "copy": "bro_copy_kernel", "fused": "bro_decode_warp_kernel"}
    kernels = {}
    for k, ms in kms.items():
        if ms <= 0.0:
            continue
        a = algo.get(k)
        kernels[names[k]] = {"ms": ms, "algorithmic_bytes": a, "achieved_gbs": (a / (ms * 1e-3) / 1e9) if a else None,
                             "frac": (a / (ms * 1e-3) / 1e9 / peak) if a else None}
    dom = max((k for k in kms if kms[k] > 0.0), key=lambda k: kms[k])
    kern_ms = kms[dom]
    algo_bytes = algo.get(dom) or (comp_bytes + uncomp_bytes)
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(args.workload, algo_bytes, names[dom]), "peak_source": peak_src,
                "kernel": names[dom], "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": algo_bytes,
                "kernel_share_of_step": kern_ms / float(np.mean(step_ms)), "kernels": kernels,
                "whole_step": {"ms": float(np.mean(step_ms)), "algorithmic_bytes": comp_bytes + uncomp_bytes,
                               "frac": (comp_bytes + uncomp_bytes) / (float(np.mean(step_ms)) * 1e-3) / 1e9 / peak},
                "batch_stats": stats}
    if dom == "parse":
        roofline["note"] = ("bro_parse_kernel (entropy decode, one thread per stream) moves few bytes and is latency-bound, not "
                            "HBM-bound (DESIGN.md 3.3); the HBM-bound kernel of the step is bro_copy_kernel, see kernels")

    # ---- end to end through the C ABI with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.e2e_steps or min(args.steps, 5)
        # Host memory for the end-to-end leg is bounded per BOX: with N ranks each pins 1/N of a BASELINE-size batch (at
        # weak scaling the first 1/N of its shard), so that N = 8 does not pin 8 x 26 GB.
        ne = n if (world == 1 or args.scaling == "strong") else max(1, n // world)
        e_in, e_out = int(in_off[ne]), int(out_off[ne])
        h_in = torch.empty(e_in, dtype=torch.uint8).pin_memory()
        h_in.copy_(d_in[:e_in])
        h_out = torch.empty(e_out, dtype=torch.uint8).pin_memory()
        hin_np, hout_np = h_in.numpy(), h_out.numpy()
        in_off_u, out_off_u = in_off[: ne + 1].astype(np.uint64), out_off[: ne + 1].astype(np.uint64)
        dec.decode_host(hin_np, in_off_u, out_off_u, out=hout_np)        # warm-up (also sizes the device staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            _, out_len, status = dec.decode_host(hin_np, in_off_u, out_off_u, out=hout_np)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e_uncomp = float(ulen_out[uidx[:ne]].sum())
        assert (status == wl["status"][uidx[:ne]]).all() and float(out_len[status == 0].sum()) == e_uncomp
        for k in check[check < ne][-8:]:
            u = int(uidx[k])
            if raws[u] is not None:
                assert hout_np[int(out_off[k]): int(out_off[k]) + int(ulen_out[u])].tobytes() == raws[u]
        e2e_t = max_over_ranks(t1 - t0)
        e2e = {"value": sum_over_ranks(e_uncomp) * e2e_steps / e2e_t / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": int(e_in + 2 * 8 * (ne + 1)), "d2h_bytes_per_step": int(e_out + 12 * ne),
               "steps": e2e_steps, "streams_per_rank": int(ne), "api": "bro_batch_decode_host (pinned host buffers, per rank)"}
        del h_in, h_out

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores, bounded sample ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = min(os.cpu_count() or 1, n_streams)      # a stream is decoded by one thread
        sample = args.cpu_sample_streams or cpu_sample_size(wl, cores, 12.0, n_streams)
        b, t = cpu_baseline_run(wl, cores, sample)
        cpu = {"value": b / t / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
               "sample": "%d of %d streams, oracle/brotli_oracle.c (C restatement of brotli-rs 0.3.23; rustc absent), %d threads, %.1f s"
                         % (sample, n_streams, cores, t)}

    if rank == 0:
        line = {
            "metric": "uncompressed GB/s (many-stream batch)", "value": value, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "streams": n_streams,
                       "unique_streams": len(streams), "streams_per_rank": n, "mode": args.mode,
                       "compressed_bytes": all_comp, "uncompressed_bytes": all_uncomp,
                       "l2_policy": "inputs and outputs far larger than L2 (no flush needed)",
                       "parallelism": ("%d rank(s), one %d-stream shard each (weak scaling), no data-path collective" % (world, n))
                                      if args.scaling == "weak" else
                                      ("one batch split over %d rank(s) (strong scaling), no data-path collective" % world)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

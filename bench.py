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
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling record")
    ap.add_argument("--no-exchange", action="store_true", help="N > 1: skip the scatter / gather measurement")
    ap.add_argument("--no-write-roof", action="store_true", help="skip the fill (write-only bandwidth) measurement")
    ap.add_argument("--extra-workloads", action=argparse.BooleanOptionalAction, default=None,
                    help="also run the other BASELINE configurations, a few steps each (default: on for the default run)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-streams", type=int, default=None)
    args = ap.parse_args()
    if args.extra_workloads is None:
        # the default run (what the driver launches) carries the other configurations; tuning runs ask for one workload
        args.extra_workloads = args.workload == "c4_highratio_w16" and not args.no_e2e and args.impl == "b200"
    return args


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
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  The sampler is started before
    the warm-up (nvidia-smi takes a few hundred ms to come up, longer with a process per rank on an 8-GPU box) and every row
    carries its timestamp; the rows that fall inside the timed window are the ones reported.  A timed region shorter than the
    sampling interval falls back to the rows of the warm-up that ran right before it under the same load, and says so."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = self.t1 = None

    def start(self):
        if self.p is not None:
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "25"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def has_rows(self):
        try:
            return os.path.getsize(self.f.name) > 0
        except OSError:
            return False

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    @staticmethod
    def _when(text):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except Exception:
            return None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        parsed = []
        for r in rows:
            try:
                parsed.append((self._when(r[0]), float(r[1]), float(r[2]),
                               [nm for k, nm in enumerate(self.NAMES) if len(r) > 5 + k and r[5 + k].strip().lower() == "active"]))
            except Exception:
                continue
        inside = [p for p in parsed if p[0] is not None and self.t0 is not None and self.t0 <= p[0] <= self.t1]
        window = "timed region"
        if not inside:
            inside, window = parsed, "warm-up + timed region (no nvidia-smi row fell inside the timed region itself)"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median([p[1] for p in inside])), "sm_max_mhz": float(max(p[2] for p in inside)),
                "reasons": sorted({nm for p in inside for nm in p[3]}), "samples": len(inside), "window": window}


def combine_clocks(h, clocks):
    """every rank sampled its own GPU: the line reports the slowest rank's median clock and the union of the throttle reasons"""
    if h.world == 1:
        return clocks
    every = [None] * h.world
    h.dist.all_gather_object(every, clocks)
    have = [c for c in every if c.get("sm_mhz") is not None]
    if not have:
        return every[0]
    out = {"sm_mhz": min(c["sm_mhz"] for c in have), "sm_max_mhz": max(c["sm_max_mhz"] for c in have),
           "reasons": sorted({r for c in every for r in c["reasons"]}), "samples": sum(c["samples"] for c in have),
           "window": have[0]["window"] if all(c["window"] == have[0]["window"] for c in have) else "mixed",
           "ranks": "min of the ranks' median SM clocks; reasons of all %d ranks" % h.world}
    return out


DEFAULT_STREAMS = {"c6_text_q11_w16": 20000, "c7_far_w22": 8000, "c4_highratio_w16": 100000, "c5_stored_10k": 100000, "c5b_literals_10k": 100000,
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
    elif name == "c6_text_q11_w16":
        # round 2 (VERDICT "decide with data"): alice29.txt at quality 11, lgwin 16 -- context-modelled, 11 % of the output from
        # the static dictionary: the parse kernel's immediate mode (bro_parse.h), or the fused kernel's general loop (--mode warp)
        raw = open(os.path.join(data, "alice29.txt"), "rb").read()
        enc = w.libbrotli_enc()
        streams, raws, status = [w.compress(enc, raw, 11, 16)], [raw], [0]
        desc = "copies of alice29.txt at quality 11 / lgwin 16 (%d -> %d B; literal context modelling, static dictionary)" % (len(streams[0]), len(raw))
        gidx = np.zeros(n_streams, dtype=np.int64)
    elif name == "c3_corpus_x1000":
        _, streams, raws, status = w.corpus_workload(data)
        desc = "every data/*compressed* stream (52, 9 invalid) x replicas, shuffled with default_rng(0)"
        gidx = np.tile(np.arange(len(streams)), (n_streams + len(streams) - 1) // len(streams))[:n_streams]
        np.random.default_rng(0).shuffle(gidx)
    else:
        streams, raws = w.make_unique_streams(name, N_UNIQUE if name != "c7_far_w22" else 64)
        status = [0] * len(streams)
        desc = w.WORKLOADS[name][5]
        gidx = np.arange(n_streams) % len(streams)
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


def config_record(args, desc, n_streams, n_unique, per_rank, comp, uncomp, world):
    """the same keys on both arms (the driver compares them)"""
    return {"workload": args.workload, "description": desc, "streams": n_streams, "unique_streams": n_unique,
            "streams_per_rank": per_rank, "mode": args.mode, "compressed_bytes": comp, "uncompressed_bytes": uncomp,
            "l2_policy": "inputs and outputs far larger than L2 (no flush needed)",
            "parallelism": ("%d rank(s), one %d-stream shard each (weak scaling), no data-path collective" % (world, per_rank))
                           if args.scaling == "weak" else
                           ("one batch split over %d rank(s) (strong scaling), no data-path collective" % world)}


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
        "config": config_record(args, desc, n_streams, len(wl["streams"]), n_streams // world if args.scaling == "weak" else n_streams,
                                float(sum(len(wl["streams"][i]) for i in wl["gidx"])),
                                float(sum(len(wl["raws"][i]) for i in wl["gidx"] if wl["raws"][i] is not None)), world),
        "sample": "%d of %d streams per step (a thread pool's throughput does not depend on the batch size)" % (sample, n_streams),
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": "%d of %d streams per step, oracle/brotli_oracle.c (C restatement of brotli-rs 0.3.23), %d threads"
                                   % (sample, n_streams, cores)},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


class Harness:
    """One rank's device, distributed helpers and decoder."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from brotli_rs_b200 import BatchDecoder
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU decode path)"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.dec = BatchDecoder(self.local_rank, mode={"auto": None, "warp": BatchDecoder.MODE_WARP,
                                                       "twophase": BatchDecoder.MODE_TWOPHASE}[args.mode])

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, x):
        return self.reduce(x, self.dist.ReduceOp.MAX if self.world > 1 else None)

    def sum_over_ranks(self, x):
        return self.reduce(x, self.dist.ReduceOp.SUM if self.world > 1 else None)


def measure_fill_gbs(h, nbytes=8 << 30):
    """Write-only bandwidth of this GPU, measured the way MEASURED_PEAKS.json measures the copy peak (torch, CUDA events,
    best of 5): the roof of a kernel that mostly WRITES, which a read+write copy figure overstates by its read half."""
    torch = h.torch
    buf = torch.empty(nbytes, dtype=torch.uint8, device=h.dev)
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); buf.fill_(7); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    del buf
    return nbytes / best / 1e6


class Batch:
    """A workload's batch (or this rank's shard of it) resident on the device, with what is needed to check it."""

    def __init__(self, h, wl, mine):
        from brotli_rs_b200.batch import slot_offsets
        torch = h.torch
        streams, raws, gidx = wl["streams"], wl["raws"], wl["gidx"]
        self.wl, self.h = wl, h
        self.ulen_in = np.array([len(s) for s in streams], dtype=np.int64)
        self.ulen_out = np.array([len(r) if r is not None else 0 for r in raws], dtype=np.int64)
        self.ucap = np.array([len(r) if r is not None else 70000 for r in raws], dtype=np.int64)
        self.uidx = gidx[mine]
        self.n = len(self.uidx)
        self.in_off = np.concatenate([[0], np.cumsum(self.ulen_in[self.uidx])]).astype(np.int64)
        self.out_off = slot_offsets(self.ucap[self.uidx]).astype(np.int64)
        ubuf = torch.from_numpy(np.frombuffer(b"".join(streams), dtype=np.uint8).copy()).to(h.dev)
        uoff = np.concatenate([[0], np.cumsum(self.ulen_in)]).astype(np.int64)
        # gather the tiled compressed batch on the device (+ 16 readable bytes behind the last stream: include/brotli_b200.h)
        delta = torch.from_numpy(uoff[self.uidx] - self.in_off[:-1]).to(h.dev)
        src_idx = torch.repeat_interleave(delta, torch.from_numpy(self.ulen_in[self.uidx]).to(h.dev)) + torch.arange(int(self.in_off[-1]), device=h.dev)
        self.d_in = torch.zeros(int(self.in_off[-1]) + 16, dtype=torch.uint8, device=h.dev)
        self.d_in[: int(self.in_off[-1])] = ubuf[src_idx]
        del src_idx, delta, ubuf
        self.d_in_off = torch.from_numpy(self.in_off).to(h.dev)
        self.d_out_off = torch.from_numpy(self.out_off).to(h.dev)
        self.d_out = torch.empty(int(self.out_off[-1]), dtype=torch.uint8, device=h.dev)
        self.d_len = torch.empty(self.n, dtype=torch.int64, device=h.dev)
        self.d_st = torch.empty(self.n, dtype=torch.int32, device=h.dev)
        self.comp_bytes = float(self.in_off[-1])
        self.uncomp_bytes = float(self.ulen_out[self.uidx].sum())
        self.want_st = torch.from_numpy(wl["status"][self.uidx]).to(h.dev)
        self.check = np.unique(np.concatenate([np.arange(min(self.n, 2048)), np.random.default_rng(h.rank).integers(0, self.n, 2048)]))

    def decode(self):
        self.h.dec.decode_device(self.d_in, self.d_in_off, self.d_out, self.d_out_off, self.d_len, self.d_st)

    def parity_gate(self):
        """a timing is only reported if the results are right: statuses, lengths and ~4,000 sampled slots"""
        torch, raws = self.h.torch, self.wl["raws"]
        self.decode()
        torch.cuda.synchronize()
        assert bool((self.d_st == self.want_st).all()), "status mismatch for some stream"
        ok_mask = self.want_st == 0
        assert float(self.d_len[ok_mask].sum().item()) == self.uncomp_bytes
        d_raw = torch.from_numpy(np.frombuffer(b"".join(r for r in raws if r is not None), dtype=np.uint8).copy()).to(self.h.dev)
        roff = np.concatenate([[0], np.cumsum(self.ulen_out)]).astype(np.int64)
        for k in self.check:
            u = int(self.uidx[k])
            if raws[u] is None:
                continue
            a = self.d_out[int(self.out_off[k]): int(self.out_off[k]) + int(self.ulen_out[u])]
            assert torch.equal(a, d_raw[int(roff[u]): int(roff[u + 1])]), "GPU output differs from the expected bytes (stream %d)" % k

    def timed(self, steps, warmup, sampler=None, collective=True):
        """-> (total ms of `steps` launches on this rank, per-step ms, launches): CUDA events on the launching stream.
        collective=False: a measurement of this rank alone (no barrier: the other ranks are not in it)"""
        h, torch = self.h, self.h.torch
        sync = h.barrier if collective else torch.cuda.synchronize
        if sampler:
            sampler.start()
        for _ in range(warmup):
            self.decode()
        if sampler and sampler.p is not None:
            # keep the load on (untimed) until nvidia-smi has delivered its first row, so that the timed region is sampled
            for _ in range(400):
                torch.cuda.synchronize()
                if sampler.has_rows():
                    break
                self.decode()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        sync()
        if sampler:
            sampler.mark_begin()
        launches0 = h.dec.launch_count
        ev[0].record()
        for k in range(steps):
            self.decode()
            ev[k + 1].record()
        sync()
        if sampler:
            sampler.mark_end()
        assert bool((self.d_st == self.want_st).all())
        return ev[0].elapsed_time(ev[-1]), [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)], h.dec.launch_count - launches0

    def kernel_times(self, steps):
        """per-kernel times (CUDA events recorded by the library around its kernels) and the batch's counters"""
        dec = self.h.dec
        dec.set_timing(True)
        ksum = {"order": 0.0, "parse": 0.0, "copy": 0.0, "fused": 0.0}
        for _ in range(steps):
            self.decode()
            for k, v in dec.last_kernel_ms().items():
                ksum[k] += v
        dec.set_timing(False)
        return {k: v / steps for k, v in ksum.items()}, dec.last_batch_stats()


def quick_workload(h, name, steps, warmup):
    """one of the other BASELINE configurations on this rank's GPU, device-resident: -> a small record"""
    n_streams = DEFAULT_STREAMS[name]
    wl = build_workload(name, n_streams)
    b = Batch(h, wl, np.arange(n_streams))
    b.parity_gate()
    total_ms, step_ms, _ = b.timed(steps, warmup, collective=False)      # rank 0 alone runs the side measurements
    kms, stats = b.kernel_times(min(steps, 3))
    peak, _ = measured_peaks()
    ms = total_ms / steps
    rec = {"workload": name, "streams": n_streams, "ms_per_step": ms, "value": b.uncomp_bytes / (ms * 1e-3) / 1e9, "unit": "GB/s",
           "whole_step_frac": (b.comp_bytes + b.uncomp_bytes) / (ms * 1e-3) / 1e9 / peak,
           "kernel_ms": {k: v for k, v in kms.items() if v > 0.004}, "retried_streams": stats["retried_streams"],
           "gated_to_fused": stats["gated_to_fused"], "parity": "statuses, lengths and sampled slots checked"}
    del b
    h.torch.cuda.empty_cache()
    return rec


def pcie_ceiling(h, h_out, d_out, h_in, d_in):
    """bare pinned cudaMemcpyAsync of the step's bytes (H2D and D2H), EVERY rank at once: each of the 3 trials starts at a
    barrier and counts as its slowest rank's time (the ranks share the host's memory and PCIe roots); best trial"""
    torch = h.torch
    best = 1e9
    for _ in range(3):
        h.barrier()
        t0 = time.perf_counter()
        d_in.copy_(h_in, non_blocking=True)
        h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, h.max_over_ranks(time.perf_counter() - t0))
    return best


def bind_to_gpu_numa_node(torch, local_rank):
    """Run this rank's host threads (and so first-touch its pinned staging) on the NUMA node its GPU hangs off: with one
    process per GPU on a two-socket box the copies otherwise cross the socket link.  -> a note for the bench line."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return "numa node of %s unknown: not bound" % bdf
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "no allowed cpu on node %d: not bound" % node
        os.sched_setaffinity(0, cpus)
        return "host threads bound to NUMA node %d (%d cpus) of GPU %s" % (node, len(cpus), bdf)
    except Exception as e:      # sysfs layout differs, no permission ...: measured unbound
        return "not bound (%s)" % type(e).__name__


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    from brotli_rs_b200 import shard_streams
    h = Harness(args)
    torch, dist, rank, world, dev, dec = h.torch, h.dist, h.rank, h.world, h.dev, h.dec

    # ---- workload: N_UNIQUE distinct streams tiled to the job's batch; this rank's shard ----
    per_rank = args.streams or DEFAULT_STREAMS[args.workload]
    n_streams = per_rank * world if args.scaling == "weak" else per_rank
    wl = build_workload(args.workload, n_streams)
    streams, raws, desc, gidx = wl["streams"], wl["raws"], wl["desc"], wl["gidx"]
    ulen_in = np.array([len(s) for s in streams], dtype=np.int64)
    ucap = np.array([len(r) if r is not None else 70000 for r in raws], dtype=np.int64)
    if args.scaling == "weak":
        mine = np.arange(rank * per_rank, (rank + 1) * per_rank)      # every rank: one BASELINE-size shard of the job
    else:
        mine = shard_streams(ulen_in[gidx], ucap[gidx], world, rank)
    bt = Batch(h, wl, mine)
    n, uidx, in_off, out_off, ulen_out = bt.n, bt.uidx, bt.in_off, bt.out_off, bt.ulen_out
    comp_bytes, uncomp_bytes = bt.comp_bytes, bt.uncomp_bytes
    bt.parity_gate()

    # ---- timed region: K launches, CUDA events on the launching stream ----
    sampler = ClockSampler(h.local_rank)
    total_ms, step_ms, launches = bt.timed(args.steps, args.warmup, sampler)
    clocks = combine_clocks(h, sampler.stop())
    total_ms_max = h.max_over_ranks(total_ms)
    all_uncomp = h.sum_over_ranks(uncomp_bytes)
    all_comp = h.sum_over_ranks(comp_bytes)
    value = all_uncomp * args.steps / (total_ms_max * 1e-3) / 1e9

    # ---- per-kernel times and the roofline of the dominant kernel on this rank ----
    peak, peak_src = measured_peaks()
    kms, stats = bt.kernel_times(args.steps)
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
    mean_step = float(np.mean(step_ms))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(args.workload, algo_bytes, names[dom]), "peak_source": peak_src,
                "kernel": names[dom], "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": algo_bytes,
                "kernel_share_of_step": kern_ms / mean_step, "kernels": kernels,
                "whole_step": {"ms": mean_step, "algorithmic_bytes": comp_bytes + uncomp_bytes,
                               "frac": (comp_bytes + uncomp_bytes) / (mean_step * 1e-3) / 1e9 / peak},
                "batch_stats": stats}
    if rank == 0 and not args.no_write_roof:
        # The step WRITES 67 x what it reads from DRAM (26 GB of output for 0.4 GB of input; copy sources are re-reads of
        # recent output, mostly L2 hits).  The read+write copy figure above is the contract's denominator; the roof that
        # physically bounds a write-dominated kernel is the write-only bandwidth, measured here the same way.
        fill = measure_fill_gbs(h)
        roofline["write_roof"] = {"fill_gbs": fill, "how": "torch fill_ of 8 GiB, CUDA events, best of 5, this run",
                                  "whole_step_frac_of_write_roof": uncomp_bytes / (mean_step * 1e-3) / 1e9 / fill}
        if two_phase and kms["copy"] > 0.0:
            roofline["write_roof"]["copy_kernel_frac_of_write_roof"] = stats["copy_bytes"] / (kms["copy"] * 1e-3) / 1e9 / fill
            # SURVEY.md 8(d) counts the copy phase as 2 x copy bytes (source read + destination write), which is what a
            # plain device copy of the same bytes moves through the SMs
            kernels[names["copy"]]["frac_counting_source_reads"] = (2.0 * stats["copy_bytes"] + rec_bytes) / (kms["copy"] * 1e-3) / 1e9 / peak
    if dom == "parse":
        roofline["note"] = ("bro_parse_kernel (entropy decode, one thread per stream) moves few bytes and is latency-bound, not "
                            "HBM-bound (DESIGN.md 3.3); the HBM-bound kernel of the step is bro_copy_kernel, see kernels")

    # ---- strong scaling (N > 1): ONE BASELINE-size batch split over the ranks, next to the weak headline above ----
    strong = None
    if world > 1 and args.scaling == "weak" and not args.no_strong:
        wl1 = build_workload(args.workload, per_rank)
        mine_s = shard_streams(ulen_in[wl1["gidx"]], ucap[wl1["gidx"]], world, rank)
        bs = Batch(h, wl1, mine_s)
        bs.parity_gate()
        s_ms, _, _ = bs.timed(args.steps, args.warmup)
        s_ms_max = h.max_over_ranks(s_ms) / args.steps
        s_uncomp = h.sum_over_ranks(bs.uncomp_bytes)
        skms, sstats = bs.kernel_times(min(args.steps, 3))
        strong = {"streams": per_rank, "streams_per_rank": bs.n, "ms_per_step": s_ms_max, "value": s_uncomp / (s_ms_max * 1e-3) / 1e9,
                  "unit": "GB/s", "speedup_vs_one_gpu": (total_ms_max / args.steps) / s_ms_max,
                  "speedup_basis": "this run's weak step (every rank decodes the whole %d-stream batch's worth: the N = 1 workload) / the strong step" % per_rank,
                  "kernel_ms_rank0": {k: v for k, v in skms.items() if v > 0.004}, "gated_to_fused": sstats["gated_to_fused"]}
        del bs
        torch.cuda.empty_cache()

    # ---- the one exchange SURVEY.md 8(e) names: the compressed batch is resident on rank 0 and the decoded streams are wanted
    # back there.  scatter (NCCL point-to-point, one message per peer) -> decode -> gather, timed separately ----
    exchange = None
    if world > 1 and args.scaling == "weak" and not args.no_exchange:
        from brotli_rs_b200.batch import slot_offsets
        from brotli_rs_b200.shard import gather_outputs, scatter_batch
        wl1 = build_workload(args.workload, per_rank)
        g1 = wl1["gidx"]
        buf = lens1 = caps1 = None
        if rank == 0:
            full = Batch(h, wl1, np.arange(per_rank))
            buf, lens1, caps1 = full.d_in[: int(full.in_off[-1])], ulen_in[g1], ucap[g1]

        def wall(fn):
            h.barrier()
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            return r, h.max_over_ranks(dt)

        best = {"scatter": 1e9, "decode": 1e9, "gather": 1e9}
        for _ in range(3):
            (mine_x, my_buf, my_lens, my_caps), t_sc = wall(lambda: scatter_batch(buf, lens1, caps1, src=0, device=dev))
            x_in_off = torch.from_numpy(np.concatenate([[0], np.cumsum(my_lens)]).astype(np.int64)).to(dev)
            x_out_off_h = slot_offsets(my_caps).astype(np.int64)
            x_out_off = torch.from_numpy(x_out_off_h).to(dev)
            x_in = torch.zeros(my_buf.numel() + 16, dtype=torch.uint8, device=dev)
            x_in[: my_buf.numel()] = my_buf
            x_out = torch.empty(int(x_out_off_h[-1]), dtype=torch.uint8, device=dev)
            x_len = torch.empty(len(mine_x), dtype=torch.int64, device=dev)
            x_st = torch.empty(len(mine_x), dtype=torch.int32, device=dev)
            _, t_de = wall(lambda: dec.decode_device(x_in, x_in_off, x_out, x_out_off, x_len, x_st))
            res, t_ga = wall(lambda: gather_outputs(mine_x, x_out, x_out_off_h, x_len, x_st, per_rank, dst=0))
            if rank == 0:
                st_all, len_all, rank_of, slot_off, buffers = res
                assert (st_all == wl1["status"][g1]).all() and int(len_all.sum()) == int(bt.ulen_out[g1].sum())
                k = int(np.random.default_rng(1).integers(0, per_rank))
                u = int(g1[k])
                assert buffers[rank_of[k]][int(slot_off[k]): int(slot_off[k]) + int(len_all[k])].cpu().numpy().tobytes() == wl1["raws"][u]
                del buffers, res
            best = {"scatter": min(best["scatter"], t_sc), "decode": min(best["decode"], t_de), "gather": min(best["gather"], t_ga)}
            del x_out, x_in, my_buf
            torch.cuda.empty_cache()
        comp1 = float(ulen_in[g1].sum())
        out1 = float(slot_offsets(ucap[g1])[-1])
        exchange = {"streams": per_rank, "scatter_ms": 1e3 * best["scatter"], "decode_ms": 1e3 * best["decode"], "gather_ms": 1e3 * best["gather"],
                    "scatter_bytes_sent_by_rank0": comp1 * (world - 1) / world, "gather_bytes_received_by_rank0": out1 * (world - 1) / world,
                    "scatter_gbs_out_of_rank0": comp1 * (world - 1) / world / best["scatter"] / 1e9,
                    "gather_gbs_into_rank0": out1 * (world - 1) / world / best["gather"] / 1e9,
                    "gather_gbs_per_peer_link": out1 / world / best["gather"] / 1e9,
                    "how": "wall clock around barrier + synchronize, max over ranks, best of 3; torch.distributed batch_isend_irecv over NCCL "
                           "(one message per peer); the scatter includes the per-shard gather of streams on rank 0 and the 16 B / stream size broadcast"}
        if rank == 0:
            del full
        torch.cuda.empty_cache()

    # ---- end to end through the C ABI with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.e2e_steps or min(args.steps, 5)
        # Host memory for the end-to-end leg is bounded per BOX: with N ranks each pins 1/N of a BASELINE-size batch (at
        # weak scaling the first 1/N of its shard), so that N = 8 does not pin 8 x 26 GB.
        ne = n if (world == 1 or args.scaling == "strong") else max(1, n // world)
        numa_note = bind_to_gpu_numa_node(torch, h.local_rank) if world > 1 else "single rank: not bound"
        e_in, e_out = int(in_off[ne]), int(out_off[ne])
        h_in = torch.empty(e_in, dtype=torch.uint8).pin_memory()
        h_in.copy_(bt.d_in[:e_in])
        h_out = torch.empty(e_out, dtype=torch.uint8).pin_memory()
        hin_np, hout_np = h_in.numpy(), h_out.numpy()
        in_off_u, out_off_u = in_off[: ne + 1].astype(np.uint64), out_off[: ne + 1].astype(np.uint64)
        dec.decode_host(hin_np, in_off_u, out_off_u, out=hout_np)        # warm-up (also sizes the device staging)
        h.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            _, out_len, status = dec.decode_host(hin_np, in_off_u, out_off_u, out=hout_np)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e_uncomp = float(ulen_out[uidx[:ne]].sum())
        assert (status == wl["status"][uidx[:ne]]).all() and float(out_len[status == 0].sum()) == e_uncomp
        for k in bt.check[bt.check < ne][-8:]:
            u = int(uidx[k])
            if raws[u] is not None:
                assert hout_np[int(out_off[k]): int(out_off[k]) + int(ulen_out[u])].tobytes() == raws[u]
        e2e_t = h.max_over_ranks(t1 - t0)
        # what the link alone gives: the same bytes with bare pinned copies, all ranks at once
        h.barrier()
        ceil_t = pcie_ceiling(h, h_out, bt.d_out[:e_out], h_in, bt.d_in[:e_in])
        sum_uncomp = h.sum_over_ranks(e_uncomp)
        e2e = {"value": sum_uncomp * e2e_steps / e2e_t / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": int(e_in + 2 * 8 * (ne + 1)), "d2h_bytes_per_step": int(e_out + 12 * ne),
               "steps": e2e_steps, "streams_per_rank": int(ne), "api": "bro_batch_decode_host (pinned host buffers, per rank; "
               "slices of the batch: the copy out of slice k overlaps the decode of slice k + 1)",
               "ceiling_gbs": sum_uncomp / ceil_t / 1e9,
               "ceiling": "the same H2D + D2H bytes as bare pinned cudaMemcpyAsync on every rank at once (each trial from a barrier, "
                          "its slowest rank; best of 3), in the metric's unit",
               "numa": numa_note,
               "frac_of_ceiling": (sum_uncomp * e2e_steps / e2e_t) / (sum_uncomp / ceil_t)}
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

    # ---- the other BASELINE configurations, device-resident, a few steps each (rank 0's GPU; N = 8 adds only C5) ----
    extra = None
    if args.extra_workloads and args.workload == "c4_highratio_w16":
        del bt
        torch.cuda.empty_cache()
        names_x = ["c1_alice29_single", "c2_quickfox_x10k", "c3_corpus_x1000", "c5_stored_10k", "c5b_literals_10k",
                   "c6_text_q11_w16", "c7_far_w22"] if world == 1 else ["c5_stored_10k"]
        extra = []
        if rank == 0:
            for nm in names_x:
                try:
                    extra.append(quick_workload(h, nm, 5, 3))
                except Exception as e:     # a side measurement never takes the headline line down
                    extra.append({"workload": nm, "error": repr(e)[:200]})
        h.barrier()

    if rank == 0:
        line = {
            "metric": "uncompressed GB/s (many-stream batch)", "value": value, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_record(args, desc, n_streams, len(streams), n, all_comp, all_uncomp, world),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if strong is not None:
            line["strong"] = strong
        if exchange is not None:
            line["exchange"] = exchange
        if extra is not None:
            line["extra_workloads"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py -- BLAST hits/s through blastfilter + scan + liftover (bit-exact) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--hits H]

A "step" is one pass of the whole hot path over one synthetic hit table per GPU:
    jx_blastfilter(raw .last text) -> jx_scan_* (filtered hits) -> jx_liftover_text(raw text, anchors)
Workload at N = 1: BASELINE.json configs[1] -- synthetic diploid pair, 40 k genes each, 20 M
hits with tandem arrays (seeded, jcvi_b200/synth.py).  At N > 1 every rank processes its OWN
genome pair of that size (the pangenome sharding of SURVEY.md 8(e): no data-path collective),
so scaling is weak and `value` = all ranks' hits / max-over-ranks time.

`value`  : text already resident in HBM (jx_text_upload before the timed region), in-memory
           hand-over of the filtered hits to scan (jx_scan_filtered == writing `.filtered` with
           "%.3g" and reading it back); liftover reuses the records blastfilter tokenized from the
           same resident table in the same step (the reference parses that file twice), so a step
           tokenizes the 1.43 GB table exactly once.
`e2e`    : the same three C-ABI calls with HOST buffers -- the pinned raw table is copied
           host->device inside the timed region (once, jx_text_upload: blastfilter and liftover
           both read it), the `.filtered` text is formatted, copied back and handed to
           jx_scan_text as a host buffer, all results are copied back.
`roofline`: the tokenizer + hash join kernel (k_tokenize), algorithmic bytes = text bytes read
           + 16 B record written per line, over its CUDA-event time on the launching stream.
`cpu_baseline` / `--impl reference`: the CPU oracle (oracle/jcvi_oracle.cpp, a sequential
           restatement of the reference's algorithm; the Python reference itself cannot travel
           to the GPU box) on a bounded sample of the same workload: one independent copy of the sample per
           host thread (the reference is sequential per table; many cores = many pair files).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "BLAST hits/s through filter+scan+liftover (bit-exact)"
GENES = 40000
HITS = 20_000_000
CPU_SAMPLE_LINES = 500_000


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_workload(seed, hits, genes=GENES, outdir=None):
    from jcvi_b200 import synth
    from jcvi_b200.formats.bed import Bed
    outdir = outdir or tempfile.mkdtemp(prefix="jcvi_b200_bench_")
    rng = np.random.default_rng(seed)
    gq = synth.Genome("A", genes, 12, rng)
    gs = synth.Genome("B", genes, 12, rng)
    arrays = synth.make_hits(gq, gs, hits, rng)
    text = synth.format_hits(gq, gs, arrays)
    qbed, sbed = os.path.join(outdir, "A.bed"), os.path.join(outdir, "B.bed")
    gq.write_bed(qbed, np.random.default_rng(seed + 7))
    gs.write_bed(sbed, np.random.default_rng(seed + 8))
    return dict(dir=outdir, qbed_path=qbed, sbed_path=sbed, qbed=Bed(qbed), sbed=Bed(sbed), text=text,
                n_hits=int(len(arrays[0])))


def head_lines(text, nlines):
    """First `nlines` lines of a uint8 text (cut at a line boundary)."""
    nl = np.flatnonzero(text[: min(len(text), nlines * 200)] == 10)
    if len(nl) <= nlines:
        return text
    return text[: nl[nlines - 1] + 1]


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region: NVML (same counters nvidia-smi prints
    as clocks.sm / clocks_event_reasons.*) every 5 ms, nvidia-smi as a fallback."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                     getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                     getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                     getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap"}
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag.is_set():
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(mx))
                r = int(get_reasons(h))
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                self.stop_flag.wait(0.005)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                r = [x.strip() for x in out.split(",")]
                self.sm.append(float(r[0])); self.mx.append(float(r[1]))
                for i in range(4):
                    if r[2 + i].lower().startswith("active"):
                        self.reasons.add(names[i])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_cpu_oracle(w, sample_text, steps, warmup, threads=1):
    """hits/s of the CPU oracle on the sample: blastfilter -> scan -> liftover, file to file.
    threads > 1: that many INDEPENDENT copies of the sample run side by side, one per host thread -- the
    reference's algorithm is sequential per table (blastfilter.py:48-95), so the way it uses a many-core host is
    one genome-pair file per core (config C4).  Returns (lines per copy, wall seconds per step)."""
    import threading
    import oracle
    d = w["dir"]
    n = int((sample_text == 10).sum())
    lasts = []
    for t in range(threads):
        last = os.path.join(d, "A.B.sample.last" if t == 0 else "A.B.sample.t%d.last" % t)
        sample_text.tofile(last)
        lasts.append(last)

    def one(last):          # ctypes releases the GIL for the duration of each call
        stem = last[:-len(".last")]
        oracle.blastfilter(last, w["qbed_path"], w["sbed_path"], out_path=last + ".filtered")
        oracle.scan(last + ".filtered", stem + ".anchors", w["qbed_path"], w["sbed_path"])
        oracle.liftover(last, stem + ".anchors", w["qbed_path"], w["sbed_path"], out_path=stem + ".lifted.anchors")

    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if threads == 1:
            one(lasts[0])
        else:
            th = [threading.Thread(target=one, args=(l,)) for l in lasts]
            for x in th:
                x.start()
            for x in th:
                x.join()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    for last in lasts[1:]:
        stem = last[:-len(".last")]
        for f in (last, last + ".filtered", stem + ".anchors", stem + ".lifted.anchors"):
            try:
                os.remove(f)
            except OSError:
                pass
    return n, times


def impl_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    hits = args.hits or HITS
    sample = min(CPU_SAMPLE_LINES, hits)
    # the sample is the head of the rank-0 workload (same seed, same table as the GPU arm)
    w = make_workload(100, hits)
    st = head_lines(w["text"], sample + 3)
    warm = min(args.warmup, 1)
    T = host_threads()
    n1, t1 = run_cpu_oracle(w, st, 1, 0, threads=1)                 # one table on one core, for the record
    n, times = run_cpu_oracle(w, st, args.steps, warm, threads=T)   # every core busy with its own table
    ms = 1000.0 * sum(times) / len(times)
    val = T * n / (ms / 1000.0)
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "hits/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 text / int32 ranks / f32 scores / f64 C-score", "data": "synthetic",
        "config": {"workload": "C2: synthetic diploid pair, 40k genes each, 20M hits with tandem arrays",
                   "sample": "%d independent copies of the first %d lines per step, one per host thread" % (T, n)},
        "cpu_baseline": {"value": val, "unit": "hits/s", "cores": T, "kind": "port",
                         "sample": "first %d lines of the C2 table, blastfilter+scan+liftover file to file, "
                                   "oracle/jcvi_oracle.cpp; %d independent copies side by side, one per host thread (the "
                                   "reference is sequential per table: a many-core host works on one pair file per core); "
                                   "one copy on one thread: %.4g hits/s" % (n, T, n1 / t1[0])},
        "e2e": {"value": val, "unit": "hits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)


def impl_ours(args):
    import torch
    import torch.distributed as dist
    from jcvi_b200.engine import get_engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("BENCH_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hits = args.hits or HITS
    t0 = time.time()
    w = make_workload(100 + rank, hits)
    text = w["text"]
    n_hits = w["n_hits"]
    log("[rank %d] workload: %d hits, %.1f MB, generated in %.1fs" % (rank, n_hits, len(text) / 1e6, time.time() - t0))

    eng = get_engine(local)
    eng.load_pair(w["qbed"], w["sbed"], False)
    stream = torch.cuda.ExternalStream(int(eng.lib.jx_stream_handle(eng.ctx)), device=local)
    pin_text = torch.from_numpy(text).pin_memory()
    dev_text = eng.upload(pin_text)        # resident before the timed region of `value`
    torch.cuda.synchronize()

    def anchors_of(s):
        blk = np.repeat(np.arange(s.n_blocks, dtype=np.int32), np.diff(s.block_start))
        return s.q_rank, s.s_rank, blk

    def step_device():
        # one library call (jx_pipeline): survivors and anchors stay on the device between the stages; the anchors and the
        # lifted pairs come back to the host inside the step, the `.filtered` text is not formatted (`e2e` includes it)
        f, s, l = eng.pipeline(dev_text, want_text=False, skip_arrays=True, liftover_dist=10)
        return s, l

    e2e_log = []

    def step_e2e():
        t0 = time.perf_counter()
        dt = eng.upload(pin_text)                      # one asynchronous H2D of the raw table ...
        f = eng.blastfilter(dt, want_text=True, skip_arrays=True)   # ... overlapped with tokenizing; `.filtered` text comes back
        t1 = time.perf_counter()
        s = eng.scan_text(f.text)                      # the `.filtered` bytes, as a host buffer
        t2 = time.perf_counter()
        aq, as_, ab = anchors_of(s)
        l = eng.liftover_text(dt, aq, as_, ab, dist=10)  # second pass over the raw table, already resident
        t3 = time.perf_counter()
        e2e_log.append((t1 - t0, t2 - t1, t3 - t2))
        h2d = len(text) + len(f.text) + 3 * 4 * len(aq)
        d2h = (len(f.text) + 12 * s.n_anchors + 16 * l.n_lifted + 5 * len(aq))
        return s, l, f, h2d, d2h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        t_wall = time.perf_counter()
        for a, b in ev:
            a.record(stream)
            fn()
            b.record(stream)
        barrier()
        wall = time.perf_counter() - t_wall
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        # a step is a sequence of launches and small synchronous result copies on one stream;
        # the event span is the device-side clock, the wall clock is its upper bound
        return dev_ms, wall * 1000.0

    # ---- warm-up (also parity self-check: device path == host-buffer path) -----------------------
    for _ in range(max(args.warmup, 3)):
        s0, l0 = step_device()
    s1, l1, f1, h2d, d2h = step_e2e()
    assert np.array_equal(s0.q_rank, s1.q_rank) and np.array_equal(s0.s_rank, s1.s_rank) and \
        np.array_equal(s0.block_start, s1.block_start), "device and host-buffer paths disagree (scan)"
    assert np.array_equal(l0.q_rank, l1.q_rank) and np.array_equal(l0.block, l1.block), "paths disagree (liftover)"
    step_e2e()
    dev_text = eng.upload(pin_text)        # the e2e steps replaced the held text
    torch.cuda.synchronize()
    for _ in range(2):
        step_device()

    # ---- timed: device-resident ------------------------------------------------------------------------
    eng.tokenizer_timing(reset=True)
    l0n = eng.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    dev_ms, wall_ms = timed(step_device, args.steps)
    launches = eng.launch_count() - l0n
    tok_ms, tok_launches, tok_bytes, tok_lines = eng.tokenizer_timing(reset=True)
    # ---- the same with record reuse off: liftover tokenizes the table again (two passes per step, like the
    # reference's two parses of the raw file); reported next to `value`, not instead of it
    eng.set_record_reuse(False)
    step_device()
    noreuse_ms, _ = timed(step_device, args.steps)
    eng.set_record_reuse(True)
    eng.tokenizer_timing(reset=True)
    # ---- timed: end to end with host buffers --------------------------------------------------------------
    e2e_ms, e2e_wall = timed(lambda: step_e2e(), args.steps)
    clocks = sampler.summary()          # sampled across both timed regions
    log("[rank %d] e2e sub-steps (s) blastfilter+upload / scan_text / liftover: %s" % (
        rank, ["%.4f %.4f %.4f" % x for x in e2e_log[-args.steps:]]))

    t = torch.tensor([dev_ms, e2e_ms, noreuse_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, noreuse_ms = float(t[0]), float(t[1]), float(t[2])
    ms_per_step = dev_ms / args.steps
    value = world * n_hits / (ms_per_step / 1000.0)
    e2e_value = world * n_hits / (e2e_ms / args.steps / 1000.0)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        try:   # DRAM bytes per 64 MiB launch from the committed `ncu --set full` capture (profiles/)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_tokenize_traffic.json")))
            traffic = float(tj["dram_bytes_per_launch"])
        except Exception:
            pass
        alg_bytes = tok_bytes + 16 * tok_lines
        achieved = alg_bytes / (tok_ms / 1000.0) / 1e9 if tok_ms > 0 else 0.0
        roof = {"bound": "hbm", "kernel": "k_tokenize (K1 tokenizer + K2 hash join, fused K4 best-score)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "B200_PROFILING.md fallback",
                "traffic": traffic, "traffic_note": "dram__bytes_read+write of one launch over the whole resident table (1426.7 MB text + 320.0 MB records algorithmic), ncu capture in profiles/r1_tokenize_traffic.json",
                "launches": tok_launches, "avg_launch_ms": tok_ms / max(tok_launches, 1),
                "algorithmic_bytes_per_launch": alg_bytes / max(tok_launches, 1),
                "share_of_step": tok_ms / dev_ms if dev_ms else None}
        # CPU baseline: bounded sample of the same table, + parity of the GPU path on that sample
        sample = head_lines(text, CPU_SAMPLE_LINES + 3)
        T = host_threads()
        ncpu, times = run_cpu_oracle(w, sample, 1, 0, threads=1)
        _, times_all = run_cpu_oracle(w, sample, 1, 0, threads=T) if T > 1 else (ncpu, times)
        cpu_val = T * ncpu / times_all[0]
        from jcvi_b200.compara import blastfilter as gbf, synteny as gsyn
        d = w["dir"]
        last = os.path.join(d, "A.B.sample.last")
        rd = lambda p: open(p, "rb").read()
        ref_filtered = rd(last + ".filtered")
        t_cli = time.perf_counter()
        gbf.main([last, "--qbed=" + w["qbed_path"], "--sbed=" + w["sbed_path"]])
        parity = rd(last + ".filtered") == ref_filtered
        out_l = gsyn.scan([last + ".filtered", os.path.join(d, "g.anchors"), "--qbed=" + w["qbed_path"],
                           "--sbed=" + w["sbed_path"], "--liftover=" + last, "--liftover_dist=10"])
        parity = parity and rd(os.path.join(d, "g.anchors")) == rd(os.path.join(d, "A.B.sample.anchors"))
        t_cli = time.perf_counter() - t_cli
        parity = parity and rd(out_l) == rd(os.path.join(d, "A.B.sample.lifted.anchors"))
        out = {
            "metric": METRIC, "value": value, "unit": "hits/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 text / int32 ranks / f32 scores / f64 C-score", "data": "synthetic",
            "config": {"workload": "C2: synthetic diploid pair, 40k genes each, 20M hits with tandem arrays",
                       "hits_per_gpu": n_hits, "text_bytes_per_gpu": int(len(text)), "genes": [GENES, GENES],
                       "options": "cscore=0.7 tandem_Nmax=10 dist=20 min_size=4 liftover_dist=10",
                       "parallelism": "one genome pair per GPU (no data-path collective)",
                       "tokenizer_passes_per_step": 1,
                       "record_reuse": "liftover re-uses the records blastfilter tokenized from the same resident "
                                       "table in the same step (outputs identical); with reuse off (two passes, "
                                       "jx_set_record_reuse(0)): %.3f ms/step = %.4g hits/s" % (
                                           noreuse_ms / args.steps, world * n_hits / (noreuse_ms / args.steps / 1000.0)),
                       "l2": "inputs (%.0f MB text per pass) are larger than the 126 MB L2" % (len(text) / 1e6),
                       "outputs": {"filtered": int(f1.n), "anchors": int(s0.n_anchors), "blocks": int(s0.n_blocks),
                                   "lifted": int(l0.n_lifted)}},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "hits/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": {"value": cpu_val, "unit": "hits/s", "cores": T, "kind": "port",
                             "sample": "first %d lines of this table, blastfilter+scan+liftover file to file, "
                                       "oracle/jcvi_oracle.cpp: %d independent copies side by side, one per host thread "
                                       "(%.2f s; one copy on one thread %.2f s = %.4g hits/s); the drop-in CLI functions on the "
                                       "same files, file to file incl. BED load and Python: %.2f s; GPU output on the sample "
                                       "byte-identical: %s" % (ncpu, T, times_all[0], times[0], ncpu / times[0], t_cli, parity)},
            "wall_ms_per_step": wall_ms / args.steps,
        }
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line: everything else any library prints on fd 1 (NCCL's version
    banner, ...) is sent to stderr; emit() writes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--hits", type=int, default=0, help="override hits per GPU (debugging only)")
    args = ap.parse_args()
    if args.impl == "reference":
        impl_reference(args)
    else:
        impl_ours(args)


if __name__ == "__main__":
    main()

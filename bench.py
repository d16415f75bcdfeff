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
    if os.environ.get("BENCH_QUERY_GROUPED"):
        # measurement aid (not the bench workload): the same hits in the order lastal / BLAST write them, all alignments of
        # a query together, instead of the generator's uniformly shuffled lines
        o = np.argsort(arrays[0], kind="stable")
        arrays = tuple(a[o] for a in arrays)
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


# ---- strong scaling: ONE genome pair sharded by chromosome pair over the ranks (BASELINE configs[2] and [4]) -----------
SS_CHUNKS = 8
SS_CONFIGS = {
    # name: (genes q, chromosomes q, genes s, chromosomes s, options)
    "c3": dict(Gq=110_000, Kq=21, Gs=35_000, Ks=7, seed=3, cscore=0.7, tandem_Nmax=10, dist=20, N=4, liftover_dist=10,
               label="C3: synthetic hexaploid-vs-diploid (110k vs 35k genes, three homeologous copies), C-score 0.7"),
    "c5": dict(Gq=60_000, Kq=16, Gs=60_000, Ks=16, seed=5, cscore=0.7, tandem_Nmax=10, dist=20, N=4, liftover_dist=20,
               label="C5: liftover-heavy -- sparse anchors (min_size=4), raw hits rescued within dist=20"),
}


def ss_workload(cfg, hits, world, rank, outdir):
    """This rank's slice of the table (chunks [8 r / N, 8 (r + 1) / N) of 8 seeded chunks: the same table for every N)
    + the two BEDs.  -> (text uint8, n_hits of the slice, qbed, sbed)"""
    from concurrent.futures import ThreadPoolExecutor
    from jcvi_b200 import synth
    from jcvi_b200.formats.bed import Bed
    c = SS_CONFIGS[cfg]
    rng = np.random.default_rng(c["seed"])
    gq = synth.Genome("W", c["Gq"], c["Kq"], rng)
    gs = synth.Genome("H", c["Gs"], c["Ks"], rng)
    per = hits // SS_CHUNKS
    mine = list(range(SS_CHUNKS * rank // world, SS_CHUNKS * (rank + 1) // world))

    def one(ch):
        arrays = synth.make_hits_chunk(gq, gs, per, ch, SS_CHUNKS, 1000 + c["seed"])
        return synth.format_hits(gq, gs, arrays, header=(ch == 0), nthreads=max(2, host_threads() // max(1, min(4, len(mine)) * max(1, min(world, 4)))))

    with ThreadPoolExecutor(max_workers=min(4, len(mine))) as ex:
        parts = list(ex.map(one, mine))
    text = np.concatenate(parts) if len(parts) > 1 else parts[0]
    qp, sp = os.path.join(outdir, "W.bed"), os.path.join(outdir, "H.bed")
    gq.write_bed(qp, np.random.default_rng(10)); gs.write_bed(sp, np.random.default_rng(11))
    return text, per * len(mine), Bed(qp), Bed(sp)


def _digest(*arrays):
    import hashlib
    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def strong_scaling(cfg, hits, steps, warmup, world, rank, local, verify=False, trace=False):
    """ms per step of blastfilter -> scan -> liftover on ONE genome pair of `hits` raw hits whose table is spread over the
    `world` ranks (slices resident in HBM; N = 1: the single-GPU jx_pipeline).  Collective over the ranks; the dict is
    complete on rank 0.  `checksum` digests the anchors and the lifted pairs: identical for every N."""
    import torch
    import torch.distributed as dist
    from jcvi_b200 import sharded
    from jcvi_b200.engine import get_engine
    c = SS_CONFIGS[cfg]
    d = tempfile.mkdtemp(prefix="jcvi_b200_ss_")
    t0 = time.time()
    text, n_mine, qbed, sbed = ss_workload(cfg, hits, world, rank, d)
    gen_s = time.time() - t0
    eng = get_engine(local)
    eng.load_pair(qbed, sbed, False)
    dev = eng.upload(text)
    torch.cuda.synchronize()
    log("[rank %d] %s slice: %d hits, %.2f GB, generated in %.1fs" % (rank, cfg, n_mine, len(text) / 1e9, gen_s))
    comm = sharded.DistComm()
    kw = dict(cscore=c["cscore"], tandem_Nmax=c["tandem_Nmax"], xdist=c["dist"], ydist=c["dist"], N=c["N"], liftover_dist=c["liftover_dist"])
    marks = []
    ss_mode = ["single GPU"]

    nosync = bool(os.environ.get("SS_MARK_NOSYNC"))       # host-side phase times only (no added synchronisation)

    def mark(name):
        if trace:
            if not nosync:
                torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))

    def step():
        if world == 1:
            f, s, l = eng.pipeline(dev, want_text=False, skip_arrays=True, cscore=kw["cscore"], tandem_Nmax=kw["tandem_Nmax"], xdist=kw["xdist"],
                                   ydist=kw["ydist"], N=kw["N"], liftover_dist=kw["liftover_dist"])
            return f.n, s, l
        del marks[:]
        mark("start")
        r = sharded.run_sharded([dev], qbed, sbed, comm=comm, engines=[eng], want_filtered=False, mark=mark,
                                mode=os.environ.get("SS_MODE", "auto"), **kw)
        ss_mode[0] = r.mode
        ss_mode[1:] = [r.n_candidates]
        return None, r.scan, r.lifted

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 1)):
        nf, s, l = step()
    sampler = ClockSampler(local)
    sampler.start()
    times = []
    for _ in range(steps):
        barrier()
        t0 = time.perf_counter()
        nf, s, l = step()
        barrier()
        times.append(time.perf_counter() - t0)
    clocks = sampler.summary()
    t = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) * 1e3
    out = {"workload": c["label"], "hits": int(hits), "n_gpus": world, "ms_per_step": ms, "value": hits / (ms / 1e3), "unit": "hits/s",
           "steps": steps, "scaling": "strong", "clocks": clocks,
           "sharding": "single GPU: jx_pipeline" if world == 1 else (
                       "text slices per rank; ncclAllReduce(max) of the best-score table; all-to-all of 16-byte candidates to the owner of "
                       "their chromosome pair (blastfilter, liftover); all-gather of tandem edges and anchors" if ss_mode[0] == "owners" else
                       "text slices per rank (tokenizer, C-score predicate and liftover prefilter are sharded); ncclAllReduce(max) of the "
                       "best-score table; all-gather of the 16-byte candidates, every rank runs pair de-duplication / tandem / scan on "
                       "all of them; the raw records near an anchor are gathered and lifted on rank 0"),
           "mode": ss_mode[0], "candidates": (ss_mode[1] if len(ss_mode) > 1 else None),
           "options": "cscore=%g tandem_Nmax=%d dist=%d min_size=%d liftover_dist=%d" % (c["cscore"], c["tandem_Nmax"], c["dist"], c["N"], c["liftover_dist"])}
    if rank == 0:
        out["outputs"] = {"anchors": int(s.n_anchors), "blocks": int(s.n_blocks), "lifted": int(l.n_lifted)}
        out["checksum"] = _digest(s.q_rank, s.s_rank, s.score, s.block_start, l.q_rank, l.s_rank, l.score, l.block, l.accepted, l.raw_score)
        if trace and marks:
            out["phases_ms"] = {marks[i][0]: round((marks[i][1] - marks[i - 1][1]) * 1e3, 3) for i in range(1, len(marks))}
    if verify and world > 1:
        # rank 0 builds the whole table and runs the single-GPU pipeline on it
        if rank == 0:
            eng.shard_reset()
            whole, _, _, _ = ss_workload(cfg, hits, 1, 0, d)
            f0, s0, l0 = eng.pipeline(whole, want_text=False, skip_arrays=True, cscore=kw["cscore"], tandem_Nmax=kw["tandem_Nmax"],
                                      xdist=kw["xdist"], ydist=kw["ydist"], N=kw["N"], liftover_dist=kw["liftover_dist"])
            out["verified_against_single_gpu"] = bool(
                _digest(s0.q_rank, s0.s_rank, s0.score, s0.block_start, l0.q_rank, l0.s_rank, l0.score, l0.block, l0.accepted, l0.raw_score)
                == out["checksum"])
        dist.barrier()
    eng.shard_reset()
    eng.release_text()
    return out


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


REF_SAMPLE_LINES = 150_000


def reference_built():
    return os.path.exists(os.path.join(ROOT, "baseline", "_ref", ".built"))


def run_cpu_reference(w, sample_text, steps, warmup, threads=1):
    """hits/s of the UNMODIFIED reference (baseline/_ref: jcvi's own blastfilter.main -> synteny.scan --liftover with the
    compiled cblast, one Python process per table -- the reference is single-threaded) on the sample; threads > 1: that
    many independent copies side by side, one process per host thread.  -> (lines per copy, [wall seconds per step], outputs dir)"""
    d = w["dir"]
    n = int((sample_text == 10).sum())
    lasts = []
    for t in range(threads):
        last = os.path.join(d, "A.B.ref%d.last" % t)
        sample_text.tofile(last)
        lasts.append(last)
    script = os.path.join(ROOT, "baseline", "time_ref.py")
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        procs = [subprocess.Popen([sys.executable, script, l, w["qbed_path"], w["sbed_path"], "10"], stdout=subprocess.PIPE,
                                  stderr=subprocess.DEVNULL) for l in lasts]
        outs = [p.communicate()[0] for p in procs]
        dt = time.perf_counter() - t0
        if any(p.returncode != 0 for p in procs):
            raise RuntimeError("the reference failed on the sample")
        inner = max(json.loads(o.decode().strip().splitlines()[-1])["total"] for o in outs)     # without interpreter start-up
        if it >= warmup:
            times.append(inner)
    for last in lasts[1:]:
        stem = last[:-len(".last")]
        for f in (last, last + ".filtered", stem + ".ref.anchors", stem + ".ref.lifted.anchors"):
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
    warm = min(args.warmup, 1)
    T = host_threads()
    if reference_built():
        # the reference's own implementation (Python + compiled cblast), one process per host thread
        st = head_lines(w["text"], min(REF_SAMPLE_LINES, hits) + 3)
        n1, t1 = run_cpu_reference(w, st, 1, 0, threads=1)
        n, times = run_cpu_reference(w, st, args.steps, warm, threads=T)
        ms = 1000.0 * sum(times) / len(times)
        val = T * n / (ms / 1000.0)
        out = {
            "impl": "reference", "metric": METRIC, "value": val, "unit": "hits/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 text / int32 ranks / f32 scores / f64 C-score", "data": "synthetic",
            "config": {"workload": "C2: synthetic diploid pair, 40k genes each, 20M hits with tandem arrays",
                       "sample": "%d independent copies of the first %d lines per step, one Python process per host thread" % (T, n)},
            "cpu_baseline": {"value": val, "unit": "hits/s", "cores": T, "kind": "reference",
                             "sample": "first %d lines of the C2 table through the UNMODIFIED reference (baseline/_ref: "
                                       "jcvi.compara.blastfilter.main -> jcvi.compara.synteny.scan --liftover, compiled cblast), file to "
                                       "file; %d independent copies side by side, one process per host thread (the reference is "
                                       "single-threaded: a many-core host works on one pair file per core); one copy on one core: "
                                       "%.4g hits/s" % (n, T, n1 / t1[0])},
            "e2e": {"value": val, "unit": "hits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        emit(out)
        return
    st = head_lines(w["text"], sample + 3)
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
        # NCCL's own log (communicator size, transports) goes to stderr: stdout is claimed for the one JSON line
        os.environ["NCCL_DEBUG"] = os.environ.get("BENCH_NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,ENV")
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
        # ONE C-ABI call with host buffers (jx_pipeline, text_on_device = 0): the pinned table is copied to the device inside
        # the call (64 MiB chunks, tokenized as they land), the `.filtered` text, the anchors and the lifted pairs come back
        # to host memory before it returns (the text crosses PCIe while scan and liftover run)
        f, s, l = eng.pipeline(pin_text, want_text=True, skip_arrays=True, liftover_dist=10)
        h2d = len(text)
        d2h = len(f.text) + 20 * s.n_anchors + 16 * l.n_lifted + 5 * s.n_anchors
        return s, l, f, h2d, d2h

    def step_e2e_calls():
        # the same through the three stage calls, `.filtered` handed to scan as the host text a user of the files would have
        t0 = time.perf_counter()
        dt = eng.upload(pin_text)
        f = eng.blastfilter(dt, want_text=True, skip_arrays=True)
        t1 = time.perf_counter()
        s = eng.scan_text(f.text)
        t2 = time.perf_counter()
        aq, as_, ab = anchors_of(s)
        l = eng.liftover_text(dt, aq, as_, ab, dist=10)
        t3 = time.perf_counter()
        e2e_log.append((t1 - t0, t2 - t1, t3 - t2))
        return s, l, f

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
    s2, l2, f2 = step_e2e_calls()
    assert bytes(memoryview(f2.text)) == bytes(memoryview(f1.text)) and np.array_equal(s2.q_rank, s1.q_rank) and \
        np.array_equal(l2.q_rank, l1.q_rank), "one call and three calls disagree"
    del s2, l2, f2
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
    def step_separate():        # the three C-ABI calls one by one, every stage tokenizing for itself
        r = eng.blastfilter(dev_text, want_text=False, raw=True, skip_arrays=True)
        try:
            s = eng.scan_filtered(r)
        finally:
            eng.free_filter(r)
        aq, as_, ab = anchors_of(s)
        return s, eng.liftover_text(dev_text, aq, as_, ab, dist=10)

    eng.set_record_reuse(False)
    step_separate()
    noreuse_ms, _ = timed(step_separate, args.steps)
    eng.set_record_reuse(True)
    eng.tokenizer_timing(reset=True)
    # ---- timed: end to end with host buffers --------------------------------------------------------------
    e2e_ms, e2e_wall = timed(lambda: step_e2e(), args.steps)
    clocks = sampler.summary()          # sampled across both timed regions
    calls_ms, _ = timed(lambda: step_e2e_calls(), args.steps)
    log("[rank %d] three-call e2e sub-steps (s) blastfilter+upload / scan_text / liftover: %s" % (
        rank, ["%.4f %.4f %.4f" % x for x in e2e_log[-args.steps:]]))

    t = torch.tensor([dev_ms, e2e_ms, noreuse_ms, calls_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, noreuse_ms, calls_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    ms_per_step = dev_ms / args.steps
    value = world * n_hits / (ms_per_step / 1000.0)
    e2e_value = world * n_hits / (e2e_ms / args.steps / 1000.0)

    # ---- strong scaling: ONE genome pair (BASELINE configs[2], 200 M hits) over the `world` GPUs ------------------------
    strong = None
    ss_hits = args.ss_hits if args.ss_hits >= 0 else int(os.environ.get("BENCH_SS_HITS", "200000000"))
    if ss_hits > 0:
        try:
            eng.release_text()
            del pin_text, dev_text
            strong = strong_scaling("c3", ss_hits, max(3, min(args.steps, 5)), 2, world, rank, local)
        except Exception as e:                       # the weak-scaling line above stands on its own
            strong = {"error": "%s: %s" % (type(e).__name__, e)}
            log("[rank %d] strong-scaling section failed: %s" % (rank, strong["error"]))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        try:   # DRAM bytes per 64 MiB launch from the committed `ncu --set full` capture (profiles/)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_tokenize_traffic.json")))
            traffic = float(tj["dram_bytes_per_launch"])
        except Exception:
            pass
        alg_bytes = tok_bytes + 16 * tok_lines
        achieved = alg_bytes / (tok_ms / 1000.0) / 1e9 if tok_ms > 0 else 0.0
        roof = {"bound": "hbm", "kernel": "k_tokenize (K1 tokenizer + K2 hash join, fused K4 best-score)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "B200_PROFILING.md fallback",
                "traffic": traffic, "traffic_note": "dram__bytes_read+write of one launch over the whole resident table (algorithmic: 1426.7 MB text + 320.0 MB records; the measured write side also holds the 'no record' fillers of the fixed slots, 72 slots per 3840-byte sub-tile), ncu capture in profiles/r2_tokenize_traffic.json",
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
        cpu = {"value": cpu_val, "unit": "hits/s", "cores": T, "kind": "port",
               "sample": "first %d lines of this table, blastfilter+scan+liftover file to file, "
                         "oracle/jcvi_oracle.cpp: %d independent copies side by side, one per host thread "
                         "(%.2f s; one copy on one thread %.2f s = %.4g hits/s); the drop-in CLI functions on the "
                         "same files, file to file incl. BED load and Python: %.2f s; GPU output on the sample "
                         "byte-identical: %s" % (ncpu, T, times_all[0], times[0], ncpu / times[0], t_cli, parity)}
        if reference_built():
            # the reference itself (Python + compiled cblast) on a smaller head of the same table, and the GPU drop-in
            # against ITS output files
            rs = head_lines(text, REF_SAMPLE_LINES + 3)
            nr, tr1 = run_cpu_reference(w, rs, 1, 0, threads=1)
            _, trT = run_cpu_reference(w, rs, 1, 0, threads=T) if T > 1 else (nr, tr1)
            rlast = os.path.join(d, "A.B.ref0.last")
            ref_f, ref_a, ref_l = rd(rlast + ".filtered"), rd(os.path.join(d, "A.B.ref0.ref.anchors")), rd(os.path.join(d, "A.B.ref0.ref.lifted.anchors"))
            gbf.main([rlast, "--qbed=" + w["qbed_path"], "--sbed=" + w["sbed_path"]])
            out_r = gsyn.scan([rlast + ".filtered", os.path.join(d, "gr.anchors"), "--qbed=" + w["qbed_path"], "--sbed=" + w["sbed_path"],
                               "--liftover=" + rlast, "--liftover_dist=10"])
            rpar = rd(rlast + ".filtered") == ref_f and rd(os.path.join(d, "gr.anchors")) == ref_a and rd(out_r) == ref_l
            cpu = {"value": T * nr / trT[0], "unit": "hits/s", "cores": T, "kind": "reference",
                   "sample": "first %d lines of this table through the UNMODIFIED reference (baseline/_ref: jcvi.compara.blastfilter.main "
                             "-> synteny.scan --liftover, compiled cblast), file to file: %d independent copies side by side, one "
                             "Python process per host thread (%.2f s; one copy on one core %.2f s = %.4g hits/s); GPU drop-in output "
                             "on the same files byte-identical to the reference's: %s" % (nr, T, trT[0], tr1[0], nr / tr1[0], rpar),
                   "port": cpu}
        out = {
            "metric": METRIC, "value": value, "unit": "hits/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 text / int32 ranks / f32 scores / f64 C-score", "data": "synthetic",
            "config": {"workload": "C2: synthetic diploid pair, 40k genes each, 20M hits with tandem arrays",
                       "hits_per_gpu": n_hits, "text_bytes_per_gpu": int(len(text)), "genes": [GENES, GENES],
                       "options": "cscore=0.7 tandem_Nmax=10 dist=20 min_size=4 liftover_dist=10",
                       "parallelism": "one genome pair per GPU (no data-path collective)",
                       "tokenizer_passes_per_step": 1,
                       "value_is": "one jx_pipeline call per step on the table resident in HBM: blastfilter -> scan -> liftover with "
                                   "survivors and anchors left on the device, anchors and lifted pairs copied to the host, `.filtered` "
                                   "text not formatted (`e2e` is the same call on the HOST table: host->device copy, `.filtered` text and all results back)",
                       "record_reuse": "liftover re-uses the records blastfilter tokenized from the same resident "
                                       "table in the same step (outputs identical); as three separate C-ABI calls with reuse off "
                                       "(two tokenizer passes, jx_set_record_reuse(0)): %.3f ms/step = %.4g hits/s" % (
                                           noreuse_ms / args.steps, world * n_hits / (noreuse_ms / args.steps / 1000.0)),
                       "l2": "inputs (%.0f MB text per pass) are larger than the 126 MB L2" % (len(text) / 1e6),
                       "outputs": {"filtered": int(f1.n), "anchors": int(s0.n_anchors), "blocks": int(s0.n_blocks),
                                   "lifted": int(l0.n_lifted)}},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "hits/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps,
                    "call": "jx_pipeline(host text, want_text=1): pinned host table in, `.filtered` text + anchors + lifted pairs "
                            "out, one C-ABI call; copies inside the timed region",
                    "three_stage_calls_ms_per_step": calls_ms / args.steps},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "wall_ms_per_step": wall_ms / args.steps,
            "strong_scaling": strong,
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
    ap.add_argument("--ss-hits", type=int, default=-1, dest="ss_hits",
                    help="raw hits of the strong-scaling table (one genome pair over all GPUs); 0 skips the section; default 200 M")
    args = ap.parse_args()
    if args.impl == "reference":
        impl_reference(args)
    else:
        impl_ours(args)


if __name__ == "__main__":
    main()

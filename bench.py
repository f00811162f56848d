#!/usr/bin/env python
"""Benchmark of the filter -> encode -> detect hot path (BASELINE.json metric:
audio-hours/sec at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sweep|filter|crnn|wavenet]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU restatement of the reference path, host cores

One "step" = one pass of the hot path over one batch of synthetic 16 kHz int16 PCM.
Default workload "sweep" (BASELINE config 5, one time-chunk of it): S streams x T s ->
fused filter kernel -> CRNN and WaveNet encode+detect over all hop-2 windows
(get_posterior semantics) -> FAR/FRR counters.  Per-GPU work is fixed (weak scaling);
streams shard over ranks and the only collective is one all-reduce of the int64 counters.
`value` = audio-hours of PCM all ranks processed per second with the PCM resident in
HBM; `e2e` = same with the PCM in pinned host memory, H2D + D2H of the posteriors and
counters inside the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_WINDOW = {"CRNN": 8043648, "Wavenet": 20663552}
BYTES_PER_FRAME = 480                                         # 320 B int16 in + 160 B mel out
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {k: float(d[k]) for k in FALLBACK_PEAKS}, "measured"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------- CPU arm
def _cpu_stream_worker(args):
    """One stream through the CPU restatement of the reference path (oracle), both models."""
    seed, n_samples, models = args
    from threadpoolctl import threadpool_limits
    from oracle import restated as R
    from wakeword_detection_b200 import synth, weights as W
    with threadpool_limits(limits=1):
        pcm = synth.stream_int16(n_samples, seed % synth.N_CLASSES, 11, seed)
        x = R.int16_to_float(pcm)
        n = 0
        for m in models:
            w = W.load_model_dir(os.path.join(ROOT, "weights", m), m)
            mel = R.mel_stream(x, w)
            L = int(w["mel_length"])
            nw = R.eval_windows(mel.shape[0], L)
            j = np.arange(nw)
            for i in range(0, nw, 64):
                jj = j[i:i + 64]
                R.posterior(mel[(2 * jj)[:, None] + np.arange(L)[None, :]], w)
            n += nw
    return n


def cpu_reference_rate(models, n_streams, seconds_per_stream, processes):
    """audio-hours/sec of the oracle on `processes` host processes (1 BLAS thread each)."""
    import multiprocessing as mp
    n_samples = int(seconds_per_stream * 16000)
    jobs = [(s, n_samples, models) for s in range(n_streams)]
    t0 = time.perf_counter()
    if processes == 1:
        for j in jobs:
            _cpu_stream_worker(j)
    else:
        with mp.get_context("spawn").Pool(processes) as pool:
            pool.map(_cpu_stream_worker, jobs)
    dt = time.perf_counter() - t0
    return n_streams * seconds_per_stream / 3600.0 / dt, dt


def measured_traffic(kernel, S, N):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/r1_traffic.json); only valid for the shape it was captured on."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(kernel)
        if d and d["streams"] == S and abs(d["seconds"] - N / 16000.0) < 1e-9:
            return int(d["bytes_per_launch"])
    except Exception:
        pass
    return None


def workload_models(workload):
    return {"sweep": ["CRNN", "Wavenet"], "crnn": ["CRNN"], "wavenet": ["Wavenet"], "filter": []}[workload]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    models = workload_models(args.workload) or ["CRNN"]
    cores = os.cpu_count() or 1
    sec = 30.0 if args.workload != "filter" else 300.0
    rates = []
    for i in range(args.warmup + args.steps):
        if args.workload == "filter":
            r, dt = cpu_filter_rate(cores, sec)
        else:
            r, dt = cpu_reference_rate(models, cores, sec, cores)
        if i >= args.warmup:
            rates.append((r, dt))
    value = float(np.mean([r for r, _ in rates]))
    ms = float(np.mean([dt for _, dt in rates]) * 1e3)
    sample = "%d streams x %.0f s per step, %d processes x 1 thread, numpy restatement of the TFLite graphs" % (cores, sec, cores)
    line = {"impl": "reference", "metric": "audio_hours_per_sec", "value": value, "unit": "audio-h/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(args, None, None),
            "cpu_baseline": {"value": value, "unit": "audio-h/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "audio-h/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def _cpu_filter_worker(args):
    seed, n_samples = args
    from threadpoolctl import threadpool_limits
    from oracle import restated as R
    from wakeword_detection_b200 import synth, weights as W
    with threadpool_limits(limits=1):
        w = W.load_model_dir(os.path.join(ROOT, "weights", "CRNN"), "CRNN")
        pcm = synth.stream_int16(16000, seed % synth.N_CLASSES, 11, seed)
        pcm = np.tile(pcm, n_samples // 16000 + 1)[:n_samples]
        R.mel_stream(R.int16_to_float(pcm), w)
    return 1


def cpu_filter_rate(processes, seconds_per_stream):
    import multiprocessing as mp
    n_samples = int(seconds_per_stream * 16000)
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(processes) as pool:
        pool.map(_cpu_filter_worker, [(s, n_samples) for s in range(processes)])
    dt = time.perf_counter() - t0
    return processes * seconds_per_stream / 3600.0 / dt, dt


def config_dict(args, S, N):
    names = {"sweep": "config5-step: FAR/FRR sweep time-chunk, filter->encode->detect, CRNN + WaveNet, hop-2 windows",
             "crnn": "filter->encode->detect, CRNN, hop-2 windows", "wavenet": "filter->encode->detect, WaveNet, hop-2 windows",
             "filter": "config2: mel filterbank extraction only"}
    c = {"workload": names[args.workload], "sample_rate": 16000, "pcm": "int16", "hop_frames": 2,
         "precision": args.precision, "l2": "inputs larger than L2 (PCM batch > 126 MB)" if S and S * N * 2 > 126e6
         else "L2 flushed between steps"}
    if S:
        c.update({"streams_per_gpu": S, "seconds_per_stream": N / 16000.0})
    return c


# --------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sweep", choices=["sweep", "filter", "crnn", "wavenet"])
    ap.add_argument("--streams", type=int, default=0, help="streams per GPU (0 = workload default)")
    ap.add_argument("--seconds", type=float, default=0.0, help="seconds per stream (0 = workload default)")
    ap.add_argument("--precision", default="tc", choices=["f32", "tc", "tc_fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from wakeword_detection_b200 import _cabi, weights as W, synth, dist as wdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    models = workload_models(args.workload)
    if args.workload == "filter":
        S, N = args.streams or 100000, int((args.seconds or 2.0) * 16000)       # BASELINE config 2
    else:
        S, N = args.streams or 512, int((args.seconds or 10.0) * 16000)
    engines = {m: _cabi.Engine(W.load_model_dir(os.path.join(ROOT, "weights", m), m), local, args.precision)
               for m in (models or ["CRNN"])}
    first = next(iter(engines.values()))
    F = first.num_frames(N)
    nwin = {m: e.num_windows(F, 2) for m, e in engines.items()} if models else {}

    pcm_dev = synth.device_pcm(S, N, seed=1234, device=dev, first_stream=rank * S)
    pcm_host = torch.empty((S, N), dtype=torch.int16).pin_memory()
    pcm_host.copy_(pcm_dev)
    pcm_stage = torch.empty_like(pcm_dev)
    mel = torch.empty((S, F, 40), dtype=torch.float32, device=dev)
    post = {m: torch.empty((S, nwin[m]), dtype=torch.float32, device=dev) for m in models}
    post_host = {m: torch.empty((S, nwin[m]), dtype=torch.float32).pin_memory() for m in models}
    thr = np.arange(0.5, 0.99999, 0.005)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if S * N * 2 <= 126e6 else None
    stage_ms = {"filter": 0.0, **{m: 0.0 for m in models}, "counts": 0.0}
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def counts():
        counters = []
        for m in models:
            seg = np.arange(S + 1, dtype=np.int64) * nwin[m]
            counters.append(engines[m].eval_counts(post[m], seg, thr, "far_edges"))
            counters.append(engines[m].eval_counts(post[m], seg, thr, "frr_max"))
        return counters

    def step(src, timed_stages=None):
        marks = [ev()]
        marks[0].record()
        first.filter(src, 0.0, out=mel)
        marks.append(ev()); marks[-1].record()
        counters = []
        for m in models:
            engines[m].posteriors(mel, 2, out=post[m])
            marks.append(ev()); marks[-1].record()
        for m in models:
            seg = np.arange(S + 1, dtype=np.int64) * nwin[m]
            counters.append(engines[m].eval_counts(post[m], seg, thr, "far_edges"))
            counters.append(engines[m].eval_counts(post[m], seg, thr, "frr_max"))
        marks.append(ev()); marks[-1].record()
        if timed_stages is not None:
            timed_stages.append(marks)
        return counters

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(run, steps):
        sync_all()
        t0, t1 = ev(), ev()
        t0.record()
        for _ in range(steps):
            run()
        t1.record()
        sync_all()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def dev_step(stages=None):
        if flush is not None:
            flush.fill_(1)
        c = step(pcm_dev, stages)
        if world > 1 and c:
            wdist.all_reduce_counters(*c)

    # End to end: the streams are pushed through in E2E_CHUNKS slices; the host->device copy of slice i+1 (copy stream)
    # overlaps filter -> encode -> detect of slice i (compute stream), as a caller feeding host buffers would do it.
    # Three slices growing geometrically (g, 2g, rest; g = the engines' stream granule, i.e. whole waves of the persistent
    # kernels): the first copy is the only one nothing can hide, so it is small, and each later copy is finished long
    # before the kernels of the slices in front of it are (measured: 27.14 ms against 27.44 ms for g / 4g / 4g / rest).
    gran = max([e.stream_granule(F, 2) for e in engines.values()] or [1])
    if gran > 1 and S >= 4 * gran:
        bounds = [0, gran, 3 * gran, S]
    elif S >= 8:
        bounds = [S * i // 4 for i in range(5)]      # no granule (filter-only workloads are copy-bound): four equal slices
    else:
        bounds = [0, S]
    if os.environ.get("WWB_E2E_BOUNDS"):
        bounds = [int(x) for x in os.environ["WWB_E2E_BOUNDS"].split(",")]
    E2E_CHUNKS = len(bounds) - 1
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event() for _ in range(E2E_CHUNKS)]

    def e2e_step():
        main = torch.cuda.current_stream(dev)
        copy_stream.wait_stream(main)          # the staging buffer of the previous step is free
        with torch.cuda.stream(copy_stream):
            for i in range(E2E_CHUNKS):
                sl = slice(bounds[i], bounds[i + 1])
                pcm_stage[sl].copy_(pcm_host[sl], non_blocking=True)
                copied[i].record(copy_stream)
        for i in range(E2E_CHUNKS):
            main.wait_event(copied[i])
            sl = slice(bounds[i], bounds[i + 1])
            first.filter(pcm_stage[sl], 0.0, out=mel[sl])
            for m in models:
                engines[m].posteriors(mel[sl], 2, out=post[m][sl])
        c = counts()
        if world > 1 and c:
            c = wdist.all_reduce_counters(*c)
        for m in models:
            post_host[m].copy_(post[m], non_blocking=True)
        if c:
            torch.stack(list(c)).cpu()
        else:
            mel[:, :1].cpu()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        dev_step()
    launches0 = sum(e.launch_count() for e in engines.values())
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    stages = []
    total_ms = timed(lambda: dev_step(stages), args.steps)
    launches = sum(e.launch_count() for e in engines.values()) - launches0
    clocks = sampler.stop() if rank == 0 else None
    for marks in stages:
        names = ["filter"] + models + ["counts"]
        for i, nme in enumerate(names):
            stage_ms[nme] += marks[i].elapsed_time(marks[i + 1])
    flush_ms = 0.0
    if flush is not None:      # the L2 flush is not part of the hot path: subtract its measured cost
        sync_all()
        a, b = ev(), ev()
        a.record()
        for _ in range(args.steps):
            flush.fill_(1)
        b.record()
        torch.cuda.synchronize()
        flush_ms = a.elapsed_time(b)
    ms_per_step = (total_ms - flush_ms) / args.steps
    audio_h = world * S * N / 16000.0 / 3600.0
    value = audio_h / (ms_per_step / 1e3)

    for _ in range(2):
        e2e_step()
    e2e_ms = timed(e2e_step, args.steps) / args.steps
    e2e_value = audio_h / (e2e_ms / 1e3)

    pk, pk_src = peaks()
    per = {k: v / args.steps for k, v in stage_ms.items()}
    dom = max((k for k in per if k != "counts"), key=lambda k: per[k])
    if dom == "filter":
        ach = S * F * BYTES_PER_FRAME / (per[dom] / 1e3) / 1e9
        roof = {"bound": "hbm", "kernel": "filter_kernel", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "traffic": None}
    else:
        ach = S * nwin[dom] * FLOP_PER_WINDOW[dom] / (per[dom] / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": dom.lower() + " encode+detect", "achieved": ach,
                "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"],
                "traffic": measured_traffic(dom.lower() + " encode+detect", S, N)}
    roof["peak_source"] = pk_src + (" (sustained bf16)" if dom != "filter" else " (copy)")
    extra = {"ms_per_stage": per,
             "filter_GBps": S * F * BYTES_PER_FRAME / (per["filter"] / 1e3) / 1e9,
             "filter_frac_of_hbm": S * F * BYTES_PER_FRAME / (per["filter"] / 1e3) / 1e9 / pk["hbm_gbs"]}
    for m in models:
        tf = S * nwin[m] * FLOP_PER_WINDOW[m] / (per[m] / 1e3) / 1e12
        extra[m + "_TFLOPs"] = tf
        extra[m + "_frac_of_tensor"] = tf / pk["bf16_tflops_sustained"]
        extra[m + "_windows_per_step"] = S * nwin[m]
    if "CRNN" in models and args.precision != "f32" and nwin["CRNN"] > 72:
        # Sliding windows share the conv + GRU-1 projection columns (crnn_tc.cu, CrnnShare): the figures above use the
        # ALGORITHMIC flops of the per-window formulation; these are the flops the kernels execute.
        q = 4                                                     # hop 2: position indices per conv step
        nsp = -(-((nwin["CRNN"] - 1 + 18 * q) // q + 1) // 126)   # strips of 126 conv steps per stream and phase
        cols = 3 * q * nsp * 126                                  # interior + two padded-column variants
        exe = cols * (2432000 + 4669440) / 19.0 / nwin["CRNN"] + 2 * 233472 + 466944 + 8320
        extra["CRNN_shared_columns"] = {"executed_flop_per_window": exe,
                                        "executed_TFLOPs": S * nwin["CRNN"] * exe / (per["CRNN"] / 1e3) / 1e12,
                                        "columns_per_window": cols / nwin["CRNN"]}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            cm = models or ["CRNN"]
            if args.workload == "filter":
                r, dt = cpu_filter_rate(1, 1800.0)
                cpu = {"value": r, "unit": "audio-h/s", "cores": 1, "kind": "port",
                       "sample": "1 stream x 1800 s, filter only, numpy restatement, 1 thread (%.1f s)" % dt}
            else:
                sec = 120.0 * (2.0 / len(cm)) if cm else 120.0
                r, dt = cpu_reference_rate(cm, 1, sec, 1)
                cpu = {"value": r, "unit": "audio-h/s", "cores": 1, "kind": "port",
                       "sample": "1 stream x %.0f s through %s, numpy restatement of the TFLite graphs, 1 thread (%.1f s)"
                                 % (sec, "+".join(cm), dt)}
        h2d = S * N * 2
        d2h = sum(S * nwin[m] * 4 for m in models) + len(models) * 2 * thr.size * 8
        line = {"metric": "audio_hours_per_sec", "value": value, "unit": "audio-h/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "f32" else "f16x2->f32",
                "data": "synthetic", "config": config_dict(args, S, N), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "audio-h/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "extra": extra}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

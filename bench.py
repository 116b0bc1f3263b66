#!/usr/bin/env python
"""bench.py — Gsamples/s of 8-bit baseband through the full coherent-dedispersion chain
(unpack -> fft_r2c -> rfi_s1 -> dedisperse -> watfft -> rfi_s2 -> signal_detect) on B200,
with per-stage achieved HBM GB/s against the measured copy peak, next to the restated
reference CPU path timed on the same box.

Contract (driver): python bench.py --gpus N --steps K --warmup W [--impl reference]
  * a "step" = one block of synthetic baseband through the whole chain on each rank;
  * workload = BASELINE.json configs[1]: 2^24-sample blocks, 8-bit, one stream, single DM,
    C = 2^11 channels (srtb_config.cfg thresholds); N > 1 shards independent blocks across ranks
    (weak scaling, no collective on the data path);
  * `value`  = samples / time with the blocks already resident in HBM (ring of 16 distinct blocks,
    256 MiB > L2, so successive steps never re-read a cached input); blocks alternate over 6 contexts
    (CUDA streams) per GPU, two blocks in flight per context, so one block's small detector-tail kernels
    overlap the next block's FFT sweeps — every block still runs the whole chain and its result is read back;
  * `e2e`    = the same from pinned HOST buffers through srtb_b200_submit_block()/collect_block() (the
    pinned-host ring: H2D of block i overlaps the compute of block i-1); every block's H2D and the D2H
    of its detector result are inside the timed region;
  * `roofline` = dominant stage: algorithmic bytes / CUDA-event time vs the measured
    copy peak (MEASURED_PEAKS.json hbm_gbs, else 6650 fallback); `stages` has every stage;
  * `cpu_baseline` = the CPU oracle (port of the reference operators, OpenMP, all host cores) on one
    block of the same workload (rank 0, N = 1 only).
--impl reference times that CPU path alone (rank 0) with the same JSON shape.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "simple-radio-telescope-backend_b200"))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402

SRTB_RING_SLOTS = 3   # SRTB_B200_RING_SLOTS in include/srtb_b200.h
METRIC = "Gsamples/s 8-bit baseband through full dedisperse chain"
UNIT = "Gsamples/s"

WORKLOADS = {
    # BASELINE.json configs[0] shape (srtb_config_1644-4559.cfg): 2^30 two-bit samples, inverted 64 MHz band
    "config1": dict(log2n=30, bits=2, fmt="simple", channels=1 << 11, dm=-478.80, f_low=1437.0, bw=-64.0,
                    fs=128e6, avg_thr=1.5, sk_thr=1.05, snr=8.0, chan_thr=0.9, maxbox=256, freq_list="1418-1422"),
    # BASELINE.json configs[1]
    "config2": dict(log2n=24, bits=-8, fmt="simple", channels=1 << 11, dm=56.778, f_low=1000.0, bw=500.0,
                    fs=1e9, avg_thr=5.0, sk_thr=1.05, snr=8.0, chan_thr=0.9, maxbox=256, freq_list=""),
    # BASELINE.json configs[2]: dual-pol 400 MHz, 2^26 per stream, DM 562.05, full RFI + detect
    "config3": dict(log2n=26, bits=-8, fmt="naocpsr_snap1", channels=1 << 11, dm=562.05, f_low=1000.0,
                    bw=400.0, fs=8e8, avg_thr=1.5, sk_thr=1.05, snr=8.0, chan_thr=0.9, maxbox=256,
                    freq_list="1018-1022"),
}

STAGES = ["unpack", "fft_r2c", "rfi_s1", "dedisperse", "watfft", "rfi_s2", "signal_detect"]


def stage_bytes(n: int, bits: int) -> dict:
    """algorithmic bytes per stream per block, SURVEY.md §8(d)"""
    b = abs(bits)
    return {"unpack": n * b / 8 + 4 * n, "fft_r2c": 8 * n, "rfi_s1": 12 * n, "dedisperse": 8 * n,
            "watfft": 8 * n, "rfi_s2": 4 * n, "signal_detect": 4 * n}


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# kernels of each per-pipe stage, for summing the ncu DRAM traffic of profiles/traffic.json
STAGE_KERNELS = {
    "unpack": r"^unpack_",
    "fft_r2c": r"^(fft_col(16)?_tma_kernel<.*, 0>|fft_trans(16)?_tma_kernel|fft_trans_r2c(16)?_tma_kernel|fft_pass_kernel|"
               r"r2c_post_kernel|r2c_col0_fixup_kernel)",
    "rfi_s1": r"^(power_sum_kernel|rfi_s1_apply_kernel|rfi_zero_ranges_kernel)",
    "dedisperse": r"^dedisperse_kernel<0>",
    "watfft": r"^fft_row(16)?_tma_kernel<.*, 0>$",
    "rfi_s2": r"^sk_kernel",
    "signal_detect": r"^(colsum_|detect_)",
}


_NCU_US = {}   # per-stage sum of kernel durations in the same committed capture (no launch / event overhead)


def stage_traffic():
    """per-stage DRAM bytes (read + write) per launch from the committed ncu capture, or {}"""
    import re
    p = ROOT / "profiles" / "traffic.json"
    if not p.exists():
        return {}, None
    try:
        d = json.loads(p.read_text())
        out = {}
        seen_fused = False
        for k in d["kernels"]:
            # the capture ends with one fused process_block: stop at its first kernel (RAW first sweep)
            if re.search(r"fft_col(16)?_tma_kernel<.*, [12]>$", k["kernel"]):
                seen_fused = True
            if seen_fused:
                continue
            for stage, pat in STAGE_KERNELS.items():
                if re.search(pat, k["kernel"]):
                    out[stage] = out.get(stage, 0.0) + k["dram_read"] + k["dram_write"]
                    _NCU_US[stage] = _NCU_US.get(stage, 0.0) + k.get("time_us", 0.0)
        return out, d.get("tag")
    except Exception:
        return {}, None


def synth_block(n_samples: int, streams: int, seed: int, bits: int = -8) -> np.ndarray:
    """V1/V2-style synthetic voltage: Gaussian sigma 20 + CW tone + a short burst,
    int8, clipped; multi-stream blocks are laid out by the caller. Sub-byte widths: uniform random
    packed samples (white noise)."""
    rng = np.random.default_rng(0x53525442 + seed)
    if abs(bits) < 8:
        return rng.integers(0, 256, n_samples * streams * abs(bits) // 8, dtype=np.uint8).view(np.int8)
    v = rng.standard_normal(n_samples * streams, dtype=np.float32) * 20.0
    t = np.arange(n_samples * streams, dtype=np.float32)
    v += 30.0 * np.cos(np.float32(2 * np.pi * 0.1185) * (t % 4096))
    mid = (n_samples * streams) // 2
    v[mid:mid + 2048] *= 4.0
    return np.clip(np.rint(v), -127, 127).astype(np.int8)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def snapshot(self):
        """one immediate query (used when the timed region was shorter than the sampling period)"""
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                  "-i", str(self.gpu_index)], capture_output=True, text=True, timeout=10).stdout
            for line in out.strip().splitlines():
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if not self.rows:
            self.snapshot()
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                  ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def oracle_chain_config(w: dict, n: int, pairs):
    import oracle_lib
    oc = oracle_lib.ChainConfig()
    oc.baseband_input_count = n
    oc.baseband_input_bits = w["bits"]
    oc.window = 0
    oc.baseband_freq_low, oc.baseband_bandwidth = w["f_low"], w["bw"]
    oc.baseband_sample_rate, oc.dm = w["fs"], w["dm"]
    oc.baseband_reserve_sample = 0
    oc.rfi_average_threshold, oc.rfi_sk_threshold = w["avg_thr"], w["sk_thr"]
    oc.spectrum_channel_count = w["channels"]
    oc.snr_threshold, oc.channel_threshold = w["snr"], w["chan_thr"]
    oc.max_boxcar_length = w["maxbox"]
    flat = [v for p in pairs for v in p]
    arr = (C.c_float * max(1, len(flat)))(*flat)
    oc._keep = arr
    oc.rfi_pairs = C.cast(arr, C.POINTER(C.c_float))
    oc.n_rfi_pairs = len(pairs)
    return oc


_CPU_THREADS = None


def pick_cpu_threads(w: dict) -> int:
    """the box may expose more logical CPUs than the container may use: time a small block at a few
    thread counts once and keep the fastest ("all the host threads it can use")"""
    global _CPU_THREADS
    if _CPU_THREADS is not None:
        return _CPU_THREADS
    import oracle_lib
    o = oracle_lib.load()
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    cands = sorted({t for t in (avail, avail // 2, avail // 4, 32, 16, 8) if 1 <= t <= avail})
    cfg = oracle_chain_config(w, 1 << 20, [])
    blk = synth_block(1 << 20, 1, 7, w["bits"]).view(np.uint8)
    best, best_t = None, None
    for t in cands:
        o.set_threads(t)
        o.chain(blk, cfg)
        t0 = time.perf_counter()
        o.chain(blk, cfg)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, best_t = dt, t
    o.set_threads(best_t)
    _CPU_THREADS = best_t
    return best_t


def cpu_chain_seconds(w: dict, n: int, reps: int, streams: int):
    """time the CPU oracle chain (restated reference operators, OpenMP) on `reps` blocks of n samples
    per stream. Multi-stream formats are timed as `streams` independent simple-format streams (the
    de-interleave is a negligible part of the CPU time)."""
    import oracle_lib
    o = oracle_lib.load()
    pick_cpu_threads(w)
    pairs = o.eval_rfi_ranges(w["freq_list"]) if w["freq_list"] else []
    cfg = oracle_chain_config(w, n, pairs)
    blocks = [synth_block(n, 1, 1000 + i, w["bits"]) for i in range(min(reps, 2))]
    t0 = time.perf_counter()
    stage = np.zeros(7)
    for i in range(reps):
        for _ in range(streams):
            _, _, _, st = o.chain(blocks[i % len(blocks)].view(np.uint8), cfg)
            stage += st
    dt = time.perf_counter() - t0
    return dt, o.num_threads(), (stage / reps).tolist()


def run_reference(args, w, wname):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = 1 << w["log2n"]
    streams = 2 if w["fmt"] != "simple" else 1
    sample_n = min(n, 1 << 24)      # one step = one block of at most 2^24 samples per stream
    for _ in range(args.warmup):
        cpu_chain_seconds(w, min(sample_n, 1 << 20), 1, 1)
    dt, threads, stage = cpu_chain_seconds(w, sample_n, args.steps, streams)
    samples = sample_n * streams * args.steps
    value = samples / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (fp64 chirp phase)",
        "data": "synthetic",
        "config": {"workload": f"{wname}: 2^{w['log2n']}-sample blocks x{streams} stream(s), "
                               f"{abs(w['bits'])}-bit, C=2^11, DM={w['dm']}",
                   "sample": f"each step = one block of 2^{int(np.log2(sample_n))} samples/stream"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} block(s) of 2^{int(np.log2(sample_n))} samples x{streams} "
                                   "stream(s); restated reference operators + naive radix-2 FFT "
                                   "(what srtb runs without FFTW), OpenMP",
                         "stage_seconds_per_block": dict(zip(STAGES, stage))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))
    return 0


_RESULT_OUT = None


def claim_stdout():
    """Keep the process's stdout for the ONE JSON line: everything else that writes to fd 1 during the run (NCCL's
    version banner, library chatter) goes to stderr."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(text: str):
    out = _RESULT_OUT or sys.stdout
    out.write(text + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default=os.environ.get("SRTB_BENCH_WORKLOAD", "config2"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stage-iters", type=int, default=5)
    ap.add_argument("--contexts", type=int, default=int(os.environ.get("SRTB_BENCH_CONTEXTS", "0")),
                    help="contexts (CUDA streams) per GPU that blocks alternate over (0 = 6 up to 2^24-sample blocks, 4 up "
                         "to 2^27, 2 above: long sweeps are pure HBM streams with little left to overlap)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    wname = args.workload
    w = WORKLOADS[wname]
    if args.impl == "reference":
        return run_reference(args, w, wname)

    import torch
    import srtb_b200

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    n = 1 << w["log2n"]
    if args.contexts <= 0:
        args.contexts = 6 if w["log2n"] <= 24 else (4 if w["log2n"] < 28 else 2)
    fmt = srtb_b200.FORMAT_BY_NAME[w["fmt"]]
    streams = srtb_b200.FORMAT_STREAMS[fmt]
    block_bytes = n * streams * abs(w["bits"]) // 8
    ring = max(2, min(16, (288 << 20) // block_bytes))      # > L2 (126 MB) of distinct input
    stream = torch.cuda.current_stream()
    ctx = srtb_b200.Context(local_rank, stream.cuda_stream)
    # optional extra contexts on their own streams: consecutive blocks alternate over them so the small
    # detector-tail kernels of one block overlap the FFT sweeps of the next
    extra_streams = [torch.cuda.Stream() for _ in range(max(0, args.contexts - 1))]
    ctxs = [ctx] + [srtb_b200.Context(local_rank, st.cuda_stream) for st in extra_streams]

    pairs = srtb_b200.eval_rfi_ranges(w["freq_list"]) if w["freq_list"] else []
    cfg = srtb_b200.BlockConfig()
    cfg.baseband_input_count = n
    cfg.baseband_input_bits = w["bits"]
    cfg.baseband_format = fmt
    cfg.window = 0
    cfg.baseband_reserve_sample = 0
    cfg.baseband_freq_low, cfg.baseband_bandwidth = w["f_low"], w["bw"]
    cfg.baseband_sample_rate, cfg.dm = w["fs"], w["dm"]
    cfg.mitigate_rfi_average_method_threshold = w["avg_thr"]
    cfg.mitigate_rfi_spectral_kurtosis_threshold = w["sk_thr"]
    cfg.spectrum_channel_count = w["channels"]
    cfg.signal_detect_signal_noise_threshold = w["snr"]
    cfg.signal_detect_channel_threshold = w["chan_thr"]
    cfg.signal_detect_max_boxcar_length = w["maxbox"]
    flat = [v for p in pairs for v in p]
    arr = (C.c_float * max(1, len(flat)))(*flat)
    cfg.rfi_freq_pairs = C.cast(arr, C.POINTER(C.c_float))
    cfg.n_rfi_freq_pairs = len(pairs)

    # synthetic blocks: distinct per rank and per ring slot
    host_blocks = []
    for i in range(ring):
        b = synth_block(n, streams, seed=rank * 1000 + i, bits=w["bits"])
        host_blocks.append(torch.from_numpy(b.view(np.uint8)).pin_memory())
    dev_blocks = [hb.cuda(non_blocking=True) for hb in host_blocks]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None, body=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for st in extra_streams:
            st.wait_stream(stream)
        if body:
            body(steps)
        else:
            for i in range(steps):
                fn(i)
        if finish:
            finish()
        for st in extra_streams:
            stream.wait_stream(st)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    detections = [0]

    # device-resident blocks go through the same ring API as the e2e path (two blocks in flight, so the
    # host enqueues block i+1 while block i runs); every block's detector result is still read back
    dtickets = []

    def _dcollect():
        c, t = dtickets.pop(0)
        res = c.collect_block(t)
        detections[0] += sum(int(r.signal_count[b]) for r in res for b in range(r.n_boxcars))

    def step_device(i):
        c = ctxs[i % len(ctxs)]
        dtickets.append((c, c.submit_block_device(cfg, dev_blocks[i % ring], block_bytes)))
        if len(dtickets) >= 2 * len(ctxs):
            _dcollect()

    def drain_device():
        while dtickets:
            _dcollect()

    def device_body(steps):
        """one host thread per context (ctypes releases the GIL inside the C calls): the ~10 launches per block are
        enqueued in parallel, so a slow host core does not cap the device-resident figure. Context c takes blocks
        c, c + n, c + 2n, ... and keeps two in flight."""
        found = [0] * len(ctxs)

        def worker(ci):
            c, mine, det = ctxs[ci], [], 0
            for i in range(ci, steps, len(ctxs)):
                mine.append(c.submit_block_device(cfg, dev_blocks[i % ring], block_bytes))
                if len(mine) >= 2:
                    det += sum(int(r.signal_count[b]) for r in c.collect_block(mine.pop(0)) for b in range(r.n_boxcars))
            while mine:
                det += sum(int(r.signal_count[b]) for r in c.collect_block(mine.pop(0)) for b in range(r.n_boxcars))
            found[ci] = det

        workers = [threading.Thread(target=worker, args=(ci,)) for ci in range(len(ctxs))]
        for t_ in workers:
            t_.start()
        for t_ in workers:
            t_.join()
        detections[0] += sum(found)

    # e2e goes through the pipelined ingest API (pinned-host ring): block i's H2D runs on the copy stream
    # while block i-1 computes; every block's detector result is read back on the host
    tickets = []

    def _collect():
        c, t = tickets.pop(0)
        res = c.collect_block(t)
        detections[0] += sum(int(r.signal_count[b]) for r in res for b in range(r.n_boxcars))

    def step_e2e(i):
        c = ctxs[i % len(ctxs)]
        tickets.append((c, c.submit_block(cfg, host_blocks[i % ring], block_bytes)))
        if len(tickets) >= 2 * len(ctxs):
            _collect()

    def drain_e2e():
        while tickets:
            _collect()

    # ---- device-resident throughput (`value`)
    # untimed warm-up: at least W steps, and enough that EVERY context has seen every ring slot once (first use
    # allocates scratch, builds twiddle tables and sets kernel attributes: cudaMalloc would stall the timed region)
    warm = max(args.warmup, (SRTB_RING_SLOTS + 1) * len(ctxs))
    for i in range(warm):
        step_device(i)
    drain_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = sum(c.launch_count for c in ctxs)
    # one host thread per context was measured slower and erratic under the GIL (100..136 vs a steady 139): opt-in only
    threaded = len(ctxs) > 1 and os.environ.get("SRTB_BENCH_THREADS", "0") == "1"
    ms_total = timed(step_device, args.steps, drain_device, body=device_body if threaded else None)
    launches = sum(c.launch_count for c in ctxs) - l0
    if rank == 0 and not sampler.rows:
        sampler.snapshot()          # short run: take one sample while the GPU is still under load
    ms_per_step = ms_total / args.steps
    samples_per_step = n * streams * world
    value = samples_per_step / (ms_per_step * 1e-3) / 1e9

    # ---- end to end from pinned host memory (`e2e`)
    for i in range(warm):
        step_e2e(i)
    drain_e2e()
    e2e_runs = [timed(step_e2e, args.steps, drain_e2e) / args.steps for _ in range(3)]
    ms_e2e = float(np.median(e2e_runs))           # host-side jitter (PCIe, the feeding thread): median of three
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = samples_per_step / (ms_e2e * 1e-3) / 1e9
    d2h = C.sizeof(srtb_b200.DetectResult) * streams

    # ---- per-stage CUDA-event timing (rank 0): each stage called through the C ABI on one stream's data
    stages = {}
    roofline = None
    if rank == 0:
        peak, peak_src = hbm_peak()
        bytes_per = stage_bytes(n, w["bits"])
        nc = n // 2
        batch = min(w["channels"], nc)
        L = nc // batch
        coef = srtb_b200.norm_coefficient(nc, w["channels"])
        bins = [r for r in (srtb_b200.rfi_range_to_bins(a, b, w["f_low"], w["bw"], nc) for a, b in pairs) if r]
        f_min, bw = np.float32(w["f_low"]), np.float32(w["bw"])
        f_c, df = float(f_min + bw), float(bw / np.float32(nc))
        outs = [torch.empty(n + 2, dtype=torch.float32, device="cuda") for _ in range(streams)]
        buf = outs[0]
        calls = {
            "unpack": lambda i: ctx.unpack(dev_blocks[i % ring], block_bytes, w["bits"], fmt, 0, outs, n),
            "fft_r2c": lambda i: ctx.fft_r2c_inplace(buf, n),
            "rfi_s1": lambda i: ctx.rfi_s1(buf, nc, w["avg_thr"], coef, bins),
            "dedisperse": lambda i: ctx.dedisperse(buf, nc, float(f_min), f_c, df, w["dm"]),
            "watfft": lambda i: ctx.watfft_c2c_backward(buf, L, batch),
            "rfi_s2": lambda i: ctx.rfi_s2_sk(buf, L, batch, w["sk_thr"]),
            "signal_detect": lambda i: ctx.signal_detect(buf, L, batch, 0, w["snr"], w["chan_thr"], w["maxbox"]),
        }
        flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
        for idx, name in enumerate(STAGES):
            times = []
            for it in range(args.stage_iters + 1):
                for prev in STAGES[:idx]:           # realistic input: run the chain up to this stage
                    calls[prev](it)
                flush.fill_(it & 0xFF)              # flush L2 between timed launches
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                calls[name](it)
                e1.record(stream)
                torch.cuda.synchronize()
                if it > 0:
                    times.append(e0.elapsed_time(e1))
            ms = float(np.mean(times))
            per_call_streams = streams if name == "unpack" else 1
            gbs = bytes_per[name] * per_call_streams / (ms * 1e-3) / 1e9
            stages[name] = {"ms": ms, "bytes": bytes_per[name] * per_call_streams, "gbs": gbs, "frac": gbs / peak}
        chain_bytes = sum(bytes_per.values()) * streams
        dom = max((s for s in STAGES), key=lambda s: stages[s]["ms"])
        traffic, traffic_tag = stage_traffic() if wname == "config2" else ({}, None)
        for k_, v_ in traffic.items():
            if k_ in stages:
                stages[k_]["dram_traffic"] = v_
                if _NCU_US.get(k_):   # offline: kernel time under ncu (cold cache), for comparison with the live `ms`
                    stages[k_]["ncu_kernel_us"] = _NCU_US[k_]
                    stages[k_]["ncu_frac"] = stages[k_]["bytes"] / (_NCU_US[k_] * 1e-6) / 1e9 / peak
        roofline = {"bound": "hbm", "kernel": dom, "achieved": stages[dom]["gbs"], "peak": peak,
                    "unit": "GB/s", "frac": stages[dom]["frac"], "traffic": traffic.get(dom),
                    "traffic_source": (f"ncu --set full capture profiles/{traffic_tag}_summary.md (dram read+write of the "
                                       "stage's kernels, one launch each, L2 flushed)") if traffic_tag else None,
                    "peak_source": peak_src,
                    "chain": {"bytes_per_sample": chain_bytes / (n * streams),
                              "achieved": chain_bytes / (ms_per_step * 1e-3) / 1e9,
                              "frac": chain_bytes / (ms_per_step * 1e-3) / 1e9 / peak}}

    # ---- CPU baseline (oracle port) on a bounded sample of the same workload
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sample_n = min(n, 1 << 24)
            cpu_chain_seconds(w, 1 << 18, 1, 1)        # warm the OpenMP pool
            dt, threads, stage = cpu_chain_seconds(w, sample_n, 2, streams)
            fast = None
            try:  # the same chain with a fast CPU FFT (pocketfft) in place of the naive radix-2 for the two FFT stages
                import scipy.fft as sfft
                xs = synth_block(sample_n, 1, 1, w["bits"]).astype(np.float32) if abs(w["bits"]) == 8 else \
                    np.random.default_rng(1).standard_normal(sample_n).astype(np.float32)
                nc_ = sample_n // 2
                cols = min(w["channels"], nc_)
                sfft.rfft(xs[:1 << 16], workers=threads)
                t0 = time.perf_counter()
                spec = sfft.rfft(xs, workers=threads)[:nc_].astype(np.complex64)
                t_r2c = time.perf_counter() - t0
                t0 = time.perf_counter()
                sfft.ifft(spec.reshape(cols, nc_ // cols), axis=1, workers=threads, norm="forward")
                t_wat = time.perf_counter() - t0
                per_block = dt / (2 * streams) - stage[1] / streams - stage[4] / streams + t_r2c + t_wat
                fast = {"value": sample_n / per_block / 1e9, "unit": UNIT,
                        "fft": f"scipy.fft (pocketfft, workers={threads}) for fft_r2c and watfft, other stages as above",
                        "fft_r2c_s": t_r2c, "watfft_s": t_wat}
            except Exception as e:
                fast = {"value": None, "fft": f"failed: {e}"}
            cpu_baseline = {"value": sample_n * streams * 2 / dt / 1e9, "unit": UNIT, "cores": threads,
                            "fast_fft": fast,
                            "kind": "port",
                            "sample": f"2 blocks of 2^{int(np.log2(sample_n))} samples x{streams} stream(s) of "
                                      "this workload; restated reference operators with the in-tree naive "
                                      "radix-2 FFT, OpenMP over all host cores",
                            "stage_seconds_per_block": dict(zip(STAGES, stage))}
        except Exception as e:  # the oracle is test infrastructure: report, never fail the bench
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (fp64 chirp phase)", "data": "synthetic",
            "config": {"workload": f"{wname}: 2^{w['log2n']}-sample blocks x{streams} stream(s) per GPU, "
                                   f"{abs(w['bits'])}-bit {w['fmt']}, C=2^11, DM={w['dm']}, full RFI + detect",
                       "parallelism": f"block-sharded x{world} (no collective)",
                       "l2": f"inputs larger than L2: ring of {ring} distinct blocks ({ring * block_bytes >> 20} MiB)",
                       "contexts_per_gpu": len(ctxs), "submit_threads_per_gpu": len(ctxs) if threaded else 1, "warmup_steps_run": warm,
                       "detections": detections[0]},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                    "runs_ms_per_step": e2e_runs, "note": "median of three runs of K steps each",
                    "h2d_bytes_per_step": block_bytes * world, "d2h_bytes_per_step": d2h * world},
            "gpu_launches": launches,
            "roofline": roofline,
            "stages": stages,
            "cpu_baseline": cpu_baseline,
        }
        emit(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    for c in ctxs:
        c.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())

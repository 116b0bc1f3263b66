#!/usr/bin/env python
"""Runs every per-pipe stage of the config-2 workload exactly once inside a cudaProfiler range, then one
fused process_block — the target of the `ncu --set full --profile-from-start off` capture whose per-kernel
DRAM bytes become profiles/traffic.json (bench.py's roofline.traffic)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "simple-radio-telescope-backend_b200"))
import bench  # noqa: E402
import srtb_b200  # noqa: E402
import ctypes as C  # noqa: E402

w = bench.WORKLOADS["config2"]
n = 1 << w["log2n"]
torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
ctx = srtb_b200.Context(0, stream.cuda_stream)
blk = torch.from_numpy(bench.synth_block(n, 1, 0).view(np.uint8)).cuda()
buf = torch.empty(n + 2, dtype=torch.float32, device="cuda")
nc, C_ = n // 2, w["channels"]
L = nc // C_
coef = srtb_b200.norm_coefficient(nc, C_)
f_min, bw = np.float32(w["f_low"]), np.float32(w["bw"])
f_c, df = float(f_min + bw), float(bw / np.float32(nc))
cfg = srtb_b200.BlockConfig()
cfg.baseband_input_count, cfg.baseband_input_bits, cfg.baseband_format = n, w["bits"], 0
cfg.baseband_freq_low, cfg.baseband_bandwidth, cfg.baseband_sample_rate, cfg.dm = w["f_low"], w["bw"], w["fs"], w["dm"]
cfg.mitigate_rfi_average_method_threshold, cfg.mitigate_rfi_spectral_kurtosis_threshold = w["avg_thr"], w["sk_thr"]
cfg.spectrum_channel_count = C_
cfg.signal_detect_signal_noise_threshold, cfg.signal_detect_channel_threshold = w["snr"], w["chan_thr"]
cfg.signal_detect_max_boxcar_length = w["maxbox"]
flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")


def stages():
    ctx.unpack(blk, n, w["bits"], 0, 0, [buf], n)
    flush.fill_(1)
    ctx.fft_r2c_inplace(buf, n)
    flush.fill_(2)
    ctx.rfi_s1(buf, nc, w["avg_thr"], coef, [])
    flush.fill_(3)
    ctx.dedisperse(buf, nc, float(f_min), f_c, df, w["dm"])
    flush.fill_(4)
    ctx.watfft_c2c_backward(buf, L, C_)
    flush.fill_(5)
    ctx.rfi_s2_sk(buf, L, C_, w["sk_thr"])
    flush.fill_(6)
    ctx.signal_detect(buf, L, C_, 0, w["snr"], w["chan_thr"], w["maxbox"])
    flush.fill_(7)
    ctx.process_block(cfg, blk, n, None, on_device=True)


for _ in range(3):
    stages()
torch.cuda.synchronize()
torch.cuda.profiler.start()
stages()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("stage_once ok")

#!/usr/bin/env python
"""BASELINE config #4 shape: one 2^27-sample 8-bit block swept over 21 trial DMs (0..1000, step 50) with
srtb_b200_process_block_dm_sweep (unpack + R2C once, then per DM: s1 + chirp + waterfall FFT + SK + detect).
Prints one JSON line: blocks/s, trial-Gsamples/s (samples x trials / time) and ms per trial."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "simple-radio-telescope-backend_b200"))
import bench  # noqa: E402
import srtb_b200  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = 1 << log2n
dms = [50.0 * i for i in range(21)]
torch.cuda.set_device(0)
ctx = srtb_b200.Context(0, torch.cuda.current_stream().cuda_stream)
cfg = srtb_b200.BlockConfig()
cfg.baseband_input_count, cfg.baseband_input_bits, cfg.baseband_format = n, -8, srtb_b200.FORMAT_SIMPLE
cfg.baseband_freq_low, cfg.baseband_bandwidth, cfg.baseband_sample_rate, cfg.dm = 1000.0, 500.0, 1e9, 0.0
cfg.mitigate_rfi_average_method_threshold, cfg.mitigate_rfi_spectral_kurtosis_threshold = 5.0, 1.05
cfg.spectrum_channel_count = 1 << 11
cfg.signal_detect_signal_noise_threshold, cfg.signal_detect_channel_threshold = 8.0, 0.9
cfg.signal_detect_max_boxcar_length = 256
blk = torch.from_numpy(bench.synth_block(n, 1, 4).view(np.uint8)).cuda()
ctx.process_block_dm_sweep(cfg, blk, n, dms, on_device=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    res = ctx.process_block_dm_sweep(cfg, blk, n, dms, on_device=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
print(json.dumps({"workload": f"config4 shape: 2^{log2n} samples x {len(dms)} trial DMs", "ms_per_block": dt * 1e3,
                  "ms_per_trial": dt * 1e3 / len(dms), "trial_gsamples_per_s": n * len(dms) / dt / 1e9,
                  "block_gsamples_per_s": n / dt / 1e9}))

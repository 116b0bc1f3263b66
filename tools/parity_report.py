#!/usr/bin/env python
"""Measured parity numbers at BASELINE config #2's full size (2^24 samples, C = 2^11, DM 56.778): every stage through
the C ABI against float64 truth / the CPU oracle, and the fused chain against the oracle chain. Prints a small table
(copied to profiles/<tag>_parity.txt). Test infrastructure: uses oracle/ as the checker."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "simple-radio-telescope-backend_b200"))
import oracle_lib  # noqa: E402
import srtb_b200  # noqa: E402
import test_gpu_parity as T  # noqa: E402

o = oracle_lib.load()
torch.cuda.set_device(0)
ctx = srtb_b200.Context(0, torch.cuda.current_stream().cuda_stream)
n, C_, dm = 1 << 24, 1 << 11, 56.778
nc, L = n // 2, n // 2 // C_
rl2 = lambda a, b: float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(np.asarray(b).ravel()))
bb = T.synth_baseband(n, seed=24)
rows = []
# unpack
buf = torch.empty(n + 2, dtype=torch.float32, device="cuda")
ctx.unpack(T.dev(bb.view(np.uint8)), n, -8, srtb_b200.FORMAT_SIMPLE, 0, [buf], n)
rows.append(("unpack (int8 -> f32)", "mismatching samples", int((buf[:n].cpu().numpy() != bb.astype(np.float32)).sum())))
# r2c vs float64
ctx.fft_r2c_inplace(buf, n)
torch.cuda.synchronize()
X = buf.cpu().numpy().view(np.complex64)
truth = np.fft.rfft(bb.astype(np.float64))
rows.append(("fft_1d_r2c 2^24", "rel-L2 vs float64 FFT", rl2(X, truth)))
rows.append(("fft_1d_r2c 2^24", "max |err| / max |X|", float(np.abs(X - truth).max() / np.abs(truth).max())))
# dedisperse vs fp64 chirp
spec = X[:nc].copy()
d = T.dev(spec)
f_min, bw = np.float32(1000.0), np.float32(500.0)
f_c, df = float(f_min + bw), float(bw / np.float32(nc))
ctx.dedisperse(d, nc, float(f_min), f_c, df, dm)
torch.cuda.synchronize()
f = np.float64(f_min) + np.float64(np.float32(df)) * np.arange(nc)
k = 4.148808e3 * 1e6 * np.float64(np.float32(dm)) / f * ((f - np.float64(np.float32(f_c))) / np.float64(np.float32(f_c))) ** 2
ded_truth = spec.astype(np.complex128) * np.exp(-2j * np.pi * (k - np.trunc(k)))
rows.append(("dedisperse", "rel-L2 vs float64 chirp", rl2(d.cpu().numpy(), ded_truth)))
rows.append(("dedisperse", "rel-L2 vs CPU oracle (fp64 phase, f32 sincos)", rl2(d.cpu().numpy(), o.dedisperse(spec, float(f_min), f_c, df, dm))))
# watfft
ctx.watfft_c2c_backward(d, L, C_)
torch.cuda.synchronize()
wt = np.fft.ifft(ded_truth.reshape(C_, L), axis=1) * L
rows.append(("watfft_1d_c2c [2048][4096]", "rel-L2 vs float64 (chirp + FFT)", rl2(d.cpu().numpy().reshape(C_, L), wt)))
# chain vs oracle
cfg = T.make_block_config(n, -8, srtb_b200.FORMAT_SIMPLE, C_, dm, avg_thr=5.0, sk_thr=1.3, snr=6.0, maxbox=256,
                          pairs=[(1200.0, 1201.0)])
work, eres, eseries, _ = o.chain(bb.view(np.uint8), T.oracle_chain_config(cfg))
hs = np.zeros((srtb_b200.MAX_BOXCARS, L), np.float32)
res = ctx.process_block(cfg, torch.from_numpy(bb.view(np.uint8).copy()).pin_memory(), n, hs, copy_all=True)[0]
got = T._from_device_ptr(ctx.block_spectrum_ptr(0), nc).reshape(C_, L)
esp = work[:n].view(np.complex64).reshape(C_, L)
gz, ez = np.all(got == 0, axis=1), np.all(esp == 0, axis=1)
rows.append(("fused chain (process_block)", "SK/manual zap decisions differing from the oracle (of 2048 channels)", int((gz != ez).sum())))
same = gz == ez
rows.append(("fused chain (process_block)", "dynamic spectrum rel-L2 vs oracle chain", rl2(got[same], esp[same])))
rows.append(("fused chain (process_block)", "zero_count GPU / oracle", f"{res.zero_count} / {eres.zero_count}"))
rows.append(("fused chain (process_block)", "boxcar series rel-L2 vs oracle (boxcar 1)", rl2(hs[0][:eres.series_length[0]], eseries[0][:eres.series_length[0]])))
rows.append(("fused chain (process_block)", "signal counts GPU", [int(res.signal_count[b]) for b in range(res.n_boxcars)]))
rows.append(("fused chain (process_block)", "signal counts oracle", [int(eres.signal_count[b]) for b in range(eres.n_boxcars)]))
print("| stage | quantity | value |\n|---|---|---|")
for a, b, c in rows:
    print(f"| {a} | {b} | {c if not isinstance(c, float) else f'{c:.3e}'} |")

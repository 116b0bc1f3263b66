#!/usr/bin/env python
"""Generates tests/golden/srtb_golden.npz from the REFERENCE'S OWN code (oracle/_ref/libsrtb_ref.so,
the reference headers compiled through oracle/ref_shim). Run in the build container, where
/root/reference exists:   make -C oracle ref && python tests/golden/make_golden.py
The GPU box has no /root/reference; the committed .npz is what travels."""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import ref_lib  # noqa: E402


def main():
    ref = ref_lib.load()
    assert ref is not None, "build oracle/_ref first: make -C oracle ref"
    rng = np.random.default_rng(20260921)
    g = {}
    raw = rng.integers(0, 256, 512, dtype=np.uint8)
    g["unpack_raw"] = raw
    for bits in (1, 2, 4, 8, -8, 16, -16):
        g[f"unpack_b{bits}"] = ref.unpack(raw, raw.size * 8 // abs(bits), bits)
    g["unpack_b-8_hamming"] = ref.unpack(raw, raw.size, -8, 2)
    g["unpack_b2_hann"] = ref.unpack(raw, raw.size * 4, 2, 1)
    a, b = ref.unpack_interleaved_2(raw, raw.size // 2, -8)
    g["il2_a"], g["il2_b"] = a, b
    a, b = ref.unpack_snap1(raw, raw.size // 2)
    g["snap1_a"], g["snap1_b"] = a, b
    for s in (2, 4):
        for i, o in enumerate(ref.unpack_gznupsr_a1(raw, raw.size // s, s)):
            g[f"gznu{s}_{i}"] = o
    for w in (1, 2):
        g[f"window{w}_16"] = np.array([ref.window(w, i, 16) for i in range(16)], np.float32)
    x = (rng.uniform(-1, 1, 256) + 1j * rng.uniform(-1, 1, 256)).astype(np.complex64)
    g["fft_x"] = x
    g["fft_fwd"] = ref.fft_c2c(x, 1)
    g["fft_bwd"] = ref.fft_c2c(x, -1)
    xr = rng.integers(-128, 128, 512).astype(np.float32)
    g["r2c_x"] = xr
    g["r2c_X"] = ref.fft_r2c(xr)
    g["watfft_8x32"] = ref.watfft(x, 32, 8)
    spec = ((rng.standard_normal(2048) + 1j * rng.standard_normal(2048)) * 100).astype(np.complex64)
    spec[rng.integers(0, 2048, 8)] *= 40
    g["s1_x"] = spec
    g["s1_y"] = ref.rfi_s1_pipe(spec, 1.5, 16, 1000.0, 500.0, "1100-1110, 1300.5-1302")
    y = (rng.standard_normal(2048) + 1j * rng.standard_normal(2048)).astype(np.complex64)
    g["dd_x"] = y
    for tag, (fl, bw, dm) in {"a": (1000.0, 500.0, 56.778), "b": (1000.0, 400.0, 562.05), "c": (1437.0, -64.0, -478.80)}.items():
        g[f"dd_{tag}"] = ref.dedisperse_pipe(y, fl, bw, dm)
        g[f"dd_{tag}_params"] = np.array([fl, bw, dm], np.float64)
    C_, L = 16, 256
    d = (rng.standard_normal((C_, L)) + 1j * rng.standard_normal((C_, L))).astype(np.complex64)
    d[2, :] = 3
    d[5, ::8] *= 9
    d[7, :] = 0
    g["s2_x"] = d.reshape(-1)
    g["s2_y"] = ref.rfi_s2_pipe(d.reshape(-1), L, C_, 1.2)
    e = (rng.standard_normal((C_, L)) + 1j * rng.standard_normal((C_, L))).astype(np.complex64)
    e[:, 40:48] *= 10
    e[3, :] = 0
    g["det_x"] = e.reshape(-1)
    hs = ref.signal_detect_pipe(e.reshape(-1), L, C_, 2 * C_ * L, False, 1000.0, 500.0, 1e9, 0.0, 6.0, 0.9, 32)
    g["det_boxcar"] = np.array([h["boxcar"] for h in hs], np.int64)
    g["det_count"] = np.array([h["count"] for h in hs], np.int64)
    g["det_length"] = np.array([h["length"] for h in hs], np.int64)
    for h in hs:
        g[f"det_series_{h['boxcar']}"] = h["series"]
    np.savez_compressed(HERE / "srtb_golden.npz", **g)
    print("wrote", HERE / "srtb_golden.npz", sum(v.nbytes for v in g.values()), "bytes in", len(g), "arrays")
    chain_golden(ref)


def chain_golden(ref):
    """One block through the reference's pipes composed as main.cpp:170-204 wires them (every stage = the reference's
    own code through the shim): the fixture the GPU box checks srtb_b200_process_block against."""
    rng = np.random.default_rng(20260922)
    n, C_, dm = 1 << 15, 16, 0.0
    nc, L = n // 2, n // 2 // C_
    v = np.clip(np.round(rng.standard_normal(n) * 20), -127, 127)
    v[n // 2:n // 2 + 48] += np.round(rng.standard_normal(48) * 90)
    raw = np.clip(v, -127, 127).astype(np.int8).view(np.uint8)
    f_low, bw, fs, avg_thr, sk_thr, snr, chan_thr, maxbox = 1000.0, 500.0, 1e9, 5.0, 1.3, 6.0, 0.9, 64
    spec = ref.fft_r2c(ref.unpack(raw, n, -8))[:nc]
    spec = ref.rfi_s1_pipe(spec, avg_thr, C_, f_low, bw, "1200-1201")
    spec = ref.dedisperse_pipe(spec, f_low, bw, dm)
    spec = ref.watfft(spec, L, C_)
    spec = ref.rfi_s2_pipe(spec, L, C_, sk_thr)
    hs = ref.signal_detect_pipe(spec, L, C_, n, False, f_low, bw, fs, dm, snr, chan_thr, maxbox)
    g = {"raw": raw, "params": np.array([n, C_, dm, f_low, bw, fs, avg_thr, sk_thr, snr, chan_thr, maxbox], np.float64),
         "freq_pairs": np.array([1200.0, 1201.0], np.float32), "spectrum": spec.reshape(C_, L),
         "zero_count": np.array([int(np.sum(np.abs(spec.reshape(C_, L)[:, 0]) ** 2 == 0))], np.int64),
         "det_boxcar": np.array([h["boxcar"] for h in hs], np.int64),
         "det_count": np.array([h["count"] for h in hs], np.int64),
         "det_length": np.array([h["length"] for h in hs], np.int64)}
    np.savez_compressed(HERE / "srtb_chain_golden.npz", **g)
    print("wrote", HERE / "srtb_chain_golden.npz", sum(v_.nbytes for v_ in g.values()), "bytes;",
          "candidates at boxcars", g["det_boxcar"].tolist(), "counts", g["det_count"].tolist())


if __name__ == "__main__":
    main()

"""Golden fixtures generated from the reference's own code (tests/golden/make_golden.py ->
tests/golden/srtb_golden.npz). CPU: the oracle reproduces them; GPU: libsrtb_b200.so reproduces them."""
from pathlib import Path

import numpy as np
import pytest

G = np.load(Path(__file__).resolve().parent / "golden" / "srtb_golden.npz")


def rel(a, b):
    a, b = np.asarray(a).astype(np.complex128).ravel(), np.asarray(b).astype(np.complex128).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


# ------------------------------------------------------------------ CPU: oracle vs golden
def test_oracle_unpack_golden(oracle):
    raw = G["unpack_raw"]
    for bits in (1, 2, 4, 8, -8, 16, -16):
        assert np.array_equal(oracle.unpack(raw, raw.size * 8 // abs(bits), bits), G[f"unpack_b{bits}"])
    assert np.array_equal(oracle.unpack(raw, raw.size, -8, 2), G["unpack_b-8_hamming"])
    assert np.array_equal(oracle.unpack(raw, raw.size * 4, 2, 1), G["unpack_b2_hann"])
    a, b = oracle.unpack_interleaved_2(raw, raw.size // 2, -8)
    assert np.array_equal(a, G["il2_a"]) and np.array_equal(b, G["il2_b"])
    a, b = oracle.unpack_snap1(raw, raw.size // 2)
    assert np.array_equal(a, G["snap1_a"]) and np.array_equal(b, G["snap1_b"])
    for s in (2, 4):
        for i, o in enumerate(oracle.unpack_gznupsr_a1(raw, raw.size // s, s)):
            assert np.array_equal(o, G[f"gznu{s}_{i}"])
    for w in (1, 2):
        assert np.array_equal(np.array([oracle.window(w, i, 16) for i in range(16)], np.float32), G[f"window{w}_16"])


def test_oracle_fft_golden(oracle):
    assert np.array_equal(oracle.fft_c2c(G["fft_x"], 1), G["fft_fwd"])
    assert np.array_equal(oracle.fft_c2c(G["fft_x"], -1), G["fft_bwd"])
    assert np.array_equal(oracle.fft_r2c(G["r2c_x"]), G["r2c_X"])
    assert np.array_equal(oracle.watfft(G["fft_x"], 32, 8), G["watfft_8x32"])


def test_oracle_stages_golden(oracle):
    y, _, _ = oracle.rfi_s1_average(G["s1_x"], 1.5, 16)
    y = oracle.rfi_manual(y, 1000.0, 500.0, oracle.eval_rfi_ranges("1100-1110, 1300.5-1302"))
    assert np.array_equal(y == 0, G["s1_y"] == 0) and np.array_equal(y, G["s1_y"])
    for tag in "abc":
        fl, bw, dm = G[f"dd_{tag}_params"]
        f_min, f_c = np.float32(fl), np.float32(np.float32(fl) + np.float32(bw))
        df = np.float32(np.float32(bw) / np.float32(2048))
        assert np.array_equal(oracle.dedisperse(G["dd_x"], float(f_min), float(f_c), float(df), float(dm)), G[f"dd_{tag}"])
    y, _, _ = oracle.rfi_s2(G["s2_x"], 256, 16, 1.2)
    assert np.array_equal(y, G["s2_y"])
    res, series = oracle.signal_detect(G["det_x"], 256, 16, 0, 6.0, 0.9, 32)
    got = {int(res.boxcar_length[b]): b for b in range(res.n_boxcars) if res.signal_count[b] > 0}
    assert sorted(got) == sorted(G["det_boxcar"].tolist())
    for bc, cnt, ln in zip(G["det_boxcar"], G["det_count"], G["det_length"]):
        b = got[int(bc)]
        assert res.signal_count[b] == cnt and res.series_length[b] == ln
        assert np.allclose(series[b, :ln], G[f"det_series_{bc}"], rtol=0, atol=1e-3)


# ------------------------------------------------------------------ GPU: CUDA path vs golden
@pytest.mark.gpu
def test_gpu_unpack_golden(ctx):
    import torch
    import srtb_b200
    raw = G["unpack_raw"]
    d = torch.from_numpy(raw).cuda()
    for bits in (1, 2, 4, 8, -8, 16, -16):
        n = raw.size * 8 // abs(bits)
        out = torch.zeros(n, dtype=torch.float32, device="cuda")
        ctx.unpack(d, raw.size, bits, 0, 0, [out], n)
        assert np.array_equal(out.cpu().numpy(), G[f"unpack_b{bits}"])
    out = torch.zeros(raw.size, dtype=torch.float32, device="cuda")
    ctx.unpack(d, raw.size, -8, 0, 2, [out], raw.size)
    assert np.allclose(out.cpu().numpy(), G["unpack_b-8_hamming"], rtol=1e-6, atol=1e-5)
    o = [torch.zeros(raw.size // 2, dtype=torch.float32, device="cuda") for _ in range(4)]
    ctx.unpack(d, raw.size, -8, srtb_b200.FORMAT_INTERLEAVED_2, 0, o[:2], raw.size // 2)
    assert np.array_equal(o[0].cpu().numpy(), G["il2_a"]) and np.array_equal(o[1].cpu().numpy(), G["il2_b"])
    ctx.unpack(d, raw.size, -8, srtb_b200.FORMAT_NAOCPSR_SNAP1, 0, o[:2], raw.size // 2)
    assert np.array_equal(o[0].cpu().numpy(), G["snap1_a"]) and np.array_equal(o[1].cpu().numpy(), G["snap1_b"])
    ctx.unpack(d, raw.size, 8, srtb_b200.FORMAT_GZNUPSR_A1_2, 0, o[:2], raw.size // 2)
    assert all(np.array_equal(o[i].cpu().numpy(), G[f"gznu2_{i}"]) for i in range(2))
    ctx.unpack(d, raw.size, 8, srtb_b200.FORMAT_GZNUPSR_A1_4, 0, o, raw.size // 4)
    assert all(np.array_equal(o[i].cpu().numpy()[:raw.size // 4], G[f"gznu4_{i}"]) for i in range(4))


@pytest.mark.gpu
def test_gpu_fft_golden(ctx):
    import torch
    for key, d in (("fft_fwd", 1), ("fft_bwd", -1)):
        x = torch.from_numpy(G["fft_x"]).cuda()
        ctx.fft_c2c(x, 256, 1, d)
        assert rel(x.cpu().numpy(), G[key]) < 1e-5       # the golden itself is the naive f32 radix-2
    buf = torch.zeros(514, dtype=torch.float32, device="cuda")
    buf[:512] = torch.from_numpy(G["r2c_x"]).cuda()
    ctx.fft_r2c_inplace(buf, 512)
    assert rel(buf.cpu().numpy().view(np.complex64), G["r2c_X"]) < 1e-5
    x = torch.from_numpy(G["fft_x"]).cuda()
    ctx.watfft_c2c_backward(x, 32, 8)
    assert rel(x.cpu().numpy(), G["watfft_8x32"]) < 1e-5


@pytest.mark.gpu
def test_gpu_stages_golden(ctx):
    import torch
    import srtb_b200
    x = torch.from_numpy(G["s1_x"]).cuda()
    bins = [srtb_b200.rfi_range_to_bins(a, b, 1000.0, 500.0, 2048) for a, b in
            srtb_b200.eval_rfi_ranges("1100-1110, 1300.5-1302")]
    ctx.rfi_s1(x, 2048, 1.5, srtb_b200.norm_coefficient(2048, 16), bins)
    got = x.cpu().numpy()
    assert np.array_equal(got == 0, G["s1_y"] == 0)
    assert rel(got, G["s1_y"]) < 1e-6
    for tag in "abc":
        fl, bw, dm = G[f"dd_{tag}_params"]
        f_min, f_c = np.float32(fl), np.float32(np.float32(fl) + np.float32(bw))
        df = np.float32(np.float32(bw) / np.float32(2048))
        x = torch.from_numpy(G["dd_x"]).cuda()
        ctx.dedisperse(x, 2048, float(f_min), float(f_c), float(df), float(dm))
        assert rel(x.cpu().numpy(), G[f"dd_{tag}"]) < 1e-5
    x = torch.from_numpy(G["s2_x"]).cuda()
    ctx.rfi_s2_sk(x, 256, 16, 1.2)
    assert np.array_equal(x.cpu().numpy(), G["s2_y"])
    h = np.zeros((srtb_b200.MAX_BOXCARS, 256), np.float32)
    res = ctx.signal_detect(torch.from_numpy(G["det_x"]).cuda(), 256, 16, 0, 6.0, 0.9, 32, h, copy_all=True)
    got = {int(res.boxcar_length[b]): b for b in range(res.n_boxcars) if res.signal_count[b] > 0}
    assert sorted(got) == sorted(G["det_boxcar"].tolist())
    for bc, cnt, ln in zip(G["det_boxcar"], G["det_count"], G["det_length"]):
        b = got[int(bc)]
        assert res.signal_count[b] == cnt and res.series_length[b] == ln
        gold = G[f"det_series_{bc}"]
        assert np.abs(h[b, :ln] - gold).max() < 2e-6 * np.abs(gold).max() * np.sqrt(float(bc)) + 1e-3


CHAIN = np.load(Path(__file__).resolve().parent / "golden" / "srtb_chain_golden.npz")


def _chain_params():
    n, C_, dm, f_low, bw, fs, avg_thr, sk_thr, snr, chan_thr, maxbox = CHAIN["params"]
    return int(n), int(C_), float(dm), float(f_low), float(bw), float(fs), float(avg_thr), float(sk_thr), float(snr), \
        float(chan_thr), int(maxbox)


def _borderline_by_boxcar(gold, snr, maxbox, border=1e-3):
    """number of boxcar-series elements within `border` (relative) of the detection threshold, from the
    reference's own dynamic spectrum in float64 (signal_detect_pipe.hpp:296-423): only those may flip a count"""
    ts = (np.abs(gold.astype(np.complex128)) ** 2).sum(axis=0)
    ts -= ts.mean()
    acc = np.cumsum(ts)
    out, b, series = {}, 1, ts
    while True:
        thr = snr * np.sqrt(np.mean(series ** 2))
        out[b] = int((np.abs(series - thr) < border * thr).sum())
        b *= 2
        if b > maxbox or b >= ts.size:
            return out
        series = acc[b:] - acc[:-b]


def _check_chain(spec, zero_count, counts_by_boxcar, lengths_by_boxcar):
    gold = CHAIN["spectrum"]
    gz, ez = np.all(spec == 0, axis=1), np.all(gold == 0, axis=1)
    assert np.array_equal(gz, ez)
    assert rel(spec[~gz], gold[~ez]) < 5e-6
    assert zero_count == int(CHAIN["zero_count"][0])
    params = _chain_params()
    borderline = _borderline_by_boxcar(gold, snr=params[8], maxbox=params[10])
    for bc, cnt, ln in zip(CHAIN["det_boxcar"], CHAIN["det_count"], CHAIN["det_length"]):
        assert lengths_by_boxcar[int(bc)] == int(ln)
        # counts equal the reference's except for elements sitting on the threshold (policy: SURVEY 8c)
        assert abs(counts_by_boxcar[int(bc)] - int(cnt)) <= borderline[int(bc)], (bc, borderline)


def test_oracle_chain_golden(oracle):
    """the oracle's one-call chain against the reference's own pipes composed as main.cpp wires them"""
    import ctypes as CT
    import oracle_lib
    n, C_, dm, f_low, bw, fs, avg_thr, sk_thr, snr, chan_thr, maxbox = _chain_params()
    cfg = oracle_lib.ChainConfig()
    cfg.baseband_input_count, cfg.baseband_input_bits, cfg.window = n, -8, 0
    cfg.baseband_freq_low, cfg.baseband_bandwidth, cfg.baseband_sample_rate, cfg.dm = f_low, bw, fs, dm
    cfg.rfi_average_threshold, cfg.rfi_sk_threshold, cfg.spectrum_channel_count = avg_thr, sk_thr, C_
    cfg.snr_threshold, cfg.channel_threshold, cfg.max_boxcar_length = snr, chan_thr, maxbox
    arr = (CT.c_float * 2)(*CHAIN["freq_pairs"].tolist())
    cfg.rfi_pairs, cfg.n_rfi_pairs = CT.cast(arr, CT.POINTER(CT.c_float)), 1
    work, res, _, _ = oracle.chain(CHAIN["raw"], cfg)
    L = n // 2 // C_
    _check_chain(work[:n].view(np.complex64).reshape(C_, L), int(res.zero_count),
                 {int(res.boxcar_length[b]): int(res.signal_count[b]) for b in range(res.n_boxcars)},
                 {int(res.boxcar_length[b]): int(res.series_length[b]) for b in range(res.n_boxcars)})


@pytest.mark.gpu
def test_gpu_chain_golden(ctx):
    """srtb_b200_process_block (fused kernels) against the same reference-generated fixture: the GPU box has no
    /root/reference, the committed vectors carry the reference's own outputs there"""
    import ctypes as CT
    import torch
    import srtb_b200
    n, C_, dm, f_low, bw, fs, avg_thr, sk_thr, snr, chan_thr, maxbox = _chain_params()
    cfg = srtb_b200.BlockConfig()
    cfg.baseband_input_count, cfg.baseband_input_bits, cfg.baseband_format, cfg.window = n, -8, srtb_b200.FORMAT_SIMPLE, 0
    cfg.baseband_freq_low, cfg.baseband_bandwidth, cfg.baseband_sample_rate, cfg.dm = f_low, bw, fs, dm
    cfg.mitigate_rfi_average_method_threshold, cfg.mitigate_rfi_spectral_kurtosis_threshold = avg_thr, sk_thr
    cfg.spectrum_channel_count = C_
    cfg.signal_detect_signal_noise_threshold, cfg.signal_detect_channel_threshold = snr, chan_thr
    cfg.signal_detect_max_boxcar_length = maxbox
    arr = (CT.c_float * 2)(*CHAIN["freq_pairs"].tolist())
    cfg.rfi_freq_pairs, cfg.n_rfi_freq_pairs = CT.cast(arr, CT.POINTER(CT.c_float)), 1
    res = ctx.process_block(cfg, torch.from_numpy(CHAIN["raw"].copy()).pin_memory(), n, None)[0]
    nc, L = n // 2, n // 2 // C_
    out = np.empty(nc, np.complex64)
    cudart = CT.CDLL("libcudart.so")
    cudart.cudaMemcpy.argtypes = [CT.c_void_p, CT.c_void_p, CT.c_size_t, CT.c_int]
    assert cudart.cudaMemcpy(out.ctypes.data, ctx.block_spectrum_ptr(0), out.nbytes, 2) == 0
    _check_chain(out.reshape(C_, L), int(res.zero_count),
                 {int(res.boxcar_length[b]): int(res.signal_count[b]) for b in range(res.n_boxcars)},
                 {int(res.boxcar_length[b]): int(res.series_length[b]) for b in range(res.n_boxcars)})

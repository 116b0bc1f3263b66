"""Pins the CPU oracle (oracle/srtb_oracle.cpp) against the reference's own operator/pipe headers,
compiled from /root/reference through the host shim (oracle/_ref/libsrtb_ref.so, see
oracle/ref_shim/README.md). Skipped when that library is absent. Same seeded inputs; bit-exact
wherever the reference's arithmetic order is fully specified, tolerance only where a reduction's
order is the runtime's choice."""
import numpy as np
import pytest

import ref_lib

ref = ref_lib.load()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref/libsrtb_ref.so not built (no /root/reference)")


def rel(a, b):
    a, b = np.asarray(a).astype(np.complex128).ravel(), np.asarray(b).astype(np.complex128).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("bits", [1, 2, 4, 8, -8, 16, -16, 32, 64])
def test_unpack_bit_exact(oracle, bits):
    rng = np.random.default_rng(abs(bits))
    nbytes = 1 << 12
    raw = (rng.standard_normal(nbytes // 4).astype(np.float32).view(np.uint8) if bits == 32 else
           rng.standard_normal(nbytes // 8).view(np.uint8) if bits == 64 else rng.integers(0, 256, nbytes, dtype=np.uint8))
    n = nbytes * 8 // abs(bits)
    assert np.array_equal(oracle.unpack(raw, n, bits), ref.unpack(raw, n, bits))
    if bits in (1, 2, 4):     # generic == handwritten, as test-unpack.cpp:211-254 checks
        assert np.array_equal(ref.unpack_handwritten(raw, n, bits), ref.unpack(raw, n, bits))


def test_unpack_multistream_bit_exact(oracle):
    rng = np.random.default_rng(9)
    raw = rng.integers(0, 256, 1 << 13, dtype=np.uint8)
    for bits in (8, -8, 16, -16):
        n = raw.size * 8 // abs(bits) // 2
        for a, b in zip(oracle.unpack_interleaved_2(raw, n, bits), ref.unpack_interleaved_2(raw, n, bits)):
            assert np.array_equal(a, b)
    n = raw.size // 2
    for a, b in zip(oracle.unpack_snap1(raw, n), ref.unpack_snap1(raw, n)):
        assert np.array_equal(a, b)
    for streams in (2, 4):
        n = raw.size // streams
        for a, b in zip(oracle.unpack_gznupsr_a1(raw, n, streams), ref.unpack_gznupsr_a1(raw, n, streams)):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("window", [0, 1, 2])
def test_window_bit_exact(oracle, window):
    for n in (16, 1000):
        assert [oracle.window(window, i, n) for i in range(n)] == [ref.window(window, i, n) for i in range(n)]
    raw = np.random.default_rng(3).integers(0, 256, 512, dtype=np.uint8)
    assert np.array_equal(oracle.unpack(raw, 512, -8, window), ref.unpack(raw, 512, -8, window))


@pytest.mark.parametrize("k", [1, 4, 10, 14])
def test_naive_fft_bit_exact(oracle, k):
    rng = np.random.default_rng(k)
    n = 1 << k
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
    for d in (1, -1):
        assert np.array_equal(oracle.fft_c2c(x, d), ref.fft_c2c(x, d))
    xr = rng.uniform(-1, 1, 2 * n).astype(np.float32)
    assert np.array_equal(oracle.fft_r2c(xr), ref.fft_r2c(xr))
    assert np.array_equal(oracle.watfft(x, n // 2 if n > 1 else 1, 2 if n > 1 else 1),
                          ref.watfft(x, n // 2 if n > 1 else 1, 2 if n > 1 else 1))


@pytest.mark.parametrize("nc,C_", [(1 << 10, 16), ((1 << 14) + 3, 64), (1 << 18, 2048)])
def test_rfi_s1_pipe(oracle, nc, C_):
    rng = np.random.default_rng(nc)
    x = ((rng.standard_normal(nc) + 1j * rng.standard_normal(nc)) * 100).astype(np.complex64)
    x[rng.integers(0, nc, 12)] *= 40
    thr = 1.5
    r = ref.rfi_s1_pipe(x, thr, C_, 1000.0, 500.0, "1100-1101, 1300.5-1302")
    o, mean, mask = oracle.rfi_s1_average(x, thr, C_)
    o = oracle.rfi_manual(o, 1000.0, 500.0, oracle.eval_rfi_ranges("1100-1101, 1300.5-1302"))
    p = np.abs(x.astype(np.complex128)) ** 2
    border = np.abs(p / (thr * p.mean()) - 1) < 1e-4       # the mean's summation order is the runtime's
    assert np.array_equal((r == 0)[~border], (o == 0)[~border])
    keep = (r != 0) & (o != 0)
    assert np.array_equal(r[keep], o[keep])                 # normalised values are bit-identical
    assert ((o == 0) & (x != 0)).sum() >= 10


def test_rfi_ranges_and_manual_zap(oracle):
    for s in ["11-12, 15-90, 233-235, 1176-1177", "", "1418-1422", "1-2-3, 5-6", " 7 - 8 ,9-10", "3-4,"]:
        assert oracle.eval_rfi_ranges(s) == ref.eval_rfi_ranges(s), s
    x = np.ones(1500, np.complex64)
    rr = ref.eval_rfi_ranges("11-12, 15-90, 233-235, 1176-1177")
    assert np.array_equal(oracle.rfi_manual(x, 0.0, 1499.0, rr), ref.rfi_manual(x, 0.0, 1499.0, rr))
    x = np.ones(1 << 12, np.complex64)
    for pairs, fl, bw in [([(1418.0, 1422.0)], 1437.0, -64.0), ([(1422.0, 1418.0)], 1437.0, -64.0),
                          ([(100.0, 200.0)], 1000.0, 500.0), ([(1400.0, 1600.0)], 1000.0, 500.0)]:
        assert np.array_equal(oracle.rfi_manual(x, fl, bw, pairs), ref.rfi_manual(x, fl, bw, pairs))


@pytest.mark.parametrize("nc,f_low,bw,dm", [(1 << 14, 1000.0, 500.0, 56.778), ((1 << 12) + 1, 1000.0, 400.0, 562.05),
                                            (1 << 16, 1437.0, -64.0, -478.80), (1 << 10, 1000.0, 500.0, 0.0)])
def test_dedisperse_pipe_bit_exact(oracle, nc, f_low, bw, dm):
    rng = np.random.default_rng(nc)
    x = (rng.standard_normal(nc) + 1j * rng.standard_normal(nc)).astype(np.complex64)
    r = ref.dedisperse_pipe(x, f_low, bw, dm)
    f_min, f_c = np.float32(f_low), np.float32(np.float32(f_low) + np.float32(bw))
    df = np.float32(np.float32(bw) / np.float32(nc))
    o = oracle.dedisperse(x, float(f_min), float(f_c), float(df), dm)
    assert np.array_equal(r, o)
    assert np.array_equal(ref.dedisperse(x, float(f_min), float(f_c), float(df), dm), o)


def test_nsamps_reserved(oracle):
    for args in [(1 << 26, 1 << 11, 1000.0, 500.0, 1e9, 5.0, True), (1 << 24, 1 << 11, 1000.0, 500.0, 1e9, 56.778, True),
                 (1 << 30, 1 << 11, 1437.0, -64.0, 128e6, -478.80, True), (1 << 30, 1 << 11, 1437.0, -64.0, 128e6, -478.80, False),
                 (1 << 28, 1 << 15, 1000.0, 500.0, 1e9, 100.0, True), (1 << 28, 1 << 15, 1000.0, 500.0, 1e9, 1.0, True)]:
        assert oracle.nsamps_reserved(*args) == ref.nsamps_reserved(*args), args


def test_rfi_s2_pipe(oracle):
    rng = np.random.default_rng(21)
    C_, L = 48, 1024
    x = (rng.standard_normal((C_, L)) + 1j * rng.standard_normal((C_, L))).astype(np.complex64)
    x[5, :] = 3
    x[9, ::8] *= 9
    x[11, :] = 0
    thr = 1.05
    r = ref.rfi_s2_pipe(x.reshape(-1), L, C_, thr).reshape(C_, L)
    o, sk, zap = oracle.rfi_s2(x.reshape(-1), L, C_, thr)
    o = o.reshape(C_, L)
    lo, hi = oracle.sk_thresholds(L, thr)
    fin = np.isfinite(sk)
    border = np.zeros(C_, bool)
    border[fin] = (np.abs(sk[fin] / hi - 1) < 1e-4) | (np.abs(sk[fin] / lo - 1) < 1e-4)
    rz = np.all(r == 0, axis=1)
    oz = np.all(o == 0, axis=1)
    assert np.array_equal(rz[~border], oz[~border])
    assert rz[5] and rz[9] and rz[11] and zap[11] == 0     # q5: the all-zero row was left alone, not "zapped"
    same = rz == oz
    assert np.array_equal(r[same], o[same])


@pytest.mark.parametrize("C_,L,maxbox,reserve", [(16, 256, 16, False), (64, 2048, 256, False), (32, 1000, 64, True)])
def test_signal_detect_pipe(oracle, C_, L, maxbox, reserve):
    rng = np.random.default_rng(C_ * L)
    x = (rng.standard_normal((C_, L)) + 1j * rng.standard_normal((C_, L))).astype(np.complex64)
    x[:, L // 8:L // 8 + 8] *= 10
    x[1, :] = 0
    n_input = 2 * C_ * L
    f_low, bw, fs, dm, snr, chan_thr = 1000.0, 500.0, 1e9, 0.005, 6.0, 0.9
    holders = ref.signal_detect_pipe(x.reshape(-1), L, C_, n_input, reserve, f_low, bw, fs, dm, snr, chan_thr, maxbox)
    reserved = oracle.nsamps_reserved(n_input, C_, f_low, bw, fs, dm, reserve) // C_
    if reserve:
        assert reserved > 0
    res, series = oracle.signal_detect(x.reshape(-1), L, C_, reserved, snr, chan_thr, maxbox)
    assert res.detect_enabled == 1
    got = {h["boxcar"]: h for h in holders}
    exp = {int(res.boxcar_length[b]): b for b in range(res.n_boxcars) if res.signal_count[b] > 0}
    # every series the reference emits exists in the oracle with (near-)identical values and counts
    for bc, h in got.items():
        b = [i for i in range(res.n_boxcars) if res.boxcar_length[i] == bc][0]
        assert h["length"] == res.series_length[b]
        scale = np.sqrt(np.mean(h["series"].astype(np.float64) ** 2))
        assert np.abs(h["series"] - series[b, :h["length"]]).max() < 1e-4 * scale * np.sqrt(bc)
        assert abs(h["count"] - int(res.signal_count[b])) <= 1
    assert set(exp) - set(got) <= {bc for bc, b in exp.items() if res.signal_count[b] <= 1}
    assert len(got) > 0
    assert ref.count_signal(series[0, :int(res.series_length[0])], snr) == res.signal_count[0]


@pytest.mark.parametrize("logn,C_,dm,bits", [(14, 16, 0.0, -8), (15, 32, 0.02, -8), (14, 8, 0.0, 2)])
def test_whole_chain_composition(oracle, logn, C_, dm, bits):
    """The reference's pipes composed as main.cpp:170-204 wires them (unpack -> R2C -> drop Nyquist -> s1 pipe ->
    dedisperse pipe -> waterfall FFT -> s2 pipe -> signal_detect_pipe_2), each stage being the reference's OWN code
    through the shim, against the oracle's one-call chain: pins sizes, the Nyquist drop, the [C][L] row layout, the
    normalisation coefficient and the detector's view of the spectrum — not only the stages in isolation."""
    import oracle_lib
    n = 1 << logn
    nc, L = n // 2, n // 2 // C_
    rng = np.random.default_rng(logn * 7 + C_)
    if bits == -8:
        v = np.clip(np.round(rng.standard_normal(n) * 20), -127, 127)
        v[n // 2:n // 2 + 32] += np.round(rng.standard_normal(32) * 90)
        raw = np.clip(v, -127, 127).astype(np.int8).view(np.uint8)
    else:
        raw = rng.integers(0, 256, n * bits // 8, dtype=np.uint8)
    f_low, bw, fs = 1000.0, 500.0, 1e9
    avg_thr, sk_thr, snr, chan_thr, maxbox = 5.0, 1.3, 6.0, 0.9, 32
    # --- the reference, stage by stage
    x = ref.unpack(raw, n, bits)
    spec = ref.fft_r2c(x)[:nc]                                   # fft_pipe.hpp:75-77: count = N/2
    spec = ref.rfi_s1_pipe(spec, avg_thr, C_, f_low, bw, "1200-1201")
    spec = ref.dedisperse_pipe(spec, f_low, bw, dm)
    spec = ref.watfft(spec, L, C_)
    spec = ref.rfi_s2_pipe(spec, L, C_, sk_thr).reshape(C_, L)
    holders = ref.signal_detect_pipe(spec.reshape(-1), L, C_, n, False, f_low, bw, fs, dm, snr, chan_thr, maxbox)
    # --- the oracle chain
    import ctypes as CT
    cfg = oracle_lib.ChainConfig()
    cfg.baseband_input_count, cfg.baseband_input_bits, cfg.window = n, bits, 0
    cfg.baseband_freq_low, cfg.baseband_bandwidth, cfg.baseband_sample_rate, cfg.dm = f_low, bw, fs, dm
    cfg.baseband_reserve_sample = 0
    cfg.rfi_average_threshold, cfg.rfi_sk_threshold = avg_thr, sk_thr
    cfg.spectrum_channel_count = C_
    cfg.snr_threshold, cfg.channel_threshold, cfg.max_boxcar_length = snr, chan_thr, maxbox
    arr = (CT.c_float * 2)(1200.0, 1201.0)
    cfg.rfi_pairs, cfg.n_rfi_pairs = CT.cast(arr, CT.POINTER(CT.c_float)), 1
    work, res, series, _ = oracle.chain(raw, cfg)
    ospec = work[:n].view(np.complex64).reshape(C_, L)
    rz, oz = np.all(spec == 0, axis=1), np.all(ospec == 0, axis=1)
    assert (rz != oz).sum() <= 1                                  # SK border only
    same = rz == oz
    assert rel(ospec[same], spec[same]) < 2e-7                    # same arithmetic; reduction order is the runtime's
    if np.array_equal(rz, oz):
        assert res.zero_count == int(np.sum(np.abs(spec[:, 0]) ** 2 == 0))
        got = {h["boxcar"]: h for h in holders}
        for b in range(res.n_boxcars):
            bc = int(res.boxcar_length[b])
            if res.signal_count[b] > 1:
                assert bc in got and abs(got[bc]["count"] - int(res.signal_count[b])) <= 1
        for bc, h in got.items():
            b = [i for i in range(res.n_boxcars) if res.boxcar_length[i] == bc][0]
            assert h["length"] == res.series_length[b]


def _refft_layout_block(rng, nt, nf):
    """spectra [time][frequency]: noise, one steady tone (low kurtosis), one bursty channel (high kurtosis), one
    manually zapped channel, and a broadband burst over a few consecutive spectra"""
    x = (rng.standard_normal((nt, nf)) + 1j * rng.standard_normal((nt, nf))).astype(np.complex64)
    x[:, 5] = 3
    x[::8, 9] *= 9
    x[:, 11] = 0
    x[nt // 3:nt // 3 + 16, :] *= 1.6   # mild enough to stay inside the kurtosis thresholds of the detector tests
    return x


@pytest.mark.parametrize("nt,nf", [(256, 64), (1000, 48)])
def test_sk_v1(oracle, nt, nf):
    """alternate f-4: mitigate_rfi_spectral_kurtosis_method (v1) on the [time][frequency] layout of the refft path"""
    x = _refft_layout_block(np.random.default_rng(nt + nf), nt, nf)
    thr = 1.1
    r = ref.sk_v1(x.reshape(-1), nf, nt, thr).reshape(nt, nf)
    o, sk, zap = oracle.sk_v1(x.reshape(-1), nf, nt, thr)
    o = o.reshape(nt, nf)
    lo, hi = oracle.sk_thresholds(nt, thr)
    fin = np.isfinite(sk)
    border = np.zeros(nf, bool)
    border[fin] = (np.abs(sk[fin] / hi - 1) < 1e-4) | (np.abs(sk[fin] / lo - 1) < 1e-4)
    rz, oz = np.all(r == 0, axis=0), np.all(o == 0, axis=0)
    assert np.array_equal(rz[~border], oz[~border])
    assert rz[5] and rz[9] and rz[11] and zap[11] == 0      # the all-zero channel is NaN -> left alone
    same = rz == oz
    assert np.array_equal(r[:, same], o[:, same])            # same serial fp32 sums: identical where decisions agree


@pytest.mark.parametrize("nt,nf,maxbox", [(512, 64, 32), (1000, 128, 256)])
def test_signal_detect_pipe_v1(oracle, nt, nf, maxbox):
    """alternate f-4: the reference's signal_detect_pipe (v1: SK v1 + per-spectrum sums + boxcars) against the oracle.
    The spectrum length is a multiple of the device's work-group size here: multi_mapreduce cuts the flat array into
    pieces of ceil(size / (items * groups)) * items elements (algorithm/multi_reduce.hpp:93-101), which are the rows
    only then; the oracle restates the intended per-spectrum sum."""
    x = _refft_layout_block(np.random.default_rng(7 * nt + nf), nt, nf)
    sk_thr, snr, chan_thr = 1.4, 5.0, 0.9
    rspec, holders = ref.signal_detect_pipe_v1(x.reshape(-1), nf, nt, sk_thr, snr, chan_thr, maxbox)
    ospec, res, series = oracle.signal_detect_v1(x.reshape(-1), nf, nt, sk_thr, snr, chan_thr, maxbox)
    assert np.array_equal(rspec.reshape(nt, nf), ospec.reshape(nt, nf))
    assert res.detect_enabled == 1 and res.zero_count == int(np.sum(np.abs(ospec.reshape(nt, nf)[0]) == 0))
    got = {h["boxcar"]: h for h in holders}
    exp = {int(res.boxcar_length[b]): b for b in range(res.n_boxcars) if res.signal_count[b] > 0}
    assert len(got) > 0
    for bc, h in got.items():
        b = [i for i in range(res.n_boxcars) if res.boxcar_length[i] == bc][0]
        assert h["length"] == res.series_length[b]
        if bc == 1:   # (the reference re-uses one device buffer for every boxcar > 1 and copies it asynchronously: SURVEY q3)
            scale = np.sqrt(np.mean(h["series"].astype(np.float64) ** 2))
            assert np.abs(h["series"] - series[b, :h["length"]]).max() < 1e-4 * scale
            assert abs(h["count"] - int(res.signal_count[b])) <= 1
    assert set(exp) - set(got) <= {bc for bc, b in exp.items() if res.signal_count[b] <= 1}

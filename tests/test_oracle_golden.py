"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for
this path, and against float64 truth where the reference only cross-checks.
CPU only."""
import numpy as np
import pytest

KAT_BYTES = np.array([0b01100011, 0b10110110, 0b00001000, 0b10011101], np.uint8)


# --- userspace/tests/test-unpack.cpp:62-140 (single byte) --------------------------------
@pytest.mark.parametrize("bits,byte,expected", [
    (1, 0b01100011, [0, 1, 1, 0, 0, 0, 1, 1]),
    (2, 0b10110110, [2, 3, 1, 2]),
    (4, 0b00001000, [0, 8]),
    (8, 0b10011101, [157]),
])
def test_unpack_single_byte_kat(oracle, bits, byte, expected):
    out = oracle.unpack(np.array([byte], np.uint8), 8 // bits, bits)
    assert out.tolist() == [float(v) for v in expected]


# --- userspace/tests/test-unpack.cpp:142-210 (4 bytes through the kernel) -------------------
@pytest.mark.parametrize("bits,expected", [
    (1, [0, 1, 1, 0, 0, 0, 1, 1, 1, 0, 1, 1, 0, 1, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 1, 1, 1, 0, 1]),
    (2, [1, 2, 0, 3, 2, 3, 1, 2, 0, 0, 2, 0, 2, 1, 3, 1]),
    (4, [6, 3, 11, 6, 0, 8, 9, 13]),
    (8, [99, 182, 8, 157]),
])
def test_unpack_four_byte_kat(oracle, bits, expected):
    out = oracle.unpack(KAT_BYTES, 32 // bits, bits)
    assert out.tolist() == [float(v) for v in expected]


def test_unpack_signed_and_wide(oracle):
    rng = np.random.default_rng(1)
    raw = rng.integers(0, 256, 64, dtype=np.uint8)
    assert np.array_equal(oracle.unpack(raw, 64, -8), raw.view(np.int8).astype(np.float32))
    assert np.array_equal(oracle.unpack(raw, 32, 16), raw.view(np.uint16).astype(np.float32))
    assert np.array_equal(oracle.unpack(raw, 32, -16), raw.view(np.int16).astype(np.float32))
    f = rng.standard_normal(16).astype(np.float32)
    assert np.array_equal(oracle.unpack(f.view(np.uint8), 16, 32), f)
    d = rng.standard_normal(16)
    assert np.array_equal(oracle.unpack(d.view(np.uint8), 16, 64), d.astype(np.float32))
    with pytest.raises(ValueError):
        oracle.unpack(raw, 8, 3)  # "[unpack pipe] unsupported baseband_input_bits" (unpack_pipe.hpp:123-127)


def test_unpack_multistream_layouts(oracle):
    raw = np.arange(32, dtype=np.uint8)
    o1, o2 = oracle.unpack_interleaved_2(raw, 16, 8)          # "1 2 1 2" unpack.hpp:221-229
    assert o1.tolist() == list(range(0, 32, 2)) and o2.tolist() == list(range(1, 32, 2))
    s1, s2 = oracle.unpack_snap1(raw.view(np.int8), 16)       # "1 1 2 2" unpack.hpp:255-268
    assert s1[:4].tolist() == [0, 1, 4, 5] and s2[:4].tolist() == [2, 3, 6, 7]
    g = oracle.unpack_gznupsr_a1(raw.view(np.int8), 16, 2)     # words of 4, 2 streams unpack.hpp:338-369
    assert g[0][:8].tolist() == [0, 1, 2, 3, 8, 9, 10, 11] and g[1][:8].tolist() == [4, 5, 6, 7, 12, 13, 14, 15]
    raw4 = np.array([0, 127, 128, 255] * 4 + list(range(16, 64)), np.uint8).view(np.int8)
    g4 = oracle.unpack_gznupsr_a1(raw4, 16, 4)                 # int(int8) ^ 0x80, unpack.hpp:315-316
    assert g4[0][:4].tolist() == [128.0, 255.0, -256.0, -129.0]


# --- userspace/tests/test-fft_window.cpp:39-48 (numpy.hamming(16), a0 = 25/46) ----------------
HAMMING16 = [0.08, 0.11976909, 0.23219992, 0.39785218, 0.58808309, 0.77, 0.91214782, 0.9899479,
             0.9899479, 0.91214782, 0.77, 0.58808309, 0.39785218, 0.23219992, 0.11976909, 0.08]


def test_hamming_window_kat(oracle):
    expected2 = [(25.0 / 46.0) - ((0.54 - e) / 0.46) * (21.0 / 46.0) for e in HAMMING16]
    got = [oracle.window(2, i, 16) for i in range(16)]
    assert np.allclose(got, expected2, atol=1e-6)       # the test's own threshold
    # window fused into unpack<1> of all-ones bytes (test-fft_window.cpp:98-121)
    out = oracle.unpack(np.array([0xFF, 0xFF], np.uint8), 16, 1, window=2)
    assert np.allclose(out, expected2, atol=1e-6)
    assert [oracle.window(0, i, 16) for i in range(16)] == [1.0] * 16
    hann = [oracle.window(1, i, 16) for i in range(16)]
    assert np.allclose(hann, np.hanning(16), atol=1e-6)


# --- userspace/tests/test-rfi_mitigation.cpp:21-71 -------------------------------------------
def test_manual_rfi_zap_kat(oracle):
    ranges = oracle.eval_rfi_ranges("11-12, 15-90, 233-235, 1176-1177")
    assert ranges == [(11.0, 12.0), (15.0, 90.0), (233.0, 235.0), (1176.0, 1177.0)]
    n = 1500
    x = np.ones(n, np.complex64)
    y = oracle.rfi_manual(x, 0.0, float(n - 1), ranges)
    expected = np.ones(n, np.complex64)
    for a, b in ranges:
        expected[int(a):int(b) + 1] = 0
    assert np.array_equal(y, expected)


def test_manual_rfi_signed_band_and_bounds(oracle):
    # inverted band of the J1644 cfg (srtb_config_1644-4559.cfg:24-29): low 1437, bw -64
    n = 1 << 12
    got = oracle.rfi_range_to_bins(1418.0, 1422.0, 1437.0, -64.0, n)
    lo = round((1422.0 - 1437.0) / -64.0 * (n - 1))
    hi = round((1418.0 - 1437.0) / -64.0 * (n - 1))
    assert got == (lo, hi)
    assert oracle.rfi_range_to_bins(100.0, 200.0, 1000.0, 500.0, n) is None      # below the band
    assert oracle.rfi_range_to_bins(1400.0, 1600.0, 1000.0, 500.0, n) is None    # above the band
    assert oracle.eval_rfi_ranges("") == []
    assert oracle.eval_rfi_ranges("1-2-3, 5-6") == [(5.0, 6.0)]


# --- FFT: the reference pins naive == FFTW (test-naive_fft.cpp:148-157) and dispatcher == FFTW
#     with tolerance clamp(eps*n/2, 1e-5, 0.05) (test-fft_wrappers.cpp:109-110). Truth = float64.
@pytest.mark.parametrize("k", [2, 5, 12])
def test_naive_c2c_vs_float64(oracle, k):
    rng = np.random.default_rng(k)
    n = 1 << k
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
    truth_f = np.fft.fft(x.astype(np.complex128))
    truth_b = np.fft.ifft(x.astype(np.complex128)) * n
    f = oracle.fft_c2c(x, +1)
    b = oracle.fft_c2c(x, -1)
    tol = float(np.clip(np.finfo(np.float32).eps * n / 2, 1e-5, 0.05))
    assert np.linalg.norm(f - truth_f) / np.linalg.norm(truth_f) < tol
    assert np.linalg.norm(b - truth_b) / np.linalg.norm(truth_b) < tol


@pytest.mark.parametrize("k,batch", [(2, 8), (2, 16), (12, 2), (16, 1)])
def test_naive_r2c_vs_float64(oracle, k, batch):
    # sizes follow userspace/tests/CMakeLists.txt:22-32 (bits 2 and 21 there; 12/16 here for time)
    rng = np.random.default_rng(233)
    n = 1 << k
    for _ in range(batch):
        x = rng.uniform(-1, 1, n).astype(np.float32)
        got = oracle.fft_r2c(x)
        truth = np.fft.rfft(x.astype(np.float64))
        tol = float(np.clip(np.finfo(np.float32).eps * n / 2, 1e-5, 0.05))
        assert got.size == n // 2 + 1
        assert np.linalg.norm(got - truth) / np.linalg.norm(truth) < tol


def test_watfft_layout(oracle):
    rng = np.random.default_rng(3)
    C_, L = 8, 64
    x = (rng.standard_normal(C_ * L) + 1j * rng.standard_normal(C_ * L)).astype(np.complex64)
    y = oracle.watfft(x, L, C_).reshape(C_, L)
    truth = np.fft.ifft(x.astype(np.complex128).reshape(C_, L), axis=1) * L
    assert np.linalg.norm(y - truth) / np.linalg.norm(truth) < 1e-5


# --- chirp: parameter set of userspace/tests/test-df64.cpp:30-33 vs float64/longdouble truth --
def test_dedisperse_against_truth(oracle):
    n = 1 << 14
    f_min, f_max, dm = np.float32(1000), np.float32(1500), np.float32(56.778)
    df = np.float32((f_max - f_min) / np.float32(n))
    x = np.ones(n, np.complex64)
    y = oracle.dedisperse(x, f_min, f_max, df, dm)
    i = np.arange(n, dtype=np.longdouble)
    f = np.longdouble(f_min) + np.longdouble(df) * i
    k = np.longdouble(4.148808e3) * 1e6 * np.longdouble(dm) / f * ((f - np.longdouble(f_max)) / np.longdouble(f_max)) ** 2
    frac = (k - np.trunc(k)).astype(np.float64)
    truth = np.exp(-2j * np.pi * frac)
    assert np.abs(y - truth).max() < 2e-5   # |k| ~ 1e7 here: fp64 rounding of k dominates
    assert np.allclose(np.abs(y), 1.0, atol=1e-6)


def test_norm_coefficient_and_sk_thresholds(oracle):
    # SURVEY §8 anchors: config #3 coef = 2^-19.5, SK window for t=1.05, M=2^14
    assert oracle.norm_coefficient(1 << 25, 1 << 11) == pytest.approx(2.0 ** -19.5, rel=1e-6)
    lo, hi = oracle.sk_thresholds(1 << 14, 1.05)
    assert lo == pytest.approx(1.94988, abs=2e-5) and hi == pytest.approx(2.04987, abs=2e-5)


def test_nsamps_reserved(oracle):
    N = 1 << 26
    assert oracle.nsamps_reserved(N, 1 << 11, 1000.0, 500.0, 1e9, 5.0, False) == 0
    r = oracle.nsamps_reserved(N, 1 << 11, 1000.0, 500.0, 1e9, 5.0, True)
    dt = 4.148808e3 * 5.0 * (1 / 1000.0 ** 2 - 1 / 1500.0 ** 2)
    minimal = 2 * round(dt * 1e9)
    keep = ((N - minimal) // (1 << 12)) * (1 << 12)
    assert abs(r - (N - keep)) <= (1 << 12)   # f32 rounding of dt*fs can move one bin
    assert r >= minimal - 64 and (N - r) % (1 << 12) == 0
    # block shorter than the smear: overlap disabled (coherent_dedispersion.hpp:118-127)
    assert oracle.nsamps_reserved(1 << 24, 1 << 11, 1000.0, 500.0, 1e9, 56.778, True) == 0


def test_signal_detect_semantics(oracle):
    rng = np.random.default_rng(7)
    C_, L = 16, 256
    x = (rng.standard_normal((C_, L)) + 1j * rng.standard_normal((C_, L))).astype(np.complex64)
    x[:, 100:104] *= 12.0                      # a 4-sample pulse in every channel
    x[3, :] = 0                                # one zapped channel
    res, series = oracle.signal_detect(x.reshape(-1), L, C_, 0, 6.0, 0.9, 16)
    assert res.zero_count == 1 and res.time_series_count == L and res.detect_enabled == 1
    assert res.n_boxcars == 5 and list(res.boxcar_length[:5]) == [1, 2, 4, 8, 16]
    assert list(res.series_length[:5]) == [L, L - 2, L - 4, L - 8, L - 16]
    ts = (np.abs(x.astype(np.complex128)) ** 2).sum(0)
    ts -= ts.mean()
    assert np.allclose(series[0, :L], ts, rtol=1e-4, atol=1e-3)
    acc = np.cumsum(ts)
    assert np.allclose(series[2, :L - 4], acc[4:] - acc[:-4], rtol=1e-4, atol=2e-2)
    for b, v in ((0, ts), (2, acc[4:] - acc[:-4])):
        thr = 6.0 * np.sqrt(np.mean(v ** 2))
        assert res.threshold[b] == pytest.approx(thr, rel=1e-4)
        assert res.signal_count[b] == int((v > thr).sum()) and res.signal_count[b] >= 1
    # too many zapped channels -> detection disabled (signal_detect_pipe.hpp:344-345)
    x[:15, :] = 0
    res2, _ = oracle.signal_detect(x.reshape(-1), L, C_, 0, 6.0, 0.9, 16)
    assert res2.zero_count == 15 and res2.detect_enabled == 0 and res2.n_boxcars == 0

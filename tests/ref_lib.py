"""ctypes access to oracle/_ref/libsrtb_ref.so — the REFERENCE'S OWN headers compiled through the
host shim (oracle/ref_shim/README.md). Test infrastructure only. `load()` returns None when the
library is absent (e.g. no /root/reference to build it from)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "_ref" / "libsrtb_ref.so"
MAXS = 32


class Ref:
    def __init__(self):
        self.lib = C.CDLL(str(LIB))
        L = self.lib
        P, SZ, F, I = C.c_void_p, C.c_size_t, C.c_float, C.c_int
        L.srtb_ref_unpack.argtypes = [P, SZ, I, I, P]
        L.srtb_ref_unpack_handwritten.argtypes = [P, SZ, I, P]
        L.srtb_ref_unpack_interleaved_2.argtypes = [P, SZ, I, I, P, P]
        L.srtb_ref_unpack_snap1.argtypes = [P, SZ, I, P, P]
        L.srtb_ref_unpack_gznupsr_a1.argtypes = [P, SZ, I, I, C.POINTER(P)]
        L.srtb_ref_window.restype = F
        L.srtb_ref_window.argtypes = [I, SZ, SZ]
        L.srtb_ref_fft_c2c.argtypes = [P, SZ, I]
        L.srtb_ref_fft_r2c.argtypes = [P, SZ]
        L.srtb_ref_watfft.argtypes = [P, SZ, SZ]
        L.srtb_ref_rfi_s1_pipe.argtypes = [P, SZ, F, SZ, F, F, C.c_char_p]
        L.srtb_ref_eval_rfi_ranges.restype = SZ
        L.srtb_ref_eval_rfi_ranges.argtypes = [C.c_char_p, P, SZ]
        L.srtb_ref_rfi_manual.argtypes = [P, SZ, F, F, P, SZ]
        L.srtb_ref_dedisperse_pipe.argtypes = [P, SZ, F, F, F]
        L.srtb_ref_dedisperse.argtypes = [P, SZ, F, F, F, F]
        L.srtb_ref_nsamps_reserved.restype = SZ
        L.srtb_ref_nsamps_reserved.argtypes = [SZ, SZ, F, F, F, F, I]
        L.srtb_ref_rfi_s2_pipe.argtypes = [P, SZ, SZ, F]
        L.srtb_ref_signal_detect_pipe.argtypes = [P, SZ, SZ, SZ, I, F, F, F, F, F, F, SZ, P, P, P, P, I]
        L.srtb_ref_count_signal.restype = C.c_ulonglong
        L.srtb_ref_count_signal.argtypes = [P, SZ, F]

    def unpack(self, raw, out_count, bits, window=0):
        raw = np.ascontiguousarray(raw)
        out = np.zeros(out_count, np.float32)
        rc = self.lib.srtb_ref_unpack(raw.ctypes.data, out_count, bits, window, out.ctypes.data)
        if rc != 0:
            raise ValueError(bits)
        return out

    def unpack_handwritten(self, raw, out_count, bits):
        raw = np.ascontiguousarray(raw)
        out = np.zeros(out_count, np.float32)
        assert self.lib.srtb_ref_unpack_handwritten(raw.ctypes.data, out_count, bits, out.ctypes.data) == 0
        return out

    def unpack_interleaved_2(self, raw, out_count, bits, window=0):
        raw = np.ascontiguousarray(raw)
        o1, o2 = np.zeros(out_count, np.float32), np.zeros(out_count, np.float32)
        assert self.lib.srtb_ref_unpack_interleaved_2(raw.ctypes.data, out_count, bits, window, o1.ctypes.data,
                                                      o2.ctypes.data) == 0
        return o1, o2

    def unpack_snap1(self, raw, out_count, window=0):
        raw = np.ascontiguousarray(raw)
        o1, o2 = np.zeros(out_count, np.float32), np.zeros(out_count, np.float32)
        self.lib.srtb_ref_unpack_snap1(raw.ctypes.data, out_count, window, o1.ctypes.data, o2.ctypes.data)
        return o1, o2

    def unpack_gznupsr_a1(self, raw, out_count, streams, window=0):
        raw = np.ascontiguousarray(raw)
        outs = [np.zeros(out_count, np.float32) for _ in range(streams)]
        arr = (C.c_void_p * streams)(*[o.ctypes.data for o in outs])
        assert self.lib.srtb_ref_unpack_gznupsr_a1(raw.ctypes.data, out_count, streams, window, arr) == 0
        return outs

    def window(self, window, i, n):
        return float(self.lib.srtb_ref_window(window, i, n))

    def fft_c2c(self, x, direction):
        y = np.ascontiguousarray(x, dtype=np.complex64).copy()
        self.lib.srtb_ref_fft_c2c(y.ctypes.data, y.size, direction)
        return y

    def fft_r2c(self, x):
        n = x.size
        buf = np.zeros(n + 2, np.float32)
        buf[:n] = x
        self.lib.srtb_ref_fft_r2c(buf.ctypes.data, n)
        return buf.view(np.complex64).copy()

    def watfft(self, x, length, batch):
        y = np.ascontiguousarray(x, dtype=np.complex64).copy()
        self.lib.srtb_ref_watfft(y.ctypes.data, length, batch)
        return y

    def rfi_s1_pipe(self, x, threshold, channel_count, freq_low, bandwidth, freq_list=""):
        y = np.ascontiguousarray(x, dtype=np.complex64).copy()
        self.lib.srtb_ref_rfi_s1_pipe(y.ctypes.data, y.size, threshold, channel_count, freq_low, bandwidth,
                                      freq_list.encode())
        return y

    def eval_rfi_ranges(self, s):
        buf = np.zeros(128, np.float32)
        n = self.lib.srtb_ref_eval_rfi_ranges(s.encode(), buf.ctypes.data, 64)
        return [(float(buf[2 * i]), float(buf[2 * i + 1])) for i in range(min(n, 64))]

    def rfi_manual(self, x, freq_low, bandwidth, pairs):
        y = np.ascontiguousarray(x, dtype=np.complex64).copy()
        p = np.asarray(pairs, np.float32).reshape(-1)
        self.lib.srtb_ref_rfi_manual(y.ctypes.data, y.size, freq_low, bandwidth, p.ctypes.data, p.size // 2)
        return y

    def dedisperse_pipe(self, x, freq_low, bandwidth, dm):
        y = np.ascontiguousarray(x, dtype=np.complex64).copy()
        self.lib.srtb_ref_dedisperse_pipe(y.ctypes.data, y.size, freq_low, bandwidth, dm)
        return y

    def dedisperse(self, x, f_min, f_c, df, dm):
        y = np.ascontiguousarray(x, dtype=np.complex64).copy()
        self.lib.srtb_ref_dedisperse(y.ctypes.data, y.size, f_min, f_c, df, dm)
        return y

    def nsamps_reserved(self, n, c, freq_low, bw, fs, dm, reserve):
        return int(self.lib.srtb_ref_nsamps_reserved(n, c, freq_low, bw, fs, dm, int(reserve)))

    def rfi_s2_pipe(self, x, time_count, chan_count, thr):
        y = np.ascontiguousarray(x, dtype=np.complex64).copy()
        self.lib.srtb_ref_rfi_s2_pipe(y.ctypes.data, time_count, chan_count, thr)
        return y

    def signal_detect_pipe(self, x, time_count, chan_count, n_input, reserve, freq_low, bw, fs, dm, snr, chan_thr,
                           max_boxcar):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        bl = np.zeros(MAXS, np.uint64)
        sl = np.zeros(MAXS, np.uint64)
        sc = np.zeros(MAXS, np.uint64)
        series = np.zeros((MAXS, time_count), np.float32)
        n = self.lib.srtb_ref_signal_detect_pipe(x.ctypes.data, time_count, chan_count, n_input, int(reserve),
                                                 freq_low, bw, fs, dm, snr, chan_thr, max_boxcar, bl.ctypes.data,
                                                 sl.ctypes.data, sc.ctypes.data, series.ctypes.data, MAXS)
        return [dict(boxcar=int(bl[i]), length=int(sl[i]), count=int(sc[i]), series=series[i, :int(sl[i])].copy())
                for i in range(n)]

    def sk_v1(self, x, fft_bins, time_counts, thr):
        y = np.ascontiguousarray(x, dtype=np.complex64).copy()
        self.lib.srtb_ref_sk_v1(C.c_void_p(y.ctypes.data), C.c_size_t(fft_bins), C.c_size_t(time_counts), C.c_float(thr))
        return y

    def signal_detect_pipe_v1(self, x, count_per_batch, batch_size, sk_thr, snr, chan_thr, max_boxcar):
        """the reference's signal_detect_pipe (v1): returns (spectrum after its SK v1, holders)"""
        y = np.ascontiguousarray(x, dtype=np.complex64).copy()
        bl = np.zeros(MAXS, np.uint64)
        sl = np.zeros(MAXS, np.uint64)
        sc = np.zeros(MAXS, np.uint64)
        series = np.zeros((MAXS, batch_size), np.float32)
        self.lib.srtb_ref_signal_detect_pipe_v1.restype = C.c_int
        n = self.lib.srtb_ref_signal_detect_pipe_v1(
            C.c_void_p(y.ctypes.data), C.c_size_t(count_per_batch), C.c_size_t(batch_size), C.c_float(sk_thr),
            C.c_float(snr), C.c_float(chan_thr), C.c_size_t(max_boxcar), C.c_void_p(bl.ctypes.data),
            C.c_void_p(sl.ctypes.data), C.c_void_p(sc.ctypes.data), C.c_void_p(series.ctypes.data), C.c_int(MAXS))
        return y, [dict(boxcar=int(bl[i]), length=int(sl[i]), count=int(sc[i]), series=series[i, :int(sl[i])].copy())
                   for i in range(n)]

    def count_signal(self, v, snr):
        v = np.ascontiguousarray(v, dtype=np.float32)
        return int(self.lib.srtb_ref_count_signal(v.ctypes.data, v.size, snr))


_inst = None


def load():
    global _inst
    if _inst is None and LIB.exists():
        _inst = Ref()
    return _inst

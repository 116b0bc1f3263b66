"""The reference's main.cpp wiring on the re-hosted pipes (tests/cpp/pipeline_main.cpp): one thread per
pipe + work queues (drop-in mode), and the same stages as one stream-ordered composite_pipe. Results
are compared with the CPU oracle block by block."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "tests" / "cpp" / "pipeline_main"


def _build():
    subprocess.run(["make", "-C", str(BIN.parent), "pipeline_main"], check=True, capture_output=True)
    assert BIN.exists(), "pipeline_main was not built (libsrtb_b200.so missing?)"


def _block(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(n) * 20
    v += 40 * np.cos(2 * np.pi * 0.1185 * np.arange(n))
    v[n // 2:n // 2 + 256] += rng.standard_normal(256) * 110
    return np.clip(np.round(v), -127, 127).astype(np.int8)


def _oracle_cfg(n, C_, dm, freq_pairs):
    import ctypes as CT
    import oracle_lib
    oc = oracle_lib.ChainConfig()
    oc.baseband_input_count, oc.baseband_input_bits, oc.window = n, -8, 0
    oc.baseband_freq_low, oc.baseband_bandwidth, oc.baseband_sample_rate, oc.dm = 1000.0, 500.0, 1e9, dm
    oc.baseband_reserve_sample = 0
    oc.rfi_average_threshold, oc.rfi_sk_threshold = 5.0, 1.3
    oc.spectrum_channel_count = C_
    oc.snr_threshold, oc.channel_threshold, oc.max_boxcar_length = 6.0, 0.9, 64
    flat = [v for p in freq_pairs for v in p]
    arr = (CT.c_float * max(1, len(flat)))(*flat)
    oc._keep = arr
    oc.rfi_pairs = CT.cast(arr, CT.POINTER(CT.c_float))
    oc.n_rfi_pairs = len(freq_pairs)
    return oc


def _run(tmp_path, fmt, raw, logn, C_, dm, composite, freq_list="1200-1201", extra=()):
    inp = tmp_path / f"bb_{fmt}_{composite}.bin"
    raw.tofile(inp)
    prefix = tmp_path / f"dump_{fmt}_{composite}_"
    cmd = [str(BIN), "--input", str(inp), "--log2n", str(logn), "--bits", "-8", "--format", fmt, "--channels",
           str(C_), "--dm", str(dm), "--avg-thr", "5", "--sk-thr", "1.3", "--snr", "6", "--max-boxcar", "64",
           "--freq-list", freq_list, "--dump-prefix", str(prefix), "--composite", str(composite), *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    works = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    return works, prefix


def _check_work(w, prefix, bb, oracle, n, C_, dm):
    work, eres, eseries, _ = oracle.chain(bb.view(np.uint8), _oracle_cfg(n, C_, dm, [(1200.0, 1201.0)]))
    L = n // 2 // C_
    assert w["count"] == L and w["batch_size"] == C_
    spec = np.fromfile(f"{prefix}{w['block']}.{w['stream']}.bin", dtype=np.complex64).reshape(C_, L)
    espec = work[:n].view(np.complex64).reshape(C_, L)
    gz, ez = np.all(spec == 0, axis=1), np.all(espec == 0, axis=1)
    nz = ~(gz | ez)
    whole = np.linalg.norm(spec[nz].astype(np.complex128) - espec[nz]) / max(1e-30, np.linalg.norm(espec[nz].astype(np.complex128)))
    assert (gz != ez).sum() <= 1, (int(gz.sum()), int(ez.sum()), float(whole), w)
    same = gz == ez
    den = np.linalg.norm(espec[same].astype(np.complex128))
    if den == 0:        # everything zapped in both
        assert not spec[same].any()
    else:
        assert np.linalg.norm(spec[same].astype(np.complex128) - espec[same]) / den < 5e-5
    if np.array_equal(gz, ez):
        assert w["zero_count"] == eres.zero_count
        got = {s["boxcar"]: s["count"] for s in w["series"]}
        exp = {int(eres.boxcar_length[b]): int(eres.signal_count[b]) for b in range(eres.n_boxcars)
               if eres.signal_count[b] > 0}
        assert set(got) ^ set(exp) <= {b for b in set(got) | set(exp) if abs(got.get(b, 0) - exp.get(b, 0)) <= 1}
        for b in set(got) & set(exp):
            assert abs(got[b] - exp[b]) <= 1
    return len(w["series"])


@pytest.mark.parametrize("composite", [0, 1])
def test_pipeline_simple_three_blocks(tmp_path, oracle, composite):
    _build()
    logn, C_, dm = 18, 64, 0.0
    n = 1 << logn
    blocks = [_block(n, s) for s in (1, 2, 3)]
    works, prefix = _run(tmp_path, "simple", np.concatenate(blocks), logn, C_, dm, composite)
    assert sorted(w["block"] for w in works) == [0, 1, 2]          # every block came out, once
    detected = 0
    for w in works:
        assert w["stream"] == 0
        detected += _check_work(w, prefix, blocks[w["block"]], oracle, n, C_, dm)
    assert detected > 0                                             # the injected bursts were found


def test_pipeline_udp_shaped_with_packet_loss(tmp_path, oracle):
    """BASELINE config #5 / SURVEY 8(f-2): the stream arrives as fastmb_roach2 packets (8-byte counter + 4096
    bytes), every 7th packet is lost; udp_receiver_pipe assembles blocks by counter and zero-fills the holes
    (io/udp/udp_receiver.hpp:180-272). Each block's results equal the oracle's on the zero-filled block."""
    _build()
    logn, C_, dm = 18, 64, 0.0
    n = 1 << logn
    blocks = [_block(n, s) for s in (21, 22, 23)]
    works, prefix = _run(tmp_path, "simple", np.concatenate(blocks), logn, C_, dm, 1,
                         extra=("--udp-shaped", "1", "--udp-drop-every", "7"))
    ppb = n // 4096                                                  # packets per block
    assert sorted(w["block"] for w in works) == [0, ppb, 2 * ppb]    # keyed by first packet counter
    for w in works:
        k = w["block"] // ppb
        holed = blocks[k].copy().reshape(ppb, 4096)
        for pkt in range(ppb):
            if (k * ppb + pkt) % 7 == 6:
                holed[pkt] = 0
        _check_work(w, prefix, holed.reshape(-1), oracle, n, C_, dm)


@pytest.mark.parametrize("fused", [1, 3])
def test_pipeline_fused_chain_pipe(tmp_path, oracle, fused):
    """the device chain as ONE pipe (baseband_chain_pipe = srtb_b200_process_block behind the pipe
    API), one or three of them on their own queues fed from one MPMC queue. Same checks as the per-stage pipeline."""
    _build()
    logn, C_, dm = 18, 64, 0.0
    n = 1 << logn
    blocks = [_block(n, s) for s in (31, 32, 33, 34, 35)]
    works, prefix = _run(tmp_path, "simple", np.concatenate(blocks), logn, C_, dm, 0, extra=("--fused", str(fused)))
    assert sorted(w["block"] for w in works) == list(range(5))
    detected = 0
    for w in works:
        assert w["stream"] == 0
        detected += _check_work(w, prefix, blocks[w["block"]], oracle, n, C_, dm)
    assert detected > 0


def test_pipeline_fused_chain_pipe_ring_equals_synchronous(tmp_path):
    """baseband_chain_pipe over the submit/collect ring (H2D of block k overlaps block k-1; result headers, the
    positive boxcar series and the dynamic spectrum all come from the ring slot — no block is run twice) must report
    exactly what the synchronous pipe reports, series included (their peaks are printed by the sink)."""
    _build()
    logn, C_ = 18, 64
    n = 1 << logn
    blocks = [_block(n, s) for s in range(51, 58)]
    inp = tmp_path / "bb_ring.bin"
    np.concatenate(blocks).tofile(inp)
    outs = []
    for extra in (("--fused", "1"), ("--fused", "1", "--ring", "3"), ("--fused", "2", "--ring", "2")):
        cmd = [str(BIN), "--input", str(inp), "--log2n", str(logn), "--bits", "-8", "--format", "simple", "--channels",
               str(C_), "--dm", "0", "--avg-thr", "5", "--sk-thr", "1.3", "--snr", "6", "--max-boxcar", "64", *extra]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        works = sorted((json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")), key=lambda w: w["block"])
        assert [w["block"] for w in works] == list(range(7))
        outs.append([(w["zero_count"], [(s_["boxcar"], s_["count"], s_["length"], s_["peak"]) for s_ in w["series"]])
                     for w in works])
    assert outs[0] == outs[1] == outs[2]
    assert any(series for _, series in outs[0])                       # the bursts are candidates


def test_pipeline_fused_chain_pipe_dual_pol(tmp_path, oracle):
    _build()
    logn, C_, dm = 16, 16, 0.0
    n = 1 << logn
    a, b = _block(n, 41), _block(n, 42)
    raw = np.empty(2 * n, np.int8)
    raw.reshape(-1, 4)[:, 0:2] = a.reshape(-1, 2)
    raw.reshape(-1, 4)[:, 2:4] = b.reshape(-1, 2)
    works, prefix = _run(tmp_path, "naocpsr_snap1", raw, logn, C_, dm, 0, extra=("--fused", "1"))
    assert sorted((w["block"], w["stream"]) for w in works) == [(0, 0), (0, 1)]
    for w in works:
        _check_work(w, prefix, (a, b)[w["stream"]], oracle, n, C_, dm)


def test_pipeline_dual_pol_fanout(tmp_path, oracle):
    """naocpsr_snap1: one unpack work fans out into two fft works with data_stream_id 2*id + s
    (unpack_pipe.hpp:249-258)"""
    _build()
    logn, C_, dm = 16, 16, 0.0
    n = 1 << logn
    a, b = _block(n, 11), _block(n, 12)
    raw = np.empty(2 * n, np.int8)
    raw.reshape(-1, 4)[:, 0:2] = a.reshape(-1, 2)
    raw.reshape(-1, 4)[:, 2:4] = b.reshape(-1, 2)
    works, prefix = _run(tmp_path, "naocpsr_snap1", raw, logn, C_, dm, 0)
    assert sorted((w["block"], w["stream"]) for w in works) == [(0, 0), (0, 1)]
    for w in works:
        _check_work(w, prefix, (a, b)[w["stream"]], oracle, n, C_, dm)


def test_pipeline_cfg_file_reader_overlap_and_candidate_sink(tmp_path, oracle):
    """the pipeline driven by a reference-style .cfg (expression values), fed by
    read_file_pipe (overlap-save rewind) and drained by write_signal_pipe (.bin/.npy/.tim)."""
    _build()
    logn, C_ = 16, 16
    n = 1 << logn
    rng = np.random.default_rng(5)
    total = 3 * n
    v = rng.standard_normal(total) * 20
    v[n // 2:n // 2 + 128] += rng.standard_normal(128) * 110          # burst in block 0
    raw = np.clip(np.round(v), -127, 127).astype(np.int8)
    inp = tmp_path / "bb.bin"
    raw.tofile(inp)
    prefix = tmp_path / "cand_"
    cfg = tmp_path / "srtb_test.cfg"
    cfg.write_text(f"""# reference-style config: every numeric value is an expression
baseband_input_count = 2 ** {logn}
baseband_input_bits = -8
baseband_format_type = simple
baseband_freq_low = 1000 + (0 / 2)
baseband_bandwidth = 500
baseband_sample_rate = 1000 * 1e6
baseband_reserve_sample = 1
dm = 0.0005
spectrum_channel_count = 2 ** 4
mitigate_rfi_average_method_threshold = 10
mitigate_rfi_spectral_kurtosis_threshold = 1.5
signal_detect_signal_noise_threshold = 6
signal_detect_max_boxcar_length = 2 ** 5
input_file_path = {inp}
baseband_output_file_prefix = {prefix}
""")
    r = subprocess.run([str(BIN), "--config_file_name", str(cfg), "--write-candidates", "1", "--dump-prefix",
                        str(tmp_path / "dump_")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    works = sorted((json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")), key=lambda w: w["block"])
    # overlap: with reserve on, consecutive blocks start (n - reserved) samples apart -> more than 3 blocks
    reserved = oracle.nsamps_reserved(n, C_, 1000.0, 500.0, 1e9, 0.0005, True)
    assert 0 < reserved < n
    stride = n - reserved
    expect_blocks = 1
    pos = stride
    while pos < total:
        expect_blocks += 1
        pos += stride
    assert len(works) == expect_blocks > 3
    L = n // 2 // C_
    assert all(w["count"] == L and w["batch_size"] == C_ for w in works)
    det = [w for w in works if w["series"]]
    assert det and det[0]["block"] == 0
    # candidate files of the first detection
    w = det[0]
    stem = f"{prefix}{w['block']}"
    bb = np.fromfile(stem + ".bin", dtype=np.int8)
    assert np.array_equal(bb, raw[:n])                                   # raw baseband of that block
    spec = np.load(stem + ".0.npy")
    assert spec.shape == (C_, L) and spec.dtype == np.complex64          # plot_spectrum.py's [freq][time]
    dumped = np.fromfile(f"{tmp_path / 'dump_'}{w['block']}.0.bin", dtype=np.complex64).reshape(C_, L)
    assert np.array_equal(spec, dumped)
    for sinfo in w["series"]:
        tim = np.fromfile(f"{stem}.{sinfo['boxcar']}.tim", dtype=np.float32)
        assert tim.size == sinfo["length"] and abs(float(tim.max()) - sinfo["peak"]) < 1e-3 * abs(sinfo["peak"]) + 1e-3
    # blocks without a detection leave no files (file mode keeps only positives)
    quiet = [w for w in works if not w["series"]]
    assert quiet and not Path(f"{prefix}{quiet[0]['block']}.bin").exists()


def test_alternate_pipes_ifft_refft():
    """ifft_1d_c2c_pipe -> refft_1d_c2c_pipe (the reference's alternative back half, fft_pipe.hpp:88-278) against a
    float64 DFT, including the overlap-save tail cut (tests/cpp/test_alt_pipes.cpp)"""
    d = BIN.parent
    subprocess.run(["make", "-C", str(d), "test_alt_pipes"], check=True, capture_output=True)
    r = subprocess.run([str(d / "test_alt_pipes")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "alt pipes ok" in r.stdout, (r.stdout[-500:], r.stderr[-1500:])


def test_product_executable_replays_a_file(tmp_path, oracle):
    """src/srtb_b200 — the drop-in executable: reference-style config file (expressions), file source, fused chain on
    every visible GPU, candidate sink. Its .tim / .npy / .bin output for a block with a burst must match what the
    oracle finds in that block."""
    exe = ROOT / "src" / "srtb_b200"
    subprocess.run(["make", "-C", str(ROOT / "src")], check=True, capture_output=True)
    logn, C_ = 18, 64
    n = 1 << logn
    blocks = [_block(n, s) for s in range(61, 65)]
    inp = tmp_path / "bb.bin"
    np.concatenate(blocks).tofile(inp)
    prefix = tmp_path / "cand_"
    cfg = tmp_path / "replay.cfg"
    cfg.write_text(f"""# replay
baseband_input_count = 2 ** {logn}
baseband_input_bits = -8
baseband_format_type = simple
baseband_freq_low = 1000.0
baseband_bandwidth = 500.0
baseband_sample_rate = 1000 * 1e6
baseband_reserve_sample = 0
dm = 0
spectrum_channel_count = 2 ** 6
mitigate_rfi_average_method_threshold = 5
mitigate_rfi_spectral_kurtosis_threshold = 1.3
signal_detect_signal_noise_threshold = 6
signal_detect_max_boxcar_length = 64
input_file_path = {inp}
baseband_output_file_prefix = {prefix}
log_level = 2
""")
    r = subprocess.run([str(exe), "--config_file_name", str(cfg), "--chains_per_gpu", "2", "--ring_depth", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "4 block(s) x 1 stream(s)" in r.stderr
    tims = sorted(tmp_path.glob("cand_*.tim"))
    npys = sorted(tmp_path.glob("cand_*.npy"))
    assert tims and npys, list(tmp_path.iterdir())
    # every block has the burst: one spectrum per block, and the boxcar-1 series of block 0 equals the oracle's
    assert len(npys) == 4
    spec = np.load(npys[0])
    assert spec.shape == (C_, n // 2 // C_) and spec.dtype == np.complex64
    import oracle_lib
    oc = oracle_lib.ChainConfig()
    oc.baseband_input_count, oc.baseband_input_bits, oc.window = n, -8, 0
    oc.baseband_freq_low, oc.baseband_bandwidth, oc.baseband_sample_rate, oc.dm = 1000.0, 500.0, 1e9, 0.0
    oc.baseband_reserve_sample = 0
    oc.rfi_average_threshold, oc.rfi_sk_threshold, oc.spectrum_channel_count = 5.0, 1.3, C_
    oc.snr_threshold, oc.channel_threshold, oc.max_boxcar_length = 6.0, 0.9, 64
    oc.n_rfi_pairs = 0
    work, eres, eseries, _ = oracle.chain(blocks[0].view(np.uint8), oc)
    espec = work[:n].view(np.complex64).reshape(C_, -1)
    # file replays carry no packet counter: the candidate files are named after the block's timestamp
    # (write_signal_pipe.hpp:145-148), so block 0's files are found by content
    keep = ~np.all(espec == 0, axis=1)
    match = [p for p in npys if np.linalg.norm(np.load(p)[keep] - espec[keep]) / np.linalg.norm(espec[keep]) < 5e-5]
    assert len(match) == 1, [p.name for p in npys]
    stem = match[0].name.split(".")[0]
    assert (tmp_path / f"{stem}.bin").stat().st_size == n                      # the raw block next to it
    assert eres.signal_count[0] > 0
    ts = np.fromfile(tmp_path / f"{stem}.1.tim", np.float32)
    assert ts.size == int(eres.series_length[0])
    assert np.abs(ts - eseries[0, :ts.size]).max() < 2e-4 * np.sqrt(np.mean(eseries[0, :ts.size].astype(np.float64) ** 2))

#!/bin/bash
# round-2 session Y: host-to-host throughput through the C++ pipes (pipeline_main, preloaded pinned blocks), config-3 and config-2 shapes; short soak
nvidia-smi -L
python - <<'PY'
import numpy as np
rng = np.random.default_rng(5)
for name, n in (("c3", 4 * (1 << 27)), ("c2", 16 * (1 << 24))):      # 4 dual-pol blocks of 2 x 2^26; 16 blocks of 2^24
    out = np.empty(n, np.int8)
    for i in range(0, n, 1 << 26):
        out[i:i + (1 << 26)] = np.clip(np.round(rng.standard_normal(1 << 26, dtype=np.float32) * 20), -127, 127).astype(np.int8)
    out.tofile(f"/tmp/bb_{name}.bin")
PY
P=tests/cpp/pipeline_main
C3="--input /tmp/bb_c3.bin --log2n 26 --bits -8 --format naocpsr_snap1 --channels 2048 --dm 562.05 --freq-low 1000 --bandwidth 400 --sample-rate 8e8 --avg-thr 1.5 --sk-thr 1.05 --snr 8 --max-boxcar 256 --freq-list 1018-1022 --preload 1 --repeat 40"
C2="--input /tmp/bb_c2.bin --log2n 24 --bits -8 --format simple --channels 2048 --dm 56.778 --avg-thr 5 --sk-thr 1.05 --snr 8 --max-boxcar 256 --preload 1 --repeat 16"
{
for mode in "--fused 1" "--fused 1 --ring 3" "--fused 2 --ring 3" "--composite 0" "--composite 1"; do
  echo "== config 3 shape (dual-pol 2 x 2^26 per block), $mode"; $P $C3 $mode 2>&1 >/dev/null | grep pipeline_main
done
for mode in "--fused 1 --ring 3" "--fused 3 --ring 3" "--composite 0"; do
  echo "== config 2 shape (2^24 per block), $mode"; $P $C2 $mode 2>&1 >/dev/null | grep pipeline_main
done
} | tee gpurun_out/r02y_cpp_pipeline_throughput.txt
python bench.py --steps 4000 --warmup 6 --no-cpu-baseline --stage-iters 1 --secondary none > gpurun_out/bench_r02y_soak.json 2> gpurun_out/bench_r02y_soak.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r02y_soak.json').read().strip().splitlines()[-1]); print('soak', d['steps'], round(d['value'],2), round(d['e2e']['value'],2), d['clocks'], d['config']['detections'])" || tail -3 gpurun_out/bench_r02y_soak.err

#!/bin/bash
# A/B of library build variants (tile width of the COL/TRANS FFT passes) in one GPU session
mkdir -p gpurun_out
for v in libsrtb_b200.so libsrtb_b200_t8.so libsrtb_b200_t4.so; do
  echo "=== $v"
  SRTB_B200_LIB=$v python -m pytest tests/test_gpu_parity.py -q -k "fft_c2c_vs_float64 or fft_r2c_inplace or chain_vs_oracle" --timeout 300 2>&1 | tail -2
  SRTB_B200_LIB=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
print("$v", "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), {k: round(s["ms"]*1e3,1) for k,s in d["stages"].items()})
PY
done

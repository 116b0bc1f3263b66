#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pipeline.py -q --timeout 300 2>&1 | tail -3
for c in 1 2 3; do
  python bench.py --steps 200 --warmup 5 --no-cpu-baseline --contexts $c > gpurun_out/ctx_$c.json 2>gpurun_out/ctx_$c.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ctx_$c.json").read().strip().splitlines()[-1])
print("contexts $c value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "ms", round(d["ms_per_step"],4))
PY
done
SRTB_BENCH_WORKLOAD=config3 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --contexts 1 > gpurun_out/cfg3.json 2>gpurun_out/cfg3.err; tail -c 1800 gpurun_out/cfg3.json; tail -3 gpurun_out/cfg3.err

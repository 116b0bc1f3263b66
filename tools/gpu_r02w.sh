#!/bin/bash
# round-2 session W: 2^25 = 2^8 * 2^9 * 2^8 (short raw first sweep, now also 3 CTAs/SM) against the default 2^9 * 2^8 * 2^8
nvidia-smi -L
B="bench.py --steps 60 --warmup 6 --no-cpu-baseline --stage-iters 1 --secondary none"
for fs in 1 0; do
  SRTB_B200_PLAN_FIRST_SHORT=$fs python $B > gpurun_out/bench_r02w_fs$fs.json 2> gpurun_out/bench_r02w_fs$fs.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_r02w_fs$fs.json').read().strip().splitlines()[-1]); print('first_short=$fs', round(d['value'],2), round(d['ms_per_step'],4), round(d['e2e']['value'],2))" || tail -3 gpurun_out/bench_r02w_fs$fs.err
done
SRTB_B200_PLAN_FIRST_SHORT=1 SRTB_B200_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02w_fs1.csv \
  python $B --steps 2 --warmup 3 --contexts 1 --no-pulse > gpurun_out/ncu_r02w.log 2>&1
python - <<'PY'
import csv
f='gpurun_out/launches_r02w_fs1.csv'
lines=[l for l in open(f) if not l.startswith('==')]
rows=[(x['Kernel Name'][:70], float(x['Metric Value'])/1000) for x in csv.DictReader(lines)]
for n,t in rows[-16:-8]: print('  %-72s %8.1f us'%(n,t))
PY
SRTB_B200_PLAN_FIRST_SHORT=1 python -m pytest tests -m gpu -q --timeout 900 -x -k "config3 or chain_vs_oracle or ring" 2>&1 | tail -3
python -m pytest tests -m gpu -q --timeout 900 -x -k "fft or r2c or c2c or watfft or pipeline_simple or config1" 2>&1 | tail -3

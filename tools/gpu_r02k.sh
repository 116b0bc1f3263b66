#!/bin/bash
# round-2 session K: long-row fused waterfall (rows 2^15..2^18), full-size config 1, configs 1/4 benches
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout 1800 -x 2>&1 | tail -8 | tee gpurun_out/pytest_r02k.log
for w in config1 config4; do for lf in 1 0; do
  SRTB_B200_LONG_FUSED=$lf python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --stage-iters 1 --secondary none > gpurun_out/bench_r02k_${w}_lf$lf.json 2> gpurun_out/bench_r02k_${w}_lf$lf.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_r02k_${w}_lf$lf.json').read().strip().splitlines()[-1]); print('$w long_fused=$lf', round(d['value'],2), round(d['ms_per_step'],3), d['gpu_launches'], round(d['e2e']['value'],2))" || tail -5 gpurun_out/bench_r02k_${w}_lf$lf.err
done; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02k_c1.csv \
  python bench.py --workload config1 --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 --secondary none > gpurun_out/ncu_r02k.log 2>&1
python - <<'PY'
import csv
f='gpurun_out/launches_r02k_c1.csv'
lines=[l for l in open(f) if not l.startswith('==')]
rows=[(x['Kernel Name'][:70], float(x['Metric Value'])/1000) for x in csv.DictReader(lines)]
for n,t in rows[-26:-13]: print('  %-72s %8.1f us'%(n,t))
PY

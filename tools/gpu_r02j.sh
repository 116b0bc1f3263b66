#!/bin/bash
# round-2 session J: coalesced detector tail, alternate pipes, product executable test
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout 1800 -x -k "not config1_full_size" 2>&1 | tail -8 | tee gpurun_out/pytest_r02j.log
run() { tag=$1; shift
  env "$@" python bench.py --steps 60 --warmup 3 --no-cpu-baseline --stage-iters 1 --secondary config2 > gpurun_out/bench_r02j_$tag.json 2> gpurun_out/bench_r02j_$tag.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/bench_r02j_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value'],2), round(d['single_context']['value'],2), d['gpu_launches'], round(d['e2e']['value'],2), {k:round(v['ms'],3) for k,v in d['roofline']['fused'].items()}, d['config']['detections'], round(d['secondary']['value'],2))" || tail -5 gpurun_out/bench_r02j_$tag.err
}
run default
run ctx2 SRTB_BENCH_CONTEXTS=2
run ctx3 SRTB_BENCH_CONTEXTS=3
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_r02j_c3.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 --secondary none --no-pulse > gpurun_out/ncu_r02j.log 2>&1
python - <<'PY'
import csv
f='gpurun_out/launches_r02j_c3.csv'
lines=[l for l in open(f) if not l.startswith('==')]
rows=[(x['Kernel Name'][:70], float(x['Metric Value'])/1000) for x in csv.DictReader(lines)]
for n,t in rows[32:41]: print('  %-72s %8.1f us'%(n,t))
PY

#!/bin/bash
# round-2 session C: whole-row kernel with per-chunk load pipelining; plan A/B; launch list
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout 900 -x -k "fft_c2c or watfft or chain or fused_chirp or golden or dm_sweep or ring" 2>&1 | tail -6 | tee gpurun_out/pytest_r02c.log
run() { # tag, env...
  tag=$1; shift
  env "$@" python bench.py --workload config3 --steps 40 --warmup 3 --no-cpu-baseline --stage-iters 1 > gpurun_out/bench_r02c_$tag.json 2> gpurun_out/bench_r02c_$tag.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/bench_r02c_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value'],2), round(d['ms_per_step'],4), d['gpu_launches'], round(d['e2e']['value'],2))" || tail -5 gpurun_out/bench_r02c_$tag.err
}
run ctx4 SRTB_BENCH_CONTEXTS=4
run ctx1 SRTB_BENCH_CONTEXTS=1
run ctx2 SRTB_BENCH_CONTEXTS=2
run ctx4_first_short SRTB_BENCH_CONTEXTS=4 SRTB_B200_PLAN_FIRST_SHORT=1
run ctx1_first_short SRTB_BENCH_CONTEXTS=1 SRTB_B200_PLAN_FIRST_SHORT=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r02c_c3.csv \
  python bench.py --workload config3 --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 > gpurun_out/ncu_r02c.log 2>&1
SRTB_B200_PLAN_FIRST_SHORT=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r02c_c3_fs.csv \
  python bench.py --workload config3 --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 > gpurun_out/ncu_r02c_fs.log 2>&1
python - <<'PY'
import csv
for f in ('gpurun_out/launches_r02c_c3.csv','gpurun_out/launches_r02c_c3_fs.csv'):
    lines=[l for l in open(f) if not l.startswith('==')]
    rows=[(x['Kernel Name'][:70], float(x['Metric Value'])/1000) for x in csv.DictReader(lines)]
    print(f)
    for n,t in rows[40:58]: print('  %-72s %8.1f us'%(n,t))
PY

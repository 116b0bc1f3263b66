#!/bin/bash
# round-2 session O: phases loaded one pass ahead, split reads per half warp, context lock; full suite
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout 1800 -x 2>&1 | tail -8 | tee gpurun_out/pytest_r02o.log
for c in 1 2; do
  python bench.py --workload config3 --steps 60 --warmup 6 --no-cpu-baseline --stage-iters 1 --contexts $c --secondary none > gpurun_out/bench_r02o_c$c.json 2> gpurun_out/bench_r02o_c$c.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_r02o_c$c.json').read().strip().splitlines()[-1]); print('ctx=$c', round(d['value'],2), round(d['ms_per_step'],4), d['gpu_launches'], round(d['e2e']['value'],2), round(d['single_context']['value'],2))" || tail -5 gpurun_out/bench_r02o_c$c.err
done
for w in config2 config1 config4; do for tab in 1 0; do
  SRTB_B200_CHIRP_TABLE=$tab python bench.py --workload $w --steps 20 --warmup 4 --no-cpu-baseline --stage-iters 1 --secondary none > gpurun_out/bench_r02o_${w}_t$tab.json 2> gpurun_out/bench_r02o_${w}_t$tab.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_r02o_${w}_t$tab.json').read().strip().splitlines()[-1]); print('$w table=$tab', round(d['value'],2), round(d['ms_per_step'],4), d['gpu_launches'], round(d['e2e']['value'],2))" || tail -5 gpurun_out/bench_r02o_${w}_t$tab.err
done; done
SRTB_B200_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02o_c3.csv \
  python bench.py --workload config3 --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 --secondary none --no-pulse > gpurun_out/ncu_r02o.log 2>&1
python - <<'PY'
import csv
f='gpurun_out/launches_r02o_c3.csv'
lines=[l for l in open(f) if not l.startswith('==')]
rows=[(x['Kernel Name'][:70], float(x['Metric Value'])/1000) for x in csv.DictReader(lines)]
for n,t in rows[-16:-8]: print('  %-72s %8.1f us'%(n,t))
PY
SRTB_B200_LANES=1 ncu --set full --clock-control none --import-source on -k regex:'fft_bigrow|fft_trans_r2c16' -s 6 -c 2 -o gpurun_out/prof_r02o \
  python bench.py --workload config3 --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 --secondary none --no-pulse > gpurun_out/ncu_r02o_full.log 2>&1
ls -la gpurun_out/prof_r02o.ncu-rep

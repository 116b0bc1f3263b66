#!/bin/bash
# round-2 session N: tabulated chirp phases in the whole-row kernel; the per-pipe pipeline test that failed in session M
nvidia-smi -L
python -m pytest tests/test_gpu_pipeline.py -q --timeout 900 2>&1 | tail -40 | tee gpurun_out/pytest_r02n_pipeline.log
python -m pytest tests -m gpu -q --timeout 1800 -x -k "chain or ring or golden or config3 or fused_chirp" 2>&1 | tail -8 | tee gpurun_out/pytest_r02n.log
for tab in 1 0; do for c in 1 2; do
  SRTB_B200_CHIRP_TABLE=$tab python bench.py --workload config3 --steps 60 --warmup 6 --no-cpu-baseline --stage-iters 1 --contexts $c --secondary none > gpurun_out/bench_r02n_t${tab}_c$c.json 2> gpurun_out/bench_r02n_t${tab}_c$c.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_r02n_t${tab}_c$c.json').read().strip().splitlines()[-1]); print('table=$tab ctx=$c', round(d['value'],2), round(d['ms_per_step'],4), d['gpu_launches'], round(d['e2e']['value'],2), round(d['single_context']['value'],2))" || tail -5 gpurun_out/bench_r02n_t${tab}_c$c.err
done; done
SRTB_B200_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02n_c3.csv \
  python bench.py --workload config3 --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 --secondary none --no-pulse > gpurun_out/ncu_r02n.log 2>&1
python - <<'PY'
import csv
f='gpurun_out/launches_r02n_c3.csv'
lines=[l for l in open(f) if not l.startswith('==')]
rows=[(x['Kernel Name'][:70], float(x['Metric Value'])/1000) for x in csv.DictReader(lines)]
for n,t in rows[-16:-8]: print('  %-72s %8.1f us'%(n,t))
PY
SRTB_B200_LANES=1 ncu --set full --clock-control none --import-source on -k regex:'fft_bigrow' -s 4 -c 1 -o gpurun_out/prof_r02n \
  python bench.py --workload config3 --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 --secondary none --no-pulse > gpurun_out/ncu_r02n_full.log 2>&1
ls -la gpurun_out/prof_r02n.ncu-rep

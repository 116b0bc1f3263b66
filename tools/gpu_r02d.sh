#!/bin/bash
# round-2 session D: full suite (incl. full-size config 3 / config 1 oracle comparisons), new bench line, ncu of the whole-row kernel
nvidia-smi -L; free -g | head -2; nproc
python -m pytest tests -m gpu -q --timeout 1800 -s -k "full_size_vs_oracle" 2>&1 | grep -E "config|passed|failed|Error|error|assert" | tail -20 | tee gpurun_out/pytest_r02d_fullsize.log
python -m pytest tests -m gpu -q --timeout 900 -k "not full_size_vs_oracle" 2>&1 | tail -8 | tee gpurun_out/pytest_r02d.log
python bench.py --steps 60 --warmup 3 > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err; tail -c 6000 gpurun_out/bench_r02d.json; tail -5 gpurun_out/bench_r02d.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_r02d_c3.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 --secondary none --no-pulse > gpurun_out/ncu_r02d.log 2>&1
python - <<'PY'
import csv
f='gpurun_out/launches_r02d_c3.csv'
lines=[l for l in open(f) if not l.startswith('==')]
rows=[(x['Kernel Name'][:70], float(x['Metric Value'])/1000) for x in csv.DictReader(lines)]
for n,t in rows[32:50]: print('  %-72s %8.1f us'%(n,t))
PY
ncu --set full --clock-control none --import-source on -k regex:fft_bigrow -s 2 -c 1 -o gpurun_out/prof_r02d_bigrow -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 --secondary none --no-pulse > gpurun_out/ncu_r02d_full.log 2>&1
tail -2 gpurun_out/ncu_r02d_full.log

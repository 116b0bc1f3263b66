#!/bin/bash
# round-2 session F: whole-row kernel with unpadded chunks (16 TMA copies per row); tail scan fix; R2C kernel captures
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout 1800 -x -k "fft_c2c or watfft or chain or fused_chirp or golden or dm_sweep or ring or signal_detect or config3_full" 2>&1 | tail -6 | tee gpurun_out/pytest_r02f.log
run() { tag=$1; shift
  env "$@" python bench.py --steps 40 --warmup 3 --no-cpu-baseline --stage-iters 1 --secondary none > gpurun_out/bench_r02f_$tag.json 2> gpurun_out/bench_r02f_$tag.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/bench_r02f_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value'],2), round(d['single_context']['value'],2), d['gpu_launches'], round(d['e2e']['value'],2), {k:round(v['ms'],3) for k,v in d['roofline']['fused'].items()}, d['config']['detections'])" || tail -5 gpurun_out/bench_r02f_$tag.err
}
run ctx4 SRTB_BENCH_CONTEXTS=4
run ctx2 SRTB_BENCH_CONTEXTS=2
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_r02f_c3.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 --secondary none --no-pulse > gpurun_out/ncu_r02f.log 2>&1
python - <<'PY'
import csv
f='gpurun_out/launches_r02f_c3.csv'
lines=[l for l in open(f) if not l.startswith('==')]
rows=[(x['Kernel Name'][:70], float(x['Metric Value'])/1000) for x in csv.DictReader(lines)]
for n,t in rows[32:42]: print('  %-72s %8.1f us'%(n,t))
PY
ncu --set full --clock-control none --import-source on -k regex:"fft_bigrow|fft_trans_r2c16|fft_col16" -s 8 -c 4 -o gpurun_out/prof_r02f -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 --secondary none --no-pulse > gpurun_out/ncu_r02f_full.log 2>&1
tail -2 gpurun_out/ncu_r02f_full.log

#!/bin/bash
# round-2 session G: full suite, bench workloads config3/4/5, tail kernel capture
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout 1800 -k "not config1_full_size" 2>&1 | tail -6 | tee gpurun_out/pytest_r02g.log
python bench.py --steps 60 --warmup 3 > gpurun_out/bench_r02g_c3.json 2> gpurun_out/bench_r02g_c3.err; tail -3 gpurun_out/bench_r02g_c3.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r02g_c3.json').read().strip().splitlines()[-1]); print('c3', round(d['value'],2), round(d['single_context']['value'],2), round(d['e2e']['value'],2), d['secondary']['value'], d['config']['detections'])"
python bench.py --workload config4 --steps 10 --warmup 3 > gpurun_out/bench_r02g_c4.json 2> gpurun_out/bench_r02g_c4.err; tail -3 gpurun_out/bench_r02g_c4.err; head -c 1800 gpurun_out/bench_r02g_c4.json; echo
python bench.py --workload config5 > gpurun_out/bench_r02g_c5.json 2> gpurun_out/bench_r02g_c5.err; tail -3 gpurun_out/bench_r02g_c5.err; head -c 1800 gpurun_out/bench_r02g_c5.json; echo
ncu --set full --clock-control none --import-source on -k regex:"colsum_final_scan|detect_boxcar" -s 4 -c 2 -o gpurun_out/prof_r02g_tail -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 --secondary none --no-pulse > gpurun_out/ncu_r02g_full.log 2>&1
tail -2 gpurun_out/ncu_r02g_full.log

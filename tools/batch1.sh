#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 600 -k "dispersed or stage_stats or j1644" 2>&1 | tail -4
for c in 1 2 3 4 6; do python bench.py --contexts $c --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ctx$c.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_ctx$c.json'));print('contexts $c value %.2f e2e %.2f'%(d['value'],d['e2e']['value']), [round(x*1e3) for x in d['e2e']['runs_ms_per_step']])"; done
python bench.py --workload config1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config1.json 2>gpurun_out/bench_config1.err; python -c "
import json;d=json.load(open('gpurun_out/bench_config1.json'));print('config1 value %.2f e2e %.2f'%(d['value'],d['e2e']['value']), {k:round(s['ms']*1e3) for k,s in d['stages'].items()})"; tail -3 gpurun_out/bench_config1.err
tools/sanitize.sh r01l

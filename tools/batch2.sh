#!/bin/bash
python -m pytest tests -m gpu -q --timeout 900 -k "fft_r2c or fft_c2c or j1644 or tones" 2>&1 | tail -3
for v in 0 1; do
SRTB_B200_WIDE_COL=$v python bench.py --workload config1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c1_w$v.json 2>gpurun_out/bench_c1_w$v.err; python -c "
import json;d=json.load(open('gpurun_out/bench_c1_w$v.json'));print('config1 WIDE=$v value %.2f e2e %.2f'%(d['value'],d['e2e']['value']), {k:round(s['ms']*1e3) for k,s in d['stages'].items()})"; tail -2 gpurun_out/bench_c1_w$v.err
done
for v in 0 1; do
SRTB_B200_WIDE_COL=$v python bench.py --workload config3 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_w$v.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_c3_w$v.json'));print('config3 WIDE=$v value %.2f e2e %.2f'%(d['value'],d['e2e']['value']), {k:round(s['ms']*1e3) for k,s in d['stages'].items()})"
done

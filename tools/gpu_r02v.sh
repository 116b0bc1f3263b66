#!/bin/bash
# round-2 session V (final state, 1 GPU): full suite, smoke(), DRAM traffic captures, the driver's own commands
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout 1800 2>&1 | tail -6 | tee gpurun_out/pytest_r02v.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for w in config3 config2; do
  n=$([ $w = config3 ] && echo 134217728 || echo 16777216)
  SRTB_B200_LANES=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/traffic_r02v_$w.csv \
    python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 --secondary none --no-pulse > gpurun_out/ncu_r02v_$w.log 2>&1
  python tools/make_traffic.py $w gpurun_out/traffic_r02v_$w.csv $n r02v && cp profiles/traffic_$w.json gpurun_out/
done
python bench.py > gpurun_out/bench_r02v_default.json 2> gpurun_out/bench_r02v_default.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r02v_default.json').read().strip().splitlines()[-1]); r=d['roofline']; print('default', round(d['value'],2), round(d['ms_per_step'],4), d['gpu_launches'], round(d['e2e']['value'],2), round(d['secondary']['value'],2), {k:r[k] for k in ('kernel','achieved','frac','traffic','share_of_stream_time')}, r['chain']['dram_measured'], d['cpu_baseline']['value'])" || tail -5 gpurun_out/bench_r02v_default.err
python bench.py --impl reference > gpurun_out/bench_r02v_reference.json 2> gpurun_out/bench_r02v_reference.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r02v_reference.json').read().strip().splitlines()[-1]); print('reference', d['value'], d['cpu_baseline']['cores'])"

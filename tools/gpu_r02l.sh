#!/bin/bash
# round-2 session L: two lanes per context (odd data streams on a second CUDA stream), DRAM traffic capture of config 3
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout 1800 -x -k "chain or ring or golden or pipeline or config3 or executable or alt" 2>&1 | tail -8 | tee gpurun_out/pytest_r02l.log
for lanes in 2 1; do for c in 1 2 4; do
  SRTB_B200_LANES=$lanes python bench.py --workload config3 --steps 60 --warmup 6 --no-cpu-baseline --stage-iters 1 --contexts $c --secondary none > gpurun_out/bench_r02l_l${lanes}_c$c.json 2> gpurun_out/bench_r02l_l${lanes}_c$c.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_r02l_l${lanes}_c$c.json').read().strip().splitlines()[-1]); print('lanes=$lanes ctx=$c', round(d['value'],2), round(d['ms_per_step'],4), d['gpu_launches'], round(d['e2e']['value'],2), d.get('single_context'))" || tail -5 gpurun_out/bench_r02l_l${lanes}_c$c.err
done; done
SRTB_B200_LANES=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/traffic_r02l_c3.csv \
  python bench.py --workload config3 --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 --secondary none --no-pulse > gpurun_out/ncu_r02l.log 2>&1
python tools/make_traffic.py config3 gpurun_out/traffic_r02l_c3.csv 134217728 r02l && cp profiles/traffic_config3.json gpurun_out/
SRTB_B200_LANES=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/traffic_r02l_c2.csv \
  python bench.py --workload config2 --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 --secondary none --no-pulse > gpurun_out/ncu_r02l2.log 2>&1
python tools/make_traffic.py config2 gpurun_out/traffic_r02l_c2.csv 16777216 r02l && cp profiles/traffic_config2.json gpurun_out/

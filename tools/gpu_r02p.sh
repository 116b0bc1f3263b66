#!/bin/bash
# round-2 session P: v1 alternates on the GPU, then the driver's own commands (default bench + reference arm)
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout 1800 -x -k "v1 or alt or executable or pipeline_simple or chain" 2>&1 | tail -8 | tee gpurun_out/pytest_r02p.log
python bench.py > gpurun_out/bench_r02p_default.json 2> gpurun_out/bench_r02p_default.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r02p_default.json').read().strip().splitlines()[-1]); print('default', round(d['value'],2), round(d['ms_per_step'],4), d['gpu_launches'], round(d['e2e']['value'],2), d.get('single_context'), d.get('secondary'), d['roofline'], d['cpu_baseline'])" || tail -5 gpurun_out/bench_r02p_default.err
python bench.py --impl reference > gpurun_out/bench_r02p_reference.json 2> gpurun_out/bench_r02p_reference.err
tail -c 1500 gpurun_out/bench_r02p_reference.json

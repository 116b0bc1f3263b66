#!/bin/bash
# round-2 session A: whole-row waterfall kernel — parity of the new sizes, config-3 bench A/B, launch list
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout 600 -x -k "fft_c2c or watfft or chain or fused_chirp or golden or dm_sweep" 2>&1 | tail -25 | tee gpurun_out/pytest_r02a.log
for br in 1 0; do
  SRTB_B200_BIGROW=$br python bench.py --workload config3 --steps 40 --warmup 3 --no-cpu-baseline --stage-iters 2 \
    > gpurun_out/bench_r02a_c3_bigrow$br.json 2> gpurun_out/bench_r02a_c3_bigrow$br.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/bench_r02a_c3_bigrow$br.json').read().strip().splitlines()[-1]); print('bigrow=$br', d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'], {k:round(v['ms'],3) for k,v in d['stages'].items()})" || tail -5 gpurun_out/bench_r02a_c3_bigrow$br.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r02a_c3.csv \
  python bench.py --workload config3 --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 > gpurun_out/ncu_r02a.log 2>&1
tail -3 gpurun_out/ncu_r02a.log

#!/bin/bash
# round-2 session B: full GPU suite after the ring refactor, config-3 bench (1 and 4 contexts), ncu of the whole-row kernel
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 | tee gpurun_out/pytest_r02b.log
for c in 4 1; do
  python bench.py --workload config3 --steps 40 --warmup 3 --no-cpu-baseline --stage-iters 2 --contexts $c \
    > gpurun_out/bench_r02b_c3_ctx$c.json 2> gpurun_out/bench_r02b_c3_ctx$c.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/bench_r02b_c3_ctx$c.json').read().strip().splitlines()[-1]); print('ctx=$c', d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'])" || tail -5 gpurun_out/bench_r02b_c3_ctx$c.err
done
SRTB_B200_FUSE_CHIRP=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r02b_c3_nochirp.csv \
  python bench.py --workload config3 --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 > gpurun_out/ncu_r02b_nochirp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_bigrow -s 2 -c 2 -o gpurun_out/prof_r02b_bigrow -f \
  python bench.py --workload config3 --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 --contexts 1 > gpurun_out/ncu_r02b_full.log 2>&1
tail -3 gpurun_out/ncu_r02b_full.log; ls -la gpurun_out/*.ncu-rep

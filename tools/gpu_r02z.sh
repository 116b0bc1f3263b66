#!/bin/bash
# round-2 session Z (8 GPUs): config 3 with the lazy NCCL group + gloo barriers (RankSync) — the last open point of the scaling table
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8 --master-port 29571 bench.py --gpus 8 --steps 60 --warmup 6 --no-cpu-baseline --stage-iters 1 --secondary none > gpurun_out/bench_r02z_c3_8gpu.json 2> gpurun_out/bench_r02z_c3_8gpu.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r02z_c3_8gpu.json').read().strip().splitlines()[-1]); print('c3_8gpu', d['n_gpus'], round(d['value'],2), round(d['ms_per_step'],4), round(d['e2e']['value'],2), d['config'].get('ms_per_step_by_rank'))" || tail -5 gpurun_out/bench_r02z_c3_8gpu.err

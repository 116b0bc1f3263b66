#!/bin/bash
# round-2 session U (2 GPUs): lazy NCCL group + gloo barriers in bench.py; second lane at low stream priority (experiment)
nvidia-smi -L
B="bench.py --steps 60 --warmup 6 --no-cpu-baseline --stage-iters 1 --secondary none"
show() { python -c "import json,sys; d=json.loads(open('gpurun_out/bench_r02u_$1.json').read().strip().splitlines()[-1]); print('$1', d['n_gpus'], round(d['value'],2), round(d['ms_per_step'],4), round(d['e2e']['value'],2), d['config'].get('ms_per_step_by_rank'))" || tail -3 gpurun_out/bench_r02u_$1.err; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
$TR --master-port 29561 $B --gpus 2 > gpurun_out/bench_r02u_default.json 2> gpurun_out/bench_r02u_default.err; show default
SRTB_B200_LANE_PRIORITY=1 $TR --master-port 29562 $B --gpus 2 > gpurun_out/bench_r02u_prio.json 2> gpurun_out/bench_r02u_prio.err; show prio
SRTB_B200_LANE_PRIORITY=1 CUDA_DEVICE_MAX_CONNECTIONS=32 $TR --master-port 29563 $B --gpus 2 > gpurun_out/bench_r02u_prio32.json 2> gpurun_out/bench_r02u_prio32.err; show prio32
SRTB_B200_LANE_PRIORITY=1 SRTB_BENCH_EAGER_NCCL=1 $TR --master-port 29564 $B --gpus 2 > gpurun_out/bench_r02u_prio_eager.json 2> gpurun_out/bench_r02u_prio_eager.err; show prio_eager

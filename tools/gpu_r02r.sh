#!/bin/bash
# round-2 session R (2 GPUs): why is the per-GPU rate lower under torchrun? separates box / dist / lanes effects
nvidia-smi -L
B="bench.py --steps 60 --warmup 6 --no-cpu-baseline --stage-iters 1 --secondary none"
show() { python -c "import json,sys; d=json.loads(open('gpurun_out/bench_r02r_$1.json').read().strip().splitlines()[-1]); print('$1', d['n_gpus'], round(d['value'],2), round(d['ms_per_step'],4), round(d['e2e']['value'],2))" || tail -3 gpurun_out/bench_r02r_$1.err; }
# (1) one process, one GPU of this box
python $B > gpurun_out/bench_r02r_single.json 2> gpurun_out/bench_r02r_single.err; show single
# (2) two independent single-GPU processes at the same time (no torch.distributed at all)
CUDA_VISIBLE_DEVICES=0 python $B > gpurun_out/bench_r02r_indep0.json 2> gpurun_out/bench_r02r_indep0.err &
CUDA_VISIBLE_DEVICES=1 python $B > gpurun_out/bench_r02r_indep1.json 2> gpurun_out/bench_r02r_indep1.err &
wait; show indep0; show indep1
# (3) torchrun, 2 ranks: default, one lane + two contexts, and with NCCL told to stay off the SMs between barriers
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
$TR --master-port 29521 $B --gpus 2 > gpurun_out/bench_r02r_tr_default.json 2> gpurun_out/bench_r02r_tr_default.err; show tr_default
SRTB_B200_LANES=1 $TR --master-port 29522 $B --gpus 2 --contexts 2 > gpurun_out/bench_r02r_tr_lanes1.json 2> gpurun_out/bench_r02r_tr_lanes1.err; show tr_lanes1
SRTB_BENCH_GLOO_BARRIER=1 $TR --master-port 29523 $B --gpus 2 > gpurun_out/bench_r02r_tr_gloo.json 2> gpurun_out/bench_r02r_tr_gloo.err; show tr_gloo

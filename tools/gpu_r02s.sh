#!/bin/bash
# round-2 session S (2 GPUs): NCCL in the process costs the two-lane path 8 % — hardware-queue aliasing? (CUDA_DEVICE_MAX_CONNECTIONS)
nvidia-smi -L
B="bench.py --steps 60 --warmup 6 --no-cpu-baseline --stage-iters 1 --secondary none"
show() { python -c "import json,sys; d=json.loads(open('gpurun_out/bench_r02s_$1.json').read().strip().splitlines()[-1]); print('$1', d['n_gpus'], round(d['value'],2), round(d['ms_per_step'],4), round(d['e2e']['value'],2), d['config'].get('lost_packets'), d['config'].get('real_time'))" || tail -3 gpurun_out/bench_r02s_$1.err; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
for mc in 32 8 1; do
  CUDA_DEVICE_MAX_CONNECTIONS=$mc $TR --master-port 2953$mc $B --gpus 2 > gpurun_out/bench_r02s_mc$mc.json 2> gpurun_out/bench_r02s_mc$mc.err; show mc$mc
done
NCCL_LAUNCH_ORDER_IMPLICIT=0 TORCH_NCCL_ENABLE_MONITORING=0 TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC=100000 CUDA_DEVICE_MAX_CONNECTIONS=32 $TR --master-port 29541 $B --gpus 2 > gpurun_out/bench_r02s_nomon.json 2> gpurun_out/bench_r02s_nomon.err; show nomon
$TR --master-port 29542 bench.py --gpus 2 --workload config5 --no-cpu-baseline --secondary none > gpurun_out/bench_r02s_c5.json 2> gpurun_out/bench_r02s_c5.err; show c5

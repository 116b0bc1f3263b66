#!/bin/bash
# round-2 session Q (8 GPUs of one box): block-sharded scaling of config 3, the DM sweep (config 4) and the live stream (config 5)
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 60 --warmup 6 --no-cpu-baseline --stage-iters 1 --secondary none > gpurun_out/bench_r02q_c3_8gpu.json 2> gpurun_out/bench_r02q_c3_8gpu.err
$TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 60 --warmup 6 --no-cpu-baseline --stage-iters 1 --secondary none > gpurun_out/bench_r02q_c3_2gpu.json 2> gpurun_out/bench_r02q_c3_2gpu.err
$TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --workload config4 --steps 10 --warmup 3 --no-cpu-baseline --stage-iters 1 --secondary none > gpurun_out/bench_r02q_c4_8gpu.json 2> gpurun_out/bench_r02q_c4_8gpu.err
$TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --workload config5 --no-cpu-baseline --secondary none > gpurun_out/bench_r02q_c5_8gpu.json 2> gpurun_out/bench_r02q_c5_8gpu.err
for f in c3_8gpu c3_2gpu c4_8gpu c5_8gpu; do
  python -c "import json; d=json.loads(open('gpurun_out/bench_r02q_$f.json').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], round(d['value'],2), round(d['ms_per_step'],4), round(d['e2e']['value'],2), {k:v for k,v in d['config'].items() if k in ('dm_trial_gsamples_per_s','lost_packets','real_time','max_sustained_gsamples_per_s','target_gsamples_per_s')})" || tail -5 gpurun_out/bench_r02q_$f.err
done
nproc; free -g | head -2

#!/bin/bash
# round-2 session T (8 GPUs): config 3 and the live stream again, with CUDA_DEVICE_MAX_CONNECTIONS pinned and the warm-up in the executable
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 60 --warmup 6 --no-cpu-baseline --stage-iters 1 --secondary none > gpurun_out/bench_r02t_c3_8gpu.json 2> gpurun_out/bench_r02t_c3_8gpu.err
$TR --nproc-per-node 4 --master-port 29552 bench.py --gpus 4 --steps 60 --warmup 6 --no-cpu-baseline --stage-iters 1 --secondary none > gpurun_out/bench_r02t_c3_4gpu.json 2> gpurun_out/bench_r02t_c3_4gpu.err
$TR --nproc-per-node 8 --master-port 29553 bench.py --gpus 8 --workload config5 --no-cpu-baseline --secondary none > gpurun_out/bench_r02t_c5_8gpu.json 2> gpurun_out/bench_r02t_c5_8gpu.err
for f in c3_8gpu c3_4gpu c5_8gpu; do
  python -c "import json; d=json.loads(open('gpurun_out/bench_r02t_$f.json').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], round(d['value'],2), round(d['ms_per_step'],4), round(d['e2e']['value'],2), {k:v for k,v in d['config'].items() if k in ('lost_packets','real_time','max_sustained_gsamples_per_s','target_gsamples_per_s','received_packets')})" || tail -5 gpurun_out/bench_r02t_$f.err
done

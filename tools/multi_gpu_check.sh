#!/bin/bash
# N = 2 scaling check exactly as the driver launches it
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/scale_2.json 2> gpurun_out/scale_2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/scale_2_ref.json 2>> gpurun_out/scale_2.err
for f in scale_1 scale_2 scale_2_ref; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$f.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("$f", "n_gpus", d["n_gpus"], "value", round(d["value"],3), "e2e", round(d["e2e"]["value"],3), "ms", round(d["ms_per_step"],4), d.get("clocks"))
except Exception as e:
    print("$f failed", e)
PY
done
tail -5 gpurun_out/scale_2.err

#!/bin/bash
# A/B of two builds of the library: tools/ab_lib.sh <tag> <variant.so>
TAG=$1; VAR=$2
for lib in libsrtb_b200.so "$VAR" libsrtb_b200.so "$VAR"; do
  SRTB_B200_LIB=$lib python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/bench_${TAG}.json'));print('lib=$lib value %.2f e2e %.2f'%(d['value'],d['e2e']['value']))"
done

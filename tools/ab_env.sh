#!/bin/bash
# A/B of one environment switch on the GPU box: parity subset, then bench with VAR=0 and VAR=1, then
# an ncu launch list with VAR=1.  usage: tools/ab_env.sh <tag> <VAR> [pytest -k expr]
TAG=$1; VAR=$2; KEXPR=${3:-"fft or chain or golden or full or block or sweep"}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 300 -k "$KEXPR" 2>&1 | tail -6 | tee gpurun_out/pytest_${TAG}.log
for v in 0 1 0 1; do
  env $VAR=$v python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_$v.json 2>gpurun_out/bench_${TAG}_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_$v.json"))
print("$VAR=$v value %.2f e2e %.2f" % (d["value"], d["e2e"]["value"]), {k: round(s["ms"]*1e3,1) for k,s in d["stages"].items()})
PY
done
env $VAR=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_${TAG}.csv")) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ik][:70]].append(float(r[iv].replace(",",""))/1e3)
    except: pass
for k,v in sorted(d.items(), key=lambda kv:-sum(kv[1])): print("%-72s n=%3d avg %.2f us" % (k,len(v),sum(v)/len(v)))
PY

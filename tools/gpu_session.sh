#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, ncu launch list, one full ncu capture.
# usage: tools/gpu_session.sh <tag> [pytest -k expr]
TAG=${1:-r01}
KEXPR=${2:-}
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout 300 ${KEXPR:+-k "$KEXPR"} 2>&1 | tee gpurun_out/pytest_gpu_${TAG}.log | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_${TAG}.log
python bench.py --steps 200 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; tail -c 600 gpurun_out/bench_ref_${TAG}.json
# launch list (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 > gpurun_out/ncu_bench_${TAG}.log 2>&1
# full capture of the FFT pass kernels + the heaviest elementwise kernels of one step
ncu --set full --clock-control none --import-source on -k regex:"fft_.*kernel|dedisperse_kernel" \
    -s 36 -c 10 -o gpurun_out/prof_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --stage-iters 1 > gpurun_out/ncu_full_${TAG}.log 2>&1
# per-stage capture: every per-pipe kernel once + one fused block (feeds profiles/traffic.json)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -o gpurun_out/prof_stages_${TAG} -f python tools/stage_once.py > gpurun_out/ncu_stages_${TAG}.log 2>&1
du -sh gpurun_out; ls -la gpurun_out | tail -12

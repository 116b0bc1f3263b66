#!/bin/bash
# round-2 session X: compute-sanitizer (memcheck, racecheck) over the kernels new in round 2, small parity cases
nvidia-smi -L
K='(chain_vs_oracle and (20-64-10 or 20-32-56 or 20-256-10 or 18-128 or 21-8-3)) or (fused_chirp_waterfall_vs_float64 and 20-32-562) or ring_returns or ring_ticket or (sk_v1_vs_oracle and 256-64) or (signal_detect_v1_vs_oracle and 512-64) or (packed_samples and 2]) or (watfft_layout)'
for tool in memcheck racecheck; do
  timeout 330 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "$K" > gpurun_out/sanitize_r02x_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|error|Error" gpurun_out/sanitize_r02x_$tool.log | tail -5
done

#!/bin/bash
# throughput of the C++ pipe mirror (tests/cpp/pipeline_main) on 16 pinned blocks of 2^24 8-bit samples replayed 16 times (H2D + chain + D2H per block):
# per-stage pipes (thread per pipe), stream-ordered composite, and the fused chain pipe on 1/3/4 queues
make -C tests/cpp pipeline_main > /dev/null 2>&1
python - <<'PY'
import numpy as np
rng = np.random.default_rng(1)
blk = np.clip(np.rint(rng.standard_normal(1 << 24) * 20), -127, 127).astype(np.int8)
with open("/dev/shm/srtb_bb.bin", "wb") as f:
    for i in range(16):
        f.write(np.roll(blk, i * 4099).tobytes())
PY
COMMON="--input /dev/shm/srtb_bb.bin --log2n 24 --bits -8 --format simple --channels 2048 --dm 56.778 --avg-thr 5 --sk-thr 1.05 --snr 8 --max-boxcar 256"
for mode in "--composite 0" "--composite 1" "--fused 1" "--fused 1 --ring 3" "--fused 2 --ring 3" "--fused 3" "--fused 4"; do
  echo "== $mode"; SRTB_LOG_LEVEL=1 ./tests/cpp/pipeline_main $COMMON $mode --preload 1 --repeat 16 2>&1 >/dev/null | grep pipeline_main
done
rm -f /dev/shm/srtb_bb.bin

#!/usr/bin/env python
"""profiles/traffic_<workload>.json from an ncu launch list that carries DRAM bytes:
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c N --csv \\
      --log-file gpurun_out/traffic.csv python bench.py --workload W --steps 2 --warmup 3 --contexts 1 ...
  tools/make_traffic.py W gpurun_out/traffic.csv <samples per block> <tag>
Takes the LAST complete block of the capture (from the first-sweep kernel that reads raw bytes to the boxcar kernel of
the last stream) and sums DRAM read + write bytes of its kernels; bench.py reports it as roofline.chain.dram_measured."""
import csv
import json
import re
import sys
from pathlib import Path

wname, path, samples, tag = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
lines = [l for l in open(path) if not l.startswith("==")]
rows = {}
order = []
for x in csv.DictReader(lines):
    i = int(x["ID"])
    if i not in rows:
        rows[i] = {"kernel": x["Kernel Name"]}
        order.append(i)
    v = float(x["Metric Value"].replace(",", ""))
    unit = x["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(unit, 1)
    rows[i][x["Metric Name"]] = v * scale
ks = [rows[i] for i in order]
# blocks end with detect_boxcar_kernel of the last stream; a block starts at a RAW first sweep (template ends "..., 1>" / "2>" / "3>")
is_raw = lambda k: re.search(r"fft_col(16)?_tma_kernel<.*, [123](, 0)?>", k["kernel"]) is not None
ends = [i for i, k in enumerate(ks) if "detect_boxcar_kernel" in k["kernel"]]
starts = [i for i, k in enumerate(ks) if is_raw(k)]
streams = 2 if wname == "config3" else 1
end = ends[-1]
start = [s for s in starts if s < end][-streams]
blk = ks[start:end + 1]
tot = sum(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0) for k in blk)
out = {"workload": wname, "tag": tag, "block_samples": samples, "block_dram_bytes": tot,
       "kernels": [{"kernel": k["kernel"][:100], "time_us": k.get("gpu__time_duration.sum"),
                    "dram_read": k.get("dram__bytes_read.sum"), "dram_write": k.get("dram__bytes_write.sum")} for k in blk]}
p = Path(__file__).resolve().parent.parent / "profiles" / f"traffic_{wname}.json"
p.write_text(json.dumps(out, indent=1))
print(p, "block DRAM bytes", tot, "=", tot / samples, "B/sample over", len(blk), "kernels")

#!/usr/bin/env python
"""Turn one GPU session's scratch output (gpurun_out/) into the tracked evidence under profiles/:
  profiles/<tag>_launches.csv   the ncu launch list (gpu__time_duration.sum, --clock-control none)
  profiles/<tag>_bench.json     the bench.py JSON line of the same session
  profiles/<tag>_summary.md     per-kernel share of a step + key metrics of the ncu --set full capture
usage: tools/summarize_profiles.py <tag>"""
import collections
import csv
import json
import re
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf"),
]


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("srtb_b200::", "")
    return name[:70]


def main():
    tag = sys.argv[1]
    PROF.mkdir(exist_ok=True)
    lines = [f"# profile summary {tag}", ""]
    bench = OUT / f"bench_{tag}.json"
    if bench.exists():
        shutil.copy(bench, PROF / f"{tag}_bench.json")
        try:
            d = json.loads(bench.read_text().strip().splitlines()[-1])
            lines += [f"bench: **{d['value']:.2f} {d['unit']}** device-resident, e2e {d['e2e']['value']:.2f}, "
                      f"{d['ms_per_step']:.4f} ms/step, {d['gpu_launches']} launches in {d['steps']} steps; "
                      f"clocks {d.get('clocks')}", "",
                      "| stage | ms (CUDA events, L2 flushed) | algorithmic MB | GB/s | frac of measured HBM peak |",
                      "|---|---|---|---|---|"]
            for k, v in d.get("stages", {}).items():
                lines.append(f"| {k} | {v['ms']:.4f} | {v['bytes'] / 1e6:.1f} | {v['gbs']:.0f} | {v['frac']:.3f} |")
            lines.append("")
        except Exception as e:  # noqa
            lines.append(f"(bench json unreadable: {e})")
    launches = OUT / f"launches_{tag}.csv"
    if launches.exists():
        shutil.copy(launches, PROF / f"{tag}_launches.csv")
        rows = list(csv.DictReader(l for l in launches.open() if not l.startswith("==")))
        agg = collections.OrderedDict()
        for r in rows:
            v = float(r["Metric Value"].replace(",", ""))
            v = v / 1000 if r["Metric Unit"] == "ns" else (v * 1000 if r["Metric Unit"] == "ms" else v)
            agg.setdefault(short(r["Kernel Name"]), []).append(v)
        ours = {k: v for k, v in agg.items() if not k.startswith("at::")}
        tot = sum(sum(v) for v in ours.values())
        lines += ["## launch list (ncu gpu__time_duration.sum, cold cache, serialised — compare shares)", "",
                  "| kernel | launches | avg us | share of our kernels |", "|---|---|---|---|"]
        for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
            lines.append(f"| `{k}` | {len(v)} | {sum(v) / len(v):.2f} | {100 * sum(v) / tot:.1f}% |")
        lines.append("")
    rep = OUT / f"prof_{tag}.ncu-rep"
    if rep.exists():
        raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) > 2:
            hdr, units, data = rows[0], rows[1], rows[2:]
            idx = {h: i for i, h in enumerate(hdr)}
            lines += ["## ncu --set full (one capture per kernel; per launch)", "",
                      "| kernel | " + " | ".join(m[1] for m in METRICS) + " |", "|---|" + "---|" * len(METRICS)]
            seen = set()
            for r in data:
                name = short(r[idx["Kernel Name"]])
                if name in seen:
                    continue
                seen.add(name)
                cells = []
                for m, _ in METRICS:
                    if m in idx:
                        val = r[idx[m]]
                        try:
                            f = float(val.replace(",", ""))
                            val = f"{f:.3g}" if f < 1e6 else f"{f:.4g}"
                        except ValueError:
                            pass
                        cells.append(f"{val} {units[idx[m]]}".strip())
                    else:
                        cells.append("-")
                lines.append(f"| `{name}` | " + " | ".join(cells) + " |")
            lines.append("")
    # per-kernel DRAM traffic of the dedicated per-stage capture -> profiles/traffic.json (bench.py reads it)
    srep = OUT / f"prof_stages_{tag}.ncu-rep"
    if srep.exists():
        raw = subprocess.run(["ncu", "-i", str(srep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) > 2:
            hdr, units, data = rows[0], rows[1], rows[2:]
            idx = {h: i for i, h in enumerate(hdr)}
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            traffic = []
            lines += ["## per-stage capture (ncu dram byte counters, L2 flushed between stages; one launch each)", "",
                      "| # | kernel | time us | dram read MB | dram write MB |", "|---|---|---|---|---|"]
            for n_, r in enumerate(data):
                rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * scale.get(units[idx["dram__bytes_read.sum"]], 1)
                wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * scale.get(units[idx["dram__bytes_write.sum"]], 1)
                t = r[idx["gpu__time_duration.sum"]]
                name = short(r[idx["Kernel Name"]])
                traffic.append({"kernel": name, "dram_read": rd, "dram_write": wr, "time_us": float(t.replace(",", ""))})
                lines.append(f"| {n_} | `{name}` | {t} | {rd / 1e6:.1f} | {wr / 1e6:.1f} |")
            lines.append("")
            (PROF / f"{tag}_traffic.json").write_text(json.dumps(traffic, indent=1))
            (PROF / "traffic.json").write_text(json.dumps({"tag": tag, "kernels": traffic}, indent=1))
    (PROF / f"{tag}_summary.md").write_text("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()

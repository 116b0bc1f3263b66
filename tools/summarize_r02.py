#!/usr/bin/env python
"""Round-2 profile summaries: tools/summarize_r02.py <tag> [--launches csv] [--bench json ...] [--ncu rep ...] [--note text]
writes profiles/<tag>_summary.md and copies the launch list / bench lines next to it (gpurun_out/ is scratch)."""
import argparse
import csv
import json
import shutil
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PROF = ROOT / "profiles"
KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB read"), ("dram__bytes_write.sum", "MB written"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("smsp__inst_executed.sum", "warp instructions"), ("launch__registers_per_thread", "registers"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]


def launches_table(path, first, count):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [(x["Kernel Name"], float(x["Metric Value"]) / 1000) for x in csv.DictReader(lines)]
    rows = rows[first:first + count] if count else rows[first:]
    out = ["| # | kernel | us |", "|---|---|---|"]
    for i, (n, t) in enumerate(rows):
        out.append(f"| {first + i} | `{n[:110]}` | {t:.1f} |")
    out.append(f"| | **sum** | **{sum(t for _, t in rows):.1f}** |")
    return out


def ncu_table(rep):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    out = ["| kernel | " + " | ".join(k for _, k in KEYS) + " |", "|---|" + "---|" * len(KEYS)]
    for r in rows[2:]:
        vals = []
        for key, _ in KEYS:
            try:
                vals.append(r[hdr.index(key)])
            except ValueError:
                vals.append("-")
        out.append(f"| `{r[hdr.index('Kernel Name')][:80]}` | " + " | ".join(vals) + " |")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--count", type=int, default=0)
    ap.add_argument("--bench", nargs="*", default=[])
    ap.add_argument("--ncu", nargs="*", default=[])
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    PROF.mkdir(exist_ok=True)
    md = [f"# {a.tag}", "", a.note, ""]
    for b in a.bench:
        d = json.loads(open(b).read().strip().splitlines()[-1])
        dst = PROF / f"{a.tag}_{Path(b).stem.split('_', 2)[-1]}.json"
        dst.write_text(json.dumps(d) + "\n")
        md += [f"## bench line `{dst.name}`", "",
               f"- workload: {d['config'].get('workload')}",
               f"- value {d['value']:.2f} {d['unit']} ({d['ms_per_step']:.4f} ms/step, {d.get('gpu_launches')} launches / {d['steps']} steps), "
               f"e2e {d['e2e']['value']:.2f}"]
        if d.get("single_context"):
            md.append(f"- single context: {d['single_context']['value']:.2f}")
        if d.get("secondary"):
            md.append(f"- secondary {d['secondary']['workload'].split(':')[0]}: {d['secondary']['value']:.2f} (e2e {d['secondary']['e2e']['value']:.2f})")
        if d.get("roofline") and d["roofline"].get("fused"):
            md.append("- fused groups (CUDA events inside the library, L2 flushed): " + ", ".join(
                f"{k} {v['ms'] * 1e3:.0f} us = {v['gbs']:.0f} GB/s ({v['frac']:.2f} of peak)" for k, v in d["roofline"]["fused"].items()))
        if d.get("clocks"):
            md.append(f"- clocks: {d['clocks']}")
        md.append("")
    if a.launches:
        dst = PROF / f"{a.tag}_launches.csv"
        shutil.copy(a.launches, dst)
        md += [f"## launch list (`{dst.name}`, ncu gpu__time_duration, cold cache, serialised)", ""]
        md += launches_table(a.launches, a.first, a.count) + [""]
    for rep in a.ncu:
        md += [f"## ncu --set full: `{Path(rep).name}`", ""] + ncu_table(rep) + [""]
    (PROF / f"{a.tag}_summary.md").write_text("\n".join(md))
    print("wrote", PROF / f"{a.tag}_summary.md")


if __name__ == "__main__":
    main()

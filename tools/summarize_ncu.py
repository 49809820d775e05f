#!/usr/bin/env python
"""Turns the raw ncu artefacts of one GPU session (gpurun_out/) into the small tracked files
under profiles/:  <tag>_launches.csv (the launch list as captured), <tag>_summary.json and
<tag>_summary.md (per-kernel time shares from the launch list; key metrics, stall reasons,
pipe utilisation and the hottest SASS regions of the full capture)."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


def launch_shares(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    per = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r[4].split("(")[0]
        per[name][0] += 1
        per[name][1] += num(r[-1])
    total = sum(v[1] for v in per.values()) or 1.0
    return {k: dict(launches=v[0], total_ms=v[1] / 1e6, share=v[1] / total) for k, v in per.items()}


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True)
    return list(csv.reader(out.stdout.splitlines()))


def main():
    tag = sys.argv[1]
    out_dir = os.path.join(ROOT, "profiles")
    src = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    summary = {"tag": tag}
    launches = os.path.join(src, f"launches_{tag}.csv")
    if os.path.exists(launches):
        shutil.copy(launches, os.path.join(out_dir, f"{tag}_launches.csv"))
        summary["launch_list"] = launch_shares(launches)
    rep = os.path.join(src, f"prof_keyed_{tag}.ncu-rep")
    if os.path.exists(rep):
        raw = ncu_csv(rep, "raw")
        hdr, units, vals = raw[0], raw[1], raw[2]
        metrics = dict(zip(hdr, zip(units, vals)))
        keep = {}
        for key in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
                    "sm__warps_active.avg.pct_of_peak_sustained_active",
                    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
                    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
                    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
                    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
                    "smsp__thread_inst_executed_per_inst_executed.ratio",
                    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
                    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum"):
            if key in metrics:
                keep[key] = {"unit": metrics[key][0], "value": metrics[key][1]}
        stalls = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): num(v[1])
                  for k, v in metrics.items() if k.startswith("smsp__average_warps_issue_stalled_")}
        summary["full_capture"] = {"metrics": keep, "stalls_per_issue": dict(sorted(stalls.items(), key=lambda kv: -kv[1]))}

        def as_bytes(key):
            unit, value = metrics.get(key, ("", "0"))
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            return num(value) * scale
        summary["dram_bytes_per_launch"] = as_bytes("dram__bytes_read.sum") + as_bytes("dram__bytes_write.sum")
        # hottest SASS regions
        src_rows = ncu_csv(rep, "source")
        h = src_rows[1]
        data = [dict(zip(h, r)) for r in src_rows[2:] if len(r) == len(h)]
        total = sum(num(d["Instructions Executed"]) for d in data) or 1.0
        ops = collections.Counter()
        for d in data:
            text = re.sub(r"^@!?U?P\d+\s+", "", d["Source"].strip())
            ops[text.split()[0].split(".")[0] if text else "?"] += num(d["Instructions Executed"])
        summary["full_capture"]["opcode_mix"] = {k: v / total for k, v in ops.most_common(14)}
        regions = []
        for i in range(0, len(data), 64):
            chunk = data[i:i + 64]
            c = sum(num(d["Instructions Executed"]) for d in chunk)
            t = sum(num(d["Thread Instructions Executed"]) for d in chunk)
            if c / total > 0.01:
                regions.append(dict(first_sass_index=i, share=c / total, avg_lanes=t / c,
                                    first_instruction=chunk[0]["Source"].strip()))
        summary["full_capture"]["hot_regions_64_instr"] = regions
    json.dump(summary, open(os.path.join(out_dir, f"{tag}_summary.json"), "w"), indent=1)
    with open(os.path.join(out_dir, f"{tag}_summary.md"), "w") as md:
        md.write(f"# ncu summary {tag}\n\n")
        if "launch_list" in summary:
            md.write("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, "
                     "`python bench.py --steps 2 --warmup 1 --no-cpu-baseline`)\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
            for k, v in sorted(summary["launch_list"].items(), key=lambda kv: -kv[1]["share"]):
                md.write(f"| `{k}` | {v['launches']} | {v['total_ms']:.3f} | {v['share'] * 100:.2f} % |\n")
        if "full_capture" in summary:
            fc = summary["full_capture"]
            md.write("\n## Full capture of the megakernel (`ncu --set full --clock-control none --import-source on`)\n\n")
            for k, v in fc["metrics"].items():
                md.write(f"* `{k}` = {v['value']} {v['unit']}\n")
            md.write(f"* DRAM bytes per launch (read + write) = {summary['dram_bytes_per_launch']:.4g}\n")
            md.write("\n### Stall reasons (warps per issue-active cycle)\n\n")
            for k, v in fc["stalls_per_issue"].items():
                md.write(f"* {k}: {v:.3f}\n")
            md.write("\n### Opcode mix (share of executed warp instructions)\n\n")
            for k, v in fc["opcode_mix"].items():
                md.write(f"* {k}: {v * 100:.1f} %\n")
            md.write("\n### Hot SASS regions (64-instruction chunks > 1 % of executed instructions)\n\n| first SASS index | share | avg active lanes | first instruction |\n|---|---|---|---|\n")
            for r in fc["hot_regions_64_instr"]:
                md.write(f"| {r['first_sass_index']} | {r['share'] * 100:.1f} % | {r['avg_lanes']:.1f} | `{r['first_instruction'][:60]}` |\n")
    print("wrote", os.path.join(out_dir, f"{tag}_summary.md"))


if __name__ == "__main__":
    main()

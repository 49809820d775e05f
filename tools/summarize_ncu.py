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
    """Per kernel: launches, total time and share; DRAM bytes when the list carries those metrics
    (one CSV row per launch and metric: ..., metric name, unit, value)."""
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    per = collections.defaultdict(lambda: dict(launches=0, ns=0.0, dram_read=0.0, dram_write=0.0))
    for r in rows:
        name, metric, value = r[4].split("(")[0], r[-3], num(r[-1])
        if metric == "gpu__time_duration.sum":
            per[name]["launches"] += 1
            per[name]["ns"] += value
        elif metric == "dram__bytes_read.sum":
            per[name]["dram_read"] += value
        elif metric == "dram__bytes_write.sum":
            per[name]["dram_write"] += value
    total = sum(v["ns"] for v in per.values()) or 1.0
    return {k: dict(launches=v["launches"], total_ms=v["ns"] / 1e6, share=v["ns"] / total,
                    dram_read_bytes=v["dram_read"], dram_write_bytes=v["dram_write"]) for k, v in per.items()}


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True)
    return list(csv.reader(out.stdout.splitlines()))


def main():
    tag = sys.argv[1]
    label = sys.argv[2] if len(sys.argv) > 2 else ""  # e.g. "suzanne": gpurun_out/prof_suzanne_<tag>.ncu-rep only
    out_dir = os.path.join(ROOT, "profiles")
    src = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    summary = {"tag": tag}
    launches = os.path.join(src, f"launches_{tag}.csv")
    if os.path.exists(launches) and not label:
        shutil.copy(launches, os.path.join(out_dir, f"{tag}_launches.csv"))
        summary["launch_list"] = launch_shares(launches)
    rep = os.path.join(src, f"prof_{label or 'subpath'}_{tag}.ncu-rep")
    if not os.path.exists(rep):
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
                    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
                    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
                    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
                    "smsp__warps_eligible.avg.per_cycle_active",
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
    if label:
        tag = f"{tag}_{label}"
    json.dump(summary, open(os.path.join(out_dir, f"{tag}_summary.json"), "w"), indent=1)
    # what bench.py quotes next to its live numbers (roofline.traffic, roofline.ncu)
    quoted_path = os.path.join(out_dir, "ncu_summary.json")
    quoted = json.load(open(quoted_path)) if os.path.exists(quoted_path) else {}
    ours = {k: v for k, v in summary.get("launch_list", {}).items() if "ptb200::" in k and ("Hits" in k or "subPath" in k or "resolve" in k)}
    if ours and any(v["dram_read_bytes"] or v["dram_write_bytes"] for v in ours.values()):
        steps = max(1, int(os.environ.get("NCU_STEPS", "2")))  # --steps 1: the value loop and the e2e loop each render once
        quoted.setdefault("dram_bytes_per_step", {})["1"] = {
            "bytes": sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in ours.values()) / steps,
            "per_kernel_per_launch": {k: (v["dram_read_bytes"] + v["dram_write_bytes"]) / max(1, v["launches"]) for k, v in ours.items()},
            "source": f"profiles/{tag}_launches.csv: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every "
                      f"path-tracing launch of one full-size step (640x480 @ 256 spp)"}
    if "full_capture" in summary and not label:
        m = summary["full_capture"]["metrics"]
        quoted["latest_full_capture"] = {
            "source": f"profiles/{tag}_summary.json (ncu --set full of {m['Kernel Name']['value'].strip()}, BENCH_SPP=16 launch of the bench workload)",
            "issue_slots_busy_pct": num(m["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"]),
            "fp64_pipe_active_pct": num(m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]["value"]),
            "fma_pipe_pct": num(m["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]["value"]),
            "alu_pipe_pct": num(m["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]["value"]),
            "active_lanes_per_warp_instruction": num(m["smsp__thread_inst_executed_per_inst_executed.ratio"]["value"]),
            "warps_active_pct_of_peak": num(m["sm__warps_active.avg.pct_of_peak_sustained_active"]["value"]),
            "registers_per_thread": num(m["launch__registers_per_thread"]["value"])}
    quoted.pop("dram_bytes_per_launch", None)
    if not label:
        json.dump(quoted, open(quoted_path, "w"), indent=1)
    with open(os.path.join(out_dir, f"{tag}_summary.md"), "w") as md:
        md.write(f"# ncu summary {tag}\n\n")
        if "launch_list" in summary:
            md.write("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, "
                     "`python bench.py --steps 1 --warmup 0 --no-cpu-baseline`)\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
            for k, v in sorted(summary["launch_list"].items(), key=lambda kv: -kv[1]["share"]):
                md.write(f"| `{k}` | {v['launches']} | {v['total_ms']:.3f} | {v['share'] * 100:.2f} % |\n")
            if any(v["dram_read_bytes"] or v["dram_write_bytes"] for v in summary["launch_list"].values()):
                md.write("\nDRAM traffic of the same launches (`dram__bytes_read.sum`, `dram__bytes_write.sum`):\n\n"
                         "| kernel | read GB | written GB | per launch MB |\n|---|---|---|---|\n")
                for k, v in sorted(summary["launch_list"].items(), key=lambda kv: -kv[1]["share"]):
                    both = v["dram_read_bytes"] + v["dram_write_bytes"]
                    md.write(f"| `{k}` | {v['dram_read_bytes'] / 1e9:.3f} | {v['dram_write_bytes'] / 1e9:.3f} | "
                             f"{both / max(1, v['launches']) / 1e6:.1f} |\n")
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

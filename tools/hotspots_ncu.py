#!/usr/bin/env python
"""Source-level hot spots of one full ncu capture: splits the kernel's SASS into regions of equal
execution frequency (loop bodies, branches), and reports per region its share of executed warp
instructions, its share of warp-state samples (≈ time), active lanes and the dominant stall reasons;
then the individual instructions with the most samples.

    python tools/hotspots_ncu.py <tag>      reads gpurun_out/prof_keyed_<tag>.ncu-rep
                                            writes profiles/<tag>_hotspots.md
"""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tag = sys.argv[1]
    rep = os.path.join(ROOT, "gpurun_out", f"prof_keyed_{tag}.ncu-rep")
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernel, hdr, data = rows[0][1], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def num(r, h):
        try:
            return float(r[col[h]] or 0)
        except ValueError:
            return 0.0

    total_inst = sum(num(r, "Instructions Executed") for r in data) or 1.0
    total_samples = sum(num(r, "# Samples") for r in data) or 1.0
    # one "iteration" = one trip of the megakernel's main loop = executions of the hottest
    # full-warp instruction outside the inner loops: take the sweep's first F2F conversion
    iters = next((num(r, "Instructions Executed") for r in data if "F2F.F32.F64" in r[col["Source"]]), 1.0)

    regions, cur = [], None
    for i, r in enumerate(data):
        e = num(r, "Instructions Executed")
        if cur and abs(cur["e"] - e) <= 0.02 * max(e, cur["e"]):
            cur["rows"].append(i)
        else:
            cur = {"e": e, "rows": [i]}
            regions.append(cur)

    lines = [f"# Source-level hot spots, capture {tag}", "", f"Kernel: `{kernel}`", "",
             f"{total_inst:.4g} warp instructions, {total_samples:.4g} warp-state samples, "
             f"{iters:.4g} main-loop iterations (≈31 casts each).", "",
             "## Regions of equal execution frequency (> 0.8 % of the samples or of the instructions)", "",
             "| SASS lines | instr. | trips / iteration | warp instr. / iteration | lanes | % instr. | % samples | main stalls | first instruction |",
             "|---|---|---|---|---|---|---|---|---|"]
    for reg in regions:
        rs = [data[i] for i in reg["rows"]]
        inst = sum(num(r, "Instructions Executed") for r in rs)
        thr = sum(num(r, "Thread Instructions Executed") for r in rs)
        smp = sum(num(r, "# Samples") for r in rs)
        if inst / total_inst < 0.008 and smp / total_samples < 0.008:
            continue
        stalls = sorted(((sum(num(r, h) for r in rs), h[6:]) for h in stall_cols), reverse=True)
        top = ", ".join(f"{name} {100 * v / smp:.0f}%" for v, name in stalls[:3] if smp and v / smp > 0.08)
        first = rs[0][col["Source"]].strip()[:48]
        lines.append(f"| {reg['rows'][0]}–{reg['rows'][-1]} | {len(rs)} | {reg['e'] / iters:.2f} | "
                     f"{inst / iters:.0f} | {thr / inst if inst else 0:.1f} | {100 * inst / total_inst:.1f} | "
                     f"{100 * smp / total_samples:.1f} | {top} | `{first}` |")
    lines += ["", "## Instructions with the most samples", "",
              "| SASS line | % samples | trips / iteration | lanes | main stall | instruction |", "|---|---|---|---|---|---|"]
    ranked = sorted(range(len(data)), key=lambda i: -num(data[i], "# Samples"))[:30]
    for i in ranked:
        r = data[i]
        smp = num(r, "# Samples")
        e = num(r, "Instructions Executed")
        stalls = sorted(((num(r, h), h[6:]) for h in stall_cols), reverse=True)
        lines.append(f"| {i} | {100 * smp / total_samples:.2f} | {e / iters:.2f} | "
                     f"{num(r, 'Thread Instructions Executed') / e if e else 0:.1f} | {stalls[0][1]} | "
                     f"`{r[col['Source']].strip()[:70]}` |")
    path = os.path.join(ROOT, "profiles", f"{tag}_hotspots.md")
    open(path, "w").write("\n".join(lines) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()

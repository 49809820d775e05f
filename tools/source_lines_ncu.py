#!/usr/bin/env python
"""Per SOURCE LINE cost of one full ncu capture (compiled with -lineinfo, captured with
--import-source on): executed warp instructions, warp-state samples (~time) and active lanes,
aggregated over every SASS instruction the line produced, inlined copies included.

    python tools/source_lines_ncu.py gpurun_out/<name>.ncu-rep [top_n] [iterations]

`iterations` (main-loop trips of the kernel) turns instruction counts into instructions per trip."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    iters = float(sys.argv[3]) if len(sys.argv) > 3 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    current = None
    col = {}
    lines = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            current = r[1].split("/")[-1]
        elif r[0] == "Line No":
            col = {}
            for i, h in enumerate(r):
                col.setdefault(h, i)
        elif r[0] not in ("", "Function Name") and current and col:
            try:
                lines.append(dict(file=current, line=int(r[0]), text=r[1].strip(),
                                  samples=float(r[col["# Samples"]] or 0),
                                  inst=float(r[col["Instructions Executed"]] or 0),
                                  thr=float(r[col["Thread Instructions Executed"]] or 0)))
            except ValueError:
                continue
    tot_i = sum(l["inst"] for l in lines) or 1.0
    tot_s = sum(l["samples"] for l in lines) or 1.0
    print(f"{tot_i:.4g} warp instructions, {tot_s:.4g} samples, {len(lines)} source lines")
    per_file = {}
    for l in lines:
        f = per_file.setdefault(l["file"], [0.0, 0.0, 0.0])
        f[0] += l["inst"]; f[1] += l["samples"]; f[2] += l["thr"]
    for f, (i, s, t) in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:18s} {100 * i / tot_i:5.1f} % instr  {100 * s / tot_s:5.1f} % samples  lanes {t / max(i, 1):4.1f}")
    print("| file:line | % instr | instr/trip | % samples | lanes | source |")
    print("|---|---|---|---|---|---|")
    for l in sorted(lines, key=lambda l: -l["inst"])[:top]:
        per = f"{l['inst'] / iters:7.1f}" if iters else "-"
        print(f"| {l['file']}:{l['line']} | {100 * l['inst'] / tot_i:.2f} | {per} | {100 * l['samples'] / tot_s:.2f} | "
              f"{l['thr'] / max(l['inst'], 1):.1f} | `{l['text'][:90]}` |")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the hot path: Msamples/s on CornellBox-Original 640x480 @ 256 spp.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path

A *step* is one full pass of the hot path over one batch of synthetic input: one render of the
BASELINE.json workload (configs[1]).  A *sample* is one camera ray with its whole 16-way
sub-path tree, the unit ArrayOutput::totalSamples() counts (src/util/ArrayOutput.cpp:58-63,
src/main/main.cpp:464-473); Msamples/s = samples / seconds / 1e6.

Our arm, one process per GPU (torchrun for N > 1), weak scaling: every rank renders the rows
y = rank (mod N) of the 640x480 frame for 256*N passes, i.e. the same number of samples per
GPU at every N; no collective touches the data path, only a host-side final gather of rows.
  value   kernels only, scene already resident in HBM (uploaded once before the timed region),
          timed with CUDA events on the launching stream, max over ranks;
  e2e     the reference-facing C-ABI call ptb200_render() with HOST buffers each step: scene
          H2D + kernels + framebuffer D2H, plus the host-side gather for N > 1;
  roofline       the path-tracing megakernel: algorithmic sweep bytes (72 B/triangle +
          32 B/sphere per ray cast) x counted casts / its CUDA-event time, against
          the measured HBM copy bandwidth — see DESIGN.md for why this "logical" figure is far
          above 1 (the primitive list is staged once into shared memory by TMA and swept from
          there); the fp64 figure next to it is the binding one;
  cpu_baseline   the reference's dod renderer on this box's host cores, bounded sample.

The reference arm times oracle/_ref/ref_tool (the reference's own sources, unmodified
dod::Scene::render with maxCpus = all host threads) when it was built, else the oracle port.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_JSON_OUT = sys.stdout
WIDTH, HEIGHT, SPP, SEED = 640, 480, int(os.environ.get("BENCH_SPP", "256")), 1  # BENCH_SPP: profiling only
SCENE = "cornell"
METRIC = "Msamples/sec on CornellBox 640x480"
UNIT = "Msamples/s"


def scene_path(name=SCENE):
    return os.path.join(ROOT, "tests", "golden", "scenes", name + ".ptscene")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            data = json.load(open(path))
            return float(data["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows = []
        self.proc = None
        self.device_index = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device_index), f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([f.strip() for f in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()  # exact PID we started
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, flag in zip(names, r[4:8]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx),
                "power_w_max": max(power), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ---------------------------------------------------------------------------------------------
def cpu_reference_step(threads, width=256, height=192):
    """One bounded sample of the workload on the host: same scene, same aspect ratio, reduced
    resolution, spp = 2 x threads so the reference's own pass scheduler (one std::async task
    per pass, Scene.cpp:208-229) has two waves of work.  Returns (Msamples/s, info)."""
    from oracle import oracle_binding as ob
    spp = max(2, 2 * threads)
    if ob.have_ref_tool():
        t0 = time.perf_counter()
        res = ob.ref_render(scene_path(), width, height, spp, threads, SEED)
        wall = time.perf_counter() - t0
        value = res["total_samples"] / res["seconds"] / 1e6
        kept = res["total_samples"] / float(width * height)
        info = {"kind": "reference", "cores": threads,
                "sample": (f"{SCENE} {width}x{height} (same 4:3 aspect as 640x480), spp={spp}, "
                           f"unmodified dod::Scene::render maxCpus={threads}; it kept {kept:.0f} of "
                           f"{spp} passes (Scene.cpp:251 drops in-flight passes); {res['seconds']:.2f} s "
                           f"inside render, {wall:.2f} s process")}
        return value, info
    # oracle port, fair scheduling
    from pt_three_ways_b200 import scenefile
    scene = scenefile.load(scene_path())
    osc = ob.OracleScene(scene)
    t0 = time.perf_counter()
    osc.render(scene.camera(width, height), ob.params_array(width, height, spp=spp, seed=SEED),
               ob.RNG_MT19937_SEQUENTIAL, threads=threads)
    sec = time.perf_counter() - t0
    value = width * height * spp / sec / 1e6
    return value, {"kind": "port", "cores": threads,
                   "sample": f"{SCENE} {width}x{height}, spp={spp}, oracle port, {threads} threads, {sec:.2f} s"}


def oracle_fair_rate(threads, width=160, height=120):
    """The oracle restatement with passes spread fairly over all host threads, all passes kept."""
    from oracle import oracle_binding as ob
    from pt_three_ways_b200 import scenefile
    scene = scenefile.load(scene_path())
    osc = ob.OracleScene(scene)
    spp = max(2, 2 * threads)
    t0 = time.perf_counter()
    osc.render(scene.camera(width, height), ob.params_array(width, height, spp=spp, seed=SEED),
               ob.RNG_MT19937_SEQUENTIAL, threads=threads)
    return width * height * spp / (time.perf_counter() - t0) / 1e6


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # under torchrun only rank 0 runs the CPU reference
    threads = os.cpu_count() or 1
    values, infos = [], []
    for _ in range(args.warmup):
        cpu_reference_step(threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, info = cpu_reference_step(threads)
        values.append(v)
        infos.append(info)
    wall = time.perf_counter() - t0
    value = statistics.mean(values)
    info = infos[-1]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / max(1, args.steps) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (CornellBox-Original scene arrays from the reference loader, fixture)",
        "config": {"workload": "CornellBox-Original.obj 640x480 256 spp (BASELINE configs[1]); "
                               "each step is a bounded sample of it: " + info["sample"]},
        "cpu_baseline": {"value": value, "unit": UNIT, **info},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from pt_three_ways_b200 import capi, partition, scenefile

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    gloo = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = dist.new_group(backend="gloo")  # the host-side final gather

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    scene = scenefile.load(scene_path())
    marshalled = capi.MarshalledScene(scene)
    camera = scene.camera(WIDTH, HEIGHT)
    spp_total = SPP * world  # weak scaling: per-GPU samples fixed
    params = capi.make_params(WIDTH, HEIGHT, spp=spp_total, seed=SEED)
    options = capi.make_options(rng_mode=capi.RNG_KEYED_PHILOX, device=local_rank,
                                row_begin=rank, row_step=world)
    own_rows = len(range(rank, HEIGHT, world))
    samples_rank = own_rows * WIDTH * spp_total
    samples_total = WIDTH * HEIGHT * spp_total

    ctx = capi.Context(local_rank)
    ctx.upload_scene(marshalled)  # resident: "uploaded once to HBM"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    # ---- value: kernels only, inputs resident ----
    for _ in range(args.warmup):
        ctx.render(camera, params, options)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    device_ms, sweep_ms, casts, launches = 0.0, 0.0, 0, 0
    for _ in range(args.steps):
        flush.fill_(1)  # L2 flush between timed iterations (not inside the event-timed region)
        torch.cuda.synchronize()
        st = ctx.render(camera, params, options)  # CUDA events on the launching stream inside
        device_ms += st["kernel_ms"]
        sweep_ms += st["sweep_kernel_ms"]
        casts += st["casts"]
        launches += st["kernel_launches"]
    barrier()
    wall = time.perf_counter() - wall0
    sampler.stop()
    device_s = max_over_ranks(device_ms * 1e-3)
    sweep_s = max_over_ranks(sweep_ms * 1e-3)
    casts_total = sum_over_ranks(float(casts))
    launches_total = int(sum_over_ranks(float(launches)))
    value = samples_total * args.steps / device_s / 1e6

    # ---- e2e: host buffers through the C ABI, H2D + kernels + D2H (+ host gather) ----
    pixels = None
    for _ in range(min(args.warmup, 2)):
        capi.render(marshalled, camera, params, options)
    barrier()
    e2e0 = time.perf_counter()
    for _ in range(args.steps):
        pixels, _ = capi.render(marshalled, camera, params, options)
        if world > 1:  # the host-side final gather of the disjoint rows (no collective on the data path)
            frame = partition.gather_rows(pixels, rank, world, group=gloo)
            if rank == 0:
                pixels = frame
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - e2e0)
    e2e_value = samples_total * args.steps / e2e_s / 1e6
    h2d = (marshalled.tv.nbytes + marshalled.tm.nbytes + marshalled.sc.nbytes + marshalled.sm.nbytes
           + marshalled.mats.nbytes + 24 + 144 + 36)
    d2h = WIDTH * HEIGHT * 32

    if rank == 0:
        assert int(pixels["n"].min()) == spp_total and int(pixels["n"].max()) == spp_total

    # ---- roofline of the megakernel ----
    hbm_peak, peak_source = measured_peaks()
    casts_per_step_rank = casts / max(1, args.steps)
    sweep_ms_per_launch = sweep_ms / max(1, args.steps)
    logical_gbs = casts_per_step_rank * scene.sweep_bytes() / (sweep_ms_per_launch * 1e-3) / 1e9
    fp64_tflops = casts_per_step_rank * scene.sweep_flops() / (sweep_ms_per_launch * 1e-3) / 1e12
    fp64_peak, _ = capi.measure_fp64_peak(local_rank)
    traffic, ncu_capture = None, None
    ncu_summary = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(ncu_summary):
        try:
            summary = json.load(open(ncu_summary))
            traffic = summary.get("dram_bytes_per_launch")
            ncu_capture = summary.get("latest_full_capture")  # committed ncu figures, not measured now
        except Exception:
            traffic = None

    line = None
    if rank == 0:
        clocks = sampler.summary()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": device_s / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (CornellBox-Original scene arrays from the reference loader, fixture)",
            "config": {
                "workload": f"CornellBox-Original.obj {WIDTH}x{HEIGHT} {SPP} spp per GPU-share "
                            f"(BASELINE configs[1]); {world} GPU(s): rows y%{world}==rank, "
                            f"{spp_total} passes, 4x4 first bounce, maxDepth 5, seed {SEED}",
                "rng": "keyed Philox4x32-10 (parity: bit-exact vs the oracle run with the same policy)",
                "l2": "256 MB flush buffer written between timed steps; the scene is "
                      "shared-memory resident by design, the 1.9 GB sample buffer exceeds L2",
                "samples_per_step": samples_total, "casts_per_sample": casts_total / (samples_total * args.steps),
                "wall_s_timed_region": wall,
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": launches_total,
            "roofline": {
                "bound": "hbm", "achieved": logical_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": logical_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_source,
                "kernel": "renderKeyedKernel",
                "note": "achieved = counted casts x (72 B/triangle + 32 B/sphere) / CUDA-event time "
                        "of the megakernel per launch: LOGICAL sweep bytes, served from shared memory "
                        "after one TMA stage per CTA, hence >> HBM peak; the binding roof is fp64",
                "fp64": {"achieved_tflops": fp64_tflops, "peak_tflops": fp64_peak,
                         "frac": fp64_tflops / fp64_peak if fp64_peak else None,
                         "peak_source": "self-measured DFMA loop (ptb200_measure_fp64_peak)",
                         "flops_per_cast": scene.sweep_flops()},
                "ms_per_launch": sweep_ms_per_launch, "bytes_per_cast": scene.sweep_bytes(),
                "ncu": ncu_capture,
            },
        }
        if world == 1 and not args.no_cpu_baseline:
            # The other RNG mode, for the record: the reference's exact mt19937 stream (one pass
            # per warp, parallel over passes only) on a bounded sample of the same workload.
            ew, eh = 160, 120
            est = ctx.render(scene.camera(ew, eh), capi.make_params(ew, eh, spp=SPP, seed=SEED),
                             capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL, device=local_rank))
            line["exact_stream_mode"] = {
                "value": est["samples"] / est["kernel_ms"] / 1e3, "unit": UNIT,
                "sample": f"{SCENE} {ew}x{eh} (same aspect), {SPP} passes, PTB200_RNG_MT19937_SEQUENTIAL",
                "parity": "bit-exact vs the oracle's sequential policy, which equals the reference's "
                          "per-pass images (tests/test_oracle_golden.py)",
                "note": "parallel over passes only (SURVEY.md section 0 item 3); not the timed headline"}
            # ... and the reference's `fp` way (mt19937 per pass and pixel): exact AND parallel
            # over pixels, so it runs in the megakernel on the full workload.
            fst = ctx.render(camera, capi.make_params(WIDTH, HEIGHT, spp=SPP, seed=SEED),
                             capi.make_options(rng_mode=capi.RNG_MT19937_PER_PIXEL, device=local_rank))
            line["fp_way_mode"] = {
                "value": fst["samples"] / fst["kernel_ms"] / 1e3, "unit": UNIT,
                "sample": f"{SCENE} {WIDTH}x{HEIGHT}, {SPP} passes, PTB200_RNG_MT19937_PER_PIXEL",
                "casts_per_sample": fst["casts"] / fst["samples"],
                "parity": "bit-exact vs the oracle's fp policy, which equals fp::render of the reference "
                          "(src/fp/Render.cpp) image for image (tests/golden/fp_pass_*.npy)",
                "note": "the reference's --way fp semantics; not the timed headline"}
            # ... and its `oo` way: the sequential stream again with the oo estimator.
            ost = ctx.render(scene.camera(ew, eh), capi.make_params(ew, eh, spp=SPP, seed=SEED),
                             capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL_OO, device=local_rank))
            line["oo_way_mode"] = {
                "value": ost["samples"] / ost["kernel_ms"] / 1e3, "unit": UNIT,
                "sample": f"{SCENE} {ew}x{eh} (same aspect), {SPP} passes, PTB200_RNG_MT19937_SEQUENTIAL_OO",
                "parity": "bit-exact vs the oracle's oo policy, which equals oo::Renderer::radiance of the "
                          "reference (src/oo/Renderer.cpp) image for image (tests/golden/oo_pass_*.npy)",
                "note": "the reference's --way oo semantics, parallel over passes only; not the timed headline"}
            threads = os.cpu_count() or 1
            v, info = cpu_reference_step(threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, **info,
                                    "oracle_fair_value": oracle_fair_rate(threads)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    return 0


def main():
    # Rank 0 must print exactly ONE JSON line on stdout.  Libraries underneath (NCCL prints
    # "NCCL version ..." on stdout at the first collective) must not add to it: park the real
    # stdout, point fd 1 at stderr for the whole run, and write the JSON line to the parked fd.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Benchmark of the hot path: Msamples/s on CornellBox-Original 640x480 @ 256 spp.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path
    python bench.py --config {1,2,3,4} ...                   # another BASELINE.json workload

A *step* is one full pass of the hot path over one batch of synthetic input: one render of the
workload (default BASELINE.json configs[1]).  A *sample* is one camera ray with its whole 16-way
sub-path tree, the unit ArrayOutput::totalSamples() counts (src/util/ArrayOutput.cpp:58-63,
src/main/main.cpp:464-473); Msamples/s = samples / seconds / 1e6.

Our arm, one process per GPU (torchrun for N > 1), STRONG scaling: the workload is fixed and its
framebuffer is tile-partitioned — rank r renders the rows y = r (mod N) for all passes; no
collective touches the data path, the final gather is every rank's D2H landing in one shared
host frame.
  value         kernels only, scene already resident in HBM (uploaded once before the timed
                region), timed with CUDA events on the launching stream, max over ranks;
  e2e           the reference-facing C-ABI call ptb200_render() with HOST buffers each step:
                scene H2D + kernels + framebuffer D2H of the rank's rows into the shared frame;
  roofline      the path-tracing kernels against the roof that binds them: algorithmic
                fp64 flops of the reference's sweep (46 flop/triangle + 16 flop/sphere per ray cast)
                x casts counted on the device / their CUDA-event time, against a DFMA peak measured
                in the same run; `hbm` carries the logical sweep bandwidth north_star asks for
                (72 B/triangle + 32 B/sphere per cast; served from shared memory after one TMA
                stage per CTA, hence far above the HBM peak) and the measured DRAM traffic;
  weak_scaling  (N > 1) the per-GPU share of N = 1 kept fixed: 256*N passes of the same frame;
  config4       (N = 8, config 1) BASELINE configs[4]: CornellBox 1920x1080 @ 4096 spp over 8 GPUs;
  cpu_baseline  the reference's dod renderer on this box's host cores, bounded sample.

The reference arm times oracle/_ref/ref_tool — the reference's own sources: its radiance() and
Camera::randomRay() for every pass, passes spread over all host threads and all of them kept
(`passes`); the unmodified dod::Scene::render entry point, which abandons the passes still in
flight when the last one is launched (Scene.cpp:251), is reported next to it (`as_is`).
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
SEED = 1
METRIC = "Msamples/sec on CornellBox 640x480"
UNIT = "Msamples/s"
# BASELINE.json configs[1..4]; BENCH_SPP shortens a run for profiling (never for a reported value)
WORKLOADS = {
    1: dict(scene="cornell", width=640, height=480, spp=256,
            name="CornellBox-Original.obj 640x480 256 spp (BASELINE configs[1])"),
    2: dict(scene="suzanne", width=640, height=480, spp=256,
            name="suzanne.obj 640x480 256 spp (BASELINE configs[2])"),
    3: dict(scene="ce", width=1280, height=720, spp=1024,
            name="ce.obj 1280x720 1024 spp (BASELINE configs[3])"),
    4: dict(scene="cornell", width=1920, height=1080, spp=4096,
            name="CornellBox-Original.obj 1920x1080 4096 spp, framebuffer tiled across the GPUs (BASELINE configs[4])"),
}
REF_FLAGS = "g++ -std=c++17 -O3 -DNDEBUG -march=x86-64-v3 -funsafe-math-optimizations (CMakeLists.txt:21 + Release; x86-64-v3 for -march=native)"


def workload(config):
    w = dict(WORKLOADS[config])
    if os.environ.get("BENCH_SPP"):
        w["spp"] = int(os.environ["BENCH_SPP"])
        w["name"] += f" [BENCH_SPP={w['spp']}: shortened, not a reportable value]"
    return w


def scene_path(name):
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


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows = []
        self.window = []
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
            self.rows.append((time.time(), [f.strip() for f in line.split(",")]))

    def mark(self):
        """Wall-clock window of the load: the first call opens it, the second closes it.  nvidia-smi
        itself is started at process start (it needs ~1 s before its first sample, longer than a
        whole 8-GPU timed region)."""
        self.window.append(time.time())

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
        lo = self.window[0] if self.window else 0.0
        hi = self.window[1] if len(self.window) > 1 else float("inf")
        inside = [row for row in self.rows if lo <= row[0] <= hi + 0.25]
        for stamp, r in inside or self.rows:  # (a window shorter than one sampling period: whole process)
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
def cpu_reference_step(threads, work, with_as_is=True):
    """One bounded sample of the workload on the host: same scene, same aspect ratio, reduced
    resolution, spp = 2 x threads.  Returns (Msamples/s, info)."""
    from oracle import oracle_binding as ob
    scene = work["scene"]
    height = {"cornell": 192, "suzanne": 96, "ce": 27}[scene]
    width = height * work["width"] // work["height"]
    spp = max(2, 2 * threads) if scene != "ce" else max(2, threads)
    base = {"cores": threads, "cpu": cpu_model(), "compiler_flags": REF_FLAGS}
    if ob.have_ref_tool():
        res = ob.ref_passes(scene_path(scene), width, height, spp, threads, SEED)
        value = res["total_samples"] / res["seconds"] / 1e6
        info = dict(base, kind="reference",
                    sample=(f"{scene} {width}x{height} (aspect of {work['width']}x{work['height']}), spp={spp}: the "
                            f"reference's own radiance()/Camera::randomRay() (compiled from its sources) for every "
                            f"pass, passes spread over {threads} threads, all kept; {res['seconds']:.2f} s"))
        if with_as_is:
            asis = ob.ref_render(scene_path(scene), width, height, spp, threads, SEED)
            kept = asis["total_samples"] / float(width * height)
            info["as_is"] = {
                "value": asis["total_samples"] / asis["seconds"] / 1e6, "unit": UNIT,
                "note": (f"unmodified dod::Scene::render, maxCpus={threads}: kept {kept:.0f} of {spp} passes "
                         f"(Scene.cpp:251 abandons the passes in flight), {asis['seconds']:.2f} s; samples "
                         f"counted as main.cpp:464-473 does")}
        return value, info
    # oracle port, fair scheduling
    from pt_three_ways_b200 import scenefile
    loaded = scenefile.load(scene_path(scene))
    osc = ob.OracleScene(loaded)
    t0 = time.perf_counter()
    osc.render(loaded.camera(width, height), ob.params_array(width, height, spp=spp, seed=SEED),
               ob.RNG_MT19937_SEQUENTIAL, threads=threads)
    sec = time.perf_counter() - t0
    value = width * height * spp / sec / 1e6
    return value, dict(base, kind="port", compiler_flags="g++ -O2 -march=x86-64-v3 -ffp-contract=off (oracle/Makefile)",
                       sample=f"{scene} {width}x{height}, spp={spp}, oracle port, {threads} threads, {sec:.2f} s")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # under torchrun only rank 0 runs the CPU reference
    work = workload(args.config)
    threads = os.cpu_count() or 1
    values, infos = [], []
    for _ in range(args.warmup):
        cpu_reference_step(threads, work, with_as_is=False)
    t0 = time.perf_counter()
    for i in range(args.steps):
        v, info = cpu_reference_step(threads, work, with_as_is=(i == args.steps - 1))
        values.append(v)
        infos.append(info)
    wall = time.perf_counter() - t0
    value = statistics.mean(values)
    info = infos[-1]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / max(1, args.steps) * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (scene arrays as the reference's loader built them, fixture)",
        "config": {"workload": work["name"] + "; each step is a bounded sample of it: " + info["sample"]},
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
        gloo = dist.new_group(backend="gloo")  # host-side barrier of the final gather

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        if world > 1:
            dist.barrier(group=gloo)

    def reduce_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.MAX if world > 1 else None)

    def sum_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.SUM if world > 1 else None)

    work = workload(args.config)
    width, height, spp = work["width"], work["height"], work["spp"]
    scene = scenefile.load(scene_path(work["scene"]))
    marshalled = capi.MarshalledScene(scene)
    ctx = capi.Context(local_rank)
    ctx.upload_scene(marshalled)  # resident: "uploaded once to HBM"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    tag = os.environ.get("MASTER_PORT", str(os.getpid()))

    sampler = ClockSampler(local_rank)
    sampler.start()

    def timed(width, height, spp, steps, warmup, with_e2e=True, sampler=None):
        """Row-partitioned render of one frame at `spp` over all ranks: kernels-only and e2e rates."""
        camera = scene.camera(width, height)
        params = capi.make_params(width, height, spp=spp, seed=SEED)
        options = capi.make_options(rng_mode=capi.RNG_KEYED_PHILOX, device=local_rank,
                                    row_begin=rank, row_step=world)
        samples_total = width * height * spp
        if sampler:
            sampler.mark()  # the device is under this load from here (warm-up) to the end of the e2e loop
        for _ in range(warmup):
            ctx.render(camera, params, options)
        barrier()
        wall0 = time.perf_counter()
        device_ms, sweep_ms, casts, launches = 0.0, 0.0, 0, 0
        for _ in range(steps):
            flush.fill_(1)  # L2 flush between timed iterations (not inside the event-timed region)
            torch.cuda.synchronize()
            st = ctx.render(camera, params, options)  # CUDA events on the launching stream inside
            device_ms += st["kernel_ms"]
            sweep_ms += st["sweep_kernel_ms"]
            casts += st["casts"]
            launches += st["kernel_launches"]
        barrier()
        wall = time.perf_counter() - wall0
        device_s = max_over_ranks(device_ms * 1e-3)
        out = {
            "value": samples_total * steps / device_s / 1e6, "device_s": device_s, "wall_s": wall,
            "sweep_ms_rank": sweep_ms, "casts_rank": casts, "samples_total": samples_total,
            "casts_total": sum_over_ranks(float(casts)), "launches_total": int(sum_over_ranks(float(launches))),
            "params": params, "options": options, "camera": camera,
        }
        if with_e2e:
            # e2e: host buffers through the C ABI; every rank's D2H lands in ONE shared host frame
            shared = partition.SharedFrame(height, width, capi.PIXEL_DTYPE, f"{tag}_{width}x{height}", rank, host_barrier)
            frame = None
            for _ in range(min(warmup, 2)):
                capi.render(marshalled, camera, params, options, out=shared.frame)
            barrier()
            e2e0 = time.perf_counter()
            for _ in range(steps):
                capi.render(marshalled, camera, params, options, out=shared.frame)
                frame = shared.gathered()  # barrier; rank 0 now holds every row
            barrier()
            e2e_s = max_over_ranks(time.perf_counter() - e2e0)
            out["e2e_value"] = samples_total * steps / e2e_s / 1e6
            out["e2e_s"] = e2e_s
            if rank == 0:
                counts = np.asarray(frame["n"])
                out["n_min"], out["n_max"] = int(counts.min()), int(counts.max())
                out["finite"] = bool(np.isfinite(np.asarray(frame["sum"])).all())
            shared.close()
        if sampler:
            sampler.mark()
            sampler.stop()
        return out

    main = timed(width, height, spp, args.steps, args.warmup, sampler=sampler)
    if rank == 0:
        assert main["n_min"] == spp and main["n_max"] == spp, (main["n_min"], main["n_max"], spp)
    h2d = (marshalled.tv.nbytes + marshalled.tm.nbytes + marshalled.sc.nbytes + marshalled.sm.nbytes
           + marshalled.mats.nbytes + 24 + 144 + 36)
    own_rows = len(range(rank, height, world))
    d2h = own_rows * width * 32

    # ---- roofline of the path-tracing kernels (this rank's launches) ----
    hbm_peak, peak_source = measured_peaks()
    sweep_ms_per_step = main["sweep_ms_rank"] / max(1, args.steps)
    casts_per_step_rank = main["casts_rank"] / max(1, args.steps)
    logical_gbs = casts_per_step_rank * scene.sweep_bytes() / (sweep_ms_per_step * 1e-3) / 1e9
    fp64_tflops = casts_per_step_rank * scene.sweep_flops() / (sweep_ms_per_step * 1e-3) / 1e12
    fp64_peak, fp64_probe_ms = capi.measure_fp64_peak(local_rank)
    # Triangle-heavy scenes are bound by the FP32 stage-0 sweep: 21 FMA-pipe operations per
    # (ray, triangle) — 15 for det, X, Y in moment form, 6 for the conservative tests
    fp32_peak, fp32_probe_ms = capi.measure_fp32_peak(local_rank)
    stage0_ops = 21
    fp32_tflops = (casts_per_step_rank * scene.num_triangles * stage0_ops * 2) / (sweep_ms_per_step * 1e-3) / 1e12
    sweep_bound = scene.num_triangles > 64
    traffic, traffic_source, ncu_capture = None, None, None
    ncu_summary = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(ncu_summary):
        try:
            summary = json.load(open(ncu_summary))
            entry = summary.get("dram_bytes_per_step", {}).get(str(args.config))
            if entry:
                traffic, traffic_source = entry["bytes"], entry["source"]
            ncu_capture = summary.get("latest_full_capture")  # committed ncu figures, not measured now
        except Exception:
            traffic = None

    line = None
    if rank == 0:
        clocks = sampler.summary()
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["device_s"] / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (scene arrays as the reference's loader built them, fixture)",
            "config": {
                "workload": f"{work['name']}; {world} GPU(s): the framebuffer is tile-partitioned, rank r renders "
                            f"rows y%{world}==r for all {spp} passes (fixed total work); 4x4 first bounce, "
                            f"maxDepth 5, seed {SEED}",
                "rng": "keyed Philox4x32-10 (parity: bit-exact vs the oracle run with the same policy)",
                "l2": "256 MB flush buffer written between timed steps; the scene is shared-memory resident "
                      "by design, the per-batch record/term buffers (~1 GB) exceed L2",
                "samples_per_step": main["samples_total"],
                "casts_per_sample": main["casts_total"] / (main["samples_total"] * args.steps),
                "wall_s_timed_region": main["wall_s"],
            },
            "clocks": clocks,
            "e2e": {"value": main["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": main["e2e_s"] / args.steps * 1e3,
                    "note": "per rank: scene H2D + kernels + D2H of its own rows into the shared host frame"},
            "gpu_launches": main["launches_total"],
            "roofline": {
                "bound": "fp32" if sweep_bound else "fp64",
                "achieved": fp32_tflops if sweep_bound else fp64_tflops,
                "peak": fp32_peak if sweep_bound else fp64_peak, "unit": "TFLOP/s",
                "frac": (fp32_tflops / fp32_peak if fp32_peak else None) if sweep_bound
                        else (fp64_tflops / fp64_peak if fp64_peak else None),
                "traffic": traffic, "traffic_source": traffic_source,
                "kernel": "primaryHitsKernel + subPathKernel + resolveSamplesKernel (pt_split.cu)",
                "note": "achieved = casts counted on the device x the reference sweep's algorithmic fp64 flops "
                        "(46/triangle + 16/sphere, SURVEY.md 8d) / CUDA-event time of the path-tracing kernels; "
                        "most of that work is executed as a conservative FP32 stage 0, so this is throughput in "
                        "reference flops, not pipe utilisation; peak = DFMA loop measured in this run",
                "peak_source": f"self-measured: fp64PeakKernel, 2*8*16*4096 flop x 512 threads x 4 CTAs/SM in "
                               f"{fp64_probe_ms:.3f} ms (ptb200_measure_fp64_peak)",
                "flops_per_cast": scene.sweep_flops(), "ms_per_step": sweep_ms_per_step,
                "fp64": {"achieved_tflops": fp64_tflops, "peak_tflops": fp64_peak,
                         "frac": fp64_tflops / fp64_peak if fp64_peak else None,
                         "note": "reference flops (46/triangle + 16/sphere per cast) against the DFMA peak; above 1 "
                                 "means the FP32 stage 0 does work the reference does in FP64"},
                "fp32": {"achieved_tflops": fp32_tflops, "peak_tflops": fp32_peak,
                         "frac": fp32_tflops / fp32_peak if fp32_peak else None,
                         "ray_triangle_tests_per_s": casts_per_step_rank * scene.num_triangles / (sweep_ms_per_step * 1e-3),
                         "ops_per_test": stage0_ops,
                         "peak_source": f"self-measured: FFMA2 loop, 4 flop each, {fp32_probe_ms:.3f} ms "
                                        f"(ptb200_measure_fp32_peak)",
                         "note": "the FP32 stage-0 sweep as executed: 21 FMA-pipe operations (2 flop each) per "
                                 "(ray, triangle) pair; binds scenes of more than a few hundred triangles"},
                "hbm": {"logical_sweep_gbs": logical_gbs, "peak_gbs": hbm_peak, "peak_source": peak_source,
                        "logical_over_peak": logical_gbs / hbm_peak, "bytes_per_cast": scene.sweep_bytes(),
                        "note": "LOGICAL bytes (72 B/triangle + 32 B/sphere per cast) served from shared memory "
                                "after one TMA stage per CTA: not a roofline fraction; `traffic` is the DRAM "
                                "traffic ncu measured"},
                "ncu": ncu_capture,
            },
        }

    # ---- side measurements ----
    if world > 1 and args.config == 1:
        weak = timed(width, height, spp * world, max(1, min(args.steps, 2)), 1, with_e2e=True)
        if rank == 0:
            line["weak_scaling"] = {
                "value": weak["value"], "e2e_value": weak["e2e_value"], "unit": UNIT,
                "workload": f"{width}x{height}, {spp * world} passes: the per-GPU share of N=1 kept fixed"}
    if world == 8 and args.config == 1 and not os.environ.get("BENCH_SPP"):
        w4 = WORKLOADS[4]
        c4 = timed(w4["width"], w4["height"], w4["spp"], 1, 0, with_e2e=True)
        if rank == 0:
            line["config4"] = {
                "workload": w4["name"], "value": c4["value"], "e2e_value": c4["e2e_value"], "unit": UNIT,
                "ms_per_step": c4["device_s"] * 1e3, "samples": c4["samples_total"],
                "casts_per_sample": c4["casts_total"] / c4["samples_total"],
                "checks": {"n_min": c4["n_min"], "n_max": c4["n_max"], "expected_n": w4["spp"], "finite": c4["finite"],
                           "casts_per_sample_expected": 33.69}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        camera = scene.camera(width, height)
        if args.config == 1:
            # The other policies, for the record: the reference's exact mt19937 stream (parallel over
            # passes only) on a bounded sample, its `fp` way (exact AND parallel over pixels) on the
            # full workload, its `oo` way on the sequential stream.
            ew, eh = 160, 120
            est = ctx.render(scene.camera(ew, eh), capi.make_params(ew, eh, spp=spp, seed=SEED),
                             capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL, device=local_rank))
            line["exact_stream_mode"] = {
                "value": est["samples"] / est["kernel_ms"] / 1e3, "unit": UNIT,
                "sample": f"cornell {ew}x{eh} (same aspect), {spp} passes, PTB200_RNG_MT19937_SEQUENTIAL",
                "parity": "bit-exact vs the oracle's sequential policy, which equals the reference's "
                          "per-pass images (tests/test_oracle_golden.py)",
                "note": "parallel over passes only (SURVEY.md section 0 item 3); not the timed headline"}
            many = 4096  # a launch that fills the machine: sub-warp pass groups (PtRenderOptions.lanesPerPass = 0)
            est = ctx.render(scene.camera(ew, eh), capi.make_params(ew, eh, spp=many, seed=SEED),
                             capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL, device=local_rank))
            line["exact_stream_mode"]["at_4096_passes"] = {
                "value": est["samples"] / est["kernel_ms"] / 1e3, "unit": UNIT,
                "sample": f"cornell {ew}x{eh}, {many} passes, lanes per pass chosen by the library"}
            fst = ctx.render(camera, capi.make_params(width, height, spp=spp, seed=SEED),
                             capi.make_options(rng_mode=capi.RNG_MT19937_PER_PIXEL, device=local_rank))
            line["fp_way_mode"] = {
                "value": fst["samples"] / fst["kernel_ms"] / 1e3, "unit": UNIT,
                "sample": f"cornell {width}x{height}, {spp} passes, PTB200_RNG_MT19937_PER_PIXEL",
                "casts_per_sample": fst["casts"] / fst["samples"],
                "parity": "bit-exact vs the oracle's fp policy, which equals fp::render of the reference "
                          "(src/fp/Render.cpp) image for image (tests/golden/fp_pass_*.npy)",
                "note": "the reference's --way fp semantics; not the timed headline"}
            ost = ctx.render(scene.camera(ew, eh), capi.make_params(ew, eh, spp=spp, seed=SEED),
                             capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL_OO, device=local_rank))
            line["oo_way_mode"] = {
                "value": ost["samples"] / ost["kernel_ms"] / 1e3, "unit": UNIT,
                "sample": f"cornell {ew}x{eh} (same aspect), {spp} passes, PTB200_RNG_MT19937_SEQUENTIAL_OO",
                "parity": "bit-exact vs the oracle's oo policy, which equals oo::Renderer::radiance of the "
                          "reference (src/oo/Renderer.cpp) image for image (tests/golden/oo_pass_*.npy)",
                "note": "the reference's --way oo semantics, parallel over passes only; not the timed headline"}
        threads = os.cpu_count() or 1
        v, info = cpu_reference_step(threads, work)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, **info}
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
    ap.add_argument("--config", type=int, choices=sorted(WORKLOADS), default=1,
                    help="BASELINE.json configs[N]; the driver's runs use the default, configs[1]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

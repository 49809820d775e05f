#!/usr/bin/env python
"""Rates of the exact-stream policies (one mt19937 per pass) against the number of passes and the
lanes that share a pass.  Run on the GPU box:
    python tools/sequential_rates.py [scene] [width] [height] [passes,passes,...] [lanes,lanes,...]
prints one JSON line per measurement (lanes 0 = the library's own choice)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pt_three_ways_b200 import capi, scenefile  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cornell"
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 160
    h = int(sys.argv[3]) if len(sys.argv) > 3 else 120
    passes = [int(p) for p in (sys.argv[4] if len(sys.argv) > 4 else "256,4096").split(",")]
    lanes = [int(p) for p in (sys.argv[5] if len(sys.argv) > 5 else "0,32,16,8,4").split(",")]
    modes = os.environ.get("SEQUENTIAL_MODES", "dod").split(",")
    scene = scenefile.load(os.path.join(ROOT, "tests/golden/scenes", name + ".ptscene"))
    ctx = capi.Context(0)
    ctx.upload_scene(scene)
    cam = scene.camera(w, h)
    reference = {}
    for mode_name in modes:
        mode = capi.RNG_MT19937_SEQUENTIAL if mode_name == "dod" else capi.RNG_MT19937_SEQUENTIAL_OO
        for spp in passes:
            params = capi.make_params(w, h, spp=spp, seed=1)
            for group in lanes:
                opts = capi.make_options(rng_mode=mode, lanes_per_pass=group)
                best = None
                for _ in range(2):
                    st = ctx.render(cam, params, opts)
                    if best is None or st["sweep_kernel_ms"] < best["sweep_kernel_ms"]:
                        best = st
                image = ctx.download()["sum"]
                key = (mode_name, spp)
                same = True
                if key in reference:
                    same = bool((image == reference[key]).all())
                else:
                    reference[key] = image.copy()
                ms = best["sweep_kernel_ms"]
                print(json.dumps(dict(scene=name, w=w, h=h, way=mode_name, passes=spp, lanes=group, ms=round(ms, 3),
                                      msamples_s=round(best["samples"] / ms / 1e3, 3),
                                      mcasts_s=round(best["casts"] / ms / 1e3, 2),
                                      us_per_cast_per_pass=round(ms * 1e3 * spp / best["casts"], 4),
                                      identical_to_first=same)), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generates tests/golden/* from the REFERENCE ITSELF (oracle/_ref/ref_tool, built from the
unmodified sources under /root/reference by oracle/Makefile).  Run in the build container:

    make -C oracle && python tests/golden/make_golden.py

Outputs (all small, committed):
  scenes/<name>.ptscene     what the reference's OBJ loader + scene recipes fed the SceneBuilder
  cameras.json              Camera state (18 doubles, hex) for several scenes and image sizes
  pass_<case>.npy           per-pass images from the reference's radiance()/randomRay()
  hits_<scene>.npz          rays + the reference's Scene::intersect records
  render_asis_cornell.raw   what the unmodified dod::Scene::render returns (raw format)
  fp_pass_<case>.npy        one whole-screen pass of the reference's `fp` way (fp::render, spp 1)
  fp_render_cornell.raw     the unmodified fp::render, 16x16, 5 spp, --max-cpus 1, seed 3
  oo_pass_<case>.npy        per-pass images from the reference's oo::Renderer::radiance()/randomRay()
  oo_render_asis_cornell.raw  what the unmodified oo::Renderer::render returns (16x16, 6 spp asked)
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_binding as ob  # noqa: E402
from pt_three_ways_b200 import scenefile  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENES = ["cornell", "suzanne", "ce", "single-sphere", "multi-sphere", "example1", "bbc-owl"]

# (case name, scene, width, height, seed, pass, firstU, firstV, maxDepth, preview)
PASS_CASES = [
    ("cornell_32x24_s1_p0", "cornell", 32, 24, 1, 0, 4, 4, 5, 0),
    ("cornell_32x24_s1_p1", "cornell", 32, 24, 1, 1, 4, 4, 5, 0),
    ("cornell_32x24_s1_p2", "cornell", 32, 24, 1, 2, 4, 4, 5, 0),
    ("cornell_20x20_s7_p0_u2v3_d3", "cornell", 20, 20, 7, 0, 2, 3, 3, 0),
    ("cornell_16x12_s3_p0_d1", "cornell", 16, 12, 3, 0, 4, 4, 1, 0),
    ("cornell_16x12_s3_p0_preview", "cornell", 16, 12, 3, 0, 4, 4, 5, 1),
    ("suzanne_24x18_s2_p0", "suzanne", 24, 18, 2, 0, 4, 4, 5, 0),
    ("ce_8x6_s1_p0", "ce", 8, 6, 1, 0, 4, 4, 5, 0),
    ("single-sphere_24x18_s4_p0", "single-sphere", 24, 18, 4, 0, 4, 4, 5, 0),
    ("multi-sphere_24x18_s5_p0", "multi-sphere", 24, 18, 5, 0, 4, 4, 5, 0),
    ("example1_24x18_s6_p0", "example1", 24, 18, 6, 0, 4, 4, 5, 0),
    ("bbc-owl_24x18_s8_p0", "bbc-owl", 24, 18, 8, 0, 4, 4, 5, 0),
]
# The `oo` way (src/oo/Renderer.cpp) walks the same stream as dod: the same cases.
OO_PASS_CASES = PASS_CASES
# (case name, scene, width, height, seed, firstU, firstV, maxDepth, preview): src/fp/Render.cpp
FP_PASS_CASES = [
    ("cornell_32x24_s1", "cornell", 32, 24, 1, 4, 4, 5, 0),
    ("cornell_32x24_s2", "cornell", 32, 24, 2, 4, 4, 5, 0),
    ("cornell_20x20_s7_u2v3_d3", "cornell", 20, 20, 7, 2, 3, 3, 0),
    ("cornell_12x9_s5_u8v8_d6", "cornell", 12, 9, 5, 8, 8, 6, 0),   # > 624 words per engine
    ("cornell_16x12_s3_d1", "cornell", 16, 12, 3, 4, 4, 1, 0),
    ("cornell_16x12_s3_preview", "cornell", 16, 12, 3, 4, 4, 5, 1),
    ("suzanne_24x18_s2", "suzanne", 24, 18, 2, 4, 4, 5, 0),
    ("ce_8x6_s1", "ce", 8, 6, 1, 4, 4, 5, 0),
    ("single-sphere_24x18_s4", "single-sphere", 24, 18, 4, 4, 4, 5, 0),
    ("multi-sphere_24x18_s5", "multi-sphere", 24, 18, 5, 4, 4, 5, 0),
    ("example1_24x18_s6", "example1", 24, 18, 6, 4, 4, 5, 0),
    ("bbc-owl_24x18_s8", "bbc-owl", 24, 18, 8, 4, 4, 5, 0),
]
CAMERA_SIZES = [(64, 48), (640, 480), (1280, 720), (1920, 1080), (256, 256), (16, 16)]


def golden_rays(scene, n, seed):
    rng = np.random.default_rng(seed)
    pts = []
    if scene.num_triangles:
        pts.append(scene.triangle_vertices.reshape(-1, 3))
    c = scene.sphere_centre_radius
    small = c[c[:, 3] < 50]
    if len(small):
        pts.append(small[:, :3] + small[:, 3:4])
        pts.append(small[:, :3] - small[:, 3:4])
    pts = np.concatenate(pts)
    lo, hi = pts.min(0) - 0.5, pts.max(0) + 0.5
    o = rng.uniform(lo, hi, size=(n, 3))
    if scene.num_triangles:
        k = n // 3
        tri = scene.triangle_vertices[rng.integers(0, scene.num_triangles, k)].reshape(k, 3, 3)
        a, b = rng.uniform(size=(2, k))
        flip = a + b > 1
        a[flip], b[flip] = 1 - a[flip], 1 - b[flip]
        o[:k] = tri[:, 0] + a[:, None] * (tri[:, 1] - tri[:, 0]) + b[:, None] * (tri[:, 2] - tri[:, 0])
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], axis=1)


def main():
    assert ob.have_ref_tool() and ob.have_reference_tree(), "build oracle/_ref first"
    os.makedirs(os.path.join(GOLDEN, "scenes"), exist_ok=True)
    print(ob.ref_tool("check-recipes"))
    for name in SCENES:
        print(name, ob.ref_tool("scene", name, os.path.join(GOLDEN, "scenes", name + ".ptscene")).strip())
    cameras = {}
    for name in SCENES:
        for w, h in CAMERA_SIZES:
            cameras[f"{name}:{w}x{h}"] = ob.ref_tool("camera", name, w, h).split()
    json.dump(cameras, open(os.path.join(GOLDEN, "cameras.json"), "w"), indent=0)
    with tempfile.TemporaryDirectory() as tmp:
        for case, scene, w, h, seed, p, fu, fv, depth, preview in PASS_CASES:
            img = ob.ref_pass(scene, w, h, seed, p, tmp, fu, fv, depth, preview)
            np.save(os.path.join(GOLDEN, f"pass_{case}.npy"), img)
            print(case, float(img.min()), float(img.max()))
        for i, name in enumerate(SCENES):
            scene = scenefile.load(os.path.join(GOLDEN, "scenes", name + ".ptscene"))
            rays = golden_rays(scene, 240, 100 + i)
            hits = ob.ref_intersect(name, rays, tmp)
            np.savez_compressed(os.path.join(GOLDEN, f"hits_{name}.npz"), rays=rays, hits=hits)
            print(name, "hits", int(hits[:, 0].sum()), "of", len(rays))
        # The unmodified render entry point, as test/seed_tests.sh drives it (16x16, 16 spp,
        # --max-cpus 1, seed 1) but in the raw format.
        out = os.path.join(GOLDEN, "render_asis_cornell.raw")
        print(ob.ref_render("cornell", 16, 16, 16, 1, 1, out))
        # The `fp` way, through its unmodified entry point fp::render.
        for case, scene, w, h, seed, fu, fv, depth, preview in FP_PASS_CASES:
            img = ob.ref_fp_pass(scene, w, h, seed, tmp, fu, fv, depth, preview)
            np.save(os.path.join(GOLDEN, f"fp_pass_{case}.npy"), img)
            print("fp", case, float(img.min()), float(img.max()))
        print(ob.ref_fp_render("cornell", 16, 16, 5, 1, 3, os.path.join(GOLDEN, "fp_render_cornell.raw")))
        # The `oo` way: its own radiance() per pass, and its unmodified entry point.
        for case, scene, w, h, seed, p, fu, fv, depth, preview in OO_PASS_CASES:
            img = ob.ref_oo_pass(scene, w, h, seed, p, tmp, fu, fv, depth, preview)
            np.save(os.path.join(GOLDEN, f"oo_pass_{case}.npy"), img)
            print("oo", case, float(img.min()), float(img.max()))
        print(ob.ref_oo_render("cornell", 16, 16, 6, 1, 2, os.path.join(GOLDEN, "oo_render_asis_cornell.raw")))


if __name__ == "__main__":
    main()

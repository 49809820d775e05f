"""The drop-in boundary: libptb200.so loads and exports every symbol include/ptb200.h declares;
struct layouts agree between the header, the ctypes binding and the reference's types; and
without a CUDA device the compute entry points fail loudly (no CPU fallback).  No GPU needed."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "ptb200.h")).read()


def declared_functions():
    return sorted(set(re.findall(r"\b(ptb200_[a-z0-9_]+)\s*\(", HEADER)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for needed in ("ptb200_render", "ptb200_render_multi", "ptb200_intersect", "ptb200_last_error",
                   "ptb200_context_create", "ptb200_context_upload_scene", "ptb200_context_render",
                   "ptb200_context_download", "ptb200_context_destroy"):
        assert needed in names


def test_library_exports_every_declared_symbol(capi):
    lib = capi.lib()
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/ptb200.h but not exported"
    assert sorted(capi.EXPORTS) == declared_functions()


def test_struct_layouts(capi):
    assert ctypes.sizeof(capi.PtMaterial) == 72      # MaterialSpec: 9 doubles (MaterialSpec.h:7-12)
    assert ctypes.sizeof(capi.PtCamera) == 144       # Camera: 18 doubles (Camera.h:11-18)
    assert ctypes.sizeof(capi.PtRenderParams) == 36  # RenderParams (RenderParams.h:3-13)
    assert ctypes.sizeof(capi.PtRenderOptions) == 32
    assert capi.PIXEL_DTYPE.itemsize == 32           # SampledPixel: Vec3 + size_t (SampledPixel.h:5-7)
    assert capi.HIT_DTYPE.itemsize == 72
    assert ctypes.sizeof(capi.PtScene) == 4 * 4 + 5 * 8 + 24
    assert ctypes.sizeof(capi.PtStats) == 40


def test_header_has_no_cxx_or_torch_types():
    body = re.sub(r"/\*.*?\*/", "", HEADER.split('extern "C" {', 1)[1], flags=re.S)  # code only
    for forbidden in ("std::", "torch", "at::", "template", "class "):
        assert forbidden not in body


def test_no_cpu_fallback_without_a_device(capi, scenes):
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    scene = scenes["cornell"]
    with pytest.raises(capi.Ptb200Error) as err:
        capi.render(scene, scene.camera(8, 6), capi.make_params(8, 6, spp=1))
    assert err.value.code == 2 and "no CPU fallback" in str(err.value)  # PTB200_ECUDA
    with pytest.raises(capi.Ptb200Error):
        capi.intersect(scene, np.zeros((1, 6)))
    with pytest.raises(capi.Ptb200Error):
        capi.Context(0)
    with pytest.raises(capi.Ptb200Error):
        capi.render(scene, scene.camera(8, 6), capi.make_params(8, 6, spp=1), devices="all")


def test_argument_validation_precedes_device_use(capi, scenes):
    scene = scenes["cornell"]
    cam = scene.camera(8, 6)
    for bad in (capi.make_params(0, 6), capi.make_params(8, 6, max_depth=1000),
                capi.make_params(8, 6, first_u=0)):
        with pytest.raises(capi.Ptb200Error) as err:
            capi.render(scene, cam, bad)
        assert err.value.code == 1  # PTB200_EINVAL
    with pytest.raises(capi.Ptb200Error) as err:
        capi.render(scene, cam, capi.make_params(8, 6),
                    capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL, row_begin=1, row_step=2))
    assert err.value.code == 1 and "partition passes" in str(err.value)
    with pytest.raises(capi.Ptb200Error) as err:
        capi.render(scene, cam, capi.make_params(8, 6),
                    capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL_OO, row_begin=1, row_step=2))
    assert err.value.code == 1 and "partition passes" in str(err.value)
    with pytest.raises(capi.Ptb200Error) as err:
        capi.render(scene, cam, capi.make_params(8, 6), capi.make_options(rng_mode=4))
    assert err.value.code == 1 and "unknown rngMode" in str(err.value)
    for lanes in (1, 5, 64, -8):
        with pytest.raises(capi.Ptb200Error) as err:
            capi.render(scene, cam, capi.make_params(8, 6),
                        capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL, lanes_per_pass=lanes))
        assert err.value.code == 1 and "lanesPerPass" in str(err.value)
    assert capi.PtRenderOptions.lanesPerPass.offset == 24  # the first of the two formerly reserved words


def test_product_does_not_touch_the_oracle():
    """Nothing under pt_three_ways_b200/ may include, import or link anything under oracle/."""
    pkg = os.path.join(ROOT, "pt_three_ways_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cuh", ".cu", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_binding" not in text and "liboracle" not in text, f
                assert not re.search(r'#include\s+"[^"]*oracle', text), f

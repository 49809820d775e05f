"""pt_three_ways_b200 — B200-native backend for the `dod` path tracer of pt-three-ways.

The product is `libptb200.so` (C ABI in include/ptb200.h, CUDA sources in csrc/) plus the C++
host side in host/.  The Python modules here are plumbing for tests and bench.py:

    capi       ctypes binding of the C ABI (raises if the library is missing: there is no
               Python or CPU implementation of the renderer)
    scenefile  PTSCENE2 fixtures (flat SoA scene arrays + recipe camera)
    partition  framebuffer row partition across GPUs and the host-side gather
"""
__all__ = ["capi", "scenefile", "partition"]

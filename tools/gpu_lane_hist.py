"""Scratch: distribution of traversal passes over the number of lanes still alive in their batch (library built with -DRT_LANE_HIST).
usage (GPU box): B200RT_LIB=.../libb200rt_hist.so python tools/gpu_lane_hist.py c5 [c4 ...]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ray_tracing_gallery_b200 import abi, native  # noqa: E402
from ray_tracing_gallery_b200.scene import build_scene  # noqa: E402

SIZES = {"c1": (1280, 720), "c2": (1920, 1080), "c3": (1920, 1080), "c4": (3840, 2160), "c5": (3840, 2160)}
for w in sys.argv[1:] or ["c5"]:
    gpu = native.Renderer(0)
    sg = build_scene(gpu, w, *SIZES[w])
    gpu.render(sg.uniforms(frame_index=1), sg.params(), want=("ray_counts",))
    hist = (C.c_ulonglong * 66)()
    gpu.lib.rt_debug_lane_hist(hist, 1)
    gpu.render(sg.uniforms(frame_index=2), sg.params(), want=("ray_counts",))
    gpu.lib.rt_debug_lane_hist(hist, 1)
    h = np.array(list(hist), dtype=np.float64).reshape(2, 33)
    for kind, name in ((0, "closest-hit rays"), (1, "shadow rays")):
        t = h[kind].sum()
        if t == 0:
            continue
        cum = np.cumsum(h[kind]) / t
        lanes = (h[kind] * np.arange(33)).sum() / t
        print(f"{w} {name}: {t:.3e} warp passes, mean alive lanes {lanes:.1f}; share of passes with <= 2 / 4 / 8 / 16 / 24 lanes alive: "
              f"{cum[2]:.1%} / {cum[4]:.1%} / {cum[8]:.1%} / {cum[16]:.1%} / {cum[24]:.1%};  32 alive: {h[kind][32] / t:.1%}")
    gpu.close()

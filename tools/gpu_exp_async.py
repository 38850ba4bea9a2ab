"""Development aid: throughput of rt_render_async (two overlapping frame slots) under different render flags."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ray_tracing_gallery_b200 import abi, native  # noqa: E402
from ray_tracing_gallery_b200.scene import build_scene  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
gpu = native.Renderer(0)
s = build_scene(gpu, wl)
fbs = [torch.zeros((s.height, s.width, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
rcs = [torch.zeros(2, dtype=torch.int64).pin_memory() for _ in range(2)]


def run(flags, n=200):
    pending, rays = [], 0
    t0 = time.perf_counter()
    for k in range(n):
        b = k & 1
        if len(pending) == 2:
            sl, pb = pending.pop(0)
            gpu.wait_frame(sl)
            rays += int(rcs[pb].sum())
        pending.append((gpu.render_async(s.uniforms(frame_index=1 + k), s.params(flags=flags), fbs[b].data_ptr(), rcs[b].data_ptr()), b))
    for sl, pb in pending:
        gpu.wait_frame(sl)
        rays += int(rcs[pb].sum())
    dt = time.perf_counter() - t0
    return rays / dt / 1e6, dt / n * 1e3


for name, fl in (("default", 0), ("split-tail", abi.RT_RENDER_SPLIT_TAIL), ("no-pdl", abi.RT_RENDER_NO_PDL), ("default", 0),
                 ("split-tail", abi.RT_RENDER_SPLIT_TAIL)):
    run(fl, 20)
    v, ms = run(fl)
    print(f"[{wl}] {name}: {v:.0f} Mrays/s, {ms:.4f} ms/frame", flush=True)
gpu.close()

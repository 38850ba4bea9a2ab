"""compute-sanitizer target: a few small frames through every kernel family of the path (wavefront with cooperative tail and split tail,
megakernel, heat map, refit + rebuild, the box-test audit) and every path of the builder: single-launch SAH (<= 65 536 primitives), the
host-followed level loop with grid-wide splits of large nodes (80 k instances), radix-tree rebuilds above the single-launch limit, exact
instance bounds.  usage: compute-sanitizer --tool memcheck python tools/gpu_sanitize.py
SANITIZE_BUILD_ONLY=1: builds and updates only (racecheck: the shared-memory bins and scans of the SAH kernels)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ray_tracing_gallery_b200 import abi, native  # noqa: E402
from ray_tracing_gallery_b200.scene import build_scene  # noqa: E402

BUILD_ONLY = os.environ.get("SANITIZE_BUILD_ONLY") == "1"
for cfg, kw, size in (("c5", dict(num_instances=20000), (320, 180)), ("c5", dict(num_instances=80000), (160, 90)), ("c4", dict(num_instances=1500), (320, 180)),
                      ("c3", {}, (320, 180)), ("c2", {}, (320, 180)), ("default", {}, (320, 180))):
    gpu = native.Renderer(0)
    sg = build_scene(gpu, cfg, *size, **kw)
    if BUILD_ONLY or kw.get("num_instances", 0) > 65536:
        for mode in (abi.RT_UPDATE_REFIT, abi.RT_UPDATE_REBUILD):
            gpu.update_instances(0, sg.instances); gpu.update_tlas(mode)
        if BUILD_ONLY:
            print(cfg, kw, "built", gpu.stats().tlas_nodes, "TLAS nodes", flush=True)
            gpu.close()
            continue
    for flags in (0, abi.RT_RENDER_SPLIT_TAIL, abi.RT_RENDER_COOP_TAIL | abi.RT_RENDER_COUNTERS):
        out = gpu.render(sg.uniforms(frame_index=2), sg.params(flags=flags))
    gpu.render(sg.uniforms(frame_index=2), sg.params(pipeline=abi.RT_PIPELINE_MEGAKERNEL))
    if sg.dynamic:
        for tick, mode in ((1, abi.RT_UPDATE_REFIT), (2, abi.RT_UPDATE_REBUILD), (3, abi.RT_UPDATE_AUTO)):
            gpu.update_instances(0, sg.animate(tick)); gpu.update_tlas(mode)
            gpu.render(sg.uniforms(frame_index=tick), sg.params())
    st = gpu.stats()
    rays = np.tile(np.array([[0.3, 2.0, -4.0, 0.001, 0.1, -0.3, 0.9, 1e4]], np.float32), (8, 1))
    gpu.box_test(rays, 0, min(st.tlas_nodes, 16), tlas=True)
    print(cfg, "ok", out["ray_counts"].tolist(), flush=True)
    gpu.close()

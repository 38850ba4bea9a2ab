"""Development aid: per-kernel times and per-segment queue fills of one workload.
usage: python tools/gpu_exp.py c4 [--width W --height H --instances N --pipeline mega --frames K]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ray_tracing_gallery_b200 import abi, native  # noqa: E402
from ray_tracing_gallery_b200.scene import build_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--width", type=int, default=0)
ap.add_argument("--height", type=int, default=0)
ap.add_argument("--instances", type=int, default=0)
ap.add_argument("--pipeline", default="wavefront")
ap.add_argument("--frames", type=int, default=5)
ap.add_argument("--shadow-rays", type=int, default=0)
ap.add_argument("--split-tail", action="store_true")
ap.add_argument("--tile", default="", help="x0,y0,w,h")
ap.add_argument("--emulate-rank-of", type=int, default=1, help="render only the strips rank 0 of N would render")
a = ap.parse_args()
gpu = native.Renderer(0)
s = build_scene(gpu, a.workload, a.width or None, a.height or None, num_instances=a.instances or None)
if a.shadow_rays:
    s.shadow_rays = a.shadow_rays
pipe = abi.RT_PIPELINE_MEGAKERNEL if a.pipeline == "mega" else abi.RT_PIPELINE_WAVEFRONT
names = ["trace0", "prep", "shadow", "resolve", "mega", "tail"]
gpu.render(s.uniforms(), s.params(pipeline=pipe, flags=abi.RT_RENDER_COUNTERS), want=("rgba8",))
st = gpu.stats()
rays = st.primary_rays + st.shadow_rays
print(f"[{a.workload}] {s.width}x{s.height} instances {len(s.instances)} tlas_nodes {st.tlas_nodes} blas_nodes {st.blas_nodes} tris {st.num_triangles}")
print(f"  TLAS build {st.last_tlas_ms:.3f} ms")
print(f"  rays: primary {st.primary_rays} shadow {st.shadow_rays} textured hits {st.textured_hits}")
for k, nm in enumerate(("closest-hit", "shadow")):
    n = (st.primary_rays, st.shadow_rays)[k] or 1
    print(f"  {nm}: nodes/ray {st.nodes_visited[k]/n:.2f} instances/ray {st.instances_entered[k]/n:.2f} tris/ray {st.triangles_tested[k]/n:.2f} anyhit/ray {st.anyhit_calls[k]/n:.3f}")
print("  segment bounce rays", list(st.segment_rays)[:4], "segment hits", list(st.segment_hits)[:4])
acc = np.zeros(6)
tot = 0.0
from ray_tracing_gallery_b200.dist import Partition  # noqa: E402
part = Partition.make(s.width, s.height, a.emulate_rank_of, 0)
tile = {}
if a.tile:
    x0, y0, w, h = (int(v) for v in a.tile.split(","))
    tile = dict(tile_x0=x0, tile_y0=y0, tile_w=w, tile_h=h)
for i in range(a.frames):
    gpu.render(s.uniforms(frame_index=2 + i), part.apply(s.params(pipeline=pipe, flags=abi.RT_RENDER_TIMING | (abi.RT_RENDER_SPLIT_TAIL if a.split_tail else 0), **tile)), want=("rgba8",))
    st = gpu.stats()
    acc += np.array(list(st.kernel_ms))
    tot += st.last_render_ms
print("  kernel ms/frame:", {n: round(float(v) / a.frames, 4) for n, v in zip(names, acc)}, "render ms", round(tot / a.frames, 4),
      "Mrays/s", round(rays / (tot / a.frames) / 1e3, 1))
# whole-frame A/B without per-kernel events: programmatic dependent launches on / off
for name, fl in (("pdl", 0), ("no-pdl", abi.RT_RENDER_NO_PDL), ("pdl", 0), ("no-pdl", abi.RT_RENDER_NO_PDL)):
    ms = []
    for i in range(max(a.frames, 20)):
        gpu.render(s.uniforms(frame_index=2 + i), part.apply(s.params(pipeline=pipe, flags=fl, **tile)), want=("rgba8",))
        ms.append(gpu.stats().last_render_ms)
    print(f"  frame ms ({name}): median {np.median(ms):.4f} min {min(ms):.4f}")
if a.workload in ("c4", "c5"):
    import time
    for mode, name in ((abi.RT_UPDATE_REFIT, "refit"), (abi.RT_UPDATE_REBUILD, "rebuild")):
        ms = []
        for i in range(5):
            gpu.update_instances(0, s.instances)
            gpu.update_tlas(mode)
            ms.append(gpu.stats().last_tlas_ms)
        print(f"  TLAS {name}: {min(ms):.3f} ms (best of 5, CUDA events)")
gpu.close()

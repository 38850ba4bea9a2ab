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
print(f"  rays: primary {st.primary_rays} shadow {st.shadow_rays} textured hits {st.textured_hits}")
for k, nm in enumerate(("closest-hit", "shadow")):
    n = (st.primary_rays, st.shadow_rays)[k] or 1
    print(f"  {nm}: nodes/ray {st.nodes_visited[k]/n:.2f} instances/ray {st.instances_entered[k]/n:.2f} tris/ray {st.triangles_tested[k]/n:.2f} anyhit/ray {st.anyhit_calls[k]/n:.3f}")
print("  segment bounce rays", list(st.segment_rays)[:4], "segment hits", list(st.segment_hits)[:4])
acc = np.zeros(6)
tot = 0.0
for i in range(a.frames):
    gpu.render(s.uniforms(frame_index=2 + i), s.params(pipeline=pipe, flags=abi.RT_RENDER_TIMING | (abi.RT_RENDER_SPLIT_TAIL if a.split_tail else 0)), want=("rgba8",))
    st = gpu.stats()
    acc += np.array(list(st.kernel_ms))
    tot += st.last_render_ms
print("  kernel ms/frame:", {n: round(float(v) / a.frames, 4) for n, v in zip(names, acc)}, "render ms", round(tot / a.frames, 4),
      "Mrays/s", round(rays / (tot / a.frames) / 1e3, 1))
gpu.close()

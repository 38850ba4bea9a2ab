"""Traversal counters per ray TYPE (closest-hit = ray-gen segments, first-hit = shadow rays) for the BASELINE workloads.
usage: [B200RT_LIB=...] python tools/gpu_per_ray.py [c5 c4 ...]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ray_tracing_gallery_b200 import abi, native
from ray_tracing_gallery_b200.scene import build_scene

for cfg in (sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]):
    gpu = native.Renderer(0)
    s = build_scene(gpu, cfg)
    gpu.render(s.uniforms(), s.params(flags=abi.RT_RENDER_COUNTERS), want=("ray_counts",))
    st = gpu.stats()
    rays = (max(1, st.primary_rays), max(1, st.shadow_rays))
    for i, name in enumerate(("closest-hit rays", "first-hit (shadow) rays")):
        print(f"{cfg} {name:>24}: {rays[i]:>10} rays | per ray: nodes {st.nodes_visited[i] / rays[i]:6.2f}  instances entered {st.instances_entered[i] / rays[i]:5.2f}  "
              f"triangles {st.triangles_tested[i] / rays[i]:5.2f}  any-hit calls {st.anyhit_calls[i] / rays[i]:5.3f}")
    gpu.close()

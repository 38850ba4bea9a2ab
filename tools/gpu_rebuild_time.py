"""rt_update_tlas REBUILD / rt_build_tlas times for small and medium instance counts (the single-launch SAH build).  usage: [B200RT_LIB=...] python tools/gpu_rebuild_time.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from ray_tracing_gallery_b200 import abi, native
from ray_tracing_gallery_b200.scene import build_scene

tag = os.path.basename(os.environ.get("B200RT_LIB", "libb200rt.so"))
for cfg, n in (("c4", 1000), ("c4", 10000), ("c4", 60000)):
    gpu = native.Renderer(0)
    s = build_scene(gpu, cfg, num_instances=n)
    best = 1e9
    for _ in range(5):
        gpu.update_instances(0, s.instances); gpu.update_tlas(abi.RT_UPDATE_REBUILD)
        best = min(best, gpu.stats().last_tlas_ms)
    print(f"{tag} {n + 1} instances: rt_update_tlas REBUILD {best:.3f} ms (best of 5), {gpu.stats().tlas_nodes} nodes")
    gpu.close()

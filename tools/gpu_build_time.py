"""Build times on the GPU box: every BLAS of the bundled assets (rt_create_model, wall clock) and the static TLAS build of C4 / C5
(rt_build_tlas: `last_tlas_ms` of rt_get_stats = CUDA events around the build), then
rebuild / refit through rt_update_tlas.   usage: [B200RT_LIB=...] python tools/gpu_build_time.py"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np

from ray_tracing_gallery_b200 import abi
from ray_tracing_gallery_b200 import native
from ray_tracing_gallery_b200.scene import build_scene, load_model, push_builtin_images

tag = os.path.basename(os.environ.get("B200RT_LIB", "libb200rt.so"))
gpu = native.Renderer(0)
push_builtin_images(gpu)
for name in ("plane.glb", "tori.glb", "fence.glb", "lain.glb"):
    t = time.perf_counter()
    load_model(gpu, name, 0)
    print(f"{tag} rt_create_model {name}: {(time.perf_counter() - t) * 1e3:.2f} ms wall (upload + BLAS build), {gpu.stats().blas_nodes} BLAS nodes so far")
gpu.close()
for cfg in ("c4", "c5"):
    gpu = native.Renderer(0)
    s = build_scene(gpu, cfg)
    st = gpu.stats()
    print(f"{tag} {cfg}: rt_build_tlas of {len(s.instances)} instances {st.last_tlas_ms:.3f} ms, {st.tlas_nodes} TLAS nodes")
    t = time.perf_counter(); gpu.build_tlas(s.instances); w = (time.perf_counter() - t) * 1e3
    print(f"{tag} {cfg}: second rt_build_tlas {gpu.stats().last_tlas_ms:.3f} ms (wall {w:.2f} ms incl. the upload)")
    for mode, mname in ((abi.RT_UPDATE_REBUILD, "rebuild"), (abi.RT_UPDATE_REFIT, "refit")):
        gpu.update_instances(0, s.instances); gpu.update_tlas(mode)
        gpu.update_instances(0, s.instances); gpu.update_tlas(mode)
        print(f"{tag} {cfg}: rt_update_tlas {mname} {gpu.stats().last_tlas_ms:.3f} ms")
    gpu.close()

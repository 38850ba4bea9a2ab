"""Experiment (not a product path): how much would opening the tori-pair instance into its two rings buy?  C5 / C4 rendered as
shipped (one instance = one BLAS holding both interlocked tori) and with every instance replaced by two instances of single-ring
models with the same transform — the geometry, the rays and the image are the same, only the TLAS leaves are tighter.
usage: python tools/gpu_exp_split_tori.py [c5 c4]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np

from ray_tracing_gallery_b200 import abi, native
from ray_tracing_gallery_b200.gltf import Geometry, ModelArrays, load_gltf
from ray_tracing_gallery_b200.scene import ASSET_DIR as ASSETS, build_scene


def split_components(m):
    idx = np.concatenate([g.indices for g in m.geometries]).reshape(-1, 3)
    P = m.positions
    parent = np.arange(len(P))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    key = {}
    for i, p in enumerate(P):
        k = tuple(np.round(p, 5))
        if k in key:
            parent[find(i)] = find(key[k])
        else:
            key[k] = i
    for a, b, c in idx:
        parent[find(a)] = find(b); parent[find(b)] = find(c)
    comp = np.array([find(t[0]) for t in idx])
    out = []
    g0 = m.geometries[0]
    for v in sorted(set(comp.tolist())):
        out.append(ModelArrays(f"{m.name}_{len(out)}", m.positions, m.normals, m.uvs,
                               [Geometry(idx[comp == v].reshape(-1).astype(np.uint32), g0.opaque, g0.diffuse_image_index, g0.metallic_roughness_image_index, g0.normal_map_image_index)]))
    return out


def timed(gpu, s, frames=5):
    best = 1e9
    for f in range(frames):
        gpu.render(s.uniforms(frame_index=1 + f), s.params(), want=("ray_counts",))
        best = min(best, gpu.stats().last_render_ms)
    gpu.render(s.uniforms(), s.params(flags=abi.RT_RENDER_COUNTERS), want=("ray_counts",))
    st = gpu.stats()
    rays = st.primary_rays + st.shadow_rays
    return best, sum(st.nodes_visited) / rays, sum(st.instances_entered) / rays, sum(st.triangles_tested) / rays, rays


for cfg in (sys.argv[1:] or ["c5", "c4"]):
    gpu = native.Renderer(0)
    s = build_scene(gpu, cfg)
    ms, nodes, inst, tris, rays = timed(gpu, s)
    print(f"{cfg} shipped (pair per instance): {ms:.3f} ms, {rays} rays, per ray nodes {nodes:.2f} instances {inst:.2f} triangles {tris:.2f}")
    # the same scene with two single-ring instances per record (record 0 is the ground plane)
    tori = load_gltf(open(os.path.join(ASSETS, "tori.glb"), "rb").read(), "tori", 1, lambda *a, **k: 1)
    rings = split_components(tori)
    handles = [gpu.create_model(r) for r in rings]
    recs = [s.instances[:1]]
    for mid, h in handles:
        r = s.instances[1:].copy()
        r["custom_index_and_mask"] = (r["custom_index_and_mask"] & np.uint32(0xFF000000)) | np.uint32(mid)
        r["blas"] = np.uint64(h)
        recs.append(r)
    s.instances = np.concatenate(recs)
    gpu.build_tlas(s.instances)
    ms2, nodes, inst, tris, rays2 = timed(gpu, s)
    print(f"{cfg} opened  (one ring per instance, {len(s.instances)} instances): {ms2:.3f} ms ({(ms2 / ms - 1) * 100:+.1f} %), {rays2} rays, per ray nodes {nodes:.2f} instances {inst:.2f} triangles {tris:.2f}")
    gpu.close()

"""Development aid: render the configs with libb200rt and with the oracle, print agreement
statistics, save images under gpurun_out/.  (Uses oracle/ as the checker -> tools/tests only.)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.binding import Oracle  # noqa: E402
from ray_tracing_gallery_b200 import abi, native  # noqa: E402
from ray_tracing_gallery_b200.scene import build_scene  # noqa: E402


def compare(name, a, b, max_segments):
    ids_a, ids_b = a["hit_ids"], b["hit_ids"]
    px_equal = np.all(ids_a == ids_b, axis=(2, 3))
    seg0 = np.all(ids_a[:, :, 0] == ids_b[:, :, 0], axis=2)
    ra, rb = a["radiance"].astype(np.float64), b["radiance"].astype(np.float64)
    fin = np.isfinite(ra) & np.isfinite(rb)
    diff = np.abs(ra - rb)
    rel = diff / np.maximum(np.abs(rb), 1e-3)
    mse = np.mean(np.where(fin, diff, 0) ** 2)
    psnr = 99.0 if mse == 0 else 10 * np.log10(1.0 / mse)
    d8 = np.abs(a["rgba8"].astype(int) - b["rgba8"].astype(int))
    print(f"[{name}] hit-id agreement all segs {px_equal.mean()*100:.4f}%  seg0 {seg0.mean()*100:.4f}%  "
          f"radiance max rel {np.nanmax(np.where(fin, rel, 0)):.3e}  frac>1e-3 {np.mean(rel > 1e-3)*100:.4f}%  PSNR {psnr:.1f} dB  "
          f"rgba8 max diff {d8.max()}  px differing {np.mean(d8.max(axis=2) > 0)*100:.4f}%  "
          f"rays gpu {a['ray_counts']} cpu {b['ray_counts']}", flush=True)
    return px_equal


def main():
    os.makedirs("gpurun_out", exist_ok=True)
    from PIL import Image
    cfgs = sys.argv[1:] or ["c1", "c2", "c3", "default"]
    for cfg in cfgs:
        size = {"c1": (640, 360), "c2": (480, 270), "c3": (640, 360), "default": (640, 360), "c4": (480, 270), "c5": (480, 270)}[cfg]
        kw = {}
        if cfg == "c4" and not os.environ.get("FULL_SCALE"):
            kw["num_instances"] = 2000
        if cfg == "c5" and not os.environ.get("FULL_SCALE"):
            kw["num_instances"] = 20000
        orc = Oracle()
        so = build_scene(orc, cfg, *size, **kw)
        t0 = time.time()
        rb = orc.render(so.uniforms(), so.params())
        t_cpu = time.time() - t0
        for pipeline in (abi.RT_PIPELINE_MEGAKERNEL, abi.RT_PIPELINE_WAVEFRONT):
            gpu = native.Renderer(0)
            sg = build_scene(gpu, cfg, *size, **kw)
            p = sg.params(pipeline=pipeline, flags=abi.RT_RENDER_COUNTERS)
            t0 = time.time()
            ra = gpu.render(sg.uniforms(), p)
            t_gpu = time.time() - t0
            st = gpu.stats()
            nm = f"{cfg}/{'mega' if pipeline else 'wave'}"
            eq = compare(nm, ra, rb, p.max_segments)
            rays = st.primary_rays + st.shadow_rays
            print(f"    render {st.last_render_ms:.3f} ms (cpu oracle {t_cpu*1e3:.0f} ms)  tlas {st.last_tlas_ms:.3f} ms  nodes/ray {sum(st.nodes_visited)/max(rays,1):.1f} "
                  f"inst/ray {sum(st.instances_entered)/max(rays,1):.2f} tris/ray {sum(st.triangles_tested)/max(rays,1):.1f} anyhit {sum(st.anyhit_calls)} "
                  f"tlas_nodes {st.tlas_nodes} blas_nodes {st.blas_nodes} tris {st.num_triangles}", flush=True)
            Image.fromarray(ra["rgba8"]).save(f"gpurun_out/{cfg}_{'mega' if pipeline else 'wave'}.png")
            if not eq.all():
                bad = np.argwhere(~eq)[:5]
                for y, x in bad:
                    print("    mismatch at", (x, y), "gpu", ra["hit_ids"][y, x].tolist(), "cpu", rb["hit_ids"][y, x].tolist())
            gpu.close()
        Image.fromarray(rb["rgba8"]).save(f"gpurun_out/{cfg}_oracle.png")
        orc.close()


if __name__ == "__main__":
    main()

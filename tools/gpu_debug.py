import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ray_tracing_gallery_b200 import abi, native
from ray_tracing_gallery_b200.scene import build_scene

cfg = sys.argv[1] if len(sys.argv) > 1 else "default"
gpu = native.Renderer(0)
s = build_scene(gpu, cfg, 640, 360)
res = {}
for name, pl in (("mega", 1), ("wave1", 0), ("wave2", 0), ("mega2", 1)):
    res[name] = gpu.render(s.uniforms(), s.params(pipeline=pl))
def diff(a, b):
    d = np.abs(res[a]["radiance"] - res[b]["radiance"]).max(axis=2)
    ys, xs = np.nonzero(d > 1e-4)
    print(a, b, "differing px:", len(ys))
    return ys, xs
diff("mega", "mega2")
diff("wave1", "wave2")
ys, xs = diff("mega", "wave1")
for y, x in list(zip(ys, xs))[:12]:
    print((x, y), "tile lane", (x % 8) + 8 * (y % 4), "mega", res["mega"]["radiance"][y, x], "wave", res["wave1"]["radiance"][y, x], res["mega"]["hit_ids"][y, x, 0])

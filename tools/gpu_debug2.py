import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ray_tracing_gallery_b200 import abi, native
from ray_tracing_gallery_b200.scene import build_scene
cfg = sys.argv[1]; segs = int(sys.argv[2]); W, H = int(sys.argv[3]), int(sys.argv[4])
gpu = native.Renderer(0)
s = build_scene(gpu, cfg, W, H)
s.max_segments = segs
a = gpu.render(s.uniforms(), s.params(pipeline=1))
b = gpu.render(s.uniforms(), s.params(pipeline=0))
c = gpu.render(s.uniforms(), s.params(pipeline=0))
d = lambda x, y: int((np.abs(x["radiance"] - y["radiance"]).max(axis=2) > 1e-4).sum())
print(cfg, "segs", segs, "mega-vs-wave", d(a, b), "wave-vs-wave", d(b, c))

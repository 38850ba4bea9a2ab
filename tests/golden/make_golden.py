"""Regenerates tests/golden/*.{json,png} from the CPU oracle.

The reference ships no golden images and cannot run here (Rust + Vulkan ray tracing), so these
fixtures pin the ORACLE against regressions; what pins the oracle to the reference is listed in
oracle/rt_oracle.cpp's header and tests/test_oracle.py.  Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from PIL import Image  # noqa: E402

from oracle.binding import Oracle  # noqa: E402
from ray_tracing_gallery_b200.scene import build_scene  # noqa: E402

CASES = {"c1": {}, "c2": {}, "c3": {"num_instances": 60}, "default": {"num_instances": 60}}

for cfg, kw in CASES.items():
    o = Oracle()
    s = build_scene(o, cfg, 160, 90, **kw)
    r = o.render(s.uniforms(), s.params())
    Image.fromarray(r["rgba8"]).save(os.path.join(HERE, f"{cfg}.png"))
    meta = {
        "config": cfg, "width": 160, "height": 90, "kwargs": kw,
        "shadow_rays": s.shadow_rays, "sun_radius": s.sun_radius, "max_segments": s.max_segments, "frame_index": s.frame_index,
        "hit_ids_sha256": hashlib.sha256(r["hit_ids"].tobytes()).hexdigest(),
        "ray_counts": [int(x) for x in r["ray_counts"]],
    }
    json.dump(meta, open(os.path.join(HERE, f"{cfg}.json"), "w"), indent=1)
    print(cfg, meta["ray_counts"])
    o.close()

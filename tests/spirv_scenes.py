"""Scenes and pixel grids shared by tests/golden/make_spirv_golden.py and tests/test_spirv_pin.py (TEST INFRASTRUCTURE).

Every scene is rendered with the shader's own constants: 2 shadow rays per textured hit (closest_hit_textured.glsl:195)
and a 3-segment ray-gen loop (lib.rs:144), whatever the BASELINE configuration of the same name uses."""
import numpy as np

from ray_tracing_gallery_b200.scene import build_scene

# name -> (build_scene config or "bumpy", width, height, pixel stride, frame_index)
SCENES = {
    "c1": ("c1", 96, 54, 2, 1),          # tori on the plane, hard shadow: table fallbacks (green / pink 1x1 images), ambient term
    "c2": ("c2", 96, 54, 2, 5),          # lain: 2048^2 sRGB diffuse, bilinear; soft shadows, animated blue noise (frame 5)
    "c3": ("c3", 96, 54, 2, 33),         # fence (alpha-clip any-hit) + mirror tori: closest_hit_mirror, second / third segment
    "default": ("default", 96, 54, 2, 2),  # the reference's DefaultScene: portal, fence, lain, textured + mirror tori
    "bumpy": ("bumpy", 80, 45, 2, 7),    # synthetic two-material model: normal map (cotangent frame), magFilter NEAREST
}
NONE = 0xFFFFFFFF


def build(backend, name):
    cfg, width, height, stride, frame = SCENES[name]
    if cfg == "bumpy":
        import synth_assets

        setup = synth_assets.build_bumpy_scene(backend, width, height)
    else:
        setup = build_scene(backend, cfg, width, height)
    setup.shadow_rays = 2
    setup.max_segments = 3
    setup.frame_index = frame
    return setup, width, height, stride


def pixel_grid(width, height, stride):
    return [(x, y) for y in range(stride // 2, height, stride) for x in range(stride // 2, width, stride)]


def run_pixels(pipe, setup, width, height, stride):
    """Run ray_generation.spv (and whatever it traces into) for the grid; arrays for the golden file."""
    pipe.set_uniforms(setup.uniforms(), width, height)
    xy = pixel_grid(width, height, stride)
    n = len(xy)
    image = np.zeros((n, 4), np.float32)
    colour = np.zeros((n, 3), np.float32)
    hits = np.full((n, 3, 3), NONE, np.uint32)
    n_primary, n_shadow = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    for i, (x, y) in enumerate(xy):
        texel, log = pipe.pixel(x, y)
        image[i] = texel
        segs = [e for e in log if not e["shadow"]]
        colour[i] = segs[-1]["payload"][0]
        n_primary[i] = len(segs)
        n_shadow[i] = len(log) - len(segs)
        for s, e in enumerate(segs):
            if e["hit"]:
                hits[i, s] = e["ids"]
    return {"xy": np.array(xy, np.int32), "image": image, "colour": colour, "hit_ids": hits, "n_primary": n_primary, "n_shadow": n_shadow,
            "size": np.array([width, height], np.int32), "frame_index": np.array([setup.frame_index], np.int32)}


def anyhit_candidates(setup, rec, n=400, seed=11):
    """Seeded candidates on the non-opaque geometry of the scene: rows (instance_id, geometry, primitive, u, v)."""
    rng = np.random.default_rng(seed)
    rows = []
    inst_ids = []
    for i, r in enumerate(rec.instances):
        model = rec.models[int(r["custom_index_and_mask"]) & 0xFFFFFF]
        for g, geo in enumerate(model.geometries):
            if not geo.opaque:
                inst_ids.append((i, g, len(geo.indices) // 3))
    assert inst_ids, "scene has no non-opaque geometry"
    for _ in range(n):
        i, g, nprim = inst_ids[rng.integers(len(inst_ids))]
        u = rng.uniform(0, 1)
        v = rng.uniform(0, 1 - u)
        rows.append((i, g, rng.integers(nprim), u, v))
    return np.array(rows, np.float64)

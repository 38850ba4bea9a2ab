"""glTF loading with `Model::load_gltf` semantics (src/util_structs.rs:1049-1156) and scene recipes."""
import os

import numpy as np
import pytest

from ray_tracing_gallery_b200 import abi
from ray_tracing_gallery_b200.gltf import load_gltf
from ray_tracing_gallery_b200.scene import ASSET_DIR, Camera, Sun, build_scene, hash_uniform, make_uniforms


class ImageLog:
    def __init__(self, start=4):
        self.images, self.start = [], start

    def __call__(self, texels, fmt, linear):
        self.images.append((texels.shape, texels.dtype, fmt, bool(linear), texels.copy()))
        return self.start + len(self.images) - 1


def load(name, fallback, log):
    with open(os.path.join(ASSET_DIR, name), "rb") as f:
        return load_gltf(f.read(), name, fallback, log)


def test_asset_facts():
    # SURVEY.md A.6
    log = ImageLog()
    plane = load("plane.glb", 0, log)
    assert plane.positions.shape == (4, 3) and plane.num_triangles == 2
    assert np.allclose(np.abs(plane.positions[:, [0, 2]]), 1) and np.allclose(plane.positions[:, 1], 0)  # node scale ignored
    tori = load("tori.glb", 1, log)
    assert tori.positions.shape == (1274, 3) and tori.num_triangles == 2304
    lain = load("lain.glb", 1, log)
    assert lain.positions.shape == (32349, 3) and lain.num_triangles == 45448
    fence = load("fence.glb", 0, log)
    assert fence.positions.shape == (4, 3) and fence.num_triangles == 2


def test_material_less_model_gets_default_geometry():
    log = ImageLog()
    tori = load("tori.glb", 1, log)
    (g,) = tori.geometries
    assert g.opaque and g.diffuse_image_index == 1 and g.normal_map_image_index == -1
    assert g.metallic_roughness_image_index == 4
    shape, dtype, fmt, linear, texel = log.images[0]
    assert shape == (1, 1, 4) and dtype == np.float32 and fmt == abi.RT_FORMAT_RGBA32_SFLOAT and not linear
    assert texel.reshape(4).tolist() == [1.0, 1.0, 0.0, 1.0]  # (1, roughness 1, metallic 0, 1)


def test_textured_models_push_images_in_material_order():
    log = ImageLog()
    lain = load("lain.glb", 1, log)
    (g,) = lain.geometries
    assert g.opaque and (g.diffuse_image_index, g.metallic_roughness_image_index, g.normal_map_image_index) == (4, 5, -1)
    d, mr = log.images
    assert d[0] == (2048, 2048, 4) and d[1] == np.uint8 and d[2] == abi.RT_FORMAT_RGBA8_SRGB and d[3]  # magFilter LINEAR
    assert mr[2] == abi.RT_FORMAT_RGBA32_SFLOAT and mr[4].reshape(4).tolist() == [1.0, 0.5, 0.0, 1.0]
    log = ImageLog()
    fence = load("fence.glb", 0, log)
    assert not fence.geometries[0].opaque  # alphaMode MASK -> any-hit
    assert log.images[0][0] == (512, 512, 4)
    alpha = log.images[0][4][..., 3]
    assert (alpha < 128).any() and (alpha >= 128).any()


def test_indices_are_rebased_and_in_range():
    log = ImageLog()
    for name in ("plane.glb", "tori.glb", "lain.glb", "fence.glb"):
        m = load(name, 0, log)
        for g in m.geometries:
            assert g.indices.dtype == np.uint32 and len(g.indices) % 3 == 0
            assert g.indices.max() < len(m.positions)
        assert m.normals.shape == m.positions.shape and m.uvs.shape == (len(m.positions), 2)


def test_default_scene_texture_indices():
    """With built-ins 0..3, DefaultScene indices are plane mr 4, tori mr 5, lain 6/7, fence 8/9."""

    class Fake:
        def __init__(self):
            self.n, self.models = 0, 0

        def push_image(self, t, f, l):
            self.n += 1
            return self.n - 1

        def create_model(self, m):
            self.models += 1
            return self.models - 1, 1000 + self.models

        def build_tlas(self, inst):
            self.inst = inst

    f = Fake()
    s = build_scene(f, "default")
    geo = {k: v[2].geometries[0] for k, v in s.models.items()}
    assert geo["plane"].metallic_roughness_image_index == 4 and geo["plane"].diffuse_image_index == 0
    assert geo["tori"].metallic_roughness_image_index == 5 and geo["tori"].diffuse_image_index == 1
    assert (geo["lain"].diffuse_image_index, geo["lain"].metallic_roughness_image_index) == (6, 7)
    assert (geo["fence"].diffuse_image_index, geo["fence"].metallic_roughness_image_index) == (8, 9)
    assert len(s.instances) == 105  # src/scene.rs:95-156
    moved = s.animate(3)  # DefaultScene::update x3: only lain's 48 transform bytes change (src/scene.rs:173-181)
    diff = [i for i in range(105) if moved[i].tobytes() != s.instances[i].tobytes()]
    assert diff == [2] and moved[2]["blas"] == s.instances[2]["blas"]
    kinds = s.instances["sbt_offset_and_flags"] & 0xFFFFFF
    assert kinds[3] == abi.RT_HIT_PORTAL and set(kinds[5:].tolist()) == {abi.RT_HIT_TEXTURED, abi.RT_HIT_MIRROR}


def test_uniform_defaults_and_sun():
    u = make_uniforms(Camera(), Sun(), 1280, 720, 0.05, 1)
    assert u.blue_noise_texture_index == 2 and u.ggx_lut_texture_index == 3 and u.frame_index == 1
    want = [np.cos(0.5) * np.sin(1.0), np.sin(0.5), np.cos(0.5) * np.cos(1.0)]  # src/main.rs:1190-1196
    assert np.allclose(list(u.sun_dir), want, atol=1e-6)
    vi = np.array(list(u.view_inverse)).reshape(4, 4).T
    assert np.allclose(vi[:3, 3], [0, 2, -5], atol=1e-6)


def test_seeded_streams_are_reproducible():
    a, b = hash_uniform(7, 0, 1000), hash_uniform(7, 0, 1000)
    assert np.array_equal(a, b) and 0 <= a.min() and a.max() < 1
    assert abs(a.mean() - 0.5) < 0.05 and not np.array_equal(a, hash_uniform(7, 1, 1000))


@pytest.mark.parametrize("cfg,n", [("c4", 10001), ("c5", 1000001)])
def test_large_configs_have_the_advertised_instance_counts(cfg, n):
    class Fake:
        def __init__(self):
            self.n = self.m = 0

        def push_image(self, *a):
            self.n += 1
            return self.n - 1

        def create_model(self, m):
            self.m += 1
            return self.m - 1, self.m

        def build_tlas(self, inst):
            pass

    s = build_scene(Fake(), cfg)
    assert len(s.instances) == n and s.instances.dtype == abi.INSTANCE_DTYPE
    if cfg == "c4":
        moved = s.animate(3)
        assert not np.array_equal(moved["transform"][1:], s.instances["transform"][1:])
        assert np.array_equal(moved["transform"][0], s.instances["transform"][0])


def test_synthetic_multi_material_model():
    """one geometry per material, images in material order (diffuse, metal-rough, [normal]), sampler choice,
    u16/u32 indices rebased per primitive, primitive -> geometry by material index, node transform ignored."""
    from synth_assets import bumpy_two_material_glb

    log = ImageLog()
    m = load_gltf(bumpy_two_material_glb(), "bumpy", 1, log)
    n = 10
    assert m.positions.shape == (2 * (n + 1) ** 2, 3)          # the two primitives each append the shared vertex block
    assert np.abs(m.positions).max() <= 1.0 + 1e-6              # node scale 3 ignored (src/util_structs.rs:1113)
    g0, g1 = m.geometries
    assert g0.opaque and not g1.opaque
    assert (g0.diffuse_image_index, g0.metallic_roughness_image_index, g0.normal_map_image_index) == (4, 5, 6)
    assert (g1.diffuse_image_index, g1.metallic_roughness_image_index, g1.normal_map_image_index) == (7, 8, -1)
    fmts = [(im[0], im[2], im[3]) for im in log.images]
    assert fmts[0] == ((12, 20, 4), abi.RT_FORMAT_RGBA8_SRGB, False)     # magFilter NEAREST
    assert fmts[1] == ((8, 8, 4), abi.RT_FORMAT_RGBA8_SRGB, True)        # metal-rough is sRGB too (reference quirk)
    assert fmts[2] == ((24, 32, 4), abi.RT_FORMAT_RGBA8_UNORM, True)     # normal map UNORM
    assert fmts[3] == ((16, 16, 4), abi.RT_FORMAT_RGBA8_SRGB, True)      # texture without sampler -> linear
    assert fmts[4][1] == abi.RT_FORMAT_RGBA32_SFLOAT and log.images[4][4].reshape(4).tolist() == [1.0, np.float32(0.6), 0.25, 1.0]
    # primitive 0 (material 1) came first: its indices are not rebased, primitive 1 (material 0) is rebased by (n+1)^2
    assert len(g0.indices) + len(g1.indices) == n * n * 6
    assert g1.indices.max() < (n + 1) ** 2 <= g0.indices.min()

"""Pins for the CPU oracle: the reference's only unit test, constants of the shipped .spv files,
independent float64 re-derivations of every shader formula, BVH-vs-brute-force equality, and the
committed golden fixtures (tests/golden/, made by tests/golden/make_golden.py)."""
import hashlib
import json
import math
import os
import struct

import numpy as np
import pytest

from conftest import make_oracle
from ray_tracing_gallery_b200 import abi
from ray_tracing_gallery_b200.gltf import load_png_file_rgba8
from ray_tracing_gallery_b200.scene import ASSET_DIR, Camera, Sun, build_scene, make_uniforms

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PHI = np.float32(0.618033988749)


@pytest.fixture(scope="module")
def orc():
    o = make_oracle()
    from ray_tracing_gallery_b200.scene import push_builtin_images

    push_builtin_images(o)
    yield o
    o.close()


# ---------------------------------------------------------------- reference KAT + .spv constants
def test_ggx_division_by_zero(orc):
    """shaders/ray-tracing/src/pbr.rs:52-69 — V_SmithGGXCorrelated is finite at roughness 0."""
    v = orc.v_smith_ggx((0, 1, 0), (1, 0, 0), (0, 1, 0), 0.0)
    assert math.isfinite(v)
    # NoV clamps to 10.0e-10 (pbr.glsl:33), NoL = 1: V = 0.5 / (1*sqrt(NoV^2) + NoV*1) = 0.25 / NoV
    assert v == pytest.approx(0.25 / 1e-9, rel=1e-5)


def _spv_constants(path):
    words = np.frombuffer(open(path, "rb").read(), dtype="<u4")
    assert words[0] == 0x07230203
    i, floats, ints = 5, [], []
    while i < len(words):
        wc, op = int(words[i]) >> 16, int(words[i]) & 0xFFFF
        if op == 43 and wc == 4:  # OpConstant (32-bit)
            ints.append(int(words[i + 3]))
            floats.append(struct.unpack("<f", struct.pack("<I", int(words[i + 3])))[0])
        i += max(wc, 1)
    return floats, ints


@pytest.mark.skipif(not os.path.exists("/root/reference/shaders/closest_hit_textured.spv"), reason="reference tree not present")
def test_shipped_spirv_constants_match_the_restatement():
    floats, ints = _spv_constants("/root/reference/shaders/closest_hit_textured.spv")
    for want in (0.618033988749, 1e-9, 0.1, 0.04, 0.001, 10000.0, 64.0, 0.5):
        assert any(abs(f - want) <= 1e-6 * max(1.0, abs(want)) for f in floats), want
    for want in (13, 41, 32, 2):  # blue-noise offsets, frame modulus, shadow-ray loop bound
        assert want in ints
    floats, ints = _spv_constants("/root/reference/shaders/ray_generation.spv")
    for want in (0.01, 10000.0, 0.0031308, 12.92, 1.055, 0.055):
        assert any(abs(f - want) <= 1e-6 * max(1.0, abs(want)) for f in floats), want


def _spv_opcounts(path):
    words = np.frombuffer(open(path, "rb").read(), dtype="<u4")
    i, ops = 5, {}
    while i < len(words):
        wc, op = int(words[i]) >> 16, int(words[i]) & 0xFFFF
        ops[op] = ops.get(op, 0) + 1
        i += max(wc, 1)
    return ops


@pytest.mark.skipif(not os.path.exists("/root/reference/shaders/any_hit_alpha_clip.spv"), reason="reference tree not present")
def test_shipped_spirv_stage_structure_matches_the_restatement():
    """The other shipped stages: what they sample, trace, ignore and which constants they carry — the facts the oracle's
    stages (and the CUDA ones) are written to."""
    OP_SAMPLE_IMPLICIT, OP_SAMPLE_EXPLICIT, OP_TRACE_RAY, OP_IGNORE, OP_READ_CLOCK = 87, 88, 4445, 4448, 5056
    sh = "/root/reference/shaders/"
    # any-hit: ONE explicit-LOD texture tap, ignoreIntersection, threshold 0.5 (any_hit_alpha_clip.glsl:24-27)
    ops = _spv_opcounts(sh + "any_hit_alpha_clip.spv")
    assert ops.get(OP_SAMPLE_EXPLICIT) == 1 and ops.get(OP_IGNORE) == 1 and OP_SAMPLE_IMPLICIT not in ops
    assert 0.5 in _spv_constants(sh + "any_hit_alpha_clip.spv")[0]
    # textured closest-hit: blue noise x2 + diffuse + metal-rough + normal map = 5 taps, all LOD 0; one trace call site (the shadow ray)
    ops = _spv_opcounts(sh + "closest_hit_textured.spv")
    assert ops.get(OP_SAMPLE_EXPLICIT) == 5 and OP_SAMPLE_IMPLICIT not in ops and ops.get(OP_TRACE_RAY) == 1
    floats, ints = _spv_constants(sh + "closest_hit_textured.spv")
    assert 12 in ints  # gl_RayFlagsTerminateOnFirstHitEXT (4) | gl_RayFlagsSkipClosestHitShaderEXT (8)
    assert np.float32(1.0 / 64.0) in np.asarray(floats, np.float32)  # blue-noise texel size
    # mirror: no texture, no trace; portal: +5 on y; miss: the sky's 0.05 blue and the sun's 1.0
    ops = _spv_opcounts(sh + "closest_hit_mirror.spv")
    assert OP_SAMPLE_EXPLICIT not in ops and OP_TRACE_RAY not in ops
    def real(path):  # 32-bit constants that are not small integers read as denormals
        return sorted({float(np.float32(f)) for f in _spv_constants(path)[0] if abs(f) > 1e-30})

    assert real(sh + "closest_hit_portal.spv") == [5.0]
    assert real(sh + "primary_ray_miss.spv") == [float(np.float32(0.05)), 1.0]
    assert real(sh + "shadow_ray_miss.spv") == []
    # ray generation: one trace call site inside the segment loop, the clock read twice (show_heatmap)
    ops = _spv_opcounts(sh + "ray_generation.spv")
    assert ops.get(OP_TRACE_RAY) == 1 and ops.get(OP_READ_CLOCK) == 2


# ---------------------------------------------------------------- shading formulas in float64
def brdf64(n, v, l, base, pr, m, sun):
    n, v, l, base = (np.asarray(x, np.float64) for x in (n, v, l, base))
    h = (v + l) / np.linalg.norm(v + l)
    a = pr * pr
    NoV = min(max(n @ v, 1e-9), 1.0)
    NoH, NoL, LoH = (min(max(x, 0.0), 1.0) for x in (n @ h, n @ l, l @ h))
    D = (a / (1 - NoH * NoH + (NoH * a) ** 2)) ** 2 / math.pi
    V = 0.5 / (NoL * math.sqrt(NoV * NoV * (1 - a * a) + a * a) + NoV * math.sqrt(NoL * NoL * (1 - a * a) + a * a))
    f90 = 0.5 + 2 * a * LoH * LoH
    f0 = 0.04 * (1 - m) + base * m
    F = f0 + (f90 - f0) * (1 - LoH) ** 5
    Fd = (1 + (f90 - 1) * (1 - NoL) ** 5) * (1 + (f90 - 1) * (1 - NoV) ** 5) / math.pi
    return sun * NoL * (base * Fd + D * V * F)


def test_brdf_against_float64(orc):
    rng = np.random.default_rng(7)
    for _ in range(200):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        v = rng.normal(size=3); v /= np.linalg.norm(v)
        l = rng.normal(size=3); l /= np.linalg.norm(l)
        if n @ v < 0.05 or n @ l < 0.05:
            continue
        base, pr, m, sun = rng.uniform(0, 1, 3), rng.uniform(0.05, 1), rng.uniform(0, 1), rng.choice([0.0, 0.5, 1.0])
        got = orc.brdf(n, v, l, base, float(pr), float(m), float(sun))
        want = brdf64(n.astype(np.float32), v.astype(np.float32), l.astype(np.float32), base.astype(np.float32), np.float32(pr), np.float32(m), sun)
        assert np.allclose(got, want, rtol=2e-4, atol=1e-6)


def test_linear_to_srgb_and_unorm8(orc):
    lib = orc.lib
    assert lib.orc_linear_to_srgb(0.0) == 0.0
    assert lib.orc_linear_to_srgb(0.0031308) == pytest.approx(0.0031308 * 12.92, rel=1e-6)
    assert lib.orc_linear_to_srgb(1.0) == pytest.approx(1.0, abs=1e-6)
    for c in (0.004, 0.05, 0.2, 0.5, 0.9):
        assert lib.orc_linear_to_srgb(c) == pytest.approx(1.055 * c ** (1 / 2.4) - 0.055, rel=1e-5)
    # SKY_COLOUR (0,0,0.05) -> navy (0,0,63), lib.rs:38
    assert lib.orc_unorm8(lib.orc_linear_to_srgb(0.05)) == 63
    assert lib.orc_unorm8(-1.0) == 0 and lib.orc_unorm8(2.0) == 255 and lib.orc_unorm8(float("nan")) == 0
    assert lib.orc_unorm8(0.5) == 128


def test_blue_noise_sequence(orc):
    """closest_hit_textured.glsl:99-120: texel (px + k*(13,41)) mod 64, fract(bn + (frame % 32) * phi)."""
    tex = load_png_file_rgba8(os.path.join(ASSET_DIR, "blue_noise_64x64.png"))[..., 0].astype(np.float32) / np.float32(255)
    for px, py, it, frame in [(0, 0, 0, 1), (5, 7, 1, 1), (1279, 719, 3, 31), (100, 200, 15, 32), (63, 64, 2, 77)]:
        k = np.float32(frame % 32) * PHI
        a = tex[(py + 2 * it * 41) % 64, (px + 2 * it * 13) % 64] + k
        b = tex[(py + (2 * it + 1) * 41) % 64, (px + (2 * it + 1) * 13) % 64] + k
        want = np.array([a - np.floor(a), b - np.floor(b)], np.float32)
        assert np.array_equal(orc.blue_noise_xi(px, py, it, frame), want)


def test_blue_noise_png_is_16_bit_grey_narrowed_by_shift():
    from PIL import Image

    img = Image.open(os.path.join(ASSET_DIR, "blue_noise_64x64.png"))
    assert img.size == (64, 64) and img.mode.startswith("I")
    raw = np.asarray(img).astype(np.uint32)
    rgba = load_png_file_rgba8(os.path.join(ASSET_DIR, "blue_noise_64x64.png"))
    assert np.array_equal(rgba[..., 0], (raw >> 8).astype(np.uint8)) and np.all(rgba[..., 3] == 255)


def test_sample_directional_light(orc):
    c = np.array([0.7384603, 0.47942555, 0.47415987], np.float32)
    # radius 0 (hard shadows, C1): the jitter vanishes exactly, whatever the blue-noise sample
    d0 = orc.sample_directional_light((0.3, 0.8), c, 0.0)
    assert np.array_equal(d0, orc.sample_directional_light((0.9, 0.1), c, 0.0))
    assert np.allclose(d0, c / np.linalg.norm(c), atol=1e-7)
    rng = np.random.default_rng(3)
    for _ in range(50):
        xi = rng.uniform(0, 1, 2)
        d = orc.sample_directional_light(xi, c, 0.05).astype(np.float64)
        assert abs(np.linalg.norm(d) - 1) < 1e-6
        cn = c.astype(np.float64) / np.linalg.norm(c)
        t = np.cross(cn, [0, 1, 0]); t /= np.linalg.norm(t)
        b = np.cross(t, cn); b /= np.linalg.norm(b)
        r, ang = math.sqrt(np.float32(xi[0])), float(np.float32(xi[1])) * 2 * math.pi
        want = c.astype(np.float64) + 0.05 * r * (math.cos(ang) * t + math.sin(ang) * b)
        want /= np.linalg.norm(want)
        assert np.allclose(d, want, atol=2e-6)


def test_texture_sampling_rules():
    o = make_oracle()
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, size=(4, 8, 4), dtype=np.uint8)
    lin_unorm = o.push_image(img, abi.RT_FORMAT_RGBA8_UNORM, True)
    near_srgb = o.push_image(img, abi.RT_FORMAT_RGBA8_SRGB, False)
    const = o.push_image(np.array([[[1.0, 0.5, 0.25, 1.0]]], np.float32), abi.RT_FORMAT_RGBA32_SFLOAT, False)
    f = img.astype(np.float64) / 255

    def srgb(c):
        return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)

    for u, v in [(0.1, 0.2), (0.99, 0.01), (-0.3, 1.7), (3.25, -2.5), (0.5, 0.5), (0.0625, 0.125)]:
        x, y = u * 8 - 0.5, v * 4 - 0.5
        x0, y0 = math.floor(x), math.floor(y)
        ax, ay = x - x0, y - y0
        t = lambda i, j: f[(y0 + j) % 4, (x0 + i) % 8]
        want = t(0, 0) * (1 - ax) * (1 - ay) + t(1, 0) * ax * (1 - ay) + t(0, 1) * (1 - ax) * ay + t(1, 1) * ax * ay
        assert np.allclose(o.sample_texture(lin_unorm, u, v), want, atol=3e-6)
        tx = f[math.floor(v * 4) % 4, math.floor(u * 8) % 8]
        want_n = np.concatenate([srgb(tx[:3]), tx[3:]])
        assert np.allclose(o.sample_texture(near_srgb, u, v), want_n, atol=3e-6)
        assert np.array_equal(o.sample_texture(const, u, v), [1.0, 0.5, 0.25, 1.0])
    assert np.array_equal(o.sample_texture(99, 0.5, 0.5), [0, 0, 0, 0])  # null descriptor
    o.close()


def test_intersection_and_inverse_against_float64(orc):
    rng = np.random.default_rng(5)
    hits = 0
    for _ in range(400):
        a, b, c = rng.uniform(-1, 1, (3, 3)).astype(np.float32)
        o = rng.uniform(-3, 3, 3).astype(np.float32)
        tgt = (a + b + c) / 3 + rng.normal(scale=0.3, size=3)
        d = (tgt - o).astype(np.float32)
        hit, tuv = orc.intersect_triangle(o, d, a, b, c)
        A, B, Cc, O, D = (x.astype(np.float64) for x in (a, b, c, o, d))
        M = np.stack([-D, B - A, Cc - A], axis=1)
        if abs(np.linalg.det(M)) < 1e-9:
            continue
        t, u, v = np.linalg.solve(M, O - A)
        inside = u >= 0 and v >= 0 and u + v <= 1
        if min(abs(u), abs(v), abs(1 - u - v)) < 1e-4:
            continue
        assert hit == inside
        if hit:
            hits += 1
            assert np.allclose(tuv, [t, u, v], rtol=1e-4, atol=1e-5)
    assert hits > 50
    for _ in range(50):
        m = np.concatenate([rng.normal(size=(3, 3)) + 2 * np.eye(3), rng.normal(size=(3, 1))], axis=1).astype(np.float32)
        inv = orc.invert_3x4(m).reshape(3, 4)
        full = np.vstack([m.astype(np.float64), [0, 0, 0, 1]])
        assert np.allclose(inv, np.linalg.inv(full)[:3], rtol=1e-4, atol=1e-5)


def test_primary_ray_matches_camera_model(orc):
    cam, sun = Camera(), Sun()
    u = make_uniforms(cam, sun, 1280, 720, 0.05, 1)
    o, d = orc.primary_ray(u, 640, 360, 1280, 720)
    assert np.allclose(o, [0, 2, -5], atol=1e-6)
    # yaw = pi looks down +z; image row 0 is up (Vulkan clip space, reversed-z projection)
    assert np.allclose(d, [0, 0, 1], atol=2e-3)
    _, d_top = orc.primary_ray(u, 640, 0, 1280, 720)
    _, d_right = orc.primary_ray(u, 1279, 360, 1280, 720)
    assert d_top[1] > 0.4 and abs(np.linalg.norm(d_top) - 1) < 1e-6
    half = math.tan(math.radians(59) / 2)
    assert d_top[1] / d_top[2] == pytest.approx(half * (1 - 1 / 720), rel=1e-4)
    assert abs(d_right[0] / d_right[2]) == pytest.approx(half * 1280 / 720 * (1 - 1 / 1280), rel=1e-4)


def test_shadow_terminator_fix(orc):
    """RTG-II 4.3 (closest_hit_textured.glsl:13-39): flat normals -> no offset; curved normals -> the
    origin moves to the convex side of the triangle."""
    pos = [0, 0, 0, 1, 0, 0, 0, 1, 0]
    ident = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]
    bary = (0.2, 0.3, 0.5)
    flat = orc.terminator_origin(pos, [0, 0, 1] * 3, bary, ident)
    assert np.allclose(flat, [0.3, 0.5, 0.0], atol=1e-7)
    curved = orc.terminator_origin(pos, [-0.5, -0.5, 1, 0.8, 0, 1, 0, 0.8, 1], bary, ident)
    assert curved[2] > 1e-3
    moved = orc.terminator_origin(pos, [0, 0, 1] * 3, bary, [2, 0, 0, 5, 0, 2, 0, 6, 0, 0, 2, 7])
    assert np.allclose(moved, [5.6, 7.0, 7.0], atol=1e-6)


# ---------------------------------------------------------------- trace semantics
def test_bvh_equals_brute_force():
    for cfg, kw in (("c3", {"num_instances": 30}), ("default", {"num_instances": 12})):
        a, b = make_oracle(), make_oracle(brute_force=True)
        sa, sb = build_scene(a, cfg, 96, 54, **kw), build_scene(b, cfg, 96, 54, **kw)
        ra, rb = a.render(sa.uniforms(), sa.params()), b.render(sb.uniforms(), sb.params())
        assert np.array_equal(ra["hit_ids"], rb["hit_ids"])
        assert np.array_equal(ra["radiance"].view(np.uint32), rb["radiance"].view(np.uint32))
        assert np.array_equal(ra["ray_counts"], rb["ray_counts"])
        a.close(); b.close()


def test_interval_is_exclusive_and_ties_go_to_lowest_ids():
    from ray_tracing_gallery_b200.scene import load_model, make_instance, mat_identity, mat_translation, push_builtin_images

    o = make_oracle()
    push_builtin_images(o)
    pid, ph, _ = load_model(o, "plane.glb", 0)
    # two coincident planes: instance 0 must win the exact tie; plane is y = 0, x,z in [-1,1]
    o.build_tlas(np.stack([make_instance(mat_identity(), pid, ph, 0), make_instance(mat_identity(), pid, ph, 0)]))
    hit, ids, tuv = o.trace((0.25, 1, 0.25), (0, -1, 0), 0.01, 100.0)
    assert hit and ids[0] == 0 and tuv[0] == 1.0
    hit, ids, _ = o.trace((0.25, 1, 0.25), (0, -1, 0), 1.0, 100.0)  # t == tmin is not a hit
    assert not hit
    hit, ids, _ = o.trace((0.25, 1, 0.25), (0, -1, 0), 0.01, 1.0)  # t == tmax is not a hit
    assert not hit
    # no face culling: the plane is hit from below too
    hit, ids, _ = o.trace((0.25, -1, 0.25), (0, 1, 0), 0.01, 100.0)
    assert hit
    # direction is not normalised and t is preserved through the instance transform
    o.build_tlas(np.stack([make_instance(mat_translation(0, -3, 0), pid, ph, 0)]))
    hit, ids, tuv = o.trace((0, 1, 0), (0, -2, 0), 0.01, 100.0)
    assert hit and tuv[0] == 2.0
    o.close()


def test_ray_counts_and_segments():
    o = make_oracle()
    s = build_scene(o, "c1", 160, 90)
    r = o.render(s.uniforms(), s.params())
    hit_px = int(np.sum(r["hit_ids"][:, :, 0, 0] != abi.MISS_ID))
    assert r["ray_counts"][0] == 160 * 90  # no mirrors in C1: one segment per pixel
    assert r["ray_counts"][1] == hit_px * s.shadow_rays
    sky = r["rgba8"][0, 0]
    assert tuple(sky) == (0, 0, 63, 255)
    o.close()


def test_alpha_clip_anyhit_lets_rays_through_the_fence():
    o = make_oracle()
    s = build_scene(o, "c3", 64, 36, num_instances=0)
    fence_idx = 1  # second instance in c3
    # straight at the fence quad centred at (2, 1, 2): some rays pass (alpha < 0.5), some stop
    seen = set()
    for i in range(40):
        x = 1.1 + 1.8 * i / 39
        hit, ids, _ = o.trace((x, 1.0, -3.0), (0, 0, 1), 0.01, 100.0)
        seen.add(ids[0] if hit else None)
    assert fence_idx in seen and (seen - {fence_idx}), seen
    o.close()


def test_frame_against_an_independent_float64_brute_force():
    """A second, independent statement of the whole segment-0 path in numpy float64 — ray generation (lib.rs:126-142), instance
    transform, brute force over every triangle in WORLD space, exclusive interval, closest t — against the oracle's frame: hit IDs
    must agree except on silhouette pixels (fp32 against fp64), the hard-shadow flag of C1 decides lit/unlit per pixel, and the
    trace-call count follows from the hit mask."""
    o = make_oracle()
    W, H = 128, 72
    s = build_scene(o, "c1", W, H)
    r = o.render(s.uniforms(), s.params())
    u = s.uniforms()
    Vi = np.array(list(u.view_inverse), np.float64).reshape(4, 4).T   # column-major -> matrix
    Pi = np.array(list(u.proj_inverse), np.float64).reshape(4, 4).T
    sun = np.array(list(u.sun_dir), np.float64)
    xs, ys = np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5)
    ndc = np.stack([xs / W * 2 - 1, ys / H * 2 - 1, np.ones_like(xs), np.ones_like(xs)], axis=-1)
    target = ndc @ Pi.T
    ld = target[..., :3] / np.linalg.norm(target[..., :3], axis=-1, keepdims=True)
    d = ld @ Vi[:3, :3].T
    org = np.broadcast_to(Vi[:3, 3], d.shape)

    # world-space triangles of every instance, with their (instance, geometry, primitive) ids
    tris, ids = [], []
    names = {mid: arrays for (mid, _, arrays) in s.models.values()}
    for ii, rec in enumerate(s.instances):
        m = names[int(rec["custom_index_and_mask"]) & 0xFFFFFF]
        T = rec["transform"].astype(np.float64).reshape(3, 4)
        P = m.positions.astype(np.float64) @ T[:, :3].T + T[:, 3]
        for gi, g in enumerate(m.geometries):
            idx = np.asarray(g.indices, np.int64).reshape(-1, 3)
            tris.append(P[idx])
            ids.append(np.stack([np.full(len(idx), ii), np.full(len(idx), gi), np.arange(len(idx))], axis=1))
    tris, ids = np.concatenate(tris), np.concatenate(ids)

    def closest(o3, d3, tmin, tmax):
        """(N,3) rays against all triangles: index of the closest valid candidate (-1 = miss) and its t."""
        v0, e1, e2 = tris[:, 0], tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]
        best_t, best_i = np.full(len(o3), np.inf), np.full(len(o3), -1)
        for a in range(0, len(tris), 256):
            sl = slice(a, a + 256)
            p = np.cross(d3[:, None, :], e2[None, sl])
            det = np.einsum("ntk,tk->nt", p, e1[sl])
            with np.errstate(divide="ignore", invalid="ignore"):
                inv = 1.0 / det
                tv = o3[:, None, :] - v0[None, sl]
                uu = np.einsum("ntk,ntk->nt", tv, p) * inv
                q = np.cross(tv, e1[None, sl])
                vv = np.einsum("nk,ntk->nt", d3, q) * inv
                tt = np.einsum("ntk,tk->nt", q, e2[sl]) * inv
            ok = (det != 0) & (uu >= 0) & (vv >= 0) & (uu + vv <= 1) & (tt > tmin) & (tt < tmax)
            tt = np.where(ok, tt, np.inf)
            k = tt.argmin(axis=1)
            t_here = tt[np.arange(len(o3)), k]
            better = t_here < best_t
            best_t = np.where(better, t_here, best_t)
            best_i = np.where(better, a + k, best_i)
        return best_i, best_t

    bi, bt = closest(org.reshape(-1, 3), d.reshape(-1, 3), 0.01, 10000.0)
    want_ids = np.where(bi[:, None] >= 0, ids[np.maximum(bi, 0)], abi.MISS_ID).reshape(H, W, 3)
    got_ids = r["hit_ids"][:, :, 0, :].astype(np.int64)
    agree = np.all(got_ids == want_ids, axis=2)
    assert agree.mean() >= 0.995, agree.mean()
    assert (bi >= 0).mean() > 0.3 and (bi < 0).mean() > 0.1           # the frame shows ground, tori and sky
    hit = (bi >= 0).reshape(H, W)
    assert abs(int(r["ray_counts"][0]) - W * H) == 0 and abs(int(r["ray_counts"][1]) - int(hit.sum())) <= (~agree).sum()
    # sky pixels: SKY_COLOUR or the sun disc (lib.rs:38-51); sun_radius 0 -> cos = 1 -> never the disc
    sky = ~hit & agree
    assert np.all(r["rgba8"][sky] == np.array([0, 0, 63, 255], np.uint8))
    # hard shadow (sun_radius 0): a pixel whose shadow ray (from the hit point, nudged along the normal side) is blocked gets only
    # the 0.1 * base ambient term; check the darkest / brightest split on the ground plane against the float64 shadow test
    ground = (want_ids[..., 0] == 0) & agree
    hp = (org.reshape(-1, 3) + d.reshape(-1, 3) * bt[:, None]).reshape(H, W, 3)
    gidx = np.argwhere(ground)[::7]
    so = hp[gidx[:, 0], gidx[:, 1]] + np.array([0.0, 1e-4, 0.0])
    si, _ = closest(so, np.broadcast_to(sun, so.shape).copy(), 0.001, 10000.0)
    lum = r["radiance"][gidx[:, 0], gidx[:, 1]].sum(axis=1)
    shadowed, lit = lum[si >= 0], lum[si < 0]
    assert len(shadowed) > 20 and len(lit) > 20
    # allow the few samples on the shadow edge to fall on either side
    thresh = 0.5 * (np.median(shadowed) + np.median(lit))
    assert np.median(lit) > 2 * np.median(shadowed)
    assert np.isclose(np.median(shadowed), lum.min(), rtol=1e-6)  # blocked -> exactly the 0.1 * base ambient term (glsl:222-225)
    assert (shadowed < thresh).mean() > 0.97 and (lit > thresh).mean() > 0.97
    # and the whole pixel, closest_hit_textured.glsl:174-226 for the ground: base = green.png (103, 234, 64) sRGB-decoded, material
    # fallback roughness 1 / metallic 0 (util_structs.rs:1090-1111), n = +y, v = -d, l = sun_dir, sun_factor from the float64
    # shadow test; colour = brdf + 0.1 * base
    srgb = lambda c: c / 12.92 if c <= 0.04045 else ((c + 0.055) / 1.055) ** 2.4
    base = np.array([srgb(103 / 255), srgb(234 / 255), srgb(64 / 255)])
    dd = d[gidx[:, 0], gidx[:, 1]]
    got_rad = r["radiance"][gidx[:, 0], gidx[:, 1]].astype(np.float64)
    close = 0
    for k in range(len(gidx)):
        want = brdf64((0.0, 1.0, 0.0), -dd[k], sun, base, 1.0, 0.0, 0.0 if si[k] >= 0 else 1.0) + 0.1 * base
        close += bool(np.all(np.abs(got_rad[k] - want) <= 1e-3 * np.maximum(np.abs(want), 1e-3)))
    assert close >= 0.97 * len(gidx), (close, len(gidx))  # all but the samples on a shadow edge
    o.close()


def test_soft_shadow_samples_against_an_independent_float64_brute_force():
    """closest_hit_textured.glsl:77-120, 159-203 end to end for the ground pixels of a C1 scene with a sun disc: blue-noise taps by
    global pixel and sample index (raw 16-bit PNG narrowed by the host, recorded at push_image), golden-ratio animation,
    disc -> direction, one first-hit test per sample in numpy float64, sun_factor = lit / N read back out of the oracle's radiance."""
    o = make_oracle()
    images = []
    push = o.push_image

    def recording_push(texels, fmt, linear):
        images.append(np.array(texels))
        return push(texels, fmt, linear)

    o.push_image = recording_push
    W, H, N, FRAME = 96, 54, 4, 37
    s = build_scene(o, "c1", W, H)
    s.shadow_rays, s.sun_radius = N, 0.08
    r = o.render(s.uniforms(frame_index=FRAME), s.params())
    u = s.uniforms(frame_index=FRAME)
    noise = images[u.blue_noise_texture_index][..., 0].astype(np.float64) / 255.0
    assert noise.shape == (64, 64)
    Vi = np.array(list(u.view_inverse), np.float64).reshape(4, 4).T
    Pi = np.array(list(u.proj_inverse), np.float64).reshape(4, 4).T
    sun = np.array(list(u.sun_dir), np.float64)
    by_id = {mid: arrays for (mid, _, arrays) in s.models.values()}
    tris = []
    for rec in s.instances:
        m = by_id[int(rec["custom_index_and_mask"]) & 0xFFFFFF]
        T = rec["transform"].astype(np.float64).reshape(3, 4)
        P = m.positions.astype(np.float64) @ T[:, :3].T + T[:, 3]
        for g in m.geometries:
            tris.append(P[np.asarray(g.indices, np.int64).reshape(-1, 3)])
    tris = np.concatenate(tris)
    v0, e1, e2 = tris[:, 0], tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]

    def any_hit(oo, dd, tmin):
        hit = np.zeros(len(oo), bool)
        for a in range(0, len(tris), 512):
            sl = slice(a, a + 512)
            p = np.cross(dd[:, None, :], e2[None, sl])
            det = np.einsum("ntk,tk->nt", p, e1[sl])
            with np.errstate(divide="ignore", invalid="ignore"):
                inv = 1.0 / det
                tv = oo[:, None, :] - v0[None, sl]
                uu = np.einsum("ntk,ntk->nt", tv, p) * inv
                q = np.cross(tv, e1[None, sl])
                vv = np.einsum("nk,ntk->nt", dd, q) * inv
                tt = np.einsum("ntk,tk->nt", q, e2[sl]) * inv
            hit |= np.any((det != 0) & (uu >= 0) & (vv >= 0) & (uu + vv <= 1) & (tt > tmin) & (tt < 10000.0), axis=1)
        return hit

    ground = np.argwhere(r["hit_ids"][:, :, 0, 0] == 0)[::5]          # instance 0 = the plane
    py, px = ground[:, 0], ground[:, 1]
    ndc = np.stack([(px + 0.5) / W * 2 - 1, (py + 0.5) / H * 2 - 1, np.ones(len(px)), np.ones(len(px))], axis=-1) @ Pi.T
    ld = ndc[:, :3] / np.linalg.norm(ndc[:, :3], axis=1, keepdims=True)
    d = ld @ Vi[:3, :3].T
    t = -Vi[1, 3] / d[:, 1]                                           # the plane is y = 0
    origin = Vi[:3, 3] + d * t[:, None]                               # flat ground: the terminator fix adds nothing
    tangent = np.cross(sun, (0.0, 1.0, 0.0)); tangent /= np.linalg.norm(tangent)
    bitangent = np.cross(tangent, sun); bitangent /= np.linalg.norm(bitangent)
    lit = np.zeros(len(px))
    for i in range(N):
        ax, ay = (px + 2 * i * 13) % 64, (py + 2 * i * 41) % 64
        bx, by = (px + (2 * i + 1) * 13) % 64, (py + (2 * i + 1) * 41) % 64
        # f32 like the shader for the fract() wrap, float64 afterwards
        shift = np.float32(FRAME % 32) * np.float32(0.618033988749)
        xi_x = (noise[ay, ax].astype(np.float32) + shift).astype(np.float32); xi_x = (xi_x - np.floor(xi_x)).astype(np.float64)
        xi_y = (noise[by, bx].astype(np.float32) + shift).astype(np.float32); xi_y = (xi_y - np.floor(xi_y)).astype(np.float64)
        rad, ang = np.sqrt(xi_x), xi_y * 2.0 * math.pi
        pt = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1) * s.sun_radius
        sd = sun + pt[:, :1] * tangent + pt[:, 1:] * bitangent
        sd /= np.linalg.norm(sd, axis=1, keepdims=True)
        lit += ~any_hit(origin + np.array([0.0, 1e-5, 0.0]), sd, 0.001)
    want_factor = lit / N
    srgb = lambda c: c / 12.92 if c <= 0.04045 else ((c + 0.055) / 1.055) ** 2.4
    base = np.array([srgb(103 / 255), srgb(234 / 255), srgb(64 / 255)])
    got_rad = r["radiance"][py, px].astype(np.float64)
    got_factor = np.array([(got_rad[k, 1] - 0.1 * base[1]) / brdf64((0.0, 1.0, 0.0), -d[k], sun, base, 1.0, 0.0, 1.0)[1] for k in range(len(px))])
    assert np.all(np.abs(got_factor * N - np.round(got_factor * N)) < 1e-2)        # multiples of 1 / N
    assert len(set(np.round(got_factor * N).astype(int).tolist())) >= 3              # a penumbra: several levels present
    assert np.mean(np.abs(got_factor - want_factor) < 1e-2) >= 0.97, np.mean(np.abs(got_factor - want_factor) < 1e-2)
    assert int(r["ray_counts"][1]) == N * int((r["hit_ids"][:, :, 0, 0] != abi.MISS_ID).sum())
    o.close()


def test_mirror_bounce_and_alpha_clip_against_an_independent_float64_brute_force():
    """Same idea for the rest of the trace semantics, on a small C3: the alpha-clip any-hit (any_hit_alpha_clip.glsl:11-28: a
    candidate on non-opaque geometry counts only if the bilinear alpha at its uv is >= 0.5), closest_hit_mirror
    (closest_hit_mirror.glsl:11-29: n = normalize(inverse-transpose * interpolated normal), reflect, origin = o + d t) and the
    second ray-gen segment — restated in numpy float64 over world-space triangles, with raw texels recorded at push_image."""
    o = make_oracle()
    images = []
    push = o.push_image

    def recording_push(texels, fmt, linear):
        images.append(np.array(texels))
        return push(texels, fmt, linear)

    o.push_image = recording_push
    W, H = 64, 36
    s = build_scene(o, "c3", W, H, num_instances=12)
    r = o.render(s.uniforms(), s.params())
    u = s.uniforms()
    Vi = np.array(list(u.view_inverse), np.float64).reshape(4, 4).T
    Pi = np.array(list(u.proj_inverse), np.float64).reshape(4, 4).T
    xs, ys = np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5)
    target = np.stack([xs / W * 2 - 1, ys / H * 2 - 1, np.ones_like(xs), np.ones_like(xs)], axis=-1) @ Pi.T
    ld = target[..., :3] / np.linalg.norm(target[..., :3], axis=-1, keepdims=True)
    d0 = (ld @ Vi[:3, :3].T).reshape(-1, 3)
    o0 = np.broadcast_to(Vi[:3, 3], d0.shape).copy()

    by_id = {mid: arrays for (mid, _, arrays) in s.models.values()}
    V0, E1, E2, IDS, NRM, UV, ALPHA_TEX, KIND, NMAT = [], [], [], [], [], [], [], [], []
    for ii, rec in enumerate(s.instances):
        m = by_id[int(rec["custom_index_and_mask"]) & 0xFFFFFF]
        T = rec["transform"].astype(np.float64).reshape(3, 4)
        P = m.positions.astype(np.float64) @ T[:, :3].T + T[:, 3]
        nmat = np.linalg.inv(T[:, :3]).T  # mat3(gl_WorldToObject3x4EXT) * n == inverse-transpose * n
        for gi, g in enumerate(m.geometries):
            idx = np.asarray(g.indices, np.int64).reshape(-1, 3)
            V0.append(P[idx[:, 0]]); E1.append(P[idx[:, 1]] - P[idx[:, 0]]); E2.append(P[idx[:, 2]] - P[idx[:, 0]])
            IDS.append(np.stack([np.full(len(idx), ii), np.full(len(idx), gi), np.arange(len(idx))], axis=1))
            NRM.append(m.normals.astype(np.float64)[idx]); UV.append(m.uvs.astype(np.float64)[idx])
            ALPHA_TEX.append(np.full(len(idx), -1 if g.opaque else g.diffuse_image_index))
            KIND.append(np.full(len(idx), int(rec["sbt_offset_and_flags"]) & 0xFFFFFF))
            NMAT.append(np.broadcast_to(nmat, (len(idx), 3, 3)))
    V0, E1, E2, IDS, NRM, UV, ALPHA_TEX, KIND, NMAT = (np.concatenate(a) for a in (V0, E1, E2, IDS, NRM, UV, ALPHA_TEX, KIND, NMAT))
    assert (ALPHA_TEX >= 0).sum() >= 2  # the fence quads

    def alpha_at(tex, uv):
        a = images[tex][..., 3].astype(np.float64) / 255.0
        h, w = a.shape
        x, y = uv[:, 0] * w - 0.5, uv[:, 1] * h - 0.5
        x0, y0 = np.floor(x), np.floor(y)
        fx, fy = x - x0, y - y0
        xi0, xi1, yi0, yi1 = (x0.astype(int) % w), ((x0.astype(int) + 1) % w), (y0.astype(int) % h), ((y0.astype(int) + 1) % h)
        return (a[yi0, xi0] * (1 - fx) + a[yi0, xi1] * fx) * (1 - fy) + (a[yi1, xi0] * (1 - fx) + a[yi1, xi1] * fx) * fy

    def closest(oo, dd):
        n = len(oo)
        best_t, best_i, best_u, best_v = np.full(n, np.inf), np.full(n, -1), np.zeros(n), np.zeros(n)
        for a in range(0, len(V0), 512):
            sl = slice(a, min(a + 512, len(V0)))
            p = np.cross(dd[:, None, :], E2[None, sl])
            det = np.einsum("ntk,tk->nt", p, E1[sl])
            with np.errstate(divide="ignore", invalid="ignore"):
                inv = 1.0 / det
                tv = oo[:, None, :] - V0[None, sl]
                uu = np.einsum("ntk,ntk->nt", tv, p) * inv
                q = np.cross(tv, E1[None, sl])
                vv = np.einsum("nk,ntk->nt", dd, q) * inv
                tt = np.einsum("ntk,tk->nt", q, E2[sl]) * inv
            ok = (det != 0) & (uu >= 0) & (vv >= 0) & (uu + vv <= 1) & (tt > 0.01) & (tt < 10000.0)
            for col in np.nonzero(ALPHA_TEX[sl] >= 0)[0]:  # any-hit: ignoreIntersection when alpha < 0.5
                rows = np.nonzero(ok[:, col])[0]
                if len(rows):
                    w = np.stack([1 - uu[rows, col] - vv[rows, col], uu[rows, col], vv[rows, col]], axis=1)
                    uv = np.einsum("nk,kc->nc", w, UV[a + col])
                    ok[rows, col] = alpha_at(int(ALPHA_TEX[a + col]), uv) >= 0.5
            tt = np.where(ok, tt, np.inf)
            k = tt.argmin(axis=1)
            rng = np.arange(n)
            better = tt[rng, k] < best_t
            best_t = np.where(better, tt[rng, k], best_t)
            best_i = np.where(better, a + k, best_i)
            best_u = np.where(better, uu[rng, k], best_u)
            best_v = np.where(better, vv[rng, k], best_v)
        return best_i, best_t, best_u, best_v

    i0, t0, u0, v0 = closest(o0, d0)
    want0 = np.where(i0[:, None] >= 0, IDS[np.maximum(i0, 0)], abi.MISS_ID)
    got = r["hit_ids"].reshape(-1, 3, 3).astype(np.int64)
    agree0 = np.all(got[:, 0] == want0, axis=1)
    assert agree0.mean() >= 0.99, agree0.mean()
    fence_insts = sorted(set(IDS[ALPHA_TEX >= 0][:, 0].tolist()))
    seen_fence = np.isin(want0[:, 0], fence_insts).sum()
    assert 0 < seen_fence < (W * H) // 2          # the fence is in view, and rays get through its holes

    mirror = (i0 >= 0) & (KIND[np.maximum(i0, 0)] == abi.RT_HIT_MIRROR) & agree0
    assert mirror.sum() > 30
    mi = i0[mirror]
    w = np.stack([1 - u0[mirror] - v0[mirror], u0[mirror], v0[mirror]], axis=1)
    n_obj = np.einsum("nk,nkc->nc", w, NRM[mi])
    n_w = np.einsum("nij,nj->ni", NMAT[mi], n_obj)
    n_w /= np.linalg.norm(n_w, axis=1, keepdims=True)
    d1 = d0[mirror] - 2.0 * np.einsum("nk,nk->n", n_w, d0[mirror])[:, None] * n_w
    o1 = o0[mirror] + d0[mirror] * t0[mirror][:, None]
    i1, _, _, _ = closest(o1, d1)
    want1 = np.where(i1[:, None] >= 0, IDS[np.maximum(i1, 0)], abi.MISS_ID)
    agree1 = np.all(got[mirror, 1] == want1, axis=1)
    assert agree1.mean() >= 0.95, agree1.mean()
    # pixels whose first hit is not a mirror / portal never trace a second segment
    assert np.all(got[(i0 >= 0) & (KIND[np.maximum(i0, 0)] == abi.RT_HIT_TEXTURED) & agree0, 1] == abi.MISS_ID)
    o.close()


# ---------------------------------------------------------------- show_heatmap (heatmap.rs, lib.rs:120-124, 174-186)
HEAT_COLOURS = np.array([(0, 2, 91), (0, 108, 251), (0, 221, 221), (51, 221, 0), (255, 252, 0), (255, 180, 0), (255, 104, 0),
                         (226, 22, 0), (191, 0, 83), (145, 0, 65)], np.float64) / 255.0


def heatmap64(heat):
    """float64 re-derivation of heatmap.rs:5-54 (with `cur` clamped to the table, see the oracle)."""
    def sat(x):
        return min(max(x, 0.0), 1.0)

    def smooth(e0, e1, x):
        t = sat((x - e0) / (e1 - e0))
        return t * t * (3.0 - 2.0 * t)

    # the function jumps where heat * 10 crosses an integer (exact integers take the floor == ceil branch), so the
    # product is rounded to f32 like the shader's; everything after it is float64
    h = float(np.float32(sat(heat)) * np.float32(10.0))
    idx = int(h)
    cur, prv, nxt = min(idx, 9), max(idx - 1, 0), min(idx + 1, 9)
    lo, hi = math.floor(h), math.ceil(h)
    s_lo, s_hi = smooth(lo - 0.8, lo + 0.8, h), smooth(hi - 0.8, hi + 0.8, h)
    r = s_lo * (1.0 - s_hi) * HEAT_COLOURS[cur] + (1.0 - s_lo) * HEAT_COLOURS[prv] + s_hi * HEAT_COLOURS[nxt]
    return np.clip(r, 0.0, 1.0)


@pytest.mark.skipif(not os.path.exists("/root/reference/shaders/ray_generation.spv"), reason="reference tree not present")
def test_shipped_ray_generation_spirv_holds_the_heatmap_constants():
    floats, ints = _spv_constants("/root/reference/shaders/ray_generation.spv")
    f32s = np.asarray(floats, np.float32)
    for want in (1_000_000.0, 0.000001, 0.8, 10.0, 3.0, 2.0, 0.01, 10000.0, 0.0031308, 12.92, 1.055, 0.055):
        assert np.any(f32s == np.float32(want)), want
    for c in sorted({int(v) for v in (HEAT_COLOURS * 255.0).round().ravel()} - {0, 255}):
        assert np.any(f32s == np.float32(c) / np.float32(255.0)), c  # the table entries, bit-exact f32 quotients
    assert 9 in ints and 10 in ints


def test_heatmap_temperature_against_float64(orc):
    heats = np.concatenate([np.linspace(-0.2, 1.2, 1401), np.arange(0, 11) / 10.0, [0.05, 0.149999, 0.150001, 0.999999]])
    for h in heats:
        got = orc.heatmap_temperature(float(np.float32(h)))
        want = heatmap64(float(np.float32(h)))
        assert np.all(np.abs(got - want) <= 2e-6), (h, got, want)
    # cold end: heat 0 -> floor == ceil == 0: s_lo = s_hi = 0.5 -> 0.25*c0 + 0.5*c0 + 0.5*c1
    assert np.allclose(orc.heatmap_temperature(0.0), 0.75 * HEAT_COLOURS[0] + 0.5 * HEAT_COLOURS[1], atol=1e-6)
    # hot end: one past the table in the reference; clamped here -> 1.25 * c9, clamped to [0, 1]
    assert np.allclose(orc.heatmap_temperature(1.0), np.clip(1.25 * HEAT_COLOURS[9], 0, 1), atol=1e-6)
    assert np.array_equal(orc.heatmap_temperature(7.0), orc.heatmap_temperature(1.0))  # saturate
    # lib.rs:182: the payload colour only tints the heat by 1e-6
    base = orc.heatmap_pixel(250_000, (0.0, 0.0, 0.0))
    assert np.array_equal(base, orc.heatmap_temperature(0.25))
    assert np.allclose(orc.heatmap_pixel(250_000, (1.0, 0.5, 0.25)) - base, (1e-6, 0.5e-6, 0.25e-6), atol=1e-7)
    assert np.array_equal(orc.heatmap_pixel(500, (0, 0, 0), scale=1000.0), orc.heatmap_temperature(0.5))


def test_oracle_heatmap_frame_uses_the_stand_in_clock():
    o = make_oracle()
    s = build_scene(o, "c1", 96, 54)
    plain = o.render(s.uniforms(), s.params())
    u = s.uniforms()
    u.show_heatmap = 1
    heat = o.render(u, s.params(), want=("rgba8", "radiance", "hit_ids", "ray_counts", "cost_cycles"))
    assert np.array_equal(heat["hit_ids"], plain["hit_ids"]) and np.array_equal(heat["ray_counts"], plain["ray_counts"])
    hit = plain["hit_ids"][:, :, 0, 0] != abi.MISS_ID
    assert np.all(heat["cost_cycles"][~hit] == 20000) and np.all(heat["cost_cycles"][hit] == 20000 * (1 + s.shadow_rays))
    assert int(heat["cost_cycles"].sum()) == 20000 * int(plain["ray_counts"].sum())
    for (y, x) in ((0, 0), (53, 95), (40, 48), (27, 48)):
        want = o.heatmap_pixel(heat["cost_cycles"][y, x], plain["radiance"][y, x])
        assert np.array_equal(heat["radiance"][y, x], want)
        assert tuple(heat["rgba8"][y, x][:3]) == tuple(o.unorm8(o.linear_to_srgb(c)) for c in want)
    # without show_heatmap the cost output is left alone
    assert not np.any(o.render(s.uniforms(), s.params(), want=("cost_cycles",))["cost_cycles"])
    o.close()


# ---------------------------------------------------------------- committed golden fixtures
@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "default"])
def test_golden_fixture(cfg):
    from PIL import Image

    meta = json.load(open(os.path.join(GOLDEN, f"{cfg}.json")))
    o = make_oracle()
    s = build_scene(o, cfg, meta["width"], meta["height"], **meta.get("kwargs", {}))
    r = o.render(s.uniforms(), s.params())
    assert hashlib.sha256(r["hit_ids"].tobytes()).hexdigest() == meta["hit_ids_sha256"]
    assert [int(x) for x in r["ray_counts"]] == meta["ray_counts"]
    want = np.asarray(Image.open(os.path.join(GOLDEN, f"{cfg}.png")))
    assert np.abs(r["rgba8"].astype(int) - want.astype(int)).max() <= 1
    o.close()


def test_synthetic_model_exercises_normal_map_nearest_and_multi_material():
    """SURVEY 8f-1 paths on the oracle: both geometries of the two-material model are hit, the masked one runs the
    any-hit (rays pass through its holes), the normal map changes the shading, NEAREST differs from LINEAR."""
    import synth_assets
    from oracle.binding import Oracle

    o = Oracle()
    s = synth_assets.build_bumpy_scene(o, 320, 180)
    r = o.render(s.uniforms(), s.params())
    assert np.isfinite(r["radiance"]).all()
    ids = r["hit_ids"][:, :, 0, :]
    on_model = (ids[..., 0] == 1) | (ids[..., 0] == 2)
    assert on_model.sum() > 1500
    assert set(np.unique(ids[on_model][:, 1]).tolist()) == {0, 1}           # gl_GeometryIndexEXT: one geometry per material
    g1 = on_model & (ids[..., 1] == 1)
    g0 = on_model & (ids[..., 1] == 0)
    assert 0.25 < g1.sum() / max(g0.sum(), 1) < 0.9                          # holes of the masked checker let rays through
    # without the normal map the opaque geometry shades differently
    arrays = s.models["bumpy"][2]
    o2 = Oracle()
    import ray_tracing_gallery_b200.gltf as gltf
    orig = gltf.load_gltf

    def no_normal_map(*a, **k):
        m = orig(*a, **k)
        m.geometries[0].normal_map_image_index = -1
        return m

    gltf.load_gltf = no_normal_map
    try:
        s2 = synth_assets.build_bumpy_scene(o2, 320, 180)
    finally:
        gltf.load_gltf = orig
    r2 = o2.render(s2.uniforms(), s2.params())
    assert np.array_equal(r2["hit_ids"], r["hit_ids"])
    d = np.abs(r2["radiance"] - r["radiance"]).max(axis=2)
    assert d[g0].mean() > 1e-3 and d[g1].max() == 0.0
    assert arrays.geometries[0].normal_map_image_index >= 0
    o.close(); o2.close()


def test_random_stress_scene_tie_rules_on_the_oracle():
    """Exact ties by construction: a duplicated instance and twin triangles.  Lowest (instance, geometry, primitive) wins."""
    import synth_assets
    from oracle.binding import Oracle

    o = Oracle()
    s = synth_assets.build_random_scene(o, 1)
    r = o.render(s.uniforms(), s.params())
    ids = r["hit_ids"].reshape(-1, 3)
    hit = ids[ids[:, 0] != 0xFFFFFFFF]
    assert 15 not in set(hit[:, 0].tolist()) and 3 in set(hit[:, 0].tolist())      # instance 15 is an exact copy of instance 3
    twins = hit[np.isin(hit[:, 0], [i for i in range(1, 15) if (i - 1) % 3 == 2])]      # instances of the twin-triangle model
    assert len(twins) > 0 and twins[:, 2].max() < 12                                    # the second copy (prims 12..23) never wins
    assert np.isfinite(r["radiance"]).all()
    o.close()

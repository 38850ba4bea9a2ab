"""Synthetic GLB models for the loader / shading paths that none of the reference's bundled assets exercises
(SURVEY.md 8f item 1): several materials in one model (one geometry per material, src/util_structs.rs:1079-1137),
a normal map (shaders/closest_hit_textured.glsl:141-157), a NEAREST magFilter (src/util_structs.rs:954-955),
non-power-of-two images, uvs outside [0,1] (REPEAT), u16 and u32 indices, un-normalised vertex normals, an
alpha-masked material next to an opaque one.  Everything is generated from a fixed seed."""
import io
import json
import struct

import numpy as np


def _png(rgba: np.ndarray) -> bytes:
    from PIL import Image

    buf = io.BytesIO()
    Image.fromarray(rgba, "RGBA").save(buf, format="PNG")
    return buf.getvalue()


def _pad4(b: bytes, fill=b"\x00") -> bytes:
    return b + fill * ((4 - len(b) % 4) % 4)


def bumpy_two_material_glb(seed: int = 7, n: int = 10) -> bytes:
    """An n x n quad grid over [-1,1]^2 in xz with a height field; the quads are dealt to two materials in a
    checkerboard (primitive 0: u16 indices, material 1; primitive 1: u32 indices, material 0)."""
    rng = np.random.default_rng(seed)
    g = np.linspace(-1.0, 1.0, n + 1, dtype=np.float32)
    xx, zz = np.meshgrid(g, g, indexing="xy")
    yy = (0.15 * np.sin(3.0 * xx) * np.cos(2.0 * zz)).astype(np.float32)
    pos = np.stack([xx, yy, zz], axis=-1).reshape(-1, 3).astype(np.float32)
    # analytic normals, deliberately NOT unit length (the shaders normalise after the transform)
    dydx = 0.45 * np.cos(3.0 * xx) * np.cos(2.0 * zz)
    dydz = -0.30 * np.sin(3.0 * xx) * np.sin(2.0 * zz)
    nrm = np.stack([-dydx, np.ones_like(xx), -dydz], axis=-1).reshape(-1, 3).astype(np.float32) * np.float32(1.7)
    uv = np.stack([(xx + 1.0) * 1.3 - 0.4, (zz + 1.0) * 0.9 - 0.3], axis=-1).reshape(-1, 2).astype(np.float32)  # leaves [0,1]

    def quad(ix, iz):
        a = iz * (n + 1) + ix
        return [a, a + n + 1, a + 1, a + 1, a + n + 1, a + n + 2]

    idx0, idx1 = [], []
    for iz in range(n):
        for ix in range(n):
            (idx0 if (ix + iz) % 2 == 0 else idx1).extend(quad(ix, iz))
    idx0 = np.asarray(idx0, np.uint16)
    idx1 = np.asarray(idx1, np.uint32)

    # images: 0 = base colour of material 0 (20x12, NEAREST), 1 = metallic-roughness (8x8, LINEAR),
    #         2 = normal map (32x24, LINEAR), 3 = base colour + alpha mask of material 1 (16x16, LINEAR)
    img0 = rng.integers(40, 255, (12, 20, 4), dtype=np.uint8); img0[..., 3] = 255
    img1 = rng.integers(0, 255, (8, 8, 4), dtype=np.uint8); img1[..., 3] = 255
    ny, nx = np.meshgrid(np.linspace(0, 4 * np.pi, 24), np.linspace(0, 6 * np.pi, 32), indexing="ij")
    nvec = np.stack([0.35 * np.sin(nx), 0.35 * np.cos(ny), np.ones_like(nx)], axis=-1)
    nvec /= np.linalg.norm(nvec, axis=-1, keepdims=True)
    img2 = np.concatenate([np.round((nvec * 0.5 + 0.5) * 255), np.full(nx.shape + (1,), 255.0)], axis=-1).astype(np.uint8)
    img3 = rng.integers(30, 255, (16, 16, 4), dtype=np.uint8)
    cy, cx = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    img3[..., 3] = np.where(((cx // 4) + (cy // 4)) % 2 == 0, 255, 20)  # MASK: coarse checker of holes
    pngs = [_png(img0), _png(img1), _png(img2), _png(img3)]

    chunks, views = [], []

    def add_view(b: bytes, target=None):
        off = sum(len(c) for c in chunks)
        chunks.append(_pad4(b))
        v = {"buffer": 0, "byteOffset": off, "byteLength": len(b)}
        if target:
            v["target"] = target
        views.append(v)
        return len(views) - 1

    v_pos, v_nrm, v_uv = add_view(pos.tobytes(), 34962), add_view(nrm.tobytes(), 34962), add_view(uv.tobytes(), 34962)
    v_i0, v_i1 = add_view(idx0.tobytes(), 34963), add_view(idx1.tobytes(), 34963)
    v_img = [add_view(p) for p in pngs]
    nv = len(pos)
    accessors = [
        {"bufferView": v_pos, "componentType": 5126, "count": nv, "type": "VEC3", "min": pos.min(0).tolist(), "max": pos.max(0).tolist()},
        {"bufferView": v_nrm, "componentType": 5126, "count": nv, "type": "VEC3"},
        {"bufferView": v_uv, "componentType": 5126, "count": nv, "type": "VEC2"},
        {"bufferView": v_i0, "componentType": 5123, "count": len(idx0), "type": "SCALAR"},
        {"bufferView": v_i1, "componentType": 5125, "count": len(idx1), "type": "SCALAR"},
    ]
    doc = {
        "asset": {"version": "2.0", "generator": "tests/synth_assets.py"},
        "buffers": [{"byteLength": sum(len(c) for c in chunks)}],
        "bufferViews": views,
        "accessors": accessors,
        "images": [{"bufferView": v, "mimeType": "image/png"} for v in v_img],
        "samplers": [{"magFilter": 9728, "minFilter": 9728}, {"magFilter": 9729, "minFilter": 9729}],
        "textures": [{"source": 0, "sampler": 0}, {"source": 1, "sampler": 1}, {"source": 2, "sampler": 1}, {"source": 3}],
        "materials": [
            {"name": "opaque_normal_mapped", "pbrMetallicRoughness": {"baseColorTexture": {"index": 0}, "metallicRoughnessTexture": {"index": 1}},
             "normalTexture": {"index": 2}},
            {"name": "masked", "alphaMode": "MASK", "alphaCutoff": 0.5, "doubleSided": True,
             "pbrMetallicRoughness": {"baseColorTexture": {"index": 3}, "metallicFactor": 0.25, "roughnessFactor": 0.6}},
        ],
        "meshes": [{"primitives": [
            {"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3, "material": 1},
            {"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 4, "material": 0},
        ]}],
        "nodes": [{"mesh": 0, "scale": [3.0, 3.0, 3.0]}],  # node transforms are ignored by the loader (:1113)
        "scenes": [{"nodes": [0]}],
        "scene": 0,
    }
    js = _pad4(json.dumps(doc).encode("utf-8"), b" ")
    blob = b"".join(chunks)
    total = 12 + 8 + len(js) + 8 + len(blob)
    return struct.pack("<III", 0x46546C67, 2, total) + struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(blob), 0x004E4942) + blob


def build_bumpy_scene(backend, width=640, height=360, shadow_rays=2):
    """plane + two instances of the synthetic model (one non-uniformly scaled and rotated) + a mirror torus."""
    import numpy as np

    from ray_tracing_gallery_b200 import abi
    from ray_tracing_gallery_b200.gltf import load_gltf
    from ray_tracing_gallery_b200.scene import (Camera, SceneSetup, Sun, load_model, make_instance, mat_rotation_y, mat_scale,
                                                mat_translation, push_builtin_images)

    push_builtin_images(backend)
    pid, ph, _ = load_model(backend, "plane.glb", 0)
    tid, th, _ = load_model(backend, "tori.glb", 1)
    arrays = load_gltf(bumpy_two_material_glb(), "bumpy", 1, backend.push_image)
    bid, bh = backend.create_model(arrays)
    squash = np.diag([1.6, 0.8, 1.1, 1.0]).astype(np.float32)
    inst = np.stack([
        make_instance(mat_scale(10.0), pid, ph, abi.RT_HIT_TEXTURED),
        make_instance(mat_translation(-1.2, 0.8, 0.5) @ mat_rotation_y(0.4) @ squash, bid, bh, abi.RT_HIT_TEXTURED, True),
        make_instance(mat_translation(1.6, 1.4, 1.5) @ mat_rotation_y(-0.9), bid, bh, abi.RT_HIT_TEXTURED, True),
        make_instance(mat_translation(0.5, 1.0, 4.0) @ mat_rotation_y(1.1), tid, th, abi.RT_HIT_MIRROR),
    ])
    s = SceneSetup("bumpy", inst, Camera(eye=(0.0, 3.0, -4.5), pitch=-0.3), Sun(), width, height, shadow_rays=shadow_rays, sun_radius=0.05,
                   description="synthetic two-material normal-mapped model")
    s.models = {"bumpy": (bid, bh, arrays)}
    backend.build_tlas(s.instances)
    return s


def build_random_scene(backend, seed: int, width=256, height=144, offset=(0.0, 0.0, 0.0)):
    """Seeded random stress scene for the traversal semantics: a triangle soup in two geometries (one alpha-masked with a
    random NEAREST texture), a second model with exactly coincident duplicate triangles (exact t ties: lowest ids must win),
    instances with arbitrary rotations, non-uniform and NEGATIVE scales, mirror / portal / textured hit groups, one
    degenerate (zero-area) triangle, a camera orbiting the origin.  `offset` moves the whole scene (instances and camera) away
    from the origin: large coordinates against small node grids are where the traversal's box test has least room."""
    from ray_tracing_gallery_b200 import abi
    from ray_tracing_gallery_b200.gltf import Geometry, ModelArrays
    from ray_tracing_gallery_b200.scene import Camera, SceneSetup, Sun, load_model, make_instance, mat_scale, push_builtin_images

    rng = np.random.default_rng(seed)
    push_builtin_images(backend)
    pid, ph, _ = load_model(backend, "plane.glb", 0)

    def soup(n_tris, spread, size):
        c = rng.uniform(-spread, spread, (n_tris, 1, 3))
        v = c + rng.uniform(-size, size, (n_tris, 3, 3))
        pos = v.reshape(-1, 3).astype(np.float32)
        nrm = np.repeat(np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]), 3, axis=0).astype(np.float32)
        nrm[np.all(nrm == 0, axis=1)] = (0, 1, 0)
        uv = rng.uniform(-1.5, 2.5, (n_tris * 3, 2)).astype(np.float32)
        return pos, nrm, uv

    # model A: soup, geometry 0 opaque with a random linear sRGB texture, geometry 1 alpha-masked with a NEAREST one
    pos, nrm, uv = soup(160, 1.0, 0.35)
    pos[9:12] = pos[9]  # a degenerate triangle (three equal vertices): never hit
    tex_a = rng.integers(0, 255, (9, 13, 4), dtype=np.uint8); tex_a[..., 3] = 255
    tex_b = rng.integers(0, 255, (8, 8, 4), dtype=np.uint8)
    ia = backend.push_image(tex_a, abi.RT_FORMAT_RGBA8_SRGB, True)
    ib = backend.push_image(tex_b, abi.RT_FORMAT_RGBA8_SRGB, False)
    mr = backend.push_image(np.asarray([1.0, 0.45, 0.3, 1.0], np.float32).reshape(1, 1, 4), abi.RT_FORMAT_RGBA32_SFLOAT, False)
    idx = np.arange(160 * 3, dtype=np.uint32)
    model_a = ModelArrays("soup", pos, nrm, uv, [Geometry(idx[: 100 * 3], True, ia, mr, -1), Geometry(idx[100 * 3 :], False, ib, mr, -1)])
    aid, ah = backend.create_model(model_a)
    # model B: 12 triangles, every one stored twice (exactly coincident): the tie rule decides which primitive id is reported
    pos, nrm, uv = soup(12, 0.6, 0.5)
    pos2, nrm2, uv2 = np.concatenate([pos, pos]), np.concatenate([nrm, nrm]), np.concatenate([uv, uv])
    model_b = ModelArrays("twins", pos2, nrm2, uv2, [Geometry(np.arange(72, dtype=np.uint32), True, 1, mr, -1)])
    bid, bh = backend.create_model(model_b)

    def random_transform():
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        w, x, y, z = q
        rot = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        sc = rng.uniform(0.4, 1.6, 3) * rng.choice([1.0, 1.0, 1.0, -1.0], 3)  # some axes mirrored: negative determinants
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] = (rot * sc).astype(np.float32)
        m[:3, 3] = rng.uniform(-3.0, 3.0, 3).astype(np.float32) + np.float32([0, 2.0, 0])
        return m

    inst = [make_instance(mat_scale(10.0), pid, ph, abi.RT_HIT_TEXTURED)]
    for k in range(14):
        model, handle = ((aid, ah), (bid, bh))[k % 3 == 2]
        kind = (abi.RT_HIT_TEXTURED, abi.RT_HIT_TEXTURED, abi.RT_HIT_MIRROR, abi.RT_HIT_PORTAL)[int(rng.integers(0, 4))]
        inst.append(make_instance(random_transform(), model, handle, kind, True))
    inst.append(inst[3].copy())  # an exact duplicate INSTANCE: every hit on it ties with instance 3, the lower gl_InstanceID wins
    off = np.float32(offset)
    if np.any(off != 0):
        for rec in inst:
            rec["transform"].reshape(3, 4)[:, 3] += off  # row-major 3x4: translation column
    ang = rng.uniform(0, 2 * np.pi)
    cam = Camera(eye=(float(9 * np.sin(ang)) + float(off[0]), float(rng.uniform(2.5, 5.0)) + float(off[1]), float(-9 * np.cos(ang)) + float(off[2])),
                 pitch=-0.25, yaw=float(np.pi - ang))
    s = SceneSetup(f"random{seed}", np.stack(inst), cam, Sun(pitch=float(rng.uniform(0.3, 1.2)), yaw=float(rng.uniform(0, 6.28))), width, height,
                   shadow_rays=int(rng.integers(1, 6)), sun_radius=float(rng.uniform(0.0, 0.08)), description="seeded random stress scene")
    backend.build_tlas(s.instances)
    return s


def build_rod_scene(backend, seed=11, width=320, height=180):
    """Instances whose exact world bounds are far smaller than the box of their transformed object-space corners: thin rods along
    the body diagonal of their box, under arbitrary rotations, shears and mirrorings (k_tighten_instance_boxes walks the vertices of
    models with <= 4096 of them).  The rod model carries two vertices no triangle references — one NaN, one 1e30 — and a second
    soup model has 4 500 vertices, so its instances keep the corner box."""
    from ray_tracing_gallery_b200 import abi
    from ray_tracing_gallery_b200.gltf import Geometry, ModelArrays
    from ray_tracing_gallery_b200.scene import Camera, SceneSetup, Sun, load_model, make_instance, mat_scale, push_builtin_images

    rng = np.random.default_rng(seed)
    push_builtin_images(backend)
    pid, ph, _ = load_model(backend, "plane.glb", 0)
    mr = backend.push_image(np.asarray([1.0, 0.5, 0.2, 1.0], np.float32).reshape(1, 1, 4), abi.RT_FORMAT_RGBA32_SFLOAT, False)

    def soup(centres, size):
        v = centres[:, None, :] + rng.uniform(-size, size, (len(centres), 3, 3))
        pos = v.reshape(-1, 3).astype(np.float32)
        nrm = np.repeat(np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]), 3, axis=0).astype(np.float32)
        nrm[np.all(nrm == 0, axis=1)] = (0, 1, 0)
        return pos, nrm, rng.uniform(0, 1, (len(pos), 2)).astype(np.float32)

    t = rng.uniform(-1.0, 1.0, (220, 1))
    pos, nrm, uv = soup(t * np.float64([1.5, 1.0, 1.2]), 0.06)
    n_ref = len(pos)
    pos = np.concatenate([pos, np.float32([[np.nan, 0, 0], [1e30, -1e30, 1e30]])])
    nrm = np.concatenate([nrm, np.float32([[0, 1, 0], [0, 1, 0]])])
    uv = np.concatenate([uv, np.zeros((2, 2), np.float32)])
    rid, rh = backend.create_model(ModelArrays("rod", pos, nrm, uv, [Geometry(np.arange(n_ref, dtype=np.uint32), True, 1, mr, -1)]))
    pos, nrm, uv = soup(rng.uniform(-1.0, 1.0, (1500, 3)) * np.float64([1.0, 0.15, 1.0]), 0.05)
    bid, bh = backend.create_model(ModelArrays("slab", pos, nrm, uv, [Geometry(np.arange(len(pos), dtype=np.uint32), True, 0, mr, -1)]))

    inst = [make_instance(mat_scale(10.0), pid, ph, abi.RT_HIT_TEXTURED)]
    for k in range(40):
        a = rng.normal(size=(3, 3))
        q, _ = np.linalg.qr(a)
        shear = np.eye(3); shear[0, 1] = rng.uniform(-0.5, 0.5)
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] = (q @ shear * rng.uniform(0.3, 1.0, 3) * rng.choice([1.0, -1.0], 3)).astype(np.float32)
        m[:3, 3] = (rng.uniform(-4.0, 4.0, 3) * np.float64([1, 0.3, 1]) + np.float64([0, 1.8, 0])).astype(np.float32)
        model, handle = ((rid, rh), (bid, bh))[k % 5 == 4]
        inst.append(make_instance(m, model, handle, (abi.RT_HIT_TEXTURED, abi.RT_HIT_MIRROR)[k % 7 == 3], True))
    s = SceneSetup("rods", np.stack(inst), Camera(eye=(0.5, 4.0, -9.0), pitch=-0.3), Sun(pitch=0.8, yaw=1.0), width, height,
                   shadow_rays=3, sun_radius=0.04, description="rods: exact instance bounds against corner boxes")
    backend.build_tlas(s.instances)
    return s


def build_mosaic_scene(backend, n_images=72, cells=(36, 20), width=480, height=270, seed=5):
    """A wall of small quads, each cell with its own material out of `n_images` real (non-1x1) images, dealt so that
    neighbouring cells differ: the 8x4-pixel tile a warp traces covers half a dozen images.  Every third material is
    alpha-masked (random 0 / 255 alpha, NEAREST) so that the any-hit stage samples per-lane images inside traversal; the
    others alternate LINEAR / NEAREST and sRGB diffuse + metal-rough pairs.  Uses the bindless table up to its 128 entries."""
    import numpy as np

    from ray_tracing_gallery_b200 import abi
    from ray_tracing_gallery_b200.gltf import Geometry, ModelArrays
    from ray_tracing_gallery_b200.scene import Camera, SceneSetup, Sun, load_model, make_instance, mat_rotation_y, mat_scale, mat_translation, push_builtin_images

    rng = np.random.default_rng(seed)
    push_builtin_images(backend)
    pid, ph, _ = load_model(backend, "plane.glb", 0)
    images = []
    for k in range(n_images):
        masked = k % 3 == 0
        tex = rng.integers(0, 256, (4, 4, 4), dtype=np.uint8)
        tex[..., 3] = rng.choice([0, 255], (4, 4)) if masked else 255
        images.append(backend.push_image(tex, abi.RT_FORMAT_RGBA8_SRGB, linear=(k % 2 == 1) and not masked))
    mr = backend.push_image(np.array([[[1.0, 0.6, 0.1, 1.0]]], np.float32), abi.RT_FORMAT_RGBA32_SFLOAT, False)
    nx, ny = cells
    pos, nrm, uvs = [], [], []
    geo_idx = [[] for _ in range(n_images)]
    for j in range(ny):
        for i in range(nx):
            k = (i * 7 + j * 13 + (i * j) % 5) % n_images
            x0, x1, y0, y1 = i / nx * 6 - 3, (i + 1) / nx * 6 - 3, j / ny * 3.2 + 0.2, (j + 1) / ny * 3.2 + 0.2
            base = len(pos)
            pos += [(x0, y0, 0), (x1, y0, 0), (x1, y1, 0), (x0, y1, 0)]
            nrm += [(0, 0.1, -1)] * 4
            uvs += [(0, 0), (1, 0), (1, 1), (0, 1)]
            geo_idx[k] += [base, base + 1, base + 2, base, base + 2, base + 3]
    geos = [Geometry(np.array(ix, np.uint32), opaque=(k % 3 != 0), diffuse_image_index=images[k], metallic_roughness_image_index=mr,
                     normal_map_image_index=-1) for k, ix in enumerate(geo_idx)]
    arrays = ModelArrays("mosaic", np.array(pos, np.float32), np.array(nrm, np.float32), np.array(uvs, np.float32), geos)
    mid, mh = backend.create_model(arrays)
    inst = np.stack([
        make_instance(mat_scale(10.0), pid, ph, abi.RT_HIT_TEXTURED),
        make_instance(mat_translation(0, 0, 2.0) @ mat_rotation_y(0.15), mid, mh, abi.RT_HIT_TEXTURED, True),
        make_instance(mat_translation(0.4, 0, 3.5) @ mat_rotation_y(-0.3), mid, mh, abi.RT_HIT_TEXTURED, True),
    ])
    s = SceneSetup("mosaic", inst, Camera(eye=(0.0, 1.8, -3.0)), Sun(), width, height, shadow_rays=2, sun_radius=0.05,
                   description=f"wall of quads over {n_images} real images, a third alpha-masked")
    backend.build_tlas(s.instances)
    return s


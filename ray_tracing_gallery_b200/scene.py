"""Host-side scene description: the reference's camera / sun / instance recipes
and the five BASELINE.json configurations.

Reference call sites restated here
  FirstPersonCamera::as_view_matrix     src/main.rs:1206-1231
  Sun::as_normal                        src/main.rs:1189-1197
  Uniforms initial values, fov, near    src/main.rs:568-598
  AccelerationStructureInstance::new    src/gpu_structs.rs:28-52
  DefaultScene instances                src/scene.rs:89-156
  LoadedModelScene                      src/scene.rs:219-254
  DefaultScene::update / write_resources  src/scene.rs:162-204
  built-in textures 0..3                src/main.rs:416-460

The reference's random tori use an unseeded `rand::thread_rng()`
(src/scene.rs:136); here every random scene comes from `hash_uniform`, a
counter-based splitmix64 stream with an explicit seed, so frames are repeatable.
All matrix maths is done in float32 like `ultraviolet` does, but exact crate
behaviour is outside the contract: the C ABI takes the finished matrices.
"""
import math
import os
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import abi
from .gltf import load_gltf, load_png_file_rgba8

ASSET_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")
F = np.float32


# ----------------------------------------------------------------------------- matrices (column vectors, M @ v)
def mat_identity():
    return np.eye(4, dtype=F)


def mat_scale(s):
    m = np.eye(4, dtype=F)
    m[0, 0] = m[1, 1] = m[2, 2] = F(s)
    return m


def mat_translation(x, y, z):
    m = np.eye(4, dtype=F)
    m[0, 3], m[1, 3], m[2, 3] = F(x), F(y), F(z)
    return m


def mat_rotation_y(angle):
    """ultraviolet `Mat4::from_rotation_y`: columns (c,0,-s), (0,1,0), (s,0,c)."""
    s, c = F(math.sin(F(angle))), F(math.cos(F(angle)))
    m = np.eye(4, dtype=F)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    return m


def transpose_matrix_for_instance(m) -> np.ndarray:
    """src/gpu_structs.rs:55-58: rows 0..2 of the object->world matrix, row-major 3x4."""
    return np.ascontiguousarray(np.asarray(m, dtype=F)[:3, :4]).reshape(12)


def make_instance(transform, model_id: int, blas_handle: int, hit_shader: int, double_sided: bool = False):
    """src/gpu_structs.rs:28-52."""
    rec = np.zeros((), abi.INSTANCE_DTYPE)
    rec["transform"] = transpose_matrix_for_instance(transform)
    rec["custom_index_and_mask"] = (model_id & 0xFFFFFF) | (0xFF << 24)
    flags = abi.RT_INSTANCE_TRIANGLE_FACING_CULL_DISABLE if double_sided else 0
    rec["sbt_offset_and_flags"] = (hit_shader & 0xFFFFFF) | (flags << 24)
    rec["blas"] = blas_handle
    return rec


# ----------------------------------------------------------------------------- camera / sun / uniforms
@dataclass
class Camera:
    """FirstPersonCamera (src/main.rs:1206-1231); positive pitch looks up."""

    eye: Tuple[float, float, float] = (0.0, 2.0, -5.0)
    pitch: float = 0.0
    yaw: float = math.pi
    fov_deg: float = 59.0
    near: float = 0.1

    def view_matrix(self):
        ex, ey, ez = (F(v) for v in self.eye)
        p, y = F(self.pitch), F(self.yaw)
        sp, cp, sy, cy = F(math.sin(p)), F(math.cos(p)), F(math.sin(y)), F(math.cos(y))
        xa = np.array([cy, 0, -sy], F)
        ya = np.array([sy * sp, cp, cy * sp], F)
        za = np.array([sy * cp, -sp, cp * cy], F)
        eye = np.array([ex, ey, ez], F)
        m = np.eye(4, dtype=F)
        m[0, :3], m[1, :3], m[2, :3] = xa, ya, za
        m[0, 3], m[1, 3], m[2, 3] = -xa.dot(eye), -ya.dot(eye), -za.dot(eye)
        return m

    def view_inverse(self):
        v = self.view_matrix()
        r = v[:3, :3]
        inv = np.eye(4, dtype=F)
        inv[:3, :3] = r.T
        inv[:3, 3] = -(r.T @ v[:3, 3])
        return inv

    def proj_inverse(self, width, height):
        """Inverse of ultraviolet's `perspective_reversed_infinite_z_vk(fov, aspect, near)`:
        P = cols (sx,0,0,0),(0,-sy,0,0),(0,0,0,-1),(0,0,near,0)."""
        t = F(math.tan(F(math.radians(self.fov_deg)) / F(2)))
        sy = F(1) / t
        sx = sy / (F(width) / F(height))
        m = np.zeros((4, 4), F)
        m[0, 0] = F(1) / sx
        m[1, 1] = F(-1) / sy
        m[2, 3] = F(-1)
        m[3, 2] = F(1) / F(self.near)
        return m


@dataclass
class Sun:
    pitch: float = 0.5
    yaw: float = 1.0

    def as_normal(self):
        p, y = F(self.pitch), F(self.yaw)
        return np.array([math.cos(p) * math.sin(y), math.sin(p), math.cos(p) * math.cos(y)], F)


def make_uniforms(camera: Camera, sun: Sun, width: int, height: int, sun_radius: float, frame_index: int) -> abi.RtUniforms:
    u = abi.RtUniforms()
    vi = camera.view_inverse()
    pi = camera.proj_inverse(width, height)
    u.view_inverse[:] = [float(x) for x in vi.T.reshape(16)]  # column-major
    u.proj_inverse[:] = [float(x) for x in pi.T.reshape(16)]
    sd = sun.as_normal()
    u.sun_dir[:] = [float(x) for x in sd]
    u.sun_radius = float(sun_radius)
    u.blue_noise_texture_index = 2
    u.ggx_lut_texture_index = 3
    u.frame_index = frame_index
    u.show_heatmap = 0
    return u


# ----------------------------------------------------------------------------- the reference's per-tick control integration
@dataclass
class KbdState:
    """`KbdState` of the reference (src/main.rs: the W/S/A/D and arrow-key flags set by the KeyboardInput events)."""

    forward: bool = False
    back: bool = False
    left: bool = False
    right: bool = False
    sun_up: bool = False
    sun_down: bool = False
    sun_cw: bool = False
    sun_ccw: bool = False


@dataclass
class Controls:
    """Camera and sun velocities carried from tick to tick (`camera_velocity`, `sun_velocity`, src/main.rs:600-610)."""

    camera_velocity: np.ndarray = field(default_factory=lambda: np.zeros(3, F))
    sun_velocity: np.ndarray = field(default_factory=lambda: np.zeros(2, F))


def integrate_controls(camera: Camera, sun: Sun, ctl: Controls, kbd: KbdState):
    """One `Event::MainEventsCleared` tick of the reference (src/main.rs:845-910), float32 like ultraviolet:
    acceleration 0.005 along the camera's local axes (forward follows the pitch), rotated by `Mat3::from_rotation_y(yaw)`,
    speed clamped to 0.2, eye += velocity, velocity *= 0.9; sun: acceleration 0.002, speed clamped to 0.05,
    yaw -= v.x, pitch = clamp(pitch + v.y, 0, pi/2), velocity *= 0.95."""
    acc, vmax = F(0.005), F(0.2)
    lv = np.zeros(3, F)
    cp, sp = F(math.cos(F(camera.pitch))), F(math.sin(F(camera.pitch)))
    if kbd.forward:
        lv[2] -= acc * cp
        lv[1] += acc * sp
    if kbd.back:
        lv[2] += acc * cp
        lv[1] -= acc * sp
    if kbd.left:
        lv[0] -= acc
    if kbd.right:
        lv[0] += acc
    ctl.camera_velocity = (ctl.camera_velocity + (mat_rotation_y(camera.yaw)[:3, :3] @ lv).astype(F)).astype(F)
    mag = F(np.sqrt(np.sum(ctl.camera_velocity * ctl.camera_velocity, dtype=F)))
    if mag > vmax:
        ctl.camera_velocity = (ctl.camera_velocity * (min(mag, vmax) / mag)).astype(F)
    camera.eye = tuple(float(v) for v in (np.asarray(camera.eye, F) + ctl.camera_velocity).astype(F))
    ctl.camera_velocity = (ctl.camera_velocity * F(0.9)).astype(F)

    acc, vmax = F(0.002), F(0.05)
    sv = ctl.sun_velocity.copy()
    if kbd.sun_up:
        sv[1] += acc
    if kbd.sun_down:
        sv[1] -= acc
    if kbd.sun_cw:
        sv[0] += acc
    if kbd.sun_ccw:
        sv[0] -= acc
    mag = F(np.sqrt(np.sum(sv * sv, dtype=F)))
    if mag > vmax:
        sv = (sv * (min(mag, vmax) / mag)).astype(F)
    sun.yaw = float(F(sun.yaw) - sv[0])
    sun.pitch = float(max(min(F(sun.pitch) + sv[1], F(math.pi / 2.0)), F(0.0)))
    ctl.sun_velocity = (sv * F(0.95)).astype(F)


def scripted_keys(tick: int) -> KbdState:
    """A fixed key schedule for headless animated runs (rt_demo --animate, tests): walk forward, strafe right while the sun
    turns clockwise and rises, then back off."""
    return KbdState(forward=tick < 20, right=10 <= tick < 30, sun_cw=20 <= tick < 40, sun_up=30 <= tick < 45, back=40 <= tick < 50,
                    left=50 <= tick < 55, sun_ccw=45 <= tick < 50, sun_down=50 <= tick < 60)


# ----------------------------------------------------------------------------- seeded streams
def hash_uniform(seed: int, stream: int, n: int) -> np.ndarray:
    """n float64 values in [0,1): splitmix64 over the counter (seed, stream, i)."""
    with np.errstate(over="ignore"):
        z = (np.arange(n, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = z + np.uint64(seed) * np.uint64(0xD1342543DE82EF95) + np.uint64(stream) * np.uint64(0xA24BAED4963EE407)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def instances_from_trs(pos, rot_y, scale, model_id, blas_handle, hit_shader) -> np.ndarray:
    """Vectorised translate * rotation_y * scale records (src/scene.rs:139-145)."""
    n = len(pos)
    rec = np.zeros(n, abi.INSTANCE_DTYPE)
    s = np.asarray(scale, F)
    c, sn = np.cos(np.asarray(rot_y, F)).astype(F), np.sin(np.asarray(rot_y, F)).astype(F)
    t = rec["transform"]
    t[:, 0], t[:, 2], t[:, 3] = c * s, sn * s, np.asarray(pos[:, 0], F)
    t[:, 5], t[:, 7] = s, np.asarray(pos[:, 1], F)
    t[:, 8], t[:, 10], t[:, 11] = -sn * s, c * s, np.asarray(pos[:, 2], F)
    rec["custom_index_and_mask"] = np.uint32((model_id & 0xFFFFFF) | (0xFF << 24))
    rec["sbt_offset_and_flags"] = np.asarray(hit_shader, np.uint32) & np.uint32(0xFFFFFF)
    rec["blas"] = np.uint64(blas_handle)
    return rec


# ----------------------------------------------------------------------------- scene assembly
@dataclass
class SceneSetup:
    """Everything one `rt_render` needs, after models and images went through a backend."""

    name: str
    instances: np.ndarray
    camera: Camera
    sun: Sun
    width: int
    height: int
    shadow_rays: int
    sun_radius: float
    max_segments: int = 3
    frame_index: int = 1
    models: dict = field(default_factory=dict)  # name -> (model_id, blas_handle, ModelArrays)
    dynamic: bool = False
    base_rot: Optional[np.ndarray] = None
    base_pos: Optional[np.ndarray] = None
    base_scale: Optional[np.ndarray] = None
    description: str = ""

    def uniforms(self, frame_index: Optional[int] = None, width=None, height=None) -> abi.RtUniforms:
        return make_uniforms(self.camera, self.sun, width or self.width, height or self.height, self.sun_radius,
                             self.frame_index if frame_index is None else frame_index)

    def params(self, width=None, height=None, **kw) -> abi.RtRenderParams:
        p = abi.RtRenderParams()
        p.width, p.height = width or self.width, height or self.height
        p.max_segments, p.shadow_rays = self.max_segments, self.shadow_rays
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def animate(self, tick: int) -> np.ndarray:
        """The instance records after `tick` updates.
        default: DefaultScene::update + write_resources (src/scene.rs:162-181): lain's rotation += 0.05 per tick, its
                 48-byte transform rewritten at instance 2, everything else untouched;
        C4:      every torus gets rotation_y += 0.05 per tick."""
        assert self.dynamic
        rec = self.instances.copy()
        if self.name == "default":
            rec["transform"][LAIN_INSTANCE] = lain_transform(tick)
            return rec
        moved = instances_from_trs(self.base_pos, self.base_rot + F(0.05) * F(tick), self.base_scale, 0, 0, 0)
        n = len(moved)
        rec["transform"][-n:] = moved["transform"]
        return rec


def push_builtin_images(backend):
    """Texture indices 0..3, src/main.rs:416-460."""
    i0 = backend.push_image(load_png_file_rgba8(os.path.join(ASSET_DIR, "green.png")), abi.RT_FORMAT_RGBA8_SRGB, False)
    i1 = backend.push_image(load_png_file_rgba8(os.path.join(ASSET_DIR, "pink.png")), abi.RT_FORMAT_RGBA8_SRGB, False)
    i2 = backend.push_image(load_png_file_rgba8(os.path.join(ASSET_DIR, "blue_noise_64x64.png")), abi.RT_FORMAT_RGBA8_UNORM, False)
    i3 = backend.push_image(load_png_file_rgba8(os.path.join(ASSET_DIR, "flipped_ggx_lut.png")), abi.RT_FORMAT_RGBA8_UNORM, True)
    assert (i0, i1, i2, i3) == (0, 1, 2, 3)


def load_model(backend, filename: str, fallback_image_index: int):
    with open(os.path.join(ASSET_DIR, filename), "rb") as f:
        data = f.read()
    arrays = load_gltf(data, filename, fallback_image_index, backend.push_image)
    model_id, handle = backend.create_model(arrays)
    return model_id, handle, arrays


def _mirror_field(seed, n, model_id, handle, lo=0.01, hi=0.1, extent=10.0, mirror_fraction=0.5):
    """Distribution of src/scene.rs:138-155 with a seeded stream."""
    pos = np.stack(
        [hash_uniform(seed, 0, n) * 2 * extent - extent, hash_uniform(seed, 1, n) * 2.0 + 0.5, hash_uniform(seed, 2, n) * 2 * extent - extent],
        axis=1,
    )
    rot = hash_uniform(seed, 3, n) * 100.0
    scale = hash_uniform(seed, 4, n) * (hi - lo) + lo
    kind = np.where(hash_uniform(seed, 5, n) < mirror_fraction, abi.RT_HIT_MIRROR, abi.RT_HIT_TEXTURED)
    return instances_from_trs(pos, rot, scale, model_id, handle, kind), pos, rot, scale


def build_scene(backend, config: str, width: Optional[int] = None, height: Optional[int] = None,
                num_instances: Optional[int] = None) -> SceneSetup:
    """Load images + models into `backend`, build its TLAS and return the render setup.
    `config` is one of c1..c5 (BASELINE.json configs[0..4]) or "default" (the reference's DefaultScene)."""
    config = config.lower()
    push_builtin_images(backend)
    models = {}

    def need(name, fallback):
        models[name] = load_model(backend, name + ".glb", fallback)
        return models[name][0], models[name][1]

    if config == "c1":
        # src/scene.rs:96-109: plane scale(10), torus translate(0,1,0); hard shadow.
        pid, ph = need("plane", 0)
        tid, th = need("tori", 1)
        inst = np.stack([
            make_instance(mat_scale(10.0), pid, ph, abi.RT_HIT_TEXTURED),
            make_instance(mat_translation(0, 1, 0), tid, th, abi.RT_HIT_TEXTURED),
        ])
        s = SceneSetup("c1", inst, Camera(), Sun(), width or 1280, height or 720, shadow_rays=1, sun_radius=0.0,
                       description="tori.glb on plane.glb, 1 primary + 1 hard-shadow ray/px")
    elif config == "c2":
        # LoadedModelScene recipe (identity, Textured) standing on the plane; 4 soft-shadow rays.
        pid, ph = need("plane", 0)
        lid, lh = need("lain", 1)
        inst = np.stack([
            make_instance(mat_scale(10.0), pid, ph, abi.RT_HIT_TEXTURED),
            make_instance(mat_identity(), lid, lh, abi.RT_HIT_TEXTURED),
        ])
        s = SceneSetup("c2", inst, Camera(eye=(0.0, 2.5, -7.0)), Sun(), width or 1920, height or 1080, shadow_rays=4,
                       sun_radius=0.05, description="lain.glb textured PBR, blue-noise soft shadows (4 shadow rays/px)")
    elif config == "c3":
        pid, ph = need("plane", 0)
        tid, th = need("tori", 1)
        fid, fh = need("fence", 0)
        n = num_instances or 100
        field_inst, _, _, _ = _mirror_field(0xC0FFEE, n, tid, th, mirror_fraction=1.0)
        big = np.stack([
            make_instance(mat_translation(-3.0, 1.25, 3.0) @ mat_rotation_y(0.6), tid, th, abi.RT_HIT_MIRROR),
            make_instance(mat_translation(4.5, 1.25, 5.0) @ mat_rotation_y(2.2), tid, th, abi.RT_HIT_MIRROR),
            make_instance(mat_translation(0.5, 1.25, 7.0) @ mat_rotation_y(1.3) @ mat_scale(1.5), tid, th, abi.RT_HIT_MIRROR),
        ])
        head = np.stack([
            make_instance(mat_scale(10.0), pid, ph, abi.RT_HIT_TEXTURED),
            make_instance(mat_translation(2, 0, 2), fid, fh, abi.RT_HIT_TEXTURED, True),
            make_instance(mat_translation(-1, 0, 1.0) @ mat_rotation_y(0.5), fid, fh, abi.RT_HIT_TEXTURED, True),
        ])
        inst = np.concatenate([head, big, field_inst])
        s = SceneSetup("c3", inst, Camera(), Sun(), width or 1920, height or 1080, shadow_rays=2, sun_radius=0.05,
                       description="fence.glb alpha-clip any-hit + mirror tori (2 bounces)")
    elif config == "c4":
        pid, ph = need("plane", 0)
        tid, th = need("tori", 1)
        n = num_instances or 10000
        side = int(math.ceil(math.sqrt(n)))
        gi = np.arange(n)
        cell = 100.0 / side
        px = (gi % side + 0.5) * cell - 50.0 + (hash_uniform(4, 0, n) - 0.5) * cell * 0.5
        pz = (gi // side + 0.5) * cell - 50.0 + (hash_uniform(4, 1, n) - 0.5) * cell * 0.5
        py = hash_uniform(4, 2, n) * 1.5 + 0.6
        pos = np.stack([px, py, pz], axis=1)
        rot = (hash_uniform(4, 3, n) * 100.0).astype(F)
        scale = (hash_uniform(4, 4, n) * 0.2 + 0.15).astype(F)
        kind = np.where(gi % 2 == 0, abi.RT_HIT_TEXTURED, abi.RT_HIT_MIRROR)
        tori = instances_from_trs(pos, rot, scale, tid, th, kind)
        inst = np.concatenate([np.stack([make_instance(mat_scale(60.0), pid, ph, abi.RT_HIT_TEXTURED)]), tori])
        s = SceneSetup("c4", inst, Camera(eye=(0.0, 14.0, -62.0), pitch=-0.28), Sun(), width or 3840, height or 2160,
                       shadow_rays=2, sun_radius=0.05, dynamic=True, base_rot=rot, base_pos=pos, base_scale=scale,
                       description="10k instanced tori, all transforms updated every frame (TLAS rebuild/refit)")
    elif config == "c5":
        pid, ph = need("plane", 0)
        tid, th = need("tori", 1)
        n = num_instances or 1000000
        ext = 60.0
        pos = np.stack([hash_uniform(5, 0, n) * 2 * ext - ext, hash_uniform(5, 1, n) * 6.0 + 0.2, hash_uniform(5, 2, n) * 2 * ext - ext], axis=1)
        rot = (hash_uniform(5, 3, n) * 100.0).astype(F)
        scale = (hash_uniform(5, 4, n) * 0.09 + 0.01).astype(F)
        tori = instances_from_trs(pos, rot, scale, tid, th, abi.RT_HIT_TEXTURED)
        inst = np.concatenate([np.stack([make_instance(mat_scale(80.0), pid, ph, abi.RT_HIT_TEXTURED)]), tori])
        s = SceneSetup("c5", inst, Camera(eye=(0.0, 16.0, -70.0), pitch=-0.3), Sun(), width or 3840, height or 2160,
                       shadow_rays=16, sun_radius=0.05,
                       description="1M-instance synthetic scene, 16 soft-shadow rays/px")
    elif config == "default":
        # src/scene.rs:35-159 with the random tori drawn from the seeded stream.
        pid, ph = need("plane", 0)
        tid, th = need("tori", 1)
        lid, lh = need("lain", 1)
        fid, fh = need("fence", 0)
        lain_base = mat_translation(-2.0, 0.0, -1.0) @ mat_scale(0.5)
        head = np.stack([
            make_instance(mat_scale(10.0), pid, ph, abi.RT_HIT_TEXTURED),
            make_instance(mat_translation(0, 1, 0), tid, th, abi.RT_HIT_TEXTURED),
            make_instance(lain_base @ mat_rotation_y(math.radians(150.0)), lid, lh, abi.RT_HIT_TEXTURED),
            make_instance(mat_translation(0, 1, 0), pid, ph, abi.RT_HIT_PORTAL, True),
            make_instance(mat_translation(2, 0, 2), fid, fh, abi.RT_HIT_TEXTURED, True),
        ])
        field_inst, _, _, _ = _mirror_field(0xD5CE, num_instances or 100, tid, th)
        inst = np.concatenate([head, field_inst])
        s = SceneSetup("default", inst, Camera(), Sun(), width or 1280, height or 720, shadow_rays=2, sun_radius=0.05, dynamic=True,
                       description="reference DefaultScene (seeded tori), lain rotating: one instance record + TLAS update per frame")
    else:
        raise ValueError(f"unknown config {config!r}")
    s.models = models
    backend.build_tlas(s.instances)
    return s


LAIN_INSTANCE = 2  # `lain_instance_offset` of DefaultScene (src/scene.rs:115-120): the record the per-frame update rewrites


def lain_transform(tick: int) -> np.ndarray:
    """DefaultScene::write_resources: the 48-byte transform written at instance 2 each frame (src/scene.rs:173-181)."""
    lain_base = mat_translation(-2.0, 0.0, -1.0) @ mat_scale(0.5)
    return transpose_matrix_for_instance(lain_base @ mat_rotation_y(F(math.radians(150.0)) + F(0.05) * F(tick)))

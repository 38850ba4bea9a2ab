"""The C++ host side (ray_tracing_gallery_b200/host_cpp) against the Python host: same GLB -> same arrays, same images in
the same order with the same formats and samplers; same scene recipes -> same instance records and uniforms.
CPU only: `host_dump` runs the C++ loader / recipes against a recording backend."""
import os
import struct
import subprocess

import numpy as np
import pytest

from ray_tracing_gallery_b200 import abi
from ray_tracing_gallery_b200.gltf import load_gltf
from ray_tracing_gallery_b200.scene import ASSET_DIR, build_scene

HOST_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ray_tracing_gallery_b200", "host_cpp")


@pytest.fixture(scope="module")
def host_dump():
    subprocess.check_call(["make", "-C", HOST_DIR, "host_dump", "rt_demo"], stdout=subprocess.DEVNULL)
    return os.path.join(HOST_DIR, "host_dump")


def read_sections(path):
    out, data, p = {}, open(path, "rb").read(), 0
    while p < len(data):
        tag, n = struct.unpack_from("<IQ", data, p)
        out[tag] = data[p + 12 : p + 12 + n]
        p += 12 + n
    return out


class ImageLog:
    def __init__(self, start):
        self.images, self.start = [], start

    def __call__(self, texels, fmt, linear):
        self.images.append((np.ascontiguousarray(texels), fmt, bool(linear)))
        return self.start + len(self.images) - 1


def compare_model(host_dump, tmp_path, glb_path, name, fallback):
    out = str(tmp_path / (name + ".bin"))
    subprocess.check_call([host_dump, "model", glb_path, str(fallback), "4", out])
    sec = read_sections(out)
    log = ImageLog(4)
    with open(glb_path, "rb") as f:
        m = load_gltf(f.read(), name, fallback, log)
    assert sec[1] == m.positions.tobytes() and sec[2] == m.normals.tobytes() and sec[3] == m.uvs.tobytes()
    for g, geo in enumerate(m.geometries):
        hdr = struct.unpack("<4i", sec[100 + g])
        assert hdr == (int(geo.opaque), geo.diffuse_image_index, geo.metallic_roughness_image_index, geo.normal_map_image_index)
        assert sec[200 + g] == geo.indices.tobytes()
    assert 100 + len(m.geometries) not in sec
    for i, (texels, fmt, linear) in enumerate(log.images):
        w, h, f, lin = struct.unpack("<4I", sec[1000 + 4 + i])
        assert (h, w) == texels.shape[:2] and f == fmt and bool(lin) == linear, (name, i)
        assert sec[2000 + 4 + i] == texels.tobytes(), (name, i)  # PNG decode incl. the 16-bit narrowing, byte for byte
    assert 1000 + 4 + len(log.images) not in sec
    return m


@pytest.mark.parametrize("name,fallback", [("plane.glb", 0), ("tori.glb", 1), ("lain.glb", 1), ("fence.glb", 0)])
def test_cpp_loader_equals_python_loader_on_the_reference_assets(host_dump, tmp_path, name, fallback):
    compare_model(host_dump, tmp_path, os.path.join(ASSET_DIR, name), name, fallback)


def test_cpp_loader_on_the_synthetic_multi_material_model(host_dump, tmp_path):
    from synth_assets import bumpy_two_material_glb

    p = tmp_path / "bumpy.glb"
    p.write_bytes(bumpy_two_material_glb())
    m = compare_model(host_dump, tmp_path, str(p), "bumpy", 1)
    assert len(m.geometries) == 2 and m.geometries[0].normal_map_image_index >= 0


def test_cpp_png_decoder_on_the_builtin_images(host_dump, tmp_path):
    """The four built-ins (8-bit RGB, 16-bit grey, 8-bit RGBA) come out of build_scene's push_builtin_images."""
    out = str(tmp_path / "scene.bin")
    subprocess.check_call([host_dump, "scene", "c1", ASSET_DIR, out])
    sec = read_sections(out)
    from ray_tracing_gallery_b200.gltf import load_png_file_rgba8

    for i, (f, fmt, lin) in enumerate([("green.png", abi.RT_FORMAT_RGBA8_SRGB, 0), ("pink.png", abi.RT_FORMAT_RGBA8_SRGB, 0),
                                       ("blue_noise_64x64.png", abi.RT_FORMAT_RGBA8_UNORM, 0), ("flipped_ggx_lut.png", abi.RT_FORMAT_RGBA8_UNORM, 1)]):
        img = load_png_file_rgba8(os.path.join(ASSET_DIR, f))
        assert struct.unpack("<4I", sec[1000 + i]) == (img.shape[1], img.shape[0], fmt, lin)
        assert sec[2000 + i] == img.tobytes(), f


@pytest.mark.parametrize("mode,depth16", [("L", False), ("LA", False), ("RGB", False), ("RGBA", False), ("P", False), ("1", False), ("I;16", True)])
def test_cpp_png_decoder_colour_types(host_dump, tmp_path, mode, depth16):
    """Every PNG colour type the `image` crate would hand to `to_rgba8()`: the C++ decoder against the Python host's decoder."""
    from PIL import Image

    from ray_tracing_gallery_b200.gltf import decode_png_rgba8

    rng = np.random.default_rng(3)
    w, h = 37, 23  # odd sizes: sub-byte rows end mid-byte
    if mode == "I;16":
        img = Image.fromarray(rng.integers(0, 65535, (h, w), dtype=np.uint16))
    elif mode == "1":
        img = Image.fromarray((rng.integers(0, 2, (h, w)) * 255).astype(np.uint8)).convert("1")
    elif mode == "P":
        img = Image.fromarray(rng.integers(0, 255, (h, w, 3), dtype=np.uint8), "RGB").convert("P", palette=Image.ADAPTIVE, colors=17)
    else:
        ch = {"L": 1, "LA": 2, "RGB": 3, "RGBA": 4}[mode]
        a = rng.integers(0, 255, (h, w, ch), dtype=np.uint8)
        img = Image.fromarray(a[..., 0] if ch == 1 else a, mode)
    p = tmp_path / "t.png"
    img.save(p, format="PNG")
    out = str(tmp_path / "t.bin")
    subprocess.check_call([host_dump, "png", str(p), out])
    sec = read_sections(out)
    want = decode_png_rgba8(p.read_bytes())
    assert struct.unpack("<4I", sec[1000])[:2] == (w, h)
    assert sec[2000] == want.tobytes()


def test_cpp_png_decoder_rgb_colour_key_and_bad_depth(host_dump, tmp_path):
    """tRNS on colour type 2 = a colour key: that exact RGB is transparent (image's to_rgba8 gives alpha 0 — it matters
    for alpha-clipped materials).  An IHDR bit depth that is illegal for the colour type must be refused, not crash."""
    import zlib

    from PIL import Image

    from ray_tracing_gallery_b200.gltf import decode_png_rgba8

    a = np.zeros((5, 7, 3), np.uint8)
    a[..., 0] = np.arange(7)[None, :] * 30
    a[2, 3] = (9, 8, 7)
    a[4, 6] = (9, 8, 7)
    p = tmp_path / "key.png"
    Image.fromarray(a, "RGB").save(p, format="PNG", transparency=(9, 8, 7))
    out = str(tmp_path / "key.bin")
    subprocess.check_call([host_dump, "png", str(p), out])
    want = decode_png_rgba8(p.read_bytes())
    assert want[2, 3, 3] == 0 and want[4, 6, 3] == 0 and want[0, 0, 3] == 255
    assert read_sections(out)[2000] == want.tobytes()

    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body))

    for depth, ctype in ((0, 0), (3, 0), (32, 2), (4, 6)):
        bad = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, depth, ctype, 0, 0, 0)) + \
            chunk(b"IDAT", zlib.compress(b"\0" * 64)) + chunk(b"IEND", b"")
        q = tmp_path / f"bad_{depth}_{ctype}.png"
        q.write_bytes(bad)
        r = subprocess.run([host_dump, "png", str(q), str(tmp_path / "bad.bin")], capture_output=True)
        assert r.returncode not in (0, -8, -11), (depth, ctype, r.returncode)  # an error exit, not SIGFPE / SIGSEGV


def test_control_integration_matches_the_python_host_and_float64(host_dump, tmp_path):
    """`Event::MainEventsCleared` of the reference (src/main.rs:845-910): camera / sun velocity integration over the fixed key
    schedule — the C++ host against the Python host tick for tick, and the Python host against a float64 restatement."""
    import math

    from ray_tracing_gallery_b200.scene import Camera, Controls, Sun, integrate_controls, scripted_keys

    out = str(tmp_path / "ctl.bin")
    ticks = 70
    subprocess.check_call([host_dump, "controls", str(ticks), out])
    cpp = np.frombuffer(read_sections(out)[6000], np.float32).reshape(ticks, 7)
    cam, sun, ctl = Camera(), Sun(), Controls()
    eye, pitch, yaw, sp, sy = np.array([0.0, 2.0, -5.0]), 0.0, math.pi, 0.5, 1.0   # float64 twin
    cv, sv = np.zeros(3), np.zeros(2)
    for t in range(ticks):
        k = scripted_keys(t)
        integrate_controls(cam, sun, ctl, k)
        got = np.array([*cam.eye, cam.pitch, cam.yaw, sun.pitch, sun.yaw])
        assert np.allclose(cpp[t], got, rtol=2e-6, atol=2e-6), (t, cpp[t], got)
        lv = np.zeros(3)
        if k.forward: lv += (0, 0.005 * math.sin(pitch), -0.005 * math.cos(pitch))
        if k.back: lv += (0, -0.005 * math.sin(pitch), 0.005 * math.cos(pitch))
        if k.left: lv[0] -= 0.005
        if k.right: lv[0] += 0.005
        s, c = math.sin(yaw), math.cos(yaw)
        cv += np.array([c * lv[0] + s * lv[2], lv[1], -s * lv[0] + c * lv[2]])
        m = np.linalg.norm(cv)
        if m > 0.2: cv *= 0.2 / m
        eye = eye + cv
        cv *= 0.9
        sv += np.array([(0.002 if k.sun_cw else 0) - (0.002 if k.sun_ccw else 0), (0.002 if k.sun_up else 0) - (0.002 if k.sun_down else 0)])
        m = np.linalg.norm(sv)
        if m > 0.05: sv *= 0.05 / m
        sy -= sv[0]
        sp = max(min(sp + sv[1], math.pi / 2), 0.0)
        sv *= 0.95
        assert np.allclose(got, [*eye, pitch, yaw, sp, sy], rtol=1e-4, atol=1e-4), t
    assert abs(cam.eye[2] + 5.0) > 0.5 and abs(sun.yaw - 1.0) > 0.2  # the schedule really moved camera and sun


class Recorder:
    """The Python twin of host_dump's recording backend."""

    def __init__(self):
        self.n = self.m = 0

    def push_image(self, *a):
        self.n += 1
        return self.n - 1

    def create_model(self, m):
        self.m += 1
        return self.m - 1, 1000 + self.m - 1

    def build_tlas(self, inst):
        self.inst = inst


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "default"])
def test_cpp_scene_recipes_equal_python_recipes(host_dump, tmp_path, cfg):
    out = str(tmp_path / "scene.bin")
    subprocess.check_call([host_dump, "scene", cfg, ASSET_DIR, out])
    sec = read_sections(out)
    s = build_scene(Recorder(), cfg)
    got = np.frombuffer(sec[5000], abi.INSTANCE_DTYPE)
    assert len(got) == len(s.instances)
    for k in ("custom_index_and_mask", "sbt_offset_and_flags", "blas"):
        assert np.array_equal(got[k], s.instances[k]), k
    # transforms: float32 on both sides; numpy's vectorised cos/sin and matmul may differ from libm in the last bits
    assert np.allclose(got["transform"], s.instances["transform"], rtol=2e-6, atol=2e-6)
    u = abi.RtUniforms.from_buffer_copy(sec[5001])
    w = s.uniforms()
    for field in ("view_inverse", "proj_inverse", "sun_dir"):
        assert np.allclose(list(getattr(u, field)), list(getattr(w, field)), rtol=1e-6, atol=1e-6), field
    assert (u.sun_radius, u.blue_noise_texture_index, u.ggx_lut_texture_index, u.frame_index) == (w.sun_radius, 2, 3, 1)
    assert struct.unpack("<4I", sec[5002]) == (s.width, s.height, s.shadow_rays, s.max_segments)


# ----------------------------------------------------------------------------------------------- on the GPU
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ray_tracing_gallery_b200", "csrc", "libb200rt.so")


@pytest.fixture(scope="module")
def host_parity(host_dump):
    exe = os.path.join(ROOT, "tests", "cpp", "host_parity")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "host_parity.cpp"), "-ldl", "-lz"])
    from oracle import binding

    return exe, binding.build()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,size", [("c1", (640, 360)), ("c3", (640, 360)), ("default", (640, 360)), ("c2", (960, 540))])
def test_cpp_host_parity_against_the_oracle(host_parity, cfg, size):
    """The parity bars, with both the CUDA path and the oracle driven from the C++ host through the C ABI."""
    import json

    exe, liborc = host_parity
    r = subprocess.run([exe, LIB, liborc, ASSET_DIR, cfg, str(size[0]), str(size[1])], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["ray_counts_equal"] and out["hit_id_agreement"] >= 0.9999


@pytest.mark.gpu
def test_rt_demo_frame_loop(host_dump, tmp_path):
    """rt_demo: the reference's headless frame loop from C++ (two frames in flight); its last frame equals the frame the
    Python host renders for the same frame index (scene floats agree to the last bits, so allow a handful of edge pixels)."""
    import json

    from conftest import make_renderer

    ppm = str(tmp_path / "c1.ppm")
    r = subprocess.run([os.path.join(HOST_DIR, "rt_demo"), "--config", "c1", "--width", "640", "--height", "360", "--frames", "6", "--out", ppm],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["frames"] == 6 and line["rays"] > 6 * 640 * 360 and line["mrays_per_s"] > 0
    data = open(ppm, "rb").read()
    assert data.startswith(b"P6\n640 360\n255\n")
    img = np.frombuffer(data[len(b"P6\n640 360\n255\n"):], np.uint8).reshape(360, 640, 3)
    gpu = make_renderer()
    s = build_scene(gpu, "c1", 640, 360)
    want = gpu.render(s.uniforms(frame_index=6), s.params(), want=("rgba8",))["rgba8"][..., :3]
    gpu.close()
    assert np.mean(np.abs(img.astype(int) - want.astype(int)).max(axis=2) <= 1) >= 0.999


@pytest.mark.gpu
def test_rt_demo_animated_run_matches_the_python_host(host_dump, tmp_path):
    """rt_demo --animate: the reference's whole per-tick loop from C++ (src/main.rs:845-948, src/scene.rs:162-204): control
    integration from the key schedule, lain's one-record update + in-place TLAS refit, frame_index + 1, two frames in
    flight.  Its last frame must equal the frame the Python host renders after the same ticks."""
    import json

    from conftest import make_renderer
    from ray_tracing_gallery_b200 import abi
    from ray_tracing_gallery_b200.scene import LAIN_INSTANCE, Controls, integrate_controls, scripted_keys

    frames = 24
    ppm = str(tmp_path / "anim.ppm")
    r = subprocess.run([os.path.join(HOST_DIR, "rt_demo"), "--config", "default", "--width", "640", "--height", "360", "--frames", str(frames), "--animate",
                        "--out", ppm], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["animated"] and line["frames"] == frames
    data = open(ppm, "rb").read()
    img = np.frombuffer(data[len(b"P6\n640 360\n255\n"):], np.uint8).reshape(360, 640, 3)
    gpu = make_renderer()
    s = build_scene(gpu, "default", 640, 360)
    ctl = Controls()
    for k in range(frames):
        integrate_controls(s.camera, s.sun, ctl, scripted_keys(k))
    assert np.allclose(line["eye"], s.camera.eye, atol=2e-4) and np.allclose(line["sun"], [s.sun.pitch, s.sun.yaw], atol=2e-4)
    gpu.update_instances(LAIN_INSTANCE, s.animate(frames)[LAIN_INSTANCE:LAIN_INSTANCE + 1])
    gpu.update_tlas(abi.RT_UPDATE_REFIT)
    want = gpu.render(s.uniforms(frame_index=frames), s.params(), want=("rgba8",))["rgba8"][..., :3]
    gpu.close()
    assert np.mean(np.abs(img.astype(int) - want.astype(int)).max(axis=2) <= 1) >= 0.995

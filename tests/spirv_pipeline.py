"""The reference's ray-tracing pipeline assembled from its own shipped SPIR-V stages, run on the CPU by
tests/spirv_interp.py.  TEST INFRASTRUCTURE (never imported by the product).

What runs as compiled reference code (the .spv binaries under /root/reference/shaders, read in place):
    ray_generation, primary_ray_miss, shadow_ray_miss, closest_hit_portal      (rust-gpu, shaders/ray-tracing/src/lib.rs)
    closest_hit_textured, closest_hit_mirror, any_hit_alpha_clip               (glslang, shaders/*.glsl)
What the Vulkan driver / RT hardware does for the reference and is therefore a callback here:
    * OpTraceRayKHR: the intersection itself comes from `tracer(origin, dir, tmin, tmax, any)` (the tests pass the
      oracle's two-level trace, which has its own float64 brute-force pins), then the hit group is chosen the way
      src/main.rs:289-305 / :360-384 lays out the SBT and the stage's module is run with the built-ins of that hit;
    * OpImageSampleExplicitLod: Vulkan's LOD-0 sampling rules in float64 (`sample_image`) — nearest / bilinear,
      REPEAT addressing, sRGB decode before filtering — over the images in push order;
    * PhysicalStorageBuffer loads: byte regions holding the reference-layout tables (ModelInfo 32 B, GeometryInfo 24 B,
      Uniforms 176 B, vertex streams, indices), built from the loader's output exactly as src/util_structs.rs:1158-1236
      uploads them.
"""
import os

import numpy as np

from ray_tracing_gallery_b200 import abi
from spirv_interp import (F32, I32, SC_HIT_ATTRIBUTE, SC_INCOMING_RAY_PAYLOAD, SC_PUSH_CONSTANT, SC_RAY_PAYLOAD, SC_UNIFORM,
                          SC_UNIFORM_CONSTANT, U32, IgnoreIntersection, Invocation, Memory, Module, SpirvError)

SHADER_DIR = "/root/reference/shaders"
STAGES = ["ray_generation", "primary_ray_miss", "shadow_ray_miss", "closest_hit_portal", "closest_hit_textured", "closest_hit_mirror",
          "any_hit_alpha_clip"]


def shaders_available():
    return all(os.path.exists(os.path.join(SHADER_DIR, s + ".spv")) for s in STAGES)


class RecordingBackend:
    """Passes the scene through to `inner` (the oracle) and keeps what a Vulkan host would have uploaded."""

    def __init__(self, inner):
        self.inner = inner
        self.images = []   # (texels, format, linear)
        self.models = []   # ModelArrays
        self.instances = None

    def push_image(self, texels, fmt, linear):
        self.images.append((np.array(texels), fmt, bool(linear)))
        return self.inner.push_image(texels, fmt, linear)

    def create_model(self, arrays):
        self.models.append(arrays)
        return self.inner.create_model(arrays)

    def build_tlas(self, instances):
        self.instances = np.array(instances)
        return self.inner.build_tlas(instances)


def srgb_eotf(c):
    c = np.asarray(c, np.float64)
    return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)


def decode_texel(texels, fmt, x, y):
    t = texels[y, x]
    if fmt == abi.RT_FORMAT_RGBA32_SFLOAT:
        return t.astype(np.float64)
    v = t.astype(np.float64) / 255.0
    if fmt == abi.RT_FORMAT_RGBA8_SRGB:
        v[:3] = srgb_eotf(v[:3])
    return v


def sample_image(image, u, v):
    """Vulkan texel addressing at LOD 0 (spec 'Texel Coordinate Systems'): normalised coords, REPEAT, texel centres at +0.5."""
    texels, fmt, linear = image
    h, w = texels.shape[:2]
    u, v = float(u), float(v)
    if not linear:
        x, y = int(np.floor(u * w)) % w, int(np.floor(v * h)) % h
        return decode_texel(texels, fmt, x, y)
    fx, fy = u * w - 0.5, v * h - 0.5
    x0, y0 = int(np.floor(fx)), int(np.floor(fy))
    ax, ay = fx - x0, fy - y0
    out = np.zeros(4)
    for dx, dy, wt in ((0, 0, (1 - ax) * (1 - ay)), (1, 0, ax * (1 - ay)), (0, 1, (1 - ax) * ay), (1, 1, ax * ay)):
        out += wt * decode_texel(texels, fmt, (x0 + dx) % w, (y0 + dy) % h)
    return out


class _Textures:
    def __getitem__(self, i):
        return ("tex", int(i))


class RefPipeline:
    """`cmd_trace_rays` over the shipped stages.  `tracer(o, d, tmin, tmax, any) -> (hit, (instance_id, geom, prim), (t, u, v))`."""

    def __init__(self, rec: RecordingBackend, tracer, anyhit_probe=None):
        self.mod = {s: Module(open(os.path.join(SHADER_DIR, s + ".spv"), "rb").read()) for s in STAGES}
        self.rec, self.tracer = rec, tracer
        self.mem = Memory()
        self.instances = rec.instances
        # ---- ModelInfo[] / GeometryInfo[] tables with device addresses inside (shared-structs/src/lib.rs:24-47)
        mi = np.zeros(len(rec.models), abi.MODEL_INFO_DTYPE) if hasattr(abi, "MODEL_INFO_DTYPE") else None
        table = bytearray()
        for m in rec.models:
            pos = self.mem.add(np.ascontiguousarray(m.positions, np.float32))
            nrm = self.mem.add(np.ascontiguousarray(m.normals, np.float32))
            uvs = self.mem.add(np.ascontiguousarray(m.uvs, np.float32))
            ginfo = bytearray()
            for g in m.geometries:
                idx = self.mem.add(np.ascontiguousarray(g.indices, np.uint32))
                ginfo += np.array([idx], np.uint64).tobytes()
                ginfo += np.array([g.diffuse_image_index, g.metallic_roughness_image_index], np.uint32).tobytes()
                ginfo += np.array([g.normal_map_image_index], np.int32).tobytes() + b"\0\0\0\0"
            gaddr = self.mem.add(np.frombuffer(bytes(ginfo) or b"\0" * 24, np.uint8))
            table += np.array([pos, nrm, uvs, gaddr], np.uint64).tobytes()
        self.model_info_addr = self.mem.add(np.frombuffer(bytes(table), np.uint8))
        self.uniform_region = np.zeros(176, np.uint8)
        self.uniforms_addr = self.mem.add(self.uniform_region)
        self.uniform_region = self.mem.regions[-1][1]
        self.push = np.array([self.model_info_addr, self.uniforms_addr, 0xACCE1], np.uint64)
        self.push_addr = self.mem.add(self.push)
        self.opaque = [[bool(g.opaque) for g in m.geometries] for m in rec.models]
        # per-instance transforms
        self.o2w, self.w2o = [], []
        for r in self.instances:
            m34 = np.asarray(r["transform"], np.float64).reshape(3, 4)
            m44 = np.vstack([m34, [0, 0, 0, 1]])
            inv = np.linalg.inv(m44)[:3]
            self.o2w.append(m34.astype(np.float32))
            self.w2o.append(inv.astype(np.float32))
        self.rays = [0, 0]          # trace calls: ray-gen segments, shadow rays
        self.anyhit_calls = 0
        self.image = {}
        self.log = []               # per trace call of the current pixel: dict(kind, hit, payload ...)
        self.clock = []             # values the next OpReadClockKHR executions return (show_heatmap frames)
        self.steps = 0

    # ------------------------------------------------------------------ environment for one invocation
    class Env:
        def __init__(self, pipe, builtins):
            self.pipe, self.b = pipe, builtins
            self.memory = pipe.mem

        def builtin(self, name, t):
            v = self.b[name]
            if t.kind == "vector":
                return np.asarray(v).astype(Module.np_type(t.elem))
            if t.kind == "matrix":
                return [np.asarray(c, np.float32) for c in v]
            return Module.np_type(t)(v)

        def bind_global(self, m, vid, pt, sc):
            p = self.pipe
            if sc == SC_PUSH_CONSTANT:
                return [m.load_phys(p.mem, p.push_addr, pt.elem)]
            if sc == SC_UNIFORM:
                return [m.load_phys(p.mem, p.uniforms_addr, pt.elem)]
            if sc == SC_UNIFORM_CONSTANT:
                if pt.elem.kind == "array":
                    return [_Textures()]
                return [("image", 0)]
            raise SpirvError(f"unbound global {vid} in storage class {sc}")

        def sample(self, handle, coord, lod):
            if float(lod) != 0.0:
                raise SpirvError("the shipped stages only sample LOD 0")
            kind, index = handle
            images = self.pipe.rec.images
            if index >= len(images):
                return np.zeros(4, np.float32)  # null descriptor (robustness2, src/main.rs:183-184)
            return sample_image(images[index], coord[0], coord[1]).astype(np.float32)

        def image_write(self, image, coord, texel):
            self.pipe.image[(int(coord[0]), int(coord[1]))] = np.asarray(texel, np.float32).copy()

        def read_clock(self):
            return self.pipe.clock.pop(0) if self.pipe.clock else 0

        def trace_ray(self, accel, flags, cull, sbt_offset, sbt_stride, miss_index, origin, tmin, direction, tmax, payload_ptr, inv):
            self.pipe.trace(self, flags, sbt_offset, miss_index, origin, tmin, direction, tmax, payload_ptr, inv)

    def _run(self, stage, builtins, cells=None, payload=None, attribs=None):
        m = self.mod[stage]
        env = RefPipeline.Env(self, builtins)
        inv = Invocation(m, env)
        for vid, (pt, sc, _) in m.globals.items():
            if sc == SC_INCOMING_RAY_PAYLOAD and payload is not None:
                inv.bind(vid, payload)
            elif sc == SC_HIT_ATTRIBUTE and attribs is not None:
                inv.bind(vid, [np.asarray(attribs, np.float32)])
        inv.run()
        self.steps += inv.steps
        return inv

    def hit_builtins(self, launch, origin, direction, tmin, t, ids):
        inst, geom, prim = ids
        r = self.instances[inst]
        o2w, w2o = self.o2w[inst], self.w2o[inst]
        return {
            "LaunchId": (launch[0], launch[1], 0), "LaunchSize": (self.size[0], self.size[1], 1),
            "WorldRayOrigin": origin, "WorldRayDirection": direction, "RayTmin": tmin, "RayTmax": t, "HitT": t,
            "InstanceCustomIndex": int(r["custom_index_and_mask"]) & 0xFFFFFF, "InstanceId": inst, "PrimitiveId": prim,
            "RayGeometryIndex": geom,
            "ObjectToWorld": [o2w[:, c] for c in range(4)],   # mat4x3: four columns of vec3
            "WorldToObject": [w2o[:, c] for c in range(4)],
            "ObjectRayOrigin": w2o[:, :3] @ np.asarray(origin, np.float32) + w2o[:, 3],
            "ObjectRayDirection": w2o[:, :3] @ np.asarray(direction, np.float32),
            "IncomingRayFlags": 0, "HitKind": 0xFE,
        }

    # ------------------------------------------------------------------ OpTraceRayKHR
    def trace(self, env, flags, sbt_offset, miss_index, origin, tmin, direction, tmax, payload_ptr, inv):
        launch = env.b["LaunchId"]
        terminate_first, skip_closest = bool(flags & 4), bool(flags & 8)
        o, d = np.asarray(origin, np.float32), np.asarray(direction, np.float32)
        self.rays[1 if terminate_first else 0] += 1
        hit, ids, tuv = self.tracer(o, d, float(tmin), float(tmax), terminate_first)
        payload_cell = [inv._load(payload_ptr)]
        entry = {"shadow": terminate_first, "origin": o.copy(), "dir": d.copy(), "hit": bool(hit), "ids": tuple(int(x) for x in ids) if hit else None,
                 "tuv": np.array(tuv, np.float32) if hit else None}
        if not hit:
            stage = "shadow_ray_miss" if miss_index == 1 else "primary_ray_miss"
            b = {"LaunchId": launch, "LaunchSize": env.b["LaunchSize"], "WorldRayOrigin": o, "WorldRayDirection": d, "RayTmin": tmin, "RayTmax": tmax}
            if stage == "shadow_ray_miss":
                # the GLSL caller's payload is {uint8_t}, the rust-gpu miss stage's is {u8}: same single member
                cell = [[np.uint8(payload_cell[0][0])]]
                self._run(stage, b, payload=cell)
                payload_cell[0][0] = type(payload_cell[0][0])(cell[0][0])
            else:
                self._run(stage, b, payload=payload_cell)
        elif not skip_closest:
            inst = ids[0]
            group = (int(self.instances[inst]["sbt_offset_and_flags"]) & 0xFFFFFF) + sbt_offset
            stage = {0: "closest_hit_textured", 1: "closest_hit_mirror", 2: "closest_hit_portal"}.get(group)
            if stage is not None:
                b = self.hit_builtins(launch, o, d, tmin, float(tuv[0]), ids)
                self._run(stage, b, payload=payload_cell, attribs=tuv[1:3])
        entry["payload"] = Module.copy(payload_cell[0])
        self.log.append(entry)
        inv._store(payload_ptr, payload_cell[0])

    # ------------------------------------------------------------------ any-hit on one candidate
    def any_hit_ignores(self, launch, origin, direction, tmin, ids, tuv):
        """Run any_hit_alpha_clip on a candidate: True when the stage calls ignoreIntersectionEXT."""
        b = self.hit_builtins(launch, np.asarray(origin, np.float32), np.asarray(direction, np.float32), tmin, float(tuv[0]), ids)
        self.anyhit_calls += 1
        try:
            self._run("any_hit_alpha_clip", b, payload=None, attribs=tuv[1:3])
        except IgnoreIntersection:
            return True
        return False

    # ------------------------------------------------------------------ cmd_trace_rays for one pixel
    def set_uniforms(self, uniforms: abi.RtUniforms, width, height):
        self.uniform_region[:] = np.frombuffer(bytes(uniforms), np.uint8)
        self.size = (width, height)

    def pixel(self, x, y):
        """Returns (vec4 written to the image, log of the trace calls of this ray-gen invocation)."""
        self.log = []
        b = {"LaunchId": (x, y, 0), "LaunchSize": (self.size[0], self.size[1], 1)}
        self._run("ray_generation", b)
        return self.image[(x, y)], self.log

"""Thin object wrapper over the C ABI of include/b200rt.h.

`CApiBackend` drives any shared library that exports the boundary's entry
points under a prefix: `rt_` for the product (libb200rt.so, see native.py) and
`orc_` for the CPU oracle (test infrastructure, bound in oracle/binding.py).
Keeping one driver for both is what lets the parity tests feed the two sides
the exact same bytes.
"""
import ctypes as C

import numpy as np

from . import abi
from .gltf import ModelArrays


class RtError(RuntimeError):
    pass


class CApiBackend:
    prefix = "rt_"

    def __init__(self, lib, ctx):
        self.lib = lib
        self.ctx = ctx
        self._keep = []

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _check(self, rc, what):
        if rc != 0:
            msg = self._fn("last_error")(self.ctx)
            raise RtError(f"{self.prefix}{what} failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if self.ctx is not None:
            self._fn("destroy")(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- ImageManager::push_image (src/util_structs.rs:1330-1349)
    def push_image(self, texels: np.ndarray, fmt: int, linear: bool) -> int:
        want = np.float32 if fmt == abi.RT_FORMAT_RGBA32_SFLOAT else np.uint8
        t = np.ascontiguousarray(texels, dtype=want)
        if t.ndim != 3 or t.shape[2] != 4:
            raise ValueError("texels must be (h, w, 4)")
        idx = C.c_uint32()
        rc = self._fn("push_image")(self.ctx, t.ctypes.data, t.shape[1], t.shape[0], fmt, int(bool(linear)), C.byref(idx))
        self._check(rc, "push_image")
        return idx.value

    # ---- Model::new (src/util_structs.rs:1158-1236)
    def create_model(self, m: ModelArrays):
        pos = np.ascontiguousarray(m.positions, np.float32)
        nrm = np.ascontiguousarray(m.normals, np.float32)
        uvs = np.ascontiguousarray(m.uvs, np.float32)
        geoms = (abi.RtGeometryDesc * len(m.geometries))()
        keep = []
        for gd, g in zip(geoms, m.geometries):
            idx = np.ascontiguousarray(g.indices, np.uint32)
            keep.append(idx)
            gd.indices = idx.ctypes.data_as(C.POINTER(C.c_uint32))
            gd.num_indices = len(idx)
            gd.opaque = 1 if g.opaque else 0
            gd.images = abi.RtGeometryImages(g.diffuse_image_index, g.metallic_roughness_image_index, g.normal_map_image_index, 0)
        desc = abi.RtModelDesc(
            pos.ctypes.data_as(C.POINTER(C.c_float)),
            nrm.ctypes.data_as(C.POINTER(C.c_float)),
            uvs.ctypes.data_as(C.POINTER(C.c_float)),
            len(pos),
            len(m.geometries),
            geoms,
        )
        mid, handle = C.c_uint32(), C.c_uint64()
        rc = self._fn("create_model")(self.ctx, C.byref(desc), C.byref(mid), C.byref(handle))
        self._check(rc, "create_model")
        return mid.value, handle.value

    # ---- build_tlas / instance writes / update_tlas
    @staticmethod
    def _records(instances) -> np.ndarray:
        a = np.ascontiguousarray(instances)
        if a.dtype != abi.INSTANCE_DTYPE:
            raise ValueError("instances must use abi.INSTANCE_DTYPE (64-byte records)")
        return a

    def build_tlas(self, instances):
        a = self._records(instances)
        self._check(self._fn("build_tlas")(self.ctx, a.ctypes.data, len(a)), "build_tlas")

    def update_instances(self, first: int, instances):
        a = self._records(instances)
        self._check(self._fn("update_instances")(self.ctx, first, len(a), a.ctypes.data), "update_instances")

    def update_tlas(self, mode: int = abi.RT_UPDATE_AUTO):
        self._check(self._fn("update_tlas")(self.ctx, mode), "update_tlas")

    # ---- cmd_trace_rays
    @staticmethod
    def rows_rendered(p: abi.RtRenderParams):
        tw, th = (p.width, p.height) if p.tile_w == 0 else (p.tile_w, p.tile_h)
        if p.strip_height and p.strip_count > 1:
            rows = sum(1 for r in range(th) if (r // p.strip_height) % p.strip_count == p.strip_index)
        else:
            rows = th
        return rows, tw

    def render(self, uniforms: abi.RtUniforms, params: abi.RtRenderParams, want=("rgba8", "radiance", "hit_ids", "ray_counts")):
        rows, tw = self.rows_rendered(params)
        out = abi.RtFrameOutputs()
        res = {}
        if "rgba8" in want:
            res["rgba8"] = np.zeros((rows, tw, 4), np.uint8)
            out.rgba8 = res["rgba8"].ctypes.data
        if "radiance" in want:
            res["radiance"] = np.zeros((rows, tw, 3), np.float32)
            out.radiance = res["radiance"].ctypes.data
        if "hit_ids" in want:
            res["hit_ids"] = np.zeros((rows, tw, params.max_segments, 3), np.uint32)
            out.hit_ids = res["hit_ids"].ctypes.data
        if "ray_counts" in want:
            res["ray_counts"] = np.zeros(2, np.uint64)
            out.ray_counts = res["ray_counts"].ctypes.data
        if "cost_cycles" in want:  # only written by show_heatmap frames
            res["cost_cycles"] = np.zeros((rows, tw), np.uint32)
            out.cost_cycles = res["cost_cycles"].ctypes.data
        rc = self._fn("render")(self.ctx, C.byref(uniforms), C.byref(params), C.byref(out))
        self._check(rc, "render")
        return res

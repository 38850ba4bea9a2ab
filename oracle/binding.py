"""ctypes binding of oracle/liborc.so — TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from ray_tracing_gallery_b200 import abi
from ray_tracing_gallery_b200.backend import CApiBackend

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "rt_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def load():
    global _LIB
    if _LIB is None:
        lib = C.CDLL(build())
        abi.declare_api(lib, "orc_")
        lib.orc_create.argtypes = [C.POINTER(C.c_void_p)]
        lib.orc_set_brute_force.argtypes = [C.c_void_p, C.c_int]
        lib.orc_set_threads.argtypes = [C.c_void_p, C.c_int]
        lib.orc_get_threads.argtypes = [C.c_void_p]
        fp = C.POINTER(C.c_float)
        lib.orc_brdf.argtypes = [fp, fp, fp, fp, C.c_float, C.c_float, C.c_float, fp]
        lib.orc_brdf.restype = None
        lib.orc_v_smith_ggx.argtypes = [fp, fp, fp, C.c_float]
        lib.orc_v_smith_ggx.restype = C.c_float
        lib.orc_heatmap_temperature.argtypes = [C.c_float, fp]
        lib.orc_heatmap_temperature.restype = None
        lib.orc_heatmap_pixel.argtypes = [C.c_uint32, C.c_float, fp, fp]
        lib.orc_heatmap_pixel.restype = None
        lib.orc_linear_to_srgb.argtypes = [C.c_float]
        lib.orc_linear_to_srgb.restype = C.c_float
        lib.orc_unorm8.argtypes = [C.c_float]
        lib.orc_unorm8.restype = C.c_uint8
        lib.orc_blue_noise_xi.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, fp]
        lib.orc_blue_noise_xi.restype = None
        lib.orc_sample_directional_light.argtypes = [fp, fp, C.c_float, fp]
        lib.orc_sample_directional_light.restype = None
        lib.orc_sample_texture.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, fp]
        lib.orc_sample_texture.restype = None
        lib.orc_intersect_triangle.argtypes = [fp, fp, fp, fp, fp, fp]
        lib.orc_invert_3x4.argtypes = [fp, fp]
        lib.orc_invert_3x4.restype = None
        lib.orc_primary_ray.argtypes = [C.POINTER(abi.RtUniforms), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, fp, fp]
        lib.orc_primary_ray.restype = None
        lib.orc_terminator_origin.argtypes = [fp, fp, fp, fp, fp]
        lib.orc_terminator_origin.restype = None
        lib.orc_anyhit_accepts.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_float]
        lib.orc_trace.argtypes = [C.c_void_p, fp, fp, C.c_float, C.c_float, C.c_int, C.POINTER(C.c_uint32), fp]
        _LIB = lib
    return _LIB


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def f32(*v):
    return np.ascontiguousarray(np.asarray(v, np.float32).reshape(-1))


class Oracle(CApiBackend):
    prefix = "orc_"

    def __init__(self, brute_force=False, threads=0):
        lib = load()
        ctx = C.c_void_p()
        lib.orc_create(C.byref(ctx))
        super().__init__(lib, ctx)
        if brute_force:
            lib.orc_set_brute_force(ctx, 1)
        if threads:
            lib.orc_set_threads(ctx, threads)

    def set_brute_force(self, on):
        self.lib.orc_set_brute_force(self.ctx, int(on))

    @property
    def threads(self):
        return self.lib.orc_get_threads(self.ctx)

    # ---- unit-level helpers
    def brdf(self, normal, view, light, base, rough, metallic, sun_factor):
        out = np.zeros(3, np.float32)
        self.lib.orc_brdf(_fp(f32(*normal)), _fp(f32(*view)), _fp(f32(*light)), _fp(f32(*base)), rough, metallic, sun_factor, _fp(out))
        return out

    def v_smith_ggx(self, normal, view, light, roughness):
        return self.lib.orc_v_smith_ggx(_fp(f32(*normal)), _fp(f32(*view)), _fp(f32(*light)), roughness)

    def heatmap_temperature(self, heat):
        out = np.zeros(3, np.float32)
        self.lib.orc_heatmap_temperature(heat, _fp(out))
        return out

    def heatmap_pixel(self, cycles, colour, scale=0.0):
        out = np.zeros(3, np.float32)
        self.lib.orc_heatmap_pixel(int(cycles), scale, _fp(f32(*colour)), _fp(out))
        return out

    def linear_to_srgb(self, c):
        return self.lib.orc_linear_to_srgb(float(c))

    def unorm8(self, c):
        return self.lib.orc_unorm8(float(c))

    def blue_noise_xi(self, px, py, iteration, frame, tex=2):
        out = np.zeros(2, np.float32)
        self.lib.orc_blue_noise_xi(self.ctx, tex, px, py, iteration, frame, _fp(out))
        return out

    def sample_directional_light(self, xi, center, radius):
        out = np.zeros(3, np.float32)
        self.lib.orc_sample_directional_light(_fp(f32(*xi)), _fp(f32(*center)), radius, _fp(out))
        return out

    def sample_texture(self, index, u, v):
        out = np.zeros(4, np.float32)
        self.lib.orc_sample_texture(self.ctx, index, u, v, _fp(out))
        return out

    def intersect_triangle(self, o, d, a, b, c):
        out = np.zeros(3, np.float32)
        hit = self.lib.orc_intersect_triangle(_fp(f32(*o)), _fp(f32(*d)), _fp(f32(*a)), _fp(f32(*b)), _fp(f32(*c)), _fp(out))
        return (hit != 0), out

    def invert_3x4(self, m):
        out = np.zeros(12, np.float32)
        self.lib.orc_invert_3x4(_fp(f32(*np.asarray(m).reshape(-1))), _fp(out))
        return out

    def primary_ray(self, uniforms, x, y, w, h):
        o, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
        self.lib.orc_primary_ray(C.byref(uniforms), x, y, w, h, _fp(o), _fp(d))
        return o, d

    def terminator_origin(self, pos9, nrm9, bary3, o2w12):
        out = np.zeros(3, np.float32)
        self.lib.orc_terminator_origin(_fp(f32(*pos9)), _fp(f32(*nrm9)), _fp(f32(*bary3)), _fp(f32(*o2w12)), _fp(out))
        return out

    def trace(self, o, d, tmin, tmax, any_hit=False):
        ids = (C.c_uint32 * 3)()
        tuv = np.zeros(3, np.float32)
        hit = self.lib.orc_trace(self.ctx, _fp(f32(*o)), _fp(f32(*d)), tmin, tmax, int(any_hit), ids, _fp(tuv))
        return (hit != 0), tuple(ids), tuv

    def anyhit_accepts(self, instance_id, geom, prim, u, v):
        """any_hit_alpha_clip on one candidate: True = kept, False = ignoreIntersectionEXT."""
        r = self.lib.orc_anyhit_accepts(self.ctx, instance_id, geom, prim, u, v)
        if r < 0:
            raise ValueError("no such candidate")
        return r == 1

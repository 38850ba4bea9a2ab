"""ctypes binding of libb200rt.so (include/b200rt.h) — the product path.

There is no CPU fallback here and nothing in this module touches oracle/: if the
shared library is missing or no CUDA device is present, `Renderer()` raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import abi
from .backend import CApiBackend, RtError

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
# B200RT_LIB: an A/B build of the same library (csrc/Makefile `variant`); still the CUDA product, never a fallback
LIB_PATH = os.environ.get("B200RT_LIB") or os.path.join(_CSRC, "libb200rt.so")
_LIB = None

# every symbol include/b200rt.h declares
EXPORTS = [
    "rt_create", "rt_destroy", "rt_last_error", "rt_set_stream", "rt_push_image", "rt_create_model", "rt_build_tlas",
    "rt_update_instances", "rt_update_instances_device", "rt_update_tlas", "rt_render", "rt_render_device", "rt_render_async",
    "rt_wait_frame", "rt_render_device_slot", "rt_host_alloc", "rt_host_free", "rt_readback",
    "rt_sync", "rt_get_stats", "rt_get_push_constants", "rt_debug_read_model_info", "rt_kernel_launches", "rt_version",
]


def build(force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into csrc/libb200rt.so (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        subprocess.check_call(["make", "-C", _CSRC, "-j4", "libb200rt.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RtError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    abi.declare_api(lib, "rt_")
    p, u32 = C.c_void_p, C.c_uint32
    lib.rt_create.argtypes = [C.c_int, C.POINTER(p)]
    lib.rt_set_stream.argtypes = [p, p]
    lib.rt_update_instances_device.argtypes = [p, u32, u32, p]
    lib.rt_render_device.argtypes = [p, C.POINTER(abi.RtUniforms), C.POINTER(abi.RtRenderParams), C.POINTER(abi.RtFrameOutputs)]
    lib.rt_render_async.argtypes = [p, C.POINTER(abi.RtUniforms), C.POINTER(abi.RtRenderParams), C.POINTER(abi.RtFrameOutputs), C.POINTER(u32)]
    lib.rt_wait_frame.argtypes = [p, u32]
    lib.rt_render_device_slot.argtypes = [p, u32, p, C.POINTER(abi.RtUniforms), C.POINTER(abi.RtRenderParams), C.POINTER(abi.RtFrameOutputs)]
    lib.rt_host_alloc.argtypes = [p, C.c_size_t, C.POINTER(p)]
    lib.rt_host_free.argtypes = [p, p]
    lib.rt_readback.argtypes = [p, p, C.c_size_t]
    lib.rt_sync.argtypes = [p]
    lib.rt_get_stats.argtypes = [p, C.POINTER(abi.RtStats)]
    lib.rt_get_push_constants.argtypes = [p, C.POINTER(abi.RtPushConstantBufferAddresses)]
    lib.rt_debug_read_model_info.argtypes = [p, u32, C.POINTER(abi.RtModelInfo), C.POINTER(abi.RtGeometryInfo), u32]
    lib.rt_version.restype = u32
    lib.rt_kernel_launches.restype = C.c_uint64
    _LIB = lib
    return lib


class Renderer(CApiBackend):
    """One context on one GPU (`rt_create`)."""

    prefix = "rt_"

    def __init__(self, device: int = 0):
        lib = load()
        ctx = C.c_void_p()
        rc = lib.rt_create(device, C.byref(ctx))
        if rc != 0:
            msg = lib.rt_last_error(None)
            raise RtError(f"rt_create({device}) failed ({rc}): {msg.decode() if msg else ''}")
        super().__init__(lib, ctx)
        self.device = device

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.rt_set_stream(self.ctx, cuda_stream), "set_stream")

    def update_instances_raw(self, first: int, count: int, host_ptr: int):
        """`rt_update_instances` from a raw host pointer (e.g. pinned memory) to `count` 64-byte records."""
        self._check(self.lib.rt_update_instances(self.ctx, first, count, host_ptr), "update_instances")

    def update_instances_device(self, first: int, count: int, device_ptr: int):
        self._check(self.lib.rt_update_instances_device(self.ctx, first, count, device_ptr), "update_instances_device")

    def render_device(self, uniforms, params, rgba8=0, radiance=0, hit_ids=0, ray_counts=0):
        """Enqueue a frame whose outputs are caller-owned device buffers (raw pointers)."""
        out = abi.RtFrameOutputs(rgba8 or None, radiance or None, hit_ids or None, ray_counts or None)
        self._check(self.lib.rt_render_device(self.ctx, C.byref(uniforms), C.byref(params), C.byref(out)), "render_device")

    def render_device_slot(self, slot: int, cuda_stream: int, uniforms, params, rgba8=0, radiance=0, hit_ids=0, ray_counts=0):
        """`rt_render_device_slot`: a frame on the caller's stream with the private queues of frame slot 0/1."""
        out = abi.RtFrameOutputs(rgba8 or None, radiance or None, hit_ids or None, ray_counts or None)
        self._check(self.lib.rt_render_device_slot(self.ctx, slot, cuda_stream, C.byref(uniforms), C.byref(params), C.byref(out)), "render_device_slot")

    def render_to_host(self, uniforms, params, host_rgba8_ptr: int, ray_counts_ptr: int = 0):
        """`rt_render` with caller-owned HOST buffers (pinned memory makes the copy asynchronous-capable)."""
        out = abi.RtFrameOutputs(host_rgba8_ptr or None, None, None, ray_counts_ptr or None)
        self._check(self.lib.rt_render(self.ctx, C.byref(uniforms), C.byref(params), C.byref(out)), "render")

    def render_async(self, uniforms, params, host_rgba8_ptr: int, ray_counts_ptr: int = 0) -> int:
        """`rt_render_async`: two frames in flight; returns the frame slot to pass to `wait_frame`."""
        out = abi.RtFrameOutputs(host_rgba8_ptr or None, None, None, ray_counts_ptr or None)
        slot = C.c_uint32()
        self._check(self.lib.rt_render_async(self.ctx, C.byref(uniforms), C.byref(params), C.byref(out), C.byref(slot)), "render_async")
        return slot.value

    def wait_frame(self, slot: int):
        self._check(self.lib.rt_wait_frame(self.ctx, slot), "wait_frame")

    def readback(self, rows: int, width: int) -> np.ndarray:
        img = np.zeros((rows, width, 4), np.uint8)
        self._check(self.lib.rt_readback(self.ctx, img.ctypes.data, img.nbytes), "readback")
        return img

    def sync(self):
        self._check(self.lib.rt_sync(self.ctx), "sync")

    def stats(self) -> abi.RtStats:
        s = abi.RtStats()
        self._check(self.lib.rt_get_stats(self.ctx, C.byref(s)), "get_stats")
        return s

    def push_constants(self) -> abi.RtPushConstantBufferAddresses:
        pc = abi.RtPushConstantBufferAddresses()
        self._check(self.lib.rt_get_push_constants(self.ctx, C.byref(pc)), "get_push_constants")
        return pc

    def read_model_info(self, model_id: int, max_geoms: int = 16):
        info = abi.RtModelInfo()
        geoms = (abi.RtGeometryInfo * max_geoms)()
        self._check(self.lib.rt_debug_read_model_info(self.ctx, model_id, C.byref(info), geoms, max_geoms), "debug_read_model_info")
        return info, geoms

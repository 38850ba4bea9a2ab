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
    "rt_sync", "rt_get_stats", "rt_get_push_constants", "rt_debug_read_model_info", "rt_debug_l2_read_bandwidth", "rt_debug_box_test", "rt_kernel_launches", "rt_version",
    "rt_set_denoise_hook", "rt_denoise_bilateral",
    "rt_group_unique_id", "rt_group_create", "rt_group_destroy", "rt_group_last_error", "rt_group_partition", "rt_group_update_instances",
    "rt_group_update_instances_device", "rt_group_build_tlas",
    "rt_group_render_device", "rt_group_render_host", "rt_group_acquire_device", "rt_group_acquire_host", "rt_group_release",
    "rt_group_readback", "rt_group_local_ray_counts", "rt_group_barrier",
]
GROUP_ID_BYTES, GROUP_STRIP_ROWS, GROUP_FRAME_SLOTS = 128, 8, 4


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
    lib.rt_debug_l2_read_bandwidth.argtypes = [p, C.c_size_t, u32, C.POINTER(C.c_float)]
    lib.rt_debug_box_test.argtypes = [p, C.c_int, u32, u32, p, u32, p, p]
    lib.rt_set_denoise_hook.argtypes = [p, p, p]
    lib.rt_denoise_bilateral.argtypes = [p, p, C.POINTER(abi.RtDenoiseBuffers)]
    u64 = C.c_uint64
    lib.rt_group_unique_id.argtypes = [p]
    lib.rt_group_create.argtypes = [p, C.c_int, C.c_int, p, u32, u32, C.POINTER(p)]
    lib.rt_group_destroy.argtypes = [p]
    lib.rt_group_destroy.restype = None
    lib.rt_group_last_error.argtypes = [p]
    lib.rt_group_last_error.restype = C.c_char_p
    lib.rt_group_partition.argtypes = [p, C.POINTER(abi.RtRenderParams)]
    lib.rt_group_partition.restype = u32
    lib.rt_group_update_instances.argtypes = [p, C.c_int, u32, u32, p, u32]
    lib.rt_group_update_instances_device.argtypes = [p, C.c_int, u32, u32, p, u32]
    lib.rt_group_build_tlas.argtypes = [p, C.c_int, p, u32, u32]
    lib.rt_group_render_device.argtypes = [p, u64, C.POINTER(abi.RtUniforms), C.POINTER(abi.RtRenderParams)]
    lib.rt_group_render_host.argtypes = [p, u64, C.POINTER(abi.RtUniforms), C.POINTER(abi.RtRenderParams)]
    lib.rt_group_acquire_device.argtypes = [p, u64, C.POINTER(p)]
    lib.rt_group_acquire_host.argtypes = [p, u64, u32, C.POINTER(p), C.POINTER(u64)]
    lib.rt_group_release.argtypes = [p, u64]
    lib.rt_group_readback.argtypes = [p, u64, p, C.c_size_t]
    lib.rt_group_local_ray_counts.argtypes = [p, C.POINTER(p)]
    lib.rt_group_barrier.argtypes = [p]
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

    def set_denoise_hook(self, fn=None, user: int = 0):
        """`rt_set_denoise_hook`.  `fn`: None (remove), "bilateral" (the library's rt_denoise_bilateral; `user` = address of a
        float sigma or 0), or a Python callable (user, cuda_stream, buffers: POINTER(RtDenoiseBuffers)) -> int."""
        if fn is None:
            ptr, self._hook = None, None
        elif fn == "bilateral":
            ptr, self._hook = C.cast(self.lib.rt_denoise_bilateral, C.c_void_p), None
        else:
            self._hook = abi.RT_DENOISE_FN(fn)  # keep the trampoline alive
            ptr = C.cast(self._hook, C.c_void_p)
        self._check(self.lib.rt_set_denoise_hook(self.ctx, ptr, user or None), "set_denoise_hook")

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

    def l2_read_bandwidth(self, nbytes: int = 32 << 20, repeats: int = 64) -> float:
        """`rt_debug_l2_read_bandwidth`: GB/s of 16-byte loads over a cache-resident buffer (bench.py's L2 denominator)."""
        out = C.c_float()
        self._check(self.lib.rt_debug_l2_read_bandwidth(self.ctx, nbytes, repeats, C.byref(out)), "debug_l2_read_bandwidth")
        return float(out.value)

    def box_test(self, rays, first_node: int, num_nodes: int, tlas: bool = False):
        """`rt_debug_box_test`: rays (n, 8) float32 {origin, tmin, direction, tmax} -> (masks (n, num_nodes, 2) uint8, node lines
        (num_nodes, 128) uint8): what the traversal's box test reports for every (ray, node) pair."""
        import numpy as np

        rays = np.ascontiguousarray(rays, dtype=np.float32)
        masks = np.zeros((len(rays), num_nodes, 2), np.uint8)
        lines = np.zeros((num_nodes, 128), np.uint8)
        self._check(self.lib.rt_debug_box_test(self.ctx, 1 if tlas else 0, first_node, num_nodes, rays.ctypes.data, len(rays), masks.ctypes.data,
                                               lines.ctypes.data), "debug_box_test")
        return masks, lines

    def push_constants(self) -> abi.RtPushConstantBufferAddresses:
        pc = abi.RtPushConstantBufferAddresses()
        self._check(self.lib.rt_get_push_constants(self.ctx, C.byref(pc)), "get_push_constants")
        return pc

    def read_model_info(self, model_id: int, max_geoms: int = 16):
        info = abi.RtModelInfo()
        geoms = (abi.RtGeometryInfo * max_geoms)()
        self._check(self.lib.rt_debug_read_model_info(self.ctx, model_id, C.byref(info), geoms, max_geoms), "debug_read_model_info")
        return info, geoms


_NCCL_PRELOADED = False


def _preload_nccl():
    """The process can hold ONE libnccl.so.2 (the loader goes by SONAME).  When PyTorch's bundled NCCL is installed, load it
    first: libtorch_cuda needs its symbols, and the C library then finds it with RTLD_NOLOAD.  Without it the C library
    falls back to the system NCCL on its own."""
    global _NCCL_PRELOADED
    if _NCCL_PRELOADED:
        return
    _NCCL_PRELOADED = True
    if os.environ.get("B200RT_NCCL_LIB"):
        return
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            path = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(path):
                C.CDLL(path, mode=C.RTLD_GLOBAL)
                return
    except (ImportError, OSError, AttributeError):
        pass


def group_unique_id() -> bytes:
    """`rt_group_unique_id` (rank 0): the 128 bytes every rank hands to `Group`."""
    _preload_nccl()
    lib = load()
    buf = C.create_string_buffer(GROUP_ID_BYTES)
    rc = lib.rt_group_unique_id(buf)
    if rc != 0:
        msg = lib.rt_group_last_error(None)
        raise RtError(f"rt_group_unique_id failed ({rc}): {msg.decode() if msg else ''}")
    return buf.raw


class Group:
    """`rt_group_*`: one frame on several GPUs of one box (one process per GPU).  Everything that crosses GPUs — the NCCL
    broadcast of instance records, the peer-memory frame on rank 0, the shared page-locked host frame — lives in the
    C library; the host only has to get `unique_id` (128 bytes) from rank 0 to every rank."""

    def __init__(self, renderer: "Renderer", n_ranks: int, rank: int, unique_id: bytes, width: int, height: int):
        _preload_nccl()
        self.lib, self.renderer, self.n, self.rank, self.width, self.height = renderer.lib, renderer, n_ranks, rank, width, height
        h = C.c_void_p()
        idbuf = C.create_string_buffer(bytes(unique_id), GROUP_ID_BYTES)
        rc = self.lib.rt_group_create(renderer.ctx, n_ranks, rank, idbuf, width, height, C.byref(h))
        if rc != 0:
            msg = self.lib.rt_group_last_error(None)
            raise RtError(f"rt_group_create failed ({rc}): {msg.decode() if msg else ''}")
        self.h = h
        self._views = {}

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.rt_group_last_error(self.h)
            raise RtError(f"rt_group_{what} failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if self.h is not None:
            self.lib.rt_group_destroy(self.h)
            self.h = None

    def partition(self, params=None) -> int:
        """Rows this rank renders; fills the strip fields of `params` when given."""
        return self.lib.rt_group_partition(self.h, C.byref(params) if params is not None else None)

    def build_tlas(self, instances=None, count: int = 0, root: int = 0, force_sharded: bool = False):
        """`rt_group_build_tlas`: the root passes the records (numpy, INSTANCE_DTYPE), the other ranks only `count`."""
        ptr = 0
        if instances is not None:
            a = np.ascontiguousarray(instances)
            ptr, count = a.ctypes.data, len(a)
        self._check(self.lib.rt_group_build_tlas(self.h, root, ptr or None, count, 1 if force_sharded else 0), "build_tlas")

    def update_instances(self, first: int, count: int, host_ptr: int, mode: int, root: int = 0):
        self._check(self.lib.rt_group_update_instances(self.h, root, first, count, host_ptr or None, mode), "update_instances")

    def update_instances_device(self, first: int, count: int, device_ptr: int, mode: int, root: int = 0):
        self._check(self.lib.rt_group_update_instances_device(self.h, root, first, count, device_ptr or None, mode), "update_instances_device")

    def render_device(self, seq: int, uniforms, params):
        self._check(self.lib.rt_group_render_device(self.h, seq, C.byref(uniforms), C.byref(params)), "render_device")

    def render_host(self, seq: int, uniforms, params):
        self._check(self.lib.rt_group_render_host(self.h, seq, C.byref(uniforms), C.byref(params)), "render_host")

    def acquire_device(self, seq: int) -> int:
        ptr = C.c_void_p()
        self._check(self.lib.rt_group_acquire_device(self.h, seq, C.byref(ptr)), "acquire_device")
        return ptr.value

    def acquire_host(self, seq: int, timeout_ms: int = 10000):
        """Rank 0: (numpy view of the finished [H, W, 4] frame in the shared host memory, summed ray counts)."""
        ptr = C.c_void_p()
        counts = (C.c_uint64 * 2)()
        self._check(self.lib.rt_group_acquire_host(self.h, seq, timeout_ms, C.byref(ptr), counts), "acquire_host")
        view = self._views.get(ptr.value)
        if view is None:  # one numpy view per frame slot of the shared segment
            buf = (C.c_uint8 * (self.width * self.height * 4)).from_address(ptr.value)
            view = self._views[ptr.value] = np.frombuffer(buf, np.uint8).reshape(self.height, self.width, 4)
        return view, (int(counts[0]), int(counts[1]))

    def readback(self, seq: int) -> np.ndarray:
        """Rank 0: wait for frame `seq` of the device path and copy it to host memory."""
        img = np.zeros((self.height, self.width, 4), np.uint8)
        self._check(self.lib.rt_group_readback(self.h, seq, img.ctypes.data, img.nbytes), "readback")
        return img

    def release(self, seq: int):
        self._check(self.lib.rt_group_release(self.h, seq), "release")

    def local_ray_counts_ptr(self) -> int:
        ptr = C.c_void_p()
        self._check(self.lib.rt_group_local_ray_counts(self.h, C.byref(ptr)), "local_ray_counts")
        return ptr.value

    def barrier(self):
        self._check(self.lib.rt_group_barrier(self.h), "barrier")


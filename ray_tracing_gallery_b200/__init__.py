"""B200-native frame hot path of expenses/ray-tracing-gallery.

Host-side mirror (Python) of the reference's scene/loader interface on top of
the C ABI in include/b200rt.h; the CUDA kernels live in csrc/ and are built
in-tree into libb200rt.so by __graft_entry__.build().  There is no CPU
fallback: `native.Renderer` raises if the library or a CUDA device is missing.
"""
from . import abi  # noqa: F401

__all__ = ["abi", "backend", "gltf", "scene", "native"]

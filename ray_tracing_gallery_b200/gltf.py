"""GLB -> model arrays, with the semantics of the reference's `Model::load_gltf`.

Mirrors src/util_structs.rs:903-1156 (ModelArrays, load_images_from_material,
load_image_from_gltf_texture, Model::load_gltf) and the PNG path of
src/util_functions.rs:239-265:

  * GLB only; images must be bufferView PNGs (:957-968);
  * one geometry per *material* (:1079-1088); a model without materials gets one
    default geometry: diffuse = caller's fallback image index, metallic-roughness
    = a 1x1 RGBA32F constant (1, roughness=1, metallic=0, 1), opaque (:1090-1111);
  * images are pushed in material order: diffuse (sRGB), metallic-roughness
    (sRGB - a reference quirk that is kept), optional normal map (UNORM) (:994-1020);
    a missing texture becomes a 1x1 RGBA32F constant, nearest-filtered (:938-950);
  * linear filtering iff the glTF sampler's magFilter != NEAREST (:954-955);
  * mesh primitives are appended in order into shared vertex arrays, their
    indices rebased by the running vertex count and appended to the geometry of
    their material (`unwrap_or(0)`), node transforms are ignored (:1113-1137).
"""
import io
import json
import struct
from dataclasses import dataclass, field
from typing import Callable, List

import numpy as np

from .abi import RT_FORMAT_RGBA8_SRGB, RT_FORMAT_RGBA8_UNORM, RT_FORMAT_RGBA32_SFLOAT

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}
_MAG_NEAREST = 9728

# `image` 0.23's `to_rgba8()` narrows 16-bit samples; the crate is not vendored in
# the reference, so the rule is switchable.  ">>8" is the default.
U16_TO_U8_RULE = "shift"


@dataclass
class Geometry:
    indices: np.ndarray  # uint32, multiple of 3, rebased to the model's vertex arrays
    opaque: bool
    diffuse_image_index: int
    metallic_roughness_image_index: int
    normal_map_image_index: int  # -1 = none


@dataclass
class ModelArrays:
    name: str
    positions: np.ndarray  # (V,3) float32
    normals: np.ndarray  # (V,3) float32
    uvs: np.ndarray  # (V,2) float32
    geometries: List[Geometry] = field(default_factory=list)

    @property
    def num_triangles(self):
        return sum(len(g.indices) // 3 for g in self.geometries)


def decode_png_rgba8(data: bytes) -> np.ndarray:
    """`image::load_from_memory_with_format(.., Png)?.to_rgba8()` (src/util_functions.rs:247-249)."""
    from PIL import Image

    img = Image.open(io.BytesIO(data))
    img.load()
    if img.mode in ("I;16", "I;16B", "I;16L", "I"):
        a = np.asarray(img).astype(np.uint32)
        if U16_TO_U8_RULE == "shift":
            g = (a >> 8).astype(np.uint8)
        else:  # rounding rule of later `image` versions
            g = ((a * 255 + 32767) // 65535).astype(np.uint8)
        out = np.empty(g.shape + (4,), np.uint8)
        out[..., 0] = out[..., 1] = out[..., 2] = g
        out[..., 3] = 255
        return out
    return np.ascontiguousarray(np.asarray(img.convert("RGBA"), dtype=np.uint8))


def load_png_file_rgba8(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        return decode_png_rgba8(f.read())


def _split_glb(data: bytes):
    magic, version, length = struct.unpack_from("<III", data, 0)
    if magic != 0x46546C67:
        raise ValueError("not a GLB container")
    off, js, blob = 12, None, None
    while off < length:
        clen, ctype = struct.unpack_from("<II", data, off)
        chunk = data[off + 8 : off + 8 + clen]
        if ctype == 0x4E4F534A:
            js = json.loads(chunk.decode("utf-8"))
        elif ctype == 0x004E4942 and blob is None:
            blob = chunk
        off += 8 + clen
    if js is None:
        raise ValueError("GLB without JSON chunk")
    if blob is None:
        raise ValueError("GLB without binary chunk")  # `gltf.blob.as_ref().unwrap()`, :1069
    return js, blob


def _read_accessor(js, blob, index) -> np.ndarray:
    acc = js["accessors"][index]
    if "sparse" in acc:
        raise ValueError("sparse accessors are not supported")
    dt = np.dtype(_COMPONENT[acc["componentType"]]).newbyteorder("<")
    ncomp = _NCOMP[acc["type"]]
    count = acc["count"]
    view = js["bufferViews"][acc["bufferView"]]
    if view.get("buffer", 0) != 0:
        raise ValueError("only buffer 0 (the GLB blob) is supported")
    base = view.get("byteOffset", 0) + acc.get("byteOffset", 0)
    elem = dt.itemsize * ncomp
    stride = view.get("byteStride", 0) or elem
    if stride == elem:
        arr = np.frombuffer(blob, dtype=dt, count=count * ncomp, offset=base).reshape(count, ncomp)
    else:
        raw = np.frombuffer(blob, dtype=np.uint8, count=(count - 1) * stride + elem, offset=base)
        idx = np.arange(count)[:, None] * stride + np.arange(elem)[None, :]
        arr = raw[idx].copy().view(dt).reshape(count, ncomp)
    if acc.get("normalized", False) and dt.kind in "ui":
        # gltf crate `into_f32()` for normalised integer texcoords
        arr = arr.astype(np.float32) / np.float32(np.iinfo(dt).max)
    return arr


def load_gltf(
    data: bytes,
    name: str,
    fallback_image_index: int,
    push_image: Callable[[np.ndarray, int, bool], int],
) -> ModelArrays:
    """`Model::load_gltf` up to (not including) the GPU upload.  `push_image(texels, format,
    linear_filter) -> index` plays `ImageManager::push_image`."""
    js, blob = _split_glb(data)

    def image_from_texture(tex_info, backup_rgba, fmt):
        if tex_info is None:
            texel = np.asarray(backup_rgba, np.float32).reshape(1, 1, 4)
            return push_image(texel, RT_FORMAT_RGBA32_SFLOAT, False)
        tex = js["textures"][tex_info["index"]]
        linear = True
        if "sampler" in tex:
            linear = js["samplers"][tex["sampler"]].get("magFilter") != _MAG_NEAREST
        img = js["images"][tex["source"]]
        if "bufferView" not in img:
            raise ValueError("Image source is a uri which we don't support")
        view = js["bufferViews"][img["bufferView"]]
        start = view.get("byteOffset", 0)
        texels = decode_png_rgba8(bytes(blob[start : start + view["byteLength"]]))
        return push_image(texels, fmt, linear)

    geoms: List[Geometry] = []
    for material in js.get("materials", []):
        pbr = material.get("pbrMetallicRoughness", {})
        base_factor = pbr.get("baseColorFactor", [1.0, 1.0, 1.0, 1.0])
        metallic = pbr.get("metallicFactor", 1.0)
        roughness = pbr.get("roughnessFactor", 1.0)
        diffuse = image_from_texture(pbr.get("baseColorTexture"), base_factor, RT_FORMAT_RGBA8_SRGB)
        mr = image_from_texture(pbr.get("metallicRoughnessTexture"), [1.0, roughness, metallic, 1.0], RT_FORMAT_RGBA8_SRGB)
        nt = material.get("normalTexture")
        nm = image_from_texture(nt, None, RT_FORMAT_RGBA8_UNORM) if nt is not None else -1
        geoms.append(
            Geometry(
                indices=np.zeros(0, np.uint32),
                opaque=material.get("alphaMode", "OPAQUE") == "OPAQUE",
                diffuse_image_index=diffuse,
                metallic_roughness_image_index=mr,
                normal_map_image_index=nm,
            )
        )
    if not geoms:
        mr = push_image(np.asarray([1.0, 1.0, 0.0, 1.0], np.float32).reshape(1, 1, 4), RT_FORMAT_RGBA32_SFLOAT, False)
        geoms.append(Geometry(np.zeros(0, np.uint32), True, fallback_image_index, mr, -1))

    positions, normals, uvs = [], [], []
    index_lists = [[] for _ in geoms]
    num_vertices = 0
    for mesh in js.get("meshes", []):
        for prim in mesh["primitives"]:
            gi = prim.get("material", 0)
            attrs = prim["attributes"]
            idx = _read_accessor(js, blob, prim["indices"]).reshape(-1).astype(np.uint32)
            pos = _read_accessor(js, blob, attrs["POSITION"]).astype(np.float32)
            nrm = _read_accessor(js, blob, attrs["NORMAL"]).astype(np.float32)
            uv = _read_accessor(js, blob, attrs["TEXCOORD_0"]).astype(np.float32)
            index_lists[gi].append(idx + np.uint32(num_vertices))
            positions.append(pos)
            normals.append(nrm)
            uvs.append(uv)
            num_vertices += len(pos)
    for g, lst in zip(geoms, index_lists):
        g.indices = np.ascontiguousarray(np.concatenate(lst) if lst else np.zeros(0, np.uint32), dtype=np.uint32)

    def cat(parts, width):
        return np.ascontiguousarray(np.concatenate(parts) if parts else np.zeros((0, width)), dtype=np.float32)

    return ModelArrays(name, cat(positions, 3), cat(normals, 3), cat(uvs, 2), geoms)

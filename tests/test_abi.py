"""Struct layouts and the C-ABI surface.  CPU only: no compute calls."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

from ray_tracing_gallery_b200 import abi, native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def off(struct, field):
    return getattr(struct, field).offset


def test_uniforms_layout():
    # shared-structs/src/lib.rs:10-19 == hit_shader_common.glsl:3-13 (scalar layout)
    assert C.sizeof(abi.RtUniforms) == 176
    assert off(abi.RtUniforms, "view_inverse") == 0
    assert off(abi.RtUniforms, "proj_inverse") == 64
    assert off(abi.RtUniforms, "sun_dir") == 128
    assert off(abi.RtUniforms, "sun_radius") == 144
    assert off(abi.RtUniforms, "blue_noise_texture_index") == 148
    assert off(abi.RtUniforms, "ggx_lut_texture_index") == 152
    assert off(abi.RtUniforms, "frame_index") == 156
    assert off(abi.RtUniforms, "show_heatmap") == 160


def test_model_geometry_push_constant_layouts():
    assert C.sizeof(abi.RtModelInfo) == 32
    assert [off(abi.RtModelInfo, f) for f in ("position_buffer_address", "normal_buffer_address", "uv_buffer_address", "geometry_info_address")] == [0, 8, 16, 24]
    assert C.sizeof(abi.RtGeometryInfo) == 24 and off(abi.RtGeometryInfo, "images") == 8
    assert C.sizeof(abi.RtGeometryImages) == 16
    assert C.sizeof(abi.RtPushConstantBufferAddresses) == 24


def test_instance_layout():
    # src/gpu_structs.rs:20-25 == VkAccelerationStructureInstanceKHR
    assert C.sizeof(abi.RtInstance) == 64
    assert off(abi.RtInstance, "instance_custom_index_and_mask") == 48
    assert off(abi.RtInstance, "sbt_record_offset_and_flags") == 52
    assert off(abi.RtInstance, "acceleration_structure_device_address") == 56
    assert abi.INSTANCE_DTYPE.itemsize == 64
    assert abi.INSTANCE_DTYPE.fields["custom_index_and_mask"][1] == 48
    assert abi.INSTANCE_DTYPE.fields["blas"][1] == 56


def test_instance_packing_matches_reference_constructor():
    from ray_tracing_gallery_b200.scene import make_instance, mat_translation

    rec = make_instance(mat_translation(1, 2, 3), model_id=5, blas_handle=0xABCDEF, hit_shader=abi.RT_HIT_MIRROR, double_sided=True)
    t = rec["transform"].reshape(3, 4)
    assert np.allclose(t[:, 3], [1, 2, 3]) and np.allclose(t[:, :3], np.eye(3))
    assert int(rec["custom_index_and_mask"]) == 5 | (0xFF << 24)
    assert int(rec["sbt_offset_and_flags"]) == 1 | (1 << 24)
    assert int(rec["blas"]) == 0xABCDEF


def test_headers_compile_as_c_and_cxx(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "b200rt.h"\nint main(void){return (int)sizeof(RtUniforms)-176;}\n')
    for cc, std in (("gcc", "-std=c11"), ("g++", "-std=c++17")):
        subprocess.check_call([cc, std, "-x", "c" if cc == "gcc" else "c++", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(tmp_path / "t")])
        assert subprocess.call([str(tmp_path / "t")]) == 0


def test_ctypes_mirrors_match_the_headers_field_by_field(tmp_path):
    """Compile a C program that prints sizeof / offsetof of every field the ctypes mirrors declare and compare: ties
    abi.py (and through test_rust_binding... the Rust file) to include/*.h mechanically."""
    structs = ["RtUniforms", "RtModelInfo", "RtGeometryImages", "RtGeometryInfo", "RtPushConstantBufferAddresses", "RtInstance",
               "RtGeometryDesc", "RtModelDesc", "RtRenderParams", "RtFrameOutputs", "RtStats"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200rt.h"', "int main(void){"]
    for name in structs:
        cls = getattr(abi, name)
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for name in structs:
        cls = getattr(abi, name)
        assert int(got[name]) == C.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(got[f"{name}.{field}"]) == getattr(cls, field).offset, f"{name}.{field}"
    assert C.sizeof(abi.RtRenderParams) == 64 and abi.RtRenderParams.heatmap_scale.offset == 52
    assert C.sizeof(abi.RtFrameOutputs) == 40 and abi.RtFrameOutputs.cost_cycles.offset == 32


def test_constants_match_the_headers(tmp_path):
    """Every RT_* integer constant that abi.py defines and include/*.h mention (enum values, flag bits, limits) must have the
    header's value: compiled and printed by a C program, like the layouts above."""
    header = open(os.path.join(ROOT, "include", "b200rt.h")).read() + open(os.path.join(ROOT, "include", "rt_abi.h")).read()
    names = sorted(n for n in dir(abi) if re.fullmatch(r"RT_[A-Z0-9_]+", n) and isinstance(getattr(abi, n), int) and re.search(r"\b%s\b" % n, header))
    assert {"RT_UPDATE_AUTO", "RT_UPDATE_REFIT", "RT_UPDATE_REBUILD", "RT_UPDATE_REBUILD_FAST", "RT_RENDER_COUNTERS", "RT_FORMAT_RGBA8_SRGB"} <= set(names)
    lines = ['#include <stdio.h>', '#include "b200rt.h"', "int main(void){"] + [f'printf("{n} %lld\\n", (long long)({n}));' for n in names] + ["return 0;}"]
    src = tmp_path / "consts.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "consts"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for n in names:
        assert int(got[n]) == getattr(abi, n), n


def declared_functions():
    text = open(os.path.join(ROOT, "include", "b200rt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 18
    assert sorted(native.EXPORTS) == names
    lib = native.load()
    for n in names:
        getattr(lib, n)  # raises AttributeError if the symbol is missing
    assert lib.rt_version() >> 16 == 1


def test_rust_binding_file_declares_every_export_and_matches_the_layouts():
    """bindings/b200rt_ffi.rs is not compiled here (no Rust toolchain); keep it in step with the header mechanically."""
    rs = open(os.path.join(ROOT, "bindings", "b200rt_ffi.rs")).read()
    declared = sorted(set(re.findall(r"pub fn (rt_[a-z_0-9]+)\s*\(", rs)))
    assert declared == declared_functions()
    # struct field order of the plain-data structs must equal the ctypes mirrors (which tests above tie to rt_abi.h / b200rt.h)
    for name, cls in (("RtRenderParams", abi.RtRenderParams), ("RtFrameOutputs", abi.RtFrameOutputs), ("RtStats", abi.RtStats)):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, rs, flags=re.S).group(1)
        fields = re.findall(r"pub ([a-z_0-9]+):", body)
        assert fields == [f[0] for f in cls._fields_], name
    for const in ("RT_RENDER_COUNTERS", "RT_RENDER_TIMING", "RT_RENDER_SPLIT_TAIL", "RT_RENDER_NO_PDL", "RT_RENDER_OUTPUT_IMAGE_ROWS", "RT_RENDER_COOP_TAIL",
                  "RT_UPDATE_AUTO", "RT_UPDATE_REFIT", "RT_UPDATE_REBUILD", "RT_UPDATE_REBUILD_FAST", "RT_FORMAT_RGBA8_UNORM", "RT_FORMAT_RGBA8_SRGB", "RT_FORMAT_RGBA32_SFLOAT"):
        m = re.search(r"pub const %s: u32 = (\d+);" % const, rs)
        assert m and int(m.group(1)) == getattr(abi, const), const


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        return
    from ray_tracing_gallery_b200.backend import RtError

    try:
        native.Renderer(0)
    except RtError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("Renderer() must fail loudly without a CUDA device")


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "ray_tracing_gallery_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liborc" not in text and "rt_oracle" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert not re.search(r'#include\s+"[^"]*oracle', text), f

"""Parity of the CUDA path (through the C ABI of libb200rt.so) against the CPU oracle.

Bars (BASELINE.json north_star): hit IDs bit-exact except tie/edge pixels (>= 99.99 %), radiance
within 1e-3 relative per pixel or PSNR >= 50 dB, identical blue-noise sequences.  The oracle is the
checker only; everything under test goes through `native.Renderer` (ctypes -> rt_*)."""
import ctypes as C

import numpy as np
import pytest

from conftest import make_oracle, make_renderer
from ray_tracing_gallery_b200 import abi
from ray_tracing_gallery_b200.backend import RtError
from ray_tracing_gallery_b200.dist import Partition, deinterleave
from ray_tracing_gallery_b200.scene import (build_scene, load_model, make_instance, mat_identity, mat_scale, mat_translation,
                                            push_builtin_images)

pytestmark = pytest.mark.gpu

PIPELINES = [abi.RT_PIPELINE_WAVEFRONT, abi.RT_PIPELINE_MEGAKERNEL]
REL_TOL = 1e-3        # north_star: radiance within 1e-3 relative per pixel ...
PSNR_MIN_DB = 50.0    # ... or PSNR >= 50 dB
ID_AGREEMENT = 0.9999


def psnr(a, b):
    fin = np.isfinite(a) & np.isfinite(b)
    mse = np.mean(np.where(fin, a.astype(np.float64) - b.astype(np.float64), 0.0) ** 2)
    return 999.0 if mse == 0 else 10 * np.log10(1.0 / mse)


def check_parity(gpu_out, cpu_out, strict_ids=True):
    ids_ok = np.all(gpu_out["hit_ids"] == cpu_out["hit_ids"], axis=(2, 3))
    assert ids_ok.mean() >= (1.0 if strict_ids else ID_AGREEMENT), f"hit-ID agreement {ids_ok.mean():.6f}"
    a, b = gpu_out["radiance"].astype(np.float64), cpu_out["radiance"].astype(np.float64)
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
    within = np.mean(np.all((rel <= REL_TOL) | ~np.isfinite(b), axis=2))
    assert within >= 0.999 or psnr(a, b) >= PSNR_MIN_DB, f"radiance: {within:.6f} of pixels within 1e-3, PSNR {psnr(a, b):.1f} dB"
    assert psnr(a, b) >= PSNR_MIN_DB
    d8 = np.abs(gpu_out["rgba8"].astype(int) - cpu_out["rgba8"].astype(int))
    assert np.mean(d8.max(axis=2) <= 1) >= 0.999
    assert np.array_equal(gpu_out["ray_counts"], cpu_out["ray_counts"])
    return ids_ok.mean(), within


def both(cfg, w, h, **kw):
    orc, gpu = make_oracle(), make_renderer()
    return orc, build_scene(orc, cfg, w, h, **kw), gpu, build_scene(gpu, cfg, w, h, **kw)


@pytest.mark.parametrize("cfg,size,kw", [
    ("c1", (1280, 720), {}),                          # BASELINE configs[0], full size
    ("c2", (1920, 1080), {}),                         # configs[1], full size
    ("c3", (1920, 1080), {}),                         # configs[2], full size
    ("default", (1280, 720), {}),                     # the reference's own DefaultScene (portal, lain, fence, 100 tori)
])
def test_frame_parity_full_size(cfg, size, kw):
    orc, so, gpu, sg = both(cfg, *size, **kw)
    want = orc.render(so.uniforms(), so.params())
    for pipeline in PIPELINES:
        got = gpu.render(sg.uniforms(), sg.params(pipeline=pipeline))
        gpu.stats()  # raises on traversal-stack overflow
        check_parity(got, want)
    orc.close(); gpu.close()


def test_synthetic_multi_material_normal_mapped_model():
    """normal map (closest_hit_textured.glsl:141-157), NEAREST sampler (util_structs.rs:954-955), non-power-of-two
    images with REPEAT, two materials = two geometries, masked geometry next to an opaque one, non-uniform instance scale."""
    import synth_assets

    orc, gpu = make_oracle(), make_renderer()
    so, sg = synth_assets.build_bumpy_scene(orc, 960, 540), synth_assets.build_bumpy_scene(gpu, 960, 540)
    want = orc.render(so.uniforms(), so.params())
    for pipeline in PIPELINES:
        got = gpu.render(sg.uniforms(), sg.params(pipeline=pipeline))
        gpu.stats()
        check_parity(got, want, strict_ids=False)
    st = gpu.stats()
    orc.close(); gpu.close()


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_stress_scenes(seed):
    """Seeded random scenes (tests/synth_assets.py): triangle soups, mirrored / non-uniformly scaled / duplicated instances,
    coincident twin triangles (exact t ties), a degenerate triangle, masked geometry with a NEAREST texture, all three hit
    groups, 1..5 shadow rays.  Same bars as the named configs."""
    import synth_assets

    orc, gpu = make_oracle(), make_renderer()
    so, sg = synth_assets.build_random_scene(orc, seed), synth_assets.build_random_scene(gpu, seed)
    want = orc.render(so.uniforms(), so.params())
    for pipeline in PIPELINES:
        got = gpu.render(sg.uniforms(), sg.params(pipeline=pipeline))
        gpu.stats()
        check_parity(got, want, strict_ids=False)
    ids = got["hit_ids"].reshape(-1, 3)
    assert 15 not in set(ids[ids[:, 0] != abi.MISS_ID][:, 0].tolist())  # the duplicated instance never wins a tie
    orc.close(); gpu.close()


def test_exact_instance_bounds_stay_conservative():
    """Instances get the world bounds of their transformed VERTICES (models up to 4 096 vertices) instead of the bounds of the eight
    transformed box corners.  Thin rods along the diagonal of their object box under rotations, shears and mirrorings are where the
    two differ most; unreferenced NaN / 1e30 vertices and a model above the vertex limit ride along.  Hit IDs and radiance against
    the oracle, after a build, a refit and a rebuild of the same records."""
    import synth_assets

    orc, gpu = make_oracle(), make_renderer()
    so, sg = synth_assets.build_rod_scene(orc), synth_assets.build_rod_scene(gpu)
    want = orc.render(so.uniforms(), so.params())
    for mode in (None, abi.RT_UPDATE_REFIT, abi.RT_UPDATE_REBUILD):
        if mode is not None:
            gpu.update_instances(0, sg.instances)
            gpu.update_tlas(mode)
        for pipeline in PIPELINES:
            got = gpu.render(sg.uniforms(), sg.params(pipeline=pipeline))
            gpu.stats()
            ids, within = check_parity(got, want, strict_ids=False)
    print(f"rods: hit IDs {ids:.4%}, radiance within 1e-3 {within:.4%}, hits {np.mean(got['hit_ids'][..., 0, 0] != abi.MISS_ID):.1%} of the pixels")
    orc.close(); gpu.close()


@pytest.mark.parametrize("offset", [(300.0, -200.0, 500.0), (4000.0, -2500.0, 7000.0)])
def test_random_scene_far_from_the_origin(offset):
    """The seeded stress scene moved hundreds / thousands of units away from the origin: node grids of a few millimetres per cell
    against coordinates whose fp32 spacing is up to half a millimetre.  The traversal's box test has to stay conservative with its
    fp32 guards (|b| 2^-21 terms) doing real work; hit IDs and radiance are held to the oracle as everywhere else."""
    import synth_assets

    orc, gpu = make_oracle(), make_renderer()
    so, sg = synth_assets.build_random_scene(orc, 3, offset=offset), synth_assets.build_random_scene(gpu, 3, offset=offset)
    want = orc.render(so.uniforms(), so.params())
    for pipeline in PIPELINES:
        got = gpu.render(sg.uniforms(), sg.params(pipeline=pipeline))
        gpu.stats()
        ids, within = check_parity(got, want, strict_ids=False)
    print(f"offset {offset}: hit IDs {ids:.4%}, radiance within 1e-3 {within:.4%}, hits {np.mean(got['hit_ids'][..., 0, 0] != abi.MISS_ID):.1%} of the pixels")
    orc.close(); gpu.close()


def test_pipelines_agree_bit_for_bit():
    gpu = make_renderer()
    s = build_scene(gpu, "default", 640, 360)
    a = gpu.render(s.uniforms(), s.params(pipeline=abi.RT_PIPELINE_WAVEFRONT))
    b = gpu.render(s.uniforms(), s.params(pipeline=abi.RT_PIPELINE_MEGAKERNEL))
    c = gpu.render(s.uniforms(), s.params(pipeline=abi.RT_PIPELINE_WAVEFRONT))
    for k in ("hit_ids", "rgba8", "ray_counts"):
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k])
    assert np.array_equal(a["radiance"].view(np.uint32), b["radiance"].view(np.uint32))
    assert np.array_equal(a["radiance"].view(np.uint32), c["radiance"].view(np.uint32))  # run-to-run deterministic
    gpu.close()


def test_bounce_segments_cooperative_tail_equals_split_launches():
    """Segments >= 1 run in one cooperative kernel by default; RT_RENDER_SPLIT_TAIL runs them as four launches each.
    Same phases, same order: the frames must be identical.  Also covers max_segments = 1 (no tail) and > 8 (counter-slot reuse)."""
    gpu = make_renderer()
    s = build_scene(gpu, "c3", 640, 360)
    for segs in (1, 3, 10):
        a = gpu.render(s.uniforms(), s.params(max_segments=segs, flags=abi.RT_RENDER_COOP_TAIL))
        if segs == 3:
            assert gpu.stats().segment_rays[0] > 0, "the config must have bounce rays for this test to mean anything"
        b = gpu.render(s.uniforms(), s.params(max_segments=segs, flags=abi.RT_RENDER_SPLIT_TAIL))
        m = gpu.render(s.uniforms(), s.params(max_segments=segs, pipeline=abi.RT_PIPELINE_MEGAKERNEL))
        for k in ("hit_ids", "rgba8", "ray_counts"):
            assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], m[k]), (segs, k)
        assert np.array_equal(a["radiance"].view(np.uint32), b["radiance"].view(np.uint32))
    gpu.close()


def test_tail_policy_follows_the_bounce_rays_of_the_latest_frame():
    """Without RT_RENDER_SPLIT_TAIL / RT_RENDER_COOP_TAIL the library picks the tail per frame: the cooperative k_tail while
    frames bounce little, separate launches once the latest frame queued many bounce rays (the count reaches the host
    through a host-mapped word, no synchronisation).  Whatever it picks, the frame is the same."""
    gpu = make_renderer()
    s = build_scene(gpu, "c3", 1920, 1080)
    timing = abi.RT_RENDER_TIMING
    ref = gpu.render(s.uniforms(), s.params(flags=abi.RT_RENDER_COOP_TAIL | timing))
    st = gpu.stats()
    assert st.kernel_launches[5] == 1 and st.kernel_launches[3] == 0  # K_TAIL, K_RESOLVE
    assert st.segment_rays[0] >= 131072, "C3 at 1080p must bounce enough for the policy to switch"
    b = gpu.render(s.uniforms(), s.params(flags=timing))  # the frame before this one bounced a lot -> separate launches
    st = gpu.stats()
    assert st.kernel_launches[5] == 0 and st.kernel_launches[3] == 3, list(st.kernel_launches)
    for k in ("hit_ids", "rgba8", "ray_counts"):
        assert np.array_equal(ref[k], b[k]), k
    assert np.array_equal(ref["radiance"].view(np.uint32), b["radiance"].view(np.uint32))
    # a scene without bounce rays goes back to the cooperative tail after one frame
    gpu2 = make_renderer()
    s2 = build_scene(gpu2, "c1", 640, 360)
    gpu2.render(s2.uniforms(), s2.params(flags=timing))
    assert gpu2.stats().kernel_launches[5] == 1
    gpu.close(); gpu2.close()


def test_blue_noise_sequence_over_frames():
    """Soft shadows must follow the reference's deterministic sequence: frame f and f+32 are identical,
    other frames differ, and every frame matches the oracle."""
    orc, so, gpu, sg = both("c2", 480, 270)
    frames = {}
    for f in (1, 2, 17, 33):
        want = orc.render(so.uniforms(frame_index=f), so.params())
        got = gpu.render(sg.uniforms(frame_index=f), sg.params())
        check_parity(got, want)
        frames[f] = got["radiance"]
    assert np.array_equal(frames[1], frames[33])
    assert not np.array_equal(frames[1], frames[2])
    orc.close(); gpu.close()


def test_hard_shadow_config_is_exact_in_ids_and_counts():
    """C1 (sun_radius 0): the shadow direction is exactly normalize(sun_dir), so every ray of the frame is
    reproducible; only pow/cos ulps remain in the radiance."""
    orc, so, gpu, sg = both("c1", 640, 360)
    want, got = orc.render(so.uniforms(), so.params()), gpu.render(sg.uniforms(), sg.params())
    assert np.array_equal(got["hit_ids"], want["hit_ids"])
    lit_gpu = got["radiance"].sum(axis=2) > 0.2 * 0.3
    lit_cpu = want["radiance"].sum(axis=2) > 0.2 * 0.3
    assert np.array_equal(lit_gpu, lit_cpu)
    assert np.max(np.abs(got["radiance"] - want["radiance"]) / np.maximum(want["radiance"], 1e-3)) < 1e-4
    orc.close(); gpu.close()


def test_tiles_and_strips_reassemble_the_full_frame():
    gpu = make_renderer()
    W, H = 320, 192
    s = build_scene(gpu, "default", W, H)
    full = gpu.render(s.uniforms(), s.params())
    # a tile uses global pixel coordinates (gl_LaunchIDEXT) for rays and blue noise
    tile = gpu.render(s.uniforms(), s.params(tile_x0=40, tile_y0=64, tile_w=120, tile_h=72))
    assert np.array_equal(tile["rgba8"], full["rgba8"][64:136, 40:160])
    assert np.array_equal(tile["hit_ids"], full["hit_ids"][64:136, 40:160])
    for world in (2, 4, 5, 8):   # 24 strips of 8 rows: even shares for 2/4/8, ragged for 5
        parts = [Partition.make(W, H, world, r) for r in range(world)]
        slabs, rays = np.zeros((world, parts[0].max_rows, W, 4), np.uint8), np.zeros(2, np.uint64)
        for r, part in enumerate(parts):
            out = gpu.render(s.uniforms(), part.apply(s.params()))
            assert out["rgba8"].shape == (part.local_rows, W, 4)
            slabs[r, : part.local_rows] = out["rgba8"]
            rays += out["ray_counts"]
        assert np.array_equal(deinterleave(slabs, parts[0]), full["rgba8"])
        assert np.array_equal(rays, full["ray_counts"])
    gpu.close()
    # a frame height that is no multiple of the strip height: the last strip is partial (100 = 12 * 8 + 4)
    gpu = make_renderer()
    s2 = build_scene(gpu, "c1", 200, 100)
    full = gpu.render(s2.uniforms(), s2.params())
    parts = [Partition.make(200, 100, 3, r) for r in range(3)]
    assert [p.local_rows for p in parts] == [36, 32, 32]
    slabs = np.zeros((3, 36, 200, 4), np.uint8)
    for r, part in enumerate(parts):
        out = gpu.render(s2.uniforms(), part.apply(s2.params()))
        assert out["rgba8"].shape == (part.local_rows, 200, 4)
        slabs[r, : part.local_rows] = out["rgba8"]
    assert np.array_equal(deinterleave(slabs, parts[0]), full["rgba8"])
    gpu.close()


def test_sah_trees_are_built_and_pay():
    """Static builds (rt_build_tlas) and per-frame rebuilds (rt_update_tlas REBUILD) both grow binned-SAH trees in one cooperative
    launch; above 8 192 primitives the large nodes of a level are split by the whole grid.  On one scene (C5's slab with 80 000 and with
    4 000 instances, i.e. with and without large nodes): build, rebuild, refit and the radix-tree rebuild (RT_UPDATE_REBUILD_FAST) must
    give identical frames (closest hit + tie rule: frames do not depend on the tree), the rebuilt tree must cost what the built one costs,
    and the SAH tree must need clearly fewer node visits per ray than the radix tree (18.9 against 32.4 at 80 000 instances): a silent
    fall-back to the radix tree (no cooperative launch) would pass every parity test and only show up here."""
    def nodes_per_ray(gpu, s):
        out = gpu.render(s.uniforms(), s.params(flags=abi.RT_RENDER_COUNTERS))
        st = gpu.stats()
        return sum(st.nodes_visited) / max(1, st.primary_rays + st.shadow_rays), out

    seen = {}
    for n in (80000, 4000):
        gpu = make_renderer()
        s = build_scene(gpu, "c5", 640, 360, num_instances=n)
        built, frame_a = nodes_per_ray(gpu, s)
        gpu.update_instances(0, s.instances); gpu.update_tlas(abi.RT_UPDATE_REBUILD)
        rebuilt, frame_b = nodes_per_ray(gpu, s)
        gpu.update_instances(0, s.instances); gpu.update_tlas(abi.RT_UPDATE_REFIT)
        refitted, frame_c = nodes_per_ray(gpu, s)
        gpu.update_instances(0, s.instances); gpu.update_tlas(abi.RT_UPDATE_REBUILD_FAST)   # the Morton radix tree
        radix, frame_d = nodes_per_ray(gpu, s)
        for k in ("rgba8", "hit_ids", "ray_counts"):
            assert np.array_equal(frame_a[k], frame_b[k]) and np.array_equal(frame_a[k], frame_c[k]) and np.array_equal(frame_a[k], frame_d[k]), k
        assert abs(rebuilt - built) < 0.02 * built   # the same algorithm (node numbering may differ, the splits do not)
        assert abs(refitted - rebuilt) < 1e-9
        seen[n] = (built, radix)
        gpu.close()
    print(f"nodes per ray, SAH / radix tree: 80 k instances {seen[80000][0]:.2f} / {seen[80000][1]:.2f}, 4 k instances {seen[4000][0]:.2f} / {seen[4000][1]:.2f}")
    assert seen[80000][0] < 0.9 * seen[80000][1] and seen[80000][0] < 24.0


@pytest.mark.parametrize("mode", [abi.RT_UPDATE_REBUILD, abi.RT_UPDATE_REFIT, abi.RT_UPDATE_AUTO])
def test_tlas_update_parity(mode):
    """C4 in small: every transform changes each tick, then rt_update_instances + rt_update_tlas
    (src/scene.rs:167-204).  Result must equal the oracle after the same update, for rebuild and refit."""
    orc, so, gpu, sg = both("c4", 480, 270, num_instances=1500)
    for tick in (1, 2, 5):
        rec_o, rec_g = so.animate(tick), sg.animate(tick)
        orc.update_instances(0, rec_o); orc.update_tlas(mode)
        gpu.update_instances(0, rec_g); gpu.update_tlas(mode)
        want, got = orc.render(so.uniforms(), so.params()), gpu.render(sg.uniforms(), sg.params())
        gpu.stats()
        check_parity(got, want)
    # partial update: one 64-byte record, like the reference's lain rotation
    one = sg.animate(9)[7:8]
    gpu.update_instances(7, one); gpu.update_tlas(mode)
    orc.update_instances(7, so.animate(9)[7:8]); orc.update_tlas(mode)
    check_parity(gpu.render(sg.uniforms(), sg.params()), orc.render(so.uniforms(), so.params()))
    orc.close(); gpu.close()


def test_default_scene_animation_loop():
    """SURVEY 8f-2, the reference's per-frame loop (src/scene.rs:162-204, src/main.rs:917-948): rotate lain, rewrite ONE
    64-byte instance record, TLAS UPDATE in place (refit), frame_index + 1, render with two frames in flight.  Every
    frame must equal the oracle's frame after the same update (the oracle rebuilds from scratch)."""
    import torch

    from ray_tracing_gallery_b200.scene import LAIN_INSTANCE

    orc, so, gpu, sg = both("default", 640, 360)
    fbs = [torch.zeros((360, 640, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
    pending, want = [], {}
    for tick in range(1, 6):
        rec = sg.animate(tick)
        assert not np.array_equal(rec[LAIN_INSTANCE], sg.instances[LAIN_INSTANCE]) and np.array_equal(rec[:2], sg.instances[:2])
        gpu.update_instances(LAIN_INSTANCE, rec[LAIN_INSTANCE:LAIN_INSTANCE + 1])
        gpu.update_tlas(abi.RT_UPDATE_REFIT)
        orc.update_instances(LAIN_INSTANCE, so.animate(tick)[LAIN_INSTANCE:LAIN_INSTANCE + 1])
        orc.update_tlas(abi.RT_UPDATE_REBUILD)
        want[tick] = orc.render(so.uniforms(frame_index=1 + tick), so.params(), want=("rgba8",))["rgba8"]
        if len(pending) == 2:
            slot, b, t = pending.pop(0)
            gpu.wait_frame(slot)
            assert np.mean(np.abs(fbs[b].numpy().astype(int) - want[t].astype(int)).max(axis=2) <= 1) >= 0.999, f"tick {t}"
        b = tick & 1
        pending.append((gpu.render_async(sg.uniforms(frame_index=1 + tick), sg.params(), fbs[b].data_ptr()), b, tick))
    for slot, b, t in pending:
        gpu.wait_frame(slot)
        assert np.mean(np.abs(fbs[b].numpy().astype(int) - want[t].astype(int)).max(axis=2) <= 1) >= 0.999, f"tick {t}"
    # and the full per-pixel bar on the last state
    check_parity(gpu.render(sg.uniforms(frame_index=6), sg.params()), orc.render(so.uniforms(frame_index=6), so.params()), strict_ids=False)
    orc.close(); gpu.close()


def test_double_buffered_tlas_updates_with_frames_in_flight():
    """SURVEY 8b (threading row): instance buffer + TLAS are double-buffered like the reference's PerFrameResources
    (src/command_buffer_recording.rs:22-30).  Updates for frame i+1 are enqueued while frame i is still in flight and go to
    the other set; every frame must show exactly the state it was enqueued with, whatever mix of full / partial writes,
    refit / rebuild, and updates without writes came before it."""
    import torch

    orc, so, gpu, sg = both("c4", 320, 180, num_instances=600)
    n = len(sg.instances)
    fbs = [torch.zeros((180, 320, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
    steps = [  # (tick, first, count, mode)
        (1, 0, n, abi.RT_UPDATE_REFIT), (2, 0, n, abi.RT_UPDATE_REFIT), (3, n - 40, 40, abi.RT_UPDATE_REFIT), (4, 0, n, abi.RT_UPDATE_REBUILD),
        (5, n - 300, 7, abi.RT_UPDATE_REFIT), (5, 0, 0, abi.RT_UPDATE_REFIT), (6, n - 300, 150, abi.RT_UPDATE_REBUILD), (7, 0, n, abi.RT_UPDATE_AUTO),
    ]
    state_o = so.instances.copy()
    pending, want = [], []

    def consume():
        slot, b, k = pending.pop(0)
        gpu.wait_frame(slot)
        assert np.mean(np.abs(fbs[b].numpy().astype(int) - want[k].astype(int)).max(axis=2) <= 1) >= 0.999, f"step {k}"

    for k, (tick, first, count, mode) in enumerate(steps):
        rec = sg.animate(tick)
        if count:
            gpu.update_instances(first, rec[first:first + count])
            state_o[first:first + count] = so.animate(tick)[first:first + count]
            orc.update_instances(0, state_o)
        gpu.update_tlas(mode)
        orc.update_tlas(abi.RT_UPDATE_REBUILD)
        want.append(orc.render(so.uniforms(frame_index=1 + k), so.params(), want=("rgba8",))["rgba8"])
        if len(pending) == 2:
            consume()
        b = k & 1
        pending.append((gpu.render_async(sg.uniforms(frame_index=1 + k), sg.params(), fbs[b].data_ptr()), b, k))
    while pending:
        consume()
    # records written without a TLAS update stay invisible (the acceleration structure keeps its own copy of the transforms)
    before = gpu.render(sg.uniforms(frame_index=9), sg.params())
    gpu.update_instances(0, sg.animate(40))
    after = gpu.render(sg.uniforms(frame_index=9), sg.params())
    for key in ("rgba8", "hit_ids", "ray_counts"):
        assert np.array_equal(before[key], after[key])
    gpu.update_tlas(abi.RT_UPDATE_REFIT)
    orc.update_instances(0, so.animate(40)); orc.update_tlas(abi.RT_UPDATE_REBUILD)
    check_parity(gpu.render(sg.uniforms(frame_index=9), sg.params()), orc.render(so.uniforms(frame_index=9), so.params()))
    gpu.stats()
    orc.close(); gpu.close()


def test_many_instances_tile_parity():
    """C5 in small (50k instances, 16 soft-shadow rays): a tile of the 4K launch against the oracle."""
    orc, so, gpu, sg = both("c5", 3840, 2160, num_instances=50000)
    tile = dict(tile_x0=1700, tile_y0=1200, tile_w=256, tile_h=96)
    want = orc.render(so.uniforms(), so.params(**tile))
    got = gpu.render(sg.uniforms(), sg.params(**tile))
    gpu.stats()
    check_parity(got, want, strict_ids=False)
    orc.close(); gpu.close()


def _decode_node_lines(lines):
    """First 128-byte lines of wide nodes -> (origin (n,3), cell (n,3), lo (n,3,8), hi (n,3,8) plane positions in float64, present (n,8))."""
    n = len(lines)
    origin = lines[:, 0:12].copy().view(np.float32).astype(np.float64)
    exp = lines[:, 12:15].astype(np.int64)
    cell = np.ldexp(1.0, exp - 127)
    meta = lines[:, 24:32]
    q = (lines[:, 32:128].copy().view(np.uint16).astype(np.uint32) << 16).view(np.float32).reshape(n, 6, 8).astype(np.float64)  # bf16 -> value
    assert np.all(q == np.round(q)) and q.min() >= 0 and q.max() <= 255
    lo = origin[:, :, None] + q[:, 0::2, :] * cell[:, :, None]
    hi = origin[:, :, None] + q[:, 1::2, :] * cell[:, :, None]
    return origin, cell, lo, hi, meta != 0


def test_box_test_is_conservative():
    """The packed-bf16 box test of a node visit (trace.cuh node_hit_mask, through rt_debug_box_test) against an exact float64 slab
    test of the same quantised child boxes: a child the ray touches on [tmin, tmax] must ALWAYS be reported, for random rays and for
    the adversarial ones — axis-parallel and nearly parallel directions, rays aimed exactly at box corners and along box faces, origins
    inside the node, origins ten thousand units away.  Also reports how many extra children the padded test lets through."""
    gpu = make_renderer()
    build_scene(gpu, "c2", 64, 64)  # lain's BLAS + a small TLAS
    st = gpu.stats()
    rng = np.random.default_rng(7)
    for tlas, first, count in ((False, 0, min(st.blas_nodes, 400)), (False, max(st.blas_nodes - 300, 0), min(st.blas_nodes, 300)), (True, 0, st.tlas_nodes)):
        _, lines = gpu.box_test(np.array([[0, 0, 0, 0.01, 0, 0, 1, 1e4]], np.float32), first, count, tlas)
        origin, cell, lo, hi, present = _decode_node_lines(lines)
        centre = origin + 127.5 * cell
        extent = 255.0 * cell
        rays = []
        for k in range(count):
            c, e = centre[k], np.maximum(extent[k], 1e-6)
            slots = np.flatnonzero(present[k])
            for _ in range(6):  # random rays through the node's neighbourhood
                o = c + (rng.random(3) - 0.5) * 3.0 * e
                t = c + (rng.random(3) - 0.5) * 1.2 * e
                rays.append((k, o, t - o))
            for s_ in slots[:4]:  # aimed exactly at a corner of a child box; along one of its faces; parallel to an axis through it
                corner = np.array([(lo if rng.random() < 0.5 else hi)[k, a, s_] for a in range(3)])
                o = c + (rng.random(3) - 0.5) * 4.0 * e
                rays.append((k, o, corner - o))
                a = rng.integers(3)
                d = rng.standard_normal(3); d[a] = 0.0
                o2 = corner - d * 2.0
                rays.append((k, o2, d))
                d3 = np.zeros(3); d3[a] = 1.0 if rng.random() < 0.5 else -1.0
                mid = 0.5 * (lo[k, :, s_] + hi[k, :, s_])
                rays.append((k, mid - d3 * 3.0 * e, d3))
                d4 = d3.copy(); d4[(a + 1) % 3] = 1e-9; d4[(a + 2) % 3] = -3e-12
                rays.append((k, mid - d4 * 2.0 * e, d4))
            rays.append((k, c + 0.1 * e * (rng.random(3) - 0.5), rng.standard_normal(3)))           # starts inside
            far = rng.standard_normal(3); far /= np.linalg.norm(far)
            rays.append((k, c - far * 1.0e4, far * (1.0 + 1e-7 * rng.standard_normal())))           # from far away
        ray_arr = np.zeros((len(rays), 8), np.float32)
        for i, (_, o, d) in enumerate(rays):
            d = d / max(np.linalg.norm(d), 1e-30)
            ray_arr[i] = (o[0], o[1], o[2], 0.001, d[0], d[1], d[2], 1.0e4 if i % 3 else 50.0)
        node_of = np.array([k for k, _, _ in rays])
        masks = np.zeros((len(rays), 2), np.uint8)
        for k0 in range(0, count, 64):  # ray x node pairs in blocks of nodes (the call tests every pair)
            sel = np.flatnonzero((node_of >= k0) & (node_of < k0 + 64))
            m, _ = gpu.box_test(ray_arr[sel], first + k0, min(64, count - k0), tlas)
            masks[sel] = m[np.arange(len(sel)), node_of[sel] - k0]
        # exact slab test in float64 on the float32 rays the device saw
        o = ray_arr[:, 0:3].astype(np.float64); d = ray_arr[:, 4:7].astype(np.float64)
        tmin = ray_arr[:, 3].astype(np.float64); tmax = ray_arr[:, 7].astype(np.float64)
        L, H = lo[node_of], hi[node_of]                    # (rays, 3, 8)
        with np.errstate(divide="ignore", invalid="ignore"):
            t0 = (L - o[:, :, None]) / d[:, :, None]
            t1 = (H - o[:, :, None]) / d[:, :, None]
        par = (d == 0.0)[:, :, None] & np.ones((1, 1, 8), bool)
        inside = (o[:, :, None] >= L) & (o[:, :, None] <= H)
        near = np.where(par, np.where(inside, -np.inf, np.inf), np.minimum(t0, t1))
        far_ = np.where(par, np.where(inside, np.inf, -np.inf), np.maximum(t0, t1))
        tn = np.maximum(near.max(axis=1), tmin[:, None]); tf = np.minimum(far_.min(axis=1), tmax[:, None])
        exact = (tn <= tf) & present[node_of] & np.all(L <= H, axis=1)
        for form in (0, 1):
            got = ((masks[:, form, None] >> np.arange(8)) & 1).astype(bool)
            missed = exact & ~got
            assert not missed.any(), f"box test missed {missed.sum()} children (form {form}, tlas {tlas}); first: ray {ray_arr[np.argwhere(missed)[0][0]]}"
            extra = (got & present[node_of] & ~exact).sum()
            print(f"{'TLAS' if tlas else 'BLAS'} nodes {first}..{first + count}: {len(rays)} rays, form {form}: {exact.sum()} children touched, all reported; "
                  f"{extra} reported but not touched ({extra / max(exact.sum(), 1):.1%} of the touched)")
    gpu.close()


def test_full_size_c5_tile_parity():
    """BASELINE configs[4] at its stated size: ALL 1 000 001 instances, 16 soft-shadow rays, a 960x270 tile of the 3840x2160
    launch against the oracle (the whole frame would be ~140 M oracle rays)."""
    orc, so, gpu, sg = both("c5", 3840, 2160)
    assert len(sg.instances) == 1_000_001 and sg.shadow_rays == 16
    tile = dict(tile_x0=1440, tile_y0=1080, tile_w=960, tile_h=270)
    want = orc.render(so.uniforms(), so.params(**tile))
    for pipeline in PIPELINES:
        got = gpu.render(sg.uniforms(), sg.params(pipeline=pipeline, **tile))
        st = gpu.stats()
        ids, within = check_parity(got, want, strict_ids=False)
        print(f"c5 full size, pipeline {pipeline}: {st.num_instances} instances, tile 960x270 of 3840x2160, rays {got['ray_counts'].tolist()}, "
              f"hit IDs {ids:.6%}, radiance within 1e-3 {within:.6%}, PSNR {psnr(got['radiance'], want['radiance']):.1f} dB")
    assert want["ray_counts"][1] > 2_000_000  # the tile really looks at the field
    orc.close(); gpu.close()


def test_full_size_c4_frame_parity_after_refits():
    """BASELINE configs[3] at its stated size: 10 001 instances, every transform updated per frame, TLAS refit (the reference's
    in-place UPDATE) for three ticks, then the whole 3840x2160 frame against the oracle (which rebuilds from scratch)."""
    orc, so, gpu, sg = both("c4", 3840, 2160)
    assert len(sg.instances) == 10_001
    for tick in (1, 2, 3):
        gpu.update_instances(0, sg.animate(tick)); gpu.update_tlas(abi.RT_UPDATE_REFIT)
        gpu.render(sg.uniforms(frame_index=tick), sg.params(), want=("ray_counts",))
    orc.update_instances(0, so.animate(3)); orc.update_tlas(abi.RT_UPDATE_REBUILD)
    want = orc.render(so.uniforms(frame_index=4), so.params())
    got = gpu.render(sg.uniforms(frame_index=4), sg.params())
    st = gpu.stats()
    ids, within = check_parity(got, want, strict_ids=False)
    print(f"c4 full size after 3 refit ticks: {st.num_instances} instances, 3840x2160, rays {got['ray_counts'].tolist()}, hit IDs {ids:.6%}, "
          f"radiance within 1e-3 {within:.6%}, PSNR {psnr(got['radiance'], want['radiance']):.1f} dB, refit {st.last_tlas_ms:.3f} ms")
    orc.close(); gpu.close()


def test_many_images_in_one_warp():
    """72 real images + the built-ins, dealt so that one warp's pixel tile covers several of them, a third alpha-masked
    (per-lane image indices inside the traversal's any-hit): the bounded bindless fetch of shade.cuh against the oracle."""
    import synth_assets

    orc, gpu = make_oracle(), make_renderer()
    so, sg = synth_assets.build_mosaic_scene(orc), synth_assets.build_mosaic_scene(gpu)
    want = orc.render(so.uniforms(), so.params())
    assert len(np.unique(want["hit_ids"][:, :, 0, 1])) > 60  # geometry index = material: most images are on screen
    for pipeline in PIPELINES:
        got = gpu.render(sg.uniforms(), sg.params(pipeline=pipeline, flags=abi.RT_RENDER_COUNTERS))
        st = gpu.stats()
        assert sum(st.anyhit_calls) > 10000
        check_parity(got, want, strict_ids=False)
    orc.close(); gpu.close()


def test_stack_overflow_is_reported():
    """A traversal stack that overflows drops a subtree, so the API must refuse the frame instead of returning a wrong image:
    the same library built with a 3-entry stack (csrc/Makefile libb200rt_stack3.so) renders C3 in a child process."""
    import os
    import subprocess
    import sys

    from ray_tracing_gallery_b200 import native

    lib = os.path.join(os.path.dirname(native.LIB_PATH), "libb200rt_stack3.so")
    if not os.path.exists(lib):
        pytest.skip("libb200rt_stack3.so not built (__graft_entry__.build())")
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from ray_tracing_gallery_b200 import native\n"
        "from ray_tracing_gallery_b200.backend import RtError\n"
        "from ray_tracing_gallery_b200.scene import build_scene\n"
        "gpu = native.Renderer(0); s = build_scene(gpu, 'c3', 640, 360)\n"
        "try:\n"
        "    gpu.render(s.uniforms(), s.params())\n"
        "    print('NO ERROR')\n"
        "except RtError as e:\n"
        "    print('ERR', e)\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, B200RT_LIB=lib), capture_output=True, text=True, timeout=300)
    assert "ERR" in out.stdout and "stack overflowed" in out.stdout, out.stdout + out.stderr


def test_group_paths_on_one_rank():
    """rt_group_* with a group of ONE rank (the 1-GPU box): NCCL comm, broadcast into the staging instance buffer + refit,
    the device-frame path (flags, acquire, readback, release with slot reuse) and the shared page-locked host frame must
    reproduce rt_render bit for bit, frame after frame."""
    from ray_tracing_gallery_b200 import native

    gpu = make_renderer()
    s = build_scene(gpu, "default", 640, 360)
    grp = native.Group(gpu, 1, 0, native.group_unique_id(), 640, 360)
    p = s.params()
    assert grp.partition(p) == 360 and p.strip_count == 0
    seq = 0
    for tick in range(1, 8):  # more frames than slots: release / reuse is exercised
        rec = np.ascontiguousarray(s.animate(tick)[2:3])
        grp.update_instances(2, 1, rec.ctypes.data, abi.RT_UPDATE_REFIT)
        u = s.uniforms(frame_index=tick)
        want = gpu.render(u, s.params(), want=("rgba8", "ray_counts"))
        seq += 1
        grp.render_device(seq, u, s.params())
        got_dev = grp.readback(seq)
        grp.release(seq)
        seq += 1
        grp.render_host(seq, u, s.params())
        got_host, counts = grp.acquire_host(seq)
        assert np.array_equal(got_dev, want["rgba8"]) and np.array_equal(got_host, want["rgba8"]), f"tick {tick}"
        assert list(counts) == want["ray_counts"].tolist()
        grp.release(seq)
    grp.barrier()
    grp.close(); gpu.close()


@pytest.mark.parametrize("cfg,kw,size", [("c5", dict(num_instances=60000), (960, 540)), ("c4", dict(num_instances=3000), (640, 360)), ("default", {}, (640, 360))])
def test_group_sharded_tlas_build_on_one_rank(cfg, kw, size):
    """SURVEY 8f-4 on the 1-GPU box: rt_group_build_tlas with the sharded path forced for a group of one rank — key histogram,
    range selection, treelet build at its place in the global leaf order, re-linking into the assembled layout, top node.
    The frame must equal the frame of the ordinary build bit for bit, and a refit of the assembled tree must still match the
    oracle."""
    from ray_tracing_gallery_b200 import native

    orc, so, gpu, sg = both(cfg, *size, **kw)
    before = gpu.render(sg.uniforms(), sg.params())
    nodes_before = gpu.stats().tlas_nodes
    grp = native.Group(gpu, 1, 0, native.group_unique_id(), *size)
    grp.build_tlas(sg.instances, force_sharded=True)
    after = gpu.render(sg.uniforms(), sg.params())
    st = gpu.stats()
    for key in ("rgba8", "hit_ids", "ray_counts", "radiance"):
        assert np.array_equal(before[key], after[key]), key
    assert st.tlas_nodes >= 2 and abs(int(st.tlas_nodes) - int(nodes_before)) <= max(8, nodes_before // 4)
    check_parity(after, orc.render(so.uniforms(), so.params()), strict_ids=False)
    if sg.dynamic:  # the assembled tree takes refits like any other
        rec_g, rec_o = sg.animate(3), so.animate(3)
        if cfg == "default":
            grp.update_instances(2, 1, np.ascontiguousarray(rec_g[2:3]).ctypes.data, abi.RT_UPDATE_REFIT)
        else:
            a = np.ascontiguousarray(rec_g)
            grp.update_instances(0, len(a), a.ctypes.data, abi.RT_UPDATE_REFIT)
        orc.update_instances(0, rec_o); orc.update_tlas(abi.RT_UPDATE_REBUILD)
        check_parity(gpu.render(sg.uniforms(), sg.params()), orc.render(so.uniforms(), so.params()), strict_ids=False)
        gpu.stats()
    grp.close(); orc.close(); gpu.close()


def test_group_two_ranks_from_cpp():
    """Two GPUs driven from the C++ host through the C ABI alone (host_cpp/rt_group_demo): NCCL broadcast of the animated
    instance record, peer-memory device frame and shared host frame, each checked bit for bit against one GPU."""
    import json
    import os
    import subprocess

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ray_tracing_gallery_b200", "host_cpp", "rt_group_demo")
    if not os.path.exists(exe):
        pytest.skip("rt_group_demo not built (__graft_entry__.build())")
    for config, size, extra in (("default", (640, 360), []), ("c3", (1001, 563), []),   # 563 rows: the last strip is partial
                                ("c3", (640, 360), ["--sharded-build", "--instances", "20000"])):  # each rank builds half of the TLAS
        out = subprocess.run([exe, "--ranks", "2", "--config", config, "--width", str(size[0]), "--height", str(size[1]), "--frames", "6"] + extra,
                             capture_output=True, text=True, timeout=400)
        assert out.returncode == 0, out.stdout + out.stderr
        line = json.loads(out.stdout.strip().splitlines()[-1])
        assert line["ok"] and line["ranks"] == 2 and line["mismatched_frames"] == 0


def test_shadow_denoise_hook():
    """SURVEY 8f-4 (readme.md:17-20, "shadow denoising"): the hook sits between segment 0's shadow rays and its resolve.  An
    identity hook must leave the frame bit-identical; a failing hook must fail the frame; the library's cross-bilateral
    filter must bring a 2-sample frame closer to the 64-sample frame than the raw 2-sample frame is."""
    import ctypes as C

    gpu = make_renderer()
    s = build_scene(gpu, "c2", 960, 540)
    s.shadow_rays = 2
    raw = gpu.render(s.uniforms(), s.params())
    seen = {}

    def identity(user, stream, buffers):
        b = buffers.contents
        seen.update(width=b.width, rows=b.rows, n=b.shadow_rays, frame=b.frame_index, ptrs=(b.sun_factor, b.position_nol), stream=stream)
        return 0

    gpu.set_denoise_hook(identity)
    same = gpu.render(s.uniforms(), s.params())
    assert seen["width"] == 960 and seen["rows"] == 540 and seen["n"] == 2 and all(seen["ptrs"]) and seen["frame"] == s.frame_index
    for key in ("rgba8", "radiance", "hit_ids", "ray_counts"):
        assert np.array_equal(raw[key], same[key]), key
    gpu.set_denoise_hook(lambda user, stream, buffers: 7)
    with pytest.raises(RtError, match="denoise hook returned 7"):
        gpu.render(s.uniforms(), s.params())
    sigma = C.c_float(0.3)
    gpu.set_denoise_hook("bilateral", C.addressof(sigma))
    den = gpu.render(s.uniforms(), s.params())
    gpu.set_denoise_hook(None)
    again = gpu.render(s.uniforms(), s.params())
    assert np.array_equal(again["radiance"], raw["radiance"])
    s.shadow_rays = 64
    ref = gpu.render(s.uniforms(), s.params())["radiance"].astype(np.float64)
    mse_raw = np.mean((raw["radiance"] - ref) ** 2)
    mse_den = np.mean((den["radiance"] - ref) ** 2)
    print(f"shadow denoise: MSE against the 64-sample frame, raw 2 samples {mse_raw:.3e}, bilateral {mse_den:.3e}")
    assert mse_den < 0.6 * mse_raw
    assert np.array_equal(den["hit_ids"], raw["hit_ids"]) and np.array_equal(den["ray_counts"], raw["ray_counts"])
    gpu.close()


def test_show_heatmap_frame():
    """SURVEY 8f-3, Uniforms.show_heatmap (lib.rs:120-124, 174-186; heatmap.rs): the frame shows
    heatmap_temperature(clock ticks of the pixel's ray-gen invocation / heatmap_scale) + 1e-6 * colour.  The clock is the
    GPU's own, so the check feeds the exported cost_cycles through the oracle's heatmap_pixel; ray work is unchanged."""
    orc, gpu = make_oracle(), make_renderer()
    s = build_scene(gpu, "default", 320, 180)
    plain = gpu.render(s.uniforms(), s.params())
    u = s.uniforms()
    u.show_heatmap = 1
    want_all = ("rgba8", "radiance", "hit_ids", "ray_counts", "cost_cycles")
    for pipeline, scale in ((abi.RT_PIPELINE_WAVEFRONT, 0.0), (abi.RT_PIPELINE_MEGAKERNEL, 200000.0)):
        p = s.params(pipeline=pipeline)
        p.heatmap_scale = scale
        heat = gpu.render(u, p, want=want_all)
        gpu.stats()
        assert np.array_equal(heat["hit_ids"], plain["hit_ids"]) and np.array_equal(heat["ray_counts"], plain["ray_counts"])
        cost = heat["cost_cycles"]
        assert cost.min() > 0
        hit = plain["hit_ids"][:, :, 0, 0] != abi.MISS_ID
        assert cost[hit].mean() > 1.5 * cost[~hit].mean()  # shaded pixels also run the hit shader and its shadow rays
        ys, xs = np.meshgrid(np.arange(0, 180, 7), np.arange(0, 320, 11), indexing="ij")
        for y, x in zip(ys.ravel(), xs.ravel()):
            want = orc.heatmap_pixel(cost[y, x], plain["radiance"][y, x], scale)
            assert np.allclose(heat["radiance"][y, x], want, rtol=0, atol=1e-6, equal_nan=True), (y, x, cost[y, x])
            enc = [orc.unorm8(orc.linear_to_srgb(c)) for c in want]
            assert np.abs(heat["rgba8"][y, x][:3].astype(int) - np.asarray(enc)).max() <= 1 and heat["rgba8"][y, x][3] == 255
    # a plain frame does not touch the cost output
    assert not np.any(gpu.render(s.uniforms(), s.params(), want=("cost_cycles",))["cost_cycles"])
    orc.close(); gpu.close()


def test_device_tables_keep_the_reference_layout():
    gpu = make_renderer()
    s = build_scene(gpu, "default", 64, 36)
    pc = gpu.push_constants()
    assert pc.model_info and pc.uniforms and pc.acceleration_structure
    for name, (mid, handle, arrays) in s.models.items():
        info, geoms = gpu.read_model_info(mid)
        assert info.position_buffer_address and info.normal_buffer_address and info.uv_buffer_address and info.geometry_info_address
        for g, gi in zip(arrays.geometries, geoms):
            assert gi.index_buffer_address
            assert (gi.images.diffuse_image_index, gi.images.metallic_roughness_image_index, gi.images.normal_map_image_index) == (
                g.diffuse_image_index, g.metallic_roughness_image_index, g.normal_map_image_index)
    st = gpu.stats()
    assert st.num_instances == 105 and st.num_triangles == 2 + 2304 + 45448 + 2 and st.tlas_nodes >= 1
    gpu.close()


def test_edge_cases():
    orc, gpu = make_oracle(), make_renderer()
    setups = []
    for b in (orc, gpu):
        push_builtin_images(b)
        pid, ph, _ = load_model(b, "plane.glb", 0)
        tid, th, _ = load_model(b, "tori.glb", 1)
        setups.append((pid, ph, tid, th))
    from ray_tracing_gallery_b200.scene import Camera, SceneSetup, Sun

    def scene(b, ids, inst):
        s = SceneSetup("edge", inst, Camera(), Sun(), 160, 90, shadow_rays=2, sun_radius=0.05)
        b.build_tlas(inst)
        return s

    def records(ids):
        pid, ph, tid, th = ids
        hidden = make_instance(mat_translation(0, 1, 0), tid, th, abi.RT_HIT_TEXTURED)
        hidden["custom_index_and_mask"] = tid  # visibility mask 0: never hit
        singular = make_instance(mat_scale(0.0), tid, th, abi.RT_HIT_TEXTURED)
        bad_handle = make_instance(mat_identity(), tid, 0x1234, abi.RT_HIT_TEXTURED)
        out_of_table = make_instance(mat_translation(3, 1, 2), tid, th, 7)  # hit group 7: payload untouched -> black
        return np.stack([make_instance(mat_scale(10.0), pid, ph, abi.RT_HIT_TEXTURED), hidden, singular, bad_handle, out_of_table])

    so, sg = scene(orc, setups[0], records(setups[0])), scene(gpu, setups[1], records(setups[1]))
    want, got = orc.render(so.uniforms(), so.params()), gpu.render(sg.uniforms(), sg.params())
    check_parity(got, want)
    assert set(np.unique(got["hit_ids"][:, :, 0, 0]).tolist()) <= {0, 4, abi.MISS_ID}
    # empty TLAS: every pixel is sky
    for b in (orc, gpu):
        b.build_tlas(np.zeros(0, abi.INSTANCE_DTYPE))
    got = gpu.render(sg.uniforms(), sg.params())
    assert np.array_equal(got["rgba8"], orc.render(so.uniforms(), so.params())["rgba8"])
    assert np.all(got["hit_ids"] == abi.MISS_ID) and got["ray_counts"].tolist() == [160 * 90, 0]
    # ragged sizes: width/height not multiples of the 8x4 warp tile, max_segments 1, 5 shadow rays
    for b, ids in ((orc, setups[0]), (gpu, setups[1])):
        b.build_tlas(records(ids))
    po, pg = so.params(width=75, height=41, max_segments=1, shadow_rays=5), sg.params(width=75, height=41, max_segments=1, shadow_rays=5)
    check_parity(gpu.render(sg.uniforms(width=75, height=41), pg), orc.render(so.uniforms(width=75, height=41), po))
    orc.close(); gpu.close()


def test_error_behaviour():
    gpu = make_renderer()
    push_builtin_images(gpu)
    from ray_tracing_gallery_b200.scene import Camera, Sun, make_uniforms

    u = make_uniforms(Camera(), Sun(), 64, 36, 0.05, 1)
    p = abi.RtRenderParams(width=64, height=36, max_segments=3, shadow_rays=2)
    with pytest.raises(RtError, match="before rt_build_tlas"):
        gpu.render(u, p)
    with pytest.raises(RtError):
        gpu.update_tlas(abi.RT_UPDATE_REBUILD)
    gpu.build_tlas(np.zeros(0, abi.INSTANCE_DTYPE))
    with pytest.raises(RtError, match="shadow_rays"):
        gpu.render(u, abi.RtRenderParams(width=64, height=36, max_segments=3, shadow_rays=0))
    with pytest.raises(RtError, match="tile outside"):
        gpu.render(u, abi.RtRenderParams(width=64, height=36, max_segments=3, shadow_rays=1, tile_x0=60, tile_w=10, tile_h=4))
    with pytest.raises(RtError, match="range outside"):
        gpu.update_instances(0, np.zeros(1, abi.INSTANCE_DTYPE))
    assert gpu.lib.rt_render(gpu.ctx, None, None, None) == -1
    gpu.render(u, p)  # the context stays usable after errors
    for _ in range(130):
        try:
            gpu.push_image(np.zeros((1, 1, 4), np.uint8), abi.RT_FORMAT_RGBA8_UNORM, False)
        except RtError as e:
            assert "table full" in str(e)
            break
    else:
        raise AssertionError("image table must be capped at 128 (src/main.rs:44)")
    gpu.close()


def test_device_output_path_and_readback():
    import torch

    gpu = make_renderer()
    s = build_scene(gpu, "c1", 320, 180)
    host = gpu.render(s.uniforms(), s.params())
    assert np.array_equal(gpu.readback(180, 320), host["rgba8"])
    gpu.set_stream(torch.cuda.current_stream().cuda_stream)
    fb = torch.zeros((180, 320, 4), dtype=torch.uint8, device="cuda")
    rad = torch.zeros((180, 320, 3), dtype=torch.float32, device="cuda")
    rays = torch.zeros(2, dtype=torch.int64, device="cuda")
    before = gpu.lib.rt_kernel_launches()
    gpu.render_device(s.uniforms(), s.params(), rgba8=fb.data_ptr(), radiance=rad.data_ptr(), ray_counts=rays.data_ptr())
    torch.cuda.synchronize()
    assert gpu.lib.rt_kernel_launches() > before
    assert np.array_equal(fb.cpu().numpy(), host["rgba8"])
    assert np.array_equal(rad.cpu().numpy().view(np.uint32), host["radiance"].view(np.uint32))
    assert rays.cpu().numpy().astype(np.uint64).tolist() == host["ray_counts"].tolist()
    gpu.close()


def test_two_frames_in_flight_match_blocking_renders():
    """rt_render_async / rt_wait_frame (the reference's two PerFrameResources + fences, src/main.rs:917-928):
    frames rendered through the two slots, copies overlapped with the next render, equal the blocking rt_render frames."""
    import torch

    gpu = make_renderer()
    s = build_scene(gpu, "c2", 480, 270)
    want = [gpu.render(s.uniforms(frame_index=1 + i), s.params(), want=("rgba8", "ray_counts")) for i in range(5)]
    fbs = [torch.zeros((270, 480, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
    rcs = [torch.zeros(2, dtype=torch.int64).pin_memory() for _ in range(2)]
    pending = []

    def consume():
        slot, b, i = pending.pop(0)
        gpu.wait_frame(slot)
        assert np.array_equal(fbs[b].numpy(), want[i]["rgba8"]), f"frame {i}"
        assert rcs[b].numpy().astype(np.uint64).tolist() == want[i]["ray_counts"].tolist()

    for i in range(5):
        b = i & 1
        if len(pending) == 2:
            consume()  # the host buffer of slot b is free again
        slot = gpu.render_async(s.uniforms(frame_index=1 + i), s.params(), fbs[b].data_ptr(), rcs[b].data_ptr())
        assert slot == b
        pending.append((slot, b, i))
    while pending:
        consume()
    with pytest.raises(RtError):
        gpu.wait_frame(2)
    gpu.sync()
    gpu.close()


def test_image_rows_output_lets_ranks_share_one_frame():
    """RT_RENDER_OUTPUT_IMAGE_ROWS: each strip share stores its rows in place into ONE [H][W] frame (what the ranks of a
    multi-GPU frame do into rank 0's peer-mapped frame, dist.SharedFrame); no gather, no de-interleave."""
    import torch

    gpu = make_renderer()
    gpu.set_stream(torch.cuda.current_stream().cuda_stream)
    W, H = 200, 100   # 13 strips, the last one partial
    s = build_scene(gpu, "default", W, H)
    full = gpu.render(s.uniforms(), s.params(), want=("rgba8", "radiance", "ray_counts"))
    frame = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    rad = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    rays = torch.zeros(2, dtype=torch.int64, device="cuda")
    total = np.zeros(2, np.uint64)
    for r in range(3):
        part = Partition.make(W, H, 3, r)
        gpu.render_device(s.uniforms(), part.apply(s.params(flags=abi.RT_RENDER_OUTPUT_IMAGE_ROWS)), rgba8=frame.data_ptr(),
                          radiance=rad.data_ptr(), ray_counts=rays.data_ptr())
        torch.cuda.synchronize()
        total += rays.cpu().numpy().astype(np.uint64)
    assert np.array_equal(frame.cpu().numpy(), full["rgba8"])
    assert np.array_equal(rad.cpu().numpy().view(np.uint32), full["radiance"].view(np.uint32))
    assert np.array_equal(total, full["ray_counts"])
    # a tile with an origin: rows are relative to the tile
    tile = torch.zeros((40, 64, 4), dtype=torch.uint8, device="cuda")
    for r in range(2):
        part = Partition.make(W, H, 2, r)
        gpu.render_device(s.uniforms(), part.apply(s.params(flags=abi.RT_RENDER_OUTPUT_IMAGE_ROWS, tile_x0=30, tile_y0=20, tile_w=64, tile_h=40)),
                          rgba8=tile.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(tile.cpu().numpy(), full["rgba8"][20:60, 30:94])
    with pytest.raises(RtError):
        gpu.render(s.uniforms(), s.params(flags=abi.RT_RENDER_OUTPUT_IMAGE_ROWS))
    gpu.close()


def test_device_slots_render_concurrently_on_caller_streams():
    """rt_render_device_slot: two frames on two caller-owned streams, each with the private queues of its slot, equal
    the frames rendered one after the other; a TLAS update issued afterwards waits for both."""
    import torch

    gpu = make_renderer()
    s = build_scene(gpu, "c3", 480, 270)
    want = [gpu.render(s.uniforms(frame_index=1 + i), s.params(), want=("rgba8", "ray_counts")) for i in range(4)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    fbs = [torch.zeros((270, 480, 4), dtype=torch.uint8, device="cuda") for _ in range(4)]
    rcs = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(4)]
    for i in range(4):  # frames 0,2 on slot/stream 0, frames 1,3 on slot/stream 1: same-slot frames are ordered by their stream
        gpu.render_device_slot(i & 1, streams[i & 1].cuda_stream, s.uniforms(frame_index=1 + i), s.params(), rgba8=fbs[i].data_ptr(),
                               ray_counts=rcs[i].data_ptr())
    gpu.update_instances(0, s.instances)        # ordered after the four frames by the library
    gpu.update_tlas(abi.RT_UPDATE_REBUILD)
    torch.cuda.synchronize()
    for i in range(4):
        assert np.array_equal(fbs[i].cpu().numpy(), want[i]["rgba8"]), f"frame {i}"
        assert rcs[i].cpu().numpy().astype(np.uint64).tolist() == want[i]["ray_counts"].tolist()
    after = gpu.render(s.uniforms(frame_index=1), s.params(), want=("rgba8",))
    assert np.array_equal(after["rgba8"], want[0]["rgba8"])
    with pytest.raises(RtError):
        gpu.render_device_slot(2, 0, s.uniforms(), s.params())
    gpu.close()

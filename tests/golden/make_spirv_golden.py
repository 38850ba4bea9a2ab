"""Golden vectors made by RUNNING THE REFERENCE'S OWN SHIPPED SHADERS (/root/reference/shaders/*.spv) on this CPU through
tests/spirv_interp.py + tests/spirv_pipeline.py.  Run in the build container (where /root/reference exists):

    python tests/golden/make_spirv_golden.py

writes tests/golden/spirv_<scene>.npz: for a strided grid of pixels of each scene, what `ray_generation.spv` wrote to
the storage image (sRGB-encoded vec4 before the UNORM8 conversion), the payload colour after the last segment, the
trace calls issued, the first-segment hit, and — for the any-hit stage — the ignore decisions on seeded fence
candidates.  Intersections (the driver's part of OpTraceRayKHR) come from the oracle's trace.

The .npz files travel with the repo, so `tests/test_spirv_pin.py` can hold the oracle (any box) and the CUDA path
(`-m gpu`, the B200 box where /root/reference does not exist) to the reference's compiled arithmetic.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from spirv_scenes import SCENES, build, run_pixels, anyhit_candidates  # noqa: E402
from spirv_pipeline import RecordingBackend, RefPipeline  # noqa: E402


def main():
    from oracle.binding import Oracle

    for name in SCENES:
        orc = Oracle()
        rec = RecordingBackend(orc)
        setup, width, height, stride = build(rec, name)
        pipe = RefPipeline(rec, lambda o, d, tmin, tmax, a: orc.trace(o, d, tmin, tmax, a))
        out = run_pixels(pipe, setup, width, height, stride)
        if name == "c3":
            cand = anyhit_candidates(setup, rec)
            out["anyhit_candidates"] = cand
            out["anyhit_ignored"] = np.array([pipe.any_hit_ignores((0, 0), (0, 0, 0), (0, 0, 1), 0.01, (int(c[0]), int(c[1]), int(c[2])), (1.0, c[3], c[4]))
                                              for c in cand], np.uint8)
        if name == "c1":
            # show_heatmap (lib.rs:120-124, 174-186): the two OpReadClockKHR results are the callback's, so the pixel is
            # heatmap_temperature(delta / 1e6) + 1e-6 * colour for a chosen delta
            u = setup.uniforms()
            u.show_heatmap = 1
            pipe.set_uniforms(u, width, height)
            # (delta >= 1 000 000 saturates heat to 1.0 and the compiled stage indexes colours[10]: rust-gpu's bounds-check panic, an
            #  endless loop in the module — tests/test_spirv_pin.py::test_saturated_heat_never_returns_in_the_reference)
            deltas = [0, 1, 99_999, 100_000, 250_000, 499_999, 650_000, 800_000, 999_999, 123_456, 777_777, 901_000, 333_333, 50_000]
            heat = []
            for k, dt in enumerate(deltas):
                x, y = (7 * k + 3) % width, (5 * k + 20) % height
                pipe.clock = [1000, 1000 + dt]
                texel, log = pipe.pixel(x, y)
                heat.append([x, y, dt, *texel, *[e for e in log if not e["shadow"]][-1]["payload"][0]])
            out["heat"] = np.array(heat, np.float64)
        path = os.path.join(HERE, f"spirv_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {len(out['xy'])} pixels, {int(out['n_primary'].sum())} ray-gen segments, {int(out['n_shadow'].sum())} shadow rays, "
              f"{pipe.steps} SPIR-V instructions executed -> {path} ({os.path.getsize(path)} bytes)")
        orc.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — Mrays/s (primary + shadow) and ms/frame of the frame hot path.

  python bench.py --gpus N --steps K --warmup W [--workload c1..c5] [--pipeline wavefront|mega]
  python bench.py --impl reference ...     # the CPU restatement of the reference shaders (oracle)

A step is one frame of the workload: (TLAS update, when the scene is dynamic) + one
`cmd_trace_rays(width, height)` + (N > 1) the gather of the framebuffer strips to rank 0.
Default workload: BASELINE.json configs[1] — lain.glb textured PBR with 4 blue-noise soft-shadow
rays per hit at 1920x1080 (the config the metric is quoted on that fits one GPU).
One JSON line on rank 0.  Timing: CUDA events per step on the launching stream, L2 flushed between
timed steps, barrier + synchronize around the timed region, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


import numpy as np  # noqa: E402

METRIC = "Mrays/s (primary+shadow)"
UNIT = "Mrays/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5", "default"])
    ap.add_argument("--pipeline", default="wavefront", choices=["wavefront", "mega"])
    ap.add_argument("--instances", type=int, default=0, help="override the instance count of c3/c4/c5")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--update-mode", default="refit", choices=["rebuild", "refit"],
                    help="dynamic scenes: refit = the reference's in-place UPDATE (src/util_structs.rs:309), rebuild = full LBVH build")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mg-overlap", default="auto", choices=["auto", "on", "off"],
                    help="N > 1 e2e: alternate frames between two streams / frame slots (rt_render_device_slot)")
    ap.add_argument("--gather", default="auto", choices=["auto", "peer", "nccl"], help="N > 1: how the frame reaches rank 0")
    return ap.parse_args()


def workload_config(setup, args, extra=None):
    cfg = {
        "workload": f"{setup.name}: {setup.description}",
        "resolution": f"{setup.width}x{setup.height}",
        "shadow_rays_per_hit": setup.shadow_rays,
        "sun_radius": setup.sun_radius,
        "max_segments": setup.max_segments,
        "instances": int(len(setup.instances)),
        "frame_index": "1..K (animated blue noise)" if setup.sun_radius > 0 else 1,
    }
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(gpu_index), "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def wait_first_sample(self, timeout=3.0):
        """nvidia-smi needs a moment to start; the frames are short, so wait until it is sampling before loading the GPU."""
        t0 = time.time()
        while self.proc is not None and time.time() - t0 < timeout:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                pass
            time.sleep(0.02)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, busy = [], [], set(), []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2]))
                except ValueError:
                    continue
                try:
                    busy.append(float(f[9]) > 0)
                except (ValueError, IndexError):
                    busy.append(True)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            loaded = [c for c, b in zip(sm, busy) if b] or sm  # median over the samples taken under load
            out.update(sm_mhz=float(np.median(loaded)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------- CPU arm
def oracle_sample(args, setup_builder, threads=0, steps=1, warmup=0):
    """Time the CPU restatement (oracle/) on a bounded sample of the workload.  Returns (Mrays/s, info)."""
    from oracle.binding import Oracle

    orc = Oracle(threads=threads)
    s = setup_builder(orc)
    # bounded sample: the full frame when it finishes in seconds, else a centred tile of the same launch
    tile = {}
    if s.name in ("c4", "c5"):
        tw, th = 960, 270
        tile = dict(tile_x0=(s.width - tw) // 2, tile_y0=(s.height - th) // 2 + s.height // 8, tile_w=tw, tile_h=th)
    p = s.params(**tile)
    rays, secs = 0, 0.0
    for i in range(warmup + steps):
        u = s.uniforms(frame_index=1 + i)
        t0 = time.perf_counter()
        r = orc.render(u, p, want=("rgba8", "ray_counts"))
        dt = time.perf_counter() - t0
        if i >= warmup:
            rays += int(r["ray_counts"].sum())
            secs += dt
    cores = orc.threads
    orc.close()
    sample = (f"{steps} frame(s) of {s.name} at {s.width}x{s.height}" +
              (f", tile {tile['tile_w']}x{tile['tile_h']} of the launch" if tile else ", full frame") + f", {rays} rays in {secs:.2f} s")
    return rays / secs / 1e6, {"cores": cores, "sample": sample, "ms_per_step": secs / steps * 1e3, "rays": rays}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs on rank 0 only
    from ray_tracing_gallery_b200.scene import build_scene

    holder = {}

    def builder(backend):
        holder["s"] = build_scene(backend, args.workload, args.width or None, args.height or None, num_instances=args.instances or None)
        return holder["s"]

    # each step is one frame on all host threads; the step count is bounded so that the arm ends within a few minutes
    cap = 60 if args.workload in ("c1", "c2", "c3", "default") else 2
    steps = max(1, min(args.steps, cap))
    warm = max(0, min(args.warmup, 2))
    value, info = oracle_sample(args, builder, steps=steps, warmup=warm)
    s = holder["s"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic: reference assets (glb/png), seeded instance transforms",
        "config": workload_config(s, args, {"note": "CPU restatement of the reference shaders (oracle/); the reference itself needs Rust + "
                                            "Vulkan ray tracing (lavapipe / host rust-gpu unavailable: no Vulkan ICD, no Rust toolchain)"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": info["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------- GPU arm
KERNELS = ["trace", "prep", "shadow", "resolve", "mega", "tail"]
KERNEL_DESC = {"trace": "k_trace (ray-gen + closest-hit traversal)", "prep": "k_prep (triangle fetch, textures, BRDF terms)",
               "shadow": "k_shadow (blue-noise shadow rays, first-hit traversal)", "resolve": "k_resolve (sun factor, sRGB store)",
               "mega": "k_mega (one thread per pixel)", "tail": "k_tail (resolve + bounce segments, cooperative)"}


def algorithmic_bytes(st, setup, pixels):
    """Algorithmic bytes of one frame per kernel (DESIGN.md 'Roofline accounting').

    Traversal: 80 B read per node visit (the five 16-byte words the box test needs), 64 B per instance entered,
    48 B per triangle tested, 12 + 24 + 16 B per any-hit call (indices, 3 uvs, one bilinear tap)."""
    n = setup.shadow_rays
    trav = [80 * st.nodes_visited[k] + 64 * st.instances_entered[k] + 48 * st.triangles_tested[k] + 52 * st.anyhit_calls[k] for k in (0, 1)]
    hits = st.textured_hits
    bounces = st.primary_rays - pixels
    # k_trace: + ray queue (32 B written and read per bounce), mirror shading (12 B indices + 36 B normals + 64 B instance),
    #          48 B hit record per textured hit, 4 B framebuffer per pixel finished here
    trace_b = trav[0] + 64 * bounces + 112 * bounces + 48 * hits + 4 * (pixels - hits)
    # k_prep: hit record 48 B in, ModelInfo 32 + GeometryInfo 24 + indices 12 + 3 vertices x 32, instance transform 48 + inverse 64,
    #         two bilinear taps (4 texels x 4 B each), 64 B record out
    prep_b = hits * (48 + 32 + 24 + 12 + 96 + 48 + 64 + 2 * 16 + 64)
    # k_shadow: per ray 16 B of the hit record (origin + launch id), 16 B of the direction table, 4 B atomic on `lit`
    shadow_b = trav[1] + st.shadow_rays * (16 + 16 + 4)
    # resolve phase (in k_tail): 52 B of the record, 4 B framebuffer
    resolve_b = hits * (52 + 4)
    mega_b = trav[0] + trav[1] + 112 * bounces + hits * (32 + 24 + 12 + 96 + 48 + 64 + 2 * 16 + 8 * n) + 4 * pixels
    return {"trace": int(trace_b), "prep": int(prep_b), "shadow": int(shadow_b), "resolve": int(resolve_b), "mega": int(mega_b),
            "tail": int(resolve_b)}  # k_tail = segment 0's resolve + bounce segments; counters are whole-frame, so the bounce
    #                                  segments' traversal is accounted under "trace"/"shadow"


def run_ours(args):
    import torch
    import torch.distributed as dist

    from ray_tracing_gallery_b200 import abi, native
    from ray_tracing_gallery_b200.dist import Partition, SharedFrame, broadcast_instances, deinterleave_into
    from ray_tracing_gallery_b200.scene import build_scene

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    args.warmup = max(args.warmup, 3)
    torch.cuda.synchronize()

    gpu = native.Renderer(local_rank)  # raises if libb200rt.so is missing: there is no fallback
    stream = torch.cuda.Stream(device=dev)  # one stream for the library's kernels, NCCL and the timing events
    torch.cuda.set_stream(stream)
    gpu.set_stream(stream.cuda_stream)
    s = build_scene(gpu, args.workload, args.width or None, args.height or None, num_instances=args.instances or None)
    W, H = s.width, s.height
    part = Partition.make(W, H, world, rank)
    pipeline = abi.RT_PIPELINE_MEGAKERNEL if args.pipeline == "mega" else abi.RT_PIPELINE_WAVEFRONT
    update_mode = abi.RT_UPDATE_REFIT if args.update_mode == "refit" else abi.RT_UPDATE_REBUILD
    rows = part.local_rows      # rows this rank renders
    slab_rows = part.max_rows   # common slab size of the all-gather (shares differ by at most one strip)
    pixels = rows * W

    fb = torch.zeros((slab_rows, W, 4), dtype=torch.uint8, device=dev)
    gathered = torch.zeros((world, slab_rows, W, 4), dtype=torch.uint8, device=dev) if world > 1 else None
    mg_final = [torch.zeros((H, W, 4), dtype=torch.uint8, device=dev) for _ in range(2)] if (world > 1 and rank == 0) else [None, None]
    rays_dev = torch.zeros(2, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    inst_dev = torch.zeros(len(s.instances) * 64, dtype=torch.uint8, device=dev) if s.dynamic else None
    host_fb = torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory()
    host_rays = torch.zeros(2, dtype=torch.int64).pin_memory()

    def params(flags=0):
        return part.apply(s.params(pipeline=pipeline, flags=flags))

    def frame_inputs(i):
        return s.uniforms(frame_index=1 + i)

    # Dynamic scene: the animated instance records of every step are produced BEFORE the timed region (they are the
    # step's input): pinned host copies for the e2e arm, device copies (rank 0) for the device-resident arm.
    n_frames = args.warmup + max(args.steps, 10) + 4
    rec_pinned, rec_dev = [], []
    if s.dynamic and rank == 0:
        for i in range(n_frames):
            t = torch.from_numpy(s.animate(i + 1).view(np.uint8).reshape(-1).copy()).pin_memory()
            rec_pinned.append(t)
            rec_dev.append(t.to(dev))

    def update_scene_device(i, from_host=False):
        """Rank 0 owns the new records; the NCCL broadcast lands them in the buffer the TLAS builder reads."""
        if rank == 0:
            inst_dev.copy_(rec_pinned[i % n_frames] if from_host else rec_dev[i % n_frames], non_blocking=True)
        if world > 1:
            broadcast_instances(inst_dev, 0)
        gpu.update_instances_device(0, len(s.instances), inst_dev.data_ptr())
        gpu.update_tlas(update_mode)

    rays_steps = torch.zeros((args.steps, 2), dtype=torch.int64, device=dev)  # per timed step {ray-gen segments, shadow rays}

    # N > 1: how the frame reaches rank 0.  "peer": every rank's render kernels store their rows straight into rank 0's
    # frame through NVLink peer memory (torch symmetric memory + RT_RENDER_OUTPUT_IMAGE_ROWS), then one cross-rank
    # barrier.  "nccl": compact slabs, all_gather_into_tensor, de-interleave on rank 0.
    shared, gather_path = None, "none"
    if world > 1:
        gather_path = "nccl all-gather + de-interleave"
        if args.gather in ("auto", "peer"):
            try:
                shared = SharedFrame(W, H, dev, slots=4)
                gather_path = "peer stores into rank 0's frame (NVLink symmetric memory) + barrier"
            except Exception as e:  # noqa: BLE001 - any failure of the optional path falls back to NCCL
                if args.gather == "peer":
                    raise
                print(f"[bench] SharedFrame unavailable ({type(e).__name__}: {e}); using NCCL all-gather", file=sys.stderr)

    def render_and_collect(i, flags, counts, slot, before_barrier=None):
        """One frame on this rank + whatever brings it to rank 0 (device side only)."""
        if shared is not None:
            gpu.render_device(frame_inputs(i), params(flags | abi.RT_RENDER_OUTPUT_IMAGE_ROWS), rgba8=shared.target_ptr(slot),
                              ray_counts=counts.data_ptr())
            if before_barrier is not None:
                stream.wait_event(before_barrier)
            shared.barrier()
            return shared.frame(slot) if rank == 0 else None
        gpu.render_device(frame_inputs(i), params(flags), rgba8=fb.data_ptr(), ray_counts=counts.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), fb.view(-1))
            if rank == 0:
                return deinterleave_into(mg_final[slot & 1], gathered, part)
        return None

    def step_device(i, flags=0, counts=None):
        if s.dynamic:
            update_scene_device(i)
        render_and_collect(i, flags, rays_dev if counts is None else counts, i & 1)

    # ---- one instrumented frame: deterministic traversal counters for the roofline accounting
    step_device(0, abi.RT_RENDER_COUNTERS)
    torch.cuda.synchronize()
    st_count = gpu.stats()
    bytes_frame = algorithmic_bytes(st_count, s, pixels)

    # ---- clocks: sampled from here (warm-up, timed region, kernel-timing pass, e2e region) — the timed region alone lasts
    #      a few tens of milliseconds, shorter than nvidia-smi's sampling period
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.wait_first_sample()

    # ---- warm-up
    for i in range(args.warmup):
        step_device(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    # ---- timed region: K steps enqueued back to back (no host synchronisation inside), a CUDA-event pair per step on the
    #      launching stream with the L2 flush between steps outside the pairs; ray counts land in a per-step device slot
    launches0 = gpu.lib.rt_kernel_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = np.zeros(len(KERNELS))
    kernel_launches = np.zeros(len(KERNELS), np.int64)
    tlas_ms = 0.0
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record(stream)
        step_device(args.warmup + i, 0, rays_steps[i])
        ev[i][1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = gpu.lib.rt_kernel_launches() - launches0
    # ---- per-kernel durations (roofline): a separate pass with CUDA events around every kernel of the frame,
    #      same stream, L2 flushed between frames; kept out of the timed region because the events serialise launches
    ksteps = max(3, min(args.steps, 10))
    for i in range(ksteps):
        flush.zero_()
        step_device(args.warmup + i, abi.RT_RENDER_TIMING)
        st = gpu.stats()
        kernel_ms += np.array(list(st.kernel_ms))
        kernel_launches += np.array(list(st.kernel_launches))
        tlas_ms += st.last_tlas_ms if s.dynamic else 0.0
    total_rays = int(rays_steps.sum().item())
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    r = torch.tensor([total_rays], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    total_ms, total_rays_all = float(t.item()), int(r.item())
    value = total_rays_all / (total_ms * 1e-3) / 1e6

    # ---- e2e: the same steps through the host-facing C ABI, host<->device copies inside the timed region.
    #      N = 1: rt_render_async, two frames in flight like the reference (src/main.rs:917-928): every step copies its
    #      uniforms (and instance records) H2D and its finished frame + ray counts D2H into pinned host memory; the copy
    #      of frame i overlaps the rendering of frame i+1; the step's result is consumed after rt_wait_frame.
    host_fbs = [host_fb, torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory()] if (world == 1 or rank == 0) else [host_fb, host_fb]
    host_rays2 = [host_rays, torch.zeros(2, dtype=torch.int64).pin_memory()]
    copy_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    render_streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)] if world > 1 else None
    mg_overlap = args.mg_overlap == "on"  # measured at 2 ranks on C2: in-order 14.9, overlapped 13.7 Grays/s (the GPU is already full)
    mg_rays = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(2)]
    mg_done = [torch.cuda.Event(), torch.cuda.Event()]
    mg_copied = [None, None]

    def e2e_run(first, count):
        rays, pending = 0, []
        for k in range(count):
            i = first + k
            if s.dynamic:
                if world == 1:
                    gpu.update_instances_raw(0, len(s.instances), rec_pinned[i % n_frames].data_ptr())  # rt_update_instances: host -> device
                    gpu.update_tlas(update_mode)
                else:
                    update_scene_device(i, from_host=True)
            if world == 1:
                b = k & 1
                slot = gpu.render_async(frame_inputs(i), params(), host_fbs[b].data_ptr(), host_rays2[b].data_ptr())
                pending.append((slot, b))
                if len(pending) == 2:
                    sl, pb = pending.pop(0)
                    gpu.wait_frame(sl)
                    rays += int(host_rays2[pb].sum().item())
            else:
                # N > 1, the same two-frames-in-flight scheme with torch streams: render + all-gather (+ de-interleave on
                # rank 0) on the main stream into device slot b, D2H copies of slot b on the copy stream
                b = k & 1
                if mg_copied[b] is not None:       # slot b's previous frame: its fence, then consume its result
                    mg_copied[b].synchronize()
                    rays += int(host_rays2[b].sum().item())
                if shared is not None:
                    # frames alternate between two render streams and the library's two frame slots, so frame i+1 starts
                    # while frame i is still draining; channel b keeps the two barriers apart
                    # Rank 0's frame buffers rotate over 4: frame k+4 reuses frame k's buffer, and its stores sit behind
                    # barrier k+2 on the same stream, which rank 0 only enters after its copy of frame k+1 (hence of frame k,
                    # the copy stream is in order) has finished.
                    rs = render_streams[b]
                    gpu.render_device_slot(b, rs.cuda_stream, frame_inputs(i), params(abi.RT_RENDER_OUTPUT_IMAGE_ROWS),
                                           rgba8=shared.target_ptr(k & 3), ray_counts=mg_rays[b].data_ptr())
                    with torch.cuda.stream(rs):
                        if mg_copied[b ^ 1] is not None:
                            rs.wait_event(mg_copied[b ^ 1])
                        shared.barrier(channel=b)
                        mg_done[b].record(rs)
                    frame = shared.frame(k & 3) if rank == 0 else None
                else:
                    # in order on one stream; peer path: rank 0 does not enter this frame's barrier before its copy of the
                    # previous frame (the slot the ranks will store into next) has finished
                    frame = render_and_collect(i, 0, mg_rays[b], b, before_barrier=mg_copied[b ^ 1])
                    mg_done[b].record(stream)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(mg_done[b])
                    if rank == 0:
                        host_fbs[b].copy_(frame, non_blocking=True)
                    host_rays2[b].copy_(mg_rays[b], non_blocking=True)
                    mg_copied[b] = torch.cuda.Event()
                    mg_copied[b].record(copy_stream)
        if world > 1:
            order = [(count & 1), ((count + 1) & 1)] if count >= 2 else [0]
            for b in order:
                if mg_copied[b] is not None:
                    mg_copied[b].synchronize()
                    rays += int(host_rays2[b].sum().item())
                    mg_copied[b] = None
        for sl, pb in pending:
            gpu.wait_frame(sl)
            rays += int(host_rays2[pb].sum().item())
        return rays

    e2e_run(0, 2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_rays = e2e_run(args.warmup, args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    r = torch.tensor([e2e_rays], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    e2e_value = int(r.item()) / float(t.item()) / 1e6
    clocks = sampler.stop() if sampler else None
    h2d = 176 + (len(s.instances) * 64 if s.dynamic else 0)
    d2h = (H * W * 4 if rank == 0 else 0) + 16

    if rank == 0:
        # ---- roofline of the dominant kernel
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        names = KERNELS
        dom = int(np.argmax(kernel_ms))
        per_launch_ms = kernel_ms[dom] / max(kernel_launches[dom], 1)
        launches_per_frame = kernel_launches[dom] / ksteps
        bytes_per_launch = bytes_frame[names[dom]] / max(launches_per_frame, 1)
        achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        rays_frame = max(int(st_count.primary_rays + st_count.shadow_rays), 1)
        ncu_note = None
        traffic = None  # DRAM bytes per launch of the dominant kernel from the committed ncu capture (same workload only)
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and world == 1:
            tj = json.load(open(tpath))
            if tj.get("workload") == s.name and not (args.width or args.height or args.instances):
                traffic = tj.get(names[dom])
                ncu_note = tj.get("ncu", {}).get(names[dom])
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "kernel": KERNEL_DESC[names[dom]],
            "peak_source": peak_src,
            "kernel_ms_per_frame": {n: float(kernel_ms[i] / ksteps) for i, n in enumerate(names)},
            "kernel_share_of_step": float((kernel_ms[dom] / ksteps) / (total_ms / args.steps)) if world == 1 else None,
            "algorithmic_bytes_per_launch": bytes_per_launch,
            "launches_per_frame": float(launches_per_frame),
            "ncu": ncu_note,  # from the committed capture of the same kernel (profiles/r01r_summary.md): what actually limits it
            "per_ray": {"nodes": float(sum(st_count.nodes_visited)) / rays_frame, "instances": float(sum(st_count.instances_entered)) / rays_frame,
                        "triangles": float(sum(st_count.triangles_tested)) / rays_frame,
                        "bytes": float(bytes_frame["mega"] if args.pipeline == "mega" else sum(bytes_frame[k] for k in KERNELS[:4])) / rays_frame},
            "note": "achieved = algorithmic bytes (per-ray node/instance/triangle fetches + queue records, DESIGN.md 3) / kernel time; the working "
                    "set of C1-C4 is cache resident (traffic = DRAM bytes of one ncu capture, far below the algorithmic bytes), the kernel is "
                    "limited by instruction issue and load latency at about half SIMD width (profiles/r01r_summary.md: 64 % issue slots busy, 18.5 of 32 lanes active, alu pipe 54 %, top stall long_scoreboard 25 %)",
        }
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            def builder(backend):
                return build_scene(backend, args.workload, args.width or None, args.height or None, num_instances=args.instances or None)
            # bounded sample, about 10 s of CPU work on the box's host cores
            v, info = oracle_sample(args, builder, steps=60 if args.workload in ("c1", "c2", "c3", "default") else 1, warmup=1)
            cpu_baseline = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": info["sample"]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic: reference assets (glb/png), seeded instance transforms",
            "config": workload_config(s, args, {
                "pipeline": args.pipeline, "partition": f"{world} rank(s), row strips of {part.strip_height or H} rows, round-robin",
                "l2": "flushed between timed steps (256 MiB device write)", "tlas_update": (args.update_mode if s.dynamic else "static"),
                "rays_per_frame": rays_frame if world == 1 else None, "gather": gather_path}),
            "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")} if clocks else None,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / args.steps * 1e3,
                    "path": "rt_render_async (pinned host buffers, two frames in flight) + rt_wait_frame" if world == 1
                    else f"rt_render_device + {gather_path} + D2H to pinned memory on rank 0 (copy stream, two frames in flight)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "tlas_update_ms_per_step": (tlas_ms / ksteps) if s.dynamic else None,
        }
        emit(line)
    gpu.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner under NCCL_DEBUG,
    # torch.distributed warnings) are sent to stderr instead.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

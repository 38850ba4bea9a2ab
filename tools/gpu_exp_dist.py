"""Development aid (multi-GPU): where a distributed frame's time goes.  Launch with torch.distributed.run."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ray_tracing_gallery_b200 import abi, native  # noqa: E402
from ray_tracing_gallery_b200.dist import Partition, deinterleave  # noqa: E402
from ray_tracing_gallery_b200.scene import build_scene  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
gpu = native.Renderer(lr)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
gpu.set_stream(stream.cuda_stream)
s = build_scene(gpu, wl)
part = Partition.make(s.width, s.height, world, rank)
rows = part.max_rows
fb = torch.zeros((rows, s.width, 4), dtype=torch.uint8, device=dev)
gathered = torch.zeros((world, rows, s.width, 4), dtype=torch.uint8, device=dev)
rays = torch.zeros(2, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def render(i):
    gpu.render_device(s.uniforms(frame_index=1 + i), part.apply(s.params()), rgba8=fb.data_ptr(), ray_counts=rays.data_ptr())


def timeit(fn, n=20, do_flush=True):
    tot = 0.0
    for i in range(n + 3):
        if do_flush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn(i)
        b.record(stream)
        b.synchronize()
        if i >= 3:
            tot += a.elapsed_time(b)
    t = torch.tensor([tot / n], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def f_render(i):
    render(i)


def f_gather(i):
    dist.all_gather_into_tensor(gathered.view(-1), fb.view(-1))


def f_render_gather(i):
    render(i)
    dist.all_gather_into_tensor(gathered.view(-1), fb.view(-1))


def f_full(i):
    render(i)
    dist.all_gather_into_tensor(gathered.view(-1), fb.view(-1))
    if rank == 0:
        deinterleave(gathered, part)


def f_barrier(i):
    dist.barrier()


# correctness: the gathered, de-interleaved frame equals the frame one GPU renders alone
render(0)
dist.all_gather_into_tensor(gathered.view(-1), fb.view(-1))
torch.cuda.synchronize()
full = None
if rank == 0:
    full = torch.zeros((s.height, s.width, 4), dtype=torch.uint8, device=dev)
    gpu.render_device(s.uniforms(frame_index=1), s.params(), rgba8=full.data_ptr(), ray_counts=rays.data_ptr())
    torch.cuda.synchronize()
    same = bool(torch.equal(full, deinterleave(gathered, part)))
    print(f"[{wl}] world {world}: gathered frame == single-GPU frame: {same}", flush=True)
# fused variant: every rank stores its rows into rank 0's frame over NVLink peer memory
shared = None
try:
    from ray_tracing_gallery_b200.dist import SharedFrame
    shared = SharedFrame(s.width, s.height, dev)
except Exception as e:  # noqa: BLE001
    if rank == 0:
        print(f"[{wl}] SharedFrame unavailable: {type(e).__name__}: {e}", flush=True)
if shared is not None:
    def f_shared(i):
        b = i & 1
        p = part.apply(s.params(flags=abi.RT_RENDER_OUTPUT_IMAGE_ROWS))
        gpu.render_device(s.uniforms(frame_index=1 + i), p, rgba8=shared.target_ptr(b), ray_counts=rays.data_ptr())
        shared.barrier()

    f_shared(0)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(f"[{wl}] world {world}: shared frame == single-GPU frame: {bool(torch.equal(full, shared.frame(0)))}", flush=True)

res = {}
if shared is not None:
    res["render+peer stores+barrier"] = timeit(f_shared)
    res["render+peer stores+barrier (no flush)"] = timeit(f_shared, do_flush=False)
for name, fn in [("render", f_render), ("all_gather only", f_gather), ("render+gather", f_render_gather), ("full", f_full)]:
    res[name] = timeit(fn)
    res[name + " (no flush)"] = timeit(fn, do_flush=False)
if rank == 0:
    print(f"[{wl}] world {world}: " + ", ".join(f"{k} {v:.3f} ms" for k, v in res.items()), flush=True)
gpu.close()
dist.barrier()
dist.destroy_process_group()

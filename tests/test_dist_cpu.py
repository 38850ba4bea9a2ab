"""N>1 host logic on CPU: world_size-2 gloo run of the strip partition, instance broadcast and
frame gather, with the oracle standing in for the per-rank renderer."""
import os
import socket

import numpy as np
import pytest

from ray_tracing_gallery_b200 import abi
from ray_tracing_gallery_b200.dist import Partition, choose_strip_height, deinterleave, deinterleave_into


def test_strips_are_eight_rows_and_shares_may_be_ragged():
    assert choose_strip_height(1080, 1) == 0 and choose_strip_height(1080, 8) == 8
    parts = [Partition.make(1920, 1080, 8, r) for r in range(8)]   # 135 strips of 8 rows over 8 ranks: 17 or 16 strips each
    rows = [p.local_rows for p in parts]
    assert rows == [136] * 7 + [128] and sum(rows) == 1080 and parts[0].max_rows == 136
    parts = [Partition.make(64, 36, 4, r) for r in range(4)]       # the last strip is partial (4 rows) and belongs to rank 0
    assert [p.local_rows for p in parts] == [12, 8, 8, 8]
    assert parts[0].global_rows().tolist() == list(range(0, 8)) + list(range(32, 36))


def test_deinterleave_inverts_the_partition():
    H, W, world = 52, 5, 4   # 7 strips: shares of 16, 16, 12 (8 + the partial 4), 8 rows
    img = np.arange(H * W * 4, dtype=np.uint32).reshape(H, W, 4)
    parts = [Partition.make(W, H, world, r) for r in range(world)]
    slabs = np.zeros((world, parts[0].max_rows, W, 4), np.uint32)
    for r, p in enumerate(parts):
        slabs[r, : p.local_rows] = img[p.global_rows()]
    assert np.array_equal(deinterleave(slabs, parts[0]), img)
    rows = np.concatenate([p.global_rows() for p in parts])
    assert sorted(rows.tolist()) == list(range(H))
    import torch

    out = torch.zeros((H, W, 4), dtype=torch.int64)
    deinterleave_into(out, torch.from_numpy(slabs.astype(np.int64)), parts[0])
    assert np.array_equal(out.numpy(), img)
    one = Partition.make(W, H, 1, 0)
    deinterleave_into(out, torch.from_numpy(img.astype(np.int64)[None]), one)
    assert np.array_equal(out.numpy(), img)


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.binding import Oracle
    from ray_tracing_gallery_b200.dist import broadcast_instances, gather_frame
    from ray_tracing_gallery_b200.scene import build_scene

    W, H = 64, 36   # ragged: rank 0 renders 20 rows, rank 1 16
    o = Oracle(threads=2)
    s = build_scene(o, "c3", W, H, num_instances=8)
    # rank 0 owns the authoritative instance records; the others start from garbage
    rec = s.instances.copy()
    if rank != 0:
        rec["transform"] = 0
    t = torch.from_numpy(rec.view(np.uint8).reshape(-1))
    broadcast_instances(t, 0)
    o.update_instances(0, rec)
    o.update_tlas(abi.RT_UPDATE_REBUILD)
    part = Partition.make(W, H, world, rank)
    p = part.apply(s.params())
    local = o.render(s.uniforms(), p, want=("rgba8",))["rgba8"]
    assert local.shape == (part.local_rows, W, 4)
    frame = gather_frame(torch.from_numpy(local), part).numpy()
    if rank == 0:
        full = o.render(s.uniforms(), s.params(), want=("rgba8",))["rgba8"]
        np.save(os.path.join(tmp, "ok.npy"), np.array([np.array_equal(frame, full)]))
    dist.barrier()
    dist.destroy_process_group()
    o.close()


def test_two_rank_gloo_frame(tmp_path):
    import torch.multiprocessing as mp

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert np.load(tmp_path / "ok.npy")[0]

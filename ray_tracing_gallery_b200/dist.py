"""Image-space partition of one frame across ranks (one process per GPU).

The reference is single-GPU; the path shards because pixels are independent
(no inter-pixel dependency in `ray_generation`, blue noise is indexed by the
global pixel coordinate: shaders/closest_hit_textured.glsl:100).  Layout:

  * rows are dealt to ranks in strips of 8 rows, round-robin, so sky and
    geometry are spread evenly (RtRenderParams.strip_*); the last strip may be
    partial and ranks may own different numbers of rows;
  * every rank holds the full scene (replicated at load);
  * on a TLAS change rank 0 broadcasts the 64-byte instance records
    (`broadcast_instances`), every rank rebuilds its replica;
  * per frame the RGBA8 strips are gathered (`all_gather_into_tensor` over
    slabs padded to the largest share) and rank 0 de-interleaves them into the
    final image with one row-gather.

torch.distributed is the plumbing (NCCL on GPUs, gloo in the CPU tests); the
collectives carry finished bytes only, there is no reduction.
"""
from dataclasses import dataclass

import numpy as np


STRIP_HEIGHT = 8  # rows per strip: two 8x4 warp tiles high, so a rank's warps trace the same compact tiles one GPU would


def choose_strip_height(height: int, world: int) -> int:
    """Strips of 8 rows whatever the image height (the last strip may be partial and ranks may own
    different numbers of rows); measured on C2 at 8 ranks: 1-row strips cost a rank 14 % (profiles/r01_notes.md)."""
    return STRIP_HEIGHT if world > 1 else 0


@dataclass
class Partition:
    width: int
    height: int
    world: int
    rank: int
    strip_height: int

    @classmethod
    def make(cls, width, height, world, rank):
        return cls(width, height, world, rank, choose_strip_height(height, world))

    def global_rows(self, rank=None) -> np.ndarray:
        """Global y of each compact local row of `rank`, ascending."""
        rank = self.rank if rank is None else rank
        y = np.arange(self.height)
        if self.world == 1:
            return y
        return y[(y // self.strip_height) % self.world == rank]

    def rows_of(self, rank) -> int:
        return len(self.global_rows(rank))

    @property
    def local_rows(self) -> int:
        """Rows this rank renders (rt_render's compact output has exactly this many)."""
        return self.rows_of(self.rank)

    @property
    def max_rows(self) -> int:
        """Rows of the largest share: the slab size of the equal-sized all-gather."""
        return max(self.rows_of(r) for r in range(self.world))

    def apply(self, params):
        """Fill the strip fields of an RtRenderParams for this rank."""
        if self.world > 1:
            params.strip_height = self.strip_height
            params.strip_count = self.world
            params.strip_index = self.rank
        return params

    def source_rows(self) -> np.ndarray:
        """For every global row y: its row in the gathered [world * max_rows] slab stack."""
        src = np.zeros(self.height, np.int64)
        m = self.max_rows
        for r in range(self.world):
            g = self.global_rows(r)
            src[g] = r * m + np.arange(len(g))
        return src


def deinterleave(gathered, part: Partition):
    """[world, max_rows, W, C] slabs (rows beyond a rank's share are padding) -> [H, W, C] image.
    Works on torch tensors and numpy arrays."""
    c = gathered.shape[-1]
    if part.world == 1:
        return gathered.reshape(part.height, part.width, c)
    flat = gathered.reshape(part.world * gathered.shape[1], part.width, c)
    src = part.source_rows()
    if isinstance(flat, np.ndarray):
        return flat[src]
    import torch

    return flat.index_select(0, torch.as_tensor(src, device=flat.device))


_INDEX_CACHE = {}


def deinterleave_into(out, gathered, part: Partition):
    """Same as `deinterleave`, written into a preallocated [H, W, C] torch tensor with ONE gather kernel."""
    import torch

    c = gathered.shape[-1]
    if part.world == 1:
        out.copy_(gathered.reshape(part.height, part.width, c))
        return out
    key = (part.width, part.height, part.world, part.strip_height, gathered.shape[1], str(gathered.device))
    idx = _INDEX_CACHE.get(key)
    if idx is None:
        idx = _INDEX_CACHE[key] = torch.as_tensor(part.source_rows(), device=gathered.device)
    torch.index_select(gathered.reshape(part.world * gathered.shape[1], part.width * c), 0, idx, out=out.view(part.height, part.width * c))
    return out


def gather_frame(local_rgba8, part: Partition, out=None):
    """All ranks contribute their [local_rows, W, 4] rows; returns the de-interleaved [H, W, 4] frame
    (meaningful on every rank; rank 0 is the consumer)."""
    import torch
    import torch.distributed as dist

    if part.world == 1:
        return local_rgba8
    m = part.max_rows
    slab = local_rgba8
    if slab.shape[0] != m:  # pad to the common slab size
        slab = torch.zeros((m,) + tuple(local_rgba8.shape[1:]), dtype=local_rgba8.dtype, device=local_rgba8.device)
        slab[: local_rgba8.shape[0]] = local_rgba8
    if out is None:
        out = torch.empty((part.world,) + tuple(slab.shape), dtype=slab.dtype, device=slab.device)
    dist.all_gather_into_tensor(out.view(-1), slab.reshape(-1))
    return deinterleave(out, part)


class SharedFrame:
    """The fused form of render + gather: every rank's render kernels store their rows straight into rank 0's
    frame through NVLink peer memory (RT_RENDER_OUTPUT_IMAGE_ROWS), so the frame needs no all-gather and no
    de-interleave — only a cross-rank barrier before rank 0 reads it.

    Built on torch's symmetric memory (`torch.distributed._symmetric_memory`): `slots` frames of [H, W, 4] bytes
    are allocated symmetrically on every rank; only rank 0's copies are written.  Raises if the platform cannot
    provide peer-mapped memory; callers fall back to `gather_frame`."""

    def __init__(self, width: int, height: int, device, group=None, slots: int = 2):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.width, self.height, self.slots = width, height, slots
        self.buf = symm_mem.empty((slots, height, width, 4), dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.rank = self.hdl.rank
        self.frame_bytes = height * width * 4
        self.root_ptr = int(self.hdl.buffer_ptrs[0])  # rank 0's buffer as seen from this rank
        torch.cuda.synchronize(device)
        self.hdl.barrier(channel=0)

    def target_ptr(self, slot: int) -> int:
        """Device pointer to pass as RtFrameOutputs.rgba8: frame `slot` of rank 0."""
        return self.root_ptr + (slot % self.slots) * self.frame_bytes

    def barrier(self, channel: int = 0):
        """All ranks' stores of the frame are complete and visible once this returns on the current stream."""
        self.hdl.barrier(channel=channel)

    def frame(self, slot: int):
        """Rank 0: the finished [H, W, 4] frame (valid after `barrier`)."""
        return self.buf[slot % self.slots]


def broadcast_instances(records_tensor, src: int = 0):
    """NCCL/gloo broadcast of the instance records (uint8 view of N x 64 bytes) from rank `src`."""
    import torch.distributed as dist

    dist.broadcast(records_tensor, src=src)
    return records_tensor

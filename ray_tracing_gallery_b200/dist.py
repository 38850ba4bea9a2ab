"""Image-space partition of one frame across ranks (one process per GPU).

The reference is single-GPU; the path shards because pixels are independent
(no inter-pixel dependency in `ray_generation`, blue noise is indexed by the
global pixel coordinate: shaders/closest_hit_textured.glsl:100).  Layout:

  * rows are dealt to ranks in strips of `strip_height` rows, round-robin, so
    sky and geometry are spread evenly (RtRenderParams.strip_*);
  * every rank holds the full scene (replicated at load);
  * on a TLAS change rank 0 broadcasts the 64-byte instance records
    (`broadcast_instances`), every rank rebuilds its replica;
  * per frame the RGBA8 strips are gathered (`all_gather_into_tensor` over
    equal-sized slabs) and rank 0 de-interleaves them into the final image.

torch.distributed is the plumbing (NCCL on GPUs, gloo in the CPU tests); the
collectives carry finished bytes only, there is no reduction.
"""
from dataclasses import dataclass

import numpy as np


def choose_strip_height(height: int, world: int, preferred=(8, 4, 2, 1)) -> int:
    for sh in preferred:
        if height % (sh * world) == 0:
            return sh
    raise ValueError(f"image height {height} cannot be dealt to {world} ranks in equal strips")


@dataclass
class Partition:
    width: int
    height: int
    world: int
    rank: int
    strip_height: int

    @classmethod
    def make(cls, width, height, world, rank):
        return cls(width, height, world, rank, choose_strip_height(height, world) if world > 1 else 0)

    @property
    def local_rows(self) -> int:
        return self.height // self.world

    def apply(self, params):
        """Fill the strip fields of an RtRenderParams for this rank."""
        if self.world > 1:
            params.strip_height = self.strip_height
            params.strip_count = self.world
            params.strip_index = self.rank
        return params

    def global_rows(self, rank=None) -> np.ndarray:
        """Global y of each compact local row of `rank`."""
        rank = self.rank if rank is None else rank
        if self.world == 1:
            return np.arange(self.height)
        sh = self.strip_height
        j = np.arange(self.local_rows) // sh
        r = np.arange(self.local_rows) % sh
        return (j * self.world + rank) * sh + r


def deinterleave(gathered, part: Partition):
    """[world, local_rows, W, C] slabs -> [H, W, C] image.  Works on torch tensors and numpy arrays."""
    w, sh = part.world, part.strip_height
    if w == 1:
        return gathered.reshape(part.height, part.width, gathered.shape[-1])
    j = part.local_rows // sh
    g = gathered.reshape(w, j, sh, part.width, gathered.shape[-1])
    if isinstance(g, np.ndarray):
        return np.ascontiguousarray(g.transpose(1, 0, 2, 3, 4)).reshape(part.height, part.width, -1)
    return g.permute(1, 0, 2, 3, 4).contiguous().view(part.height, part.width, -1)


def deinterleave_into(out, gathered, part: Partition):
    """Same as `deinterleave`, written into a preallocated [H, W, C] torch tensor with ONE strided copy."""
    w, sh = part.world, part.strip_height
    c = gathered.shape[-1]
    if w == 1:
        out.copy_(gathered.reshape(part.height, part.width, c))
        return out
    j = part.local_rows // sh
    out.view(j, w, sh, part.width, c).copy_(gathered.view(w, j, sh, part.width, c).permute(1, 0, 2, 3, 4))
    return out


def gather_frame(local_rgba8, part: Partition, out=None):
    """All ranks contribute their [local_rows, W, 4] slab; returns the de-interleaved [H, W, 4] frame
    (meaningful on every rank; rank 0 is the consumer)."""
    import torch
    import torch.distributed as dist

    if part.world == 1:
        return local_rgba8
    if out is None:
        out = torch.empty((part.world,) + tuple(local_rgba8.shape), dtype=local_rgba8.dtype, device=local_rgba8.device)
    dist.all_gather_into_tensor(out.view(-1), local_rgba8.reshape(-1))
    return deinterleave(out, part)


def broadcast_instances(records_tensor, src: int = 0):
    """NCCL/gloo broadcast of the instance records (uint8 view of N x 64 bytes) from rank `src`."""
    import torch.distributed as dist

    dist.broadcast(records_tensor, src=src)
    return records_tensor

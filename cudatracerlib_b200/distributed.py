"""Multi-GPU plumbing of the path: one process per GPU, scene replicated, image split in interleaved tiles,
one reduce of the PixelData accumulator per frame, ray counters summed.  torch.distributed is used for plumbing only.

The reference is single-GPU (no NCCL/MPI anywhere); this is the B200-native scale-out of `Tracer<true>::DoPass`
(Kernel/Tracer.h:209-248): RNG is a pure function of (pass, pixel index, dimension) (Kernel/Sampler_device.h:91-107), so
the partition does not change any path, and adding the zero-initialised accumulators of the other ranks is exact.
"""
import numpy as np

TILE = 64


def tiles_of_rank(w, h, tile_w, tile_h, rank, world):
    """Rectangles (x0, y0, x1, y1) of the tiles `ctl_render_pass_tiled(part=rank, n_parts=world)` renders, in its slot order."""
    tiles_x = (w + tile_w - 1) // tile_w
    tiles_y = (h + tile_h - 1) // tile_h
    out = []
    for t in range(rank, tiles_x * tiles_y, world):
        tx, ty = t % tiles_x, t // tiles_x
        out.append((tx * tile_w, ty * tile_h, min(w, (tx + 1) * tile_w), min(h, (ty + 1) * tile_h)))
    return out


class DistributedFrame:
    """Renders frames of `spp` passes cooperatively.  `render_pass(pass_index, new_trace)` must render THIS rank's tiles of
    one pass into `accum` (a torch tensor of 7*w*h floats on the rank's device, zeroed on new_trace); on the GPU that is
    `tracer.DoPassTiled(...)`, in the CPU tests it is the oracle looping over `tiles_of_rank`."""

    def __init__(self, accum, render_pass, rays_of_last_frame, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.accum, self.render_pass, self.rays_of_last_frame, self.group = accum, render_pass, rays_of_last_frame, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def frame(self, spp, reduce=True):
        for p in range(spp):
            self.render_pass(p, p == 0)
        if reduce and self.world > 1:
            # the ONE collective of the path: sum of the per-rank PixelData accumulators to rank 0
            self.dist.reduce(self.accum, dst=0, op=self.dist.ReduceOp.SUM, group=self.group)

    def total_rays(self):
        import torch
        t = torch.tensor([float(self.rays_of_last_frame())], dtype=torch.float64, device=self.accum.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return int(t.item())


class DistributedPasses:
    """Frames of a whole-frame integrator (WavefrontPathTracer: its queue slots are pixel indices, so the image cannot be tiled) shared by PASS
    index: rank r renders passes r, r + world, r + 2 world, ... of the frame into its own accumulator, then the same single reduce.  Passes are
    independent given their index (sample tables and the sampler skip are functions of it), so every pass is exactly the one a single GPU
    would render; only the order of the per-pixel additions differs.  On the GPU: `tracer.setParameter("PassStride", world);
    tracer.setParameter("PassPhase", rank)` and `render_pass = lambda p, new_trace: tracer.DoPass(new_trace)`; in the CPU tests the oracle
    renders pass `p`."""

    def __init__(self, accum, render_pass, rays_of_last_frame, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.accum, self.render_pass, self.rays_of_last_frame, self.group = accum, render_pass, rays_of_last_frame, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def passes_of_rank(self, spp):
        return list(range(self.rank, spp, self.world))

    def frame(self, spp, reduce=True):
        mine = self.passes_of_rank(spp)
        if not mine:
            self.accum.zero_()          # fewer passes than ranks: this rank contributes zeros
        for k, p in enumerate(mine):
            self.render_pass(p, k == 0)
        if reduce and self.world > 1:
            self.dist.reduce(self.accum, dst=0, op=self.dist.ReduceOp.SUM, group=self.group)

    total_rays = DistributedFrame.total_rays

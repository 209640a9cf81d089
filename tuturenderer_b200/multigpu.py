"""Multi-GPU rendering: split samples-per-pixel across ranks, one reduce of the fp32 accumulation
buffer at the end.  One process per GPU; torch.distributed is the plumbing
(NCCL over NVLink on the GPU box, gloo in the CPU tests).

Samples are independent in the reference (no pixel filter, no adaptive sampling:
PathTracing.hpp:508-513), and the Philox stream is keyed by (pixel, sample), so the union of the
ranks' sample ranges is exactly the single-GPU sample set.
"""
from __future__ import annotations

from typing import Callable


def split_samples(spp: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous sample range [begin, begin+count) of `rank`; remainders go to the low ranks."""
    if world < 1 or not (0 <= rank < world) or spp < 0:
        raise ValueError("bad spp/world/rank")
    base, rem = divmod(spp, world)
    count = base + (1 if rank < rem else 0)
    begin = rank * base + min(rank, rem)
    return begin, count


def reduce_sum_to_root(accum, dist=None, dst: int = 0):
    """The path's single collective: sum the per-rank accumulation buffers onto rank `dst`."""
    if dist is None:
        import torch.distributed as dist  # noqa: PLC0415
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum


def render_distributed(accumulate: Callable[[int, int, "object"], None], accum, spp: int, rank: int, world: int,
                       dist=None):
    """accumulate(sample_begin, sample_count, accum) adds this rank's samples into `accum`
    (a torch tensor of width*height*3 sums); afterwards the buffers are reduced onto rank 0.
    Returns the (reduced, on rank 0) accumulation buffer; divide by spp to get the image."""
    begin, count = split_samples(spp, world, rank)
    if count:
        accumulate(begin, count, accum)
    return reduce_sum_to_root(accum, dist)


class CudaRenderer:
    """Per-rank renderer on top of the C ABI: device accumulation buffer as a torch tensor."""

    def __init__(self, scene, device: int = 0, paths_in_flight: int = 0, profile_stages: bool = False, lanes: int = 0):
        import torch
        from . import api
        self.torch = torch
        self.api = api
        self.device = device
        torch.cuda.set_device(device)
        self.ctx = api.Context(device)
        self.ctx.upload(scene)
        self.ctx.configure(paths_in_flight, profile_stages, lanes)
        self.scene = scene
        n = scene.width * scene.height * 3
        self.accum = torch.zeros(n, dtype=torch.float32, device=f"cuda:{device}")
        self.rgb = torch.empty(n, dtype=torch.float32, device=f"cuda:{device}")

    def accumulate(self, begin: int, count: int, accum=None, seed: int = 1) -> None:
        accum = self.accum if accum is None else accum
        stream = self.api.stream_handle(self.torch.cuda.current_stream().cuda_stream)
        self.ctx.render_accumulate_device(begin, count, seed, accum.data_ptr(), stream)

    def finalize(self, spp: int, accum=None):
        accum = self.accum if accum is None else accum
        stream = self.api.stream_handle(self.torch.cuda.current_stream().cuda_stream)
        self.ctx.finalize_device(accum.data_ptr(), 1.0 / spp, self.rgb.data_ptr(), stream)
        return self.rgb.view(self.scene.height, self.scene.width, 3)

    def render_bdpt(self, spp: int, seed: int, rank: int = 0, world: int = 1):
        """One whole BDPT frame (reference include/BDPT.hpp): ranks add the strategy sums of their sample
        range (t = 1 splats land in the same full-frame buffer), one reduce, then bkgcolor + sum / spp."""
        self.accum.zero_()
        stream = lambda: self.api.stream_handle(self.torch.cuda.current_stream().cuda_stream)
        render_distributed(lambda b, c, a: self.ctx.render_bdpt_accumulate_device(b, c, seed, a.data_ptr(), stream()),
                           self.accum, spp, rank, world)
        if rank == 0:
            self.ctx.finalize_bdpt_device(self.accum.data_ptr(), 1.0 / spp, self.rgb.data_ptr(), stream())
            return self.rgb.view(self.scene.height, self.scene.width, 3)
        return None

    def render(self, spp: int, seed: int, rank: int = 0, world: int = 1):
        """One whole frame; returns the image tensor on rank 0 (None elsewhere)."""
        self.accum.zero_()
        render_distributed(lambda b, c, a: self.accumulate(b, c, a, seed), self.accum, spp, rank, world)
        if rank == 0:
            return self.finalize(spp)
        return None

"""Multi-GPU plumbing: pixel-strip partition of one frame and the single gather of the framebuffer.

The reference parallelises a frame over 128x128 tiles on std::threads sharing one scene
(src/scene.cpp:470-506).  Here the scene is replicated on every GPU, image rows are dealt out in
cyclic strips (strip s -> rank s % world), each rank renders its strips (plus a one-row halo for the
Sobel window, inside rtb_render_strips) and ONE collective brings the rows to rank 0:
`torch.distributed.gather` of equal-sized compact buffers (NCCL over NVLink on the GPU box, gloo in
the CPU tests), followed by an index_copy that puts the rows back in image order.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

DEFAULT_STRIP_ROWS = 8


def owned_rows(height: int, strip_rows: int, rank: int, world: int) -> np.ndarray:
    """Rows of the image rendered by `rank` (ascending).  Mirrors rtb_render_strips (include/rtb.h)."""
    y = np.arange(height)
    return y[(y // strip_rows) % world == rank]


def max_rows(height: int, strip_rows: int, world: int) -> int:
    return max(len(owned_rows(height, strip_rows, r, world)) for r in range(world))


def gather_frame(local: torch.Tensor, height: int, strip_rows: int, rank: int, world: int, group=None, dst: int = 0):
    """local: [max_rows, width, 3] float32 (first len(owned_rows) rows valid).  Returns the full
    [height, width, 3] frame on `dst`, None elsewhere.  One collective."""
    width = local.shape[1]
    if world == 1:
        return local[:height]
    bufs = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    dist.gather(local, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    frame = torch.empty((height, width, 3), dtype=local.dtype, device=local.device)
    for r in range(world):
        rows = torch.as_tensor(owned_rows(height, strip_rows, r, world), device=local.device, dtype=torch.long)
        frame.index_copy_(0, rows, bufs[r][: len(rows)])
    return frame

"""Multi-GPU plumbing: pixel-strip partition of one frame and the single exchange of the framebuffer.

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

DEFAULT_STRIP_ROWS = 16


def owned_rows(height: int, strip_rows: int, rank: int, world: int, origin: int = 0) -> np.ndarray:
    """Rows of the image rendered by `rank` (ascending).  Mirrors rtb_strip_rows (include/rtb.h): strips are counted from
    row `origin` (Renderer.strip_origin(): the first row that can contain geometry), floor division for the rows above it."""
    y = np.arange(height)
    return y[np.floor_divide(y - origin, strip_rows) % world == rank]


def max_rows(height: int, strip_rows: int, world: int, origin: int = 0) -> int:
    return max(len(owned_rows(height, strip_rows, r, world, origin)) for r in range(world))


def gather_frame(local: torch.Tensor, height: int, strip_rows: int, rank: int, world: int, group=None, dst: int = 0, origin: int = 0):
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
        rows = torch.as_tensor(owned_rows(height, strip_rows, r, world, origin), device=local.device, dtype=torch.long)
        frame.index_copy_(0, rows, bufs[r][: len(rows)])
    return frame


class FrameExchange:
    """Assembles the strips of all ranks into one frame on rank 0, once per frame.

    Preferred transport ("p2p"): rank 0's framebuffer lives in symmetric memory
    (torch.distributed._symmetric_memory: CUDA VMM allocations mapped into every rank over NVLink), every rank's
    output kernel stores its rows straight into it at their image position (rtb_render_strips_to_frame with the
    peer pointer), and one device-side barrier on the render stream closes the frame: the transfer is the output
    kernel's own stores, there is no separate collective and no un-permute.  Two frame buffers alternate so that
    a rank running ahead never overwrites rows rank 0 is still reading (see render()).
    Fallback ("nccl"): compact per-rank buffers, one torch.distributed.gather over NCCL, rows put back in image
    order with preallocated index tensors (gather_frame above, without its per-call allocations).
    """

    def __init__(self, height: int, width: int, strip_rows: int, rank: int, world: int, device, transport: str = "auto", origin: int = 0):
        self.h, self.w, self.strip, self.rank, self.world = height, width, strip_rows, rank, world
        self.origin = origin    # Renderer.strip_origin() of the scene / camera the exchange is used with (nccl transport needs it)
        self.device = torch.device(device)
        self.transport = "nccl"
        self.frame = None
        if transport in ("auto", "p2p") and world > 1 and self.device.type == "cuda":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                # two frames, used alternately: see render()
                self.frames = symm_mem.empty((2, height, width, 3), dtype=torch.float32, device=self.device)
                self.hdl = symm_mem.rendezvous(self.frames, dist.group.WORLD)
                self.root_ptr = int(self.hdl.buffer_ptrs[0])
                self.frame_bytes = height * width * 3 * 4
                self.parity = 0
                self.frame = self.frames[0]
                self.transport = "p2p"
            except Exception as e:   # no VMM / fabric handle support on this box: use the collective
                if transport == "p2p":
                    raise
                self.why_not_p2p = repr(e)
                self.frame = None
        # every rank must agree on the transport
        if world > 1:
            flag = torch.tensor([1 if self.transport == "p2p" else 0], device=self.device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                self.transport = "nccl"
        if self.transport == "nccl":
            self.local = torch.empty((max_rows(height, strip_rows, world, origin), width, 3), dtype=torch.float32, device=self.device)
            if rank == 0:
                self.bufs = [torch.empty_like(self.local) for _ in range(world)]
                self.rows = [torch.as_tensor(owned_rows(height, strip_rows, r, world, origin), device=self.device, dtype=torch.long) for r in range(world)]
                if self.frame is None:
                    self.frame = torch.empty((height, width, 3), dtype=torch.float32, device=self.device)

    def align(self, stream):
        """Device-side rendezvous of all ranks on `stream` (no data): lets a timing harness start every rank's frame together,
        so that a frame's closing barrier does not absorb the ranks' start skew."""
        if self.transport == "p2p":
            with torch.cuda.stream(stream):
                self.hdl.barrier(channel=1)
        elif self.world > 1:
            stream.synchronize()
            dist.barrier()

    def render(self, renderer, stream, end_event=None):
        """Render this rank's strips and exchange.  Returns (stats, frame on rank 0 / None elsewhere).  Everything is
        enqueued on `stream` (a torch.cuda.Stream); `end_event` (optional) is recorded behind the exchange, before the
        host waits for the frame's statistics."""
        import contextlib
        on = (lambda: torch.cuda.stream(stream)) if self.device.type == "cuda" else contextlib.nullcontext
        raw = stream.cuda_stream if self.device.type == "cuda" else None
        if self.transport == "p2p":
            # Frames alternate between two symmetric buffers and ONE device-side barrier on the render stream closes each
            # frame.  Write-after-read safety without a second barrier: a rank stores frame N+2 into the buffer of frame N
            # only after it passed barrier N+1, which rank 0's stream reaches after everything it enqueued before — in
            # particular whatever consumes frame N (conversion, copy, a user kernel on this stream).  A consumer on another
            # stream must order itself before rank 0's next render() call.
            k = self.parity
            self.parity ^= 1
            # the frame and its closing barrier are enqueued back to back; only then does the host wait for the frame's counters
            renderer.render_strips_to_frame_begin(self.root_ptr + k * self.frame_bytes, self.strip, self.rank, self.world, stream=raw)
            with on():
                self.hdl.barrier(channel=0)
            if end_event is not None:
                end_event.record(stream)
            st = renderer.render_end()
            self.frame = self.frames[k]
            return st, (self.frame if self.rank == 0 else None)
        st = renderer.render_strips_device(self.local.data_ptr(), self.strip, self.rank, self.world, stream=raw)
        with on():
            dist.gather(self.local, self.bufs if self.rank == 0 else None, dst=0)
            if self.rank == 0:
                for r in range(self.world):
                    self.frame.index_copy_(0, self.rows[r], self.bufs[r][: len(self.rows[r])])
        if end_event is not None:
            end_event.record(stream)
        return st, (self.frame if self.rank == 0 else None)

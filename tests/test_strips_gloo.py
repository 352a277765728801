"""Host-side multi-GPU logic on CPU: the cyclic strip partition and the one-collective frame
gather, with world_size 2 and 3 over gloo."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rendering_b200 import _ffi
from rendering_b200 import dist as rdist


def test_partition_matches_c_abi_and_covers_every_row_once():
    lib = C.CDLL(_ffi.CUDA_LIB_PATH)
    lib.rtb_strip_rows_owned.restype = C.c_int
    lib.rtb_strip_rows.restype = C.c_int
    lib.rtb_strip_rows.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    for height, strip, world in [(1080, 8, 8), (1080, 32, 4), (256, 8, 3), (7, 8, 2), (92, 5, 8)]:
        # strips counted from row `origin` (the first row that can hold geometry): any origin still covers every row once
        for origin in (0, 1, 7, height // 3, height - 1):
            seen = np.zeros(height, int)
            for r in range(world):
                rows = rdist.owned_rows(height, strip, r, world, origin)
                out = np.empty(height, np.int32)
                n = lib.rtb_strip_rows(height, strip, origin, r, world, out.ctypes.data)
                assert n == len(rows) and np.array_equal(out[:n], rows)
                if origin == 0:
                    assert n == lib.rtb_strip_rows_owned(height, strip, r, world)
                seen[rows] += 1
            assert (seen == 1).all()
            # the strip that starts at the origin belongs to rank 0
            assert origin >= height or origin in rdist.owned_rows(height, strip, 0, world, origin)
    assert lib.rtb_strip_rows_owned(100, 0, 0, 2) < 0     # bad arguments are refused
    assert lib.rtb_strip_rows(100, 0, 0, 0, 2, None) < 0


def _worker(rank, world, port, height, width, strip, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(height * width * 3, dtype=torch.float32).reshape(height, width, 3)
    rows = rdist.owned_rows(height, strip, rank, world)
    local = torch.full((rdist.max_rows(height, strip, world), width, 3), -1.0)
    local[: len(rows)] = full[torch.as_tensor(rows)]          # what rtb_render_strips would have produced
    frame = rdist.gather_frame(local, height, strip, rank, world)
    if rank == 0:
        assert frame is not None and torch.equal(frame, full)
        open(os.path.join(out_dir, "ok"), "w").write("1")
    else:
        assert frame is None
    dist.destroy_process_group()


@pytest.mark.parametrize("world,height,strip", [(2, 37, 8), (3, 50, 4)])
def test_gather_frame_gloo(tmp_path, world, height, strip):
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, height, 16, strip, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").exists()


class _FakeRenderer:
    """Stands in for rb.Renderer on the CPU box: fills the compact buffer the way rtb_render_strips would."""

    def __init__(self, full, exchange):
        self.full, self.ex = full, exchange

    def render_strips_device(self, ptr, strip, rank, world, stream=None):
        assert ptr == self.ex.local.data_ptr()
        rows = rdist.owned_rows(self.full.shape[0], strip, rank, world)
        self.ex.local[: len(rows)] = self.full[torch.as_tensor(rows)]
        return {"rows": len(rows)}


def _exchange_worker(rank, world, port, height, width, strip, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(height * width * 3, dtype=torch.float32).reshape(height, width, 3)
    ex = rdist.FrameExchange(height, width, strip, rank, world, "cpu")
    assert ex.transport == "nccl"          # the collective path (gloo here); p2p needs CUDA symmetric memory
    for _ in range(2):                     # buffers are reused frame after frame
        st, frame = ex.render(_FakeRenderer(full, ex), None)
        assert (frame is not None) == (rank == 0)
        if rank == 0:
            assert torch.equal(frame, full)
    if rank == 0:
        open(os.path.join(out_dir, "ok"), "w").write("1")
    dist.destroy_process_group()


def test_frame_exchange_collective_path_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000) + 7
    mp.spawn(_exchange_worker, args=(2, port, 37, 16, 8, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()

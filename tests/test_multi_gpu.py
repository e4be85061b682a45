"""-m gpu, needs >= 2 CUDA devices (skipped otherwise): one process per GPU, every exchange mode x partition, blocking
and pipelined frames; every rank must end with the oracle's full frame."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import zig_vulkan_b200 as zv
    from zig_vulkan_b200 import ffi, scenes
    from oracle import orc

    W, H = 256, 144  # 36 strips: divisible by 2 and 4, not by 8
    grid = scenes.build_grid(64)
    mats = zv.terrain_materials()
    sun = scenes.sun(True)
    cams = [scenes.camera_from_pose(W, H, o, q) for o, q in scenes.sweep_poses(5)]
    sc = orc.OracleScene.from_grid(grid, mats)
    refs = [sc.render(c, sun)[0] for c in cams]
    failures = []
    for partition in ("interleave", "slab"):
        for exchange in ("allgather", "peer", "peerflags"):
            if partition == "slab":
                h = H // world
                ctx = ffi.Context(W, H, len(grid.brick_indices), device=rank, rows=(rank * h, (rank + 1) * h))
            else:
                ctx = ffi.Context(W, H, len(grid.brick_indices), device=rank, part=(rank, world))
            ctx.upload_grid(grid, mats)
            ids = [ffi.Context.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, 0)
            ctx.comm_init(rank, world, ids[0])
            if exchange != "allgather":
                handles = [None] * world
                dist.all_gather_object(handles, ctx.comm_ipc_handle())
                ctx.comm_open_peers(rank, world, b"".join(handles))
                ctx.comm_set_exchange(ffi.VRT_EXCHANGE_PEER_STORE if exchange == "peer" else ffi.VRT_EXCHANGE_PEER_FLAGS)
            # blocking frames
            for cam, ref in zip(cams, refs):
                ctx.trace(cam, sun)
                ctx.sync()
                dist.barrier()
                if not np.array_equal(ctx.read_framebuffer(), ref):
                    failures.append(f"{partition}/{exchange}: blocking frame differs on rank {rank}")
                dist.barrier()
            # pipelined frames, every rank copies to its own pinned buffers
            bufs = [torch.zeros(H, W, 4, dtype=torch.uint8).pin_memory() for _ in cams]
            for cam, buf in zip(cams, bufs):
                ctx.trace_to_host_async(cam, sun, buf.data_ptr() if rank % 2 == 0 else None)
            ctx.sync()
            dist.barrier()
            if rank % 2 == 0:
                for buf, ref in zip(bufs, refs):
                    if not np.array_equal(buf.numpy(), ref):
                        failures.append(f"{partition}/{exchange}: pipelined frame differs on rank {rank}")
            dist.barrier()
            ctx.close()
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write("\n".join(failures))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_every_rank_ends_with_the_full_frame(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.txt").read() == ""

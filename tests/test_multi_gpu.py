"""-m gpu, needs >= 2 CUDA devices (skipped otherwise): one process per GPU, every exchange mode x partition x tile schedule,
blocking frames of a moving camera WITHOUT any host-side barrier, pipelined frames, and the two mixed; every rank must end
every frame with the oracle's full frame.  Logs of the 2- / 4- / 8-GPU runs are kept under profiles/."""
import os
import socket
import sys
import time

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

    W, H = 256, 144  # 36 strips: divisible by 2 and 4, not by 8 (ranks own 4 or 5 strips then)
    grid = scenes.build_grid(64)
    mats = zv.terrain_materials()
    sun = scenes.sun(True)
    cams = [scenes.camera_from_pose(W, H, o, q) for o, q in scenes.sweep_poses(7)]
    sc = orc.OracleScene.from_grid(grid, mats)
    refs = [sc.render(c, sun)[0] for c in cams]
    failures = []
    S, L, D, SH = ffi.VRT_SCHED_STATIC, ffi.VRT_SCHED_LPT, ffi.VRT_SCHED_DEAL, ffi.VRT_SCHED_SHARED
    modes = [("interleave", "allgather", S), ("interleave", "allgather", L), ("interleave", "peer", S), ("interleave", "peer", D), ("interleave", "peerflags", L),
             ("interleave", "peerflags", D), ("interleave", "peerflags", SH), ("interleave", "peer", SH), ("interleave", "peerpush", L), ("interleave", "peerpush", D), ("interleave", "peertiles", L), ("interleave", "peertiles", D),
             ("interleave", "peertiles", SH), ("slab", "peertiles", S), ("slab", "allgather", S), ("slab", "peer", L),
             ("slab", "peerflags", S), ("slab", "peerpush", S)]
    for partition, exchange, sched in modes:
        tag = f"{partition}/{exchange}/sched{sched}"
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
            ctx.comm_set_exchange({"peer": ffi.VRT_EXCHANGE_PEER_STORE, "peerflags": ffi.VRT_EXCHANGE_PEER_FLAGS, "peerpush": ffi.VRT_EXCHANGE_PEER_PUSH,
                                   "peertiles": ffi.VRT_EXCHANGE_PEER_TILES}[exchange])
        ctx.set_schedule(sched, 2)
        # blocking frames of a MOVING camera, no host-side barrier anywhere: a fast rank's next frame must not tear the frame a
        # slow rank is still reading (the ranks are skewed on purpose)
        for rep in range(3):
            for i, (cam, ref) in enumerate(zip(cams, refs)):
                ctx.trace(cam, sun)
                img = ctx.read_framebuffer()
                if (i + rep) % world == rank:
                    time.sleep(0.003)
                if not np.array_equal(img, ref):
                    failures.append(f"{tag}: blocking frame {rep}.{i} differs on rank {rank} in {(img != ref).any(axis=2).sum()} pixels")
            for i, (cam, ref) in enumerate(zip(cams, refs)):
                img = ctx.trace_to_host(cam, sun)
                if not np.array_equal(img, ref):
                    failures.append(f"{tag}: trace_to_host frame {rep}.{i} differs on rank {rank}")
        # pipelined frames, some ranks copy to their own pinned buffers, mixed with a blocking frame
        bufs = [torch.zeros(H, W, 4, dtype=torch.uint8).pin_memory() for _ in cams]
        for cam, buf in zip(cams, bufs):
            ctx.trace_to_host_async(cam, sun, buf.data_ptr() if rank % 2 == 0 else None)
        ctx.trace(cams[2], sun)
        if not np.array_equal(ctx.read_framebuffer(), refs[2]):
            failures.append(f"{tag}: blocking frame after pipelined frames differs on rank {rank}")
        ctx.sync()
        if rank % 2 == 0:
            for i, (buf, ref) in enumerate(zip(bufs, refs)):
                if not np.array_equal(buf.numpy(), ref):
                    failures.append(f"{tag}: pipelined frame {i} differs on rank {rank}")
        dist.barrier()
        # the host as the consumer (VRT_EXCHANGE_HOST): no device exchange, each rank copies exactly its own rows / strips
        if sched not in (D, SH):
            ctx.comm_set_exchange(ffi.VRT_EXCHANGE_HOST)
            mine = np.zeros(H, dtype=bool)
            if partition == "slab":
                mine[rank * (H // world):(rank + 1) * (H // world)] = True
            else:
                for t in range(rank, (H + 3) // 4, world):
                    mine[t * 4:t * 4 + 4] = True
            for i in (1, 4):
                buf = torch.full((H, W, 4), 7, dtype=torch.uint8).pin_memory()
                ctx.trace_to_host_async(cams[i], sun, buf.data_ptr())
                ctx.sync()
                got = buf.numpy()
                if not (np.array_equal(got[mine], refs[i][mine]) and (got[~mine] == 7).all()):
                    failures.append(f"{tag}: host-assembled frame {i} wrong on rank {rank}")
                got2 = np.full((H, W, 4), 7, dtype=np.uint8)
                ctx.trace_to_host(cams[i], sun, out=got2)
                if not (np.array_equal(got2[mine], refs[i][mine]) and (got2[~mine] == 7).all()):
                    failures.append(f"{tag}: host-assembled blocking frame {i} wrong on rank {rank}")
        dist.barrier()
        ctx.close()
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write("\n".join(failures))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_every_rank_ends_with_the_full_frame(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.txt").read() == ""

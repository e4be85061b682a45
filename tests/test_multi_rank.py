"""N > 1 host logic on CPU: two gloo ranks partition the frame into row slabs, each 'traces' its slab (with the oracle
standing in for the device, since there is no GPU here), and an all-gather assembles the frame — the same partition,
offsets and collective shape bench.py / vrt_comm_* use with NCCL (in-place all-gather of H/world rows per rank)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import zig_vulkan_b200 as zv
    from zig_vulkan_b200 import scenes
    from oracle import orc

    W, H = 96, 64
    grid = scenes.build_grid(64)  # replicated on every rank, like the device buffers
    mats = zv.terrain_materials()
    cam = scenes.camera(W, H, origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))
    sun = scenes.sun(True)
    rows = (rank * (H // world), (rank + 1) * (H // world))
    img, _, cnt = orc.OracleScene.from_grid(grid, mats).render(cam, sun, rows=rows, threads=1)
    slab = torch.from_numpy(img[rows[0]:rows[1]].copy())
    frame = torch.zeros(H, W, 4, dtype=torch.uint8)
    dist.all_gather_into_tensor(frame.view(-1), slab.view(-1))  # rank r's slab lands at offset r*slab_bytes
    rays = torch.tensor([cnt["rays"]], dtype=torch.int64)
    dist.all_reduce(rays)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the bench's max-over-ranks timing reduction
    if rank == 0:
        np.savez(out_path, frame=frame.numpy(), rays=rays.numpy(), tmax=t.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_row_slab_partition_and_all_gather(tmp_path, world):
    sys.path.insert(0, ROOT)
    import zig_vulkan_b200 as zv
    from zig_vulkan_b200 import scenes
    from oracle import orc

    out = str(tmp_path / "frame.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    z = np.load(out)
    grid = scenes.build_grid(64)
    cam = scenes.camera(96, 64, origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))
    ref, _, cnt = orc.OracleScene.from_grid(grid, zv.terrain_materials()).render(cam, scenes.sun(True))
    assert np.array_equal(z["frame"], ref)
    assert int(z["rays"][0]) == cnt["rays"]
    assert float(z["tmax"][0]) == float(world)

#!/usr/bin/env python
"""bench.py — Mrays/s of the voxel ray-tracing hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3] [--impl ours|reference]

A "step" is one frame of the workload: every pixel's primary ray marched through the brick grid plus, for C3/C5, one
sun ray per primary hit (brick_raytracer.comp main()).  `value` counts rays actually cast (primary + sun) per second
of device time with the grid resident in HBM; `e2e` is the same frame through vrt_trace_to_host() with the camera/sun
blocks coming from host memory and the RGBA8 frame landing in pinned host memory inside the timed region.
N > 1: one process per GPU (torchrun), the image rows are tiled across ranks and exchanged after the trace kernel;
time is the max over ranks and rays are summed (strong scaling: the frame is fixed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

POSE0 = dict(origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))
L2_FLUSH_BYTES = 144 << 20  # > 126 MB L2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--exchange", default="auto", choices=["auto", "allgather", "peer", "peerflags"],
                    help="N > 1: fused peer stores from the trace kernel (frame barrier = NCCL 4-byte all-reduce, or peer flag words: "
                         "peerflags), or an NCCL all-gather after it; auto = peerflags up to 4 GPUs, all-gather above (measured: profiles/README.md)")
    ap.add_argument("--partition", default="interleave", choices=["interleave", "slab"], help="N > 1: 4-row strips round-robin, or one row slab per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="N = 1: also time the 11 way points of the reference's benchmark fly-through (Benchmark.zig:141-173)")
    ap.add_argument("--baseline-kernel", action="store_true", help="time the reference-shape kernel instead of the tuned one")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled every few ms during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if mask & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def oracle_scene(wl):
    """Grid arrays for the oracle.  The grid builder is libvrt_host (product); the oracle only traces."""
    import zig_vulkan_b200 as zv
    from zig_vulkan_b200 import scenes
    from oracle import orc

    grid = scenes.build_grid(wl.n_voxels, wl.brick_dim, brick_alloc=alloc_for(wl))
    mats = zv.terrain_materials()
    return grid, mats, orc.OracleScene.from_grid(grid, mats)


def alloc_for(wl):
    from zig_vulkan_b200 import scenes

    if wl.n_voxels >= 1024:  # keep material_indices (brick_alloc * brick_dim^3 bytes) off the GiB scale
        return scenes.count_bricks(wl.n_voxels, wl.brick_dim)
    return 0


def time_oracle(wl, steps, warmup, budget_s=None):
    """Full frames of the workload on the host cores.  With oracle/_ref present (the reference's own shader text compiled by g++,
    oracle/ref_shim/) that library is what is timed — kind "reference"; its frame is checked against the hand-written oracle's
    once, outside the timed region.  Otherwise the oracle port is timed — kind "port"."""
    import numpy as np
    from zig_vulkan_b200 import scenes
    from oracle import ref

    grid, mats, sc = oracle_scene(wl)
    cam = scenes.camera(wl.width, wl.height, **POSE0)
    sun = scenes.sun(wl.sun)
    cores = os.cpu_count() or 1
    img, _, cnt = sc.render(cam, sun, threads=cores)
    rays = cnt["rays"]  # the shader has no counters: rays cast = pixels + one sun ray per primary hit, as the oracle counts them
    kind = "port"
    render = lambda: sc.render(cam, sun, threads=cores)
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_shader.so" if wl.brick_dim <= 8 else "libref_shader_wide.so")):
        rimg, _ = ref.render(sc, cam, sun, threads=cores)
        if not np.array_equal(rimg, img):
            raise SystemExit("bench.py: oracle/_ref (the reference's shader text) and the oracle disagree on this frame")
        kind = "reference"
        render = lambda: ref.render(sc, cam, sun, threads=cores)
    for _ in range(warmup):
        render()
    times = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        render()
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_begin > budget_s:
            break
    return rays, times, cores, kind


def run_reference(args):
    """The reference's own implementation of the path on the host cores: its compute shader (brick_raytracer.comp + rand.comp)
    compiled by g++ under oracle/ref_shim/ (oracle/_ref, kind "reference"), all host threads, one invocation per pixel.  Falls
    back to the oracle port (kind "port") only if oracle/_ref was not built."""
    from zig_vulkan_b200 import scenes

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = scenes.WORKLOADS[args.workload]
    steps = min(args.steps, 60)
    rays, times, cores, kind = time_oracle(wl, steps, min(args.warmup, 3), budget_s=150.0)
    total = sum(times)
    value = rays * len(times) / total / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": len(times), "warmup": min(args.warmup, 3),
        "ms_per_step": total / len(times) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": f"{wl.name}: {wl.description}", "pose": "pose0", "rays_per_step": rays},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind,
                         "sample": f"{len(times)} full {wl.width}x{wl.height} frames, all rows, "
                                   + ("oracle/_ref/libref_shader.so (the reference's shader text, g++)" if kind == "reference" else "oracle/liboracle.so") + f" with {cores} threads"},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # NCCL prints its version banner (and any debug output) on stdout; this program's stdout is ONE JSON line
    # (NCCL only honours NCCL_DEBUG_FILE above the VERSION level)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import numpy as np
    import torch
    import torch.distributed as dist

    import zig_vulkan_b200 as zv
    from zig_vulkan_b200 import ffi, scenes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the trace path has no CPU fallback (use --impl reference for the host-core baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.exchange == "auto":
        args.exchange = "peerflags" if world <= 4 else "allgather"
    wl = scenes.WORKLOADS[args.workload]
    W, H = wl.width, wl.height
    interleave = world > 1 and args.partition == "interleave" and not args.baseline_kernel
    if not interleave and H % world != 0:
        raise SystemExit(f"image height {H} is not divisible by {world} ranks")
    rows = (rank * (H // world), (rank + 1) * (H // world)) if (world > 1 and not interleave) else (0, 0)
    part = (rank, world) if interleave else None
    n_pixels = W * H

    grid = scenes.build_grid(wl.n_voxels, wl.brick_dim, brick_alloc=alloc_for(wl))
    mats = zv.terrain_materials()
    cam = scenes.camera(W, H, **POSE0)
    sun = scenes.sun(wl.sun)
    brick_bytes = wl.brick_dim ** 3 // 8

    flags = ffi.VRT_FLAG_BASELINE if args.baseline_kernel else 0
    ctx = ffi.Context(W, H, len(grid.brick_indices), brick_dim=wl.brick_dim, n_brick_alloc=grid.brick_alloc, device=local_rank, flags=flags,
                      rows=rows, part=part)
    stream = torch.cuda.Stream(dev)  # a non-default stream: handle 0 would mean "restore the ctx's own stream"
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)  # torch.cuda.Event only sees torch's current stream
    ctx.upload_grid(grid, mats)

    if world > 1:
        if rank == 0:
            uid = torch.frombuffer(bytearray(ffi.Context.comm_unique_id()), dtype=torch.uint8).to(dev)
        else:
            uid = torch.empty(ffi.VRT_NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
        dist.broadcast(uid, 0)
        ctx.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
        if args.exchange in ("peer", "peerflags"):
            mine = torch.frombuffer(bytearray(ctx.comm_ipc_handle()), dtype=torch.uint8).to(dev)
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)
            ctx.comm_open_peers(rank, world, b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
            ctx.comm_set_exchange(ffi.VRT_EXCHANGE_PEER_STORE if args.exchange == "peer" else ffi.VRT_EXCHANGE_PEER_FLAGS)

    # ray / request-byte counters of this rank's rows (identical to the oracle's, tests/test_golden.py): the reference-shape
    # kernel for slabs, the tuned kernel's counting variant for interleaved strips
    cctx = ffi.Context(W, H, len(grid.brick_indices), brick_dim=wl.brick_dim, n_brick_alloc=grid.brick_alloc, device=local_rank,
                       flags=ffi.VRT_FLAG_AOV | (0 if interleave else ffi.VRT_FLAG_BASELINE), rows=rows, part=part)
    cctx.upload_grid(grid, mats)
    cctx.trace(cam, sun)
    counters = cctx.counters()
    cctx.close()
    my_rays = counters["rays"]
    my_rows = (sum(min(4, H - t * 4) for t in range(rank, (H + 3) // 4, world)) if interleave else (rows[1] - rows[0] if world > 1 else H))
    my_alg_bytes = 4 * my_rows * W + 4 * counters["status_fetches"] + (4 + brick_bytes) * counters["bricks_entered"] + 25 * counters["hits"]

    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------------------------------------------------------- device-resident timing
    barrier()  # ranks enter the first exchanged frame together (set-up time differs from rank to rank)
    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        ctx.trace(cam, sun)
    launches_per_step = ctx.last_trace_launches()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # L2 flush between timed iterations (outside the per-step events)
        starts[i].record(stream)
        ctx.trace(cam, sun)
        ends[i].record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop()
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = float(sum(step_ms))

    # ---------------------------------------------------------------- end-to-end through the C ABI with host buffers
    # (a) frame latency: vrt_trace_to_host per frame, blocking (camera+sun from host structs -> frame in pinned host memory)
    host_frames = [torch.empty(n_pixels * 4, dtype=torch.uint8).pin_memory() for _ in range(2)]
    lat_s = 0.0
    n_lat = min(args.steps, 50)
    for i in range(3 + n_lat):
        flush.fill_(i & 0xFF)
        barrier()
        t0 = time.perf_counter()
        if rank == 0:
            ctx.trace_to_host(cam, sun, out_ptr=host_frames[0].data_ptr())
        else:
            ctx.trace(cam, sun)
            ctx.sync()
        if world > 1:
            dist.barrier()
        if i >= 3:
            lat_s += time.perf_counter() - t0
    # (b) frame throughput: the same call pipelined (vrt_trace_to_host_async, 2 frames in flight: the copy of frame k
    # overlaps the trace of frame k+1).  The L2 flush is enqueued between frames INSIDE the timed region.
    # Rank 0 receives the frames on its host; the other ranks take part in the same frame ring without a host copy.
    for i in range(4):
        ctx.trace_to_host_async(cam, sun, host_frames[i & 1].data_ptr() if rank == 0 else None)
    ctx.sync()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ctx.trace_to_host_async(cam, sun, host_frames[i & 1].data_ptr() if rank == 0 else None)
    ctx.sync()
    barrier()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    r = torch.tensor([float(my_rays), float(my_alg_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    total_ms, e2e_s = float(t[0]), float(t[1])
    rays, alg_bytes = float(r[0]), float(r[1])

    if rank == 0:
        peak, peak_kind = peaks()
        ms_per_step = total_ms / args.steps
        value = rays / (ms_per_step * 1e-3) / 1e6
        # roofline of the dominant kernel = the trace kernel; at N=1 the step IS that one launch
        achieved = (my_alg_bytes / (ms_per_step * 1e-3)) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(args.workload)
        except Exception:
            pass
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{wl.name}: {wl.description}", "pose": "pose0 origin (0,-10,28) pitch 25deg", "rays_per_step": int(rays),
                "grid_bricks": len(grid.brick_indices), "active_bricks": grid.active_bricks, "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB fill)",
                "kernel": "baseline" if args.baseline_kernel else "tuned", "partition": (f"4-row strips round-robin over {world} ranks" if interleave else f"{world} row slabs") if world > 1 else "whole frame",
                "exchange": args.exchange if world > 1 else "none",
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_kind": peak_kind, "algorithmic_bytes_per_launch": int(my_alg_bytes), "kernel_ms": ms_per_step,
                         "note": "request-byte model of the reference algorithm (DESIGN.md); rank 0's launch"},
            "e2e": {"value": rays * args.steps / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 128, "d2h_bytes_per_step": n_pixels * 4,
                    "ms_per_step": e2e_s / args.steps * 1e3, "frame_latency_ms": lat_s / n_lat * 1e3,
                    "how": "vrt_trace_to_host_async per frame (camera+sun host structs in, RGBA8 frame into pinned host memory, 2 frames in flight), "
                           "L2 flush enqueued between frames inside the timed region; frame_latency_ms = blocking vrt_trace_to_host"},
            "gpu_launches": launches_per_step * args.steps, "wall_ms": wall_ms, "clocks": clocks,
        }
        if world == 1 and args.sweep:
            # the reference's own benchmark camera path (offsets are in its world units; our world has the same 64-unit extent)
            per_pose = []
            for origin, yaw in scenes.sweep_poses(11):
                pcam = scenes.camera_from_pose(W, H, origin, yaw)
                cctx = ffi.Context(W, H, len(grid.brick_indices), brick_dim=wl.brick_dim, n_brick_alloc=grid.brick_alloc, device=local_rank,
                                   flags=ffi.VRT_FLAG_AOV)
                cctx.upload_grid(grid, mats)
                cctx.trace(pcam, sun)
                prays = cctx.counters()["rays"]
                cctx.close()
                ms = []
                for i in range(3 + 20):
                    flush.fill_(i & 0xFF)
                    ctx.trace(pcam, sun)
                    if i >= 3:
                        ms.append(ctx.last_trace_ms())
                per_pose.append({"origin": [round(v, 3) for v in origin], "rays": prays, "ms": sum(ms) / len(ms), "mrays_s": prays / (sum(ms) / len(ms) * 1e-3) / 1e6})
            line["sweep"] = {"poses": per_pose, "mean_mrays_s": sum(p["mrays_s"] for p in per_pose) / len(per_pose),
                             "total_mrays_s": sum(p["rays"] for p in per_pose) / sum(p["ms"] * 1e-3 for p in per_pose) / 1e6}
        if world == 1:
            # the step after the path: the reference's present pass (image.frag) over the traced frame, same resolution
            dn = []
            for i in range(3 + 20):
                flush.fill_(i & 0xFF)
                ctx._check(ctx._l.vrt_denoise(ctx.handle, ffi.DenoiseParams.default(), W, H, 0))
                if i >= 3:
                    dn.append(ctx.last_denoise_ms())
            dn_ms = sum(dn) / len(dn)
            line["denoise"] = {"ms_per_frame": dn_ms, "mpixels_s": n_pixels / (dn_ms * 1e-3) / 1e6, "params": "samples 20, bias 0.6, multiplier 1.5, tolerance 20 (GraphicsPipeline.zig:34-39)",
                               "algorithmic_gb_s": 8 * n_pixels / (dn_ms * 1e-3) / 1e9, "bound": "ALU/FMA issue (21 samples x 2 pow per pixel), not HBM"}
        if world == 1:
            # explicit-ray mode (vrt_trace_rays): this camera's primary rays as a 32 B/ray device buffer in, 32 B/ray hit records out
            jj, ii = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
            u, v = (ii / np.float32(W - 1))[..., None], (jj / np.float32(H - 1))[..., None]
            hor, ver, llc, org = (np.array(list(getattr(cam, f)), dtype=np.float32) for f in ("horizontal", "vertical", "lower_left_corner", "origin"))
            d = (hor * u + llc) + (ver * v - org)
            d = d / np.linalg.norm(d, axis=-1, keepdims=True)
            if H % 4 == 0 and W % 8 == 0:  # 8x4-pixel tiles, tile after tile: 32 consecutive rays = one coherent tile (the caller chooses the order)
                d = d.reshape(H // 4, 4, W // 8, 8, 3).transpose(0, 2, 1, 3, 4)[::-1]  # ground tiles first, like the pixel kernel
            rays_np = np.zeros(n_pixels, dtype=ffi.RAY_DTYPE)
            rays_np["origin"], rays_np["direction"] = org, d.reshape(-1, 3)
            d_rays = torch.from_numpy(rays_np.view(np.uint8).reshape(-1)).to(dev)
            d_hits = torch.zeros(n_pixels * 32, dtype=torch.uint8, device=dev)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3 + 20)]
            for i, (a, b) in enumerate(ev):
                flush.fill_(i & 0xFF)
                a.record(stream)
                ctx.trace_rays_device(d_rays.data_ptr(), d_hits.data_ptr(), n_pixels)
                b.record(stream)
            torch.cuda.synchronize(dev)
            er_ms = sum(a.elapsed_time(b) for a, b in ev[3:]) / 20
            line["explicit_rays"] = {"rays": n_pixels, "ms": er_ms, "mrays_s": n_pixels / (er_ms * 1e-3) / 1e6, "ray_io_gb_s": 64 * n_pixels / (er_ms * 1e-3) / 1e9,
                                     "note": "primary rays only, ordered in 8x4-pixel tiles; 32 B ray in + 32 B hit record out per ray, 128-bit loads / stores"}
        if world == 1 and not args.no_cpu_baseline:
            crays, ctimes, cores, kind = time_oracle(wl, 12, 1, budget_s=20.0)
            best = min(ctimes)
            line["cpu_baseline"] = {"value": crays / best / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind,
                                    "sample": f"best of {len(ctimes)} full {W}x{H} frames of the same workload, "
                                              + ("oracle/_ref/libref_shader.so = the reference's shader text compiled by g++" if kind == "reference" else "oracle/liboracle.so")
                                              + f", {cores} threads"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

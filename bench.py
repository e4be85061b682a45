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

POSE0 = dict(origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))   # the headline pose: 25 deg down onto the terrain, ~52 % of the pixels hit
POSE0_DESC = "pose0 origin (0,-10,28) pitch 25deg"
POSE_SURVEY = dict(origin=(0.0, -8.0, 0.0), euler_deg=(0.0, 0.0, 0.0))  # SURVEY.md 8(d)'s "pose 0": level view from above the grid centre
L2_FLUSH_BYTES = 144 << 20  # > 126 MB L2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--exchange", default="auto", choices=["auto", "allgather", "peer", "peerflags", "peerpush", "peertiles"],
                    help="N > 1: fused peer stores from the trace kernel (frame barrier = NCCL 4-byte all-reduce, or peer flag words: "
                         "peerflags), or an NCCL all-gather after it; auto = every combination of exchange and schedule is timed for a few "
                         "frames on this box and the fastest one runs the timed region")
    ap.add_argument("--schedule", default="auto", choices=["auto", "static", "lpt", "deal", "shared"],
                    help="tile order: static bottom-up, cost-sorted (longest first), or cost-sorted and dealt across the ranks (peer exchange modes)")
    ap.add_argument("--partition", default="interleave", choices=["interleave", "slab"], help="N > 1: 4-row strips round-robin, or one row slab per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline measurement (no sweep / REF / C5 / denoise / explicit rays / edit frames)")
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
    return rays, times, cores, kind, grid


def run_reference(args):
    """The reference's own implementation of the path on the host cores: its compute shader (brick_raytracer.comp + rand.comp)
    compiled by g++ under oracle/ref_shim/ (oracle/_ref, kind "reference"), all host threads, one invocation per pixel.  Falls
    back to the oracle port (kind "port") only if oracle/_ref was not built."""
    from zig_vulkan_b200 import scenes

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = scenes.WORKLOADS[args.workload]
    # exactly --steps timed frames after --warmup untimed ones, like the product arm (a C3 frame takes ~0.12 s on 16 threads: the
    # default 200 + 20 is half a minute); the budget is a safety net for a slow box, far above what the driver's runs need
    rays, times, cores, kind, grid = time_oracle(wl, args.steps, args.warmup, budget_s=900.0)
    total = sum(times)
    value = rays * len(times) / total / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
        "ms_per_step": total / len(times) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": f"{wl.name}: {wl.description}", "pose": POSE0_DESC, "rays_per_step": int(rays),
                                      "grid_bricks": len(grid.brick_indices), "active_bricks": grid.active_bricks},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind,
                         "sample": f"{len(times)} full {wl.width}x{wl.height} frames, all rows, "
                                   + ("oracle/_ref/libref_shader.so (the reference's shader text, g++)" if kind == "reference" else "oracle/liboracle.so") + f" with {cores} threads"},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


class Rig:
    """Per-process plumbing: device, stream, L2 flush buffer, rank barrier."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.Stream(self.dev)  # a non-default stream: handle 0 would mean "restore the ctx's own stream"
        torch.cuda.set_stream(self.stream)         # torch.cuda.Event only sees torch's current stream
        self.flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def reduce(self, values, op="max"):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(v) for v in t]

    def make_ctx(self, grid, mats, W, H, brick_dim=4, flags=0, solo=False):
        """A context on the bench stream with the grid uploaded; world > 1 (and not solo): this rank's part of the frame, NCCL
        communicator, peer mappings — every exchange mode selectable afterwards."""
        from zig_vulkan_b200 import ffi

        dist, torch = self.dist, self.torch
        multi = self.world > 1 and not solo
        interleave = multi and self.args.partition == "interleave" and not self.args.baseline_kernel
        if multi and not interleave and H % self.world != 0:
            raise SystemExit(f"image height {H} is not divisible by {self.world} ranks")
        rows = (self.rank * (H // self.world), (self.rank + 1) * (H // self.world)) if (multi and not interleave) else (0, 0)
        ctx = ffi.Context(W, H, len(grid.brick_indices), brick_dim=brick_dim, n_brick_alloc=grid.brick_alloc, device=self.local_rank, flags=flags, rows=rows,
                          part=(self.rank, self.world) if interleave else None)
        ctx.set_stream(self.stream.cuda_stream)
        ctx.upload_grid(grid, mats)
        ctx.interleaved = interleave
        if multi:
            if self.rank == 0:
                uid = torch.frombuffer(bytearray(ffi.Context.comm_unique_id()), dtype=torch.uint8).to(self.dev)
            else:
                uid = torch.empty(ffi.VRT_NCCL_ID_BYTES, dtype=torch.uint8, device=self.dev)
            dist.broadcast(uid, 0)
            ctx.comm_init(self.rank, self.world, bytes(uid.cpu().numpy().tobytes()))
            mine = torch.frombuffer(bytearray(ctx.comm_ipc_handle()), dtype=torch.uint8).to(self.dev)
            allh = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(allh, mine)
            ctx.comm_open_peers(self.rank, self.world, b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
        return ctx

    def timed(self, ctx, cam, sun, steps, warmup):
        """`steps` frames after `warmup`, L2 flushed before each (outside the per-step events).  Returns the per-step device times
        (ms, CUDA events on the launching stream) and the wall time of the region."""
        torch = self.torch
        self.barrier()  # ranks enter the first exchanged frame together
        for i in range(warmup):
            self.flush.fill_(i & 0xFF)
            ctx.trace(cam, sun)
        self.barrier()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        t0 = time.perf_counter()
        self.launches = 0  # this library's kernels enqueued inside the timed region (trace, exchange, and the schedule sort every 8th frame)
        for i in range(steps):
            self.flush.fill_(i & 0xFF)
            starts[i].record(self.stream)
            ctx.trace(cam, sun)
            ends[i].record(self.stream)
            self.launches += ctx.last_trace_launches()
        self.barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        return [s.elapsed_time(e) for s, e in zip(starts, ends)], wall_ms

    def set_mode(self, ctx, exchange, schedule, interval=8):
        from zig_vulkan_b200 import ffi

        if self.world > 1:
            ctx.comm_set_exchange({"allgather": ffi.VRT_EXCHANGE_ALLGATHER, "peer": ffi.VRT_EXCHANGE_PEER_STORE, "peerflags": ffi.VRT_EXCHANGE_PEER_FLAGS,
                                   "peerpush": ffi.VRT_EXCHANGE_PEER_PUSH, "peertiles": ffi.VRT_EXCHANGE_PEER_TILES}[exchange])
        ctx.set_schedule({"static": ffi.VRT_SCHED_STATIC, "lpt": ffi.VRT_SCHED_LPT, "deal": ffi.VRT_SCHED_DEAL, "shared": ffi.VRT_SCHED_SHARED}[schedule], interval)

    def choose_mode(self, ctx, cam, sun, want_crc=None):
        """Exchange x schedule by measurement on this box: each candidate runs 4 + 12 flushed frames, the smallest max-over-ranks mean wins.
        A candidate whose frame is not the single-GPU frame on every rank (want_crc) is out, whatever its time."""
        import zlib
        args = self.args
        if args.baseline_kernel:
            return ("allgather" if self.world > 1 else "none", "static"), {}
        # (peerpush, peertiles and the NCCL-barrier variant `peer` lose to peerflags at every N measured — profiles/README.md — and stay selectable by name)
        exchanges = (["allgather", "peerflags"] if args.exchange == "auto" else [args.exchange]) if self.world > 1 else ["none"]
        cands = []
        for ex in exchanges:
            for sc in (["static", "lpt", "deal"] if args.schedule == "auto" else [args.schedule]):  # (`shared` loses at every N: by name only)
                if sc in ("deal", "shared") and (self.world == 1 or ex == "allgather" or not ctx.interleaved):
                    continue  # these need a peer exchange (a rank's tiles are scattered over the image) and are pointless on one GPU
                cands.append((ex, sc))
        if not cands:
            raise SystemExit("no exchange / schedule combination fits these options")
        if len(cands) == 1:
            self.set_mode(ctx, *cands[0])
            return cands[0], {}
        table = {}
        for ex, sc in cands:
            self.set_mode(ctx, ex, sc)
            ms, _ = self.timed(ctx, cam, sun, 12, 4)
            mean = self.reduce([sum(ms) / len(ms)])[0]
            if want_crc is not None:
                bad = float(zlib.crc32(ctx.read_framebuffer().tobytes()) != want_crc)
                if self.reduce([bad])[0] > 0:
                    mean = None  # wrong frame on some rank
            table[f"{ex}/{sc}"] = mean
        good = {k: v for k, v in table.items() if v is not None}
        if not good:
            raise SystemExit("bench.py: no exchange / schedule combination reproduced the single-GPU frame")
        best = min(good, key=good.get)
        ex, sc = best.split("/")
        self.set_mode(ctx, ex, sc)
        return (ex, sc), table


def stats(ms):
    s = sorted(ms)
    return {"min": s[0], "median": s[len(s) // 2], "max": s[-1], "mean": sum(s) / len(s)}


def frame_counters(rig, grid, mats, W, H, brick_dim, cam, sun):
    """Rays cast and the request-byte model of the whole frame (identical to the oracle's counters, tests/test_golden.py): the
    reference-shape kernel with VRT_FLAG_AOV on this GPU."""
    from zig_vulkan_b200 import ffi

    cctx = ffi.Context(W, H, len(grid.brick_indices), brick_dim=brick_dim, n_brick_alloc=grid.brick_alloc, device=rig.local_rank, flags=ffi.VRT_FLAG_AOV | ffi.VRT_FLAG_BASELINE)
    cctx.upload_grid(grid, mats)
    cctx.trace(cam, sun)
    c = cctx.counters()
    cctx.close()
    brick_bytes = brick_dim ** 3 // 8
    c["alg_bytes"] = 4 * W * H + 4 * c["status_fetches"] + (4 + brick_bytes) * c["bricks_entered"] + 25 * c["hits"]
    return c


def measure(rig, grid, mats, W, H, brick_dim, cam, sun, steps, warmup, with_e2e=True):
    """One workload at this world size: mode selection, the timed region, per-rank kernel / exchange split, frame CRC against the
    single-GPU frame, end-to-end through host buffers."""
    import zlib

    import numpy as np
    torch, dist = rig.torch, rig.dist
    world, rank = rig.world, rig.rank
    n_pixels = W * H
    flags = 0
    if rig.args.baseline_kernel:
        from zig_vulkan_b200 import ffi
        flags = ffi.VRT_FLAG_BASELINE
    # the frame one GPU traces alone: what every rank must end up with, whatever the exchange and the schedule
    solo_crc = 0
    if rank == 0:
        solo = rig.make_ctx(grid, mats, W, H, brick_dim, flags, solo=True)
        solo_crc = zlib.crc32(solo.trace_to_host(cam, sun).tobytes())
        solo.close()
    solo_crc = int(rig.reduce([float(solo_crc)])[0])  # (a CRC32 is exact in a float64)
    ctx = rig.make_ctx(grid, mats, W, H, brick_dim, flags)
    (exchange, schedule), table = rig.choose_mode(ctx, cam, sun, want_crc=solo_crc)

    step_ms, wall_ms = rig.timed(ctx, cam, sun, steps, max(warmup, 3))
    launches_timed = rig.launches
    total_ms = rig.reduce([sum(step_ms)])[0]  # max over ranks of the summed device time
    step_max = rig.reduce(step_ms)            # per-step max over ranks

    # where the time goes: (rebuild +) trace kernel alone vs the exchange behind it, per rank (a short separate loop that reads the
    # ctx's own events after every frame; not part of `value`)
    kms, xms = [], []
    for i in range(3 + 16):
        rig.flush.fill_(i & 0xFF)
        ctx.trace(cam, sun)
        if i >= 3:
            k, t = ctx.last_trace_kernel_ms(), ctx.last_trace_ms()
            kms.append(k), xms.append(t - k)
    my_kernel = sum(kms) / len(kms)
    my_exch = sum(xms) / len(xms)
    kmax, xmax = rig.reduce([my_kernel, my_exch])
    kmin = -rig.reduce([-my_kernel])[0]

    # the frame itself: identical on every rank, identical to one GPU tracing it alone
    img = ctx.read_framebuffer()
    crc = zlib.crc32(img.tobytes())
    crcs = [crc]
    if world > 1:
        t = torch.tensor([crc], dtype=torch.int64, device=rig.dev)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        crcs = [int(o[0]) for o in out]

    res = {"ctx": ctx, "exchange": exchange, "schedule": schedule, "candidates_ms": table, "step_ms": step_ms, "step_max": step_max, "total_ms": total_ms,
           "wall_ms": wall_ms, "launches_timed": launches_timed, "kernel_ms_max": kmax, "kernel_ms_min": kmin, "exchange_ms": xmax,
           "crc": {"value": f"{crc:08x}", "all_ranks_equal": len(set(crcs)) == 1, "equals_single_gpu_frame": (solo_crc == crc) if rank == 0 else None}}
    if not with_e2e:
        return res

    # ---------------------------------------------------------------- end-to-end through the C ABI with host buffers
    # (a) frame latency: vrt_trace_to_host per frame, blocking (camera+sun from host structs -> frame in pinned host memory)
    host_frames = [torch.empty(n_pixels * 4, dtype=torch.uint8).pin_memory() for _ in range(2)]
    lat_s, n_lat = 0.0, min(steps, 50)
    for i in range(3 + n_lat):
        rig.flush.fill_(i & 0xFF)
        rig.barrier()
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
    # (b) frame throughput: the same call pipelined (vrt_trace_to_host_async, 2 frames in flight: the copy of frame k overlaps the
    # trace of frame k+1).  The L2 flush is enqueued between frames INSIDE the timed region.  Rank 0 receives the frames on its
    # host; the other ranks take part in the same frame ring without a host copy.
    for i in range(4):
        ctx.trace_to_host_async(cam, sun, host_frames[i & 1].data_ptr() if rank == 0 else None)
    ctx.sync()
    rig.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        rig.flush.fill_(i & 0xFF)
        ctx.trace_to_host_async(cam, sun, host_frames[i & 1].data_ptr() if rank == 0 else None)
    ctx.sync()
    rig.barrier()
    e2e_s = rig.reduce([time.perf_counter() - t0])[0]
    res["e2e_s"], res["latency_ms"] = e2e_s, lat_s / n_lat * 1e3
    # what the in-stream L2 flush itself costs per frame (it is inside the e2e region and serial with the trace)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(rig.stream)
    for i in range(20):
        rig.flush.fill_(i & 0xFF)
    f1.record(rig.stream)
    f1.synchronize()
    res["flush_ms"] = f0.elapsed_time(f1) / 20
    res["e2e_frame_ok"] = bool(np.array_equal(host_frames[(steps - 1) & 1].numpy().reshape(H, W, 4), img)) if rank == 0 else None
    if world > 1 and ctx.interleaved:
        res["e2e_host"] = e2e_host_assembled(rig, ctx, cam, sun, W, H, steps, img)
    return res


def e2e_host_assembled(rig, ctx, cam, sun, W, H, steps, want_img):
    """N > 1 with the host as the consumer: no device-side exchange (VRT_EXCHANGE_HOST) — every rank DMAs its own strips straight into
    ONE frame in shared pinned host memory (POSIX shm, cudaHostRegister'ed by each process), N PCIe links at once.  Two frames in
    flight per rank; the timed region ends when every rank's last copy has landed (vrt_sync + barrier)."""
    import numpy as np
    from multiprocessing import shared_memory
    from zig_vulkan_b200 import ffi

    torch, dist = rig.torch, rig.dist
    fb = W * H * 4
    name = [f"vrt_bench_{os.getpid()}" if rig.rank == 0 else None]
    dist.broadcast_object_list(name, 0)
    shm = shared_memory.SharedMemory(name=name[0], create=True, size=2 * fb) if rig.rank == 0 else None
    dist.barrier()
    if shm is None:
        shm = shared_memory.SharedMemory(name=name[0])
        try:  # this process only attaches: keep its resource tracker from unlinking (and complaining about) rank 0's segment at exit
            from multiprocessing import resource_tracker
            resource_tracker.unregister(shm._name, "shared_memory")
        except Exception:
            pass
    frames = np.frombuffer(shm.buf, dtype=np.uint8, count=2 * fb)
    ptr = frames.ctypes.data
    rt = torch.cuda.cudart()
    registered = int(rt.cudaHostRegister(ptr, 2 * fb, 0)) == 0
    out = None
    try:
        ctx.comm_set_exchange(ffi.VRT_EXCHANGE_HOST)
        ctx.set_schedule(ffi.VRT_SCHED_LPT, 8)
        if rig.rank == 0:
            frames[:] = 0
        for i in range(9):  # warm-up: the first sort of the new (local) tile space is in here
            ctx.trace_to_host_async(cam, sun, ptr + (i & 1) * fb)
        ctx.sync()
        rig.barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            rig.flush.fill_(i & 0xFF)
            ctx.trace_to_host_async(cam, sun, ptr + (i & 1) * fb)
        ctx.sync()
        rig.barrier()
        secs = rig.reduce([time.perf_counter() - t0])[0]
        ok = bool(np.array_equal(frames[((steps - 1) & 1) * fb:][:fb].reshape(H, W, 4), want_img)) if rig.rank == 0 else None
        out = {"seconds": secs, "frame_ok": ok, "pinned": registered, "d2h_bytes_per_step_per_rank": fb // rig.world}
    finally:
        rig.barrier()
        if registered:
            rt.cudaHostUnregister(ptr)
        del frames
        shm.close()
        if rig.rank == 0:
            shm.unlink()
    return out


def e2e_record(m, rays, steps, n_pixels, world):
    """The headline: frames through the C ABI with host buffers.  N = 1: vrt_trace_to_host_async into pinned memory.  N > 1: the
    host-assembled exchange (every rank's strips over its own PCIe link into one shared pinned frame); the variant that first
    assembles the frame on every GPU over NVLink and then ships it through rank 0's link alone is reported beside it."""
    funnel = {"value": rays * steps / m["e2e_s"] / 1e6, "ms_per_step": m["e2e_s"] / steps * 1e3, "frame_latency_ms": m["latency_ms"], "frame_ok": m["e2e_frame_ok"],
              "l2_flush_ms_per_step": m.get("flush_ms")}
    how = ("vrt_trace_to_host_async per frame (camera+sun host structs in, RGBA8 frame into pinned host memory, 2 frames in flight), "
           "L2 flush enqueued between frames inside the timed region (l2_flush_ms_per_step of every step is that fill); frame_latency_ms = blocking vrt_trace_to_host")
    h = m.get("e2e_host")
    if h:
        return {"value": rays * steps / h["seconds"] / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 128 * world, "d2h_bytes_per_step": n_pixels * 4,
                "ms_per_step": h["seconds"] / steps * 1e3, "frame_ok": h["frame_ok"], "pinned": h["pinned"], "l2_flush_ms_per_step": m.get("flush_ms"),
                "how": "VRT_EXCHANGE_HOST: no device-side exchange; " + how.split(";")[0] + f"; each of the {world} ranks copies its own 4-row strips into ONE "
                       "frame in shared pinned host memory over its own PCIe link; timed until every rank's last copy has landed",
                "via_rank0_after_device_exchange": funnel}
    funnel.update({"unit": "Mrays/s", "h2d_bytes_per_step": 128, "d2h_bytes_per_step": n_pixels * 4, "how": how})
    return funnel


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

    import numpy as np
    import torch

    import zig_vulkan_b200 as zv
    from zig_vulkan_b200 import ffi, scenes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the trace path has no CPU fallback (use --impl reference for the host-core baseline)")
    rig = Rig(args)
    rank, dev, stream = rig.rank, rig.dev, rig.stream
    wl = scenes.WORKLOADS[args.workload]
    W, H = wl.width, wl.height
    n_pixels = W * H
    grid = scenes.build_grid(wl.n_voxels, wl.brick_dim, brick_alloc=alloc_for(wl))
    mats = zv.terrain_materials()
    cam = scenes.camera(W, H, **POSE0)
    sun = scenes.sun(wl.sun)

    sampler = ClockSampler(rig.local_rank)
    sampler.start()
    m = measure(rig, grid, mats, W, H, wl.brick_dim, cam, sun, args.steps, args.warmup)
    clocks = sampler.stop()
    ctx = m["ctx"]
    line = None
    if rank == 0:
        cnt = frame_counters(rig, grid, mats, W, H, wl.brick_dim, cam, sun)
        rays, alg_bytes = cnt["rays"], cnt["alg_bytes"]
        peak, peak_kind = peaks()
        ms_per_step = m["total_ms"] / args.steps
        value = rays / (ms_per_step * 1e-3) / 1e6
        # roofline of the dominant kernel = the trace kernel (at N = 1 the step IS that one launch; at N > 1 each rank's launch
        # covers 1/N of the frame's algorithmic bytes and `kernel_ms` is the slowest rank's)
        kernel_ms = ms_per_step if world == 1 else m["kernel_ms_max"]
        achieved = (alg_bytes / world / (kernel_ms * 1e-3)) / 1e9
        prof = {}
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                prof = json.load(f).get(args.workload) or {}
        except Exception:
            pass
        if not isinstance(prof, dict):
            prof = {"dram_bytes": prof}
        issue = None
        if prof.get("warp_instructions") and world == 1:
            # the bound this kernel actually runs against: one warp instruction per scheduler per cycle, 4 schedulers x 148 SMs
            sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965
            floor_ms = prof["warp_instructions"] / (148 * 4 * sm_mhz * 1e6) * 1e3
            issue = {"bound": "issue", "warp_instructions": prof["warp_instructions"], "floor_ms": floor_ms, "frac": floor_ms / kernel_ms, "sm_mhz": sm_mhz,
                     "source": prof.get("source"), "note": "instruction count from the committed ncu capture of this command; clock sampled live"}
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{wl.name}: {wl.description}", "pose": POSE0_DESC, "rays_per_step": int(rays),
                "grid_bricks": len(grid.brick_indices), "active_bricks": grid.active_bricks, "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB fill)",
                "kernel": "baseline" if args.baseline_kernel else "tuned",
                "partition": ("4-row strips round-robin" if ctx.interleaved else "row slabs") + f" over {world} ranks" if world > 1 else "whole frame",
                "exchange": m["exchange"], "schedule": m["schedule"], "mode_candidates_ms": m["candidates_ms"],
            },
            "step_ms": stats(m["step_max"]),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": prof.get("dram_bytes"),
                         "traffic_source": prof.get("source"), "peak_kind": peak_kind, "algorithmic_bytes_per_launch": int(alg_bytes // world), "kernel_ms": kernel_ms,
                         "note": "request-byte model of the reference algorithm (DESIGN.md); DRAM is ~0.3 % busy — the kernel is issue-bound, see `issue`",
                         "issue": issue},
            "e2e": e2e_record(m, rays, args.steps, n_pixels, world),
            "gpu_launches": m["launches_timed"], "wall_ms": m["wall_ms"], "clocks": clocks,
            "frame_crc": m["crc"],
        }
        if world > 1:
            line["per_rank_kernel_ms"] = {"max": m["kernel_ms_max"], "min": m["kernel_ms_min"]}
            line["exchange_ms"] = m["exchange_ms"]

    extras = not args.no_extras and not args.baseline_kernel
    if extras and args.workload == "C3":
        # BASELINE config 5 beside it: the same scene at 3840x2160 (at N > 1: its own contexts, communicator and peer mappings)
        ctx.close()
        cam5 = scenes.camera(3840, 2160, **POSE0)
        m5 = measure(rig, grid, mats, 3840, 2160, wl.brick_dim, cam5, sun, max(20, args.steps // 4), 5)
        if rank == 0:
            c5 = frame_counters(rig, grid, mats, 3840, 2160, wl.brick_dim, cam5, sun)
            ms5 = m5["total_ms"] / max(20, args.steps // 4)
            line["c5"] = {"workload": scenes.WORKLOADS["C5"].description, "rays_per_step": c5["rays"], "ms_per_step": ms5, "value": c5["rays"] / (ms5 * 1e-3) / 1e6,
                          "unit": "Mrays/s", "step_ms": stats(m5["step_max"]), "exchange": m5["exchange"], "schedule": m5["schedule"], "mode_candidates_ms": m5["candidates_ms"],
                          "per_rank_kernel_ms": {"max": m5["kernel_ms_max"], "min": m5["kernel_ms_min"]}, "exchange_ms": m5["exchange_ms"], "frame_crc": m5["crc"],
                          "e2e": e2e_record(m5, c5["rays"], max(20, args.steps // 4), 3840 * 2160, rig.world)}
        m5["ctx"].close()
        ctx = None
    if rank == 0 and world == 1 and extras:
        if ctx is None:
            ctx = rig.make_ctx(grid, mats, W, H, wl.brick_dim)
            rig.set_mode(ctx, "none", m["schedule"])
        single_gpu_extras(rig, line, ctx, wl, grid, mats, cam, sun)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        crays, ctimes, cores, kind, _ = time_oracle(wl, 12, 1, budget_s=20.0)
        best = min(ctimes)
        line["cpu_baseline"] = {"value": crays / best / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind,
                                "sample": f"best of {len(ctimes)} full {W}x{H} frames of the same workload, "
                                          + ("oracle/_ref/libref_shader.so = the reference's shader text compiled by g++" if kind == "reference" else "oracle/liboracle.so")
                                          + f", {cores} threads"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if ctx is not None:
        ctx.close()
    if world > 1:
        rig.barrier()
        rig.dist.destroy_process_group()


def single_gpu_extras(rig, line, ctx, wl, grid, mats, cam, sun):
    """What sits around the headline, N = 1 only: the reference's own fly-through, the survey's pose, the reference application's
    default configuration (general shading path), scene edits, the present pass, explicit rays."""
    import numpy as np
    import torch

    import zig_vulkan_b200 as zv
    from zig_vulkan_b200 import ffi, scenes

    W, H = wl.width, wl.height
    n_pixels = W * H
    dev, stream, flush = rig.dev, rig.stream, rig.flush

    def trace_ms(c, pcam, psun, n=20, warm=3):
        ms = []
        for i in range(warm + n):
            flush.fill_(i & 0xFF)
            c.trace(pcam, psun)
            if i >= warm:
                ms.append(c.last_trace_ms())
        return ms

    # ---- the reference's own benchmark camera path (Benchmark.zig:141-173; offsets are in its world units, our world has the same
    # 64-unit extent) + the survey's pose 0.  Rays per pose from the counting kernel.
    cctx = ffi.Context(W, H, len(grid.brick_indices), brick_dim=wl.brick_dim, n_brick_alloc=grid.brick_alloc, device=rig.local_rank, flags=ffi.VRT_FLAG_AOV | ffi.VRT_FLAG_BASELINE)
    cctx.upload_grid(grid, mats)

    def pose_record(pcam, label):
        cctx.trace(pcam, sun)
        prays = cctx.counters()["rays"]
        ms = trace_ms(ctx, pcam, sun, n=24, warm=9)  # the schedule re-sorts every 8 frames: the first sort for this pose is inside the warm-up
        mean = sum(ms) / len(ms)
        return {"pose": label, "rays": prays, "ms": mean, "ms_min": min(ms), "mrays_s": prays / (mean * 1e-3) / 1e6}

    per_pose = []
    for origin, yaw in scenes.sweep_poses(11):
        per_pose.append(pose_record(scenes.camera_from_pose(W, H, origin, yaw), [round(v, 3) for v in origin]))
    line["sweep"] = {"what": "11 way points of the reference's benchmark fly-through (Benchmark.zig:141-173)", "poses": per_pose,
                     "mean_mrays_s": sum(p["mrays_s"] for p in per_pose) / len(per_pose),
                     "total_mrays_s": sum(p["rays"] for p in per_pose) / sum(p["ms"] * 1e-3 for p in per_pose) / 1e6,
                     "mean_ms": sum(p["ms"] for p in per_pose) / len(per_pose)}
    line["pose_survey"] = pose_record(scenes.camera(W, H, **POSE_SURVEY), "SURVEY 8(d) pose 0: origin (0,-8,0), yaw 0")
    cctx.close()
    ctx.trace(cam, sun)

    # ---- the reference application's default configuration (main.zig:77-81,122-135; Sun.zig:4-11): general shading path
    R = scenes.REF_DEFAULT
    rgrid = scenes.build_ref_default_grid()
    rcam = scenes.camera(R["width"], R["height"], spp=R["spp"], max_bounce=R["max_bounce"], origin=(0.0, -5.0, 14.0), euler_deg=(25.0, 0.0, 0.0))
    rsun = scenes.sun(True, R["sun_radius"])
    rc = ffi.Context(R["width"], R["height"], len(rgrid.brick_indices), device=rig.local_rank, flags=ffi.VRT_FLAG_AOV)
    rc.upload_grid(rgrid, mats)
    rc.trace(rcam, rsun)
    rrays = rc.counters()["rays"]
    rc.close()
    rctx = ffi.Context(R["width"], R["height"], len(rgrid.brick_indices), device=rig.local_rank)
    rctx.set_stream(stream.cuda_stream)
    rctx.upload_grid(rgrid, mats)
    rctx.set_schedule(ffi.VRT_SCHED_LPT, 8)
    rms = trace_ms(rctx, rcam, rsun, n=24, warm=9)
    # the same frame restricted to what the simple path covers (1 sample, no bounce, point sun), for the per-ray comparison
    scam = scenes.camera(R["width"], R["height"], spp=1, max_bounce=0, origin=(0.0, -5.0, 14.0), euler_deg=(25.0, 0.0, 0.0))
    sms = trace_ms(rctx, scam, scenes.sun(True, 0.0), n=24, warm=9)
    # ... and those same rays through the GENERAL path's code (an unused material entry of an unknown type switches the launch to it
    # without changing a pixel, tests/test_gpu_parity.py::test_simple_and_general_shading_paths_agree): what the code path itself costs
    gms = None
    unused = next((i for i in range(len(mats) - 1, 0, -1) if not (rgrid.material_indices == i).any()), None)
    if unused is not None:
        gmats = mats.copy()
        gmats[unused]["type"] = 4
        rctx.upload_materials(0, gmats)
        gms = trace_ms(rctx, scam, scenes.sun(True, 0.0), n=24, warm=9)
        rctx.upload_materials(0, mats)
    sc = ffi.Context(R["width"], R["height"], len(rgrid.brick_indices), device=rig.local_rank, flags=ffi.VRT_FLAG_AOV | ffi.VRT_FLAG_BASELINE)
    sc.upload_grid(rgrid, mats)
    sc.trace(scam, scenes.sun(True, 0.0))
    srays = sc.counters()["rays"]
    sc.close()
    rctx.close()
    rmean, smean = sum(rms) / len(rms), sum(sms) / len(sms)
    line["ref_default"] = {"workload": "reference default (main.zig:77-81,122-135): 128x64x128 bricks of 4^3 @0.5, 1024x576, spp 2, max_bounce 2, sun disc radius 5",
                           "rays_per_step": rrays, "ms_per_step": rmean, "mrays_s": rrays / (rmean * 1e-3) / 1e6, "ns_per_ray": rmean * 1e6 / rrays,
                           "simple_path_same_view": {"rays": srays, "ms": smean, "ns_per_ray": smean * 1e6 / srays},
                           "general_over_simple_per_ray": (rmean / rrays) / (smean / srays),
                           "note": "general_over_simple_per_ray compares different rays (two samples, scattered bounce rays, a sun disc against one coherent "
                                   "camera + sun ray per pixel); general_path_same_rays runs the simple configuration's rays through the general path's code"}
    if gms:
        gmean = sum(gms) / len(gms)
        line["ref_default"]["general_path_same_rays"] = {"rays": srays, "ms": gmean, "ns_per_ray": gmean * 1e6 / srays, "over_simple_path": gmean / smean}

    # ---- the step after the path: the reference's present pass (image.frag) over the traced frame, same resolution
    dn = []
    for i in range(3 + 20):
        flush.fill_(i & 0xFF)
        ctx._check(ctx._l.vrt_denoise(ctx.handle, ffi.DenoiseParams.default(), W, H, 0))
        if i >= 3:
            dn.append(ctx.last_denoise_ms())
    dn_ms = sum(dn) / len(dn)
    line["denoise"] = {"ms_per_frame": dn_ms, "mpixels_s": n_pixels / (dn_ms * 1e-3) / 1e6, "params": "samples 20, bias 0.6, multiplier 1.5, tolerance 20 (GraphicsPipeline.zig:34-39)",
                       "algorithmic_gb_s": 8 * n_pixels / (dn_ms * 1e-3) / 1e9, "bound": "ALU/FMA issue (21 samples x 2 pow per pixel), not HBM"}

    # ---- explicit-ray mode (vrt_trace_rays): this camera's primary rays as a 32 B/ray device buffer in, 32 B/ray hit records out
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

    # ---- scene edits between frames (VoxelRT.updateGridDelta, VoxelRT.zig:107-172): insert on the host grid, ship the five dirty
    # ranges through the staging ring, trace.  Timed: vrt_trace's own events (rebuild of the derived structures + trace).
    ship_delta = lambda: ctx.upload_grid_delta(grid)

    for which in range(5):
        grid.delta_reset(which)
    n = wl.n_voxels
    rng = np.random.default_rng(1)
    fresh = [(int(rng.integers(8, n - 8)), n - 8 - 4 * (k % 3), int(rng.integers(8, n - 8))) for k in range(12)]  # floating high above the terrain: empty bricks
    t_in, t_new = [], []
    for k in range(12):
        flush.fill_(k)
        x, y, z = fresh[k]
        grid.insert(x, y, z, 7)  # a new brick: status bit + brick index change -> distance planes rebuilt
        ship_delta()
        ctx.trace(cam, sun)
        t_new.append(ctx.last_trace_ms())
        flush.fill_(k)
        grid.insert(x, y, z - 1 if z % 4 else z + 1, 7)  # its neighbour voxel in the same brick: occupancy only
        ship_delta()
        ctx.trace(cam, sun)
        t_in.append(ctx.last_trace_ms())
    base = trace_ms(ctx, cam, sun, n=12, warm=2)
    line["edit_frame_ms"] = {"no_edit": sum(base) / len(base), "voxel_in_loaded_brick": sum(t_in[2:]) / len(t_in[2:]), "voxel_in_new_brick": sum(t_new[2:]) / len(t_new[2:]),
                             "what": "Grid.insert on the host, the five dirty ranges through the pinned staging ring, vrt_trace (rebuild of derived structures + trace kernel)"}


if __name__ == "__main__":
    main()

"""Device -> pinned-host copy bandwidth of an 8.3 MB frame: one copy vs split over 2 / 4 streams (is vrt_trace_to_host_async's copy link- or engine-bound?)"""
import time, torch
dev = torch.device("cuda", 0)
n = 1920 * 1080 * 4
src = torch.empty(n, dtype=torch.uint8, device=dev)
dst = torch.empty(n, dtype=torch.uint8).pin_memory()
for parts in (1, 2, 4):
    streams = [torch.cuda.Stream(dev) for _ in range(parts)]
    chunk = n // parts
    def once():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dst[i * chunk:(i + 1) * chunk].copy_(src[i * chunk:(i + 1) * chunk], non_blocking=True)
    for _ in range(5):
        once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 200
    print(f"{parts} stream(s): {dt * 1e3:.4f} ms per frame, {n / dt / 1e9:.1f} GB/s")
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev); hb = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter(); hb.copy_(big, non_blocking=True); torch.cuda.synchronize(); print("256 MiB D2H: %.1f GB/s" % ((256 << 20) / (time.perf_counter() - t0) / 1e9))

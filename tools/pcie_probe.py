"""Pinned D2H / H2D bandwidth of the box (GPU): the ceiling of bench.py's e2e arm."""
import time, torch
dev = torch.device("cuda", 0)
for mb in (64, 256, 868):
    n = mb * 1024 * 1024
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5): fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print("%s %4d MiB: %.1f GB/s" % (name, mb, n / dt / 1e9))
# two streams, both directions at once
d2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev); h2 = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("bidirectional: D2H %.1f GB/s + H2D %.1f GB/s" % (h.numel() / dt / 1e9, h2.numel() / dt / 1e9))

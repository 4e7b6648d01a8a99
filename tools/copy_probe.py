"""Host<->device copy bandwidth of the box (pinned memory), alone and both directions at once: the ceiling of
bench.py's end-to-end leg.  Usage: python tools/copy_probe.py"""
import time
import torch

n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for _ in range(2):
    d.copy_(h, non_blocking=True); h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h2.copy_(d2, non_blocking=True))):
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    print(f"{name}: {5 * n / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"both at once: {5 * n / dt / 1e9:.1f} GB/s each way")

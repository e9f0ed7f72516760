"""Measures pinned-host <-> device copy bandwidth on the box (context for bench.py's e2e number)."""
import torch, time
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        t = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    print(name, "1 GiB: %.1f GB/s" % (n / best / 1e9))
# 2-D style: 11 rows

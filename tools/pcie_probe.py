"""Dev probe: pinned-memory PCIe bandwidth (H2D, D2H, both at once) -- the floor of bench.py's e2e."""
import time

import torch

n = 1 << 28  # 1 GiB of int32
h_a = torch.empty(n, dtype=torch.int32).pin_memory()
h_b = torch.empty(n, dtype=torch.int32).pin_memory()
d_a = torch.empty(n, dtype=torch.int32, device="cuda")
d_b = torch.empty(n, dtype=torch.int32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best


gb = n * 4 / 1e9
print("H2D GB/s", gb / t(lambda: d_a.copy_(h_a, non_blocking=True)))
print("D2H GB/s", gb / t(lambda: h_b.copy_(d_b, non_blocking=True)))


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


print("duplex GB/s each way", gb / t(both))

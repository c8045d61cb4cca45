"""Dev probe: write-only and copy bandwidth of this GPU with torch kernels (roofline sanity)."""
import torch

x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")   # 4 GiB
y = torch.empty(1 << 29, dtype=torch.float32, device="cuda")
z = torch.empty(1 << 29, dtype=torch.float32, device="cuda")


def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print("fill 4 GiB   GB/s", x.numel() * 4 / t(lambda: x.fill_(1.0)) / 1e6)
print("zero 4 GiB   GB/s", x.numel() * 4 / t(lambda: x.zero_()) / 1e6)
print("copy 2+2 GiB GB/s", 2 * y.numel() * 4 / t(lambda: z.copy_(y)) / 1e6)
print("read 4 GiB (sum) GB/s", x.numel() * 4 / t(lambda: x.sum()) / 1e6)

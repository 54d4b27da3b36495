"""Write-only and copy bandwidth of this GPU (context for the emit kernel, which is 91 % stores)."""
import torch
n = 8 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(f, reps=10):
    for _ in range(3): f()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); f(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best
ms = t(lambda: a.fill_(73))
print(f"fill  (write only): {n / ms / 1e6:8.1f} GB/s")
ms = t(lambda: a.view(torch.int64).fill_(0x4949494949494949))
print(f"fill64(write only): {n / ms / 1e6:8.1f} GB/s")
ms = t(lambda: b.copy_(a))
print(f"copy  (read+write): {2 * n / ms / 1e6:8.1f} GB/s")
ms = t(lambda: torch.cuda.memset if False else a.zero_())
print(f"zero  (write only): {n / ms / 1e6:8.1f} GB/s")

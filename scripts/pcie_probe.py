"""D2H / H2D bandwidth of pinned memory on this box (context for the e2e numbers)."""
import time
import torch
d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
h = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
for name, fn in (("D2H", lambda n: h[:n].copy_(d[:n], non_blocking=True)), ("H2D", lambda n: d[:n].copy_(h[:n], non_blocking=True))):
    for n in (64 << 20, 256 << 20, 1 << 30):
        fn(n)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fn(n)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print(f"{name} {n >> 20} MiB: {n / dt / 1e9:.1f} GB/s")

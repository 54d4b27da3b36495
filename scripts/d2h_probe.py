"""Concurrent device→host bandwidth of one box: how many GB/s the HOST takes when 1, 2, 4, ... GPUs stream at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/d2h_probe.py [out.json]

No kernels of this repo run: every rank owns a pinned host ring (4 x 256 MiB, first-touched by the rank itself) and
copies a device buffer into it over and over.  In phase n only ranks < n are active (n = 1, 2, 4, ..., N); the others
idle at the barrier.  The aggregate is what the e2e path of bench.py can at most deliver to host memory on this box:
one GPU alone is bounded by its PCIe link, many GPUs together by the host's memory / root-complex bandwidth.
Also printed: each rank's NUMA node and CPU affinity (the round-1 NUMA binding never took effect: the boxes of this
pool are single-node VMs, /sys/bus/pci/devices/*/numa_node reads -1).
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device(dev))

SLOT = 256 << 20
d = torch.empty(SLOT, dtype=torch.uint8, device=dev)
d.fill_(7)
ring = [torch.empty(SLOT, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
for r in ring:
    r.fill_(0)  # first touch by the owning process


def numa_info():
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
    except Exception as e:  # noqa: BLE001
        bdf, node = "?", f"unreadable ({type(e).__name__})"
    nodes = [x for x in os.listdir("/sys/devices/system/node")] if os.path.isdir("/sys/devices/system/node") else []
    return {"rank": rank, "pci": bdf, "gpu_numa_node": node, "host_numa_nodes": sorted(n for n in nodes if n.startswith("node")),
            "cpus_allowed": len(os.sched_getaffinity(0))}


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def phase(n_active, seconds=1.5):
    active = rank < n_active
    barrier()
    nbytes = 0
    t0 = time.perf_counter()
    if active:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            k = 0
            while time.perf_counter() - t0 < seconds:
                for _ in range(4):
                    ring[k & 3].copy_(d, non_blocking=True)
                    k += 1
                    nbytes += SLOT
                s.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([nbytes / dt / 1e9 if active else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t)
    barrier()
    return float(t.item())


info = numa_info()
infos = [info]
if world > 1:
    infos = [None] * world
    dist.all_gather_object(infos, info)
res = {}
n = 1
while n <= world:
    phase(n, 0.3)  # warm-up
    res[n] = phase(n)
    n *= 2
if rank == 0:
    out = {"world": world, "slot_mib": SLOT >> 20, "aggregate_d2h_gb_per_s": res,
           "per_gpu_gb_per_s": {k: v / k for k, v in res.items()}, "ranks": infos}
    print(json.dumps(out))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
if world > 1:
    dist.destroy_process_group()

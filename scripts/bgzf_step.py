"""Times the BGZF stage on a scaled workload: device-resident chunk loop (emit + code, no host copy) and the
streamed variant through a pinned host ring.
    python scripts/bgzf_step.py [scale] [workload]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import mimick_b200 as M  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
wl = sys.argv[2] if len(sys.argv) > 2 else "c5"
desc, _, nhap, fmt, kw = bench.WORKLOADS[wl]
simkw = dict(bench.COMMON_SIM)
simkw.update(kw)
genome = bench.make_genome(wl, scale, "cuda:0", pin=True)
sim = M.Simulator(fmt, seed=42, device=0, max_chunk_bytes=int(os.environ.get("CHUNK_MIB", "1024")) << 20, **simkw)
for nm, hap, h in genome:
    sim.add_contig(nm, (h.data_ptr(), h.numel()), hap)
plan = sim.plan()
place = sim.place()
raw = place.r1_bytes + place.r2_bytes
for it in range(2):
    sim.rewind()
    ms = 0.0
    z = 0
    while True:
        ch = sim.emit_chunk_bgzf()
        ms += ch.kernel_ms
        z += ch.r1_bytes + ch.r2_bytes
        if ch.done:
            break
    print(f"resident {it}: pairs={plan.n_pairs} text={raw} coded={z} ratio={raw / z:.2f} emit+code_ms={ms:.2f} "
          f"({raw / ms / 1e6:.0f} GB/s of text)")
ms_plain = sim.emit_all()
print(f"emit only: {ms_plain:.2f} ms")
cap = sim.bgzf_bound()
ring = [torch.empty(cap, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
for it in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = [0]

    def consume(ch, p1, p2):
        got[0] += ch.r1_bytes + ch.r2_bytes
    sim.emit_stream([ring[0].data_ptr(), ring[1].data_ptr()], [ring[2].data_ptr(), ring[3].data_ptr()], cap, consume, bgzf=True)
    dt = time.perf_counter() - t0
    print(f"streamed {it}: coded={got[0]} wall={dt * 1e3:.1f} ms -> {plan.n_pairs / dt / 1e6:.1f} M pairs/s, "
          f"{got[0] / dt / 1e9:.1f} GB/s over PCIe, {raw / dt / 1e9:.1f} GB/s of text")
sim.close()

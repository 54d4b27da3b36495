"""What one rank of an N-GPU run does, timed on ONE GPU: replicated planning + placement and emission of shard r of W.
    python scripts/rank_step.py W [r] [scale] [steps]
Prints per-step wall time and the phase times (the NCCL exchange of four counters per step is not included).
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import mimick_b200 as M  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
r = int(sys.argv[2]) if len(sys.argv) > 2 else 0
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
desc, _, nhap, fmt, kw = bench.WORKLOADS["c5"]
simkw = dict(bench.COMMON_SIM)
simkw.update(kw)
genome = bench.make_genome("c5", scale, "cuda:0", pin=True)
sim = M.Simulator(fmt, seed=42, device=0, max_chunk_bytes=2048 << 20, **simkw)
for nm, hap, h in genome:
    sim.add_contig(nm, (h.data_ptr(), h.numel()), hap)
acc = [0.0, 0.0, 0.0, 0.0]
for it in range(3 + steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan = sim.plan()
    sim.set_shard(r, W)
    place = sim.place()
    ms = sim.emit_all()
    torch.cuda.synchronize()
    dt = 1e3 * (time.perf_counter() - t0)
    if it >= 3:
        for k, v in enumerate((dt, plan.host_ms, place.host_ms, ms)):
            acc[k] += v / steps
print(f"RANK world={W} rank={r} scale={scale} pairs={place.n_pairs} step_ms={acc[0]:.3f} plan_ms={acc[1]:.3f} place_ms={acc[2]:.3f} "
      f"emit_ms={acc[3]:.3f} batches={plan.n_batches} place_phases={[round(x, 2) for x in place.phase_ms]}")
sim.close()

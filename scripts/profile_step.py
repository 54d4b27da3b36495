"""One warm-up step + one step of the hot path on a scaled C5 workload — the command ncu wraps.
    python scripts/profile_step.py [scale]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mimick_b200 as M  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
wl = sys.argv[2] if len(sys.argv) > 2 else "c5"
desc, _, nhap, fmt, kw = bench.WORKLOADS[wl]
simkw = dict(bench.COMMON_SIM)
simkw.update(kw)
genome = bench.make_genome(wl, scale, "cuda:0", pin=True)
if os.environ.get("MMK_FORCE_4BIT"):  # one N makes the whole reference 4-bit
    genome[0][2][12345] = ord("N")
sim = M.Simulator(fmt, seed=42, device=0, max_chunk_bytes=2048 << 20, **simkw)
for nm, hap, h in genome:
    sim.add_contig(nm, (h.data_ptr(), h.numel()), hap)
for it in range(2):
    plan = sim.plan()
    place = sim.place()
    ms = sim.emit_all()
    print(f"step {it}: pairs={plan.n_pairs} molecules={plan.n_molecules} epochs={plan.n_epochs} batches={plan.n_batches} "
          f"bytes={place.r1_bytes + place.r2_bytes} emit_ms={ms:.3f} plan_ms={plan.host_ms:.2f} place_ms={place.host_ms:.2f} "
          f"phases={[round(x, 2) for x in place.phase_ms]}")
if os.environ.get("MMK_PROFILE_BGZF", "1") != "0":  # the same shard once more, coded as BGZF members on the device
    sim.rewind()
    ms = z = 0
    while True:
        ch = sim.emit_chunk_bgzf()
        ms += ch.kernel_ms
        z += ch.r1_bytes + ch.r2_bytes
        if ch.done:
            break
    print(f"bgzf: coded={z} ratio={(place.r1_bytes + place.r2_bytes) / z:.2f} emit+code_ms={ms:.2f}")
sim.close()

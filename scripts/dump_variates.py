"""Writes what baseline/make_julia_goldens.jl replays through the real reference on a Julia host:
    python scripts/dump_variates.py <name> <format> <fasta> [<fasta> ...] [key=value ...]
-> tests/golden/julia/<name>.variates   every variate the CUDA path consumed, in the reference's draw order (mmk_variate records)
   tests/golden/julia/<name>.args.jl    the same call as a Julia named tuple
   tests/golden/julia/<name>.cuda.sha256  SHA-256 of the CUDA path's R1 / R2 text (what the replay must reproduce)
Needs a GPU (the variates come from the kernels)."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import mimick_b200 as M  # noqa: E402

name, fmt = sys.argv[1], sys.argv[2]
fastas = [a for a in sys.argv[3:] if "=" not in a]
kw = {}
for a in sys.argv[3:]:
    if "=" in a:
        k, v = a.split("=", 1)
        kw[k] = eval(v, {})  # noqa: S307 (own command line)
kw.setdefault("seed", 0)
out = os.path.join(ROOT, "tests", "golden", "julia")
os.makedirs(out, exist_ok=True)
with M.Simulator(fmt, **kw) as sim:
    for h, p in enumerate(fastas, start=1):
        sim.load_fasta(p, h)
    sim.plan()
    sim.place()
    r1, r2 = sim.emit_to_host()
    var = np.concatenate([sim.mask_variates(), sim.variates()])
var.tofile(os.path.join(out, name + ".variates"))
jl = {True: "true", False: "false"}
items = ", ".join(f"{k}={jl.get(v, v) if isinstance(v, bool) else (list(v) if isinstance(v, tuple) else v)}" for k, v in kw.items() if k != "seed")
with open(os.path.join(out, name + ".args.jl"), "w") as fh:
    fh.write(f"(fastas={[os.path.abspath(p) for p in fastas]!r}, format={fmt!r}, kw=(; {items}))\n".replace("'", '"'))
with open(os.path.join(out, name + ".cuda.sha256"), "w") as fh:
    fh.write(f"{hashlib.sha256(r1.tobytes()).hexdigest()}  {name}.R1.fq\n{hashlib.sha256(r2.tobytes()).hexdigest()}  {name}.R2.fq\n")
print(f"{name}: {len(var)} variates, {r1.nbytes + r2.nbytes} bytes of text")

"""Wall clock of the product call a user makes — mimick([fasta1, fasta2], "haplotagging", ...) — FASTA files in, *.fq.gz out.
    python scripts/file_e2e.py [scale] [dir]
The synthetic C5 genome (bench.py) at `scale` is written as two plain FASTA files under `dir` (default /dev/shm, a tmpfs:
no disk in the measurement), mimick() reads them, simulates and writes <dir>/out/SIM.R{1,2}.fq.gz (BGZF coded on the device).
Prints one JSON line: read pairs/s over the whole call, the split load / plan+place / write, bytes written; the first
64 MiB of each output are inflated with zlib as a validity check.
"""
import json
import os
import shutil
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
import mimick_b200 as M  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
base = sys.argv[2] if len(sys.argv) > 2 else "/dev/shm"
work = os.path.join(base, "mmk_file_e2e")
shutil.rmtree(work, ignore_errors=True)
os.makedirs(os.path.join(work, "out"))
genome = bench.make_genome("c5", scale, "cuda:0", pin=False)
paths = [os.path.join(work, "hap1.fa"), os.path.join(work, "hap2.fa")]
for hap, path in enumerate(paths, start=1):
    with open(path, "wb") as f:
        for nm, h, t in genome:
            if h != hap:
                continue
            a = t.numpy()
            f.write(b">" + nm.encode() + b"\n")
            n60 = a.size // 60 * 60
            rows = np.empty((n60 // 60, 61), dtype=np.uint8)
            rows[:, :60] = a[:n60].reshape(-1, 60)
            rows[:, 60] = 10
            f.write(rows.tobytes())
            if a.size > n60:
                f.write(a[n60:].tobytes() + b"\n")
fasta_bytes = sum(os.path.getsize(p) for p in paths)
del genome
desc, _, nhap, fmt, kw = bench.WORKLOADS["c5"]
simkw = dict(bench.COMMON_SIM)
simkw.update(kw)
simkw["read_length"] = list(simkw["read_length"])
for it in range(2):  # the first call pays the CUDA context and the pinned ring; the second is the steady state
    t0 = time.perf_counter()
    outs = M.mimick(paths, fmt, prefix=os.path.join(work, "out", "SIM"), seed=42, device=0, **simkw)
    dt = time.perf_counter() - t0
o = outs[0]
written = os.path.getsize(o.r1_path) + os.path.getsize(o.r2_path)
ok = True
for p in (o.r1_path, o.r2_path):  # concatenated gzip members: inflate the first 64 MiB of the file
    raw = open(p, "rb").read(64 << 20)
    d = zlib.decompressobj(31)
    n = 0
    while raw:
        n += len(d.decompress(raw))
        if d.eof:
            raw = d.unused_data
            d = zlib.decompressobj(31)
        else:
            break
    ok = ok and n > 0
print(json.dumps({"call": "mimick([hap1.fa, hap2.fa], 'haplotagging', prefix=...) -> SIM.R1.fq.gz, SIM.R2.fq.gz on tmpfs",
                  "scale": scale, "fasta_bytes": fasta_bytes, "read_pairs": o.n_pairs, "fastq_bytes": o.r1_bytes + o.r2_bytes,
                  "file_bytes": written, "wall_s": dt, "read_pairs_per_s": o.n_pairs / dt,
                  "fastq_gb_per_s": (o.r1_bytes + o.r2_bytes) / dt / 1e9, "file_gb_per_s": written / dt / 1e9, "inflates": ok}))
shutil.rmtree(work, ignore_errors=True)

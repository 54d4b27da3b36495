"""Haplotype construction (reported separately from the emit roofline): one sample of BASELINE configs[3] —
100 Mbp master contig, 100 k SNPs + 10 k indels, diploid — through mmk_set_sample_variants (host piece list, SNP scatter,
segmented gather + 2-bit pack on the device).
    python scripts/haplotype_bench.py [reps]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mimick_b200 as M  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rng = np.random.default_rng(1004)
n = 100_000_000
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
master = acgt[rng.integers(0, 4, size=n)]
n_var = 110_000
pos = (np.sort(rng.choice(n // 64 - 2, size=n_var, replace=False)) + 1).astype(np.int64) * 64 + rng.integers(0, 16, size=n_var)
is_indel = np.zeros(n_var, dtype=bool)
is_indel[rng.choice(n_var, size=10_000, replace=False)] = True
e_pos, e_type, e_ref, e_alt, e_hap = [], [], [], [], []
for i in range(n_var):
    p = int(pos[i])
    ref = master[p - 1]
    if is_indel[i]:
        k = int(rng.integers(1, 21))
        ref_len, alt, typ = (k + 1, bytes([ref]), 1) if rng.random() < 0.5 else (1, bytes([ref]) + acgt[rng.integers(0, 4, size=k)].tobytes(), 1)
    else:
        ref_len, alt, typ = 1, bytes([acgt[(np.searchsorted(acgt, ref) + int(rng.integers(1, 4))) % 4]]), 0
    for hap in (1, 2):
        if rng.random() < 0.3:
            e_pos.append(p), e_type.append(typ), e_ref.append(ref_len), e_alt.append(alt), e_hap.append(hap)
with M.Simulator("tellseq", seed=42, coverage=20, singletons=0.0, attempts=300) as sim:
    t0 = time.perf_counter()
    sim.add_contig("chr1", master.tobytes(), 1)
    t_load = time.perf_counter() - t0
    contig = np.zeros(len(e_pos), dtype=np.int32)
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        sim.set_sample_variants(3, 2, [0], contig, e_pos, e_type, e_ref, e_alt, e_hap)   # synchronous
        best = min(best, time.perf_counter() - t0)
    lens = [s["length"] for s in sim.schemas()]
    traffic = 2 * n + sum(lens) / 4 + len(e_pos) * 24      # master ASCII read per haplotype + packed words written + records
    print(f"master upload (pageable host memory): {t_load * 1e3:.1f} ms")
    print(f"set_sample_variants: {len(e_pos)} mutations on 2 haplotypes of {n} bp -> {best * 1e3:.2f} ms "
          f"({traffic / best / 1e9:.1f} GB/s of master read + packed haplotypes written; includes the host-side piece list)")
